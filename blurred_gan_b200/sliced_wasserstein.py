"""Drop-in for the reference's ``sliced_wasserstein.py`` with every array operation on the GPU.

Same names and positional signatures as the reference; arrays may be numpy (as in the reference) or
torch tensors, results of the free functions are CUDA ``torch.Tensor`` and ``API.end`` returns a
list of Python floats like the reference.  What stays on the host is only what parity pins there:
the random draws.  Patch positions and projection directions come from numpy's *global* legacy
RandomState in exactly the reference's order -- ``randint`` x then y per level per
``feed``, 4x ``randn(147, 128)`` per level per ``end`` (also in ``end('reals')``, where the
reference draws them and gets 0) -- so that identical seeds give identical patches and projections.

Deliberate differences: inputs are never mutated (the reference overwrites float32 inputs);
the real-vs-real distance of ``end('reals')`` is returned as the exact 0 the reference computes,
without sorting anything.
"""
from __future__ import annotations

import numpy as np
import torch

from . import dist as _dist
from . import ops

# ----------------------------------------------------------------------------


def _to_device(a, device=None) -> torch.Tensor:
    if torch.is_tensor(a):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return t.to(device=device, dtype=torch.float32).contiguous()


def get_descriptors_for_minibatch(minibatch, nhood_size, nhoods_per_image, sums=None):
    """``sliced_wasserstein.py:13-23``.  Returns [n * nhoods_per_image, 3, nhood, nhood] on the GPU;
    if ``sums`` (fp64 [6] CUDA tensor) is given, per-channel sum / sum of squares are accumulated into it."""
    mb = _to_device(minibatch)
    S = mb.shape  # (minibatch, channel, height, width)
    assert len(S) == 4 and S[1] == 3
    N = nhoods_per_image * S[0]
    H = nhood_size // 2
    # same two draws, same order and shapes as the reference: columns first, then rows
    x = np.random.randint(H, S[3] - H, size=(N, 1, 1, 1))
    y = np.random.randint(H, S[2] - H, size=(N, 1, 1, 1))
    if sums is None:
        sums = torch.zeros(2 * S[1], dtype=torch.float64, device=mb.device)
    return ops.swd_extract(mb, x.reshape(-1), y.reshape(-1), nhoods_per_image, nhood_size, sums)


# ----------------------------------------------------------------------------


def finalize_descriptors(desc, sums=None):
    """``sliced_wasserstein.py:27-34``: concatenate, standardise each channel over (descriptor, i, j)
    with the population std, flatten to [rows, C * nhood * nhood]."""
    if isinstance(desc, list):
        desc = torch.cat([_to_device(d) for d in desc], dim=0)
    else:
        desc = _to_device(desc).clone()
    assert desc.dim() == 4  # (neighborhood, channel, height, width)
    rows, c, a, b = desc.shape
    if sums is None:
        d64 = desc.to(torch.float64)
        sums = torch.cat([d64.sum(dim=(0, 2, 3)), d64.square().sum(dim=(0, 2, 3))])
    ops.swd_normalise_(desc, sums, float(rows) * a * b)
    return desc.reshape(rows, -1)


# ----------------------------------------------------------------------------


def _draw_directions(dim, dirs_per_repeat):
    dirs = np.random.randn(dim, dirs_per_repeat)  # (descriptor_component, direction)
    dirs /= np.sqrt(np.sum(np.square(dirs), axis=0, keepdims=True))
    return dirs.astype(np.float32)


def _scratch_proj(scratch, name, d, rows, device):
    if scratch is None:
        return None
    t = scratch.get(name)
    if t is None or t.shape != (d, rows) or t.device != device:
        t = scratch[name] = torch.empty((d, rows), dtype=torch.float32, device=device)
    return t


def sliced_wasserstein(A, B, dir_repeats, dirs_per_repeat, group=None, scratch=None):
    """``sliced_wasserstein.py:38-51``: mean over repeats of mean |sort(A dirs) - sort(B dirs)|.

    With a process group, A and B are the full (already all-gathered) descriptor sets on every rank and each rank
    handles its own block of the ``dir_repeats * dirs_per_repeat`` directions; the per-repeat sums are all-reduced.
    ``scratch`` (a dict, optional): the two projection arrays and the sort workspace are kept in it between calls --
    ``API.end`` evaluates every pyramid level with equally shaped arrays."""
    A = _to_device(A)
    B = _to_device(B, A.device)
    assert A.dim() == 2 and A.shape == B.shape  # (neighborhood, descriptor_component)
    rows, dim = A.shape
    # every rank draws the same directions, in the reference's order: one randn(dim, dirs_per_repeat) per repeat
    dirs = np.concatenate([_draw_directions(dim, dirs_per_repeat) for _ in range(dir_repeats)], axis=1)
    rank, world = _dist.world_info(group)
    block = _dist.direction_block(dir_repeats * dirs_per_repeat, rank, world)
    partial = torch.zeros(dir_repeats, dtype=torch.float64, device=A.device)
    if block.stop > block.start:
        dirs_t = torch.from_numpy(np.ascontiguousarray(dirs[:, block])).to(A.device)
        nd = block.stop - block.start
        # [directions of this rank, rows]; sort both, sum |A - B| per repeat: one library call (the sorted B is never stored)
        projA = ops.swd_project(A, dirs_t, out=_scratch_proj(scratch, "projA", nd, rows, A.device))
        projB = ops.swd_project(B, dirs_t, out=_scratch_proj(scratch, "projB", nd, rows, A.device))
        ops.swd_sort_diff_(projA, projB, partial, block.start, dirs_per_repeat, scratch)
    return np.float32(_dist.combine_repeat_sums(partial, rows, dirs_per_repeat, group))


# ----------------------------------------------------------------------------


def downscale_minibatch(minibatch, lod):
    """``sliced_wasserstein.py:55-61`` (not used by ``API``; plain torch, kept for surface parity)."""
    if lod == 0:
        return minibatch
    t = _to_device(minibatch)
    for _ in range(lod):
        t = (t[:, :, 0::2, 0::2] + t[:, :, 0::2, 1::2] + t[:, :, 1::2, 0::2] + t[:, :, 1::2, 1::2]) * 0.25
    return torch.round(t).clip(0, 255).to(torch.uint8)


# ----------------------------------------------------------------------------

_binomial5 = np.float32([1, 4, 6, 4, 1])
gaussian_filter = np.outer(_binomial5, _binomial5).astype(np.float32) / 256.0  # same 5x5 table as the reference's :65-70


def pyr_down(minibatch):  # matches cv2.pyrDown()
    mb = _to_device(minibatch)
    assert mb.dim() == 4
    return ops.pyr_down(mb)


def pyr_up(minibatch):  # matches cv2.pyrUp()
    mb = _to_device(minibatch)
    assert mb.dim() == 4
    return ops.pyr_up(mb)


def generate_laplacian_pyramid(minibatch, num_levels):
    """``sliced_wasserstein.py:83-88``; the residual ``g - pyr_up(pyr_down(g))`` is one fused kernel."""
    pyramid = [_to_device(minibatch)]
    for _ in range(1, num_levels):
        coarse = ops.pyr_down(pyramid[-1])
        pyramid[-1] = ops.pyr_up(coarse, minuend=pyramid[-1])
        pyramid.append(coarse)
    return pyramid


def reconstruct_laplacian_pyramid(pyramid):
    minibatch = _to_device(pyramid[-1])
    for level in pyramid[-2::-1]:
        minibatch = ops.pyr_up(minibatch) + _to_device(level, minibatch.device)
    return minibatch


# ----------------------------------------------------------------------------


class _PinnedRing:
    """A few pinned int32 staging buffers for the per-feed position upload (asynchronous H2D, reused round-robin)."""

    def __init__(self, slots: int = 4):
        self.bufs = [None] * slots
        self.events = [None] * slots
        self.k = 0

    def take(self, n: int):
        k = self.k = (self.k + 1) % len(self.bufs)
        if self.events[k] is not None:
            self.events[k].synchronize()  # the copy that last read this buffer
        if self.bufs[k] is None or self.bufs[k].numel() < n:
            self.bufs[k] = torch.empty(n, dtype=torch.int32).pin_memory()
        return k, self.bufs[k][:n]

    def upload(self, k: int, host: torch.Tensor, device) -> torch.Tensor:
        dev = host.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self.events[k] = ev
        return dev


class API:
    """``sliced_wasserstein.py:98-133``: begin / feed / end state machine over a Laplacian pyramid.

    ``group``: optional ``torch.distributed`` process group (or rely on the default group when it is initialised and
    ``distributed=True``).  Every rank feeds its own shard of the images; see ``blurred_gan_b200/dist.py`` for the
    exchange and for how the random stream stays identical to a single-process run."""

    def __init__(self, image_shape, device=None, group=None, distributed=False, overlap_gather=True):
        # overlap_gather: the all-gathers of `end` run asynchronously (NCCL's stream) beside the next set's feed kernels and the
        # previous level's sort; False waits for each one on the spot (bench.py's per-phase device times)
        self.overlap_gather = overlap_gather
        self._desc_real = None
        self.nhood_size, self.nhoods_per_image = 7, 128
        self.dir_repeats, self.dirs_per_repeat = 4, 128
        self.device = device
        self.group = group
        self.rank, self.world = _dist.world_info(group) if (distributed or group is not None) else (0, 1)
        self._stage = None
        self._scratch = {}  # pyramid workspace of feed(), reused between minibatches
        # one pyramid level per power of two down to 16 px (image_shape = (C, H, W); H drives it, like the reference)
        self.resolutions = []
        size = image_shape[1]
        while size >= 16:
            self.resolutions.append(size)
            size //= 2

    def get_metric_names(self):
        return [f"SWDx1e3_{res}" for res in self.resolutions] + ["SWDx1e3_avg"]

    def get_metric_formatting(self):
        return ["%-13.4f"] * (len(self.resolutions) + 1)

    def begin(self, mode):
        assert mode in ['warmup', 'reals', 'fakes']
        self.descriptors = [[] for _ in self.resolutions]  # per level: list of chunks (level-by-level path)
        self._slabs = None  # per level: one growing [capacity, C, nhood, nhood] array, filled in place by the fused feed
        self._sums = None  # per level: fp64 [sum_c..., sumsq_c..., rows]; the statistics are accumulated on the device
        self._rows = [0 for _ in self.resolutions]
        self._feed_log = [0, 0]  # (number of feeds, rolling hash of their minibatch sizes): must agree across ranks

    def feed(self, mode, minibatch):
        mb = _to_device(minibatch, self.device)
        n, c = mb.shape[0], mb.shape[1]
        assert mb.dim() == 4 and c == 3
        self._feed_log = [self._feed_log[0] + 1, (self._feed_log[1] * 1000003 + n) % (1 << 50)]
        n_levels = len(self.resolutions)
        h, w = mb.shape[2], mb.shape[3]
        sizes = [(h >> lod, w >> lod) for lod in range(n_levels)]
        fused = h % (1 << (n_levels - 1)) == 0 and w % (1 << (n_levels - 1)) == 0  # (odd sizes: level-by-level calls below)
        if self._sums is None:
            self._sums = [torch.zeros(2 * c + 1, dtype=torch.float64, device=mb.device) for _ in range(n_levels)]
        levels = None if fused else generate_laplacian_pyramid(mb, n_levels)
        if levels is not None:
            sizes = [(lv.shape[2], lv.shape[3]) for lv in levels]
        # All draws of this feed, in the reference's order (finest level first; x then y), written straight into a pinned
        # staging buffer and uploaded asynchronously.
        count = n * self.nhoods_per_image
        if self._stage is None:
            self._stage = _PinnedRing()
        slot, host = self._stage.take(2 * n_levels * count)
        _dist.positions_packed_for_round(n, sizes, self.nhood_size, self.nhoods_per_image, self.rank, self.world, out=host.numpy())
        packed = self._stage.upload(slot, host, mb.device)
        if fused:
            # pyramid + gather of all levels: one library call (11 launches back to back), the pyramid lives in a reused workspace
            # The descriptors of a level go straight into one array that grows geometrically: no per-minibatch allocation, no
            # concatenation at the end, and no 4.8 MB blocks that would fragment the multi-GB projection / sort scratch of
            # the previous evaluation inside the caching allocator (which made `end` pay cudaMalloc / cudaFree again).
            used = self._rows[0]
            if self._slabs is None:
                assert all(not chunks for chunks in self.descriptors)
                self._slabs = [torch.empty((4 * count, c, self.nhood_size, self.nhood_size), dtype=torch.float32, device=mb.device)
                               for _ in range(n_levels)]
            if used + count > self._slabs[0].shape[0]:
                cap = max(2 * self._slabs[0].shape[0], used + count)
                for lod, old in enumerate(self._slabs):
                    new = torch.empty((cap,) + tuple(old.shape[1:]), dtype=old.dtype, device=old.device)
                    new[:used].copy_(old[:used])
                    self._slabs[lod] = new
            ops.swd_feed(mb, packed, n_levels, self.nhoods_per_image, self.nhood_size, self._sums, self._scratch,
                         outs=[slab[used : used + count] for slab in self._slabs])
            for lod in range(n_levels):
                self._rows[lod] += count
            return
        for lod, level in enumerate(levels):
            px = packed[2 * lod * count : (2 * lod + 1) * count]
            py = packed[(2 * lod + 1) * count : (2 * lod + 2) * count]
            self.descriptors[lod].append(ops.swd_extract(level, px, py, self.nhoods_per_image, self.nhood_size, self._sums[lod][: 2 * c]))
            self._rows[lod] += count

    def _finalize_all(self, local):
        """Standardise every level with the statistics of the whole (all-rank) set, then gather every rank's rows.  One
        all-reduce for the sums of all levels and one read-back of the row counts, BEFORE the first gather is enqueued: a
        read-back behind a gather would wait for it and the gathers could not run beside anything."""
        for sums, rows_local in zip(self._sums, self._rows):
            sums[-1] = float(rows_local)
        stacked = torch.stack(self._sums)  # [levels, 2 C + 1]
        _dist.all_reduce_sum_(stacked, self.group)
        rows_total = stacked[:, -1].cpu().tolist()  # exact in fp64
        out = []
        for lod, desc in enumerate(local):
            ops.swd_normalise_(desc, stacked[lod], rows_total[lod] * desc.shape[2] * desc.shape[3])
            flat = desc.reshape(desc.shape[0], -1)
            if self.overlap_gather:
                # (equal row counts on every rank: check_lockstep in end() has just verified the feed counts and sizes)
                out.append(_dist.all_gather_rows_async(flat, self.group, equal_rows=True))
            else:
                out.append(_dist.Gathered(_dist.all_gather_rows(flat, self.group)))
        return out

    @property
    def desc_real(self):
        """The standardised descriptors of the real set, one [rows, 147] tensor per level (waits for their all-gathers)."""
        return None if self._desc_real is None else [g.result() for g in self._desc_real]

    def end(self, mode):
        # every rank replays the position draws of all ranks with its OWN minibatch sizes (dist.positions_packed_for_round):
        # ranks that fed different sizes or counts would silently desynchronise the random streams (and the directions)
        _dist.check_lockstep(self._feed_log, self.group if self.world > 1 else None, self.world)
        if self._slabs is not None:
            # (a slab with more than 25 % slack is copied once so that the set kept for later comparisons does not pin the slack)
            local = [slab[:rows] if 4 * rows >= 3 * slab.shape[0] else slab[:rows].clone() for slab, rows in zip(self._slabs, self._rows)]
        else:
            local = [chunks[0] if len(chunks) == 1 else torch.cat(chunks, dim=0) for chunks in self.descriptors]
        local_dlen = int(local[0].shape[1] * local[0].shape[2] * local[0].shape[3])
        desc = self._finalize_all(local)
        del self.descriptors, local
        self._slabs = None
        if mode in ['warmup', 'reals']:
            self._desc_real = desc  # (not waited for: the gathers overlap the feed of the next set)
            # The reference evaluates sliced_wasserstein(real, real) here: dir_repeats randn draws per level and
            # a result of exactly 0.  Consume the draws (they position the stream for the fakes), skip the sorts.
            dlen = local_dlen
            for _ in desc:
                for _ in range(self.dir_repeats):
                    np.random.randn(dlen, self.dirs_per_repeat)
            dist = [np.float32(0.0)] * len(desc)
        else:
            scratch = {}  # projections + sort workspace, shared by the levels of this evaluation
            # level by level: wait for this level's two gathers only -- the later levels' gathers run beside its sort
            dist = [sliced_wasserstein(real.result(), fake.result(), self.dir_repeats, self.dirs_per_repeat, self.group, scratch)
                    for real, fake in zip(self._desc_real, desc)]
            del scratch
        scaled = [d * 1e3 for d in dist]
        return scaled + [np.mean(scaled)]
