"""Multi-GPU plumbing for the hot path: one process per GPU, `torch.distributed` (NCCL on GPUs, gloo in CPU tests).

Blur: every image plane is independent, so a batch is simply sharded across ranks -- no collective on the data path
(`shard_range`).  SWD has exactly one exchange step:

  1. all-reduce (sum, fp64) of the per-channel sum / sum-of-squares of every level, so that every rank standardises
     its descriptors with the statistics of the whole set (reference sliced_wasserstein.py:29-32);
  2. all-gather of the standardised descriptors of every level;
  3. every rank projects / sorts / diffs its own block of the dir_repeats x dirs_per_repeat directions, then the
     per-repeat partial sums are all-reduced.

The random stream stays bit-compatible with a single-process run: all ranks seed numpy identically and, at their
k-th `feed`, draw the patch positions of all `world` minibatches of round k in rank order, keeping their own
(`positions_for_round`).  The equivalent single-process run feeds the minibatches in the order
rank0[k], rank1[k], ..., rank0[k+1], ...

The helpers in this file are free of CUDA calls so that the host logic can be tested on CPU with gloo.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world_info(group=None) -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n items for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def positions_for_round(n_images: int, level_sizes, nhood_size: int, nhoods_per_image: int, rank: int, world: int):
    """Patch centres of this rank's minibatch for every pyramid level: [(x, y), ...] in level order.

    Draws, from numpy's global stream, exactly what a single process feeding the `world` minibatches of this round one
    after the other would draw -- for each minibatch, for each level (finest first), x then y
    (reference sliced_wasserstein.py:20-21,121-123) -- and keeps the draws of minibatch `rank`."""
    half = nhood_size // 2
    count = nhoods_per_image * n_images
    mine = None
    for r in range(world):
        draws = []
        for height, width in level_sizes:
            x = np.random.randint(half, width - half, size=(count, 1, 1, 1))
            y = np.random.randint(half, height - half, size=(count, 1, 1, 1))
            draws.append((x.reshape(-1), y.reshape(-1)))
        if r == rank:
            mine = draws
    return mine


def _numpy_mt19937_state():
    """(address of numpy's global legacy MT19937 state, its lock), or None when the global generator is something else.

    numpy exposes the raw state of a bit generator through its ctypes interface (`state_address`); for MT19937 that is
    `struct { uint32_t key[624]; int pos; }` (numpy/random/src/mt19937/mt19937.h)."""
    try:
        bitgen = np.random.mtrand._rand._bit_generator
        if type(bitgen).__name__ != "MT19937":
            return None
        return int(bitgen.ctypes.state_address), bitgen.lock
    except Exception:  # an unexpected numpy layout: use numpy's own randint
        return None


def positions_packed_for_round(n_images: int, level_sizes, nhood_size: int, nhoods_per_image: int, rank: int, world: int,
                               out: np.ndarray | None = None) -> np.ndarray:
    """`positions_for_round` as one int32 array [level][x | y][n_images * nhoods_per_image] (the layout `API.feed` uploads).

    Same draws from the same stream, but made by `bgan_swd_positions_i32` (csrc/hostrng.cu: a bit-exact replay of
    numpy's legacy `randint` working in place on numpy's global MT19937 state) -- numpy needs ~10 ns per draw plus ~50 us
    per call, and every rank of a `world`-rank evaluation replays the draws of all ranks.  Falls back to numpy's own
    `randint` if the global generator is not the legacy MT19937."""
    import ctypes

    from . import _lib

    count = nhoods_per_image * n_images
    levels = len(level_sizes)
    if out is None:
        out = np.empty(2 * levels * count, dtype=np.int32)
    assert out.dtype == np.int32 and out.size == 2 * levels * count and out.flags.c_contiguous
    state = _numpy_mt19937_state()
    if state is None:
        pos = positions_for_round(n_images, level_sizes, nhood_size, nhoods_per_image, rank, world)
        out[:] = np.concatenate([np.concatenate(p) for p in pos])
        return out
    addr, lock = state
    heights = (ctypes.c_int * levels)(*[int(h) for h, _ in level_sizes])
    widths = (ctypes.c_int * levels)(*[int(w) for _, w in level_sizes])
    with lock:
        rc = _lib.load().bgan_swd_positions_i32(addr, ctypes.cast(addr + 624 * 4, ctypes.POINTER(ctypes.c_int)), levels, heights, widths,
                                                nhood_size, count, world, rank, out.ctypes.data)
    _lib.check(rc, "bgan_swd_positions_i32")
    return out


def check_lockstep(values, group=None, world=None) -> None:
    """Raise unless every rank holds the same `values` (a short list of integers < 2^52, e.g. the feed count and the
    hash of the fed minibatch sizes).  `positions_packed_for_round` replays the other ranks' draws with this rank's own
    sizes, and `sliced_wasserstein` relies on all ranks drawing the same directions: both only hold in lock step."""
    if (world if world is not None else world_info(group)[1]) <= 1:
        return
    v = torch.tensor([float(x) for x in values], dtype=torch.float64)
    t = torch.cat([v, -v])
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)  # max(v) and -min(v)
    t = t.cpu()
    k = len(values)
    if not torch.equal(t[:k], -t[k:]):
        raise RuntimeError("sliced_wasserstein.API: ranks fed different numbers or sizes of minibatches (feeds, size hash: this rank %s, "
                           "max over ranks %s, min %s); a multi-rank evaluation needs the same sequence of minibatch sizes on every rank"
                           % (list(values), t[:k].tolist(), (-t[k:]).tolist()))


def direction_block(total_dirs: int, rank: int, world: int) -> slice:
    """The block of projection directions (columns of the [dim, repeats * dirs_per_repeat] matrix) owned by `rank`."""
    lo, hi = shard_range(total_dirs, rank, world)
    return slice(lo, hi)


def repeat_segments(block: slice, dirs_per_repeat: int):
    """Split a direction block into (repeat index, local start, local stop) pieces that do not straddle repeats."""
    out = []
    d = block.start
    while d < block.stop:
        rep = d // dirs_per_repeat
        stop = min(block.stop, (rep + 1) * dirs_per_repeat)
        out.append((rep, d - block.start, stop - block.start))
        d = stop
    return out


def _needs_host_staging(t: torch.Tensor, group=None) -> bool:
    """gloo (CPU tests, single-GPU multi-process tests) moves CUDA tensors through the host; NCCL does not."""
    return t.is_cuda and dist.get_backend(group) == "gloo"


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if world_info(group)[1] > 1:
        if _needs_host_staging(t, group):
            h = t.cpu()
            dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_gather_rows(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate the rows of every rank's [rows_r, k] tensor (rank-major).  Row counts may differ between ranks."""
    rank, world = world_info(group)
    if world == 1:
        return t
    if _needs_host_staging(t, group):
        return all_gather_rows(t.cpu(), group).to(t.device)
    t = t.contiguous()
    counts = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    sizes = [int(c.item()) for c in all_counts]
    if len(set(sizes)) == 1:
        out = torch.empty((sizes[0] * world,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
        return out
    # unequal shards: pad to the largest, gather, drop the padding
    biggest = max(sizes)
    padded = torch.zeros((biggest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    padded[: t.shape[0]] = t
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


class Gathered:
    """The result of `all_gather_rows_async`: `.result()` makes the current stream (NCCL) or the host (gloo) wait for the
    collective and returns the gathered tensor.  Until then the collective runs beside whatever is enqueued on the
    compute stream -- the kernels of the next `feed` calls, the sort of the previous pyramid level."""

    def __init__(self, out: torch.Tensor, work=None, keep=None):
        self._out, self._work, self._keep = out, work, keep

    def result(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = self._keep = None
        return self._out


def all_gather_rows_async(t: torch.Tensor, group=None, equal_rows: bool = False) -> Gathered:
    """`all_gather_rows` that does not wait: equal shards over NCCL are gathered with `async_op=True` (the collective runs on
    NCCL's own stream, ordered after what the current stream holds at this point); every other case is gathered at once.
    `equal_rows`: the caller knows that every rank holds the same number of rows (`API.end` after its lock-step check) --
    otherwise the row counts are exchanged first, and reading them back waits for everything NCCL's stream still holds."""
    rank, world = world_info(group)
    if world == 1 or _needs_host_staging(t, group) or not t.is_cuda:
        return Gathered(all_gather_rows(t, group))
    t = t.contiguous()
    if equal_rows:
        sizes = [t.shape[0]] * world
    else:
        counts = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        sizes = [int(c.item()) for c in all_counts]
    if len(set(sizes)) != 1:
        return Gathered(all_gather_rows(t, group))
    out = torch.empty((sizes[0] * world,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    work = dist.all_gather_into_tensor(out, t, group=group, async_op=True)
    return Gathered(out, work, keep=t)


def combine_repeat_sums(partial: torch.Tensor, rows: int, dirs_per_repeat: int, group=None) -> float:
    """partial[r] = this rank's sum of |sorted A - sorted B| over its directions of repeat r (fp64).  Returns the
    reference's `np.mean` over repeats of the per-repeat fp32 means (sliced_wasserstein.py:49-51)."""
    all_reduce_sum_(partial, group)
    per_repeat = (partial / float(rows * dirs_per_repeat)).to(torch.float32)
    return float(np.float32(per_repeat.to(torch.float64).mean().item()))


def allreduce_gradients(params, group=None) -> None:
    """Data-parallel gradient exchange of the train-step harness (BASELINE configs[2]): one flat all-reduce (sum) of
    the gradients of `params`.  The reference divides its losses by the *global* batch size (wgan.py:130,157), so the
    sum over ranks is the gradient of the global batch -- no averaging."""
    if world_info(group)[1] == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    all_reduce_sum_(flat, group)
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset:offset + n].view_as(g))
        offset += n
