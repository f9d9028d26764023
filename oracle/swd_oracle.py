"""CPU restatement (numpy only, no scipy) of the reference's sliced-Wasserstein metric.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative to
/root/reference/sliced_wasserstein.py.  Parity status: PINNED -- every function here is
checked against the unmodified reference module, run in the build container, by
tests/test_oracle_swd.py (live, when /root/reference exists) and against the committed
fixtures tests/golden/swd_*.npz (always).

Differences from the reference, on purpose:
  * inputs are never mutated (the reference's pyramid aliases and overwrites a float32 input,
    SURVEY Q9);
  * the random draws are taken from an explicit ``numpy.random.RandomState`` (default: the
    global legacy state the reference uses) so that tests can replay the exact stream.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# 5-tap binomial; the 5x5 filter of :65-70 is outer(B5, B5) / 256.
_B5 = np.array([1.0, 4.0, 6.0, 4.0, 1.0], dtype=np.float64)


def _rng(rng):
    return np.random if rng is None else rng


# --------------------------------------------------------------------------------------
# Laplacian pyramid  (:65-88)
# --------------------------------------------------------------------------------------
def _binomial5x5_mirror(x: np.ndarray, gain: float) -> np.ndarray:
    """x (*) outer(B5,B5) * gain/256 with REFLECT_101 borders, fp64 accumulate, fp32 store.

    scipy.ndimage.convolve(..., mode='mirror') as called at :74 and :81 accumulates in double
    and rounds once to the array dtype; every product and partial sum here is exact in double,
    so the summation order does not matter and the result is bit-identical.
    """
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (2, 2), (2, 2)), mode="reflect")
    h, w = x.shape[2], x.shape[3]
    acc = np.zeros(x.shape, dtype=np.float64)
    for dy in range(5):
        for dx in range(5):
            acc += (_B5[dy] * _B5[dx] * gain / 256.0) * xp[:, :, dy : dy + h, dx : dx + w]
    return acc.astype(F32)


def pyr_down(minibatch: np.ndarray) -> np.ndarray:
    """:72-74 -- blur then keep even rows/cols."""
    assert minibatch.ndim == 4
    return _binomial5x5_mirror(minibatch, 1.0)[:, :, ::2, ::2]


def pyr_up(minibatch: np.ndarray) -> np.ndarray:
    """:76-81 -- zero-stuff to 2x, blur with 4x gain."""
    assert minibatch.ndim == 4
    n, c, h, w = minibatch.shape
    z = np.zeros((n, c, 2 * h, 2 * w), dtype=minibatch.dtype)
    z[:, :, ::2, ::2] = minibatch
    return _binomial5x5_mirror(z, 4.0)


def generate_laplacian_pyramid(minibatch: np.ndarray, num_levels: int) -> list[np.ndarray]:
    """:83-88, without the reference's in-place aliasing of the caller's array."""
    levels = [np.array(minibatch, dtype=F32, copy=True)]
    for _ in range(1, num_levels):
        coarse = pyr_down(levels[-1])
        levels[-1] = levels[-1] - pyr_up(coarse)
        levels.append(coarse)
    return levels


def reconstruct_laplacian_pyramid(pyramid: list[np.ndarray]) -> np.ndarray:
    """:90-94."""
    out = pyramid[-1]
    for level in pyramid[-2::-1]:
        out = pyr_up(out) + level
    return out


def downscale_minibatch(minibatch: np.ndarray, lod: int) -> np.ndarray:
    """:55-61 -- 2x2 box filter ``lod`` times, round-to-uint8."""
    if lod == 0:
        return minibatch
    t = minibatch.astype(F32)
    for _ in range(lod):
        t = (t[:, :, 0::2, 0::2] + t[:, :, 0::2, 1::2] + t[:, :, 1::2, 0::2] + t[:, :, 1::2, 1::2]) * F32(0.25)
    return np.round(t).clip(0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------
# descriptors  (:13-34)
# --------------------------------------------------------------------------------------
def draw_patch_positions(n_images: int, height: int, width: int, nhood_size: int, nhoods_per_image: int, rng=None):
    """The two ``randint`` draws of :20-21, x (columns) first, then y (rows)."""
    half = nhood_size // 2
    count = nhoods_per_image * n_images
    px = _rng(rng).randint(half, width - half, size=(count, 1, 1, 1))
    py = _rng(rng).randint(half, height - half, size=(count, 1, 1, 1))
    return px.reshape(-1), py.reshape(-1)


def gather_descriptors(level: np.ndarray, px: np.ndarray, py: np.ndarray, nhood_size: int, nhoods_per_image: int) -> np.ndarray:
    """:13-23 -- out[m, c, i, j] = level[m // per_image, c, py[m] + j - half, px[m] + i - half].
    Axis 2 of the result walks x and axis 3 walks y (the reference's ogrid names)."""
    assert level.ndim == 4 and level.shape[1] == 3
    half = nhood_size // 2
    offs = np.arange(-half, half + 1)
    img = (np.arange(len(px)) // nhoods_per_image)[:, None, None, None]
    chan = np.arange(3)[None, :, None, None]
    cols = px[:, None, None, None] + offs[None, None, :, None]
    rows = py[:, None, None, None] + offs[None, None, None, :]
    return level[img, chan, rows, cols]


def get_descriptors_for_minibatch(minibatch, nhood_size, nhoods_per_image, rng=None):
    n, _, h, w = minibatch.shape
    px, py = draw_patch_positions(n, h, w, nhood_size, nhoods_per_image, rng)
    return gather_descriptors(minibatch, px, py, nhood_size, nhoods_per_image)


def finalize_descriptors(desc) -> np.ndarray:
    """:27-34 -- per-channel standardisation over (descriptor, i, j), population std."""
    if isinstance(desc, list):
        desc = np.concatenate(desc, axis=0)
    assert desc.ndim == 4
    desc = desc - np.mean(desc, axis=(0, 2, 3), keepdims=True)
    desc = desc / np.std(desc, axis=(0, 2, 3), keepdims=True)
    return desc.reshape(desc.shape[0], -1)


# --------------------------------------------------------------------------------------
# distance  (:38-51)
# --------------------------------------------------------------------------------------
def draw_directions(dim: int, dirs_per_repeat: int, rng=None) -> np.ndarray:
    """:42-44 -- randn in fp64, unit columns, cast to fp32."""
    d = _rng(rng).randn(dim, dirs_per_repeat)
    d /= np.sqrt(np.sum(np.square(d), axis=0, keepdims=True))
    return d.astype(F32)


def sliced_wasserstein_with_dirs(A: np.ndarray, B: np.ndarray, dirs_list) -> float:
    """:45-51 for a given list of direction matrices."""
    assert A.ndim == 2 and A.shape == B.shape
    per_repeat = []
    for dirs in dirs_list:
        pa = np.sort(A @ dirs, axis=0)
        pb = np.sort(B @ dirs, axis=0)
        per_repeat.append(np.mean(np.abs(pa - pb)))
    return np.mean(per_repeat)


def sliced_wasserstein(A, B, dir_repeats, dirs_per_repeat, rng=None) -> float:
    dirs = []
    per_repeat = []
    for _ in range(dir_repeats):  # draw-then-use per repeat, like the reference, to keep the stream order
        d = draw_directions(A.shape[1], dirs_per_repeat, rng)
        dirs.append(d)
        per_repeat.append(sliced_wasserstein_with_dirs(A, B, [d]))
    return np.mean(per_repeat)


# --------------------------------------------------------------------------------------
# API  (:98-133)
# --------------------------------------------------------------------------------------
class API:
    def __init__(self, image_shape, rng=None):
        self.nhood_size = 7
        self.nhoods_per_image = 128
        self.dir_repeats = 4
        self.dirs_per_repeat = 128
        self.resolutions = []
        self.rng = rng
        res = image_shape[1]
        while res >= 16:
            self.resolutions.append(res)
            res //= 2

    def get_metric_names(self):
        return ["SWDx1e3_%d" % res for res in self.resolutions] + ["SWDx1e3_avg"]

    def get_metric_formatting(self):
        return ["%-13.4f"] * len(self.get_metric_names())

    def begin(self, mode):
        assert mode in ["warmup", "reals", "fakes"]
        self.descriptors = [[] for _ in self.resolutions]

    def feed(self, mode, minibatch):
        pyramid = generate_laplacian_pyramid(minibatch, len(self.resolutions))
        for lod, level in enumerate(pyramid):
            self.descriptors[lod].append(
                get_descriptors_for_minibatch(level, self.nhood_size, self.nhoods_per_image, self.rng)
            )

    def end(self, mode):
        desc = [finalize_descriptors(d) for d in self.descriptors]
        del self.descriptors
        if mode in ["warmup", "reals"]:
            self.desc_real = desc
        dist = [
            sliced_wasserstein(dreal, dfake, self.dir_repeats, self.dirs_per_repeat, self.rng)
            for dreal, dfake in zip(self.desc_real, desc)
        ]
        dist = [d * 1e3 for d in dist]
        return dist + [np.mean(dist)]
