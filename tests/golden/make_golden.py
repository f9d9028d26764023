"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

* blur: the first 72 lines of /root/reference/pytorch/gaussianblur.py are exec'd verbatim
  (the module cannot be imported: its tail downloads MNIST and calls exit());
  C > 1 is obtained by folding channels into the batch (the reference is single-channel, Q5);
  gradients come from ATen autograd of the reference's own F.conv2d calls.
* SWD: /root/reference/sliced_wasserstein.py is imported unmodified; inputs are passed as
  copies because its pyramid builder overwrites float32 inputs (Q9).

Inputs are regenerated in the tests from numpy RandomState seeds (bit-stable across machines),
so only outputs are stored.  The reference has no tests/golden vectors of its own.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference_blur():
    with open(os.path.join(REF, "pytorch", "gaussianblur.py")) as f:
        head = "".join(f.readlines()[:72])
    ns = {}
    exec(compile(head, "reference:pytorch/gaussianblur.py[1:72]", "exec"), ns)
    return ns


def load_reference_swd():
    sys.path.insert(0, REF)
    import sliced_wasserstein  # noqa: E402

    return sliced_wasserstein


def blur_input(seed, shape):
    return (np.random.RandomState(seed).rand(*shape) * 2.0 - 1.0).astype(np.float32)


BLUR_CASES = [  # (name, shape NCHW, stds)
    ("mnist", (3, 1, 28, 28), [0.01, 0.05, 0.5, 1.0, 2.0, 5.0, 11.5, 23.0]),
    ("celeba", (1, 1, 128, 128), [0.5, 2.0, 5.0, 12.0, 23.0]),
    ("ragged", (2, 2, 37, 50), [0.3, 1.7, 4.0, 9.0, 30.0]),
    ("tiny", (2, 1, 5, 9), [0.4, 1.0, 3.0]),
    ("wide", (1, 1, 16, 300), [2.0, 12.0, 40.0]),
]


def make_blur(ns):
    GaussianBlur = ns["GaussianBlur"]
    out = {}
    for name, shape, stds in BLUR_CASES:
        n, c, h, w = shape
        x = blur_input(1234, shape)
        g = blur_input(4321, shape)
        for std in stds:
            mod = GaussianBlur(std)
            xt = torch.from_numpy(x).reshape(n * c, 1, h, w).requires_grad_(True)
            y = mod(xt)
            y.backward(torch.from_numpy(g).reshape(n * c, 1, h, w))
            key = "%s_std%g" % (name, std)
            out[key + "_y"] = y.detach().numpy().reshape(shape)
            out[key + "_gx"] = xt.grad.numpy().reshape(shape)
            out[key + "_taps"] = mod.kernel.numpy().reshape(-1)
    np.savez_compressed(os.path.join(HERE, "blur_torch.npz"), **out)

    # kernel-size / tap table over a dense std grid and several image sizes
    table = {}
    stds = [0.01, 0.1, 0.2, 0.23, 0.3, 0.4, 0.5, 0.6, 0.75, 1, 1.4, 2, 2.6, 3, 4, 5, 6.5, 8, 10, 12, 16, 20, 23, 23.5, 32, 48, 64]
    sizes = [(128, 128), (28, 28), (1024, 1024), (37, 50), (16, 300), (5, 9), (64, 64), (32, 32), (16, 16)]
    rows = []
    for std in stds:
        for h, w in sizes:
            mod = GaussianBlur(std)
            mod._image_size = min(h, w)
            k = mod.kernel_size
            rows.append((std, h, w, k))
            table["taps_%g_%d_%d" % (std, h, w)] = mod.kernel.numpy().reshape(-1)
    table["rows"] = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "blur_torch_taps.npz"), **table)


def swd_input(seed, shape):
    return (np.random.RandomState(seed).rand(*shape) * 2.0 - 1.0).astype(np.float32)


def make_swd(sw):
    out = {}
    # --- small, fully dumped case: 6 + 6 images of 3x64x64, fed as minibatches of 3 ---------
    reals = swd_input(11, (6, 3, 64, 64))
    fakes = (0.8 * swd_input(12, (6, 3, 64, 64)) + 0.1).astype(np.float32)
    pyr = sw.generate_laplacian_pyramid(reals.copy(), 3)
    for i, lvl in enumerate(pyr):
        out["s64_pyr%d" % i] = lvl
    out["s64_down"] = sw.pyr_down(reals.copy())
    out["s64_up"] = sw.pyr_up(out["s64_down"].copy())
    np.random.seed(7)
    d = sw.get_descriptors_for_minibatch(pyr[0], 7, 128)
    out["s64_desc0"] = d[:256].copy()  # first two images' worth
    np.random.seed(7)
    out["s64_px"] = np.random.randint(3, 64 - 3, size=(6 * 128,))
    out["s64_py"] = np.random.randint(3, 64 - 3, size=(6 * 128,))
    out["s64_final0"] = sw.finalize_descriptors(d.copy())[:256]

    def run_api(shape, reals, fakes, mb):
        np.random.seed(0)
        api = sw.API(shape)
        api.begin("reals")
        for i in range(0, len(reals), mb):
            api.feed("reals", reals[i : i + mb].copy())
        r0 = api.end("reals")
        api.begin("fakes")
        for i in range(0, len(fakes), mb):
            api.feed("fakes", fakes[i : i + mb].copy())
        r1 = api.end("fakes")
        return np.array(r0, dtype=np.float64), np.array(r1, dtype=np.float64)

    out["s64_api_reals"], out["s64_api_fakes"] = run_api((3, 64, 64), reals, fakes, 3)

    # --- CelebA-shaped case: 16 + 16 images of 3x128x128, minibatches of 8 -------------------
    reals = swd_input(21, (16, 3, 128, 128))
    fakes = np.clip(swd_input(22, (16, 3, 128, 128)) * 1.1 - 0.05, -1, 1).astype(np.float32)
    out["s128_api_reals"], out["s128_api_fakes"] = run_api((3, 128, 128), reals, fakes, 8)
    pyr = sw.generate_laplacian_pyramid(reals[:2].copy(), 4)
    for i, lvl in enumerate(pyr):
        out["s128_pyr%d_sum" % i] = np.array([lvl.astype(np.float64).sum(), np.abs(lvl.astype(np.float64)).sum()])
    out["s128_pyr3"] = pyr[3]

    # --- sliced_wasserstein() on its own, fixed A/B --------------------------------------------
    rs = np.random.RandomState(5)
    A = rs.randn(1000, 147).astype(np.float32)
    B = (rs.randn(1000, 147) * 1.2 + 0.1).astype(np.float32)
    np.random.seed(9)
    out["sw_AB"] = np.array([sw.sliced_wasserstein(A, B, 4, 128)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "swd.npz"), **out)


SWD_LARGE = dict(n=2048, mb=64, seed_reals=61, seed_fakes=62, np_seed=0)


def swd_large_inputs():
    """2 x 2048 images of 3 x 128 x 128 (BASELINE-like scale: 262,144 descriptors per level and set, i.e. 64 sort tiles per
    direction and 2048 row tiles of the tensor-core projection).  The GPU test regenerates them from the same seeds."""
    c = SWD_LARGE
    reals = swd_input(c["seed_reals"], (c["n"], 3, 128, 128))
    fakes = (0.9 * swd_input(c["seed_fakes"], (c["n"], 3, 128, 128))).astype(np.float32)
    fakes[:, :, ::2] *= 0.8
    return reals, fakes


def make_swd_large(sw):
    """The unmodified reference `API` on 2 x 2048 images (about two minutes on 8 cores): 2 x 5 floats."""
    c = SWD_LARGE
    reals, fakes = swd_large_inputs()
    np.random.seed(c["np_seed"])
    api = sw.API((3, 128, 128))
    api.begin("reals")
    for i in range(0, c["n"], c["mb"]):
        api.feed("reals", reals[i : i + c["mb"]].copy())
    r0 = api.end("reals")
    api.begin("fakes")
    for i in range(0, c["n"], c["mb"]):
        api.feed("fakes", fakes[i : i + c["mb"]].copy())
    r1 = api.end("fakes")
    np.savez_compressed(os.path.join(HERE, "swd_large.npz"), api_reals=np.array(r0, dtype=np.float64), api_fakes=np.array(r1, dtype=np.float64),
                        config=np.array([c["n"], c["mb"], c["seed_reals"], c["seed_fakes"], c["np_seed"]], dtype=np.int64))


def make_swd_downscale(sw):
    """`downscale_minibatch` (sliced_wasserstein.py:55-61; module surface, not called by `API`) on uint8-range images."""
    rs = np.random.RandomState(71)
    x = rs.randint(0, 256, size=(3, 3, 32, 48)).astype(np.uint8)
    out = {"x": x}
    for lod in (1, 2, 3):
        out["lod%d" % lod] = sw.downscale_minibatch(x, lod)
    np.savez_compressed(os.path.join(HERE, "swd_downscale.npz"), **out)


if __name__ == "__main__":
    if "--swd-downscale" in sys.argv:
        make_swd_downscale(load_reference_swd())
    elif "--swd-large" in sys.argv:  # slow; kept out of the default run
        make_swd_large(load_reference_swd())
    else:
        make_blur(load_reference_blur())
        make_swd(load_reference_swd())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
