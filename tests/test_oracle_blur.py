"""CPU: pins oracle/blur_oracle.py against the golden vectors made from the unmodified reference
(tests/golden/make_golden.py), and against the live reference when it is present."""
import numpy as np
import pytest

from conftest import needs_reference, uniform_pm1
from oracle import blur_oracle as bo

# must mirror tests/golden/make_golden.py::BLUR_CASES
BLUR_CASES = [
    ("mnist", (3, 1, 28, 28), [0.01, 0.05, 0.5, 1.0, 2.0, 5.0, 11.5, 23.0]),
    ("celeba", (1, 1, 128, 128), [0.5, 2.0, 5.0, 12.0, 23.0]),
    ("ragged", (2, 2, 37, 50), [0.3, 1.7, 4.0, 9.0, 30.0]),
    ("tiny", (2, 1, 5, 9), [0.4, 1.0, 3.0]),
    ("wide", (1, 1, 16, 300), [2.0, 12.0, 40.0]),
]
FLAT = [(n, s, std) for n, s, stds in BLUR_CASES for std in stds]


def test_kernel_size_table_matches_reference(golden_taps):
    for std, h, w, k in golden_taps["rows"]:
        kk, r = bo.torch_rule_plan(float(std), int(h), int(w))
        assert kk == int(k), (std, h, w)
        assert r == kk // 2


def test_torch_taps_match_reference(golden_taps):
    for std, h, w, k in golden_taps["rows"]:
        ref = golden_taps["taps_%g_%d_%d" % (std, h, w)]
        got = bo.torch_rule_taps(float(std), int(k))
        assert got.shape == ref.shape
        # only the order of the fp32 normalising sum differs
        np.testing.assert_allclose(got, ref, rtol=3e-7, atol=1e-9)


def test_torch_rule_tiny_std_is_identity_taps():
    for std in (0.01, 0.1):  # exp underflows to exactly 0 in fp32
        np.testing.assert_array_equal(bo.taps_for(bo.RULE_TORCH, std, 128, 128), np.float32([0, 1, 0]))
    np.testing.assert_allclose(bo.taps_for(bo.RULE_TORCH, 0.2, 128, 128), [0, 1, 0], atol=1e-10)


@pytest.mark.parametrize("name,shape,std", FLAT)
def test_blur_forward_and_gradient_match_reference(golden_blur, name, shape, std):
    x = uniform_pm1(1234, shape)
    g = uniform_pm1(4321, shape)
    key = "%s_std%g" % (name, std)
    y = bo.gaussian_blur_torch_ref(x, std)
    scale = np.abs(golden_blur[key + "_y"]).max()
    assert np.abs(y - golden_blur[key + "_y"]).max() <= 2e-6 * max(scale, 1e-3)
    # backward of a self-adjoint operator is the operator itself
    gx = bo.gaussian_blur_torch_ref(g, std)
    scale = np.abs(golden_blur[key + "_gx"]).max()
    assert np.abs(gx - golden_blur[key + "_gx"]).max() <= 2e-6 * max(scale, 1e-3)


def test_tf_rule_tap_table():
    # (std, size) -> (k, radius, taps)
    expect = {
        (0.01, 128): (3, 1), (0.2, 128): (3, 1), (0.5, 128): (4, 2), (1, 128): (7, 3), (2, 128): (13, 6), (3, 128): (19, 9),
        (5, 128): (31, 15), (8, 128): (49, 24), (12, 128): (73, 36), (16, 128): (97, 48), (23, 128): (128, 64),
        (5, 28): (28, 14), (23, 28): (28, 14), (32, 1024): (193, 96), (64, 1024): (385, 192),
    }
    for (std, size), (k, r) in expect.items():
        kk, rr, s_eff = bo.tf_rule_plan(std, size, size)
        assert (kk, rr) == (k, r), (std, size)
        assert s_eff == pytest.approx((k - 1) / 6.0, rel=1e-6)
    # std' never falls below 1/3, so the TF blur is never the identity
    taps = bo.taps_for(bo.RULE_TF, 0.01, 128, 128)
    np.testing.assert_allclose(taps, [0.0109, 0.9783, 0.0109], atol=2e-4)


def test_tf_rule_impulse_response_and_layouts():
    h, w = 20, 24
    taps = bo.taps_for(bo.RULE_TF, 1.0, h, w)
    r = len(taps) // 2
    x = np.zeros((1, 1, h, w), np.float32)
    x[0, 0, 10, 12] = 1.0
    y = bo.blur_images_tf_ref(x, 1.0)  # last dim 24 -> NCHW
    np.testing.assert_allclose(y[0, 0, 10 - r : 10 + r + 1, 12 - r : 12 + r + 1], np.outer(taps, taps), rtol=1e-6, atol=1e-9)
    # NHWC with 3 channels == the NCHW result of the transposed tensor
    xs = uniform_pm1(3, (2, h, w, 3))
    y_nhwc = bo.blur_images_tf_ref(xs, 2.0)
    y_nchw = bo.blur_nchw(xs.transpose(0, 3, 1, 2), bo.taps_for(bo.RULE_TF, 2.0, h, w), "vh")
    np.testing.assert_allclose(y_nhwc.transpose(0, 3, 1, 2), y_nchw, rtol=0, atol=0)


def test_operator_is_self_adjoint_and_darkens_borders():
    x = uniform_pm1(1, (1, 1, 17, 23))
    y = uniform_pm1(2, (1, 1, 17, 23))
    bx = bo.gaussian_blur_torch_ref(x, 2.0)
    by = bo.gaussian_blur_torch_ref(y, 2.0)
    assert abs(float((bx * y).sum()) - float((x * by).sum())) < 1e-4
    ones = np.ones((1, 1, 64, 64), np.float32)
    b1 = bo.gaussian_blur_torch_ref(ones, 2.0)
    assert b1[0, 0, 0, 0] == pytest.approx(0.310, abs=2e-3) and b1[0, 0, 32, 32] == pytest.approx(1.0, abs=1e-6)


_LIVE_REFERENCE_SCRIPT = """
import sys
import numpy as np
import torch
with open("/root/reference/pytorch/gaussianblur.py") as f:
    head = "".join(f.readlines()[:72])
ns = {}
exec(compile(head, "ref_head", "exec"), ns)
data = np.load(sys.argv[1])
out = {}
for std in data["stds"]:
    out["y_%g" % std] = ns["GaussianBlur"](float(std))(torch.from_numpy(data["x"])).numpy()
np.savez(sys.argv[2], **out)
"""


@needs_reference
def test_live_reference_agrees_with_oracle(tmp_path):
    """The unmodified reference head runs in a CHILD process (isolated interpreter, empty environment, scratch working
    directory): /root/reference is untrusted content, whatever it does at import time stays out of the test process."""
    import subprocess
    import sys

    import torch

    x = uniform_pm1(77, (4, 1, 33, 40))
    stds = np.array([0.7, 3.3, 15.0])
    np.savez(tmp_path / "in.npz", x=x, stds=stds)
    subprocess.run([sys.executable, "-I", "-c", _LIVE_REFERENCE_SCRIPT, str(tmp_path / "in.npz"), str(tmp_path / "out.npz")],
                   check=True, cwd=str(tmp_path), env={"PATH": "/usr/bin:/bin", "OMP_NUM_THREADS": "4"}, timeout=300,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref = np.load(tmp_path / "out.npz")
    for std in stds:
        y_ref = ref["y_%g" % std]
        y = bo.gaussian_blur_torch_ref(x, float(std))
        assert np.abs(y - y_ref).max() <= 2e-6 * np.abs(y_ref).max()
        y_port = bo.blur_conv2d_port(torch.from_numpy(x), float(std)).numpy()
        assert np.abs(y_port - y_ref).max() <= 2e-6 * np.abs(y_ref).max()


@pytest.mark.parametrize("shape,std", [((2, 20, 24, 3), 2.0), ((1, 128, 128, 3), 5.0), ((2, 28, 28, 1), 23.0), ((1, 37, 50, 3), 0.01)])
def test_tf_rule_oracle_against_two_independent_convolutions(shape, std):
    """TensorFlow cannot run here, so the TF-rule oracle is pinned from two sides instead: its taps by the table of SURVEY
    A.3 (above), its arithmetic -- depthwise, stride 1, SAME zero padding, vertical pass then horizontal pass,
    gaussian_blur.py:116-130 -- by ATen's grouped conv2d and by scipy.ndimage.correlate1d(mode='constant'), two library
    implementations that share no code with the oracle's numpy loops."""
    import scipy.ndimage
    import torch
    import torch.nn.functional as F

    x = uniform_pm1(9, shape)
    n, h, w, c = shape
    taps = bo.taps_for(bo.RULE_TF, std, h, w)
    r = len(taps) // 2
    got = bo.blur_images_tf_ref(x, std)
    assert got.shape == x.shape and got.dtype == np.float32
    # ATen: depthwise convolutions in fp64 (an odd kernel with padding r is exactly SAME padding)
    t = torch.from_numpy(x.astype(np.float64)).permute(0, 3, 1, 2)
    k = torch.from_numpy(taps.astype(np.float64))
    yv = F.conv2d(t, k.view(1, 1, -1, 1).repeat(c, 1, 1, 1), padding=(r, 0), groups=c)
    yt = F.conv2d(yv, k.view(1, 1, 1, -1).repeat(c, 1, 1, 1), padding=(0, r), groups=c).permute(0, 2, 3, 1).numpy()
    scale = np.abs(yt).max()
    assert np.abs(got - yt).max() <= 2e-6 * scale
    # scipy: correlate along H, then along W, zeros outside
    ys = scipy.ndimage.correlate1d(x.astype(np.float64), taps.astype(np.float64), axis=1, mode="constant", cval=0.0)
    ys = scipy.ndimage.correlate1d(ys, taps.astype(np.float64), axis=2, mode="constant", cval=0.0)
    assert np.abs(got - ys).max() <= 2e-6 * scale
    assert np.abs(yt - ys).max() <= 1e-12 * scale  # (the two references agree with each other to fp64 rounding)
