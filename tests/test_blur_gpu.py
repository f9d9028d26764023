"""GPU parity tests of the blur: CUDA path (through the public modules -> custom op -> C ABI) against
the CPU oracle and the golden vectors of the unmodified reference.

Tolerance (BASELINE.json north_star): <= 1e-5 relative on blurred pixels and gradients, measured as
max|got - want| / max|want| (the blurred field of a zero-mean input has values arbitrarily close to 0,
so a per-element relative error is not meaningful), plus an element-wise allclose(rtol=1e-5, atol=2e-6*scale).
"""
import os

import numpy as np
import pytest
import torch

from conftest import uniform_pm1
from oracle import blur_oracle as bo

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_close(got, want, what=""):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else got
    scale = max(float(np.abs(want).max()), 1e-6)
    err = float(np.abs(got - want).max()) / scale
    assert err <= RTOL, "%s: relative error %.3g > %g" % (what, err, RTOL)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=2e-6 * scale, err_msg=what)


from test_oracle_blur import FLAT  # noqa: E402  (same cases as the golden file)


@pytest.mark.parametrize("name,shape,std", FLAT)
def test_torch_module_matches_reference_golden(golden_blur, name, shape, std):
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = _dev(uniform_pm1(1234, shape)).requires_grad_(True)
    g = _dev(uniform_pm1(4321, shape))
    mod = GaussianBlur(std).cuda()
    y = mod(x)
    y.backward(g)
    key = "%s_std%g" % (name, std)
    assert_close(y, golden_blur[key + "_y"], key + " forward")
    assert_close(x.grad, golden_blur[key + "_gx"], key + " input gradient")
    assert mod.kernel_size == len(golden_blur[key + "_taps"])


def test_device_taps_match_reference(golden_taps):
    from blurred_gan_b200 import ops

    for std, h, w, k in golden_taps["rows"]:
        ref = golden_taps["taps_%g_%d_%d" % (std, h, w)]
        got = ops.device_taps(ops.RULE_TORCH, float(std), int(k) // 2, "cuda").cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-9, err_msg="std=%g k=%d" % (std, k))
    for std, size in [(0.01, 128), (0.5, 128), (5, 128), (23, 128), (64, 1024)]:
        k, r, s = bo.tf_rule_plan(std, size, size)
        got = ops.device_taps(ops.RULE_TF, s, r, "cuda").cpu().numpy()
        np.testing.assert_allclose(got, bo.tf_rule_taps(s, 2 * r + 1), rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("std", [0.01, 0.5, 1.0, 2.0, 3.0, 5.0, 8.0, 12.0, 16.0, 23.0])
def test_celeba_shape_nchw_torch_rule(std):
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = uniform_pm1(0, (4, 3, 128, 128))
    want = bo.gaussian_blur_torch_ref(x, std)
    assert_close(GaussianBlur(std).cuda()(_dev(x)), want, "std=%g" % std)


@pytest.mark.parametrize("std", [0.01, 0.5, 1.0, 2.0, 5.0, 8.0, 16.0, 23.0])
@pytest.mark.parametrize("layout", ["NHWC", "NCHW"])
def test_celeba_shape_tf_rule(std, layout):
    from blurred_gan_b200.gaussian_blur import GaussianBlur2D, blur_images

    shape = (3, 128, 128, 3) if layout == "NHWC" else (3, 3, 128, 128)
    x = uniform_pm1(1, shape)
    want = bo.blur_images_tf_ref(x, std)
    assert_close(blur_images(_dev(x), std), want, "%s std=%g" % (layout, std))
    layer = GaussianBlur2D(std).cuda()
    assert_close(layer(_dev(x)), want, "layer %s std=%g" % (layout, std))


def test_mnist_shape_annealed_std():
    """BASELINE configs[0]: 64 x 28 x 28 x 1 with std annealed 23 -> 0, both APIs."""
    from blurred_gan_b200.gaussian_blur import GaussianBlur2D
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = uniform_pm1(2, (64, 1, 28, 28))
    xd = _dev(x)
    mod = GaussianBlur(23.0).cuda()
    layer = GaussianBlur2D(23.0).cuda()
    x_nhwc = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    for std in np.linspace(23.0, 0.02, 12):
        std = float(std)
        mod.set_std(std)
        assert_close(mod(xd), bo.gaussian_blur_torch_ref(x, std), "torch std=%g" % std)
        layer.assign(std)
        assert_close(layer(_dev(x_nhwc)), bo.blur_images_tf_ref(x_nhwc, std), "tf std=%g" % std)


@pytest.mark.parametrize("shape,layout", [((2, 3, 37, 50), "NCHW"), ((2, 31, 45, 3), "NHWC"), ((3, 2, 5, 9), "NCHW"),
                                          ((1, 1, 16, 300), "NCHW"), ((2, 130, 66, 1), "NHWC"), ((1, 5, 200, 36), "NCHW"),
                                          ((1, 3, 129, 131), "NCHW"), ((1, 40, 260, 3), "NHWC")])
def test_ragged_shapes_both_rules(shape, layout):
    from blurred_gan_b200.gaussian_blur import blur_images
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = uniform_pm1(5, shape)
    for std in (0.4, 1.7, 4.0, 9.0, 30.0):
        if bo.get_data_format(shape) == layout:
            assert_close(blur_images(_dev(x), std), bo.blur_images_tf_ref(x, std), "tf %s std=%g" % (shape, std))
        if layout == "NCHW":
            assert_close(GaussianBlur(std).cuda()(_dev(x)), bo.gaussian_blur_torch_ref(x, std), "torch %s std=%g" % (shape, std))


@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_non_finite_pixel_stays_in_its_plane(bad):
    """what an Inf / NaN pixel does depends on the kernel -- its tap neighbourhood on the FFMA path (like the
    reference's convolution), its whole plane on the tensor-core path (0 x NaN inside the MMAs) -- but it never leaves its plane."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = uniform_pm1(8, (2, 3, 128, 128))
    clean = x.copy()
    x[0, 1, 40, 77] = bad
    for std, path in ((0.5, 5), (5.0, 4)):
        y = GaussianBlur(std).cuda()(_dev(x)).cpu().numpy()
        assert ops.last_blur_path() == path
        want = bo.gaussian_blur_torch_ref(clean, std)
        assert not np.isfinite(y[0, 1, 40, 77])
        for n in range(2):
            for c in range(3):
                if (n, c) != (0, 1):
                    assert np.isfinite(y[n, c]).all()
                    assert_close(y[n, c], want[n, c], "plane (%d, %d) next to a non-finite one, std=%g" % (n, c, std))
        if path == 5:  # FFMA: only the (2r + 1)^2 neighbourhood (radius 1 at std 0.5)
            finite = np.isfinite(y[0, 1])
            assert (~finite).sum() <= 9 and finite[:38].all() and finite[43:].all()


@pytest.mark.parametrize("std", [8.0, 32.0, 64.0])
def test_large_image_wide_halo(std, monkeypatch):
    """BASELINE configs[3] shape (1024 x 1024, std up to 64), reduced batch for the CPU oracle.  NCHW planes run on the
    two-pass tensor-core kernel (path 6), NHWC images on the same kernel between two layout conversions (path 7); with it
    disabled, on the CUDA-core tile kernel."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import blur_images
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = uniform_pm1(6, (1, 2, 1024, 1024))
    x[0, 1] *= 1e-5
    want = bo.gaussian_blur_torch_ref(x, std)
    got = GaussianBlur(std).cuda()(_dev(x))
    assert ops.last_blur_path() == 6
    for c in range(2):
        assert_close(got[0, c], want[0, c], "torch 1024 plane %d std=%g (tensor-core axis passes)" % (c, std))
    monkeypatch.setenv("BGAN_BLUR_NO_AXIS_TC", "1")
    assert_close(GaussianBlur(std).cuda()(_dev(x))[0, 0], want[0, 0], "torch 1024 std=%g (tile kernel)" % std)
    if std >= 64.0:
        assert ops.last_blur_path() == 3  # row pass + column pass through scratch
    monkeypatch.delenv("BGAN_BLUR_NO_AXIS_TC")
    # NHWC (the reference's own layout, gaussian_blur.py:111-130): interleaved -> planar, the same two passes, planar -> interleaved
    xh = uniform_pm1(7, (1, 1024, 1024, 3))
    xh[0, :, :, 2] *= 3e-4
    want_h = bo.blur_images_tf_ref(xh, std)
    got_h = blur_images(_dev(xh), std)
    assert ops.last_blur_path() == 7
    for c in range(3):
        assert_close(got_h[0, :, :, c], want_h[0, :, :, c], "tf NHWC 1024 channel %d std=%g" % (c, std))
    monkeypatch.setenv("BGAN_BLUR_NO_AXIS_TC", "1")
    assert_close(blur_images(_dev(xh[:, :256]), std), bo.blur_images_tf_ref(xh[:, :256], std), "tf NHWC 256 x 1024 std=%g (tile kernel)" % std)
    assert ops.last_blur_path() in (1, 2, 3)


@pytest.mark.parametrize("shape,std,unit_tiles", [((40, 1, 1024, 256), 20.0, 0), ((40, 1, 256, 1024), 7.0, 0), ((24, 1, 512, 512), 12.0, 1),
                                                   ((24, 1, 512, 768), 30.0, 2), ((5, 1, 1024, 1024), 64.0, 0), ((3, 1, 640, 1024), 50.0, 1),
                                                   ((4, 1, 896, 640), 75.0, 0)])
def test_axis_tensor_core_path_many_work_units_per_cta(shape, std, unit_tiles, monkeypatch):
    """Persistent CTAs that take several work units in a row -- whole strips (radius <= 128: 320 strips on 107 CTAs in the first
    pass of the first case), pairs of tiles (radius > 128; 7 and 5 tiles per strip give a last unit of one tile) or, with the
    unit size forced, single tiles and pairs at radius <= 128: the hand-over of row scales, accumulator slots, operand sets
    and output boxes from one unit to the next.  Every plane is checked, the planes have different magnitudes; the TF rule at
    std 50 / 64 / 75 has radius 150 / 192 / 225 (nb = 2)."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import blur_images

    if unit_tiles:
        monkeypatch.setenv("BGAN_BLUR_AXIS_UNIT_TILES", str(unit_tiles))
    x = uniform_pm1(77, shape)
    x *= (10.0 ** np.random.RandomState(3).uniform(-4, 2, size=(shape[0], 1, 1, 1))).astype(np.float32)
    want = bo.blur_images_tf_ref(np.ascontiguousarray(x.transpose(0, 2, 3, 1)), std).transpose(0, 3, 1, 2)
    # (NCHW input of more than 3 channels-last... the functional entry point sees a 4-D tensor whose last dim is not 1 or 3: NCHW)
    got = blur_images(_dev(x), std)
    assert ops.last_blur_path() == 6
    for n in range(shape[0]):
        assert_close(got[n], want[n], "plane %d of %s std=%g" % (n, shape, std))


@pytest.mark.parametrize("shape,std", [((2, 1, 256, 256), 6.0), ((1, 3, 300, 260), 9.0), ((1, 2, 512, 384), 40.0), ((3, 1, 260, 516), 20.0)])
def test_axis_tensor_core_path_ragged_planes(shape, std):
    """Planes that are not multiples of the 128 x 128 tile, both tap rules, forward and backward, asymmetric explicit taps."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import blur_images
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = uniform_pm1(50, shape)
    x[0, 0, :, : shape[3] // 2] *= 1e-4  # rows and columns of very different magnitude
    xd = _dev(x).requires_grad_(True)
    y = GaussianBlur(std).cuda()(xd)
    assert ops.last_blur_path() == 6
    assert_close(y, bo.gaussian_blur_torch_ref(x, std), "axis path torch rule %s std=%g" % (shape, std))
    g = uniform_pm1(51, shape)
    y.backward(_dev(g))
    assert_close(xd.grad, bo.gaussian_blur_torch_ref(g, std), "axis path backward %s std=%g" % (shape, std))
    assert_close(blur_images(_dev(x), std), bo.blur_images_tf_ref(x, std), "axis path tf rule %s std=%g" % (shape, std))
    asym = np.linspace(0.01, 0.05, 33).astype(np.float32)  # radius 16, not symmetric
    ya = ops.blur2d_taps(_dev(x), torch.from_numpy(asym).cuda(), ops.LAYOUT_NCHW)
    assert ops.last_blur_path() == 6
    assert_close(ya, bo.blur_nchw(x, asym), "axis path, asymmetric taps %s" % (shape,))


def test_tma_and_plain_staging_agree_bitwise():
    from blurred_gan_b200 import ops
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = _dev(uniform_pm1(8, (5, 3, 128, 128)))
    os.environ["BGAN_BLUR_NO_TC"] = os.environ["BGAN_BLUR_NO_DIRECT"] = "1"  # this test is about the CUDA-core tile kernel
    for std in (0.5, 2.0, 5.0, 23.0):
        mod = GaussianBlur(std).cuda()
        os.environ.pop("BGAN_BLUR_NO_TMA", None)
        y_tma = mod(x)
        path_tma = ops.last_blur_path()
        os.environ["BGAN_BLUR_NO_TMA"] = "1"
        try:
            y_plain = mod(x)
            path_plain = ops.last_blur_path()
        finally:
            os.environ.pop("BGAN_BLUR_NO_TMA", None)
            if std == 23.0:
                os.environ.pop("BGAN_BLUR_NO_TC", None)
                os.environ.pop("BGAN_BLUR_NO_DIRECT", None)
        assert path_tma == 1 and path_plain == 2
        assert torch.equal(y_tma, y_plain)


@pytest.mark.parametrize("th,tw", [(32, 64), (64, 128), (16, 32), (128, 64)])
def test_tilings_agree(th, tw):
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = _dev(uniform_pm1(9, (3, 2, 128, 128)))
    for std in (1.0, 5.0):
        mod = GaussianBlur(std).cuda()
        os.environ["BGAN_BLUR_NO_TC"] = os.environ["BGAN_BLUR_NO_DIRECT"] = "1"
        try:
            want = mod(x)
            os.environ["BGAN_BLUR_TH"], os.environ["BGAN_BLUR_TW"] = str(th), str(tw)
            got = mod(x)
        finally:
            os.environ.pop("BGAN_BLUR_TH", None), os.environ.pop("BGAN_BLUR_TW", None), os.environ.pop("BGAN_BLUR_NO_TC", None)
            os.environ.pop("BGAN_BLUR_NO_DIRECT", None)
        assert_close(got, want.cpu().numpy(), "tile %dx%d std=%g" % (th, tw, std))


def test_double_backward_gradient_penalty_pattern():
    """The WGAN-GP critic step differentiates through the input-gradient (reference wgan.py:234-246, :140)."""
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    std = 2.0
    x_np, w_np, z_np = uniform_pm1(10, (2, 3, 32, 40)), uniform_pm1(11, (2, 3, 32, 40)), uniform_pm1(12, (2, 3, 32, 40))
    x = _dev(x_np).requires_grad_(True)
    w = _dev(w_np).requires_grad_(True)
    mod = GaussianBlur(std).cuda()
    y = mod(x)
    (gx,) = torch.autograd.grad((y * w).sum(), x, create_graph=True)  # gx = B w
    assert_close(gx, bo.gaussian_blur_torch_ref(w_np, std), "first-order")
    (gw,) = torch.autograd.grad((gx * _dev(z_np)).sum(), w)  # d/dw <B w, z> = B z
    assert_close(gw, bo.gaussian_blur_torch_ref(z_np, std), "double backward")


def test_explicit_taps_op_and_asymmetric_adjoint():
    from blurred_gan_b200 import ops

    x_np = uniform_pm1(13, (2, 2, 24, 28))
    taps = bo.taps_for(bo.RULE_TORCH, 2.0, 24, 28)
    got = ops.blur2d_taps(_dev(x_np), _dev(taps), ops.LAYOUT_NCHW)
    assert_close(got, bo.blur_nchw(x_np, taps), "explicit symmetric taps")
    asym = np.float32([0.1, 0.2, 0.4, 0.25, 0.05])
    x = _dev(x_np).requires_grad_(True)
    y = ops.blur2d_taps(x, _dev(asym), ops.LAYOUT_NCHW)
    assert_close(y, bo.blur_nchw(x_np, asym), "explicit asymmetric taps")
    g_np = uniform_pm1(14, x_np.shape)
    y.backward(_dev(g_np))
    assert_close(x.grad, bo.blur_nchw(g_np, asym[::-1].copy()), "adjoint of asymmetric taps")


@pytest.mark.parametrize("std", [0.5, 5.0, 23.0])
def test_full_size_properties(std):
    """BASELINE configs[1] at full size (256 x 3 x 128 x 128): size-independent properties."""
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    torch.manual_seed(0)
    x = torch.rand(256, 3, 128, 128, device="cuda") * 2 - 1
    y = torch.rand(256, 3, 128, 128, device="cuda") * 2 - 1
    mod = GaussianBlur(std).cuda()
    bx, by = mod(x), mod(y)
    # self-adjoint: <Bx, y> == <x, By>
    lhs, rhs = (bx.double() * y.double()).sum().item(), (x.double() * by.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    # linear
    z = mod(2.0 * x - 0.5 * y)
    assert (z - (2.0 * bx - 0.5 * by)).abs().max().item() <= 1e-5 * max(bx.abs().max().item(), 1e-3) * 3
    # every plane is treated alike: plane 700 alone == plane 700 of the batch
    single = mod(x.view(768, 1, 128, 128)[700:701])
    assert torch.equal(single[0, 0], bx.view(768, 128, 128)[700])
    # a few planes against the oracle
    idx = [0, 383, 767]
    want = bo.gaussian_blur_torch_ref(x.view(768, 1, 128, 128)[idx].cpu().numpy(), std)
    assert_close(bx.view(768, 1, 128, 128)[idx], want, "full-size spot check")


@pytest.mark.parametrize("std", [8.0, 64.0])
def test_full_size_properties_large_images(std):
    """BASELINE configs[3] at full size (32 x 3 x 1024 x 1024, TF rule): size-independent properties of the two-pass
    tensor-core kernel (whole-strip units at std 8, pairs of tiles at std 64), and the NHWC route against the NCHW one."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import blur_images

    torch.manual_seed(0)
    x = torch.rand(32, 3, 1024, 1024, device="cuda") * 2 - 1
    y = torch.rand(32, 3, 1024, 1024, device="cuda") * 2 - 1
    bx, by = blur_images(x, std), blur_images(y, std)
    assert ops.last_blur_path() == 6
    # self-adjoint: <Bx, y> == <x, By>
    lhs, rhs = (bx.double() * y.double()).sum().item(), (x.double() * by.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    # linear
    z = blur_images(2.0 * x - 0.5 * y, std)
    assert (z - (2.0 * bx - 0.5 * by)).abs().max().item() <= 1e-5 * max(bx.abs().max().item(), 1e-3) * 3
    del z, by, y
    # every plane is treated alike: plane 77 alone (one strip unit per CTA) == plane 77 of the batch (six units per CTA)
    single = blur_images(x.view(96, 1, 1024, 1024)[77:78].reshape(1, 1, 1024, 1024).expand(1, 2, 1024, 1024).contiguous(), std)
    assert torch.equal(single[0, 0], bx.view(96, 1024, 1024)[77]) and torch.equal(single[0, 1], single[0, 0])
    # NHWC images of the same data go through the planar route (path 7): the same bits
    xh = x[:4].permute(0, 2, 3, 1).contiguous()
    yh = blur_images(xh, std)
    assert ops.last_blur_path() == 7
    assert torch.equal(yh.permute(0, 3, 1, 2), bx[:4])
    # one plane against the oracle
    xs = x[31:32, 2:3].permute(0, 2, 3, 1).cpu().numpy()
    assert_close(bx[31, 2], bo.blur_images_tf_ref(np.ascontiguousarray(xs), std)[0, :, :, 0], "full-size spot check, std=%g" % std)


def test_ones_image_border_profile_and_errors():
    from blurred_gan_b200 import ops
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    y = GaussianBlur(2.0).cuda()(torch.ones(1, 1, 64, 64, device="cuda"))
    assert y[0, 0, 0, 0].item() == pytest.approx(0.310, abs=2e-3) and y[0, 0, 32, 32].item() == pytest.approx(1.0, abs=1e-6)
    with pytest.raises(RuntimeError):
        ops.blur2d(torch.zeros(2, 8, 8, 2, device="cuda"), ops.LAYOUT_NHWC, ops.RULE_TF, 1.0, 1)  # NHWC needs C in {1,3}
    assert GaussianBlur(1.0).cuda()(torch.zeros(0, 3, 8, 8, device="cuda")).shape == (0, 3, 8, 8)
    # non-contiguous input is accepted
    xt = torch.rand(2, 16, 20, 3, device="cuda").permute(0, 3, 1, 2)
    assert_close(GaussianBlur(1.0).cuda()(xt), bo.gaussian_blur_torch_ref(xt.cpu().numpy(), 1.0), "permuted view")


@pytest.mark.parametrize("std", [0.01, 0.5, 2.0, 5.0, 12.0, 23.0, 40.0])
def test_tensor_core_path_matches_oracle_and_fma_path(std, monkeypatch):
    """128 x 128 planes run on tcgen05 (fp16 split, fp32 accumulate in TMEM): same 1e-5 bar, any input magnitude."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import blur_images
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    monkeypatch.setenv("BGAN_BLUR_NO_DIRECT", "1")  # small radii would otherwise take the FFMA path (blur_direct.cu)
    x = uniform_pm1(20, (7, 3, 128, 128))
    x[1] *= 1e-6   # tiny plane magnitudes (gradients)
    x[2] *= 3e4    # beyond fp16 range without the per-plane power-of-two scale
    x[3, 0] = 0.0  # an all-zero plane
    x[4, 1, 5:9, 100:] += 50.0
    mod = GaussianBlur(std).cuda()
    got = mod(_dev(x))
    assert ops.last_blur_path() == 4
    os.environ["BGAN_BLUR_NO_TC"] = "1"
    try:
        fma = mod(_dev(x))
        assert ops.last_blur_path() in (1, 2)
    finally:
        os.environ.pop("BGAN_BLUR_NO_TC", None)
    want = bo.gaussian_blur_torch_ref(x, std)
    for n in range(x.shape[0]):  # per image: magnitudes differ by 10 orders
        assert_close(got[n], want[n], "tensor-core path, image %d, std=%g" % (n, std))
        assert_close(fma[n], want[n], "fma path, image %d, std=%g" % (n, std))
    assert torch.count_nonzero(got[3, 0]).item() == 0
    xt = uniform_pm1(21, (2, 3, 128, 128))
    got_tf = blur_images(_dev(xt), std)
    assert ops.last_blur_path() == 4
    assert_close(got_tf, bo.blur_images_tf_ref(xt, std), "tensor-core path, tf rule, std=%g" % std)


@pytest.mark.parametrize("planes", [1, 2, 3, 147, 148, 149, 297, 445, 600])
def test_tensor_core_pipeline_plane_counts(planes):
    """The two-planes-in-flight pipeline of blur_tc128_kernel has separate code for the first, second and later planes of a
    CTA and a balanced grid (ceil(planes / rounds) CTAs): 1, 2, 3, 4 and 5 planes per CTA, with and without a short CTA.
    Every plane has its own magnitude (per-plane power-of-two scales travel through a 4-deep ring).  Checked against the
    CUDA-core path (itself checked against the oracle), bit pattern of the plane order included."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    torch.manual_seed(planes)
    x = torch.rand(planes, 1, 128, 128, device="cuda") * 2 - 1
    x *= torch.logspace(-4, 4, planes, device="cuda").view(-1, 1, 1, 1)[torch.randperm(planes, device="cuda")]
    mod = GaussianBlur(5.0).cuda()
    got = mod(x)
    assert ops.last_blur_path() == 4
    os.environ["BGAN_BLUR_NO_TC"] = "1"
    try:
        ref = mod(x)
        assert ops.last_blur_path() in (1, 2)
    finally:
        os.environ.pop("BGAN_BLUR_NO_TC", None)
    err = (got - ref).abs().amax(dim=(1, 2, 3)) / ref.abs().amax(dim=(1, 2, 3))
    assert err.max().item() <= 1e-5, "plane %d of %d: %g" % (int(err.argmax()), planes, err.max().item())
    # one oracle check on top (first and last plane)
    for n in (0, planes - 1):
        assert_close(got[n:n + 1], bo.gaussian_blur_torch_ref(x[n:n + 1].cpu().numpy(), 5.0), "plane %d of %d" % (n, planes))


@pytest.mark.parametrize("std", [0.5, 3.0, 23.0])
def test_tensor_core_path_nhwc(std, monkeypatch):
    """NHWC 128 x 128 x 3 (the layout of the reference's Keras layer): the compute warps gather one channel plane with 4-byte
    cp.async (TMA cannot walk the innermost dimension with an element stride), the stores scatter with stride 3."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import GaussianBlur2D

    monkeypatch.setenv("BGAN_BLUR_NO_DIRECT", "1")
    x = uniform_pm1(30, (5, 128, 128, 3))
    x[1] *= 1e-5
    layer = GaussianBlur2D(std).cuda()
    xd = _dev(x).requires_grad_(True)
    y = layer(xd)
    assert ops.last_blur_path() == 4
    want = bo.blur_images_tf_ref(x, std)
    for n in range(x.shape[0]):
        assert_close(y[n], want[n], "NHWC tensor-core path image %d std=%g" % (n, std))
    g = uniform_pm1(31, x.shape)
    y.backward(_dev(g))
    assert_close(xd.grad, bo.blur_images_tf_ref(g, std), "NHWC tensor-core backward std=%g" % std)
    x1 = uniform_pm1(32, (4, 128, 128, 1))
    assert_close(layer(_dev(x1)), bo.blur_images_tf_ref(x1, std), "NHWC C=1 std=%g" % std)


@pytest.mark.parametrize("std", [0.01, 0.3, 0.5, 0.7, 1.0, 1.3])
def test_direct_path_small_radius(std):
    """128 x 128 images with radius <= 4 (the small-sigma end of the annealing) run on the FFMA kernel of blur_direct.cu:
    both layouts, both tap rules, forward and backward, any plane magnitude."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import GaussianBlur2D
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    x = uniform_pm1(40, (5, 3, 128, 128))
    x[1] *= 1e-6
    x[2] *= 3e4
    x[3, 0] = 0.0
    mod = GaussianBlur(std).cuda()
    k_torch, r_torch = bo.torch_rule_plan(std, 128, 128)
    xd = _dev(x).requires_grad_(True)
    y = mod(xd)
    assert ops.last_blur_path() == (5 if r_torch <= 4 else 4)
    want = bo.gaussian_blur_torch_ref(x, std)
    for n in range(x.shape[0]):
        assert_close(y[n], want[n], "direct path NCHW image %d std=%g" % (n, std))
    g = uniform_pm1(41, x.shape)
    y.backward(_dev(g))
    assert_close(xd.grad, bo.gaussian_blur_torch_ref(g, std), "direct path NCHW backward std=%g" % std)
    # NHWC, tf rule (k = floor(6 std) + 1): radius 1 .. 4 for std <= 1.3
    xh = uniform_pm1(42, (6, 128, 128, 3))
    xh[4] *= 1e-4
    layer = GaussianBlur2D(std).cuda()
    xhd = _dev(xh).requires_grad_(True)
    yh = layer(xhd)
    assert ops.last_blur_path() == 5
    wanth = bo.blur_images_tf_ref(xh, std)
    for n in range(xh.shape[0]):
        assert_close(yh[n], wanth[n], "direct path NHWC image %d std=%g" % (n, std))
    gh = uniform_pm1(43, xh.shape)
    yh.backward(_dev(gh))
    assert_close(xhd.grad, bo.blur_images_tf_ref(gh, std), "direct path NHWC backward std=%g" % std)
    # one image / odd image counts (NHWC units are half images)
    x1 = uniform_pm1(44, (1, 128, 128, 3))
    assert_close(layer(_dev(x1)), bo.blur_images_tf_ref(x1, std), "direct path NHWC single image std=%g" % std)
    x1c = uniform_pm1(45, (3, 128, 128, 1))
    assert_close(layer(_dev(x1c)), bo.blur_images_tf_ref(x1c, std), "direct path NHWC C=1 std=%g" % std)


def test_direct_path_explicit_asymmetric_taps():
    """Explicit taps on the direct path: correlation orientation in both passes, adjoint = reversed taps."""
    from blurred_gan_b200 import ops

    asym = np.array([0.05, 0.1, 0.15, 0.3, 0.25, 0.1, 0.05], dtype=np.float32)  # radius 3
    x_np = uniform_pm1(46, (4, 3, 128, 128))
    x = _dev(x_np).requires_grad_(True)
    y = ops.blur2d_taps(x, torch.from_numpy(asym).cuda(), ops.LAYOUT_NCHW)
    assert ops.last_blur_path() == 5
    assert_close(y, bo.blur_nchw(x_np, asym), "direct path, asymmetric taps")
    g_np = uniform_pm1(47, x_np.shape)
    y.backward(_dev(g_np))
    assert_close(x.grad, bo.blur_nchw(g_np, asym[::-1].copy()), "direct path, adjoint of asymmetric taps")
    xh_np = uniform_pm1(48, (3, 128, 128, 3))
    yh = ops.blur2d_taps(_dev(xh_np), torch.from_numpy(asym).cuda(), ops.LAYOUT_NHWC)
    assert ops.last_blur_path() == 5
    assert_close(yh, bo.blur_nhwc(xh_np, asym), "direct path NHWC, asymmetric taps")


def test_randomised_kernel_selection_and_parity():
    """Seeded random shapes, layouts, radii and (asymmetric) explicit taps across every kernel of the dispatch table:
    whichever kernel runs, the result is the oracle's zero-padded correlation along W then H."""
    from blurred_gan_b200 import ops

    rs = np.random.RandomState(2024)
    seen = set()
    cases = []
    for _ in range(6):   # 128 x 128: direct (r <= 5 / 4) or tensor-core kernel
        cases.append(("NCHW", (int(rs.randint(1, 6)), int(rs.randint(1, 4)), 128, 128), int(rs.randint(1, 90))))
        cases.append(("NHWC", (int(rs.randint(1, 6)), 128, 128, 3), int(rs.randint(1, 70))))
    for _ in range(5):   # large planes: two-pass tensor-core kernel (r >= 12), tile kernel otherwise
        h, w = 4 * int(rs.randint(64, 140)), 4 * int(rs.randint(64, 140))
        cases.append(("NCHW", (1, int(rs.randint(1, 3)), h, w), int(rs.randint(2, 200))))
    for _ in range(5):   # everything else: tile kernel
        cases.append(("NCHW", (2, 2, int(rs.randint(5, 120)), int(rs.randint(5, 200))), int(rs.randint(1, 40))))
    for layout, shape, radius in cases:
        taps = rs.rand(2 * radius + 1).astype(np.float32)
        taps /= taps.sum()
        x = (rs.rand(*shape) * 2 - 1).astype(np.float32) * np.float32(10.0 ** rs.randint(-3, 4))
        lay = ops.LAYOUT_NCHW if layout == "NCHW" else ops.LAYOUT_NHWC
        got = ops.blur2d_taps(_dev(x), _dev(taps), lay)
        seen.add(ops.last_blur_path())
        want = bo.blur_nchw(x, taps) if layout == "NCHW" else bo.blur_nhwc(x, taps)
        for n in range(shape[0]):
            assert_close(got[n], want[n], "%s %s radius %d path %d" % (layout, shape, radius, ops.last_blur_path()))
    # the tensor-core kernel of 128 x 128 planes generates its taps in-kernel (tap rules only): random sigmas, both rules
    from blurred_gan_b200.gaussian_blur import blur_images
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    for _ in range(6):
        std = float(rs.uniform(1.2, 30.0))
        x = (rs.rand(int(rs.randint(1, 5)), 3, 128, 128) * 2 - 1).astype(np.float32) * np.float32(10.0 ** rs.randint(-3, 4))
        got = GaussianBlur(std).cuda()(_dev(x))
        seen.add(ops.last_blur_path())
        assert_close(got, bo.gaussian_blur_torch_ref(x, std), "torch rule std=%g" % std)
        xh = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
        got = blur_images(_dev(xh), std)
        seen.add(ops.last_blur_path())
        assert_close(got, bo.blur_images_tf_ref(xh, std), "tf rule NHWC std=%g" % std)
    assert {4, 5, 6} <= seen and (seen & {1, 2, 3}), seen


# ---- several tensors per launch, interpolates between blurred batches -----------------------------------
@pytest.mark.parametrize("std,rule", [(0.5, "torch"), (1.0, "torch"), (3.0, "torch"), (23.0, "torch"), (0.5, "tf"), (5.0, "tf")])
@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
@pytest.mark.parametrize("count,batch", [(2, 5), (3, 50), (4, 1)])
def test_multi_tensor_launch_is_bitwise_the_single_launches(std, rule, layout, count, batch):
    """One launch over `count` tensors (units numbered tensor after tensor; 3 x 50 x 3 = 450 planes makes CTAs cross tensor
    boundaries) == `count` single launches, bit for bit (same kernels, same per-plane arithmetic)."""
    from blurred_gan_b200 import ops

    if rule == "torch" and layout == "NHWC":
        pytest.skip("the torch-rule module is NCHW only")
    shape = (batch, 3, 128, 128) if layout == "NCHW" else (batch, 128, 128, 3)
    xs = [_dev(uniform_pm1(10 + t, shape)) * float(3 ** t) for t in range(count)]
    lay = ops.LAYOUT_NCHW if layout == "NCHW" else ops.LAYOUT_NHWC
    if rule == "torch":
        _, r = bo.torch_rule_plan(std, 128, 128)
        rl, s = ops.RULE_TORCH, std
    else:
        _, r, s = bo.tf_rule_plan(std, 128, 128)
        rl = ops.RULE_TF
    singles = []
    for x in xs:
        singles.append(ops.blur2d(x, lay, rl, float(s), r))
        path_single = ops.last_blur_path()
    multi = ops.blur2d_multi(xs, lay, rl, float(s), r)
    assert ops.last_blur_path() == path_single and path_single in (4, 5)
    assert len(multi) == count
    for t in range(count):
        assert torch.equal(multi[t], singles[t]), "tensor %d differs" % t


def test_multi_tensor_other_shapes_modules_and_autograd():
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import GaussianBlur2D
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    # generic shapes: one launch sequence per tensor behind the same call
    xs = [_dev(uniform_pm1(20 + t, (2, 3, 37, 50))) for t in range(3)]
    mod = GaussianBlur(2.0).cuda()
    for got, x in zip(mod.forward_many(*xs), xs):
        assert torch.equal(got, mod(x))
    layer = GaussianBlur2D(4.0).cuda()
    big = [_dev(uniform_pm1(30 + t, (1, 3, 256, 256))) for t in range(2)]
    for got, x in zip(layer.forward_many(*big), big):
        assert torch.equal(got, layer(x))
    # autograd through the multi op: backward and double backward are the same op on the gradients
    a = _dev(uniform_pm1(40, (2, 3, 128, 128))).requires_grad_(True)
    b = _dev(uniform_pm1(41, (2, 3, 128, 128))).requires_grad_(True)
    ga, gb = _dev(uniform_pm1(42, (2, 3, 128, 128))), _dev(uniform_pm1(43, (2, 3, 128, 128)))
    ya, yb = mod.forward_many(a, b)
    da, db = torch.autograd.grad([ya, yb], [a, b], [ga, gb], create_graph=True)
    assert torch.equal(da, mod(ga)) and torch.equal(db, mod(gb))
    # the backward is differentiable again (double backward): d/dg <B g, v> = B v
    va = _dev(uniform_pm1(44, (2, 3, 128, 128)))
    gr = ga.clone().requires_grad_(True)
    ya2, _ = mod.forward_many(a, b)
    (d1,) = torch.autograd.grad(ya2, a, gr, create_graph=True)
    (d2,) = torch.autograd.grad(d1, gr, va)
    assert_close(d2, mod(va).cpu().numpy(), "double backward of the multi op")
    # errors: too many tensors, mismatched shapes, CPU tensors
    with pytest.raises(RuntimeError):
        ops.blur2d_multi([a.detach()] * 5, ops.LAYOUT_NCHW, ops.RULE_TORCH, 2.0, 5)
    with pytest.raises(RuntimeError):
        ops.blur2d_multi([a.detach(), a.detach()[:1]], ops.LAYOUT_NCHW, ops.RULE_TORCH, 2.0, 5)
    with pytest.raises(RuntimeError):
        ops.blur2d_multi([a.detach().cpu()], ops.LAYOUT_NCHW, ops.RULE_TORCH, 2.0, 5)


def test_multi_tensor_double_backward():
    """Gradient-penalty pattern through the multi op: d/dx of ||B-adjoint grad||^2 equals the single-op result."""
    from blurred_gan_b200 import ops

    x = _dev(uniform_pm1(50, (2, 3, 128, 128)))
    w = _dev(uniform_pm1(51, (2, 3, 128, 128)))

    def penalty(blur_fn):
        xa = x.clone().requires_grad_(True)
        y = blur_fn(xa)
        score = (y * y * w).sum()
        (g,) = torch.autograd.grad(score, xa, create_graph=True)
        (gg,) = torch.autograd.grad(g.square().sum(), xa)
        return gg

    single = penalty(lambda t: ops.blur2d(t, ops.LAYOUT_NCHW, ops.RULE_TORCH, 3.0, 7))
    multi = penalty(lambda t: ops.blur2d_multi([t], ops.LAYOUT_NCHW, ops.RULE_TORCH, 3.0, 7)[0])
    assert torch.equal(single, multi)


@pytest.mark.parametrize("shape", [(7, 3, 128, 128), (5, 1, 28, 28), (3, 2, 5, 9), (0, 3, 8, 8)])
def test_lerp_rows_is_the_reference_interpolation(shape):
    """wgan.py:237-239: x_hat = reals + a * (fakes - reals), one weight per image; same fp32 expression, bit-exact or 1 ulp (FMA contraction)."""
    from blurred_gan_b200 import ops

    if shape[0]:
        u, v = _dev(uniform_pm1(60, shape)), _dev(uniform_pm1(61, shape))
    else:
        u, v = torch.empty(shape, device="cuda"), torch.empty(shape, device="cuda")
    a = torch.rand(shape[0], device="cuda")
    got = ops.lerp_rows(u, v, a)
    want = u.double() + a.double().view(-1, 1, 1, 1) * (v.double() - u.double())
    assert got.shape == u.shape
    if shape[0]:
        assert (got.double() - want).abs().max().item() <= 2.5e-7
        # linearity of the blur: B(lerp(x, z)) == lerp(B x, B z) to rounding -- what the fused critic step relies on
        if shape[-1] >= 28:
            r = 3
            bl = lambda t: ops.blur2d(t, ops.LAYOUT_NCHW, ops.RULE_TORCH, 1.3, r)  # noqa: E731
            lhs = bl(got)
            rhs = ops.lerp_rows(bl(u), bl(v), a)
            assert (lhs - rhs).abs().max().item() <= 1e-6
