"""CPU: the drop-in modules expose the reference's names and host-side semantics; the compute path
refuses CPU tensors (there is no fallback)."""
import numpy as np
import pytest
import torch

from blurred_gan_b200 import gaussian_blur as tfblur
from blurred_gan_b200 import sliced_wasserstein as swd
from blurred_gan_b200.pytorch import gaussianblur as tblur
from oracle import blur_oracle as bo


def test_torch_module_surface_and_state():
    m = tblur.GaussianBlur(2.0)
    assert isinstance(m.std, torch.nn.Parameter) and tuple(m.std.shape) == (1,) and not m.std.requires_grad
    assert list(m.state_dict().keys()) == ["std"]
    assert m._image_size is None
    # before any forward the size is unclipped (reference pytorch/gaussianblur.py:49-54), e.g. 119 for 23.5
    assert tblur.GaussianBlur(23.5).kernel_size == 119
    m2 = tblur.GaussianBlur(1.0)
    m2.load_state_dict(m.state_dict())
    assert m2.std_value() == 2.0 and m2.kernel_size == 11
    m2.set_std(5.0)
    assert m2.kernel_size == 25 and float(m2.std) == 5.0
    with torch.no_grad():
        m2.std.fill_(0.5)  # a controller writing the Parameter directly is picked up through its version counter
    assert m2.std_value() == 0.5 and m2.kernel_size == 3


def test_torch_module_kernel_matches_reference(golden_taps):
    for std, h, w, k in golden_taps["rows"]:
        m = tblur.GaussianBlur(float(std))
        m._image_size = min(int(h), int(w))
        assert m.kernel_size == int(k)
        got = m.kernel
        assert tuple(got.shape) == (1, int(k))
        np.testing.assert_allclose(got.numpy().ravel(), golden_taps["taps_%g_%d_%d" % (std, h, w)], rtol=3e-7, atol=1e-9)


def test_torch_helpers():
    assert tblur.nearest_odd_integer(4.2) == 5 and tblur.nearest_odd_integer(6) == 7
    assert tblur.nearest_odd_integer(128, higher=False) == 127 and tblur.nearest_odd_integer(27, higher=False) == 25
    assert tblur.appropriate_kernel_size(0.01) == 3 and tblur.appropriate_kernel_size(23) == 115
    k = tblur.gaussian_kernel_1d(2.0, 11)
    assert tuple(k.shape) == (1, 11) and float(k.sum()) == pytest.approx(1.0, abs=1e-6)


def test_cpu_tensors_are_refused():
    x = torch.zeros(1, 1, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU"):
        tblur.GaussianBlur(1.0)(x)
    with pytest.raises(RuntimeError, match="no CPU"):
        tfblur.blur_images(x, 1.0)
    with pytest.raises(RuntimeError):
        tblur.GaussianBlur(1.0)(torch.zeros(8, 8))


def test_tf_helpers_match_oracle():
    assert tfblur.get_data_format(torch.zeros(2, 8, 8, 3)) == "NHWC"
    assert tfblur.get_data_format(torch.zeros(2, 8, 8, 1)) == "NHWC"
    assert tfblur.get_data_format(torch.zeros(2, 3, 8, 8)) == "NCHW"
    assert tfblur.get_image_dims(torch.zeros(2, 10, 12, 3)) == (10, 12, 3)
    assert tfblur.get_image_dims(torch.zeros(2, 3, 10, 12)) == (10, 12, 3)
    assert tfblur.appropriate_kernel_size(0.5) == 4.0 and tfblur.appropriate_kernel_size(23) == 139.0
    assert tfblur.appropriate_std(31) == pytest.approx(5.0)
    assert tfblur.maximum_reasonable_std(128) == pytest.approx(21.0)
    for std, k in [(0.5, 4), (2.0, 13), (21.1666667, 128)]:
        np.testing.assert_allclose(tfblur.gaussian_kernel_1d(std, k).numpy(), bo.tf_rule_taps(std, k), rtol=3e-7, atol=1e-9)
    layer = tfblur.GaussianBlur2D(initial_std=0.05)
    assert list(layer.state_dict().keys()) == ["std"] and layer.std_value() == pytest.approx(0.05)
    layer.assign(3.0)
    assert layer.std_value() == 3.0


def test_swd_surface():
    for name in ["get_descriptors_for_minibatch", "finalize_descriptors", "sliced_wasserstein", "downscale_minibatch",
                 "gaussian_filter", "pyr_down", "pyr_up", "generate_laplacian_pyramid", "reconstruct_laplacian_pyramid", "API"]:
        assert hasattr(swd, name), name
    api = swd.API((3, 128, 128))
    assert (api.nhood_size, api.nhoods_per_image, api.dir_repeats, api.dirs_per_repeat) == (7, 128, 4, 128)
    assert api.resolutions == [128, 64, 32, 16]
    assert api.get_metric_names() == ["SWDx1e3_128", "SWDx1e3_64", "SWDx1e3_32", "SWDx1e3_16", "SWDx1e3_avg"]
    assert api.get_metric_formatting() == ["%-13.4f"] * 5
    assert swd.API((3, 28, 28)).resolutions == [28]
    with pytest.raises(AssertionError):
        api.begin("bogus")
    np.testing.assert_array_equal(swd.gaussian_filter * 256, np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1]))
