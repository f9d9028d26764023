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
        m2.std.fill_(0.5)  # a controller writing the Parameter directly
    assert m2.std_value() == 0.5 and m2.kernel_size == 3
    # writes through `.data` do not bump the Parameter's version counter: they must be seen all the same
    v = m2.std._version
    m2.std.data.fill_(3.0)
    assert m2.std._version == v and m2.std_value() == 3.0 and m2.kernel_size == 15
    m2.std.data.copy_(torch.tensor([8.0]))
    assert m2.std_value() == 8.0 and m2.kernel_size == 41
    m2.assign(2.0)  # the reference controllers' name (callbacks.py:60-62)
    assert m2.std_value() == 2.0
    # std stays in host memory under module moves / casts (read every forward without a device synchronisation)
    m3 = m2.to(torch.float64)
    assert m3.std.dtype == torch.float32 and m3.std.device.type == "cpu" and list(m3.state_dict().keys()) == ["std"]


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
    layer.std.data.fill_(1.5)  # `.data` writes (no version bump) are seen too
    assert layer.std_value() == 1.5
    layer.std.data.copy_(torch.tensor(0.25))
    assert layer.std_value() == 0.25
    assert layer.to(torch.float64).std.dtype == torch.float32


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


def test_api_feed_bookkeeping_with_the_library_stubbed_out(monkeypatch):
    """Host logic of ``API.feed`` / ``end`` on the fused path -- the per-level arrays that grow geometrically and are filled
    in place, the row accounting, the position stream -- with the library calls replaced by CPU stubs (the arithmetic itself
    is covered by the GPU parity tests).  Ragged last minibatch, capacity doublings, slack trimming at ``end``."""
    from blurred_gan_b200 import dist as bd
    from blurred_gan_b200 import ops

    calls = []

    def fake_feed(minibatch, positions, levels, per_image, nhood, sums, scratch=None, outs=None):
        n = minibatch.shape[0]
        assert outs is not None and len(outs) == levels and positions.numel() == 2 * levels * n * per_image
        for lod, o in enumerate(outs):
            assert o.shape == (n * per_image, 3, nhood, nhood) and o.is_contiguous()
            o.fill_(float(len(calls) * 10 + lod))  # minibatch index and level, recoverable from the array
            sums[lod][:6] += 1.0
        calls.append(n)
        return outs

    class Ring:
        def take(self, n):
            return 0, torch.empty(n, dtype=torch.int32)

        def upload(self, k, host, device):
            return host

    monkeypatch.setattr(swd, "_to_device", lambda a, device=None: torch.as_tensor(a, dtype=torch.float32))
    monkeypatch.setattr(swd, "_PinnedRing", Ring)
    monkeypatch.setattr(ops, "swd_feed", fake_feed)
    monkeypatch.setattr(ops, "swd_normalise_", lambda desc, sums, count: desc)
    api = swd.API((3, 64, 64))
    assert api.resolutions == [64, 32, 16]
    np.random.seed(3)
    state0 = np.random.get_state()[1].copy()
    api.begin("reals")
    sizes = [2, 2, 2, 2, 2, 2, 2, 2, 2, 1]  # 19 images: capacity 4 -> 8 -> 16 minibatches of 2, ragged tail
    for n in sizes:
        api.feed("reals", np.zeros((n, 3, 64, 64), np.float32))
    assert calls == sizes
    rows = sum(sizes) * 128
    assert api._rows == [rows] * 3 and all(s.shape[0] >= rows for s in api._slabs)
    assert all(not c for c in api.descriptors)
    for lod, slab in enumerate(api._slabs):
        got = slab[:rows, 0, 0, 0].reshape(-1).numpy()
        want = np.concatenate([np.full(n * 128, i * 10 + lod, np.float32) for i, n in enumerate(sizes)])
        assert np.array_equal(got, want)  # every minibatch landed behind the previous one, growth copied the old rows
    assert not np.array_equal(np.random.get_state()[1], state0)  # positions came from numpy's global stream
    # reference stream: 10 feeds x 3 levels x (x, y) randint draws
    np.random.seed(3)
    for n in sizes:
        for size in (64, 32, 16):
            np.random.randint(3, size - 3, size=(n * 128, 1, 1, 1))
            np.random.randint(3, size - 3, size=(n * 128, 1, 1, 1))
    expect_after = np.random.get_state()
    np.random.seed(3)
    bd_state = np.random.get_state()
    np.random.set_state(bd_state)
    api2 = swd.API((3, 64, 64))
    api2.begin("reals")
    calls.clear()
    for n in sizes:
        api2.feed("reals", np.zeros((n, 3, 64, 64), np.float32))
    got_after = np.random.get_state()
    assert np.array_equal(got_after[1], expect_after[1]) and got_after[2] == expect_after[2]
    monkeypatch.setattr(bd, "all_reduce_sum_", lambda t, group=None: t)
    res = api2.end("reals")
    assert res == [0.0, 0.0, 0.0, 0.0] and api2._slabs is None
    assert all(d.shape == (rows, 147) for d in api2.desc_real)
    # 19 of 32 minibatch slots used: the kept set was trimmed to its own storage
    assert all(d.untyped_storage().nbytes() == rows * 147 * 4 for d in api2.desc_real)


def test_multi_tensor_entry_points_refuse_cpu_tensors_and_bad_lists():
    """forward_many / blur_images_many / lerp_rows keep the no-CPU-fallback rule of the single-tensor calls."""
    from blurred_gan_b200 import ops

    x = torch.zeros(2, 3, 16, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        tblur.GaussianBlur(1.0).forward_many(x, x.clone())
    with pytest.raises(RuntimeError, match="CUDA"):
        tfblur.GaussianBlur2D(1.0).forward_many(x, x.clone())
    with pytest.raises(RuntimeError, match="CUDA"):
        tfblur.blur_images_many([x], 1.0)
    with pytest.raises(RuntimeError, match="4-D"):
        tfblur.blur_images_many([x[0]], 1.0)
    with pytest.raises(RuntimeError):
        ops.lerp_rows(x, x, torch.zeros(2))
    assert ops.MAX_MULTI == 4
