"""CPU: pins oracle/swd_oracle.py against golden vectors from the unmodified reference module."""
import sys

import numpy as np

from conftest import needs_reference, uniform_pm1
from oracle import swd_oracle as so


def _s64_inputs():
    reals = uniform_pm1(11, (6, 3, 64, 64))
    fakes = (0.8 * uniform_pm1(12, (6, 3, 64, 64)) + 0.1).astype(np.float32)
    return reals, fakes


def _run_api(api, reals, fakes, mb):
    api.begin("reals")
    for i in range(0, len(reals), mb):
        api.feed("reals", reals[i : i + mb].copy())
    r0 = api.end("reals")
    api.begin("fakes")
    for i in range(0, len(fakes), mb):
        api.feed("fakes", fakes[i : i + mb].copy())
    return np.array(r0, np.float64), np.array(api.end("fakes"), np.float64)


def test_pyramid_is_bit_exact(golden_swd):
    reals, _ = _s64_inputs()
    keep = reals.copy()
    pyr = so.generate_laplacian_pyramid(reals, 3)
    np.testing.assert_array_equal(reals, keep)  # the oracle must not mutate its input
    for i, lvl in enumerate(pyr):
        np.testing.assert_array_equal(lvl, golden_swd["s64_pyr%d" % i])
    np.testing.assert_array_equal(so.pyr_down(reals), golden_swd["s64_down"])
    np.testing.assert_array_equal(so.pyr_up(golden_swd["s64_down"]), golden_swd["s64_up"])
    np.testing.assert_array_equal(so.reconstruct_laplacian_pyramid(pyr), so.pyr_up(so.pyr_up(pyr[2]) + pyr[1]) + pyr[0])


def test_descriptors_and_rng_order(golden_swd):
    level0 = golden_swd["s64_pyr0"]
    rng = np.random.RandomState(7)
    px, py = so.draw_patch_positions(6, 64, 64, 7, 128, rng)
    np.testing.assert_array_equal(px, golden_swd["s64_px"])
    np.testing.assert_array_equal(py, golden_swd["s64_py"])
    desc = so.gather_descriptors(level0, px, py, 7, 128)
    np.testing.assert_array_equal(desc[:256], golden_swd["s64_desc0"])
    # transposed patch layout: axis 2 walks x, axis 3 walks y
    m, c, i, j = 5, 1, 2, 6
    assert desc[m, c, i, j] == level0[m // 128, c, py[m] + j - 3, px[m] + i - 3]
    np.testing.assert_allclose(so.finalize_descriptors(desc)[:256], golden_swd["s64_final0"], rtol=0, atol=0)


def test_api_matches_reference_small(golden_swd):
    reals, fakes = _s64_inputs()
    np.random.seed(0)
    r0, r1 = _run_api(so.API((3, 64, 64)), reals, fakes, 3)
    np.testing.assert_array_equal(r0, golden_swd["s64_api_reals"])
    np.testing.assert_allclose(r1, golden_swd["s64_api_fakes"], rtol=1e-6)


def test_api_matches_reference_celeba_shape(golden_swd):
    reals = uniform_pm1(21, (16, 3, 128, 128))
    fakes = np.clip(uniform_pm1(22, (16, 3, 128, 128)) * 1.1 - 0.05, -1, 1).astype(np.float32)
    np.random.seed(0)
    api = so.API((3, 128, 128))
    assert api.resolutions == [128, 64, 32, 16]
    assert api.get_metric_names() == ["SWDx1e3_128", "SWDx1e3_64", "SWDx1e3_32", "SWDx1e3_16", "SWDx1e3_avg"]
    r0, r1 = _run_api(api, reals, fakes, 8)
    np.testing.assert_array_equal(r0, golden_swd["s128_api_reals"])
    np.testing.assert_allclose(r1, golden_swd["s128_api_fakes"], rtol=1e-6)
    pyr = so.generate_laplacian_pyramid(reals[:2], 4)
    np.testing.assert_array_equal(pyr[3], golden_swd["s128_pyr3"])
    for i, lvl in enumerate(pyr):
        ref = golden_swd["s128_pyr%d_sum" % i]
        assert lvl.astype(np.float64).sum() == ref[0] and np.abs(lvl.astype(np.float64)).sum() == ref[1]


def test_sliced_wasserstein_known_answer(golden_swd):
    rs = np.random.RandomState(5)
    A = rs.randn(1000, 147).astype(np.float32)
    B = (rs.randn(1000, 147) * 1.2 + 0.1).astype(np.float32)
    got = so.sliced_wasserstein(A, B, 4, 128, np.random.RandomState(9))
    np.testing.assert_allclose(got, golden_swd["sw_AB"][0], rtol=1e-6)
    assert so.sliced_wasserstein(A, A, 2, 16, np.random.RandomState(1)) == 0.0


@needs_reference
def test_live_reference_agrees_with_oracle():
    sys.path.insert(0, "/root/reference")
    import sliced_wasserstein as ref

    x = uniform_pm1(5, (3, 3, 32, 48))
    for a, b in zip(ref.generate_laplacian_pyramid(x.copy(), 2), so.generate_laplacian_pyramid(x, 2)):
        np.testing.assert_array_equal(a, b)
    np.random.seed(4)
    d_ref = ref.get_descriptors_for_minibatch(x, 7, 16)
    d = so.get_descriptors_for_minibatch(x, 7, 16, np.random.RandomState(4))
    np.testing.assert_array_equal(d_ref, d)


def test_downscale_minibatch_matches_reference_golden():
    """sliced_wasserstein.py:55-61 (module surface; `API` does not call it): box filter 2^lod, round half to even, uint8."""
    import os

    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "swd_downscale.npz"))
    x = g["x"]
    assert so.downscale_minibatch(x, 0) is x
    for lod in (1, 2, 3):
        got = so.downscale_minibatch(x, lod)
        assert got.dtype == np.uint8 and got.shape == (3, 3, 32 >> lod, 48 >> lod)
        np.testing.assert_array_equal(got, g["lod%d" % lod])
