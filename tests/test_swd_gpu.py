"""GPU parity tests of the SWD kernels against the CPU oracle and the reference's golden vectors.
Integer / gather / sort work is bit-exact; the pyramid is bit-exact (fp64 accumulate, one rounding);
floating-point reductions are within 1e-6; the SWD distances within north_star's 1e-4."""
import os

import numpy as np
import pytest
import torch

from conftest import uniform_pm1
from oracle import swd_oracle as so

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [(5, 3, 64, 64), (2, 3, 128, 128), (1, 3, 16, 16), (2, 3, 32, 48)])
def test_pyramid_bit_exact(shape):
    from blurred_gan_b200 import sliced_wasserstein as sw

    x = uniform_pm1(3, shape)
    levels = 1
    while min(shape[2], shape[3]) // (2 ** levels) >= 8 and levels < 4:
        levels += 1
    xd = _dev(x)
    keep = xd.clone()
    got = sw.generate_laplacian_pyramid(xd, levels)
    assert torch.equal(xd, keep)  # inputs are not mutated
    want = so.generate_laplacian_pyramid(x, levels)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    np.testing.assert_array_equal(sw.reconstruct_laplacian_pyramid(got).cpu().numpy(), so.reconstruct_laplacian_pyramid(want))


def test_pyr_down_up_odd_sizes_and_golden(golden_swd):
    from blurred_gan_b200 import sliced_wasserstein as sw

    for shape in [(1, 3, 17, 31), (2, 1, 5, 6), (1, 2, 1, 9), (1, 1, 2, 2)]:
        x = uniform_pm1(4, shape)
        np.testing.assert_array_equal(sw.pyr_down(x).cpu().numpy(), so.pyr_down(x))
        np.testing.assert_array_equal(sw.pyr_up(x).cpu().numpy(), so.pyr_up(x))
    reals = uniform_pm1(11, (6, 3, 64, 64))
    np.testing.assert_array_equal(sw.pyr_down(reals).cpu().numpy(), golden_swd["s64_down"])
    np.testing.assert_array_equal(sw.pyr_up(golden_swd["s64_down"]).cpu().numpy(), golden_swd["s64_up"])
    for i, lvl in enumerate(sw.generate_laplacian_pyramid(reals, 3)):
        np.testing.assert_array_equal(lvl.cpu().numpy(), golden_swd["s64_pyr%d" % i])


def test_descriptor_gather_exact_and_statistics(golden_swd):
    from blurred_gan_b200 import ops
    from blurred_gan_b200 import sliced_wasserstein as sw

    level0 = golden_swd["s64_pyr0"]
    np.random.seed(7)
    desc = sw.get_descriptors_for_minibatch(level0, 7, 128)
    assert tuple(desc.shape) == (6 * 128, 3, 7, 7)
    np.testing.assert_array_equal(desc[:256].cpu().numpy(), golden_swd["s64_desc0"])
    want = so.gather_descriptors(level0, golden_swd["s64_px"], golden_swd["s64_py"], 7, 128)
    np.testing.assert_array_equal(desc.cpu().numpy(), want)
    sums = torch.zeros(6, dtype=torch.float64, device="cuda")
    ops.swd_extract(_dev(level0), golden_swd["s64_px"], golden_swd["s64_py"], 128, 7, sums)
    w64 = want.astype(np.float64)
    np.testing.assert_allclose(sums.cpu().numpy(), np.concatenate([w64.sum((0, 2, 3)), (w64 ** 2).sum((0, 2, 3))]), rtol=1e-12)
    fin = sw.finalize_descriptors(desc)
    assert tuple(fin.shape) == (768, 147)
    np.testing.assert_allclose(fin[:256].cpu().numpy(), golden_swd["s64_final0"], rtol=2e-6, atol=2e-6)
    fin2 = sw.finalize_descriptors([desc[:300], desc[300:]])
    assert torch.equal(fin, fin2)


@pytest.mark.parametrize("rows,k,d", [(1000, 147, 128), (129, 147, 512), (5, 7, 3), (4096, 147, 64), (300, 49, 130)])
def test_projection(rows, k, d):
    from blurred_gan_b200 import ops

    rs = np.random.RandomState(rows)
    A = rs.randn(rows, k).astype(np.float32)
    dirs = rs.randn(k, d).astype(np.float32)
    got = ops.swd_project(_dev(A), _dev(dirs)).cpu().numpy()
    want = (A.astype(np.float64) @ dirs.astype(np.float64)).T
    assert got.shape == (d, rows)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("rows,k,d", [(20000, 147, 512), (333, 192, 70), (1000, 16, 128), (257, 1, 1)])
def test_projection_tensor_core_path_any_magnitude(rows, k, d, monkeypatch):
    """The tcgen05 projection (csrc/project_tc.cu): rows of very different magnitudes (per-row power-of-two scales), zero
    rows, ragged last tile, direction counts that are not multiples of 128 -- against fp64 and against the FFMA kernel."""
    from blurred_gan_b200 import _lib, ops

    rs = np.random.RandomState(rows + k)
    A = rs.randn(rows, k).astype(np.float32)
    A[1::7] *= 1e-6
    A[2::7] *= 3e4
    A[3::11] = 0.0
    A[:, 0] *= 1e-3
    dirs = (rs.randn(k, d) * 0.07).astype(np.float32)
    want = (A.astype(np.float64) @ dirs.astype(np.float64)).T
    got = ops.swd_project(_dev(A), _dev(dirs)).cpu().numpy()
    assert _lib.load().bgan_swd_project_last_path() == 2
    monkeypatch.setenv("BGAN_SWD_NO_TC_PROJECT", "1")
    fma = ops.swd_project(_dev(A), _dev(dirs)).cpu().numpy()
    assert _lib.load().bgan_swd_project_last_path() == 1
    monkeypatch.delenv("BGAN_SWD_NO_TC_PROJECT")
    # per row (= per output column): error relative to that row's largest projection
    scale = np.abs(want).max(axis=0, keepdims=True) + 1e-30
    assert (np.abs(got - want) / scale).max() <= 1e-5
    assert (np.abs(fma - want) / scale).max() <= 1e-5
    assert np.all(got[:, 3::11] == 0.0)


def test_projection_long_descriptors_use_the_fma_kernel():
    from blurred_gan_b200 import _lib, ops

    rs = np.random.RandomState(5)
    A = rs.randn(300, 200).astype(np.float32)
    dirs = rs.randn(200, 40).astype(np.float32)
    got = ops.swd_project(_dev(A), _dev(dirs)).cpu().numpy()
    assert _lib.load().bgan_swd_project_last_path() == 1
    want = (A.astype(np.float64) @ dirs.astype(np.float64)).T
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("algo", ["default", "lsd", "fallback"])
@pytest.mark.parametrize("segments,rows", [(1, 1), (3, 5), (2, 4096), (5, 4097), (130, 1000), (3, 8192), (3, 8193), (4, 100000), (4, 100001),
                                           (1, 1 << 20), (2, 1 << 21), (1, (1 << 21) + 1)])
def test_segmented_sort_is_exact(segments, rows, algo, monkeypatch):
    """Three routes: the value-bucket sort (segments of up to 2^21 keys), the radix passes alone, and the
    radix passes entered through the sample sort's overflow flag."""
    from blurred_gan_b200 import _lib, ops

    if algo == "lsd":
        monkeypatch.setenv("BGAN_SORT_ALGO", "lsd")
    if algo == "fallback":
        monkeypatch.setenv("BGAN_SORT_FORCE_FALLBACK", "1")
    rs = np.random.RandomState(segments * 7919 + rows)
    keys = (rs.randn(segments, rows) * 10.0 ** rs.randint(-3, 4, size=(segments, 1))).astype(np.float32)
    if rows >= 5:
        keys[:, 0] = 0.0
        keys[:, 1] = -0.0
        keys[:, 2] = keys[:, 3]
        keys[0, 4] = np.inf
        keys[-1, 4] = -np.inf
    got = ops.swd_sort_(_dev(keys)).cpu().numpy()
    want = np.sort(keys, axis=1)
    np.testing.assert_array_equal(got, want)  # (-0.0 == 0.0 compares equal; order among them is irrelevant)
    path = _lib.load().bgan_swd_sort_last_path()
    if algo == "lsd" or rows > (1 << 21):
        assert path == 1
    elif rows <= 8192:
        assert path == 2  # one CTA per segment: nothing can overflow, nothing is read back
    elif algo == "fallback":
        assert path == 3
    else:
        assert path in (2, 3)  # (3: an infinity among the samples stretches the value range over everything else)


def test_sort_inside_a_cuda_graph_takes_the_radix_passes():
    """The value-bucket sort reads an overflow flag back and synchronises the stream; a stream that is being captured cannot be
    synchronised, so a captured sort of long segments must run the radix passes (and replay correctly)."""
    from blurred_gan_b200 import _lib, ops

    rs = np.random.RandomState(77)
    src = torch.from_numpy(rs.randn(3, 50000).astype(np.float32)).cuda()
    keys = src.clone()
    scratch = {}
    ops.swd_sort_(keys.clone(), scratch)  # workspace allocated outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.swd_sort_(keys, scratch)
    assert _lib.load().bgan_swd_sort_last_path() == 1
    for _ in range(2):
        keys.copy_(src)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(keys, torch.sort(src, dim=1).values)


def test_sort_fuzz_short():
    """Ten seconds of tools/sort_fuzz.py: random shapes (ragged blocks, runs at the 8192-key boundary, odd row counts) and
    distributions through the sort and the fused sort + distance, against torch.sort and fp64 sums."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "tools", "sort_fuzz.py"), "10", "3"], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and proc.stdout.startswith("ok:"), proc.stdout + proc.stderr


def _sort_case(name, rs, segments, rows):
    if name == "all_equal":
        return np.full((segments, rows), 0.25, np.float32)
    if name == "ten_values":  # every value far too frequent to sort: equality buckets
        return rs.randint(0, 10, size=(segments, rows)).astype(np.float32) - 4.5
    if name == "ascending":
        return np.sort(rs.randn(segments, rows).astype(np.float32), axis=1)
    if name == "descending":
        return -np.sort(rs.randn(segments, rows).astype(np.float32), axis=1)
    if name == "half_zero":  # one heavy value in a continuous distribution
        k = rs.randn(segments, rows).astype(np.float32)
        k[rs.rand(segments, rows) < 0.5] = 0.0
        return k
    if name == "outliers":  # a tight cluster and a few values 36 orders of magnitude away
        k = (rs.rand(segments, rows) * 1e-6).astype(np.float32)
        k[:, ::1000] = 1e30
        k[:, 1::1000] = -1e30
        return k
    if name == "clusters":  # near-duplicates at many scales: sub-buckets overload, the bitonic route sorts them
        k = (10.0 ** rs.randint(-30, 30, size=(segments, rows))).astype(np.float32)
        return k * (1 + 1e-6 * rs.rand(segments, rows)).astype(np.float32) * rs.choice([-1.0, 1.0], size=(segments, rows)).astype(np.float32)
    if name == "lognormal":
        return np.exp(rs.randn(segments, rows) * 8).astype(np.float32)
    if name == "nonfinite":
        k = rs.randn(segments, rows).astype(np.float32)
        k[:, ::7] = np.inf
        k[:, 3::11] = -np.inf
        return k
    if name == "nans":  # positive NaNs sort last, as in np.sort
        k = rs.randn(segments, rows).astype(np.float32)
        k[:, ::13] = np.nan
        return k
    if name == "denormals":
        return (rs.randn(segments, rows) * 1e-41).astype(np.float32)
    raise KeyError(name)


@pytest.mark.parametrize("rows", [5000, 70001, 300000])
@pytest.mark.parametrize("name", ["all_equal", "ten_values", "ascending", "descending", "half_zero", "outliers", "clusters", "lognormal",
                                  "nonfinite", "nans", "denormals"])
def test_segmented_sort_is_exact_whatever_the_distribution(name, rows):
    """The bucket boundaries come from a sample of the keys and the sub-buckets from a linear map of the value: neither may
    cost correctness when the keys are nothing like a smooth density."""
    from blurred_gan_b200 import _lib, ops

    rs = np.random.RandomState(rows + len(name))
    keys = _sort_case(name, rs, 3, rows)
    got = ops.swd_sort_(_dev(keys)).cpu().numpy()
    np.testing.assert_array_equal(got, np.sort(keys, axis=1))
    path = _lib.load().bgan_swd_sort_last_path()
    if rows > 8192 and name in ("ascending", "descending"):
        assert path == 2  # smooth densities stay on the value-bucket route
    if rows > 8192 and name in ("all_equal", "ten_values", "half_zero"):
        assert path == 2  # heavy values get buckets of their own that are never sorted: no overflow, no radix passes


def test_absdiff_sum():
    from blurred_gan_b200 import ops

    a, b = uniform_pm1(1, (128, 3001)), uniform_pm1(2, (128, 3001))
    got = ops.swd_absdiff_sum(_dev(a), _dev(b)).item()
    want = np.abs(a.astype(np.float64) - b.astype(np.float64)).sum()
    # |a-b| is formed in fp32 like the reference's np.abs(projA - projB), then summed in fp64
    assert got == pytest.approx(np.abs(a - b).astype(np.float64).sum(), rel=1e-12)
    assert got == pytest.approx(want, rel=1e-6)


def test_sliced_wasserstein_known_answer(golden_swd):
    from blurred_gan_b200 import sliced_wasserstein as sw

    rs = np.random.RandomState(5)
    A = rs.randn(1000, 147).astype(np.float32)
    B = (rs.randn(1000, 147) * 1.2 + 0.1).astype(np.float32)
    np.random.seed(9)
    got = sw.sliced_wasserstein(A, B, 4, 128)
    assert got == pytest.approx(golden_swd["sw_AB"][0], rel=1e-4)
    np.random.seed(1)
    assert sw.sliced_wasserstein(A, A, 2, 16) == 0.0


def _run_api(api, reals, fakes, mb, as_tensor=False):
    conv = (lambda a: torch.from_numpy(a).cuda()) if as_tensor else (lambda a: a.copy())
    api.begin("reals")
    for i in range(0, len(reals), mb):
        api.feed("reals", conv(reals[i : i + mb]))
    r0 = api.end("reals")
    api.begin("fakes")
    for i in range(0, len(fakes), mb):
        api.feed("fakes", conv(fakes[i : i + mb]))
    return np.array(r0, np.float64), np.array(api.end("fakes"), np.float64)


@pytest.mark.parametrize("as_tensor", [False, True])
def test_api_matches_reference_golden(golden_swd, as_tensor):
    """Identical seeds -> identical patches and projections; distances within 1e-4 (north_star)."""
    from blurred_gan_b200 import sliced_wasserstein as sw

    reals = uniform_pm1(11, (6, 3, 64, 64))
    fakes = (0.8 * uniform_pm1(12, (6, 3, 64, 64)) + 0.1).astype(np.float32)
    np.random.seed(0)
    r0, r1 = _run_api(sw.API((3, 64, 64)), reals, fakes, 3, as_tensor)
    np.testing.assert_array_equal(r0, golden_swd["s64_api_reals"])
    np.testing.assert_allclose(r1, golden_swd["s64_api_fakes"], rtol=1e-4)

    reals = uniform_pm1(21, (16, 3, 128, 128))
    fakes = np.clip(uniform_pm1(22, (16, 3, 128, 128)) * 1.1 - 0.05, -1, 1).astype(np.float32)
    np.random.seed(0)
    r0, r1 = _run_api(sw.API((3, 128, 128)), reals, fakes, 8, as_tensor)
    np.testing.assert_array_equal(r0, golden_swd["s128_api_reals"])
    np.testing.assert_allclose(r1, golden_swd["s128_api_fakes"], rtol=1e-4)
    assert len(r1) == 5 and r1[4] == pytest.approx(np.mean(r1[:4]), rel=1e-6)


def test_api_matches_oracle_larger():
    from blurred_gan_b200 import sliced_wasserstein as sw

    reals = uniform_pm1(31, (96, 3, 128, 128))
    fakes = (0.9 * uniform_pm1(32, (96, 3, 128, 128))).astype(np.float32)
    fakes[:, :, ::2] *= 0.8
    np.random.seed(3)
    want = _run_api(so.API((3, 128, 128)), reals, fakes, 32)
    np.random.seed(3)
    got = _run_api(sw.API((3, 128, 128)), reals, fakes, 32)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_allclose(got[1], want[1], rtol=1e-4)


@pytest.mark.parametrize("rows,segments,seg_offset,per_sum,n_sums", [(5000, 6, 0, 3, 2), (4096, 4, 126, 128, 2), (70001, 3, 5, 2, 4), (1, 2, 0, 1, 2),
                                                                   (262144, 8, 0, 4, 2)])
def test_fused_sort_diff_is_the_sort_then_the_distance(rows, segments, seg_offset, per_sum, n_sums):
    """bgan_swd_sort_diff_f32 (sliced_wasserstein.py:47-50 in one call): a comes back sorted bit for bit like np.sort, the
    sums are sum |sort(a) - sort(b)| per group of segments (fp64 of the fp32 differences; accumulated INTO sums)."""
    from blurred_gan_b200 import ops

    rs = np.random.RandomState(rows % 1000 + segments)
    a = (rs.randn(segments, rows) * np.exp(rs.randn(segments, 1))).astype(np.float32)
    b = (rs.randn(segments, rows) * 1.3 + 0.2).astype(np.float32)
    a[0, : min(rows, 7)] = [0.0, -0.0, np.float32(1e-40), -np.float32(1e-40), 3e38, -3e38, 1.0][: min(rows, 7)]  # signed zeros, subnormals, extremes
    sums = torch.full((n_sums,), 0.5, dtype=torch.float64, device="cuda")
    ad, bd = _dev(a), _dev(b)
    ops.swd_sort_diff_(ad, bd, sums, seg_offset, per_sum)
    sa, sb = np.sort(a, axis=1), np.sort(b, axis=1)
    got_a = ad.cpu().numpy()
    np.testing.assert_array_equal(got_a, sa)
    want = np.full(n_sums, 0.5)
    for s in range(segments):
        want[(s + seg_offset) // per_sum] += np.abs(sa[s] - sb[s]).astype(np.float64).sum()
    np.testing.assert_allclose(sums.cpu().numpy(), want, rtol=1e-12)
    # the same through the two separate calls
    ad2, bd2 = ops.swd_sort_(_dev(a)), ops.swd_sort_(_dev(b))
    assert torch.equal(ad2, ad)
    for s in range(segments):
        sep = float(ops.swd_absdiff_sum(ad2[s], bd2[s]).item())
        assert sep == pytest.approx(np.abs(sa[s] - sb[s]).astype(np.float64).sum(), rel=1e-12)


def test_api_matches_reference_golden_at_baseline_scale():
    """2 x 2048 images (262,144 descriptors per level and set: 64 sort tiles per direction, 2048 row tiles of the tensor-core
    projection) against the UNMODIFIED reference, run once in the build container (tests/golden/make_golden.py --swd-large;
    fixture: 2 x 5 floats).  Same seeds, same minibatches, same numpy stream; tolerance north_star's 1e-4."""
    import os

    from blurred_gan_b200 import sliced_wasserstein as sw
    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "swd_large.npz"))
    n, mb, seed_r, seed_f, np_seed = (int(v) for v in g["config"])
    reals = uniform_pm1(seed_r, (n, 3, 128, 128))
    fakes = (0.9 * uniform_pm1(seed_f, (n, 3, 128, 128))).astype(np.float32)
    fakes[:, :, ::2] *= 0.8
    np.random.seed(np_seed)
    r0, r1 = _run_api(sw.API((3, 128, 128)), reals, fakes, mb, True)
    np.testing.assert_array_equal(r0, g["api_reals"])
    np.testing.assert_allclose(r1, g["api_fakes"], rtol=1e-4)


def test_downscale_minibatch_values():
    """sliced_wasserstein.py:55-61 against the reference's golden (uint8 input) and the oracle (float input, ties)."""
    import os

    from blurred_gan_b200 import sliced_wasserstein as sw
    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "swd_downscale.npz"))
    x = g["x"]
    xd = torch.from_numpy(x).cuda()
    assert sw.downscale_minibatch(xd, 0) is xd
    for lod in (1, 2, 3):
        got = sw.downscale_minibatch(xd, lod)
        assert got.dtype == torch.uint8 and got.is_cuda
        np.testing.assert_array_equal(got.cpu().numpy(), g["lod%d" % lod])
        np.testing.assert_array_equal(sw.downscale_minibatch(x, lod).cpu().numpy(), g["lod%d" % lod])  # numpy in, like the reference
    # float input with exact .5 ties and values outside [0, 255]: round-half-to-even and the clip
    xf = (np.random.RandomState(72).randint(-40, 600, size=(2, 3, 16, 16)) * 0.5).astype(np.float32)
    for lod in (1, 2):
        np.testing.assert_array_equal(sw.downscale_minibatch(_dev(xf), lod).cpu().numpy(), so.downscale_minibatch(xf, lod))


def _dist_worker(rank, world, port, out_dir):
    import os
    import sys

    import torch.distributed as dist

    from conftest import ROOT

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # two ranks share cuda:0; collectives staged through the host
    from blurred_gan_b200 import sliced_wasserstein as sw

    reals = uniform_pm1(51, (16, 3, 64, 64))
    fakes = (0.85 * uniform_pm1(52, (16, 3, 64, 64))).astype(np.float32)
    mb = 4
    np.random.seed(77)
    api = sw.API((3, 64, 64), distributed=True)
    assert (api.rank, api.world) == (rank, world)
    api.begin("reals")
    for k in range(len(reals) // (mb * world)):
        api.feed("reals", reals[(k * world + rank) * mb : (k * world + rank + 1) * mb])
    r0 = api.end("reals")
    api.begin("fakes")
    for k in range(len(fakes) // (mb * world)):
        api.feed("fakes", torch.from_numpy(fakes[(k * world + rank) * mb : (k * world + rank + 1) * mb]).cuda())
    r1 = api.end("fakes")
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([r0, r1], dtype=np.float64))
    dist.destroy_process_group()


def test_distributed_api_two_ranks_match_single_process(tmp_path):
    """sharded feed, all-reduced statistics, all-gathered descriptors, direction blocks per rank."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_dist_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(2)]
    np.testing.assert_array_equal(got[0], got[1])
    reals = uniform_pm1(51, (16, 3, 64, 64))
    fakes = (0.85 * uniform_pm1(52, (16, 3, 64, 64))).astype(np.float32)
    np.random.seed(77)
    want = _run_api(so.API((3, 64, 64)), reals, fakes, 4)  # single process, same global minibatch order
    np.testing.assert_array_equal(got[0][0], want[0])
    np.testing.assert_allclose(got[0][1], want[1], rtol=1e-4)


@pytest.mark.parametrize("shape", [(6, 3, 128, 128), (3, 3, 64, 96), (2, 3, 16, 16), (4, 3, 20, 28)])
def test_fused_feed_is_bitwise_the_level_by_level_feed(shape):
    """bgan_swd_feed_f32 (pyramid + gather of all levels in one call) == the level-by-level calls: descriptors and channel
    statistics bit for bit, same position stream.  20 x 28 has one level only."""
    from blurred_gan_b200 import sliced_wasserstein as sw

    x = _dev(uniform_pm1(77, shape))
    results = []
    for fused in (True, False):
        np.random.seed(5)
        api = sw.API(shape[1:])
        api.begin("reals")
        if not fused:
            # the reference's own call sequence (sliced_wasserstein.py:121-123): pyramid, then per level x and y draws + gather
            levels = sw.generate_laplacian_pyramid(x, len(api.resolutions))
            sums = [torch.zeros(7, dtype=torch.float64, device="cuda") for _ in levels]
            descs = [sw.get_descriptors_for_minibatch(lv, 7, 128, s[:6]) for lv, s in zip(levels, sums)]
        else:
            api.feed("reals", x)
            descs = [slab[: api._rows[lod]] for lod, slab in enumerate(api._slabs)]
            assert all(not chunks for chunks in api.descriptors)
            sums = api._sums
        results.append(([d.clone() for d in descs], [s[:6].clone() for s in sums], np.random.get_state()[2]))
    (d0, s0, p0), (d1, s1, p1) = results
    assert p0 == p1, "the two paths consumed different numbers of draws"
    assert len(d0) == len(d1)
    for a, b in zip(d0, d1):
        assert a.shape == b.shape and torch.equal(a, b)
    for a, b in zip(s0, s1):
        # fp64 atomics: the accumulation order inside extract_kernel is not deterministic
        assert torch.allclose(a, b, rtol=1e-12, atol=1e-9)
    assert torch.equal(x, _dev(uniform_pm1(77, shape))), "feed must not modify its input"
