"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (blurred_gan_b200/dist.py).

The compute here is the numpy oracle (allowed in tests only); what is under test is the exchange protocol: sharding,
the replayed random stream, the all-reduce of channel statistics, the all-gather of rows, the direction blocks and
the combination of per-repeat partial sums -- against a single-process oracle run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, uniform_pm1

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from blurred_gan_b200 import dist as bd
    from oracle import swd_oracle as so

    assert bd.world_info() == (rank, world)
    # ---- all_gather_rows with different row counts --------------------------------------------------------
    t = torch.full((3 + rank, 2), float(rank))
    g = bd.all_gather_rows(t)
    assert g.shape == (3 + 4, 2) and torch.equal(g[:3], torch.zeros(3, 2)) and torch.equal(g[3:], torch.ones(4, 2))
    # the asynchronous form: CPU / gloo tensors are gathered at once, equal or not; the handle returns the same tensor every time
    for rows in (3 + rank, 5):
        h = bd.all_gather_rows_async(torch.full((rows, 2), float(rank)), equal_rows=(rows == 5))
        got = h.result()
        assert got is h.result() and torch.equal(got, bd.all_gather_rows(torch.full((rows, 2), float(rank))))
    # ---- lock-step check: equal feed logs pass, a rank that fed another minibatch size is caught on every rank ----
    bd.check_lockstep([3, 12345])
    try:
        bd.check_lockstep([3, 12345 + rank])
        raise SystemExit("check_lockstep accepted ranks with different feed logs")
    except RuntimeError as exc:
        assert "different numbers or sizes" in str(exc)
    # ---- distributed SWD protocol, oracle compute ------------------------------------------------------------
    reals = uniform_pm1(41, (8, 3, 32, 32))
    fakes = (0.7 * uniform_pm1(42, (8, 3, 32, 32)) + 0.05).astype(np.float32)
    mb, levels, per_image, nhood, repeats, dpr = 2, 2, 16, 7, 4, 8
    np.random.seed(123)

    def collect(images):
        # round k: rank r owns minibatch k * world + r
        chunks = [[] for _ in range(levels)]
        for k in range(len(images) // (mb * world)):
            mine = images[(k * world + rank) * mb : (k * world + rank + 1) * mb]
            pyr = so.generate_laplacian_pyramid(mine, levels)
            pos = bd.positions_for_round(mb, [(lv.shape[2], lv.shape[3]) for lv in pyr], nhood, per_image, rank, world)
            for lod, lv in enumerate(pyr):
                chunks[lod].append(so.gather_descriptors(lv, pos[lod][0], pos[lod][1], nhood, per_image))
        out = []
        for lod in range(levels):
            d = np.concatenate(chunks[lod]).astype(np.float64)
            sums = torch.tensor(np.concatenate([d.sum((0, 2, 3)), (d ** 2).sum((0, 2, 3)), [d.shape[0]]]))
            bd.all_reduce_sum_(sums)
            cnt = sums[-1].item() * nhood * nhood
            mean = (sums[:3] / cnt).numpy()
            std = np.sqrt(sums[3:6].numpy() / cnt - mean ** 2)
            dn = ((d - mean[None, :, None, None]) / std[None, :, None, None]).astype(np.float32)
            out.append(bd.all_gather_rows(torch.from_numpy(dn.reshape(dn.shape[0], -1))).numpy())
        return out

    def distance(A, B):
        dirs = np.concatenate([so.draw_directions(A.shape[1], dpr) for _ in range(repeats)], axis=1)
        block = bd.direction_block(repeats * dpr, rank, world)
        pa, pb = np.sort(A @ dirs[:, block], axis=0), np.sort(B @ dirs[:, block], axis=0)
        partial = torch.zeros(repeats, dtype=torch.float64)
        for rep, lo, hi in bd.repeat_segments(block, dpr):
            partial[rep] += float(np.abs(pa[:, lo:hi] - pb[:, lo:hi]).astype(np.float64).sum())
        return bd.combine_repeat_sums(partial, A.shape[0], dpr)

    d_real = collect(reals)
    for d in d_real:  # the reference's real-vs-real pass draws (and discards) its directions
        for _ in range(repeats):
            np.random.randn(d.shape[1], dpr)
    d_fake = collect(fakes)
    res = [distance(a, b) for a, b in zip(d_real, d_fake)]
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array(res))
    dist.destroy_process_group()


def test_distributed_swd_protocol_matches_single_process(tmp_path):
    from oracle import swd_oracle as so

    mp.spawn(_worker, args=(WORLD, _free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    got = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(WORLD)]
    np.testing.assert_array_equal(got[0], got[1])
    # single-process oracle, minibatches fed in the same global order, same seed
    reals = uniform_pm1(41, (8, 3, 32, 32))
    fakes = (0.7 * uniform_pm1(42, (8, 3, 32, 32)) + 0.05).astype(np.float32)
    rng = np.random.RandomState(123)

    def collect(images):
        chunks = [[] for _ in range(2)]
        for i in range(0, len(images), 2):
            for lod, lv in enumerate(so.generate_laplacian_pyramid(images[i : i + 2], 2)):
                chunks[lod].append(so.get_descriptors_for_minibatch(lv, 7, 16, rng))
        return [so.finalize_descriptors(c) for c in chunks]

    d_real = collect(reals)
    for d in d_real:
        for _ in range(4):
            rng.randn(d.shape[1], 8)
    d_fake = collect(fakes)
    want = [so.sliced_wasserstein(a, b, 4, 8, rng) for a, b in zip(d_real, d_fake)]
    np.testing.assert_allclose(got[0], want, rtol=2e-5)  # (row order and fp64 statistics differ from numpy's fp32 means)


def test_shard_and_direction_helpers():
    from blurred_gan_b200 import dist as bd

    for n in (0, 1, 7, 256, 768):
        for world in (1, 2, 3, 8):
            parts = [bd.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    covered = []
    for r in range(3):
        blk = bd.direction_block(512, r, 3)
        for rep, lo, hi in bd.repeat_segments(blk, 128):
            assert (blk.start + lo) // 128 == rep == (blk.start + hi - 1) // 128
            covered += list(range(blk.start + lo, blk.start + hi))
    assert covered == list(range(512))
    assert bd.world_info() == (0, 1)


# ---- host RNG replay (csrc/hostrng.cu): bit-exact with numpy's legacy randint, in place on numpy's global state ----------
@pytest.mark.parametrize("low,high,count", [(3, 125, 8192), (3, 61, 8192), (3, 13, 77), (3, 5, 10), (3, 4, 5), (0, 1 << 20, 3000),
                                            (-5, 7, 100), (3, 125, 65536 * 3 + 17), (3, 125, 0)])
def test_mt19937_randint_replay_is_numpy_bit_for_bit(low, high, count):
    import ctypes

    from blurred_gan_b200 import _lib

    lib = _lib.load()
    for seed in (0, 1, 123):
        np.random.seed(seed)
        np.random.randn(3)  # leaves a cached gaussian in the state: must survive
        want = np.random.randint(low, high, size=(count, 1, 1, 1)).reshape(-1)
        tail = (np.random.randn(4), np.random.randint(0, 1000, 5))
        for store in (True, False):
            np.random.seed(seed)
            np.random.randn(3)
            st = np.random.get_state(legacy=True)
            key, pos = np.array(st[1], dtype=np.uint32, copy=True), ctypes.c_int(int(st[2]))
            out = np.full(count, -1, dtype=np.int32)
            rc = lib.bgan_mt19937_randint_i32(key.ctypes.data, ctypes.byref(pos), low, high, count, out.ctypes.data if store else None)
            assert rc == 0
            np.random.set_state(("MT19937", key, pos.value, st[3], st[4]))
            if store:
                assert np.array_equal(out, want)
            got_tail = (np.random.randn(4), np.random.randint(0, 1000, 5))
            assert np.array_equal(got_tail[0], tail[0]) and np.array_equal(got_tail[1], tail[1])  # the stream continues identically


def test_mt19937_randint_argument_errors():
    import ctypes

    from blurred_gan_b200 import _lib

    lib = _lib.load()
    key, pos = np.zeros(624, dtype=np.uint32), ctypes.c_int(624)
    out = np.zeros(4, dtype=np.int32)
    assert lib.bgan_mt19937_randint_i32(key.ctypes.data, ctypes.byref(pos), 5, 5, 4, out.ctypes.data) == -1  # low >= high
    assert lib.bgan_mt19937_randint_i32(key.ctypes.data, ctypes.byref(pos), 0, 1 << 33, 4, out.ctypes.data) == -3  # 64-bit path of numpy
    pos = ctypes.c_int(700)
    assert lib.bgan_mt19937_randint_i32(key.ctypes.data, ctypes.byref(pos), 0, 5, 4, out.ctypes.data) == -1


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 1), (3, 0), (8, 5)])
def test_packed_positions_equal_the_numpy_draws(world, rank):
    from blurred_gan_b200 import dist as bd

    sizes = [(128, 128), (64, 64), (32, 48), (16, 16)]
    np.random.seed(7)
    ref = bd.positions_for_round(5, sizes, 7, 128, rank, world)
    ref_tail = np.random.randn(3)
    np.random.seed(7)
    got = bd.positions_packed_for_round(5, sizes, 7, 128, rank, world)
    got_tail = np.random.randn(3)
    assert got.dtype == np.int32
    assert np.array_equal(got, np.concatenate([np.concatenate(p) for p in ref]))
    assert np.array_equal(ref_tail, got_tail)
    # a global generator that is not the legacy MT19937 falls back to numpy's own randint
    import unittest.mock as mock

    with mock.patch.object(bd, "_numpy_mt19937_state", return_value=None):
        np.random.seed(7)
        again = bd.positions_packed_for_round(5, sizes, 7, 128, rank, world)
    assert np.array_equal(again, got)


def test_mt19937_scalar_avx2_and_avx512_paths_agree(tmp_path):
    """The replay has a scalar body, an AVX2 body and an AVX-512 body (csrc/hostrng_avx2.cpp, hostrng_avx512.cpp; the widest
    one the CPU supports is chosen at run time): same draws, same final state.  The choice is latched at the first call, so
    each body runs in its own child process (BGAN_RNG_SCALAR=1 / BGAN_RNG_NO_AVX512=1 pin the narrower ones)."""
    import subprocess
    import sys

    script = r"""
import ctypes, sys, numpy as np
sys.path.insert(0, %r)
from blurred_gan_b200 import _lib
lib = _lib.load()
st = np.random.RandomState(7).get_state()
key, pos = np.array(st[1], dtype=np.uint32, copy=True), ctypes.c_int(int(st[2]))
acc = []
for low, high, count, store in [(3, 125, 65536, True), (3, 61, 4099, False), (3, 29, 7, True), (3, 13, 8, True), (0, 1 << 31, 1000, True),
                                (3, 125, 623, False), (3, 125, 625, True), (-9, 1000003, 31337, True)]:
    out = np.full(count, -1, dtype=np.int32)
    assert lib.bgan_mt19937_randint_i32(key.ctypes.data, ctypes.byref(pos), low, high, count, out.ctypes.data if store else None) == 0
    acc.append(out)
np.savez(sys.argv[1], key=key, pos=pos.value, **{"o%%d" %% i: a for i, a in enumerate(acc)})
""" % ROOT
    results = []
    for i, pins in enumerate(({}, {"BGAN_RNG_NO_AVX512": "1"}, {"BGAN_RNG_SCALAR": "1"})):
        path = str(tmp_path / ("rng%d.npz" % i))
        env = {k: v for k, v in os.environ.items() if not k.startswith("BGAN_RNG_")}
        env.update(pins)
        subprocess.run([sys.executable, "-c", script, path], check=True, env=env, timeout=120)
        results.append(np.load(path))
    a = results[0]
    for b in results[1:]:
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert np.array_equal(a[k], b[k]), k
    # and both are numpy's draws
    rs = np.random.RandomState(7)
    assert np.array_equal(a["o0"], rs.randint(3, 125, size=65536))


# ---- data-parallel train step (BASELINE configs[2]): gradients all-reduced over gloo == the single-process step -----------
def _tiny_gan(blur):
    """A blurred WGAN with deterministic linear nets (no dropout / batch norm), so that a sharded step must reproduce the
    single-process step exactly up to the summation order of the all-reduce."""
    import torch.nn as nn

    from blurred_gan_b200.blurred_gan import BlurredWGAN

    class Gen(nn.Module):
        latent_size = 6

        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(6, 3 * 12 * 12)

        def forward(self, z):
            return torch.tanh(self.fc(z)).view(-1, 3, 12, 12)

    torch.manual_seed(11)
    gen, critic = Gen(), nn.Sequential(nn.Flatten(), nn.Linear(3 * 12 * 12, 1))
    hp = BlurredWGAN.HyperParameters(learning_rate=1e-2, batch_size=8, global_batch_size=8, initial_blur_std=1.5)
    return gen, critic, hp, blur


def _step_worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_wgan_harness import OracleBlur2D

    from blurred_gan_b200.blurred_gan import BlurredWGAN
    from blurred_gan_b200.dist import allreduce_gradients

    gen, critic, hp, blur = _tiny_gan(OracleBlur2D(1.5))
    gan = BlurredWGAN(gen, critic, hp, blur=blur, grad_sync=allreduce_gradients if world > 1 else None)
    torch.manual_seed(5)
    reals = torch.rand(8, 3, 12, 12) * 2 - 1
    latents = [torch.rand(8, 6) for _ in range(4)]  # critic step and generator step of two batches
    lo, hi = rank * 8 // world, (rank + 1) * 8 // world
    draws = iter(latents)
    gan.latents_batch = lambda: next(draws)[lo:hi]  # every rank sees its shard of the same global latents
    for _ in range(2):
        gan.train_on_batch(reals[lo:hi])
    state = {k: v.detach().clone() for k, v in list(gan.generator.state_dict().items()) + list(gan.discriminator[1].state_dict().items())}
    torch.save(state, os.path.join(out_dir, "w%d_r%d.pt" % (world, rank)))
    if world > 1:
        dist.destroy_process_group()


def test_sharded_train_step_equals_the_single_process_step(tmp_path):
    """SURVEY 8e / configs[2]: the batch is sharded over ranks, the losses are divided by the *global* batch size
    (wgan.py:130,157), the gradients are summed over ranks (dist.allreduce_gradients) -- two Adam steps later every rank
    holds the weights of the single-process run on the whole batch."""
    mp.spawn(_step_worker, args=(WORLD, _free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    _step_worker(0, 1, 0, str(tmp_path))
    single = torch.load(tmp_path / "w1_r0.pt")
    ranks = [torch.load(tmp_path / ("w%d_r%d.pt" % (WORLD, r))) for r in range(WORLD)]
    for k, want in single.items():
        assert torch.equal(ranks[0][k], ranks[1][k]), k          # the ranks stay in lock-step
        assert torch.allclose(ranks[0][k], want, rtol=1e-5, atol=1e-6), k
    # and the step did move the weights
    gen0 = _tiny_gan(None)[0]
    assert not torch.allclose(single["fc.weight"], gen0.state_dict()["fc.weight"])
