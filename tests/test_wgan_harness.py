"""Train-step harness (N1 WGAN-GP step, N2 sigma controllers, N4 SWD metric).

CPU tests run the harness with the *oracle's* blur layer (the reference's two F.conv2d calls) because the package has no
CPU path; the GPU test runs the same step through the B200 blur and checks it against the oracle-blur step with
identical weights, latents and dropout masks.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn as nn

from blurred_gan_b200 import callbacks, nets
from blurred_gan_b200.blurred_gan import BlurredWGAN, BlurredWGANGP
from blurred_gan_b200.wgan import WGANGP, gradient_penalty
from oracle import blur_oracle as bo


class OracleBlur2D(nn.Module):
    """CPU stand-in for GaussianBlur2D: TF-rule plan (gaussian_blur.py:50-80) + explicit-tap depthwise F.conv2d."""

    def __init__(self, initial_std=0.01):
        super().__init__()
        self.std = nn.Parameter(torch.tensor(float(initial_std)), requires_grad=False)
        self.calls = 0

    def assign(self, value):
        self.std.data.fill_(float(value))

    def forward(self, x):
        self.calls += 1
        n, c, h, w = x.shape
        taps = torch.from_numpy(bo.taps_for(bo.RULE_TF, float(self.std.item()), h, w)).to(x)
        r = taps.numel() // 2
        kv = taps.view(1, 1, -1, 1).repeat(c, 1, 1, 1)
        kh = taps.view(1, 1, 1, -1).repeat(c, 1, 1, 1)
        y = torch.nn.functional.conv2d(x, kv, padding=(r, 0), groups=c)  # vertical first (gaussian_blur.py:116-130)
        return torch.nn.functional.conv2d(y, kh, padding=(0, r), groups=c)


class CountingBlur(OracleBlur2D):
    """Counts blur applications including backward and double backward (the operator is self-adjoint)."""

    def forward(self, x):
        outer = self
        taps = torch.from_numpy(bo.taps_for(bo.RULE_TF, float(self.std.item()), x.shape[2], x.shape[3])).to(x)

        class Op(torch.autograd.Function):
            @staticmethod
            def forward(ctx, t):
                outer.applications += 1
                return OracleBlur2D.forward(outer, t)

            @staticmethod
            def backward(ctx, g):
                return Op.apply(g)

        return Op.apply(x)


def _mnist_gan(blur, batch=64, cls=BlurredWGANGP):
    torch.manual_seed(0)
    hp = cls.HyperParameters(batch_size=batch, global_batch_size=batch, initial_blur_std=23.0)
    return cls(nets.MnistGenerator(), nets.MnistDiscriminator(), hp, blur=blur)


def test_nets_have_the_reference_shapes():
    g, d = nets.MnistGenerator(), nets.MnistDiscriminator()
    z = torch.rand(4, 100)
    g.eval()
    assert g(z).shape == (4, 1, 28, 28) and d(g(z)).shape == (4, 1)
    gc, dc = nets.CelebaGenerator(), nets.CelebaDiscriminator()
    gc.eval()
    x = gc(torch.rand(2, 100))
    assert x.shape == (2, 3, 128, 128) and dc(x).shape == (2, 1)
    # parameter counts quoted in SURVEY section 2 (demo_celeba.py nets: ~11.7 M / ~4.4 M)
    assert abs(sum(p.numel() for p in gc.parameters()) / 1e6 - 11.7) < 0.3
    assert abs(sum(p.numel() for p in dc.parameters()) / 1e6 - 4.4) < 0.2


def test_cfg1_mnist_step_on_cpu_issues_seven_blurs_and_anneals():
    """BASELINE configs[0]: demo_mnist-shaped blurred WGAN-GP step, batch 64 x 28 x 28 x 1, std annealed 23 -> 0, on CPU."""
    blur = CountingBlur(23.0)
    blur.applications = 0
    gan = _mnist_gan(blur)
    sched = callbacks.LinearBlurSchedule(total_batches=3, max_value=23.0, min_value=0.0)
    sched.set_model(gan)
    rng = np.random.RandomState(0)
    stds = []
    for step in range(3):
        sched.on_batch_begin()
        stds.append(float(gan.std.item()))
        before = blur.applications
        reals = torch.from_numpy((rng.rand(64, 1, 28, 28) * 2 - 1).astype(np.float32))
        gan.train_on_batch(reals)
        assert blur.applications - before == 7  # 3 fwd + 1 bwd + 1 double-bwd (critic), 1 fwd + 1 bwd (generator)
        m = gan.metric_results()
        assert all(math.isfinite(v) for v in m.values()), m
        assert m["std"] == pytest.approx(stds[-1])
    assert stds == pytest.approx([23.0, 11.5, 0.0])
    assert gan.n_img == 192 and gan.n_batches == 3


def test_fused_critic_step_equals_the_reference_step_with_six_blurs():
    """one blur call for (fakes, reals) + interpolates taken between the blurred batches (the blur is
    linear) + B applied to the critic-input gradient (B is self-adjoint) == the reference's critic step."""
    results = []
    for fused in (False, True):
        blur = CountingBlur(2.5)
        blur.applications = 0
        torch.manual_seed(0)
        hp = BlurredWGANGP.HyperParameters(batch_size=16, global_batch_size=16, initial_blur_std=2.5)
        gan = BlurredWGANGP(nets.MnistGenerator(), nets.MnistDiscriminator(), hp, blur=blur, fused_critic=fused)
        torch.manual_seed(1)
        reals = torch.rand(16, 1, 28, 28) * 2 - 1
        torch.manual_seed(2)
        gan.train_on_batch(reals)
        results.append((blur.applications, gan.metric_results(), [p.detach().clone() for p in gan.discriminator[1].parameters()]))
    (n0, m0, p0), (n1, m1, p1) = results
    assert (n0, n1) == (7, 6)  # critic: 2 fwd + B(grad) + its backward instead of 3 fwd + 1 bwd + 1 double-bwd
    for k in m0:
        assert m1[k] == pytest.approx(m0[k], rel=1e-4, abs=1e-6), k
    for a, b in zip(p0, p1):
        # an Adam step moves every weight by ~lr = 1e-3 in the direction of the gradient's sign: equal steps <=> equal gradients
        assert (a - b).abs().max().item() <= 1e-4


def test_drift_term_is_a_vector_like_the_reference():
    gan = _mnist_gan(OracleBlur2D(1.0), batch=8)
    reals = torch.rand(8, 1, 28, 28) * 2 - 1
    gan.batch_size = 8
    fakes = gan.generate_samples(training=False).detach()
    loss = gan.discriminator_loss(reals, fakes, gan._scores(reals, True), gan._scores(fakes, True))
    assert loss.shape == (8,)  # wgan.py:282-284: scalar + [B] vector
    plain = BlurredWGAN(nets.MnistGenerator(), nets.MnistDiscriminator(), BlurredWGAN.HyperParameters(global_batch_size=8), blur=OracleBlur2D(1.0))
    plain.batch_size = 8
    assert plain.discriminator_loss(reals, fakes, torch.ones(8, 1), torch.zeros(8, 1)).item() == pytest.approx(-1.0)


def test_gradient_penalty_of_a_linear_critic():
    w = torch.randn(1, 1, 28, 28)
    critic = lambda x: (x * w).sum(dim=(1, 2, 3)).unsqueeze(1)  # noqa: E731
    critic.train = lambda mode=True: None
    gp = gradient_penalty(critic, torch.rand(5, 1, 28, 28), torch.rand(5, 1, 28, 28))
    assert gp.item() == pytest.approx((w.norm().item() - 1.0) ** 2, rel=1e-5)


def test_blur_decay_controller_matches_the_keras_schedule():
    ctl = callbacks.BlurDecayController(total_n_training_examples=60000 * 10, max_value=5.0)
    # tf.keras ExponentialDecay(5, decay_steps=60000, 0.96)(step) = 5 * 0.96 ** (step / 60000)
    assert ctl.schedule(0) == 5.0
    assert ctl.schedule(18750) == pytest.approx(5.0 * 0.96 ** (18750 / 60000.0))  # 10 epochs of batch 32: sigma only falls to 0.987 * max
    assert ctl.schedule(18750) / 5.0 == pytest.approx(0.9873, abs=1e-3)


def test_adaptive_controller_tracks_but_does_not_assign_by_default():
    gan = _mnist_gan(OracleBlur2D(23.5), batch=4)
    ctl = callbacks.AdaptiveBlurController(warmup_n_batches=0)
    ctl.set_model(gan)
    ctl.on_train_begin()
    assert float(gan.std.item()) == 23.5
    for b in range(250):
        ctl.on_batch_end(b, {"fake_scores": 1.0, "real_scores": 1.0})  # ratio 0.5: "stable"
    # lowered once at batch 100 and once at batch 200 (delay_between_modifications = 100)
    assert ctl.std == pytest.approx(23.5 * 0.99 ** 2)
    assert float(gan.std.item()) == 23.5  # callbacks.py:102-103: the assignment is disabled in the reference


def test_adaptive_controller_stops_the_model_below_min_value():
    """callbacks.py:133-134: once the tracked sigma falls below min_value the *model's* stop_training flag is raised."""
    gan = _mnist_gan(OracleBlur2D(1.0), batch=4)
    ctl = callbacks.AdaptiveBlurController(warmup_n_batches=0, max_value=0.0101, min_value=0.01)
    ctl.set_model(gan)
    assert not getattr(gan, "stop_training", False)
    for b in range(0, 301, 100):
        ctl.on_batch_end(b, {"fake_scores": 1.0, "real_scores": 1.0})
    assert ctl.std < 0.01 and ctl.stop_training and gan.stop_training is True


def test_assign_std_reaches_layers_without_assign():
    """a blur layer that only offers `set_std` (or nothing) must still see the controller's value."""

    class OnlySetStd(OracleBlur2D):
        assign = None

        def set_std(self, value):
            with torch.no_grad():
                self.std.fill_(float(value))

    gan = _mnist_gan(OnlySetStd(5.0), batch=4)
    gan.assign_std(2.5)
    assert float(gan.std.item()) == 2.5

    class Bare(OracleBlur2D):
        assign = None

    gan = _mnist_gan(Bare(5.0), batch=4)
    v = gan.std._version
    gan.assign_std(1.25)
    assert float(gan.std.item()) == 1.25 and gan.std._version > v  # an in-place op on the Parameter, not a `.data` write


@pytest.mark.gpu
def test_celeba_step_on_gpu_matches_the_oracle_blur_step(monkeypatch):
    """BASELINE configs[2] shape (small batch): the step through the B200 blur == the step through the oracle blur."""
    from blurred_gan_b200.gaussian_blur import GaussianBlur2D

    # the nets run on cuDNN: with TF32 convolutions a 1e-7 difference between the two blurs flips 10-bit roundings and shows up
    # as ~1e-3 in the scores, so compare the two steps in real fp32
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    results = []
    for kind in ("b200", "oracle"):
        torch.manual_seed(0)
        blur = GaussianBlur2D(3.0) if kind == "b200" else OracleBlur2D(3.0)
        hp = BlurredWGANGP.HyperParameters(batch_size=8, global_batch_size=8, initial_blur_std=3.0)
        gan = BlurredWGANGP(nets.CelebaGenerator(), nets.CelebaDiscriminator(), hp, blur=blur).cuda()
        torch.manual_seed(1)
        reals = torch.rand(8, 3, 128, 128, device="cuda") * 2 - 1
        torch.manual_seed(2)
        gan.train_on_batch(reals)
        results.append((gan.metric_results(), [p.detach().clone() for p in gan.discriminator[1].parameters()]))
    (m0, p0), (m1, p1) = results
    for k in m0:
        assert m0[k] == pytest.approx(m1[k], rel=2e-3, abs=1e-5), k
    for a, b in zip(p0, p1):
        # one Adam step moves every weight by ~lr; what is compared is the direction, which depends on the blurred gradients
        assert (a - b).abs().max().item() <= 2e-3 * 1.01
        assert ((a - b).abs() > 1e-4).float().mean().item() < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("controller", ["linear", "decay"])
def test_sigma_controllers_drive_the_device_blur_without_synchronising(controller):
    """SURVEY 8f row N2 on the device: a controller (callbacks.py:45-62 / the driver's linear 23 -> 0 anneal) writes sigma
    before every batch and the blur layer in front of the critic picks it up -- radius, kernel choice and taps -- with parity
    at every step and without a single device synchronisation (sync debug mode raises on any)."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import GaussianBlur2D

    gan = _mnist_gan(GaussianBlur2D(23.0), batch=4).cuda()
    assert gan.std.device.type == "cpu"  # sigma lives in host memory (launch parameter), the nets on the GPU
    steps = 6
    if controller == "linear":
        ctl = callbacks.LinearBlurSchedule(total_batches=steps, max_value=23.0, min_value=0.0)
        want_std = [23.0 - 23.0 * i / (steps - 1) for i in range(steps)]
    else:
        ctl = callbacks.BlurDecayController(total_n_training_examples=20, max_value=8.0)  # decay_steps = 2 batches
        want_std = [8.0 * 0.96 ** (i / 2.0) for i in range(steps)]
    ctl.set_model(gan)
    x = torch.from_numpy(np.ascontiguousarray(np.random.RandomState(5).rand(4, 128, 128, 3).astype(np.float32) * 2 - 1)).cuda()
    x_host = x.cpu().numpy()
    outs, paths = [], []
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        for i in range(steps):
            gan.n_batches = i
            ctl.on_batch_begin()
            outs.append(gan.blur(x))  # what Sequential([blur, discriminator]) feeds the critic (blurred_gan.py:30-35)
            paths.append(ops.last_blur_path())
    finally:
        torch.cuda.set_sync_debug_mode("default")
    if controller == "linear":  # 23 -> 0 crosses the kernel choices: tensor-core kernel for wide taps, direct kernel for few
        assert len(set(paths)) > 1, paths
    for i, y in enumerate(outs):
        assert float(gan.blur.std_value()) == pytest.approx(want_std[-1])
        want = bo.blur_images_tf_ref(x_host, np.float32(want_std[i]))
        err = np.abs(y.cpu().numpy() - want).max() / np.abs(want).max()
        assert err <= 1e-5, "step %d (std %.4f): %.3g" % (i, want_std[i], err)


@pytest.mark.gpu
def test_swd_metric_on_device_tensors():
    from blurred_gan_b200.metrics import SWDMetric

    np.random.seed(0)
    torch.manual_seed(0)
    m = SWDMetric((3, 64, 64))
    for _ in range(2):
        m.update_state(torch.rand(16, 3, 64, 64, device="cuda") * 2 - 1, torch.rand(16, 3, 64, 64, device="cuda") * 2 - 1)
    res = m.result()
    assert len(res) == len(m.metric_names()) and all(math.isfinite(v) and v > 0 for v in res)
    # the metric is the API fed with the same minibatches under the same numpy seed
    from blurred_gan_b200 import sliced_wasserstein as sw

    torch.manual_seed(0)
    batches = [(torch.rand(16, 3, 64, 64, device="cuda") * 2 - 1, torch.rand(16, 3, 64, 64, device="cuda") * 2 - 1) for _ in range(2)]
    np.random.seed(0)
    api = sw.API((3, 64, 64))
    for mode, k in (("reals", 0), ("fakes", 1)):
        api.begin(mode)
        for b in batches:
            api.feed(mode, b[k])
        want = api.end(mode)
    assert res == pytest.approx(want, rel=1e-6)


@pytest.mark.gpu
def test_fused_critic_step_on_gpu(monkeypatch):
    """Row N3 on the B200 path: forward_many (one launch for both batches) + lerp_rows between the blurred batches +
    B on the critic-input gradient gives the unfused step's losses; 5 library blur calls per step instead of 7."""
    from blurred_gan_b200 import ops
    from blurred_gan_b200.gaussian_blur import GaussianBlur2D

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    calls = {"single": 0, "multi": 0}
    real_single, real_multi = ops.blur2d, ops.blur2d_multi
    monkeypatch.setattr(ops, "blur2d", lambda *a, **k: (calls.__setitem__("single", calls["single"] + 1), real_single(*a, **k))[1])
    monkeypatch.setattr(ops, "blur2d_multi", lambda *a, **k: (calls.__setitem__("multi", calls["multi"] + 1), real_multi(*a, **k))[1])
    results = []
    for fused in (False, True):
        calls["single"] = calls["multi"] = 0
        torch.manual_seed(0)
        hp = BlurredWGANGP.HyperParameters(batch_size=8, global_batch_size=8, initial_blur_std=3.0)
        gan = BlurredWGANGP(nets.CelebaGenerator(), nets.CelebaDiscriminator(), hp, blur=GaussianBlur2D(3.0), fused_critic=fused).cuda()
        torch.manual_seed(1)
        reals = torch.rand(8, 3, 128, 128, device="cuda") * 2 - 1
        torch.manual_seed(2)
        gan.train_on_batch(reals)
        results.append((dict(calls), gan.metric_results(), [p.detach().clone() for p in gan.discriminator[1].parameters()]))
    (c0, m0, p0), (c1, m1, p1) = results
    # library blur calls per step, backward and double backward included (the ops' autograd formulas call the same ops)
    assert c0 == {"single": 7, "multi": 0}
    assert c1 == {"single": 4, "multi": 1}
    for k in m0:
        assert m1[k] == pytest.approx(m0[k], rel=2e-3, abs=1e-5), k
    for a, b in zip(p0, p1):
        assert (a - b).abs().max().item() <= 2e-3 * 1.01
        assert ((a - b).abs() > 1e-4).float().mean().item() < 0.02
