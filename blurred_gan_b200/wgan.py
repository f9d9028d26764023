"""WGAN / WGAN-GP training step in PyTorch: the *caller* of the blur hot path.

Restates reference ``wgan.py:28-285`` (Keras-subclassed WGAN, WGAN-GP loss with the ProGAN drift term) on torch
modules, so that BASELINE configs[0] (MNIST-shaped step) and configs[2] (CelebA-128 step, batch sharded over GPUs)
can be run around the B200 blur.  The networks' arithmetic is cuDNN / ATen library code and is not an optimisation
target; what matters here is that the step issues exactly the reference's seven blur applications (3 forward +
1 backward + 1 double-backward in the critic step, 1 forward + 1 backward in the generator step) on
tensors of the reference's shapes.

Reference behaviours that are kept on purpose (each one is visible in the losses):

* latents are ``U[0, 1)`` (``wgan.py:118``), not normal;
* the Wasserstein terms are ``sum(...) / global_batch_size`` (``wgan.py:130,157``): the one data-parallel-aware line;
* critic scores for the loss use dropout (``training=True``, ``wgan.py:136-137``), the gradient-penalty pass and the
  generator step do not (``wgan.py:163,242``); the generator runs in inference mode inside the critic step
  (``wgan.py:135``: BatchNorm uses its moving statistics there);
* the drift term ``e_drift * (|fake_score| + |real_score|)`` is a length-B vector added to the scalar loss
  (``wgan.py:282-284``); the tape differentiates its sum, so the Wasserstein and penalty terms are weighted B times
  more than the drift term.  ``WGANGP.discriminator_loss`` returns that vector, ``discriminator_step`` backpropagates
  its sum;
* optimiser: Keras ``"adam"`` defaults (``wgan.py:56-61``): lr from the hyper-parameters, betas (0.9, 0.999), eps 1e-7.

Not restated: Keras ``fit`` plumbing, TensorBoard summaries, ``simple_parsing`` flags (out of scope).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn


@dataclass
class TrainingConfig:
    """``wgan.py:19-25``."""

    log_dir: str = "results/log"
    checkpoint_dir: str = "results/log/checkpoints"
    save_image_summaries_interval: int = 50


class _Mean:
    """``tf.keras.metrics.Mean``: running mean of everything it was called with since the last reset.

    Values stay on their device until ``result()`` is asked for, so a training step never synchronises on a metric."""

    def __init__(self, name: str):
        self.name = name
        self.reset_states()

    def reset_states(self):
        self._sum, self._n = None, 0

    def __call__(self, value):
        v = value.detach().float() if torch.is_tensor(value) else torch.tensor(float(value))
        s = v.sum()
        self._sum = s if self._sum is None else self._sum + s.to(self._sum.device)
        self._n += v.numel()

    def result(self) -> float:
        return float(self._sum.item()) / self._n if self._n else 0.0


class WGAN(nn.Module):
    """``wgan.py:28-231``."""

    @dataclass
    class HyperParameters:
        """``wgan.py:34-43``."""

        learning_rate: float = 0.001
        d_steps_per_g_step: int = 1
        batch_size: int = 32
        global_batch_size: int = 32
        optimizer: str = "adam"

    def __init__(self, generator: nn.Module, discriminator: nn.Module, hyperparams: "WGAN.HyperParameters",
                 config: Optional[TrainingConfig] = None, latent_size: Optional[int] = None, grad_sync=None):
        """``grad_sync(params)``, if given, is called on the freshly computed gradients before every optimiser step
        (data-parallel runs pass ``blurred_gan_b200.dist.allreduce_gradients``: the losses are already divided by
        the *global* batch size, so the sum over ranks is the gradient of the global batch)."""
        super().__init__()
        self.grad_sync = grad_sync
        self.hparams = hyperparams
        self.config = config or TrainingConfig()
        self.generator = generator
        self.discriminator = discriminator
        if hyperparams.optimizer != "adam":
            raise ValueError("only the reference's default optimizer ('adam') is restated")
        self.generator_optimizer = torch.optim.Adam(self.generator.parameters(), lr=hyperparams.learning_rate, betas=(0.9, 0.999), eps=1e-7)
        self.discriminator_optimizer = torch.optim.Adam(self._trainable(self.discriminator), lr=hyperparams.learning_rate, betas=(0.9, 0.999),
                                                        eps=1e-7)
        self.d_steps_per_g_step = hyperparams.d_steps_per_g_step
        self.batch_size = None  # determined by the first batch (wgan.py:65)
        self.latent_size = latent_size if latent_size is not None else getattr(_unwrap(generator), "latent_size")
        self.n_img = 0      # wgan.py:70-71: progress counters (checkpointed state in the reference)
        self.n_batches = 0
        self.images = None
        self.real_scores_metric = _Mean("real_scores")
        self.fake_scores_metric = _Mean("fake_scores")
        self.gen_loss_metric = _Mean("gen_loss")
        self.disc_loss_metric = _Mean("disc_loss")

    @staticmethod
    def _trainable(module):
        return [p for p in module.parameters() if p.requires_grad]

    # ---- metrics ---------------------------------------------------------------------------------
    def _metrics(self):
        return [m for m in vars(self).values() if isinstance(m, _Mean)]

    def reset_metrics(self):
        for m in self._metrics():
            m.reset_states()

    def metric_results(self) -> dict:
        return {m.name: m.result() for m in self._metrics()}

    # ---- wgan.py:86-114 ----------------------------------------------------------------------------
    def train_on_batch(self, reals: torch.Tensor) -> dict:
        self.reset_metrics()
        self.batch_size = reals.shape[0]
        _, self.images = self.discriminator_step(reals)
        if self.n_batches % self.d_steps_per_g_step == 0:
            self.generator_step()
        self.n_img += reals.shape[0]
        self.n_batches += 1
        return {"n_img": self.n_img, "n_batches": self.n_batches}

    def latents_batch(self) -> torch.Tensor:
        assert self.batch_size is not None
        dev = next(self.generator.parameters()).device
        return torch.rand(self.batch_size, self.latent_size, device=dev)  # wgan.py:118: uniform, not normal

    def generate_samples(self, latents=None, training=False) -> torch.Tensor:
        if latents is None:
            latents = self.latents_batch()
        self.generator.train(training)
        return self.generator(latents)

    def _scores(self, images, training: bool) -> torch.Tensor:
        self.discriminator.train(training)
        return self.discriminator(images)

    # ---- critic -------------------------------------------------------------------------------------
    def _critic_scores(self, fakes, reals):
        """``wgan.py:136-137``: scores of the generated, then of the real batch, dropout active."""
        return self._scores(fakes, training=True), self._scores(reals, training=True)

    def discriminator_loss(self, reals, fakes, real_scores, fake_scores) -> torch.Tensor:
        return (fake_scores - real_scores).sum() * (1.0 / self.hparams.global_batch_size)  # wgan.py:128-130

    def discriminator_step(self, reals):
        """``wgan.py:132-151``."""
        with torch.no_grad():  # the tape only reaches the critic's variables: gradients w.r.t. the fakes are pruned
            fakes = self.generate_samples(training=False)
        fake_scores, real_scores = self._critic_scores(fakes, reals)
        disc_loss = self.discriminator_loss(reals, fakes, real_scores, fake_scores)
        self.discriminator_optimizer.zero_grad(set_to_none=True)
        # like tape.gradient(loss, trainable_variables): only the paths that reach the critic's variables are
        # differentiated (x_hat of the penalty is a leaf too; .backward() would blur once more on its way there)
        params = self._trainable(self.discriminator)
        for p, g in zip(params, torch.autograd.grad(disc_loss.sum(), params)):
            p.grad = g
        if self.grad_sync is not None:
            self.grad_sync(params)
        self.discriminator_optimizer.step()
        self.fake_scores_metric(fake_scores)
        self.real_scores_metric(real_scores)
        self.disc_loss_metric(disc_loss)
        return disc_loss.detach(), (fakes, reals)

    # ---- generator ----------------------------------------------------------------------------------
    def generator_loss(self, fake_scores) -> torch.Tensor:
        return -fake_scores.sum() * (1.0 / self.hparams.global_batch_size)  # wgan.py:155-157

    def generator_step(self):
        """``wgan.py:159-172``."""
        fakes = self.generate_samples(training=True)
        fake_scores = self._scores(fakes, training=False)
        gen_loss = self.generator_loss(fake_scores)
        self.generator_optimizer.zero_grad(set_to_none=True)
        # the critic's variables are not watched by the generator's tape: do not accumulate into their .grad
        grads = torch.autograd.grad(gen_loss, list(self.generator.parameters()))
        for p, g in zip(self.generator.parameters(), grads):
            p.grad = g
        if self.grad_sync is not None:
            self.grad_sync(list(self.generator.parameters()))
        self.generator_optimizer.step()
        self.fake_scores_metric(fake_scores)
        self.gen_loss_metric(gen_loss)
        return gen_loss.detach()

    def count_params(self) -> int:
        return sum(p.numel() for p in self.parameters())


def _unwrap(module):
    return module.module if hasattr(module, "module") and isinstance(module.module, nn.Module) else module


def gradient_penalty(discriminator, reals, fakes) -> torch.Tensor:
    """``wgan.py:234-246``: ``E[(||grad_x D(x_hat)||_2 - 1)^2]`` on random interpolates, differentiable w.r.t. the
    critic's variables (the critic -- and with it the blur -- is differentiated twice)."""
    batch_size = reals.shape[0]
    a = torch.rand(batch_size, 1, 1, 1, device=reals.device, dtype=reals.dtype)
    x_hat = (reals + a * (fakes - reals)).detach().requires_grad_(True)
    discriminator.train(False)  # wgan.py:242: training=False
    y_hat = discriminator(x_hat)
    (grad,) = torch.autograd.grad(y_hat, x_hat, grad_outputs=torch.ones_like(y_hat), create_graph=True)
    norm = grad.reshape(batch_size, -1).norm(dim=1)
    return ((norm - 1.0) ** 2).mean()


class WGANGP(WGAN):
    """``wgan.py:249-285``."""

    @dataclass
    class HyperParameters(WGAN.HyperParameters):
        e_drift: float = 1e-4
        gp_coefficient: float = 10.0

    def __init__(self, generator, discriminator, hyperparams, config=None, latent_size=None, grad_sync=None):
        super().__init__(generator, discriminator, hyperparams, config, latent_size, grad_sync)
        self.gp_term_metric = _Mean("gp_term")
        self.norm_term_metric = _Mean("norm_term")

    def _gradient_penalty(self, reals, fakes) -> torch.Tensor:
        return gradient_penalty(self.discriminator, reals, fakes)

    def discriminator_loss(self, reals, fakes, real_scores, fake_scores) -> torch.Tensor:
        disc_loss = super().discriminator_loss(reals, fakes, real_scores, fake_scores)
        gp_term = self.hparams.gp_coefficient * self._gradient_penalty(reals, fakes)
        self.gp_term_metric(gp_term)
        disc_loss = disc_loss + gp_term
        # wgan.py:282-284: a [B] vector (norm over the last axis of [B, 1] scores) added to the scalar loss
        norm_term = self.hparams.e_drift * (fake_scores.norm(dim=-1) + real_scores.norm(dim=-1))
        self.norm_term_metric(norm_term)
        return disc_loss + norm_term
