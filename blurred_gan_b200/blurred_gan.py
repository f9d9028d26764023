"""``blurred_gan.py:17-52`` on torch modules: any GAN class with the blur layer in front of its critic.

``BlurredVariant(WGANGP)`` wraps the discriminator in ``Sequential([blur, discriminator])`` (``blurred_gan.py:30-35``),
so every critic evaluation -- on fakes, reals and gradient-penalty interpolates, forward, backward and double
backward -- goes through the blur; ``model.std`` is the layer's frozen ``std`` parameter that the sigma controllers
(``blurred_gan_b200.callbacks``) overwrite every batch.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from .wgan import WGAN, WGANGP, TrainingConfig, _Mean  # noqa: F401


def BlurredVariant(some_gan_base_class):
    class BlurredGAN(some_gan_base_class):
        @dataclass
        class HyperParameters(some_gan_base_class.HyperParameters):
            initial_blur_std: float = 0.05  # blurred_gan.py:27

        def __init__(self, generator, discriminator, hyperparams, config=None, blur=None, fused_critic=False, **kwargs):
            """``blur``: the blur layer; defaults to the B200 ``GaussianBlur2D`` (CUDA only).  Tests of the harness on CPU
            pass the oracle's CPU layer instead -- the package itself has no CPU path.

            ``fused_critic`` (not in the reference): the critic step blurs the generated and
            the real batch with ONE library call (``forward_many``: one launch for 128 x 128 planes) and, because the
            blur is linear, takes the gradient-penalty interpolates between the two *blurred* batches
            (``B x_hat = B reals + a (B fakes - B reals)``) instead of blurring ``x_hat`` a third time; the gradient
            w.r.t. the unblurred ``x_hat`` that the penalty needs is ``B`` applied to the gradient w.r.t. ``B x_hat``
            (``B`` is self-adjoint).  Per step: 6 blur applications in 5 launches instead of 7 in 7, the same losses up
            to fp32 rounding (tests/test_wgan_harness.py)."""
            if blur is None:
                from .gaussian_blur import GaussianBlur2D

                blur = GaussianBlur2D(initial_std=hyperparams.initial_blur_std)
            discriminator_with_blur = nn.Sequential(blur, discriminator)
            super().__init__(generator, discriminator_with_blur, hyperparams, config, **kwargs)
            self.blur = blur
            self.std_metric = _Mean("std")
            self.fused_critic = bool(fused_critic)
            self._blurred_pair = None  # (B reals, B fakes) of the current critic step when fused

        @property
        def std(self):
            return self.blur.std

        def assign_std(self, value: float) -> None:
            """``model.std.assign(value)`` (reference callbacks.py:62)."""
            for name in ("assign", "set_std"):
                fn = getattr(self.blur, name, None)
                if fn is not None:
                    fn(value)
                    return
            with torch.no_grad():  # (an in-place op on the Parameter itself: bumps its version, unlike `.data` writes)
                self.blur.std.fill_(float(value))

        def discriminator_step(self, reals):
            try:
                disc_loss, images = super().discriminator_step(reals)
            finally:
                self._blurred_pair = None
            self.std_metric(self.std)  # blurred_gan.py:47
            return disc_loss, images

        # ---- fused critic step (row N3) ---------------------------------------------------------------------------
        def _blur_many(self, *images):
            many = getattr(self.blur, "forward_many", None)
            return list(many(*images)) if many is not None else [self.blur(x) for x in images]

        def _critic_scores(self, fakes, reals):
            if not self.fused_critic:
                return super()._critic_scores(fakes, reals)
            blurred_fakes, blurred_reals = self._blur_many(fakes, reals)
            self._blurred_pair = (blurred_reals, blurred_fakes)
            critic = self.discriminator[1]
            critic.train(True)
            return critic(blurred_fakes), critic(blurred_reals)  # same order (and dropout draws) as the unfused step

        def _gradient_penalty(self, reals, fakes):
            if not self.fused_critic or self._blurred_pair is None:
                return super()._gradient_penalty(reals, fakes)
            blurred_reals, blurred_fakes = self._blurred_pair
            batch_size = reals.shape[0]
            a = torch.rand(batch_size, 1, 1, 1, device=reals.device, dtype=reals.dtype)
            if blurred_reals.is_cuda:
                from . import ops

                u_hat = ops.lerp_rows(blurred_reals, blurred_fakes, a.reshape(-1))
            else:  # harness tests on CPU (oracle blur layer)
                u_hat = (blurred_reals + a * (blurred_fakes - blurred_reals)).detach()
            u_hat.requires_grad_(True)
            critic = self.discriminator[1]
            critic.train(False)  # wgan.py:242
            y_hat = critic(u_hat)
            (grad_u,) = torch.autograd.grad(y_hat, u_hat, grad_outputs=torch.ones_like(y_hat), create_graph=True)
            grad = self.blur(grad_u)  # d/dx_hat = B^T d/d(B x_hat), B^T = B
            norm = grad.reshape(batch_size, -1).norm(dim=1)
            return ((norm - 1.0) ** 2).mean()

    BlurredGAN.__name__ = "Blurred" + some_gan_base_class.__name__
    return BlurredGAN


BlurredWGANGP = BlurredVariant(WGANGP)
BlurredWGAN = BlurredVariant(WGAN)
