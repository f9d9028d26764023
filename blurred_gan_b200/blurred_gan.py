"""``blurred_gan.py:17-52`` on torch modules: any GAN class with the blur layer in front of its critic.

``BlurredVariant(WGANGP)`` wraps the discriminator in ``Sequential([blur, discriminator])`` (``blurred_gan.py:30-35``),
so every critic evaluation -- on fakes, reals and gradient-penalty interpolates, forward, backward and double
backward -- goes through the blur; ``model.std`` is the layer's frozen ``std`` parameter that the sigma controllers
(``blurred_gan_b200.callbacks``) overwrite every batch.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch.nn as nn

from .wgan import WGAN, WGANGP, TrainingConfig, _Mean  # noqa: F401


def BlurredVariant(some_gan_base_class):
    class BlurredGAN(some_gan_base_class):
        @dataclass
        class HyperParameters(some_gan_base_class.HyperParameters):
            initial_blur_std: float = 0.05  # blurred_gan.py:27

        def __init__(self, generator, discriminator, hyperparams, config=None, blur=None, **kwargs):
            """``blur``: the blur layer; defaults to the B200 ``GaussianBlur2D`` (CUDA only).  Tests of the harness on CPU
            pass the oracle's CPU layer instead -- the package itself has no CPU path."""
            if blur is None:
                from .gaussian_blur import GaussianBlur2D

                blur = GaussianBlur2D(initial_std=hyperparams.initial_blur_std)
            discriminator_with_blur = nn.Sequential(blur, discriminator)
            super().__init__(generator, discriminator_with_blur, hyperparams, config, **kwargs)
            self.blur = blur
            self.std_metric = _Mean("std")

        @property
        def std(self):
            return self.blur.std

        def assign_std(self, value: float) -> None:
            """``model.std.assign(value)`` (reference callbacks.py:62) without a device round trip."""
            if hasattr(self.blur, "assign"):
                self.blur.assign(value)
            else:
                self.blur.std.data.fill_(float(value))

        def discriminator_step(self, reals):
            disc_loss, images = super().discriminator_step(reals)
            self.std_metric(self.std)  # blurred_gan.py:47
            return disc_loss, images

    BlurredGAN.__name__ = "Blurred" + some_gan_base_class.__name__
    return BlurredGAN


BlurredWGANGP = BlurredVariant(WGANGP)
BlurredWGAN = BlurredVariant(WGAN)
