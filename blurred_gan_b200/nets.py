"""The reference's DCGAN generator / critic definitions as torch modules (NCHW), for the train-step harness only.

``demo_mnist.py:48-86`` (28 x 28 x 1) and ``demo_celeba.py:51-124`` (128 x 128 x 3).  Keras defaults that differ from
torch's are spelled out: ``LeakyReLU`` slope 0.3, ``BatchNormalization`` momentum 0.99 / eps 1e-3, ``padding="same"``
transposed convolutions (k = 5, stride 2 -> padding 2, output_padding 1).  Library arithmetic (cuDNN), not a target.
"""
from __future__ import annotations

import torch.nn as nn


def _bn(c, dims=2):
    cls = nn.BatchNorm2d if dims == 2 else nn.BatchNorm1d
    return cls(c, eps=1e-3, momentum=0.01)


def _up(cin, cout, stride):
    return nn.ConvTranspose2d(cin, cout, 5, stride=stride, padding=2, output_padding=stride - 1, bias=False)


class MnistGenerator(nn.Sequential):
    """``demo_mnist.py:48-71``."""

    def __init__(self, latent_size: int = 100):
        super().__init__(
            nn.Linear(latent_size, 7 * 7 * 256, bias=False), _bn(7 * 7 * 256, 1), nn.LeakyReLU(0.3), nn.Unflatten(1, (256, 7, 7)),
            _up(256, 128, 1), _bn(128), nn.LeakyReLU(0.3),
            _up(128, 64, 2), _bn(64), nn.LeakyReLU(0.3),
            _up(64, 1, 2), nn.Tanh())
        self.latent_size = latent_size


class MnistDiscriminator(nn.Sequential):
    """``demo_mnist.py:74-86``."""

    def __init__(self):
        super().__init__(
            nn.Conv2d(1, 64, 5, stride=2, padding=2), nn.LeakyReLU(0.3), nn.Dropout(0.3),
            nn.Conv2d(64, 128, 5, stride=2, padding=2), nn.LeakyReLU(0.3), nn.Dropout(0.3),
            nn.Flatten(), nn.Linear(7 * 7 * 128, 1))


class CelebaGenerator(nn.Sequential):
    """``demo_celeba.py:51-93``: Dense 4*4*512 -> six transposed convolutions -> Conv 3, tanh."""

    def __init__(self, latent_size: int = 100):
        layers = [nn.Linear(latent_size, 4 * 4 * 512, bias=False), _bn(4 * 4 * 512, 1), nn.LeakyReLU(0.3), nn.Unflatten(1, (512, 4, 4)),
                  _up(512, 512, 1), _bn(512), nn.LeakyReLU(0.3)]
        for cin, cout in ((512, 256), (256, 128), (128, 64), (64, 32), (32, 16)):
            layers += [_up(cin, cout, 2), _bn(cout), nn.LeakyReLU(0.3)]
        layers += [nn.Conv2d(16, 3, 5, padding=2, bias=False), nn.Tanh()]
        super().__init__(*layers)
        self.latent_size = latent_size


class CelebaDiscriminator(nn.Sequential):
    """``demo_celeba.py:96-124``: six stride-2 convolutions with dropout, Dense 1."""

    def __init__(self):
        layers, cin = [], 3
        for cout in (16, 32, 64, 128, 256, 512):
            layers += [nn.Conv2d(cin, cout, 5, stride=2, padding=2), nn.LeakyReLU(0.3), nn.Dropout(0.3)]
            cin = cout
        layers += [nn.Flatten(), nn.Linear(2 * 2 * 512, 1)]
        super().__init__(*layers)
