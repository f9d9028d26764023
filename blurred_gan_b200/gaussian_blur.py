"""Drop-in for the reference's ``gaussian_blur.py`` (TensorFlow 2 / Keras), re-hosted on torch tensors.

Same function names and argument meaning; tensors are CUDA ``torch.Tensor`` instead of ``tf.Tensor``
and the Keras layer becomes an ``nn.Module``.  Rules replicated from the reference:

* ``appropriate_kernel_size(std) = floor(6 std) + 1``            (``gaussian_blur.py:21-26`` -- not "nearest odd")
* ``blur_images`` clips it to ``[3, max(H, W)]`` and re-derives ``std = (k - 1) / 6`` (``:58-72``)
* taps run over ``[-(k // 2), k // 2]`` -- always an odd count -- and are a true Gaussian (``:83-88``)
* layout autodetect: last dim in {1, 3} -> NHWC, else NCHW (``:34-39``)
* depthwise, zero "SAME" padding (``:116-130``)

The reference runs the vertical pass first; the fused kernel runs the horizontal one first, which
only changes fp32 rounding order (measured <= 2e-7, far inside the 1e-5 parity bar).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import ops

_F32 = np.float32


def _as_f32(value) -> np.float32:
    if torch.is_tensor(value):
        return _F32(float(value.detach().reshape(-1)[0].item()))
    return _F32(float(value))


def maximum_reasonable_std(image_resolution: int) -> float:
    """``gaussian_blur.py:15-18``."""
    return appropriate_std(image_resolution - 1)


def appropriate_kernel_size(std):
    """``gaussian_blur.py:21-26``: ``(6 * std) * 2 // 2 + 1`` evaluated in fp32 (a float, like the reference)."""
    return float(_F32(np.floor((_F32(6.0) * _as_f32(std)) * _F32(2.0) / _F32(2.0))) + _F32(1.0))


def appropriate_std(kernel_size):
    """``gaussian_blur.py:29-31``."""
    return float((_F32(kernel_size) - _F32(1.0)) / _F32(6.0))


def get_data_format(image) -> str:
    """``gaussian_blur.py:34-39``."""
    return "NHWC" if image.shape[-1] in (1, 3) else "NCHW"


def get_image_dims(image):
    """``gaussian_blur.py:42-47`` -> (height, width, channels)."""
    if get_data_format(image) == "NHWC":
        return image.shape[1], image.shape[2], image.shape[-1]
    return image.shape[2], image.shape[-1], image.shape[1]


def gaussian_kernel_1d(std, kernel_size) -> torch.Tensor:
    """``gaussian_blur.py:83-88``: normalised Gaussian over ``[-(k // 2), k // 2]``, fp32 (host helper)."""
    half = int(kernel_size) // 2
    s = torch.tensor(float(_as_f32(std)), dtype=torch.float32)
    x = torch.arange(-half, half + 1, dtype=torch.float32)
    g = torch.exp(-(x.square() / (2 * s.square()))) / (torch.sqrt(torch.tensor(2 * math.pi, dtype=torch.float32)) * s)
    return g / g.sum()


def gaussian_blur(image: torch.Tensor, std, kernel_size) -> torch.Tensor:
    """``gaussian_blur.py:91-132``: separable depthwise blur with a Gaussian of ``std`` truncated to ``kernel_size``."""
    if not torch.is_tensor(image) or image.dim() != 4:
        raise RuntimeError("gaussian_blur expects a 4-D tensor")
    if not image.is_cuda:
        raise RuntimeError("gaussian_blur (blurred_gan_b200) runs on CUDA tensors only; there is no CPU fallback")
    data_format = get_data_format(image)
    assert data_format in {"NHWC", "NCHW"}, "invalid data format"
    layout = ops.LAYOUT_NHWC if data_format == "NHWC" else ops.LAYOUT_NCHW
    return ops.blur2d(image, layout, ops.RULE_TF, float(_as_f32(std)), int(kernel_size) // 2)


def blur_images(images: torch.Tensor, scale) -> torch.Tensor:
    """``gaussian_blur.py:50-80``: pick the kernel size for ``scale``, clip it to the image, re-derive std, blur."""
    h, w, _ = get_image_dims(images)
    full_resolution = _F32(max(h, w))
    kernel_size = _F32(appropriate_kernel_size(scale))
    kernel_size = min(max(kernel_size, _F32(3.0)), full_resolution)
    std = max(_F32(appropriate_std(kernel_size)), _F32(0.01))
    return gaussian_blur(images, std, kernel_size)


def blur_images_many(images, scale) -> list:
    """``[blur_images(x, scale) for x in images]`` for 1 .. 4 equally shaped batches through ONE call of the library
    (one launch for 128 x 128 planes).  Not in the reference: its critic step blurs the generated and the real batch back
    to back with the same sigma (``wgan.py:136-137`` through ``blurred_gan.py:30-35``); SURVEY section 8f, row N3."""
    images = list(images)
    for x in images:
        if not torch.is_tensor(x) or x.dim() != 4:
            raise RuntimeError("blur_images_many expects 4-D tensors")
        if not x.is_cuda:
            raise RuntimeError("blur_images_many (blurred_gan_b200) runs on CUDA tensors only; there is no CPU fallback")
    h, w, _ = get_image_dims(images[0])
    kernel_size = _F32(appropriate_kernel_size(scale))
    kernel_size = min(max(kernel_size, _F32(3.0)), _F32(max(h, w)))
    std = max(_F32(appropriate_std(kernel_size)), _F32(0.01))
    data_format = get_data_format(images[0])
    layout = ops.LAYOUT_NHWC if data_format == "NHWC" else ops.LAYOUT_NCHW
    return list(ops.blur2d_multi(images, layout, ops.RULE_TF, float(_as_f32(std)), int(kernel_size) // 2))


class GaussianBlur2D(nn.Module):
    """``gaussian_blur.py:135-148``.  ``std`` is a non-trainable scalar that sigma controllers overwrite
    every batch (reference callbacks.py:60-62); it is kept as a frozen Parameter so that it
    checkpoints with the model like the reference's ``tf.Variable``.  The reference's before/after
    image summaries (``:145-147``) are logging, not arithmetic, and are not emitted."""

    def __init__(self, initial_std: float = 0.01, *args, **kwargs):
        super().__init__()
        # Kept in HOST memory whatever device the layer is moved to (`_apply` below): the kernel-size rule needs sigma
        # on the host before every launch, a CPU scalar is free to read and every write to it -- `assign`, `fill_`,
        # `.data.fill_`, `load_state_dict` -- is seen by the next forward (no cached copy that could go stale).
        self.std = nn.Parameter(torch.tensor(float(initial_std), dtype=torch.float32), requires_grad=False)

    def _apply(self, fn, recurse=True):
        std = self._parameters.pop("std")
        try:
            super()._apply(fn, recurse)
        finally:
            self._parameters["std"] = std
        return self

    def std_value(self) -> float:
        """Current sigma as a Python float.  No device synchronisation unless a caller replaced `std` by a CUDA tensor."""
        return float(self.std.detach().reshape(-1)[0])

    def assign(self, value: float) -> None:
        """``model.std.assign(value)`` of the reference's controllers (``callbacks.py:60-62``)."""
        with torch.no_grad():
            self.std.fill_(float(value))

    set_std = assign

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        return blur_images(image, self.std_value())

    def forward_many(self, *images: torch.Tensor) -> list:
        """The layer applied to several equally shaped batches in one library call (see ``blur_images_many``)."""
        return blur_images_many(images, self.std_value())
