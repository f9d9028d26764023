"""Drop-in for the reference's ``pytorch/gaussianblur.py`` (lines 9-71), minus its script tail.

Same names, same arguments, same results -- including the reference's quirks, which parity is
defined against:

* the Gaussian normaliser sits inside the exponential (``pytorch/gaussianblur.py:11``);
* ``kernel_size`` is ``max(3, odd_up(ceil(5 std)))`` clipped to ``odd_up(min(H, W)) - 2`` (``:17-30,43-57``),
  and the clip uses the image size remembered from the *previous* forward until the next one (``:37,61``);
* zero "same" padding after normalisation, horizontal pass then vertical pass (``:67-70``).

Extensions over the reference (which is single-channel and CPU-only): any number of
channels (depthwise, same taps), CUDA tensors, autograd to any order.  ``forward`` runs the fused
sm_100a kernel; there is no CPU path.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from .. import ops


def gaussian_kernel_1d(std, kernel_size: int) -> torch.Tensor:
    """Taps of the reference, shape [1, kernel_size], fp32 (``pytorch/gaussianblur.py:9-14``).

    Host-side helper kept for API parity (``GaussianBlur.kernel``); the blur kernels generate the
    same taps on the device.
    """
    std_t = torch.as_tensor(std, dtype=torch.float32).detach().reshape(-1)[:1].cpu()
    half = kernel_size // 2
    offsets = torch.arange(-half, half + 1, dtype=torch.float32).unsqueeze(0)
    exponent = (offsets.square() / (2 * std_t.square())) / (math.sqrt(2 * math.pi) * std_t)
    g = torch.exp(-exponent)
    return g / g.sum()


def nearest_odd_integer(number, higher: bool = True) -> int:
    """``pytorch/gaussianblur.py:17-20``."""
    integer = int(np.ceil(float(number)))
    odd = integer if integer % 2 == 1 else integer + 1
    return odd if higher else odd - 2


def appropriate_kernel_size(std) -> int:
    """``pytorch/gaussianblur.py:23-30``: nearest odd integer above 5 std (fp32 product), at least 3."""
    five_std = float(np.float32(5.0) * np.float32(float(std)))
    return max(nearest_odd_integer(five_std), 3)


class GaussianBlur(nn.Module):
    def __init__(self, std: float = 2.0):
        super().__init__()
        # frozen Parameter named `std`, shape [1]: part of state_dict, assigned by the sigma controllers.
        # It stays in HOST memory whatever device the module is moved to (`_apply` below): the kernel-size rule
        # (`:43-57`) needs sigma on the host before every launch -- radius and kernel choice are launch parameters --
        # so a device-resident std would cost a device->host synchronisation per change, and a cached host copy goes
        # stale when the Parameter is written through `.data` (which does not bump `_version`).  Reading a CPU scalar is
        # free and sees every write, whichever way it was made.
        self.std = nn.Parameter(torch.Tensor([std]), requires_grad=False)
        self._image_size = None

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

    def set_std(self, value: float) -> None:
        """Assign a new std (what a sigma controller calls every batch; reference ``callbacks.py:60-62``)."""
        with torch.no_grad():
            self.std.fill_(float(value))

    assign = set_std  # the name the reference's controllers use (``model.std.assign``)

    @property
    def kernel(self) -> torch.Tensor:
        return gaussian_kernel_1d(std=self.std_value(), kernel_size=self.kernel_size).to(self.std.device)

    @property
    def kernel_size(self) -> int:
        k_size = appropriate_kernel_size(self.std_value())
        if self._image_size:
            k_size = min(k_size, nearest_odd_integer(self._image_size, higher=False))
        assert k_size % 2 == 1, "kernel size should be odd"
        return k_size

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if inputs.dim() != 4:
            raise RuntimeError("GaussianBlur expects [N, C, H, W], got %s" % (tuple(inputs.shape),))
        if not inputs.is_cuda:
            raise RuntimeError("GaussianBlur (blurred_gan_b200) runs on CUDA tensors only; there is no CPU fallback")
        self._image_size = min(inputs.shape[-2:])
        k = self.kernel_size
        if k < 1:
            raise AssertionError("kernel size should be odd")
        return ops.blur2d(inputs, ops.LAYOUT_NCHW, ops.RULE_TORCH, self.std_value(), k // 2)

    def forward_many(self, *inputs: torch.Tensor) -> list:
        """``[self(x) for x in inputs]`` for 1 .. 4 equally shaped batches through one library call (one launch for
        128 x 128 planes).  Not in the reference; serves a critic step that blurs real and generated batches with the
        same sigma."""
        for x in inputs:
            if x.dim() != 4:
                raise RuntimeError("GaussianBlur expects [N, C, H, W], got %s" % (tuple(x.shape),))
            if not x.is_cuda:
                raise RuntimeError("GaussianBlur (blurred_gan_b200) runs on CUDA tensors only; there is no CPU fallback")
        self._image_size = min(inputs[0].shape[-2:])
        k = self.kernel_size
        if k < 1:
            raise AssertionError("kernel size should be odd")
        return list(ops.blur2d_multi(list(inputs), ops.LAYOUT_NCHW, ops.RULE_TORCH, self.std_value(), k // 2))
