"""PyTorch custom ops over the C ABI (include/bgan.h).  PyTorch is plumbing here: device memory,
the current stream and autograd bookkeeping; every FLOP runs in libbgan_b200.so.

``bgan::blur2d`` is differentiable to any order: the blur matrix is symmetric (zero "same" padding
applied after tap normalisation), so the vector-Jacobian product of the op is the op
itself applied to the incoming gradient.  That covers the 7 blur applications of one WGAN-GP step
(reference wgan.py:136-140,163-166,234-246): 4 forward, 2 backward, 1 double-backward.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import LAYOUT_NCHW, LAYOUT_NHWC, RULE_EXPLICIT, RULE_TF, RULE_TORCH  # noqa: F401


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            "%s: expected a CUDA tensor, got device %s. blurred_gan_b200 has no CPU path "
            "(the CPU restatement lives in oracle/ and is test infrastructure only)." % (what, t.device)
        )
    if t.dtype != torch.float32:
        raise RuntimeError("%s: expected float32, got %s" % (what, t.dtype))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device.  (The raw accessor costs ~1 us; building a
    torch.cuda.Stream object costs ~15 us, which showed up as a fifth of the host time of an SWD feed.)"""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _on_device:
    """`with torch.cuda.device(t.device)` that does nothing when t's device is already current (the common case)."""

    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


# --------------------------------------------------------------------------------------------
# blur
# --------------------------------------------------------------------------------------------
def _dims(x: torch.Tensor, layout: int) -> tuple[int, int, int, int]:
    if x.dim() != 4:
        raise RuntimeError("blur2d: expected a 4-D tensor, got shape %s" % (tuple(x.shape),))
    if layout == LAYOUT_NCHW:
        n, c, h, w = x.shape
    else:
        n, h, w, c = x.shape
    return n, c, h, w


@torch.library.custom_op("bgan::blur2d", mutates_args=(), device_types="cuda")
def blur2d(x: torch.Tensor, layout: int, rule: int, std_eff: float, radius: int) -> torch.Tensor:
    """y = B x with taps generated in-kernel from (rule, std_eff, radius)."""
    _require_cuda(x, "blur2d")
    lib = _lib.load()
    x = x.contiguous()
    n, c, h, w = _dims(x, layout)
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    with _on_device(x.device):
        ws_bytes = lib.bgan_blur2d_workspace_bytes(n, c, h, w, layout, radius)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
        rc = lib.bgan_blur2d_f32(x.data_ptr(), y.data_ptr(), n, c, h, w, layout, rule, float(std_eff), radius, None,
                                 ws.data_ptr() if ws is not None else None, ws_bytes, _stream_ptr())
    _lib.check(rc, "bgan_blur2d_f32")
    return y


@blur2d.register_fake
def _(x, layout, rule, std_eff, radius):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def _blur2d_backward(ctx, grad):
    return blur2d(grad, ctx.layout, ctx.rule, ctx.std_eff, ctx.radius), None, None, None, None


def _blur2d_setup(ctx, inputs, output):
    _, ctx.layout, ctx.rule, ctx.std_eff, ctx.radius = inputs


blur2d.register_autograd(_blur2d_backward, setup_context=_blur2d_setup)


MAX_MULTI = 4  # kMaxMulti of csrc/common.cuh: tensors per bgan_blur2d_multi_f32 call


@torch.library.custom_op("bgan::blur2d_multi", mutates_args=(), device_types="cuda")
def blur2d_multi(xs: Sequence[torch.Tensor], layout: int, rule: int, std_eff: float, radius: int) -> list[torch.Tensor]:
    """[B x for x in xs] for 1 .. 4 equally shaped tensors: one launch for 128 x 128 planes."""
    if not 1 <= len(xs) <= MAX_MULTI:
        raise RuntimeError("blur2d_multi: 1 .. %d tensors per call, got %d" % (MAX_MULTI, len(xs)))
    for x in xs:
        _require_cuda(x, "blur2d_multi")
        if x.shape != xs[0].shape or x.device != xs[0].device:
            raise RuntimeError("blur2d_multi: tensors must share shape and device")
    lib = _lib.load()
    xs = [x.contiguous() for x in xs]
    n, c, h, w = _dims(xs[0], layout)
    ys = [torch.empty_like(x) for x in xs]
    if xs[0].numel() == 0:
        return ys
    count = len(xs)
    xp = (ctypes.c_void_p * count)(*[x.data_ptr() for x in xs])
    yp = (ctypes.c_void_p * count)(*[y.data_ptr() for y in ys])
    with _on_device(xs[0].device):
        ws_bytes = lib.bgan_blur2d_workspace_bytes(n, c, h, w, layout, radius)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xs[0].device) if ws_bytes else None
        rc = lib.bgan_blur2d_multi_f32(xp, yp, count, n, c, h, w, layout, rule, float(std_eff), radius, None,
                                       ws.data_ptr() if ws is not None else None, ws_bytes, _stream_ptr())
    _lib.check(rc, "bgan_blur2d_multi_f32")
    return ys


@blur2d_multi.register_fake
def _(xs, layout, rule, std_eff, radius):
    return [torch.empty_like(x, memory_format=torch.contiguous_format) for x in xs]


def _blur2d_multi_backward(ctx, grads):
    # B is self-adjoint: the gradients of the outputs that have one go through the same (single) launch
    live = [i for i, g in enumerate(grads) if g is not None]
    out = [None] * len(grads)
    if live:
        for i, g in zip(live, blur2d_multi([grads[i] for i in live], ctx.layout, ctx.rule, ctx.std_eff, ctx.radius)):
            out[i] = g
    return out, None, None, None, None


def _blur2d_multi_setup(ctx, inputs, output):
    _, ctx.layout, ctx.rule, ctx.std_eff, ctx.radius = inputs


blur2d_multi.register_autograd(_blur2d_multi_backward, setup_context=_blur2d_multi_setup)


def lerp_rows(u: torch.Tensor, v: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """``u[i] + weights[i] * (v[i] - u[i])`` per leading index (no autograd: the gradient penalty detaches its interpolates,
    reference wgan.py:237-241)."""
    _require_cuda(u, "lerp_rows")
    _require_cuda(v, "lerp_rows")
    _require_cuda(weights, "lerp_rows(weights)")
    if u.shape != v.shape or weights.numel() != u.shape[0]:
        raise RuntimeError("lerp_rows: u and v must share a shape and weights must have one entry per leading index")
    u, v, weights = u.detach().contiguous(), v.detach().contiguous(), weights.detach().contiguous()
    out = torch.empty_like(u)
    rows = u.shape[0]
    if u.numel():
        with _on_device(u.device):
            rc = _lib.load().bgan_lerp_rows_f32(u.data_ptr(), v.data_ptr(), weights.data_ptr(), out.data_ptr(), rows, u.numel() // rows,
                                                _stream_ptr())
        _lib.check(rc, "bgan_lerp_rows_f32")
    return out


@torch.library.custom_op("bgan::blur2d_taps", mutates_args=(), device_types="cuda")
def blur2d_taps(x: torch.Tensor, taps: torch.Tensor, layout: int) -> torch.Tensor:
    """y = B x with caller-supplied taps (1-D, odd length, used as given -- not renormalised)."""
    _require_cuda(x, "blur2d_taps")
    _require_cuda(taps, "blur2d_taps(taps)")
    if taps.dim() != 1 or taps.numel() % 2 != 1:
        raise RuntimeError("blur2d_taps: taps must be 1-D with odd length")
    lib = _lib.load()
    x = x.contiguous()
    taps = taps.contiguous()
    n, c, h, w = _dims(x, layout)
    radius = taps.numel() // 2
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    with _on_device(x.device):
        ws_bytes = lib.bgan_blur2d_workspace_bytes(n, c, h, w, layout, radius)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
        rc = lib.bgan_blur2d_f32(x.data_ptr(), y.data_ptr(), n, c, h, w, layout, RULE_EXPLICIT, 1.0, radius, taps.data_ptr(),
                                 ws.data_ptr() if ws is not None else None, ws_bytes, _stream_ptr())
    _lib.check(rc, "bgan_blur2d_f32")
    return y


@blur2d_taps.register_fake
def _(x, taps, layout):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def _blur2d_taps_backward(ctx, grad):
    (taps,) = ctx.saved_tensors
    # correlation with taps t is adjoint to correlation with reversed taps; symmetric taps are their own reverse
    return blur2d_taps(grad, torch.flip(taps, dims=(0,)), ctx.layout), None, None


def _blur2d_taps_setup(ctx, inputs, output):
    _, taps, ctx.layout = inputs
    ctx.save_for_backward(taps)


blur2d_taps.register_autograd(_blur2d_taps_backward, setup_context=_blur2d_taps_setup)


def device_taps(rule: int, std_eff: float, radius: int, device) -> torch.Tensor:
    """The 2*radius+1 taps exactly as the blur kernels compute them (for parity tests / `.kernel`)."""
    lib = _lib.load()
    out = torch.empty(2 * radius + 1, dtype=torch.float32, device=device)
    with _on_device(out.device):
        _lib.check(lib.bgan_blur_taps_f32(rule, float(std_eff), radius, out.data_ptr(), _stream_ptr()), "bgan_blur_taps_f32")
    return out


def last_blur_path() -> int:
    return _lib.load().bgan_blur2d_last_path()


# --------------------------------------------------------------------------------------------
# SWD building blocks (no autograd: the metric is evaluation-only, reference callbacks.py:153-179)
# --------------------------------------------------------------------------------------------
def pyr_down(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, "pyr_down")
    assert x.dim() == 4
    x = x.contiguous()
    n, c, h, w = x.shape
    y = torch.empty((n, c, (h + 1) // 2, (w + 1) // 2), dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        _lib.check(_lib.load().bgan_pyr_down_f32(x.data_ptr(), y.data_ptr(), n * c, h, w, _stream_ptr()), "bgan_pyr_down_f32")
    return y


def pyr_up(x: torch.Tensor, minuend: torch.Tensor | None = None) -> torch.Tensor:
    """pyr_up(x), or ``minuend - pyr_up(x)`` fused when ``minuend`` is given."""
    _require_cuda(x, "pyr_up")
    assert x.dim() == 4
    x = x.contiguous()
    n, c, h, w = x.shape
    y = torch.empty((n, c, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
    mptr = None
    if minuend is not None:
        _require_cuda(minuend, "pyr_up(minuend)")
        assert tuple(minuend.shape) == tuple(y.shape)
        minuend = minuend.contiguous()
        mptr = minuend.data_ptr()
    with _on_device(x.device):
        _lib.check(_lib.load().bgan_pyr_up_f32(x.data_ptr(), mptr, y.data_ptr(), n * c, h, w, _stream_ptr()), "bgan_pyr_up_f32")
    return y


def swd_extract(level: torch.Tensor, px, py, nhoods_per_image: int, nhood: int, sums: torch.Tensor) -> torch.Tensor:
    """Gather descriptors [n*per_image, C, nhood, nhood] and accumulate channel statistics into ``sums`` (fp64 [2C])."""
    _require_cuda(level, "swd_extract")
    level = level.contiguous()
    n, c, h, w = level.shape
    dev = level.device
    px_t = torch.as_tensor(np.ascontiguousarray(px, dtype=np.int32)).to(dev) if not torch.is_tensor(px) else px.to(dev, torch.int32)
    py_t = torch.as_tensor(np.ascontiguousarray(py, dtype=np.int32)).to(dev) if not torch.is_tensor(py) else py.to(dev, torch.int32)
    px_t, py_t = px_t.contiguous(), py_t.contiguous()
    assert px_t.numel() == n * nhoods_per_image and py_t.numel() == n * nhoods_per_image
    assert sums.dtype == torch.float64 and sums.numel() == 2 * c and sums.is_cuda and sums.is_contiguous()
    out = torch.empty((n * nhoods_per_image, c, nhood, nhood), dtype=torch.float32, device=dev)
    with _on_device(dev):
        rc = _lib.load().bgan_swd_extract_f32(level.data_ptr(), n, c, h, w, px_t.data_ptr(), py_t.data_ptr(), nhoods_per_image,
                                              nhood, out.data_ptr(), sums.data_ptr(), _stream_ptr())
    _lib.check(rc, "bgan_swd_extract_f32")
    return out


def swd_feed(minibatch: torch.Tensor, positions: torch.Tensor, levels: int, nhoods_per_image: int, nhood: int, sums: list,
             scratch: dict | None = None, outs: list | None = None) -> list:
    """One ``API.feed``: Laplacian pyramid + descriptor gather of every level through ONE library call.

    ``positions``: int32 CUDA tensor [levels][x | y][n * nhoods_per_image]; ``sums``: per level an fp64 CUDA tensor (>= 2C
    entries) that the channel statistics are accumulated into; ``scratch``: dict keeping the pyramid workspace between
    calls; ``outs``: optional per-level destinations (contiguous [n * nhoods_per_image, C, nhood, nhood] views, e.g. slices of a
    growing per-level array).  Returns the per-level descriptor tensors [n * nhoods_per_image, C, nhood, nhood]."""
    _require_cuda(minibatch, "swd_feed")
    minibatch = minibatch.contiguous()
    n, c, h, w = minibatch.shape
    dev = minibatch.device
    count = n * nhoods_per_image
    assert positions.is_cuda and positions.dtype == torch.int32 and positions.is_contiguous() and positions.numel() == 2 * levels * count
    assert len(sums) == levels and all(t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() >= 2 * c for t in sums)
    lib = _lib.load()
    ws_bytes = lib.bgan_swd_feed_workspace_bytes(n, c, h, w, levels)
    ws = scratch.get("feed_ws") if scratch is not None else None
    if ws is None or ws.numel() < ws_bytes or ws.device != dev:
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        if scratch is not None:
            scratch["feed_ws"] = ws
    if outs is None:
        outs = [torch.empty((count, c, nhood, nhood), dtype=torch.float32, device=dev) for _ in range(levels)]
    else:
        assert len(outs) == levels and all(t.shape == (count, c, nhood, nhood) and t.dtype == torch.float32 and t.is_contiguous() and
                                           t.device == dev for t in outs)
    out_p = (ctypes.c_void_p * levels)(*[t.data_ptr() for t in outs])
    sum_p = (ctypes.c_void_p * levels)(*[t.data_ptr() for t in sums])
    with _on_device(dev):
        rc = lib.bgan_swd_feed_f32(minibatch.data_ptr(), n, c, h, w, levels, positions.data_ptr(), nhoods_per_image, nhood, out_p, sum_p,
                                   ws.data_ptr(), ws_bytes, _stream_ptr())
    _lib.check(rc, "bgan_swd_feed_f32")
    if scratch is None:
        ws.record_stream(torch.cuda.current_stream(dev))
    return outs


def swd_normalise_(desc: torch.Tensor, sums: torch.Tensor, count: float) -> torch.Tensor:
    """In place (desc - mean_c) / std_c for desc [rows, C, a, b]; mean/std derived on device from ``sums``."""
    _require_cuda(desc, "swd_normalise_")
    assert desc.is_contiguous() and desc.dim() == 4
    rows, c, a, b = desc.shape
    with _on_device(desc.device):
        rc = _lib.load().bgan_swd_normalise_f32(desc.data_ptr(), rows, c, a * b, sums.data_ptr(), float(count), _stream_ptr())
    _lib.check(rc, "bgan_swd_normalise_f32")
    return desc


def swd_project(desc: torch.Tensor, dirs: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """desc [rows, k] @ dirs [k, d] -> direction-major [d, rows] (into ``out`` when given)."""
    _require_cuda(desc, "swd_project")
    _require_cuda(dirs, "swd_project(dirs)")
    desc = desc.contiguous()
    dirs = dirs.contiguous()
    rows, k = desc.shape
    assert dirs.shape[0] == k
    d = dirs.shape[1]
    if out is None:
        out = torch.empty((d, rows), dtype=torch.float32, device=desc.device)
    assert out.shape == (d, rows) and out.dtype == torch.float32 and out.is_contiguous() and out.device == desc.device
    with _on_device(desc.device):
        rc = _lib.load().bgan_swd_project_f32(desc.data_ptr(), rows, k, dirs.data_ptr(), d, out.data_ptr(), _stream_ptr())
    _lib.check(rc, "bgan_swd_project_f32")
    return out


def swd_sort_(keys: torch.Tensor, scratch: dict | None = None) -> torch.Tensor:
    """Ascending in-place sort of every row of keys [segments, rows].  ``scratch``: a dict in which the workspace tensor is
    kept between calls (an evaluation sorts 32 equally shaped arrays)."""
    _require_cuda(keys, "swd_sort_")
    assert keys.is_contiguous() and keys.dim() == 2
    segments, rows = keys.shape
    lib = _lib.load()
    with _on_device(keys.device):
        done = 0
        while done < segments:  # the kernel grid caps segments per launch
            seg = min(segments - done, 32768)
            ws_bytes = lib.bgan_swd_sort_workspace_bytes(rows, seg)
            ws = scratch.get("sort_ws") if scratch is not None else None
            if ws is None or ws.numel() < ws_bytes or ws.device != keys.device:
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
                if scratch is not None:
                    scratch["sort_ws"] = ws
            part = keys[done : done + seg]
            _lib.check(lib.bgan_swd_sort_f32(part.data_ptr(), rows, seg, ws.data_ptr(), ws_bytes, _stream_ptr()), "bgan_swd_sort_f32")
            done += seg
    return keys


def swd_absdiff_sum(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sum |a - b| as a 1-element fp64 device tensor."""
    _require_cuda(a, "swd_absdiff_sum")
    _require_cuda(b, "swd_absdiff_sum")
    assert a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel()
    out = torch.zeros(1, dtype=torch.float64, device=a.device)
    with _on_device(a.device):
        rc = _lib.load().bgan_swd_absdiff_sum_f32(a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), _stream_ptr())
    _lib.check(rc, "bgan_swd_absdiff_sum_f32")
    return out


def swd_sort_diff_(a: torch.Tensor, b: torch.Tensor, sums: torch.Tensor, seg_offset: int, segs_per_sum: int,
                   scratch: dict | None = None) -> torch.Tensor:
    """``sums[(s + seg_offset) // segs_per_sum] += sum |sort(a[s]) - sort(b[s])|`` for every row s of a, b [segments, rows]
    through ONE library call (``bgan_swd_sort_diff_f32``): a is sorted in place, b is clobbered, the sorted b is never
    stored.  ``sums``: fp64 CUDA tensor.  Returns ``sums``."""
    _require_cuda(a, "swd_sort_diff_")
    _require_cuda(b, "swd_sort_diff_")
    assert a.is_contiguous() and b.is_contiguous() and a.dim() == 2 and a.shape == b.shape and a.device == b.device
    assert sums.is_cuda and sums.dtype == torch.float64 and sums.is_contiguous() and sums.device == a.device
    segments, rows = a.shape
    n_sums = sums.numel()
    lib = _lib.load()
    with _on_device(a.device):
        done = 0
        while done < segments:  # the kernel grid caps segments per launch
            seg = min(segments - done, 32768)
            ws_bytes = lib.bgan_swd_sort_diff_workspace_bytes(rows, seg, n_sums)
            ws = scratch.get("sort_ws") if scratch is not None else None
            if ws is None or ws.numel() < ws_bytes or ws.device != a.device:
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
                if scratch is not None:
                    scratch["sort_ws"] = ws
            rc = lib.bgan_swd_sort_diff_f32(a[done : done + seg].data_ptr(), b[done : done + seg].data_ptr(), rows, seg, seg_offset + done,
                                            segs_per_sum, n_sums, sums.data_ptr(), ws.data_ptr(), ws_bytes, _stream_ptr())
            _lib.check(rc, "bgan_swd_sort_diff_f32")
            if scratch is None:
                ws.record_stream(torch.cuda.current_stream(a.device))
            done += seg
    return sums
