"""ctypes binding of the C ABI declared in include/bgan.h.

There is no CPU fallback: if the shared library is missing the first call raises, loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p, POINTER

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libbgan_b200.so")
ABI_VERSION = 4

RULE_TORCH, RULE_TF, RULE_EXPLICIT = 0, 1, 2
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1

# name -> (restype, argtypes); mirrors include/bgan.h one to one (tests check the two agree)
SIGNATURES = {
    "bgan_version": (c_int, []),
    "bgan_last_error": (c_char_p, []),
    "bgan_blur_plan": (c_int, [c_int, c_float, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_float)]),
    "bgan_blur_taps_f32": (c_int, [c_int, c_float, c_int, c_void_p, c_void_p]),
    "bgan_blur2d_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "bgan_blur2d_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "bgan_blur2d_multi_f32": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "bgan_lerp_rows_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "bgan_blur2d_last_path": (c_int, []),
    "bgan_pyr_down_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "bgan_pyr_up_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "bgan_swd_extract_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                     c_void_p, c_void_p]),
    "bgan_swd_feed_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "bgan_swd_feed_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, POINTER(c_void_p), POINTER(c_void_p),
                                  c_void_p, c_size_t, c_void_p]),
    "bgan_swd_normalise_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_double, c_void_p]),
    "bgan_swd_project_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "bgan_swd_project_last_path": (c_int, []),
    "bgan_swd_sort_last_path": (c_int, []),
    "bgan_swd_sort_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "bgan_swd_sort_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "bgan_swd_absdiff_sum_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "bgan_swd_sort_diff_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "bgan_swd_sort_diff_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "bgan_mt19937_randint_i32": (c_int, [c_void_p, POINTER(c_int), c_int64, c_int64, c_int64, c_void_p]),
    "bgan_swd_positions_i32": (c_int, [c_void_p, POINTER(c_int), c_int, POINTER(c_int), POINTER(c_int), c_int, c_int64, c_int, c_int,
                                       c_void_p]),
}

_lib = None


class BganError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libbgan_b200.so (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BganError(
            "blurred_gan_b200: %s is missing. Build it with `python -m blurred_gan_b200._build` "
            "(needs nvcc, targets sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.bgan_version()
    if got != ABI_VERSION:
        raise BganError("libbgan_b200.so ABI version %d, binding expects %d; rebuild" % (got, ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().bgan_last_error()
        raise BganError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def blur_plan(rule: int, std: float, height: int, width: int) -> tuple[int, int, float]:
    """(kernel_size, radius, std_eff) from the library's host-side rule (no GPU needed)."""
    k, r, s = c_int(), c_int(), c_float()
    check(load().bgan_blur_plan(rule, std, height, width, ctypes.byref(k), ctypes.byref(r), ctypes.byref(s)), "bgan_blur_plan")
    return k.value, r.value, s.value
