"""CPU: the C-ABI library loads without a GPU, exports exactly what include/bgan.h declares, and its
host-only entry points (kernel-size rules, argument validation) agree with the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from blurred_gan_b200 import _build, _lib
from oracle import blur_oracle as bo


@pytest.fixture(scope="module")
def lib():
    _build.build_library()
    return _lib.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "bgan.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bgan_[a-z0-9_]+)\s*\(", text)))


def test_header_binding_and_exports_agree(lib):
    declared = _declared_functions()
    assert declared == sorted(_lib.SIGNATURES), "include/bgan.h and _lib.SIGNATURES differ"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\bT (bgan_[a-z0-9_]+)", out)))
    assert exported == declared, "exports of libbgan_b200.so differ from include/bgan.h"


def test_version_and_plain_c_linkage(lib):
    assert lib.bgan_version() == _lib.ABI_VERSION
    # the boundary is a C ABI: no libtorch / libc10 / libpython dependency
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd and "python" not in ldd


@pytest.mark.parametrize("size", [(128, 128), (28, 28), (1024, 1024), (37, 50), (16, 300), (64, 64)])
def test_blur_plan_matches_oracle(lib, size):
    h, w = size
    for std in np.concatenate([np.linspace(0.01, 3, 60), np.linspace(3, 70, 135)]):
        std = float(np.float32(std))
        k, r, s = _lib.blur_plan(_lib.RULE_TORCH, std, h, w)
        assert (k, r) == bo.torch_rule_plan(std, h, w)
        assert s == pytest.approx(std)
        k, r, s = _lib.blur_plan(_lib.RULE_TF, std, h, w)
        ko, ro, so_ = bo.tf_rule_plan(std, h, w)
        assert (k, r) == (ko, ro)
        assert s == pytest.approx(so_, rel=1e-7)


def test_blur_plan_matches_reference_table(lib, golden_taps):
    for std, h, w, k in golden_taps["rows"]:
        assert _lib.blur_plan(_lib.RULE_TORCH, float(std), int(h), int(w))[0] == int(k)


def test_argument_errors_are_reported_without_a_gpu(lib):
    k = ctypes.c_int()
    assert lib.bgan_blur_plan(7, 1.0, 8, 8, ctypes.byref(k), None, None) < 0
    assert b"rule" in lib.bgan_last_error()
    assert lib.bgan_blur_plan(0, -1.0, 8, 8, ctypes.byref(k), None, None) < 0
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.bgan_blur2d_f32(None, None, 1, 1, 8, 8, 0, 0, 1.0, 1, None, None, 0, None) < 0
    assert lib.bgan_blur2d_f32(8, 16, 1, 2, 8, 8, _lib.LAYOUT_NHWC, 0, 1.0, 1, None, None, 0, None) < 0  # NHWC needs C in {1,3}
    assert lib.bgan_swd_sort_f32(None, 10, 1, None, 0, None) < 0
    assert lib.bgan_swd_project_f32(None, 1, 1, None, 1, None, None) < 0
    assert lib.bgan_swd_sort_workspace_bytes(1 << 21, 128) >= (1 << 21) * 128 * 4
    assert lib.bgan_blur2d_workspace_bytes(256, 3, 128, 128, 0, 57) == 0          # fused path: no scratch
    # large planes, wide blur: transposed intermediate + row maxima of the two-pass tensor-core path
    plane = 2 * 3 * 1024 * 1024 * 4
    assert plane <= lib.bgan_blur2d_workspace_bytes(2, 3, 1024, 1024, 0, 160) <= plane + 2 * 3 * 2048 * 4 + 512
    # NHWC, 3 channels: planar copy + the two-pass tensor-core kernel's scratch on the channel planes (path 7)
    assert 2 * plane <= lib.bgan_blur2d_workspace_bytes(2, 3, 1024, 1024, 1, 160) <= 2 * plane + 2 * 3 * 2048 * 4 + 1024
    assert lib.bgan_blur2d_workspace_bytes(2, 3, 1024, 1024, 1, 300) == plane                     # radius > 256: two launches of the tile kernel


def test_new_entry_points_validate_their_arguments_without_a_gpu(lib):
    """bgan_blur2d_multi_f32 / bgan_lerp_rows_f32 / bgan_swd_feed_f32: argument errors are reported before any CUDA call."""
    two = (ctypes.c_void_p * 2)(256, 512)
    outs = (ctypes.c_void_p * 2)(1024, 2048)
    five_x = (ctypes.c_void_p * 5)(*range(256, 256 * 6, 256))
    five_y = (ctypes.c_void_p * 5)(*range(4096, 4096 + 256 * 5, 256))
    assert lib.bgan_blur2d_multi_f32(None, None, 2, 1, 3, 128, 128, 0, 0, 1.0, 1, None, None, 0, None) < 0
    assert lib.bgan_blur2d_multi_f32(five_x, five_y, 5, 1, 3, 128, 128, 0, 0, 1.0, 1, None, None, 0, None) < 0   # at most 4 tensors
    assert b"tensors" in lib.bgan_last_error()
    assert lib.bgan_blur2d_multi_f32(two, outs, 0, 1, 3, 128, 128, 0, 0, 1.0, 1, None, None, 0, None) < 0
    aliased = (ctypes.c_void_p * 2)(1024, 256)                                                               # ys[1] == xs[0]
    assert lib.bgan_blur2d_multi_f32(two, aliased, 2, 1, 3, 128, 128, 0, 0, 1.0, 1, None, None, 0, None) < 0
    assert b"aliases" in lib.bgan_last_error()
    same_out = (ctypes.c_void_p * 2)(1024, 1024)
    assert lib.bgan_blur2d_multi_f32(two, same_out, 2, 1, 3, 128, 128, 0, 0, 1.0, 1, None, None, 0, None) < 0
    assert lib.bgan_blur2d_multi_f32(two, outs, 2, 1, 3, 128, 128, 0, 0, -1.0, 1, None, None, 0, None) < 0      # std must be positive
    assert lib.bgan_blur2d_multi_f32(two, outs, 2, 1, 2, 128, 128, _lib.LAYOUT_NHWC, 0, 1.0, 1, None, None, 0, None) < 0

    assert lib.bgan_lerp_rows_f32(None, None, None, None, 4, 16, None) < 0
    assert lib.bgan_lerp_rows_f32(256, 512, 768, 1024, 4, 0, None) < 0
    assert lib.bgan_lerp_rows_f32(256, 512, 768, 1024, 0, 16, None) == 0                                        # empty batch: nothing to do

    # pyramid workspace of one feed: residuals of levels 0..L-2 and Gaussian images of levels 1..L-1
    n, c = 64, 3
    want = 4 * n * c * (128 * 128 + 64 * 64 + 32 * 32 + 64 * 64 + 32 * 32 + 16 * 16)
    assert want <= lib.bgan_swd_feed_workspace_bytes(n, c, 128, 128, 4) <= want + 512
    assert lib.bgan_swd_feed_workspace_bytes(n, c, 128, 128, 0) == 0
    assert lib.bgan_swd_feed_f32(None, n, c, 128, 128, 4, None, 128, 7, None, None, None, 0, None) < 0
    four = (ctypes.c_void_p * 4)(256, 512, 768, 1024)
    assert lib.bgan_swd_feed_f32(256, n, c, 128, 128, 4, 512, 128, 7, four, four, None, 0, None) < 0            # workspace missing
    assert b"workspace" in lib.bgan_last_error()
    assert lib.bgan_swd_feed_f32(256, n, c, 100, 100, 4, 512, 128, 7, four, four, 2048, 1 << 30, None) < 0      # 100 does not halve 3 times


_lib_E_WORKSPACE = -2  # BGAN_E_WORKSPACE


def test_sort_diff_validates_its_arguments_without_a_gpu(lib):
    """bgan_swd_sort_diff_f32 (sort + |A - B| fused): argument and workspace errors before any CUDA call."""
    assert lib.bgan_swd_sort_diff_f32(None, None, 10, 1, 0, 1, 1, None, None, 0, None) < 0
    assert lib.bgan_swd_sort_diff_f32(256, 512, 10, 4, 0, 2, 1, 768, 1024, 1 << 30, None) < 0   # segments 0..3 / 2 -> sums 0..1, only 1 given
    assert b"do not map" in lib.bgan_last_error()
    assert lib.bgan_swd_sort_diff_f32(256, 512, 10, 4, 126, 128, 1, 768, 1024, 1 << 30, None) < 0  # offset pushes the last segment into sum 1
    assert lib.bgan_swd_sort_diff_f32(256, 512, 10, 4, 0, 2, 2, 768, 1024, 16, None) == _lib_E_WORKSPACE
    assert lib.bgan_swd_sort_diff_workspace_bytes(1 << 18, 64, 4) >= lib.bgan_swd_sort_workspace_bytes(1 << 18, 64) + 4 * 64 * 8
    assert lib.bgan_swd_sort_diff_workspace_bytes(0, 64, 4) == 0
