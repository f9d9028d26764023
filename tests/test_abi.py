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
    assert lib.bgan_blur2d_workspace_bytes(2, 3, 1024, 1024, 1, 160) == plane                     # NHWC: two launches of the tile kernel
