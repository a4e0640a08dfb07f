"""CPU-side checks of the drop-in boundary: the library builds/loads, exports every symbol that
include/vlg_b200.h declares, the ctypes mirror of vlg_problem_t matches the C layout, and argument
validation works without a GPU (no compute call is made here)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vlg_b200.h")

import vlg_b200  # noqa: E402
from vlg_b200 import _cabi  # noqa: E402


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlg_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _cabi.load()
    declared = _declared_symbols()
    assert set(declared) == set(_cabi.EXPORTS), (declared, _cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vlg_version() == 100


def test_problem_struct_layout_matches_c():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "vlg_b200.h"
int main(void){
  printf("%zu", sizeof(vlg_problem_t));
#define O(f) printf(" %zu", offsetof(vlg_problem_t, f));
  O(N) O(H) O(W) O(K) O(dtype) O(padding) O(coord_mode) O(flags) O(ignore_index)
  O(w_l1) O(w_gd) O(w_ssim) O(w_ce) O(w_tv) O(term_mask) O(ce_norm) O(global_N) O(ce_class_weight)
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    P = _cabi.Problem
    names = ["N", "H", "W", "K", "dtype", "padding", "coord_mode", "flags", "ignore_index", "w_l1", "w_gd",
             "w_ssim", "w_ce", "w_tv", "term_mask", "ce_norm", "global_N", "ce_class_weight"]
    assert vals[0] == C.sizeof(P)
    assert vals[1:] == [getattr(P, n).offset for n in names]


def test_header_constants_match_binding():
    text = open(HEADER).read()
    consts = dict(re.findall(r"#define\s+(VLG_[A-Z0-9_]+)\s+\(?(-?\d+)u?\)?", text))
    assert int(consts["VLG_F32"]) == _cabi.F32 and int(consts["VLG_BF16"]) == _cabi.BF16
    assert int(consts["VLG_PAD_BORDER"]) == _cabi.PAD_BORDER
    assert int(consts["VLG_COORD_GRID"]) == _cabi.COORD_GRID
    assert int(consts["VLG_NEAR_RADIUS"]) == _cabi.NEAR_RADIUS
    assert int(consts["VLG_LOSS_SLOTS"]) == _cabi.LOSS_SLOTS
    assert int(consts["VLG_LOSS_TOTAL"]) == _cabi.LOSS_TOTAL
    assert int(consts["VLG_TERM_ALL"]) == _cabi.TERM_ALL


def test_argument_validation_without_gpu():
    lib = _cabi.load()
    bad = _cabi.Problem(N=1, H=1, W=8, K=20)
    assert lib.vlg_workspace_bytes(C.byref(bad), 0) == 0
    assert b"H>=2" in lib.vlg_last_error()
    badk = _cabi.Problem(N=1, H=8, W=8, K=7)
    assert lib.vlg_workspace_bytes(C.byref(badk), 0) == 0
    assert b"K=7" in lib.vlg_last_error()
    ok = _cabi.Problem(N=2, H=128, W=256, K=20, padding=1)
    small = lib.vlg_workspace_bytes(C.byref(ok), 0)
    big = lib.vlg_workspace_bytes(C.byref(ok), 1)
    P = 2 * 128 * 256
    assert 0 < small < big
    assert big >= P * 23 * 4 + P * 23 * 8
    ok.flags = _cabi.FLAG_NO_FAR_PATH
    assert lib.vlg_workspace_bytes(C.byref(ok), 1) < big
    # NULL workspace is refused before anything is launched
    rc = lib.vlg_reduce_partials(C.byref(ok), None, None, 0, None)
    assert rc == -1


def test_no_cpu_fallback():
    import torch
    x = torch.zeros(1, 3, 8, 8)
    with pytest.raises(vlg_b200.VlgError):
        vlg_b200.L1Loss()(x, x)
    with pytest.raises(vlg_b200.VlgError):
        vlg_b200.warp(x, None, torch.zeros(1, 8, 8, 2))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; the package must not import, load or execute it."""
    pkg = os.path.join(ROOT, "video-layout-generation_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|libvlg_oracle|c_oracle|torch_oracle|oracle/")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f
