"""ORACLE (test infrastructure, NOT product code) -- ctypes loader for oracle/warp_oracle.c.

Numpy in / numpy out, NCHW fp32, grid [N,H,W,2].  See warp_oracle.c for the citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
PAD = {"zeros": 0, "border": 1}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libvlg_oracle.so")
    src = os.path.join(_HERE, "warp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libvlg_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        for name in ("vlgo_l1", "vlgo_gd", "vlgo_ssim", "vlgo_ce", "vlgo_tv"):
            getattr(_LIB, name).restype = C.c_double
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def base_grid(N, H, W):
    g = np.empty((N, H, W, 2), np.float32)
    lib().vlgo_base_grid(N, H, W, g.ctypes.data_as(C.c_void_p))
    return g


def flow_to_grid(flow):
    flow, fp = _f(flow)
    N, H, W, _ = flow.shape
    g = np.empty_like(flow)
    lib().vlgo_flow_to_grid(fp, N, H, W, g.ctypes.data_as(C.c_void_p))
    return g


def sample_coords(grid, padding="border"):
    grid, gp = _f(grid)
    N, H, W, _ = grid.shape
    ixy = np.empty((N, H, W, 2), np.float32)
    x0y0 = np.empty((N, H, W, 2), np.int32)
    w4 = np.empty((N, H, W, 4), np.float32)
    lib().vlgo_sample_coords(gp, N, H, W, PAD[padding], ixy.ctypes.data_as(C.c_void_p),
                             x0y0.ctypes.data_as(C.c_void_p), w4.ctypes.data_as(C.c_void_p))
    return ixy, x0y0, w4


def warp_fwd(src, grid, padding="border"):
    src, sp = _f(src)
    grid, gp = _f(grid)
    N, Cc, H, W = src.shape
    assert grid.shape == (N, H, W, 2)
    out = np.empty_like(src)
    lib().vlgo_warp_fwd(sp, gp, N, Cc, H, W, PAD[padding], out.ctypes.data_as(C.c_void_p))
    return out


def argmax(x):
    x, xp = _f(x)
    N, Cc, H, W = x.shape
    out = np.empty((N, H, W), np.int64)
    lib().vlgo_argmax(xp, N, Cc, H, W, out.ctypes.data_as(C.c_void_p))
    return out


def warp_bwd(src, grid, gout, padding="border"):
    src, sp = _f(src)
    grid, gp = _f(grid)
    gout, op = _f(gout)
    N, Cc, H, W = src.shape
    dsrc = np.empty_like(src)
    dgrid = np.empty_like(grid)
    lib().vlgo_warp_bwd(sp, gp, op, N, Cc, H, W, PAD[padding], dsrc.ctypes.data_as(C.c_void_p),
                        dgrid.ctypes.data_as(C.c_void_p))
    return dsrc, dgrid


def l1(a, b):
    a, ap = _f(a); b, bp = _f(b)
    return lib().vlgo_l1(ap, bp, C.c_size_t(a.size))


def gd(a, b):
    a, ap = _f(a); b, bp = _f(b)
    N, Cc, H, W = a.shape
    return lib().vlgo_gd(ap, bp, N, Cc, H, W)


def ssim(a, b):
    a, ap = _f(a); b, bp = _f(b)
    N, Cc, H, W = a.shape
    return lib().vlgo_ssim(ap, bp, N, Cc, H, W)


def ce(logits, label, ignore_index=-100):
    logits, lp = _f(logits)
    label = np.ascontiguousarray(label, dtype=np.int64)
    N, Cc, H, W = logits.shape
    return lib().vlgo_ce(lp, label.ctypes.data_as(C.c_void_p), N, Cc, H, W, C.c_int64(ignore_index))


def tv(flow):
    flow, fp = _f(flow)
    N, H, W, _ = flow.shape
    return lib().vlgo_tv(fp, N, H, W)
