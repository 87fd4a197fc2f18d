"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/c → _build/libpsl_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpsl_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th", C.c_int32), ("min_th", C.c_int32)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "c", f) for f in os.listdir(os.path.join(_HERE, "c"))]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_ic_angle.restype = C.c_float
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def params(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7) -> OrbParams:
    return OrbParams(nfeatures, scale_factor, nlevels, ini_th, min_th)


def resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def gauss_blur(src: np.ndarray, ksize: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gauss_blur_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), src.shape[1], ksize)
    return dst


def fast_atan2(y: float, x: float) -> float:
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def fast_score(img: np.ndarray, th: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_fast_score(_p(img), img.shape[1], img.shape[0], img.strides[0], th, _p(out))
    return out


def orb_tables(p: OrbParams):
    L = p.nlevels
    scale = np.empty(L, np.float32)
    inv = np.empty(L, np.float32)
    quota = np.empty(L, np.int32)
    umax = np.empty(16, np.int32)
    lib().orc_orb_tables(C.byref(p), _p(scale), _p(inv), _p(quota), _p(umax))
    return scale, inv, quota, umax


def fast_cells(img: np.ndarray, ini_th=20, min_th=7) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size // 4 + 16
    out = np.empty((cap, 3), np.float32)
    n = lib().orc_fast_cells(_p(img), img.shape[1], img.shape[0], img.strides[0], ini_th, min_th, _p(out), cap)
    return out[:n].copy()


def octree(xyr: np.ndarray, min_x, max_x, min_y, max_y, N) -> np.ndarray:
    xyr = np.ascontiguousarray(xyr, np.float32)
    out = np.empty((len(xyr) + 4, 3), np.float32)
    n = lib().orc_octree(_p(xyr), len(xyr), min_x, max_x, min_y, max_y, N, _p(out), len(out))
    assert n >= 0
    return out[:n].copy()


def orb_extract(img: np.ndarray, p: OrbParams | None = None, cap: int | None = None):
    """Returns (kps structured array [n] of KP_DTYPE, desc u8 [n,32])."""
    p = p or params()
    img = np.ascontiguousarray(img, np.uint8)
    cap = cap or (p.nfeatures + 4 * p.nlevels + 16)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = C.c_int(0)
    rc = lib().orc_orb_extract(C.byref(p), _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc),
                               cap, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"orc_orb_extract rc={rc}")
    return kps[: n.value].copy(), desc[: n.value].copy()


def orb_extract_batch_mt(frames: np.ndarray, p: OrbParams | None = None, nthreads: int = 1) -> np.ndarray:
    """CPU-baseline harness: extract every frame of [B,H,W] on `nthreads` threads; returns counts."""
    p = p or params()
    frames = np.ascontiguousarray(frames, np.uint8)
    B, H, W = frames.shape
    n = np.zeros(B, np.int32)
    x = C.c_uint32(0)
    rc = lib().orc_orb_extract_batch_mt(C.byref(p), _p(frames), B, W, H, frames.strides[1],
                                        C.c_int64(frames.strides[0]), nthreads, _p(n), C.byref(x))
    if rc != 0:
        raise RuntimeError("orc_orb_extract_batch_mt failed")
    return n
