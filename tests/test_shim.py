"""include/psl_orbslam_shim.hpp — ORB_SLAM2::ORBextractor / LINEextractor with the reference's signatures over the C-ABI —
compiled against mock cv:: / Eigen types (tests/shim) and, on the GPU, driven the way Frame::ExtractORB / ExtractLSD drive
the reference's classes (src/Frame.cc:311-317, 489-494)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

SHIM = os.path.join(ROOT, "tests", "shim")


def _build():
    subprocess.check_call(["make", "-s", "-C", SHIM])
    return os.path.join(SHIM, "libshim_test.so")


def test_shim_header_compiles_and_links():
    so = _build()
    out = subprocess.check_output(["nm", "-D", so]).decode()
    assert " T shim_extract" in out
    for sym in ("psl_orb_extract", "psl_line_extract", "psl_create", "psl_destroy", "psl_orb_tables"):
        assert f" U {sym}" in out, sym      # resolved by libpsl_frontend.so, i.e. the shim really forwards to the C-ABI


@pytest.mark.gpu
def test_shim_classes_reproduce_the_goldens(orc):
    from psl_slam_b200._lib import KEYLINE_DTYPE, KP_DTYPE
    lib = C.CDLL(_build())
    g = load_golden("orb_vga_seed1")
    nf, nl, ini, mn = [int(v) for v in g["params"]]
    img = np.ascontiguousarray(g["image"])
    h, w = img.shape
    cap, lcap = nf + 4 * nl + 64, 200
    kps, desc = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
    kl, ld, eq = np.zeros(lcap, KEYLINE_DTYPE), np.zeros((lcap, 32), np.uint8), np.zeros((lcap, 3), np.float64)
    nlines = C.c_int(0)
    n = lib.shim_extract(img.ctypes.data_as(C.c_void_p), w, h, nf, C.c_float(float(g["scale_factor"])), nl, ini, mn,
                         kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap,
                         kl.ctypes.data_as(C.c_void_p), ld.ctypes.data_as(C.c_void_p), eq.ctypes.data_as(C.c_void_p), lcap,
                         C.byref(nlines))
    assert n == len(g["kps"]), n
    assert np.array_equal(np.stack([kps[f][:n] for f in ("x", "y", "size", "angle", "response")], 1), g["kps"])
    assert np.array_equal(kps["octave"][:n], g["octave"]) and np.array_equal(desc[:n], g["desc"])
    okl, old, oeq, _ = orc.line_extract(img, lcap)
    m = nlines.value
    assert m == len(okl) and m > 5
    assert kl[:m].tobytes() == okl.tobytes() and np.array_equal(ld[:m], old) and np.array_equal(eq[:m], oeq)
