"""CPU: the device-side sequential cores (lsd_core.cuh, line_core.cuh, orb_octree_core.cuh) compiled for the
host (tests/emu/, PSL_HOST_EMU) against the oracle — the same source the CUDA kernels run, checked without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden

EMU = os.path.join(ROOT, "tests", "emu")
LINES = golden_names("line_")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU])
    return {n: C.CDLL(os.path.join(EMU, f"lib{n}_emu.so")) for n in ("octree", "lsd", "line")}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", LINES)
def test_lsd_core_vs_oracle(emu, orc, name):
    g = load_golden(name)
    scaled = orc.lsd_scaled_image(g["image"])
    out = np.zeros((8192, 4), np.float32)
    n = emu["lsd"].emu_lsd(_p(scaled), scaled.shape[1], scaled.shape[0], _p(out), 8192)
    assert np.array_equal(out[:n], g["lsd_raw"])


@pytest.mark.parametrize("name", LINES)
def test_line_core_vs_oracle(emu, orc, name):
    from psl_slam_b200._lib import KEYLINE_DTYPE
    g = load_golden(name)
    h, w = g["image"].shape
    raw = np.ascontiguousarray(g["lsd_raw"], np.float32)
    nf = int(g["nfeatures"])
    kl = np.zeros(max(nf, 1), KEYLINE_DTYPE)
    eq = np.zeros((max(nf, 1), 3), np.float64)
    ov = C.c_int(0)
    n = emu["line"].emu_frame_lines(_p(raw), len(raw), w, h, nf, _p(kl), _p(eq), len(kl), C.byref(ov))
    assert ov.value == 0 and n == len(g["keylines"])
    for f in kl.dtype.names:
        assert np.array_equal(kl[f][:n], g["keylines"][f]), f
    assert np.array_equal(eq[:n], g["lineeq"])


def test_line_core_capacity_is_reported(emu):
    from psl_slam_b200._lib import KEYLINE_DTYPE
    g = load_golden("line_textured_top60")
    h, w = g["image"].shape
    raw = np.ascontiguousarray(g["lsd_raw"], np.float32)
    kl = np.zeros(8, KEYLINE_DTYPE)
    eq = np.zeros((8, 3), np.float64)
    ov = C.c_int(0)
    assert emu["line"].emu_frame_lines(_p(raw), len(raw), w, h, 60, _p(kl), _p(eq), 8, C.byref(ov)) == -1


@pytest.mark.parametrize("nt", [32, 128, 256])
def test_octree_core_vs_oracle(emu, orc, nt):
    from psl_slam_b200 import synth
    gray, _, _ = synth.sequence(11, 1, 320, 240, poster_size=1024)
    img = gray[0]
    h, w = img.shape
    cand = orc.fast_cells(img)
    assert len(cand) > 300
    width, height = (w - 16) - 16, (h - 16) - 16  # maxBorder - minBorder, ORBextractor.cc:773-778
    for N in (40, 150, 600):
        if nt == 128 and N > 200:
            continue
        ref = orc.octree(cand, 0, width, 0, height, N)
        packed = ((cand[:, 0].astype(np.uint32) << 20) | (cand[:, 1].astype(np.uint32) << 8) |
                  cand[:, 2].astype(np.uint32)).astype(np.uint32)
        out = np.zeros(N + 64, np.uint32)
        n_ini = int(round(width / height))
        n = emu["octree"].emu_octree(_p(packed), len(packed), n_ini, C.c_float(width / n_ini), width, height, N,
                                     _p(out), len(out), nt)
        assert n == len(ref)
        got = np.stack([out[:n] >> 20, (out[:n] >> 8) & 0xFFF, out[:n] & 0xFF], 1).astype(np.float32)
        assert np.array_equal(got, ref)
