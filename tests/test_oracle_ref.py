"""CPU: the oracle against the REFERENCE'S OWN CODE.  oracle/_ref/libpsl_ref_orb.so is /root/reference/src/ORBextractor.cc
compiled unchanged (recipe: oracle/Makefile `ref`) against the OpenCV stand-in headers of oracle/ref_shim, whose primitives
(FAST per call, resize, GaussianBlur, copyMakeBorder, fastAtan2) are the cv2-4.13-verified arithmetic.  Everything
first-party — the cell loop with its threshold fallback and skip tests, DistributeOctTree / DivideNode, IC_Angle,
computeOrbDescriptor, operator() — runs as the authors wrote it.  This pins rows A0-A7 of SURVEY.md §8 to the reference.

Two builds: the pinned one (declared choices H1: node addresses in creation order, H2: correctly rounded cos / sin) must
equal the oracle bit for bit; the native one (glibc malloc addresses, libm cosf / sinf) shows how far those two choices
move the reference's own output."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
ORB = golden_names("orb_")


@pytest.fixture(scope="module")
def ref():
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    libs = {}
    for kind, name in (("pinned", "libpsl_ref_orb.so"), ("native", "libpsl_ref_orb_native.so")):
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            pytest.skip("oracle/_ref not built (needs /root/reference)")
        lib = C.CDLL(path)
        lib.ref_orb_extract.restype = C.c_int
        lib.ref_orb_extract.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_orb_tables.restype = C.c_int
        lib.ref_orb_tables.argtypes = [C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 4
        libs[kind] = lib
    return libs


def _extract(lib, img, nf, sf, nl, ini, mn):
    from psl_slam_b200._lib import KP_DTYPE
    img = np.ascontiguousarray(img)
    cap = nf + 64 * nl
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = lib.ref_orb_extract(nf, sf, nl, ini, mn, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                            kps.ctypes.data, desc.ctypes.data, cap)
    assert 0 <= n <= cap
    return kps[:n], desc[:n]


def _golden_params(g):
    nf, nl, ini, mn = [int(v) for v in g["params"]]
    return nf, float(g["scale_factor"]), nl, ini, mn


def test_tables_equal_reference_getters(ref, orc):
    for nf, sf, nl in ((1000, 1.2, 8), (4000, 1.2, 12), (500, 1.5, 5)):
        arrs = [np.zeros(nl, np.float32) for _ in range(4)]
        assert ref["pinned"].ref_orb_tables(nf, sf, nl, *[a.ctypes.data for a in arrs]) == nl
        scale, inv, _, _ = orc.orb_tables(orc.params(nf, sf, nl, 20, 7))
        assert np.array_equal(arrs[0], scale) and np.array_equal(arrs[1], inv)
        assert np.array_equal(arrs[2], scale * scale) and np.array_equal(arrs[3], np.float32(1.0) / (scale * scale))


@pytest.mark.parametrize("name", ORB)
def test_oracle_equals_reference_build(ref, orc, name):
    g = load_golden(name)
    nf, sf, nl, ini, mn = _golden_params(g)
    rk, rd = _extract(ref["pinned"], g["image"], nf, sf, nl, ini, mn)
    ok, od = orc.orb_extract(g["image"], orc.params(nf, sf, nl, ini, mn))
    assert len(rk) == len(ok)
    for f in rk.dtype.names:
        assert np.array_equal(rk[f], ok[f]), f
    assert np.array_equal(rd, od)
    # and the committed golden (cv2 primitives + Python control flow) says the same
    assert np.array_equal(np.stack([rk["x"], rk["y"], rk["size"], rk["angle"], rk["response"]], 1), g["kps"])
    assert np.array_equal(rk["octave"], g["octave"]) and np.array_equal(rd, g["desc"])


@pytest.mark.parametrize("seed,w,h", [(21, 640, 480), (22, 320, 240), (23, 752, 480)])
def test_oracle_equals_reference_build_on_fresh_frames(ref, orc, seed, w, h):
    """Frames no golden holds (the synthetic sequence generator + pure noise), strided input."""
    from psl_slam_b200 import synth
    gray, _, _ = synth.sequence(seed, 1, w, h, poster_size=2048)
    rng = np.random.default_rng(seed)
    frames = [(gray[0], 8), (rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8), 4)]
    for img, nl in frames:
        pad = np.zeros((img.shape[0], img.shape[1] + 5), np.uint8)
        pad[:, :img.shape[1]] = img
        view = pad[:, :img.shape[1]]
        rk, rd = _extract(ref["pinned"], view, 1000, 1.2, nl, 20, 7)
        ok, od = orc.orb_extract(np.ascontiguousarray(img), orc.params(1000, 1.2, nl, 20, 7))
        assert len(rk) == len(ok) and len(rk) > 50
        assert rk.tobytes() == ok.tobytes() and np.array_equal(rd, od)


def test_sensitivity_to_the_declared_choices(ref, orc):
    """H1 / H2 with the real allocator and libm: the reference itself moves by a few keypoints (reported, bounded)."""
    g = load_golden("orb_vga_seed1")
    nf, sf, nl, ini, mn = _golden_params(g)
    pk, pd = _extract(ref["pinned"], g["image"], nf, sf, nl, ini, mn)
    nk, nd = _extract(ref["native"], g["image"], nf, sf, nl, ini, mn)
    key = lambda k: set(zip(k["x"].tolist(), k["y"].tolist(), k["octave"].tolist()))
    common = key(pk) & key(nk)
    moved = len(key(pk) - common)
    pdict = {(x, y, o): d.tobytes() for x, y, o, d in zip(pk["x"], pk["y"], pk["octave"], pd)}
    ndict = {(x, y, o): d.tobytes() for x, y, o, d in zip(nk["x"], nk["y"], nk["octave"], nd)}
    desc_diff = sum(pdict[k] != ndict[k] for k in common)
    print(f"\nreference, native vs pinned build: {len(pk)} / {len(nk)} keypoints, {moved} selected differently (H1), "
          f"{desc_diff} descriptors of the common ones differ (H2)")
    assert abs(len(pk) - len(nk)) <= 8 and moved <= 0.05 * len(pk) and desc_diff <= 0.01 * len(pk)


def test_line_iterator_equals_reference_build(orc):
    """C9: the grid walk of Frame::AssignFeaturesToGridForLine — the oracle's Bresenham against the reference's
    add_src/lineIterator.cpp compiled unchanged."""
    from psl_slam_b200._lib import KEYLINE_DTYPE, make_line_frame_view
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    path = os.path.join(REF_DIR, "libpsl_ref_lineiter.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    lib = C.CDLL(path)
    lib.ref_line_iterator.restype = C.c_int
    lib.ref_line_iterator.argtypes = [C.c_double] * 4 + [C.c_void_p, C.c_int]
    rng = np.random.default_rng(5)
    n = 400
    kl = np.zeros(n, KEYLINE_DTYPE)
    # end points inside and outside the image, horizontal / vertical / single-cell / steep lines included
    kl["start_x"], kl["end_x"] = rng.uniform(-40, 680, n), rng.uniform(-40, 680, n)
    kl["start_y"], kl["end_y"] = rng.uniform(-40, 520, n), rng.uniform(-40, 520, n)
    kl["end_y"][:20] = kl["start_y"][:20]
    kl["end_x"][20:40] = kl["start_x"][20:40]
    kl["end_x"][40:50], kl["end_y"][40:50] = kl["start_x"][40:50] + 1, kl["start_y"][40:50] + 1
    bounds = (0.0, 0.0, 640.0, 480.0)
    view, keep = make_line_frame_view(kl, np.zeros((n, 32), np.uint8), np.zeros((n, 3)), None, bounds)
    gw, gh = np.float32(view.grid_w_inv), np.float32(view.grid_h_inv)
    xy = np.zeros((4096, 2), np.int32)
    for i in range(n):
        cells = orc.line_grid_cells(view, i)
        a = [float(np.float32(kl[f][i]) * g) for f, g in (("start_x", gw), ("start_y", gh), ("end_x", gw), ("end_y", gh))]
        m = lib.ref_line_iterator(*a, xy.ctypes.data, len(xy))
        assert m <= len(xy)
        ref_cells = [int(x) * 48 + int(y) for x, y in xy[:m] if 0 <= x < 64 and 0 <= y < 48]
        assert list(cells) == ref_cells, i
