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


def build_native() -> str:
    """-O3 -march=native build of the same sources ON THIS HOST, for the timed CPU baseline (the
    reference builds with -O3 -march=native).  The portable x86-64-v3 build stays the checker of the tests."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith(("flags", "model name")):
                    flags += line
                if line.strip() == "" and flags:
                    break
    except OSError:
        pass
    # one file per host CPU: a build made for another machine (e.g. in the build container) is never loaded here
    so = os.path.join(_HERE, "_build", f"libpsl_oracle_native_{hashlib.sha1(flags.encode()).hexdigest()[:10]}.so")
    subprocess.check_call(["make", "-s", "-C", _HERE, "native", f"NATIVE_OUT={so}"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        so = _SO
        if os.environ.get("PSL_ORACLE_NATIVE") == "1":
            try:
                so = build_native()
            except Exception:   # no compiler on this host: the portable build
                so = _SO
        if so == _SO and not os.path.exists(_SO):
            build()
        _lib = C.CDLL(so)
        _lib.path = so
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


# ---- matchers (oracle/c/orc_match.cpp) --------------------------------------------------------
from psl_slam_b200._lib import (MatchParams, QUERY_DTYPE, make_feature_vector,  # noqa: E402  (ABI structs only)
                                make_frame_view)


def descriptor_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    out = np.empty(len(a), np.int32)
    lib().orc_descriptor_distance(_p(a), _p(b), len(a), _p(out))
    return out


def hamming_knn2(q: np.ndarray, t: np.ndarray):
    q, t = np.ascontiguousarray(q, np.uint8), np.ascontiguousarray(t, np.uint8)
    idx = np.empty((len(q), 2), np.int32)
    dist = np.empty((len(q), 2), np.int32)
    lib().orc_hamming_knn2(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist))
    return idx, dist


def features_in_area(kps_un, desc, bounds, x, y, r, min_level=-1, max_level=-1):
    fv, keep = make_frame_view(kps_un, None, desc, bounds)
    out = np.empty(len(desc) + 1, np.int32)
    n = lib().orc_features_in_area(C.byref(fv), C.c_float(x), C.c_float(y), C.c_float(r), min_level, max_level,
                                   _p(out), len(out))
    return out[:n].copy()


def match_projection(kps_un, u_right, desc, bounds, queries, qdesc, claimed_in, mode, th_dist=100, nn_ratio=0.6,
                     check_orientation=True):
    fv, keep = make_frame_view(kps_un, u_right, desc, bounds)
    queries = np.ascontiguousarray(queries, QUERY_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    cl = None if claimed_in is None else np.ascontiguousarray(claimed_in, np.uint8)
    prm = MatchParams(mode, th_dist, nn_ratio, int(check_orientation))
    assign = np.empty(len(desc), np.int32)
    nm = C.c_int32()
    lib().orc_match_projection(C.byref(fv), _p(queries), _p(qdesc), len(queries), None if cl is None else _p(cl),
                               C.byref(prm), _p(assign), C.byref(nm))
    return assign, nm.value


def match_bow(kf_desc, kf_angle, kf_valid, kf_csr, f_desc, f_angle, f_csr, nn_ratio=0.7, th_low=50,
              check_orientation=True):
    kf_desc, f_desc = np.ascontiguousarray(kf_desc, np.uint8), np.ascontiguousarray(f_desc, np.uint8)
    kf_angle, f_angle = np.ascontiguousarray(kf_angle, np.float32), np.ascontiguousarray(f_angle, np.float32)
    kf_valid = np.ascontiguousarray(kf_valid, np.uint8)
    kfv, k1 = make_feature_vector(*kf_csr)
    ffv, k2 = make_feature_vector(*f_csr)
    match = np.empty(len(f_desc), np.int32)
    nm = C.c_int32()
    lib().orc_match_bow(_p(kf_desc), _p(kf_angle), _p(kf_valid), len(kf_desc), C.byref(kfv), _p(f_desc), _p(f_angle),
                        len(f_desc), C.byref(ffv), C.c_float(nn_ratio), th_low, int(check_orientation), _p(match),
                        C.byref(nm))
    return match, nm.value


# ---- Frame bookkeeping (oracle/c/orc_frame.cpp) -------------------------------------------------
def undistort_keypoints(kps, cam):
    """cam: psl_slam_b200._lib.Distortion.  Frame::UndistortKeyPoints."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.empty_like(kps)
    lib().orc_undistort_keypoints(_p(kps), len(kps), C.byref(cam), _p(out))
    return out


def image_bounds(cols, rows, cam):
    b = np.zeros(4, np.float32)
    lib().orc_image_bounds(int(cols), int(rows), C.byref(cam), _p(b))
    return tuple(float(v) for v in b)


def stereo_from_rgbd(kps, depth_u16, depth_factor, bf):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    depth_u16 = np.ascontiguousarray(depth_u16, np.uint16)
    ur = np.empty(len(kps), np.float32)
    z = np.empty(len(kps), np.float32)
    lib().orc_stereo_from_rgbd(_p(kps), len(kps), _p(depth_u16), depth_u16.shape[1], depth_u16.shape[0],
                               depth_u16.shape[1], C.c_float(depth_factor), C.c_float(bf), _p(ur), _p(z))
    return ur, z


def queries_from_last_frame(kps_last, z_last, Tcw_last, Tcw_cur, cam, scale_factors, th, bounds, valid=None,
                            claims=None, mono=False):
    kps_last = np.ascontiguousarray(kps_last, KP_DTYPE)
    z_last = np.ascontiguousarray(z_last, np.float32)
    Tl = np.ascontiguousarray(np.asarray(Tcw_last, np.float32)[:3, :4])
    Tc = np.ascontiguousarray(np.asarray(Tcw_cur, np.float32)[:3, :4])
    cam = np.ascontiguousarray(cam, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    va = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    cl = None if claims is None else np.ascontiguousarray(claims, np.uint8)
    q = np.zeros(len(kps_last), QUERY_DTYPE)
    lib().orc_queries_from_last_frame(_p(kps_last), _p(z_last), None if va is None else _p(va),
                                      None if cl is None else _p(cl), len(kps_last), _p(Tl), _p(Tc), _p(cam), _p(sf),
                                      C.c_float(th), int(mono), C.c_float(bounds[0]), C.c_float(bounds[1]),
                                      C.c_float(bounds[2]), C.c_float(bounds[3]), _p(q))
    return q


def track_batch_mt(gray, depth, Tcw12, cam6, p: OrbParams | None = None, th=15.0, nn_ratio=0.9, check_ori=True,
                   nthreads=1):
    """CPU-baseline harness for the point front end (extract + stereo + SearchByProjection vs previous frame)."""
    p = p or params()
    gray = np.ascontiguousarray(gray, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    T = np.ascontiguousarray(Tcw12, np.float32)
    cam = np.ascontiguousarray(cam6, np.float32)
    B, H, W = gray.shape
    n = np.zeros(B, np.int32)
    nm = np.zeros(B, np.int32)
    rc = lib().orc_track_batch_mt(C.byref(p), _p(gray), _p(depth), B, W, H, _p(T), _p(cam), C.c_float(th),
                                  C.c_float(nn_ratio), int(check_ori), nthreads, _p(n), _p(nm))
    if rc != 0:
        raise RuntimeError("orc_track_batch_mt failed")
    return n, nm


def frontend_batch_mt(gray, depth, Tcw12, cam6, p: OrbParams | None = None, th=15.0, nn_ratio=0.9, check_ori=True,
                      line_nfeatures=200, line_desc_th=0.95, nthreads=1):
    """CPU-baseline harness for the combined front end (points as track_batch_mt + LINEextractor per frame +
    SearchByGeomNApearance per consecutive pair).  Returns (n, nmatches, nl, line_nmatches)."""
    p = p or params()
    gray = np.ascontiguousarray(gray, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    T = np.ascontiguousarray(Tcw12, np.float32)
    cam = np.ascontiguousarray(cam6, np.float32)
    B, H, W = gray.shape
    n, nm, nl, lnm = (np.zeros(B, np.int32) for _ in range(4))
    rc = lib().orc_frontend_batch_mt(C.byref(p), _p(gray), _p(depth), B, W, H, _p(T), _p(cam), C.c_float(th),
                                     C.c_float(nn_ratio), int(check_ori), line_nfeatures, C.c_float(line_desc_th),
                                     nthreads, _p(n), _p(nm), _p(nl), _p(lnm))
    if rc != 0:
        raise RuntimeError("orc_frontend_batch_mt failed")
    return n, nm, nl, lnm


def rgbd_frontend_batch_mt(color, rgb_order, depth, Tcw12, cam6, p: OrbParams | None = None, th=15.0, nn_ratio=0.9,
                           check_ori=True, line_nfeatures=200, line_desc_th=0.95, nthreads: int = 1):
    """GrabImageRGBD for a batch: cvtColor -> GRAY (cv2 4.13 arithmetic), then frontend_batch_mt.  color [B,H,W,3|4] u8."""
    p = p or params()
    color = np.ascontiguousarray(color, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    B, H, W, ch = color.shape
    T = np.ascontiguousarray(Tcw12, np.float32).reshape(B, 12)
    cam = np.ascontiguousarray(cam6, np.float32)
    n, nm, nl, lnm = (np.zeros(B, np.int32) for _ in range(4))
    rc = lib().orc_rgbd_frontend_batch_mt(C.byref(p), _p(color), ch, int(bool(rgb_order)), _p(depth), B, W, H, _p(T), _p(cam),
                                          C.c_float(th), C.c_float(nn_ratio), int(check_ori), int(line_nfeatures),
                                          C.c_float(line_desc_th), int(nthreads), _p(n), _p(nm), _p(nl), _p(lnm))
    if rc != 0:
        raise RuntimeError(f"orc_rgbd_frontend_batch_mt rc={rc}")
    return n, nm, nl, lnm


def color_to_gray(color, rgb_order=True):
    color = np.ascontiguousarray(color, np.uint8)
    gray = np.empty(color.shape[:-1], np.uint8)
    lib().orc_color_to_gray(_p(color), color.shape[-1], int(bool(rgb_order)), _p(gray), C.c_int64(gray.size))
    return gray


def match_triangulation(kf1, fv1, kf2, fv2, F12, ex, ey, scale2, sigma2, only_stereo=False, check_ori=True, th_low=50):
    """kf* = (kps_un, u_right, desc, has_mappoint), fv* = (node ids, offs, idx)."""
    from psl_slam_b200._lib import make_feature_vector, make_keyframe_view
    a, k1 = make_keyframe_view(*kf1)
    b, k2 = make_keyframe_view(*kf2)
    f1, k3 = make_feature_vector(*fv1)
    f2, k4 = make_feature_vector(*fv2)
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    sc, s2 = np.ascontiguousarray(scale2, np.float32), np.ascontiguousarray(sigma2, np.float32)
    m12 = np.zeros(max(a.n, 1), np.int32)
    nm = C.c_int32()
    lib().orc_match_triangulation(C.byref(a), C.byref(f1), C.byref(b), C.byref(f2), _p(F), C.c_float(ex), C.c_float(ey),
                                  _p(sc), _p(s2), int(only_stereo), int(check_ori), th_low, _p(m12), C.byref(nm))
    return m12[: a.n].copy(), nm.value


def match_fuse(kps_un, u_right, desc, bounds, queries, qdesc, inv_sigma2, th_low=50):
    from psl_slam_b200._lib import FUSE_QUERY_DTYPE
    fv, keep = make_frame_view(kps_un, u_right, desc, bounds)
    queries = np.ascontiguousarray(queries, FUSE_QUERY_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    s2 = None if inv_sigma2 is None else np.ascontiguousarray(inv_sigma2, np.float32)   # None: the Sim3 form
    bi = np.zeros(max(len(queries), 1), np.int32)
    bd = np.zeros(max(len(queries), 1), np.int32)
    lib().orc_match_fuse(C.byref(fv), _p(queries), _p(qdesc), len(queries), None if s2 is None else _p(s2), th_low,
                         _p(bi), _p(bd))
    return bi[: len(queries)].copy(), bd[: len(queries)].copy()


def match_bow_kf(desc1, angle1, valid1, csr1, desc2, angle2, valid2, csr2, nn_ratio=0.75, th_low=50,
                 check_orientation=True):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, ...): (matches12 [n1], nmatches)."""
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    angle1, angle2 = np.ascontiguousarray(angle1, np.float32), np.ascontiguousarray(angle2, np.float32)
    valid1, valid2 = np.ascontiguousarray(valid1, np.uint8), np.ascontiguousarray(valid2, np.uint8)
    f1, k1 = make_feature_vector(*csr1)
    f2, k2 = make_feature_vector(*csr2)
    m12 = np.zeros(max(len(desc1), 1), np.int32)
    nm = C.c_int32()
    lib().orc_match_bow_kf(_p(desc1), _p(angle1), _p(valid1), len(desc1), C.byref(f1), _p(desc2), _p(angle2), _p(valid2),
                           len(desc2), C.byref(f2), C.c_float(nn_ratio), th_low, int(check_orientation), _p(m12),
                           C.byref(nm))
    return m12[: len(desc1)].copy(), nm.value


def match_sim3(kf1, kf2, bounds, q12, mp_desc1, q21, mp_desc2, th_high=100):
    """ORBmatcher::SearchBySim3 after the projections; kf* = (kps_un, desc): (matches12 [n1], nFound)."""
    from psl_slam_b200._lib import FUSE_QUERY_DTYPE
    a, ka = make_frame_view(kf1[0], None, kf1[1], bounds)
    b, kb = make_frame_view(kf2[0], None, kf2[1], bounds)
    q12, q21 = np.ascontiguousarray(q12, FUSE_QUERY_DTYPE), np.ascontiguousarray(q21, FUSE_QUERY_DTYPE)
    d1, d2 = np.ascontiguousarray(mp_desc1, np.uint8), np.ascontiguousarray(mp_desc2, np.uint8)
    assert len(q12) == a.n and len(q21) == b.n
    m12 = np.zeros(max(a.n, 1), np.int32)
    nf = C.c_int32()
    lib().orc_match_sim3(C.byref(a), C.byref(b), _p(q12), _p(d1), _p(q21), _p(d2), th_high, _p(m12), C.byref(nf))
    return m12[: a.n].copy(), nf.value


def match_initialization(kps1_un, desc1, prev_matched, f2, bounds, window=100, nn_ratio=0.9, th_low=50,
                         check_orientation=True):
    """ORBmatcher::SearchForInitialization; f2 = (kps_un, desc): (matches12, nmatches, prev_matched after)."""
    kps1 = np.ascontiguousarray(kps1_un, KP_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8)
    pm = np.array(prev_matched, np.float32, copy=True).reshape(-1, 2)
    b, kb = make_frame_view(f2[0], None, f2[1], bounds)
    m12 = np.zeros(max(len(kps1), 1), np.int32)
    nm = C.c_int32()
    lib().orc_match_initialization(_p(kps1), _p(d1), len(kps1), _p(pm), C.byref(b), int(window), C.c_float(nn_ratio),
                                   th_low, int(check_orientation), _p(m12), C.byref(nm))
    return m12[: len(kps1)].copy(), nm.value, pm


# ---- lines ------------------------------------------------------------------------------------------
from psl_slam_b200._lib import KEYLINE_DTYPE  # noqa: E402  (ABI struct only)


def lsd_detect(img: np.ndarray, order_mode: int = 1) -> np.ndarray:
    """cv::LineSegmentDetector(REFINE_STD).detect -> float32 [n,4] (x1,y1,x2,y2).
    order_mode 1 = bins descending, raster order inside a bin (matches cv2 4.13 and OpenCV 3.x);
    0 = unstable std::sort (kept only to show the sensitivity)."""
    img = np.ascontiguousarray(img, np.uint8)
    cap = 8192
    out = np.empty((cap, 4), np.float32)
    n = lib().orc_lsd_detect(_p(img), img.shape[1], img.shape[0], img.strides[0], order_mode, _p(out), cap)
    return out[:n].copy()


def line_iterator_count(w, h, x1, y1, x2, y2) -> int:
    return int(lib().orc_line_iterator_count(w, h, C.c_float(x1), C.c_float(y1), C.c_float(x2), C.c_float(y2)))


def merge_lines_lsd(lines: np.ndarray) -> np.ndarray:
    lines = np.ascontiguousarray(lines, np.float32).reshape(-1, 4)
    out = np.empty((max(len(lines), 1), 4), np.float32)
    n = lib().orc_merge_lines_lsd(_p(lines), len(lines), _p(out), len(out))
    return out[:n].copy()


def clamp_segments(lines: np.ndarray, w: int, h: int) -> np.ndarray:
    lines = np.ascontiguousarray(lines, np.float32).reshape(-1, 4).copy()
    lib().orc_clamp_segments(_p(lines), len(lines), w, h)
    return lines


def make_keylines(lines: np.ndarray, w: int, h: int, nfeatures: int = 200) -> np.ndarray:
    lines = np.ascontiguousarray(lines, np.float32).reshape(-1, 4)
    kl = np.zeros(max(len(lines), 1), KEYLINE_DTYPE)
    n = lib().orc_make_keylines(_p(lines), len(lines), w, h, nfeatures, _p(kl))
    return kl[:n].copy()


def lbd_gradients(img: np.ndarray):
    img = np.ascontiguousarray(img, np.uint8)
    dx = np.empty(img.shape, np.int16)
    dy = np.empty(img.shape, np.int16)
    lib().orc_lbd_gradients(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(dx), _p(dy))
    return dx, dy


def lbd_descriptors(dx: np.ndarray, dy: np.ndarray, kl: np.ndarray):
    """(float [n,72], binary u8 [n,32])"""
    kl = np.ascontiguousarray(kl, KEYLINE_DTYPE)
    des = np.zeros((len(kl), 72), np.float32)
    out = np.zeros((len(kl), 32), np.uint8)
    for i in range(len(kl)):
        lib().orc_lbd_one(_p(dx), _p(dy), dx.shape[1], dx.shape[0], C.c_void_p(kl.ctypes.data + 68 * i),
                          C.c_void_p(des.ctypes.data + 288 * i))
        lib().orc_lbd_binarise(C.c_void_p(des.ctypes.data + 288 * i), C.c_void_p(out.ctypes.data + 32 * i))
    return des, out


def line_extract(img: np.ndarray, nfeatures: int = 200, cap: int = 1024):
    """LINEextractor::operator(): (keylines [n], ldesc u8 [n,32], lineeq f64 [n,3], lbd f32 [n,72])."""
    img = np.ascontiguousarray(img, np.uint8)
    kl = np.zeros(cap, KEYLINE_DTYPE)
    ld = np.zeros((cap, 32), np.uint8)
    eq = np.zeros((cap, 3), np.float64)
    lbd = np.zeros((cap, 72), np.float32)
    n = C.c_int(0)
    rc = lib().orc_line_extract(_p(img), img.shape[1], img.shape[0], img.strides[0], nfeatures, _p(kl), _p(ld), _p(eq),
                                _p(lbd), cap, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"orc_line_extract rc={rc}")
    n = n.value
    return kl[:n].copy(), ld[:n].copy(), eq[:n].copy(), lbd[:n].copy()


def lsd_scaled_image(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    W, H = C.c_int(), C.c_int()
    lib().orc_lsd_scaled_image(_p(img), img.shape[1], img.shape[0], img.strides[0], None, C.byref(W), C.byref(H))
    out = np.empty((H.value, W.value), np.uint8)
    lib().orc_lsd_scaled_image(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), C.byref(W), C.byref(H))
    return out


# ---- line matchers --------------------------------------------------------------------------------
from psl_slam_b200._lib import LINE_QUERY_DTYPE, make_line_frame_view  # noqa: E402  (ABI structs only)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, np.uint8)


def line_match_nnr(d1, d2, nnr):
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros(len(d1), np.int32)
    n = lib().orc_line_match_nnr(_p(d1), len(d1), _p(d2), len(d2), C.c_float(nnr), _p(out))
    return out, int(n)


def line_search_geom(kl_last, d_last, has_ml, kl_cur, d_cur, bounds, desc_th):
    kl_last, kl_cur = np.ascontiguousarray(kl_last, KEYLINE_DTYPE), np.ascontiguousarray(kl_cur, KEYLINE_DTYPE)
    d_last, d_cur, has_ml = _u8(d_last), _u8(d_cur), _u8(has_ml)
    bounds = np.ascontiguousarray(bounds, np.float32)
    out = np.zeros(len(kl_cur), np.int32)
    n = lib().orc_line_search_geom(_p(kl_last), _p(d_last), _p(has_ml), len(kl_last), _p(kl_cur), _p(d_cur),
                                   len(kl_cur), _p(bounds), C.c_float(desc_th), _p(out))
    return out, int(n)


def line_frame_bf_match(d1, d2, nn_ratio, th):
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros(len(d1), np.int32)
    lib().orc_line_frame_bf_match(_p(d1), len(d1), _p(d2), len(d2), C.c_float(nn_ratio), C.c_float(th), _p(out))
    return out


def line_search_double(d1, d2, nn_ratio, th):
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros(len(d1), np.int32)
    n = lib().orc_line_search_double(_p(d1), len(d1), _p(d2), len(d2), C.c_float(nn_ratio), C.c_float(th), _p(out))
    return out, int(n)


def line_grid_cells(view, line):
    """Grid cells (cx * 48 + cy) a key line is registered in, Frame::AssignFeaturesToGridForLine order."""
    out = np.zeros(4096, np.int32)
    n = lib().orc_line_grid_cells(C.byref(view), int(line), _p(out), len(out))
    assert n <= len(out)
    return out[:n]


def lines_in_area(view, x1, y1, x2, y2, r, TH):
    out = np.zeros(max(view.n, 1), np.int32)
    n = lib().orc_lines_in_area(C.byref(view), C.c_float(x1), C.c_float(y1), C.c_float(x2), C.c_float(y2),
                                C.c_float(r), C.c_float(TH), _p(out), len(out))
    return out[:n].copy()


def line_match_projection(view, queries, qdesc, claimed_in, mode, nn_ratio):
    queries = np.ascontiguousarray(queries, LINE_QUERY_DTYPE)
    qdesc, claimed_in = _u8(qdesc), _u8(claimed_in)
    out = np.zeros(max(view.n, 1), np.int32)
    n = lib().orc_line_match_projection(C.byref(view), _p(queries), _p(qdesc), len(queries),
                                        None if claimed_in is None else _p(claimed_in), mode, C.c_float(nn_ratio),
                                        _p(out))
    return out[: view.n].copy(), int(n)


def plane_assoc(planes_cam, pts, Tcw, map_planes, map_bad, d_th, a_th, mode):
    planes_cam = np.ascontiguousarray(planes_cam, np.float32).reshape(-1, 4)
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 15)
    Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(4, 4)
    map_planes = np.ascontiguousarray(map_planes, np.float32).reshape(-1, 4)
    map_bad = _u8(map_bad)
    out = np.zeros(max(len(planes_cam), 1), np.int32)
    n = lib().orc_plane_assoc(_p(planes_cam), _p(pts), len(planes_cam), _p(Tcw), _p(map_planes),
                              None if map_bad is None else _p(map_bad), len(map_planes), C.c_float(d_th),
                              C.c_float(a_th), mode, _p(out))
    return out[: len(planes_cam)].copy(), int(n)


def line_fuse(kl, kf_desc, queries, qdesc, th_cos=0.998, th_low=50):
    """Window search of LSDmatcher::Fuse: (best_idx [nq], best_dist [nq], fused count)."""
    from psl_slam_b200._lib import KEYLINE_DTYPE, LINE_FUSE_QUERY_DTYPE
    kl = np.ascontiguousarray(kl, KEYLINE_DTYPE)
    queries = np.ascontiguousarray(queries, LINE_FUSE_QUERY_DTYPE)
    kf_desc, qdesc = _u8(kf_desc), _u8(qdesc)
    nq = len(queries)
    bi, bd = np.full(max(nq, 1), -1, np.int32), np.full(max(nq, 1), 256, np.int32)
    n = lib().orc_line_fuse(_p(kl), len(kl), _p(kf_desc), _p(queries), _p(qdesc), nq, C.c_float(th_cos), int(th_low),
                            _p(bi), _p(bd))
    return bi[:nq].copy(), bd[:nq].copy(), int(n)


def lines_3d(kl, depth_f32, fx, fy, cx, cy, seed):
    """Frame::isLineGood: (mvLines3D [n,6] f64, mvLineEq [n,3] f32)."""
    from psl_slam_b200._lib import KEYLINE_DTYPE
    kl = np.ascontiguousarray(kl, KEYLINE_DTYPE)
    dep = np.ascontiguousarray(depth_f32, np.float32)
    cam = np.array([fx, fy, cx, cy], np.float32)
    n = len(kl)
    l3, eq = np.zeros((max(n, 1), 6)), np.zeros((max(n, 1), 3), np.float32)
    lib().orc_lines_3d(_p(kl), n, _p(dep), dep.shape[1], dep.shape[0], dep.shape[1], _p(cam), C.c_uint32(seed), _p(l3), _p(eq))
    return l3[:n].copy(), eq[:n].copy()


def plane_hypotheses(kl_un, line_eq, lines3d, junctions, cap=None):
    """Frame::ExtractLSD plane hypotheses: (le_l [nj,6], planes [np,4], normals [np,3], junction_of [np], count)."""
    from psl_slam_b200._lib import JUNCTION_DTYPE, KEYLINE_DTYPE
    kl = np.ascontiguousarray(kl_un, KEYLINE_DTYPE)
    eq = np.ascontiguousarray(line_eq, np.float32).reshape(-1, 3)
    l3 = np.ascontiguousarray(lines3d, np.float64).reshape(-1, 6)
    js = np.ascontiguousarray(junctions, JUNCTION_DTYPE)
    nj = len(js)
    cap = nj if cap is None else cap
    le = np.zeros((max(nj, 1), 6))
    pl, nr, ow = np.zeros((max(cap, 1), 4), np.float32), np.zeros((max(cap, 1), 3)), np.zeros(max(cap, 1), np.int32)
    n = lib().orc_plane_hypotheses(_p(kl), _p(eq), _p(l3), len(kl), _p(js), nj, _p(le), _p(pl), _p(nr), _p(ow), cap)
    m = min(n, cap)
    return le[:nj].copy(), pl[:m].copy(), nr[:m].copy(), ow[:m].copy(), int(n)


def line_search_triangulation(d1, ml1, d2, ml2, nn_ratio, th, is_double):
    d1, d2, ml1, ml2 = _u8(d1), _u8(d2), _u8(ml1), _u8(ml2)
    out = np.zeros(max(len(d1), 1), np.int32)
    n = lib().orc_line_search_triangulation(_p(d1), _p(ml1), len(d1), _p(d2), _p(ml2), len(d2), C.c_float(nn_ratio),
                                            C.c_float(th), int(is_double), _p(out))
    return out[: len(d1)].copy(), int(n)


def line_junctions(kl_un, lines3d, img_w, img_h, radius=20.0, fan_thr=np.pi / 4, cap=4096):
    """CPartiallyRecoverConnectivity + Frame::convertFansToKeyLines: (fans [m,4] f32, junctions [k] JUNCTION_DTYPE)."""
    from psl_slam_b200._lib import JUNCTION_DTYPE
    kl = np.ascontiguousarray(kl_un, KEYLINE_DTYPE)
    l3 = None if lines3d is None else np.ascontiguousarray(lines3d, np.float64)
    fans = np.zeros((cap, 4), np.float32)
    js = np.zeros(cap, JUNCTION_DTYPE)
    nf, nj = C.c_int32(), C.c_int32()
    lib().orc_line_junctions(_p(kl), None if l3 is None else _p(l3), len(kl), int(img_w), int(img_h), C.c_float(radius),
                             C.c_float(fan_thr), _p(fans), None if l3 is None else _p(js), cap, C.byref(nf), C.byref(nj))
    assert nf.value <= cap
    return fans[: nf.value].copy(), js[: nj.value].copy()


def line_search_triangulation_new(kl1, d1, func1, ml1, kl2, d2, func2, ml2, F21, F12, nn_ratio, th, is_double):
    """LSDmatcher::SearchForTriangulationNew: (vMatchedPairs [n1], nmatches)."""
    kl1, kl2 = np.ascontiguousarray(kl1, KEYLINE_DTYPE), np.ascontiguousarray(kl2, KEYLINE_DTYPE)
    d1, d2 = np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(d2, np.uint8)
    f1, f2 = np.ascontiguousarray(func1, np.float64), np.ascontiguousarray(func2, np.float64)
    m1, m2 = np.ascontiguousarray(ml1, np.uint8), np.ascontiguousarray(ml2, np.uint8)
    A, B = np.ascontiguousarray(F21, np.float32).reshape(9), np.ascontiguousarray(F12, np.float32).reshape(9)
    out = np.zeros(max(len(d1), 1), np.int32)
    n = lib().orc_line_search_triangulation_new(_p(kl1), _p(d1), _p(f1), _p(m1), len(d1), _p(kl2), _p(d2), _p(f2), _p(m2),
                                                len(d2), _p(A), _p(B), C.c_float(nn_ratio), C.c_float(th), int(is_double),
                                                _p(out))
    return out[: len(d1)].copy(), int(n)


POSE_POINT_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("u_right", "<f4"), ("inv_sigma2", "<f4"), ("xw", "<f4"),
                             ("yw", "<f4"), ("zw", "<f4"), ("flags", "<u4")])  # psl_pose_point, 32 B


POSE_LIL_DTYPE = np.dtype([("line1", "<f8", (6,)), ("line2", "<f8", (6,)), ("cross", "<f8", (3,)), ("obs1", "<f8", (3,)),
                           ("obs2", "<f8", (3,)), ("ins", "<f8", (2,)), ("flags", "<u4"), ("pad_", "<u4")])  # psl_pose_lil, 192 B


def pose_optimization(Tcw, pts, fx, fy, cx, cy, bf, lils=None):
    """Optimizer::PoseOptimization (UNPINNED restatement over the vendored g2o):
    (Tcw after [4,4] f32, mvbOutlier [n] u8, nInitialCorrespondences - nBad) and, with `lils` (the structural-line edges),
    additionally mvbOutlier_Insec [n_lil] u8."""
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
    p = np.ascontiguousarray(pts, POSE_POINT_DTYPE)
    out = np.zeros(16, np.float32)
    bad = np.zeros(max(len(p), 1), np.uint8)
    if lils is None:
        n = lib().orc_pose_optimization(_p(T), _p(p), len(p), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                                        C.c_float(bf), _p(out), _p(bad))
        return out.reshape(4, 4), bad[: len(p)].copy(), int(n)
    l = np.ascontiguousarray(lils, POSE_LIL_DTYPE)
    lbad = np.zeros(max(len(l), 1), np.uint8)
    n = lib().orc_pose_optimization_lil(_p(T), _p(p), len(p), _p(l), len(l), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                                        C.c_float(cy), C.c_float(bf), _p(out), _p(bad), _p(lbad))
    return out.reshape(4, 4), bad[: len(p)].copy(), int(n), lbad[: len(l)].copy()
