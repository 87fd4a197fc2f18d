"""Host-side mirrors of ORB_SLAM2::LSDmatcher (add_inc/LSDmatcher.h:20-75 in the reference) and
ORB_SLAM2::InsectLineMatch (add_inc/InsectlineMatch.h:9-17) over the C-ABI.  The reference methods take
Frame / KeyFrame / MapLine / InsectLine objects; here the caller passes the plain arrays those objects hold —
exactly what crosses the C-ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import JUNCTION_DTYPE, KEYLINE_DTYPE, LINE_FUSE_QUERY_DTYPE, LINE_QUERY_DTYPE, make_line_frame_view
from .orb import Context, _ptr


@dataclass
class LineFrameData:
    """What the line matchers read from a Frame: mvKeylinesUn, mLdesc, mvKeyLineFunctions, mvLines3D and the
    image bounds (include/Frame.h)."""
    kl_un: np.ndarray             # KEYLINE_DTYPE [n]
    ldesc: np.ndarray             # u8 [n,32]
    lineeq: np.ndarray            # f64 [n,3]
    lines3d: np.ndarray | None    # f64 [n,6] (first xyz, second xyz)
    bounds: tuple                 # (mnMinX, mnMinY, mnMaxX, mnMaxY)


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


class LSDmatcher:
    TH_HIGH = 80  # LSDmatcher.cpp:12-14
    TH_LOW = 50
    HISTO_LENGTH = 30

    def __init__(self, nnratio: float = 0.95, checkOri: bool = True, *, ctx: Context | None = None, device: int = 0):
        """LSDmatcher(float nnratio=0.95, bool checkOri=true) — LSDmatcher.cpp:16-18."""
        if ctx is None:
            cfg = _lib.default_config()
            cfg.device = device
            ctx = Context(cfg)
        self.ctx = ctx
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)

    def match(self, desc1, desc2, nnr):
        """match / matchNNR — LSDmatcher.cpp:354-413.  Returns (matches_12, count)."""
        d1, d2 = _u8(desc1), _u8(desc2)
        out = np.full(len(d1), -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_line_match_nnr(self.ctx.handle, _ptr(d1), len(d1), _ptr(d2), len(d2),
                                                     C.c_float(nnr), _ptr(out), C.byref(nm)))
        return out, nm.value

    matchNNR = match

    def SearchByGeomNApearance(self, cur: LineFrameData, last: LineFrameData, has_mapline_last, desc_th):
        """SearchByGeomNApearance(CurrentFrame, LastFrame, desc_th) — LSDmatcher.cpp:36-110.
        Returns (assign_cur [n_cur] = last-frame line index or -1, count)."""
        kl_l, kl_c = np.ascontiguousarray(last.kl_un, KEYLINE_DTYPE), np.ascontiguousarray(cur.kl_un, KEYLINE_DTYPE)
        d_l, d_c, has = _u8(last.ldesc), _u8(cur.ldesc), _u8(has_mapline_last)
        b = np.ascontiguousarray(cur.bounds, np.float32)
        out = np.full(len(kl_c), -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_line_search_geom(self.ctx.handle, _ptr(kl_l), _ptr(d_l), _ptr(has), len(kl_l),
                                                       _ptr(kl_c), _ptr(d_c), len(kl_c), _ptr(b), C.c_float(desc_th),
                                                       _ptr(out), C.byref(nm)))
        return out, nm.value

    def FrameBFMatch(self, ldesc1, ldesc2, TH):
        """FrameBFMatch(ldesc1, ldesc2, LineMatches, TH) — LSDmatcher.cpp:492-516."""
        d1, d2 = _u8(ldesc1), _u8(ldesc2)
        out = np.full(len(d1), -1, np.int32)
        self.ctx.check(_lib.lib().psl_line_frame_bf_match(self.ctx.handle, _ptr(d1), len(d1), _ptr(d2), len(d2),
                                                          C.c_float(self.mfNNratio), C.c_float(TH), _ptr(out)))
        return out

    def SearchDouble(self, ldesc1, ldesc2):
        """SearchDouble(InitialFrame, CurrentFrame, LineMatches) — LSDmatcher.cpp:462-490."""
        d1, d2 = _u8(ldesc1), _u8(ldesc2)
        out = np.full(len(d1), -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_line_search_double(self.ctx.handle, _ptr(d1), len(d1), _ptr(d2), len(d2),
                                                         C.c_float(self.mfNNratio), C.c_float(self.TH_LOW), _ptr(out),
                                                         C.byref(nm)))
        return out, nm.value

    def SearchForTriangulation(self, ldesc1, has_mapline1, ldesc2, has_mapline2, as_pairs=True, isDouble=False):
        """SearchForTriangulation(pKF1, pKF2, vMatchedPairs) — LSDmatcher.cpp:705-741 (as_pairs: TH_LOW, mutual) and the
        vector<int> form :743-779 (TH_HIGH, mutual only when isDouble).  Returns (matches12 [n1], count)."""
        d1, d2, m1, m2 = _u8(ldesc1), _u8(ldesc2), _u8(has_mapline1), _u8(has_mapline2)
        out = np.full(len(d1), -1, np.int32)
        nm = C.c_int32()
        th = self.TH_LOW if as_pairs else self.TH_HIGH
        self.ctx.check(_lib.lib().psl_line_search_triangulation(self.ctx.handle, _ptr(d1), _ptr(m1), len(d1), _ptr(d2),
                                                                _ptr(m2), len(d2), C.c_float(self.mfNNratio),
                                                                C.c_float(th), int(as_pairs or isDouble), _ptr(out),
                                                                C.byref(nm)))
        return out, nm.value

    def SearchForTriangulationNew(self, kf1, kf2, F21, F12, isDouble=False):
        """SearchForTriangulationNew(pKF1, pKF2, vMatchedPairs, isDouble) — LSDmatcher.cpp:783-824 over FrameBFMatchNew
        (:518-581) and mutualOverlap (:583-658).  kf* = (mvKeyLines, mLineDescriptors, mvKeyLineFunctions [n,3], has MapLine
        [n]); F21 / F12 = ComputeF12(pKF2, pKF1) / ComputeF12(pKF1, pKF2), 3x3 float.  Returns (vMatchedPairs [n1], count)."""
        k1, k2 = np.ascontiguousarray(kf1[0], KEYLINE_DTYPE), np.ascontiguousarray(kf2[0], KEYLINE_DTYPE)
        d1, d2, m1, m2 = _u8(kf1[1]), _u8(kf2[1]), _u8(kf1[3]), _u8(kf2[3])
        f1 = np.ascontiguousarray(kf1[2], np.float64).reshape(-1, 3)
        f2 = np.ascontiguousarray(kf2[2], np.float64).reshape(-1, 3)
        A = np.ascontiguousarray(F21, np.float32).reshape(9)
        B = np.ascontiguousarray(F12, np.float32).reshape(9)
        out = np.full(len(d1), -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_line_search_triangulation_new(
            self.ctx.handle, _ptr(k1), _ptr(d1), _ptr(f1), _ptr(m1), len(d1), _ptr(k2), _ptr(d2), _ptr(f2), _ptr(m2),
            len(d2), _ptr(A), _ptr(B), C.c_float(self.mfNNratio), C.c_float(self.TH_LOW), int(isDouble), _ptr(out),
            C.byref(nm)))
        return out, nm.value

    def Fuse(self, keylines, kf_descriptors, queries, map_line_desc, th_cos=0.998):
        """Window search of Fuse(pKF, vpMapLines, th) — LSDmatcher.cpp:847-984: per projected MapLine (queries:
        LINE_FUSE_QUERY_DTYPE) the line of KeyFrame::GetLinesInArea (KeyFrame.cc:857-891) at level pred-1..pred with
        the smallest Hamming distance to pKF->mDescriptors.row(idx) (:938).  Returns (best_idx [nq] or -1 when the
        distance exceeds TH_LOW, best_dist [nq]); the replace-or-add bookkeeping stays with the caller."""
        kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
        kd, qd = _u8(kf_descriptors), _u8(map_line_desc)
        q = np.ascontiguousarray(queries, LINE_FUSE_QUERY_DTYPE)
        bi, bd = np.full(len(q), -1, np.int32), np.full(len(q), 256, np.int32)
        self.ctx.check(_lib.lib().psl_line_fuse(self.ctx.handle, _ptr(kl), len(kl), _ptr(kd), len(kd), _ptr(q), _ptr(qd),
                                                len(q), C.c_float(th_cos), int(self.TH_LOW), _ptr(bi), _ptr(bd)))
        return bi, bd

    def _project(self, frame: LineFrameData, queries, qdesc, claimed, mode):
        fv, keep = make_line_frame_view(frame.kl_un, frame.ldesc, frame.lineeq, frame.lines3d, frame.bounds)
        queries = np.ascontiguousarray(queries, LINE_QUERY_DTYPE)
        qdesc = _u8(qdesc)
        cl = None if claimed is None else _u8(claimed)
        assign = np.full(fv.n, -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_line_match_projection(self.ctx.handle, C.byref(fv), _ptr(queries), _ptr(qdesc),
                                                            len(queries), None if cl is None else _ptr(cl), mode,
                                                            C.c_float(self.mfNNratio), _ptr(assign), C.byref(nm)))
        return assign, nm.value

    def SearchByProjectionLastFrame(self, cur: LineFrameData, queries, last_desc, claimed=None):
        """SearchByProjection(CurrentFrame, LastFrame, th) — LSDmatcher.cpp:112-215."""
        return self._project(cur, queries, last_desc, claimed, 0)

    def SearchByProjectionMapLines(self, frame: LineFrameData, queries, ml_desc, claimed=None):
        """SearchByProjection(F, vpMapLines, eval_orient, th) — LSDmatcher.cpp:260-352."""
        return self._project(frame, queries, ml_desc, claimed, 1)


class InsectLineMatch:
    def __init__(self, dTh: float = 0.1, aTh: float = 0.86, *, ctx: Context | None = None, device: int = 0):
        """InsectLineMatch(float dTh=0.1, float aTh=0.86) — InsectlineMatch.cpp:8."""
        if ctx is None:
            cfg = _lib.default_config()
            cfg.device = device
            ctx = Context(cfg)
        self.ctx, self.dTh, self.aTh = ctx, float(dTh), float(aTh)

    def _run(self, planes_cam, pts, Tcw, map_planes, map_bad, mode):
        planes_cam = np.ascontiguousarray(planes_cam, np.float32).reshape(-1, 4)
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 15)
        Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(4, 4)
        map_planes = np.ascontiguousarray(map_planes, np.float32).reshape(-1, 4)
        bad = None if map_bad is None else _u8(map_bad)
        out = np.full(len(planes_cam), -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_plane_assoc(self.ctx.handle, _ptr(planes_cam), _ptr(pts), len(planes_cam), _ptr(Tcw),
                                                  _ptr(map_planes), None if bad is None else _ptr(bad), len(map_planes),
                                                  C.c_float(self.dTh), C.c_float(self.aTh), mode, _ptr(out), C.byref(nm)))
        return out, nm.value

    def SearchMapInsectline(self, planes_cam, pts, Tcw, map_planes, map_bad=None):
        """SearchMapInsectline(Frame&, vpMapInsectline) — InsectlineMatch.cpp:9-60."""
        return self._run(planes_cam, pts, Tcw, map_planes, map_bad, 0)

    def AssociatePlanesByBoundary(self, planes_cam, pts, Tcw, map_planes):
        """Map::AssociatePlanesByBoundary(Frame&, dTh, aTh) — src/Map.cc:204-272 (the live twin)."""
        return self._run(planes_cam, pts, Tcw, map_planes, None, 1)


def plane_hypotheses(ctx: Context, kl_un, line_eq, lines3d, junctions, cap: int | None = None):
    """The plane hypotheses Frame::ExtractLSD builds from coplanar intersecting line pairs (Frame.cc:512-645, OldPlane
    :474-487).  junctions: JUNCTION_DTYPE (intersection_lines_plane).  Returns (mvle_l [nj,6] f64, mvPlanes [np,4] f32,
    mvPlaneNormal [np,3] f64, junction index of each plane [np])."""
    kl = np.ascontiguousarray(kl_un, KEYLINE_DTYPE)
    eq = np.ascontiguousarray(line_eq, np.float32).reshape(-1, 3)
    l3 = np.ascontiguousarray(lines3d, np.float64).reshape(-1, 6)
    js = np.ascontiguousarray(junctions, JUNCTION_DTYPE)
    nj = len(js)
    cap = nj if cap is None else cap
    le = np.zeros((max(nj, 1), 6))
    pl, nr, ow = np.zeros((max(cap, 1), 4), np.float32), np.zeros((max(cap, 1), 3)), np.zeros(max(cap, 1), np.int32)
    n = C.c_int32()
    ctx.check(_lib.lib().psl_plane_hypotheses(ctx.handle, _ptr(kl), _ptr(eq), _ptr(l3), len(kl), _ptr(js), nj, _ptr(le),
                                              _ptr(pl), _ptr(nr), _ptr(ow), cap, C.byref(n)))
    return le[:nj], pl[: n.value], nr[: n.value], ow[: n.value]


def line_junctions(ctx: Context, kl_un, lines3d, img_w: int, img_h: int, radius: float = 20.0,
                   fan_thr: float = float(np.float32(0.25 * np.pi)), cap: int = 4096):
    """The junction detection of Frame::ExtractLSD (Frame.cc:504-507): CPartiallyRecoverConnectivity
    (PartiallyRecoverConnectivity.cpp:14-133) + Frame::convertFansToKeyLines (Frame.cc:426-472).  lines3d = mvLines3D
    [n,6] or None (fans only).  Returns (fans [m,4] f32 rows (x, y, i, j), intersection_lines_plane as JUNCTION_DTYPE)."""
    kl = np.ascontiguousarray(kl_un, KEYLINE_DTYPE)
    l3 = None if lines3d is None else np.ascontiguousarray(lines3d, np.float64).reshape(-1, 6)
    fans = np.zeros((max(cap, 1), 4), np.float32)
    js = np.zeros(max(cap, 1), JUNCTION_DTYPE)
    nf, nj = C.c_int32(), C.c_int32()
    ctx.check(_lib.lib().psl_line_junctions(ctx.handle, _ptr(kl), None if l3 is None else _ptr(l3), len(kl), int(img_w),
                                            int(img_h), C.c_float(radius), C.c_float(fan_thr), _ptr(fans),
                                            None if l3 is None else _ptr(js), cap, C.byref(nf), C.byref(nj)))
    return fans[: nf.value], js[: nj.value]


def lines_3d(ctx: Context, kl_un, depth_f32, fx, fy, cx, cy, seed: int = 0):
    """Frame::isLineGood (Frame.cc:662-750): (mvLines3D [n,6] f64, mvLineEq [n,3] f32) of the KeyLines from the CV_32F
    depth image; `seed` pins random_unique's rand() (see include/psl_frontend.h)."""
    kl = np.ascontiguousarray(kl_un, KEYLINE_DTYPE)
    dep = np.ascontiguousarray(depth_f32, np.float32)
    n = len(kl)
    l3, eq = np.zeros((max(n, 1), 6)), np.zeros((max(n, 1), 3), np.float32)
    ctx.check(_lib.lib().psl_lines_3d(ctx.handle, _ptr(kl), n, _ptr(dep), dep.shape[1], dep.shape[0], C.c_float(fx),
                                      C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_uint32(seed), _ptr(l3), _ptr(eq)))
    return l3[:n], eq[:n]
