"""Host-side mirror of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:41-83 in the reference) over the
C-ABI.  The reference methods take Frame / KeyFrame / MapPoint objects; here the caller passes the plain
arrays those objects hold (a FrameData view and projected queries) — exactly what crosses the C-ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import FUSE_QUERY_DTYPE, KP_DTYPE, Q_CLAIMS, Q_VALID, QUERY_DTYPE, MatchParams, make_feature_vector, make_frame_view, make_keyframe_view
from .orb import Context, _ptr


@dataclass
class FrameData:
    """What the matchers read from a Frame: mvKeysUn, mvuRight, mDescriptors and the image bounds
    (include/Frame.h:150-290)."""
    kps_un: np.ndarray            # KP_DTYPE [n]
    desc: np.ndarray              # u8 [n,32]
    u_right: np.ndarray | None    # f32 [n] (<=0: none)
    bounds: tuple                 # (mnMinX, mnMinY, mnMaxX, mnMaxY)


class ORBmatcher:
    TH_HIGH = 100  # ORBmatcher.cc:37
    TH_LOW = 50    # ORBmatcher.cc:38
    HISTO_LENGTH = 30

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, *, ctx: Context | None = None, device: int = 0):
        """ORBmatcher(float nnratio=0.6, bool checkOri=true) — ORBmatcher.cc:41-43."""
        if ctx is None:
            cfg = _lib.default_config()
            cfg.device = device
            ctx = Context(cfg)
        self.ctx = ctx
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)

    def DescriptorDistance(self, a: np.ndarray, b: np.ndarray):
        """ORBmatcher.cc:1647-1663; accepts one pair [32] or n pairs [n,32]."""
        a2 = np.ascontiguousarray(np.atleast_2d(a), np.uint8)
        b2 = np.ascontiguousarray(np.atleast_2d(b), np.uint8)
        out = np.zeros(len(a2), np.int32)
        self.ctx.check(_lib.lib().psl_descriptor_distance(self.ctx.handle, _ptr(a2), _ptr(b2), len(a2), _ptr(out)))
        return int(out[0]) if np.ndim(a) == 1 else out

    def _project(self, frame: FrameData, queries, qdesc, claimed, mode, th_dist, ratio, ori):
        fv, keep = make_frame_view(frame.kps_un, frame.u_right, frame.desc, frame.bounds)
        queries = np.ascontiguousarray(queries, QUERY_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, np.uint8)
        cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
        prm = MatchParams(mode, th_dist, ratio, int(ori))
        assign = np.full(fv.n, -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_match_projection(self.ctx.handle, C.byref(fv), _ptr(queries), _ptr(qdesc),
                                                       len(queries), None if cl is None else _ptr(cl), C.byref(prm),
                                                       _ptr(assign), C.byref(nm)))
        return assign, nm.value

    def SearchByProjectionLastFrame(self, cur: FrameData, queries, last_desc, claimed=None):
        """SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) — ORBmatcher.cc:1328-1470.
        `queries` carry what the per-point prologue (:1353-1393) computes."""
        return self._project(cur, queries, last_desc, claimed, 0, self.TH_HIGH, self.mfNNratio,
                             self.mbCheckOrientation)

    def SearchByProjectionMapPoints(self, frame: FrameData, queries, mp_desc, claimed=None):
        """SearchByProjection(Frame &F, const vector<MapPoint*>&, th) — ORBmatcher.cc:45-129."""
        return self._project(frame, queries, mp_desc, claimed, 1, self.TH_HIGH, self.mfNNratio, False)

    def SearchByProjectionKeyFrame(self, cur: FrameData, queries, kf_desc, has_mappoint_cur, ORBdist):
        """SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, ORBdist) — ORBmatcher.cc:1472-1599
        (relocalization, Tracking.cc:2142, 2156).  It is the Frame<-LastFrame form with other settings, so it runs
        through the same entry point (mode 0): no right-coordinate gate (u_right = NULL), every match claims its
        keypoint (:1545-1546 skips any keypoint that holds a MapPoint, old or just assigned) and the threshold is
        ORBdist.  queries: u, v, radius = th * mvScaleFactors[pred], levels pred-1 .. pred+1, angle of pKF's keypoint,
        PSL_Q_VALID for the points that pass :1492-1525."""
        q = np.ascontiguousarray(queries, QUERY_DTYPE).copy()
        q["flags"] = np.where(q["flags"] & Q_VALID, Q_VALID | Q_CLAIMS, 0).astype(np.uint32)
        view = FrameData(cur.kps_un, cur.desc, None, cur.bounds)
        return self._project(view, q, kf_desc, has_mappoint_cur, 0, int(ORBdist), self.mfNNratio,
                             self.mbCheckOrientation)

    def SearchByBoW(self, kf_desc, kf_angle, kf_valid, kf_fv, f_desc, f_angle, f_fv):
        """SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) — ORBmatcher.cc:159-288.
        kf_fv / f_fv: (node_ids ascending, offs, indices) CSR of DBoW2::FeatureVector."""
        kf_desc = np.ascontiguousarray(kf_desc, np.uint8)
        f_desc = np.ascontiguousarray(f_desc, np.uint8)
        kf_angle = np.ascontiguousarray(kf_angle, np.float32)
        f_angle = np.ascontiguousarray(f_angle, np.float32)
        kf_valid = np.ascontiguousarray(kf_valid, np.uint8)
        a, k1 = make_feature_vector(*kf_fv)
        b, k2 = make_feature_vector(*f_fv)
        match = np.full(len(f_desc), -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_match_bow(self.ctx.handle, _ptr(kf_desc), _ptr(kf_angle), _ptr(kf_valid),
                                                len(kf_desc), C.byref(a), _ptr(f_desc), _ptr(f_angle), len(f_desc),
                                                C.byref(b), C.c_float(self.mfNNratio), self.TH_LOW,
                                                int(self.mbCheckOrientation), _ptr(match), C.byref(nm)))
        return match, nm.value


    def SearchForTriangulation(self, kf1, fv1, kf2, fv2, F12, epipole, scale_factors2, level_sigma2_2,
                               bOnlyStereo=False):
        """SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) — ORBmatcher.cc:657-823.
        kf* = (kps_un, u_right, desc, has_mappoint); fv* = CSR FeatureVector; epipole = (ex, ey) of KF1's centre in
        image 2 (:663-670).  Returns (vMatchedPairs [m,2], matches12 [n1], count)."""
        a, k1 = make_keyframe_view(*kf1)
        b, k2 = make_keyframe_view(*kf2)
        f1, k3 = make_feature_vector(*fv1)
        f2, k4 = make_feature_vector(*fv2)
        F = np.ascontiguousarray(F12, np.float32).reshape(9)
        sc = np.ascontiguousarray(scale_factors2, np.float32)
        s2 = np.ascontiguousarray(level_sigma2_2, np.float32)
        m12 = np.full(a.n, -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_match_triangulation(self.ctx.handle, C.byref(a), C.byref(f1), C.byref(b), C.byref(f2),
                                                          _ptr(F), C.c_float(epipole[0]), C.c_float(epipole[1]), _ptr(sc),
                                                          _ptr(s2), len(sc), int(bOnlyStereo),
                                                          int(self.mbCheckOrientation), self.TH_LOW, _ptr(m12), C.byref(nm)))
        idx = np.nonzero(m12 >= 0)[0]
        return np.stack([idx, m12[idx]], 1).astype(np.int64), m12, nm.value


    def FuseSearch(self, kf: FrameData, queries, mp_desc, inv_level_sigma2):
        """The window search of Fuse(pKF, vpMapPoints, th) — ORBmatcher.cc:893-950.  Returns (best_idx [nq] = keypoint
        to fuse with or -1, best_dist [nq]); the replace-or-add bookkeeping (:953-975) is the caller's."""
        fv, keep = make_frame_view(kf.kps_un, kf.u_right, kf.desc, kf.bounds)
        queries = np.ascontiguousarray(queries, FUSE_QUERY_DTYPE)
        qd = np.ascontiguousarray(mp_desc, np.uint8)
        s2 = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
        bi = np.full(len(queries), -1, np.int32)
        bd = np.full(len(queries), 256, np.int32)
        if s2 is not None:
            nlevels = len(s2)
        else:   # no table: nlevels only bounds the octaves
            nlevels = int(kf.kps_un["octave"].max()) + 1 if fv.n else 1
        self.ctx.check(_lib.lib().psl_match_fuse(self.ctx.handle, C.byref(fv), _ptr(queries), _ptr(qd), len(queries),
                                                 None if s2 is None else _ptr(s2), nlevels, self.TH_LOW, _ptr(bi),
                                                 _ptr(bd)))
        return bi, bd

    def FuseSearchSim3(self, kf: FrameData, queries, mp_desc):
        """The window search of Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) — ORBmatcher.cc:1046-1075 (loop closing,
        LoopClosing.cc:599): the pose form without the chi-square gate."""
        return self.FuseSearch(FrameData(kf.kps_un, kf.desc, None, kf.bounds), queries, mp_desc, None)

    def SearchByProjectionSim3(self, kf: FrameData, queries, mp_desc, matched):
        """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) — ORBmatcher.cc:290-403 (LoopClosing.cc:375) after the
        projection.  queries: psl_fuse_query records (u, v, radius = th * mvScaleFactors[pred], pred_level, PSL_Q_VALID);
        matched[i] = vpMatched[i] != NULL.  Mode 0 of the projection matcher with the settings listed in the header.
        Returns (assign [n] = query written into vpMatched[i] or -1, nmatches)."""
        qf = np.ascontiguousarray(queries, FUSE_QUERY_DTYPE)
        q = np.zeros(len(qf), QUERY_DTYPE)
        q["u"], q["v"], q["radius"] = qf["u"], qf["v"], qf["radius"]
        q["min_level"], q["max_level"] = qf["pred_level"] - 1, qf["pred_level"]
        q["flags"] = np.where(qf["flags"] & Q_VALID, Q_VALID | Q_CLAIMS, 0).astype(np.uint32)
        view = FrameData(kf.kps_un, kf.desc, None, kf.bounds)
        return self._project(view, q, mp_desc, matched, 0, self.TH_LOW, self.mfNNratio, False)

    def SearchByBoWKeyFrames(self, desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2):
        """SearchByBoW(pKF1, pKF2, vpMatches12) — ORBmatcher.cc:522-655.  valid*: the keypoint holds a good MapPoint;
        fv*: CSR FeatureVector.  Returns (matches12 [n1] = KF2 index or -1, nmatches)."""
        desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
        angle1, angle2 = np.ascontiguousarray(angle1, np.float32), np.ascontiguousarray(angle2, np.float32)
        valid1, valid2 = np.ascontiguousarray(valid1, np.uint8), np.ascontiguousarray(valid2, np.uint8)
        a, k1 = make_feature_vector(*fv1)
        b, k2 = make_feature_vector(*fv2)
        m12 = np.full(len(desc1), -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_match_bow_kf(self.ctx.handle, _ptr(desc1), _ptr(angle1), _ptr(valid1), len(desc1),
                                                   C.byref(a), _ptr(desc2), _ptr(angle2), _ptr(valid2), len(desc2),
                                                   C.byref(b), C.c_float(self.mfNNratio), self.TH_LOW,
                                                   int(self.mbCheckOrientation), _ptr(m12), C.byref(nm)))
        return m12, nm.value

    def SearchBySim3(self, kf1: FrameData, kf2: FrameData, q12, mp_desc1, q21, mp_desc2):
        """SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) — ORBmatcher.cc:1102-1326 after the projections
        (one psl_fuse_query per keypoint of either keyframe).  Returns (matches12 [n1], nFound)."""
        a, ka = make_frame_view(kf1.kps_un, None, kf1.desc, kf1.bounds)
        b, kb = make_frame_view(kf2.kps_un, None, kf2.desc, kf2.bounds)
        q12 = np.ascontiguousarray(q12, FUSE_QUERY_DTYPE)
        q21 = np.ascontiguousarray(q21, FUSE_QUERY_DTYPE)
        if len(q12) != a.n or len(q21) != b.n:
            raise ValueError("SearchBySim3 takes one query per keypoint")
        d1, d2 = np.ascontiguousarray(mp_desc1, np.uint8), np.ascontiguousarray(mp_desc2, np.uint8)
        m12 = np.full(a.n, -1, np.int32)
        nf = C.c_int32()
        self.ctx.check(_lib.lib().psl_match_sim3(self.ctx.handle, C.byref(a), C.byref(b), _ptr(q12), _ptr(d1), _ptr(q21),
                                                 _ptr(d2), self.TH_HIGH, _ptr(m12), C.byref(nf)))
        return m12, nf.value

    def SearchForInitialization(self, kps1_un, desc1, f2: FrameData, vbPrevMatched, windowSize=100):
        """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — ORBmatcher.cc:405-520.
        Returns (vnMatches12 [n1], nmatches, vbPrevMatched after the update [n1, 2])."""
        k1 = np.ascontiguousarray(kps1_un, _lib.KP_DTYPE)
        d1 = np.ascontiguousarray(desc1, np.uint8)
        pm = np.array(vbPrevMatched, np.float32, copy=True).reshape(-1, 2)
        b, kb = make_frame_view(f2.kps_un, None, f2.desc, f2.bounds)
        m12 = np.full(len(k1), -1, np.int32)
        nm = C.c_int32()
        self.ctx.check(_lib.lib().psl_match_initialization(self.ctx.handle, _ptr(k1), _ptr(d1), len(k1), _ptr(pm),
                                                           C.byref(b), int(windowSize), C.c_float(self.mfNNratio),
                                                           self.TH_LOW, int(self.mbCheckOrientation), _ptr(m12),
                                                           C.byref(nm)))
        return m12, nm.value, pm


def hamming_knn2(ctx: Context, q: np.ndarray, t: np.ndarray):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) — LSDmatcher.cpp:361-362."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), -1, np.int32)
    ctx.check(_lib.lib().psl_hamming_knn2(ctx.handle, _ptr(q), len(q), _ptr(t), len(t), _ptr(idx), _ptr(dist)))
    return idx, dist
