"""TEST INFRASTRUCTURE — golden-vector generator for the point matchers.

Independent Python restatement (written from the reference source, not from oracle/c) of
ORBmatcher::SearchByProjection (src/ORBmatcher.cc:45-129 and :1328-1470), SearchByBoW (:159-288),
ComputeThreeMaxima (:1601-1642), DescriptorDistance (:1647-1663) and Frame::AssignFeaturesToGrid /
GetFeaturesInArea / PosInGrid (src/Frame.cc:269-284, 985-1050).  cv2.BFMatcher supplies the real
knnMatch used by LSDmatcher::matchNNR (add_src/LSDmatcher.cpp:354-376).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
COLS, ROWS = 64, 48
HISTO = 30


def c_round(x) -> int:
    """C round(): half away from zero."""
    x = float(x)
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def popcount_dist(a: np.ndarray, b: np.ndarray) -> int:
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


class FrameView:
    def __init__(self, kps_un, octave, u_right, desc, min_x, min_y, max_x, max_y):
        self.x = kps_un[:, 0].astype(F32)
        self.y = kps_un[:, 1].astype(F32)
        self.angle = kps_un[:, 3].astype(F32)
        self.octave = octave.astype(np.int32)
        self.u_right = None if u_right is None else u_right.astype(F32)
        self.desc = desc
        self.n = len(desc)
        self.min_x, self.min_y, self.max_x, self.max_y = F32(min_x), F32(min_y), F32(max_x), F32(max_y)
        self.w_inv = F32(COLS) / F32(self.max_x - self.min_x)  # Frame.cc:163-164
        self.h_inv = F32(ROWS) / F32(self.max_y - self.min_y)
        self.grid = [[[] for _ in range(ROWS)] for _ in range(COLS)]
        for i in range(self.n):  # AssignFeaturesToGrid
            px = c_round(F32(self.x[i] - self.min_x) * self.w_inv)
            py = c_round(F32(self.y[i] - self.min_y) * self.h_inv)
            if 0 <= px < COLS and 0 <= py < ROWS:
                self.grid[px][py].append(i)

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        x, y, r = F32(x), F32(y), F32(r)
        out = []
        c0 = max(0, int(math.floor(F32(F32(x - self.min_x) - r) * self.w_inv)))
        if c0 >= COLS:
            return out
        c1 = min(COLS - 1, int(math.ceil(F32(F32(x - self.min_x) + r) * self.w_inv)))
        if c1 < 0:
            return out
        r0 = max(0, int(math.floor(F32(F32(y - self.min_y) - r) * self.h_inv)))
        if r0 >= ROWS:
            return out
        r1 = min(ROWS - 1, int(math.ceil(F32(F32(y - self.min_y) + r) * self.h_inv)))
        if r1 < 0:
            return out
        check = (min_level > 0) or (max_level >= 0)
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for i in self.grid[ix][iy]:
                    if check:
                        if self.octave[i] < min_level:
                            continue
                        if max_level >= 0 and self.octave[i] > max_level:
                            continue
                    if abs(F32(self.x[i] - x)) < r and abs(F32(self.y[i] - y)) < r:
                        out.append(i)
        return out


def three_maxima(sizes):
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > m1:
            m3, m2, m1 = m2, m1, s
            i3, i2, i1 = i2, i1, i
        elif s > m2:
            m3, m2 = m2, s
            i3, i2 = i2, i
        elif s > m3:
            m3, i3 = s, i
    if m2 < F32(0.1) * F32(m1):
        i2 = i3 = -1
    elif m3 < F32(0.1) * F32(m1):
        i3 = -1
    return i1, i2, i3


def rot_bin(a1, a2):
    factor = F32(1.0) / F32(HISTO)
    rot = F32(F32(a1) - F32(a2))
    if rot < 0.0:
        rot = F32(rot + F32(360.0))
    b = c_round(F32(rot * factor))
    return 0 if b == HISTO else b


def search_by_projection(fv: FrameView, queries, qdesc, claimed_in, mode, th_dist=100, nn_ratio=0.6, check_ori=True):
    """queries: structured array (u,v,radius,min_level,max_level,u_right,angle,flags)."""
    assign = np.full(fv.n, -1, np.int32)
    claimed = np.zeros(fv.n, bool) if claimed_in is None else claimed_in.astype(bool).copy()
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for qi, q in enumerate(queries):
        if not (q["flags"] & 1):
            continue
        cand = fv.features_in_area(q["u"], q["v"], q["radius"], int(q["min_level"]), int(q["max_level"]))
        if not cand:
            continue
        best, best_i, best_l, best2, best2_l = 256, -1, -1, 256, -1
        for i2 in cand:
            if claimed[i2]:
                continue
            if fv.u_right is not None and fv.u_right[i2] > 0:
                if abs(F32(F32(q["u_right"]) - fv.u_right[i2])) > F32(q["radius"]):
                    continue
            d = popcount_dist(qdesc[qi], fv.desc[i2])
            if d < best:
                best2, best2_l = best, best_l
                best, best_l, best_i = d, int(fv.octave[i2]), i2
            elif mode == 1 and d < best2:
                best2, best2_l = d, int(fv.octave[i2])
        if best <= th_dist:
            if mode == 1 and best_l == best2_l and F32(best) > F32(nn_ratio) * F32(best2):
                continue
            assign[best_i] = qi
            claimed[best_i] = bool(q["flags"] & 2)
            nm += 1
            if mode == 0 and check_ori:
                hist[rot_bin(q["angle"], fv.angle[best_i])].append(best_i)
    if mode == 0 and check_ori:
        keep = three_maxima([len(h) for h in hist])
        for b in range(HISTO):
            if b not in keep:
                for i in hist[b]:
                    assign[i] = -1
                    nm -= 1
    return assign, nm


def search_by_projection_keyframe(fv: FrameView, queries, kf_desc, has_mappoint_cur, orb_dist, check_ori=True):
    """SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, ORBdist) — ORBmatcher.cc:1472-1599,
    written from that overload on its own (not through search_by_projection): candidates of
    GetFeaturesInArea(u, v, radius, pred-1, pred+1) that hold no MapPoint yet (:1545-1546, which includes the ones
    this call has just filled), smallest distance, accept <= ORBdist, rotation histogram on pKF's keypoint angle.
    queries: (u, v, radius, min_level = pred-1, max_level = pred+1, -, angle = pKF->mvKeysUn[i].angle, flags & 1)."""
    mp = [None if not h else -2 for h in has_mappoint_cur]      # mvpMapPoints: -2 = a MapPoint from before the call
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for qi, q in enumerate(queries):
        if not (q["flags"] & 1):
            continue
        cand = fv.features_in_area(q["u"], q["v"], q["radius"], int(q["min_level"]), int(q["max_level"]))
        if not cand:
            continue
        best, best_i = 256, -1
        for i2 in cand:
            if mp[i2] is not None:
                continue
            d = popcount_dist(kf_desc[qi], fv.desc[i2])
            if d < best:
                best, best_i = d, i2
        if best <= orb_dist:
            mp[best_i] = qi
            nm += 1
            if check_ori:
                hist[rot_bin(q["angle"], fv.angle[best_i])].append(best_i)
    if check_ori:
        keep = three_maxima([len(h) for h in hist])
        for b in range(HISTO):
            if b not in keep:
                for i in hist[b]:
                    mp[i] = None
                    nm -= 1
    assign = np.array([m if (m is not None and m >= 0) else -1 for m in mp], np.int32)
    return assign, nm


def search_by_bow(kf_desc, kf_angle, kf_valid, kf_fv, f_desc, f_angle, f_fv, nn_ratio=0.7, th_low=50, check_ori=True):
    """kf_fv / f_fv: dict node_id -> list of indices (DBoW2::FeatureVector)."""
    nf = len(f_desc)
    match = np.full(nf, -1, np.int32)
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for node in sorted(set(kf_fv) & set(f_fv)):
        for ikf in kf_fv[node]:
            if not kf_valid[ikf]:
                continue
            b1, bi, b2 = 256, -1, 256
            for jf in f_fv[node]:
                if match[jf] >= 0:
                    continue
                d = popcount_dist(kf_desc[ikf], f_desc[jf])
                if d < b1:
                    b2, b1, bi = b1, d, jf
                elif d < b2:
                    b2 = d
            if b1 <= th_low and F32(b1) < F32(nn_ratio) * F32(b2):
                match[bi] = ikf
                if check_ori:
                    hist[rot_bin(kf_angle[ikf], f_angle[bi])].append(bi)
                nm += 1
    if check_ori:
        keep = three_maxima([len(h) for h in hist])
        for b in range(HISTO):
            if b not in keep:
                for i in hist[b]:
                    match[i] = -1
                    nm -= 1
    return match, nm


def knn2_cv2(q, t):
    import cv2
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, 2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), -1, np.int32)
    for i, row in enumerate(m):
        for k, e in enumerate(row):
            idx[i, k], dist[i, k] = e.trainIdx, int(e.distance)
    return idx, dist


def search_for_triangulation(kps1, ur1, desc1, has_mp1, fv1, kps2, ur2, desc2, has_mp2, fv2, F12, ex, ey, scale2,
                             sigma2_2, only_stereo=False, check_ori=True, th_low=50):
    """ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:657-823) + CheckDistEpipolarLine (:140-157).
    kps*: structured keypoints; fv*: dict node -> list of indices; F12 float32 3x3.  Returns (matches12, count)."""
    n1 = len(kps1)
    m12 = np.full(n1, -1, np.int32)
    hist = [[] for _ in range(HISTO)]
    nm = 0
    F = F12.astype(F32)
    for node in sorted(set(fv1) & set(fv2)):
        for idx1 in fv1[node]:
            if has_mp1[idx1]:
                continue
            st1 = ur1[idx1] >= 0
            if only_stereo and not st1:
                continue
            x1, y1 = F32(kps1["x"][idx1]), F32(kps1["y"][idx1])
            best, bidx = th_low, -1
            for idx2 in fv2[node]:
                if has_mp2[idx2]:   # vbMatched2 is never set by the reference
                    continue
                st2 = ur2[idx2] >= 0
                if only_stereo and not st2:
                    continue
                d = popcount_dist(desc1[idx1], desc2[idx2])
                if d > th_low or d > best:
                    continue
                x2, y2, o2 = F32(kps2["x"][idx2]), F32(kps2["y"][idx2]), int(kps2["octave"][idx2])
                if not st1 and not st2:
                    dx, dy = F32(F32(ex) - x2), F32(F32(ey) - y2)
                    if F32(F32(dx * dx) + F32(dy * dy)) < F32(F32(100) * F32(scale2[o2])):
                        continue
                a = F32(F32(F32(x1 * F[0, 0]) + F32(y1 * F[1, 0])) + F[2, 0])
                b = F32(F32(F32(x1 * F[0, 1]) + F32(y1 * F[1, 1])) + F[2, 1])
                c = F32(F32(F32(x1 * F[0, 2]) + F32(y1 * F[1, 2])) + F[2, 2])
                num = F32(F32(F32(a * x2) + F32(b * y2)) + c)
                den = F32(F32(a * a) + F32(b * b))
                if den == 0:
                    continue
                dsqr = F32(F32(num * num) / den)
                if float(dsqr) < 3.84 * float(F32(sigma2_2[o2])):
                    bidx, best = idx2, d
            if bidx >= 0:
                m12[idx1] = bidx
                nm += 1
                if check_ori:
                    hist[rot_bin(kps1["angle"][idx1], kps2["angle"][bidx])].append(idx1)
    if check_ori:
        keep = three_maxima([len(h) for h in hist])
        for i in range(HISTO):
            if i in keep:
                continue
            for idx in hist[i]:
                m12[idx] = -1
                nm -= 1
    return m12, nm


def fuse_search(fv: FrameView, queries, qdesc, inv_sigma2, th_low=50):
    """The window search of ORBmatcher::Fuse (src/ORBmatcher.cc:893-950).  queries: structured (u, v, u_right, radius,
    pred_level, flags).  Returns (best_idx, best_dist)."""
    nq = len(queries)
    bi = np.full(nq, -1, np.int32)
    bd = np.full(nq, 256, np.int32)
    for q in range(nq):
        Q = queries[q]
        if not (Q["flags"] & 1):
            continue
        best, bidx = 256, -1
        for i in fv.features_in_area(Q["u"], Q["v"], Q["radius"]):
            lvl = int(fv.octave[i])
            if lvl < Q["pred_level"] - 1 or lvl > Q["pred_level"]:
                continue
            ex, ey = F32(F32(Q["u"]) - fv.x[i]), F32(F32(Q["v"]) - fv.y[i])
            if fv.u_right is not None and fv.u_right[i] >= 0:
                er = F32(F32(Q["u_right"]) - fv.u_right[i])
                e2 = F32(F32(F32(ex * ex) + F32(ey * ey)) + F32(er * er))
                if float(F32(e2 * F32(inv_sigma2[lvl]))) > 7.8:
                    continue
            else:
                e2 = F32(F32(ex * ex) + F32(ey * ey))
                if float(F32(e2 * F32(inv_sigma2[lvl]))) > 5.99:
                    continue
            d = popcount_dist(qdesc[q], fv.desc[i])
            if d < best:
                best, bidx = d, i
        bd[q] = best
        if best <= th_low:
            bi[q] = bidx
    return bi, bd


# ---------------------------------------------------------------------------------------------------------------
# Loop-closing / initialisation matchers ("next" row N1, second batch).  Each function below is written from its own
# overload in src/ORBmatcher.cc, not through the functions above.
# ---------------------------------------------------------------------------------------------------------------
def search_by_bow_kf(desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2, nn_ratio=0.75, th_low=50, check_ori=True):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) — src/ORBmatcher.cc:522-655.  valid*: the keypoint holds a
    MapPoint that is not bad; fv*: dict node -> list of indices.  Returns (matches12 [n1] = KF2 index or -1, count)."""
    n1, n2 = len(desc1), len(desc2)
    m12 = np.full(n1, -1, np.int32)
    matched2 = np.zeros(n2, bool)
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for node in sorted(set(fv1) & set(fv2)):
        for idx1 in fv1[node]:
            if not valid1[idx1]:
                continue
            b1, bi, b2 = 256, -1, 256
            for idx2 in fv2[node]:
                if matched2[idx2] or not valid2[idx2]:
                    continue
                d = popcount_dist(desc1[idx1], desc2[idx2])
                if d < b1:
                    b2, b1, bi = b1, d, idx2
                elif d < b2:
                    b2 = d
            if b1 < th_low and F32(b1) < F32(nn_ratio) * F32(b2):      # strict `<TH_LOW` here (:592), unlike :224
                m12[idx1] = bi
                matched2[bi] = True
                if check_ori:
                    hist[rot_bin(angle1[idx1], angle2[bi])].append(idx1)
                nm += 1
    if check_ori:
        keep = three_maxima([len(h) for h in hist])
        for b in range(HISTO):
            if b not in keep:
                for i in hist[b]:
                    m12[i] = -1
                    nm -= 1
    return m12, nm


def _window_best(fv: FrameView, Q, qd):
    """min-distance keypoint of KeyFrame::GetFeaturesInArea(u, v, radius) at level pred-1..pred (first one on ties)."""
    best, bidx = 2 ** 31 - 1, -1
    for i in fv.features_in_area(Q["u"], Q["v"], Q["radius"]):
        lvl = int(fv.octave[i])
        if lvl < Q["pred_level"] - 1 or lvl > Q["pred_level"]:
            continue
        d = popcount_dist(qd, fv.desc[i])
        if d < best:
            best, bidx = d, i
    return best, bidx


def search_by_sim3(fv1: FrameView, fv2: FrameView, q12, qdesc1, q21, qdesc2, th_high=100):
    """ORBmatcher::SearchBySim3 — src/ORBmatcher.cc:1102-1326 after the projections: q12[i1] = MapPoint of KF1 keypoint
    i1 projected into KF2 (flags & 1 = it exists, is not bad, not already matched and passed the depth / image /
    distance tests), q21 likewise.  Returns (matches12 [n1] = KF2 index or -1, nFound)."""
    n1, n2 = len(q12), len(q21)
    m1 = np.full(n1, -1, np.int32)
    m2 = np.full(n2, -1, np.int32)
    for i1 in range(n1):
        if q12[i1]["flags"] & 1:
            best, bidx = _window_best(fv2, q12[i1], qdesc1[i1])
            if best <= th_high:
                m1[i1] = bidx
    for i2 in range(n2):
        if q21[i2]["flags"] & 1:
            best, bidx = _window_best(fv1, q21[i2], qdesc2[i2])
            if best <= th_high:
                m2[i2] = bidx
    out = np.full(n1, -1, np.int32)
    found = 0
    for i1 in range(n1):
        idx2 = m1[i1]
        if idx2 >= 0 and m2[idx2] == i1:
            out[i1] = idx2
            found += 1
    return out, found


def fuse_search_sim3(fv: FrameView, queries, qdesc, th_low=50):
    """The window search of ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) — src/ORBmatcher.cc:1046-1075: as the
    pose form but without the chi-square gate.  Returns (best_idx, best_dist)."""
    nq = len(queries)
    bi = np.full(nq, -1, np.int32)
    bd = np.full(nq, 256, np.int32)
    for q in range(nq):
        if not (queries[q]["flags"] & 1):
            continue
        best, bidx = _window_best(fv, queries[q], qdesc[q])
        if bidx >= 0:
            bd[q] = best
            if best <= th_low:
                bi[q] = bidx
    return bi, bd


def search_by_projection_sim3(fv: FrameView, queries, qdesc, matched_in, th_low=50):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) — src/ORBmatcher.cc:290-403 after the
    projection: queries (u, v, radius, pred_level, flags & 1); matched_in[i] = vpMatched[i] != NULL before the call.
    Returns (assign [n] = query index written into vpMatched[i] by this call or -1, nmatches)."""
    matched = matched_in.astype(bool).copy()
    assign = np.full(fv.n, -1, np.int32)
    nm = 0
    for qi, Q in enumerate(queries):
        if not (Q["flags"] & 1):
            continue
        best, bidx = 256, -1
        for i in fv.features_in_area(Q["u"], Q["v"], Q["radius"]):
            if matched[i]:
                continue
            lvl = int(fv.octave[i])
            if lvl < Q["pred_level"] - 1 or lvl > Q["pred_level"]:
                continue
            d = popcount_dist(qdesc[qi], fv.desc[i])
            if d < best:
                best, bidx = d, i
        if best <= th_low:
            matched[bidx] = True
            assign[bidx] = qi
            nm += 1
    return assign, nm


def search_for_initialization(kps1, desc1, prev_matched, fv2: FrameView, window, nn_ratio=0.9, th_low=50,
                              check_ori=True):
    """ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — src/ORBmatcher.cc:405-520.
    kps1: structured keypoints of F1 (mvKeysUn); prev_matched [n1][2] float32.  Returns (matches12, nmatches,
    prev_matched after the update)."""
    n1, n2 = len(kps1), fv2.n
    INT_MAX = 2 ** 31 - 1
    m12 = np.full(n1, -1, np.int32)
    m21 = np.full(n2, -1, np.int32)
    mdist = np.full(n2, INT_MAX, np.int64)
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for i1 in range(n1):
        level1 = int(kps1["octave"][i1])
        if level1 > 0:
            continue
        cand = fv2.features_in_area(prev_matched[i1, 0], prev_matched[i1, 1], window, level1, level1)
        if not cand:
            continue
        best, best2, bidx = INT_MAX, INT_MAX, -1
        for i2 in cand:
            d = popcount_dist(desc1[i1], fv2.desc[i2])
            if mdist[i2] <= d:
                continue
            if d < best:
                best2, best, bidx = best, d, i2
            elif d < best2:
                best2 = d
        if best <= th_low and F32(best) < F32(F32(best2) * F32(nn_ratio)):
            if m21[bidx] >= 0:
                m12[m21[bidx]] = -1
                nm -= 1
            m12[i1] = bidx
            m21[bidx] = i1
            mdist[bidx] = best
            nm += 1
            if check_ori:
                hist[rot_bin(kps1["angle"][i1], fv2.angle[bidx])].append(i1)
    if check_ori:
        keep = three_maxima([len(h) for h in hist])
        for b in range(HISTO):
            if b in keep:
                continue
            for i in hist[b]:
                if m12[i] >= 0:
                    m12[i] = -1
                    nm -= 1
    pm = prev_matched.astype(F32).copy()
    for i1 in range(n1):
        if m12[i1] >= 0:
            pm[i1, 0], pm[i1, 1] = fv2.x[m12[i1]], fv2.y[m12[i1]]
    return m12, nm, pm
