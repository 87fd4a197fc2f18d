"""TEST INFRASTRUCTURE — golden-vector generator for the line matchers.

Independent Python restatement (written from the reference source, not from oracle/c) of
LSDmatcher::matchNNR / match (add_src/LSDmatcher.cpp:354-413), SearchByGeomNApearance (:36-110),
SearchByProjection (:112-215 and :260-352), FrameBFMatch + lineDescriptorMAD (:492-516, :660-685),
SearchDouble (:462-490), Frame::AssignFeaturesToGridForLine / GetFeaturesInAreaForLine
(src/Frame.cc:286-309, 752-826), ORB_SLAM2::LineIterator (add_src/lineIterator.cpp:33-77),
InsectLineMatch::SearchMapInsectline (add_src/InsectlineMatch.cpp:9-60) and Map::AssociatePlanesByBoundary
(src/Map.cc:204-272).  cv2.BFMatcher supplies the real knnMatch.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
F64 = np.float64
COLS, ROWS = 64, 48


def knn2_cv2(q, t):
    import cv2
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(np.ascontiguousarray(q), np.ascontiguousarray(t), 2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), -1.0, np.float32)
    for i, r in enumerate(m):
        for k, d in enumerate(r):
            idx[i, k] = d.trainIdx
            dist[i, k] = d.distance
    return idx, dist


def match_nnr(d1, d2, nnr):
    out = np.full(len(d1), -1, np.int32)
    if len(d1) == 0 or len(d2) < 2:
        return out, 0
    idx, dist = knn2_cv2(d1, d2)
    ok = dist[:, 0] < dist[:, 1] * F32(nnr)
    out[ok] = idx[ok, 0]
    return out, int(ok.sum())


def angle2d(ax, ay, bx, by):
    ax, ay, bx, by = float(ax), float(ay), float(bx), float(by)
    dot = ax * bx + ay * by
    return abs(dot / (math.sqrt(ax * ax + ay * ay) * math.sqrt(bx * bx + by * by)))


def search_geom(kl_last, d_last, has_ml, kl_cur, d_cur, bounds, desc_th):
    assign = np.full(len(kl_cur), -1, np.int32)
    if len(kl_cur) == 0:
        return assign, 0
    m12, _ = match_nnr(d_last, d_cur, desc_th)
    dW = float(F32(bounds[2]) - F32(bounds[0])) * 0.1
    dH = float(F32(bounds[3]) - F32(bounds[1])) * 0.1
    cth = math.cos(20.0 / 180.0 * math.pi)
    n = 0
    for i1 in range(len(kl_last)):
        if not has_ml[i1] or m12[i1] < 0:
            continue
        c, l = kl_cur[m12[i1]], kl_last[i1]
        if c["start_x"] == 0:
            continue
        a = angle2d(F32(c["e_oct_x"] - c["s_oct_x"]), F32(c["e_oct_y"] - c["s_oct_y"]),
                    F32(l["e_oct_x"] - l["s_oct_x"]), F32(l["e_oct_y"] - l["s_oct_y"]))
        if a < cth:
            continue
        far_s = abs(float(F32(c["s_oct_x"] - l["s_oct_x"]))) > dW or abs(float(F32(c["s_oct_y"] - l["s_oct_y"]))) > dH
        far_e = abs(float(F32(c["e_oct_x"] - l["e_oct_x"]))) > dW or abs(float(F32(c["e_oct_y"] - l["e_oct_y"]))) > dH
        if far_s and far_e:
            continue
        assign[m12[i1]] = i1
        n += 1
    return assign, n


def frame_bf_match(d1, d2, nn_ratio, th):
    out = np.full(len(d1), -1, np.int32)
    if len(d1) == 0 or len(d2) < 2:
        return out
    idx, dist = knn2_cv2(d1, d2)
    d0, dd1 = dist[:, 0], dist[:, 1]
    n = len(d0)
    med12 = float(np.sort(dd1 - d0)[n // 2])
    dev = np.abs((dd1 - d0 - F32(0)).astype(np.float64) - med12).astype(np.float32)  # fabsf(float - float - double)
    nn12 = 1.4826 * float(np.sort(dev)[n // 2]) * 0.5
    for i in range(n):
        if float(dd1[i] - d0[i]) > nn12 and d0[i] < F32(th) and d0[i] < F32(nn_ratio) * dd1[i]:
            out[i] = idx[i, 0]
    return out


def search_double(d1, d2, nn_ratio, th):
    out = np.full(len(d1), -1, np.int32)
    if len(d1) == 0 or len(d2) == 0:
        return out, 0
    m12 = frame_bf_match(d1, d2, nn_ratio, th)
    m21 = frame_bf_match(d2, d1, nn_ratio, th)
    n = 0
    for i, j in enumerate(m12):
        if j >= 0 and m21[j] == i:
            out[i] = j
            n += 1
    return out, n


def bresenham_cells(x1, y1, x2, y2):
    """ORB_SLAM2::LineIterator over doubles."""
    steep = abs(y2 - y1) > abs(x2 - x1)
    if steep:
        x1, y1, x2, y2 = y1, x1, y2, x2
    if x1 > x2:
        x1, x2, y1, y2 = x2, x1, y2, y1
    dx, dy = x2 - x1, abs(y2 - y1)
    err = dx / 2.0
    ystep = 1 if y1 < y2 else -1
    x, y, mx = int(x1), int(y1), int(x2)
    out = []
    while x <= mx:
        out.append((y, x) if steep else (x, y))
        err -= dy
        if err < 0:
            y += ystep
            err += dx
        x += 1
    return out


class LineFrame:
    def __init__(self, kl, desc, lineeq, lines3d, bounds):
        self.kl, self.desc, self.eq, self.l3d = kl, desc, lineeq, lines3d
        self.min_x, self.min_y, self.max_x, self.max_y = (F32(b) for b in bounds)
        self.w_inv = F32(COLS) / F32(self.max_x - self.min_x)
        self.h_inv = F32(ROWS) / F32(self.max_y - self.min_y)
        self.grid = {}
        for i, k in enumerate(kl):
            for (px, py) in bresenham_cells(float(F32(k["start_x"]) * self.w_inv), float(F32(k["start_y"]) * self.h_inv),
                                            float(F32(k["end_x"]) * self.w_inv), float(F32(k["end_y"]) * self.h_inv)):
                if 0 <= px < COLS and 0 <= py < ROWS:
                    self.grid.setdefault((px, py), []).append(i)

    def in_area(self, x1, y1, x2, y2, r, TH):
        x1, y1, x2, y2, r, TH = F32(x1), F32(y1), F32(x2), F32(y2), F32(r), F32(TH)
        xs = [x1, F32(float(F32(x1 + x2)) / 2.0), x2]
        ys = [y1, F32(float(F32(y1 + y2)) / 2.0), y2]
        d1x, d1y = F32(x1 - x2), F32(y1 - y2)
        n1 = F32(math.sqrt(float(F32(F32(d1x * d1x) + F32(d1y * d1y)))))
        d1x, d1y = F32(d1x / n1), F32(d1y / n1)
        out, seen = [], set()
        for x, y in zip(xs, ys):
            cx0 = max(0, math.floor(float(F32(F32(F32(x - self.min_x) - r) * self.w_inv))))
            if cx0 >= COLS:
                continue
            cx1 = min(COLS - 1, math.ceil(float(F32(F32(F32(x - self.min_x) + r) * self.w_inv))))
            if cx1 < 0:
                continue
            cy0 = max(0, math.floor(float(F32(F32(F32(y - self.min_y) - r) * self.h_inv))))
            if cy0 >= ROWS:
                continue
            cy1 = min(ROWS - 1, math.ceil(float(F32(F32(F32(y - self.min_y) + r) * self.h_inv))))
            if cy1 < 0:
                continue
            for ix in range(cx0, cx1 + 1):
                for iy in range(cy0, cy1 + 1):
                    for j in self.grid.get((ix, iy), []):
                        if j in seen:
                            continue
                        k = self.kl[j]
                        d2x, d2y = F32(k["start_x"] - k["end_x"]), F32(k["start_y"] - k["end_y"])
                        n2 = F32(math.sqrt(float(F32(F32(d2x * d2x) + F32(d2y * d2y)))))
                        d2x, d2y = F32(d2x / n2), F32(d2y / n2)
                        cs = abs(F32(F32(d1x * d2x) + F32(d1y * d2y)))
                        if cs < TH:
                            continue
                        L = self.eq[j]
                        dist = F32(float(L[0]) * float(x) + float(L[1]) * float(y) + float(L[2]))
                        if abs(float(dist)) < float(r):
                            out.append(j)
                            seen.add(j)
        return out


def popcount_dist(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def search_by_projection(fr: LineFrame, q, qdesc, claimed_in, mode, nn_ratio):
    n = len(fr.kl)
    assign = np.full(n, -1, np.int32)
    claimed = np.zeros(n, bool) if claimed_in is None else claimed_in.astype(bool).copy()
    c10, c15 = math.cos(10.0 / 180.0 * math.pi), math.cos(15.0 / 180.0 * math.pi)
    nm = 0
    for qi in range(len(q)):
        Q = q[qi]
        if not (Q["flags"] & 1):
            continue
        cand = fr.in_area(Q["x1"], Q["y1"], Q["x2"], Q["y2"], Q["radius"], 0.96 if mode == 0 else 0.998)
        best, best2, lvl, lvl2, bidx = 256, 256, -1, -1, -1
        for i2 in cand:
            if claimed[i2]:
                continue
            c = fr.kl[i2]
            if mode == 0:
                a = angle2d(F32(c["e_oct_x"] - c["s_oct_x"]), F32(c["e_oct_y"] - c["s_oct_y"]),
                            F32(Q["ex"] - Q["sx"]), F32(Q["ey"] - Q["sy"]))
                if a < c10:
                    continue
                d = popcount_dist(qdesc[qi], fr.desc[i2])
                mx, mn = max(F32(Q["length"]), F32(c["line_length"])), min(F32(Q["length"]), F32(c["line_length"]))
                if float(F32(mn / mx)) < 0.75:
                    continue
                if d < best:
                    best, bidx = d, i2
            else:
                v = fr.l3d[i2, :3] - fr.l3d[i2, 3:]
                nrm = Q["normal"]
                dot = F32(v[0] * nrm[0] + v[1] * nrm[1] + v[2] * nrm[2])
                mf = F32(math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))
                mm = F32(math.sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]))
                ang = abs(F32(dot / F32(mf * mm)))
                if float(ang) < c15:
                    continue
                d = popcount_dist(qdesc[qi], fr.desc[i2])
                if d < best:
                    best2, best, lvl2, lvl, bidx = best, d, lvl, int(c["octave"]), i2
                elif d < best2:
                    lvl2, best2 = int(c["octave"]), d
        if best <= 95:
            if mode == 1 and lvl == lvl2 and F32(best) > F32(nn_ratio) * F32(best2):
                continue
            assign[bidx] = qi
            if Q["flags"] & 2:
                claimed[bidx] = True
            nm += 1
    return assign, nm


def plane_assoc(planes_cam, pts, Tcw, map_planes, map_bad, d_th, a_th, mode):
    n = len(planes_cam)
    assign = np.full(n, -1, np.int32)
    nm = 0
    dTh = F32(d_th)
    for i in range(n):
        pM = (Tcw.astype(np.float64).T @ planes_cam[i].astype(np.float64)).astype(np.float32)
        ldTh = F32(d_th)
        found = False
        for m in range(len(map_planes)):
            if mode == 0 and map_bad is not None and map_bad[m]:
                continue
            pW = map_planes[m].astype(np.float32)
            if mode == 1 and pW[3] < 0:
                pW = -pW
            ang = F32(F32(F32(pM[0] * pW[0]) + F32(pM[1] * pW[1])) + F32(pM[2] * pW[2]))
            if ang > F32(a_th) or ang < -F32(a_th):
                d = [F32(float(pW[0]) * pts[i, 3 * k] + float(pW[1]) * pts[i, 3 * k + 1] + float(pW[2]) * pts[i, 3 * k + 2]
                         + float(pW[3])) for k in range(5)]
                dis = F32(F32(F32(F32(F32(d[0] + d[1]) + d[2]) + d[3]) + d[4]) / F32(5))
                thr = ldTh if mode == 0 else dTh
                if abs(dis) < thr:
                    if mode == 0:
                        ldTh = dis
                    else:
                        dTh = dis
                        nm += 1
                    assign[i] = m
                    found = True
        if mode == 0 and found:
            nm += 1
    return assign, nm


# ---------------------------------------------------------------------------------------------------
# LSDmatcher::Fuse window search (LSDmatcher.cpp:916-953) over KeyFrame::GetLinesInArea (KeyFrame.cc:857-891)
# ---------------------------------------------------------------------------------------------------
def lines_in_area(kl, x1, y1, x2, y2, r, TH=0.998):
    """All KeyLines of the KeyFrame in index order; float / double mix as written in the reference."""
    x1, y1, x2, y2, r, TH = F32(x1), F32(y1), F32(x2), F32(y2), F32(r), F32(TH)
    out = []
    with np.errstate(all="ignore"):
        d1x, d1y = F32(x1 - x2), F32(y1 - y2)
        n1 = F32(np.sqrt(F32(F32(d1x * d1x) + F32(d1y * d1y))))
        d1x, d1y = F32(d1x / n1), F32(d1y / n1)
        mx, my = 0.5 * float(F32(x1 + x2)), 0.5 * float(F32(y1 + y2))   # doubles
        for i in range(len(kl)):
            k = kl[i]
            ex, ey = mx - float(k["pt_x"]), my - float(k["pt_y"])
            distance = F32(ex * ex + ey * ey)
            if distance > F32(r * r):
                continue
            d2x, d2y = F32(k["start_x"] - k["end_x"]), F32(k["start_y"] - k["end_y"])
            n2 = F32(np.sqrt(F32(F32(d2x * d2x) + F32(d2y * d2y))))
            d2x, d2y = F32(d2x / n2), F32(d2y / n2)
            cs = F32(abs(F32(F32(d1x * d2x) + F32(d1y * d2y))))
            if cs < TH:      # NaN (degenerate direction) passes, as in the reference
                continue
            out.append(i)
    return out


def line_fuse(kl, kf_desc, queries, qdesc, th_cos=0.998, th_low=50):
    nq = len(queries)
    bi, bd = np.full(nq, -1, np.int32), np.full(nq, 256, np.int32)
    for q in range(nq):
        Q = queries[q]
        if not (int(Q["flags"]) & 1):
            continue
        best_d, best_i = 256, -1
        for idx in lines_in_area(kl, Q["u1"], Q["v1"], Q["u2"], Q["v2"], Q["radius"], th_cos):
            lvl = int(kl[idx]["octave"])
            if lvl < int(Q["pred_level"]) - 1 or lvl > int(Q["pred_level"]):
                continue
            d = int(np.unpackbits(np.bitwise_xor(qdesc[q], kf_desc[idx])).sum())
            if d < best_d:
                best_d, best_i = d, idx
        bd[q] = best_d
        if best_d <= th_low and best_i >= 0:
            bi[q] = best_i
    return bi, bd


# ---------------------------------------------------------------------------------------------------
# Plane hypotheses of Frame::ExtractLSD (Frame.cc:512-645) with Frame::OldPlane (:474-487)
# ---------------------------------------------------------------------------------------------------
def plane_hypotheses(kl_un, line_eq, lines3d, junctions):
    """junctions: structured (l1, l2, cross2d_x, cross2d_y, cross3d[3]).  Returns (le_l [nj,6] f64, planes [np,4] f32,
    normals [np,3] f64, junction_of [np] i32).  numpy scalar arithmetic in the reference's types."""
    F64 = np.float64
    le = np.zeros((len(junctions), 6), F64)
    planes, normals, owner = [], [], []
    with np.errstate(all="ignore"):
        for i, J in enumerate(junctions):
            l1, l2 = int(J["l1"]), int(J["l2"])
            for s_, l in enumerate((l1, l2)):
                k = kl_un[l]
                sp = np.array([F64(k["start_x"]), F64(k["start_y"]), F64(1.0)])
                ep = np.array([F64(k["end_x"]), F64(k["end_y"]), F64(1.0)])
                c = np.array([sp[1] * ep[2] - sp[2] * ep[1], sp[2] * ep[0] - sp[0] * ep[2], sp[0] * ep[1] - sp[1] * ep[0]])
                le[i, 3 * s_:3 * s_ + 3] = c / np.sqrt(c[0] * c[0] + c[1] * c[1])
            e1, e2 = line_eq[l1].astype(F32), line_eq[l2].astype(F32)
            if (e1 == 0).all() or (e2 == 0).all():
                continue
            A, B = lines3d[l1].astype(F64), lines3d[l2].astype(F64)
            if (np.abs(A) <= 1e-12).all() or (np.abs(B) <= 1e-12).all():
                continue
            pn = np.array([F32(F32(e1[1] * e2[2]) - F32(e1[2] * e2[1])), F32(F32(e1[2] * e2[0]) - F32(e1[0] * e2[2])),
                           F32(F32(e1[0] * e2[1]) - F32(e1[1] * e2[0]))], F32)
            norm = F32(np.sqrt(F32(F32(F32(pn[0] * pn[0]) + F32(pn[1] * pn[1])) + F32(pn[2] * pn[2]))))
            pn = np.array([F32(pn[0] / norm), F32(pn[1] / norm), F32(pn[2] / norm)], F32)
            n_ = pn.astype(F64)
            pts = [A[:3], A[3:], B[:3], B[3:], np.asarray(J["cross3d"], F64)]
            d = [F32(n_[0] * P[0] + n_[1] * P[1] + n_[2] * P[2]) for P in pts]
            dmin, dmax = F32(10000), F32(-10000)
            for v in d:
                dmin = dmin if dmin < v else v
                dmax = dmax if dmax > v else v
            if F64(F32(dmax - dmin)) > 0.05:
                continue
            dis = F32(-F32(F32(F32(F32(d[0] + d[1]) + d[2]) + d[3]) + d[4]) / F32(5))
            pl = np.array([pn[0], pn[1], pn[2], dis], F32)
            if pl[3] < 0:
                pl, n_ = -pl, -n_
            old = False
            for q in planes:
                dd = F32(pl[3] - q[3])
                ang = F32(F32(F32(pl[0] * q[0]) + F32(pl[1] * q[1])) + F32(pl[2] * q[2]))
                if F64(dd) > 0.2 or F64(dd) < -0.2:
                    continue
                if F64(ang) < 0.9397 and F64(ang) > -0.9397:
                    continue
                old = True
                break
            if old:
                continue
            planes.append(pl)
            normals.append(n_)
            owner.append(i)
    return (le, np.array(planes, F32).reshape(-1, 4), np.array(normals, F64).reshape(-1, 3), np.array(owner, np.int32))


# ---------------------------------------------------------------------------------------------------
# LSDmatcher::FrameBFMatchNew / mutualOverlap / SearchForTriangulationNew (LSDmatcher.cpp:518-658, 783-824): the
# epipolar-overlap variant of the line triangulation search (nothing in the reference calls it; "next" row N1)
# ---------------------------------------------------------------------------------------------------
def _matvec3_f32(F, x, y):
    """cv::Mat(3x3, CV_32F) * (x, y, 1): double accumulation, one rounding (cv::Mat float products)."""
    return [F32(F64(F[k, 0]) * F64(x) + F64(F[k, 1]) * F64(y) + F64(F[k, 2]) * F64(1.0)) for k in range(3)]


def _cross_f32(a, b):
    return [F32(F32(a[1] * b[2]) - F32(a[2] * b[1])), F32(F32(a[2] * b[0]) - F32(a[0] * b[2])),
            F32(F32(a[0] * b[1]) - F32(a[1] * b[0]))]


def _norm_diff(a, b):
    """cv::norm(a - b) for 3x1 CV_32F (a double): float differences, squares accumulated in double."""
    d = [F32(a[k] - b[k]) for k in range(3)]
    return F64(math.sqrt(F64(d[0]) * F64(d[0]) + F64(d[1]) * F64(d[1]) + F64(d[2]) * F64(d[2])))


def mutual_overlap(pts):
    """LSDmatcher::mutualOverlap (:583-658) on four collinear homogeneous points."""
    max_dist, o1, o2 = F32(0), 0, 3
    for i in range(3):
        for j in range(i + 1, 4):
            d = F32(_norm_diff(pts[i], pts[j]))      # float dist = norm(...)
            if d > max_dist:
                max_dist, o1, o2 = d, i, j
    if max_dist < F32(1.0):
        return F32(0)
    inner = [k for k in range(4) if k != o1 and k != o2]
    return F32(_norm_diff(pts[inner[0]], pts[inner[1]]) / F64(max_dist))     # double norm / float -> float


def frame_bf_match_new(d1, d2, kl1, kl2, func2, F, th, nn_ratio):
    """LSDmatcher::FrameBFMatchNew (:518-581): the best knn match of a line of set 1 is kept if the projections of its end
    points onto the matched line (along their epipolar lines F * p) overlap that line's segment by more than 0.8, its
    distance is < th and passes the ratio test.  F: 3x3 float32."""
    out = np.full(len(d1), -1, np.int32)
    if len(d2) < 2:
        return out          # knnMatch returns one entry: the loop `j < size() - 1` does not run
    idx, dist = knn2_cv2(d1, d2)
    F = np.asarray(F, F32)
    with np.errstate(all="ignore"):
        for q in range(len(d1)):
            t = int(idx[q, 0])
            e1 = _matvec3_f32(F, kl1["start_x"][q], kl1["start_y"][q])
            e2 = _matvec3_f32(F, kl1["end_x"][q], kl1["end_y"][q])
            l2 = [F32(func2[t, 0]), F32(func2[t, 1]), F32(func2[t, 2])]
            p1, p2 = _cross_f32(l2, e1), _cross_f32(l2, e2)
            if not (abs(float(p1[2])) > 1e-12 and abs(float(p2[2])) > 1e-12):
                continue
            s1, s2 = F32(1.0 / F64(p1[2])), F32(1.0 / F64(p2[2]))      # Mat /= s  ->  convertTo(alpha = 1 / s) in float
            p1 = [F32(v * s1) for v in p1]
            p2 = [F32(v * s2) for v in p2]
            q1 = [F32(kl2["start_x"][t]), F32(kl2["start_y"][t]), F32(1)]
            q2 = [F32(kl2["end_x"][t]), F32(kl2["end_y"][t]), F32(1)]
            score = mutual_overlap([p1, p2, q1, q2])
            d0, dd1 = F32(dist[q, 0]), F32(dist[q, 1])
            if d0 < F32(th) and float(score) > 0.8 and d0 < F32(F32(nn_ratio) * dd1):
                out[q] = t
    return out


def search_for_triangulation_new(d1, kl1, func1, ml1, d2, kl2, func2, ml2, F21, F12, nn_ratio, th, is_double):
    """LSDmatcher::SearchForTriangulationNew (:783-824).  Returns (vMatchedPairs [n1], nmatches)."""
    out = np.full(len(d1), -1, np.int32)
    if len(d1) == 0 or len(d2) == 0:
        return out, 0
    t1 = frame_bf_match_new(d1, d2, kl1, kl2, func2, F21, th, nn_ratio)
    t2 = frame_bf_match_new(d2, d1, kl2, kl1, func1, F12, th, nn_ratio)
    n = 0
    for i in range(len(d1)):
        j = t1[i]
        if j < 0:
            continue
        if is_double and t2[j] != i:
            continue
        if ml1[i] or ml2[j]:
            continue
        out[i] = j
        n += 1
    return out, n
