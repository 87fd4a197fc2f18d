"""TEST INFRASTRUCTURE — golden-vector generator for the line-junction detection of Frame::ExtractLSD ("next" row N2).

CPartiallyRecoverConnectivity's constructor (add_src/PartiallyRecoverConnectivity.cpp:14-133, called at
src/Frame.cc:504-505 with expandWidth = 20 and fanThr = pi/4, include/Frame.h:217-218) written from that file over the
REAL OpenCV primitives it calls: cv2.fastAtan2, cv2.addWeighted (what the cv::MatExpr of ptsDropInRotatedRect,
:154-160, evaluates to: `a*(X - s) + b*(Y - t)` folds into one addWeighted(X, a, Y, b, -(s*a) - (t*b)), whose CV_32F
kernel in cv2 4.13 works in double and rounds once) and cv2.determinant (2x2 CV_32F in double).  The scalar float
arithmetic follows the C++ types statement by statement; sin / cos / tan of a float are the correctly rounded values
(DESIGN.md H2).  The 3-D half (Frame::convertFansToKeyLines / Frame_shortestDistance, src/Frame.cc:380-472) solves its
2x2 system with Eigen's colPivHouseholderQr; Eigen is not available, so `cross3d_numpy` gives an independent
numpy.linalg solution that the tests compare with a tolerance.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
F64 = np.float64
CV_PI = 3.1415926535897932384626433832795


def _angles(p):
    import cv2
    dy, dx = F32(p[3] - p[1]), F32(p[2] - p[0])
    deg = F32(cv2.fastAtan2(float(dy), float(dx)))
    arc = F32(F64(F32(deg / F32(180))) * CV_PI)          # float degAng / 180 * CV_PI -> float
    return dy, dx, deg, arc


def _rect(p, radius):
    """RotatedRect of one line (:31-45): centre, half extents (CvSize truncates to int), sin / cos of its angle."""
    dy, dx, deg, arc = _angles(p)
    cx, cy = F32(F32(p[0] + p[2]) / F32(2)), F32(F32(p[1] + p[3]) / F32(2))
    length = abs(dy) if abs(F32(math.tan(float(arc)))) > 1 else abs(dx)
    height = int(F32(radius * F32(2)))                   # tsize.height = radius * 2   (int)
    width = int(F32(length + F32(F32(2) * radius)))      # tsize.width = length + 2 * radius   (int)
    ang = F32(F64(deg) * CV_PI / 180)                    # rotRect.angle * CV_PI / 180 -> float
    return cx, cy, F32(F32(width) / F32(2)), F32(F32(height) / F32(2)), F32(math.sin(float(ang))), F32(math.cos(float(ang))), arc


def _in_rect(x, y, R):
    cx, cy, hw, hh, dsin, dcos, _ = R
    fx = F32(F32(dcos * F32(x - cx)) + F32(dsin * F32(y - cy)))
    fy = F32(F32(dsin * F32(x - cx)) - F32(dcos * F32(y - cy)))
    return bool(-hw <= fx and fx < hw and -hh <= fy and fy < hh)


def _intersection(p, q):
    import cv2
    A1, B1 = F32(p[1] - p[3]), F32(p[2] - p[0])
    C1 = F32(F32(p[3] * p[0]) - F32(p[1] * p[2]))
    A2, B2 = F32(q[1] - q[3]), F32(q[2] - q[0])
    C2 = F32(F32(q[3] * q[0]) - F32(q[1] * q[2]))
    det = lambda a, b, c, d: cv2.determinant(np.array([[a, b], [c, d]], F32))
    with np.errstate(all="ignore"):
        D = F32(det(A1, B1, A2, B2))
        X = F32(F64(det(-C1, B1, -C2, B2)) / F64(D))
        Y = F32(F64(det(A1, -C1, A2, -C2)) / F64(D))
    return X, Y


def fans(lines, radius, fan_thr, img_w, img_h):
    """lines [n,4] float32 (mLines: start x, y, end x, y).  Returns fans [m,4] float32 (x, y, i, j) after the duplicate
    removal (:111-132) and the same list before it."""
    import cv2
    L = np.ascontiguousarray(lines, F32)
    n = len(L)
    radius, fan_thr = F32(radius), F32(fan_thr)
    X = np.concatenate([L[:, 0], L[:, 2]]).reshape(-1, 1).copy()      # mPts: the start points, then the end points
    Y = np.concatenate([L[:, 1], L[:, 3]]).reshape(-1, 1).copy()
    rects = [_rect(L[i], radius) for i in range(n)]
    raw = []
    for i in range(n):
        cx, cy, hw, hh, dsin, dcos, arc = rects[i]
        a, b = float(dcos), float(dsin)
        fx = cv2.addWeighted(X, a, Y, b, (-float(cx)) * a + (-float(cy)) * b)[:, 0]
        fy = cv2.addWeighted(X, b, Y, -a, (-float(cx)) * b - (-float(cy)) * a)[:, 0]
        for j in range(2 * n):
            if not (-hw <= fx[j] and fx[j] < hw and -hh <= fy[j] and fy[j] < hh):
                continue
            cur = j - n if j >= n else j
            if cur == i:
                continue
            arc1 = rects[cur][6]
            tmpa = F32(np.fmod(F32(abs(F32(arc - arc1))), F32(CV_PI)))
            if tmpa < fan_thr or CV_PI - float(tmpa) < float(fan_thr):
                continue
            ix, iy = _intersection(L[i], L[cur])
            if _in_rect(ix, iy, rects[i]) and ix >= 4 and ix < F32(img_w - 4) and iy >= 4 and iy < F32(img_h - 4):
                raw.append((ix, iy, F32(i), F32(cur)))
    raw = np.array(raw, F32).reshape(-1, 4)
    keep = []
    for i in range(len(raw)):
        s1, s2 = int(raw[i, 2]), int(raw[i, 3])
        dup = False
        for j in range(i + 1, len(raw)):
            s3, s4 = int(raw[j, 2]), int(raw[j, 3])
            if (s1 == s3 and s2 == s4) or (s1 == s4 and s2 == s3):
                dup = True
                break
        if not dup:
            keep.append(i)
    return raw[keep], raw


def cross3d_numpy(l1, l2):
    """Frame_shortestDistance (src/Frame.cc:380-424) with numpy.linalg.solve in place of Eigen's QR: (ok, point)."""
    l1, l2 = np.asarray(l1, F64), np.asarray(l2, F64)
    p1, p2, d1, d2 = l1[:3], l2[:3], l1[3:] - l1[:3], l2[3:] - l2[:3]
    w = p1 - p2
    A = np.array([[d1 @ d1, -(d1 @ d2)], [d1 @ d2, -(d2 @ d2)]])
    if A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0] == 0:
        return False, np.zeros(3)
    x = np.linalg.solve(A, np.array([-(w @ d1), -(w @ d2)]))
    r1, r2 = p1 + x[0] * d1, p2 + x[1] * d2
    mid_x, mid_y = (l1[:3] + l2[:3]) * 0.5, (l1[3:] + l2[3:]) * 0.5
    if np.linalg.norm(mid_x - mid_y) * 2 < np.linalg.norm(l1) + np.linalg.norm(l2):
        return True, (r1 + r2) * 0.5
    return False, np.zeros(3)    # the reference falls off the end of the function here (no return): pinned to "no point"
