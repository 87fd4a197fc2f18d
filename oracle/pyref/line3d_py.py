"""Independent restatement of Frame::isLineGood (src/Frame.cc:662-750) with LINEextractor::compPt3dCov,
extract3dline_mahdist, verify3dLine, computeLine3d_svd and mah_dist3d_pt_line (add_src/LineExtractor.cpp:27-323) over
the REAL cv2.SVDecomp (both decompositions), used to make tests/golden/lines3d_*.npz.

Pinned choice H6: `rand()` of random_unique (add_inc/LineExtractor.h:25-37) is the ANSI C example generator
(state = state * 1103515245 + 12345, return (state >> 16) & 0x7fff), re-seeded per line with
seed * 1000003 + line index + 1 — the reference draws from the process-wide rand() stream, which no test can pin.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

F32 = np.float32


class Rand:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFF

    def __call__(self):
        self.s = (self.s * 1103515245 + 12345) & 0xFFFFFFFF
        return (self.s >> 16) & 0x7FFF


def depth_std(d):
    return 0.00273 * d * d + 0.00074 * d + -0.00058


def mat3(a, b):  # cv::gemm of 3x3 doubles: sum over k in order
    return [[a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] for j in range(3)] for i in range(3)]


def comp_pt3d_cov(pt, f):
    x, y, z = pt
    J = [[z / f, 0.0, x / z], [0.0, z / f, y / z], [0.0, 0.0, 1.0]]
    s = depth_std(z)
    G = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, s * s]]
    Jt = [[J[j][i] for j in range(3)] for i in range(3)]
    cov = np.array(mat3(mat3(J, G), Jt), np.float64)
    w, u, _ = cv2.SVDecomp(cov)
    ws = [math.sqrt(float(w[i, 0])) for i in range(3)]
    D = [[1 / ws[0], 0.0, 0.0], [0.0, 1 / ws[1], 0.0], [0.0, 0.0, 1 / ws[2]]]
    Ut = [[float(u[j, i]) for j in range(3)] for i in range(3)]
    du = mat3(D, Ut)
    return [du[i][j] for i in range(3) for j in range(3)]


def mah_dist(pos, DU, q1, q2):
    xa, ya, za = q1
    xb, yb, zb = q2
    c1, c2, c3, c4, c5, c6, c7, c8, c9 = DU
    x1, x2, x3 = pos
    A1 = c1 * (x1 - xa) + c2 * (x2 - ya) + c3 * (x3 - za)
    A2 = c4 * (x1 - xa) + c5 * (x2 - ya) + c6 * (x3 - za)
    A3 = c7 * (x1 - xa) + c8 * (x2 - ya) + c9 * (x3 - za)
    B1 = c1 * (x1 - xb) + c2 * (x2 - yb) + c3 * (x3 - zb)
    B2 = c4 * (x1 - xb) + c5 * (x2 - yb) + c6 * (x3 - zb)
    B3 = c7 * (x1 - xb) + c8 * (x2 - yb) + c9 * (x3 - zb)
    t1 = A1 * B2 - A2 * B1
    t2 = A1 * B3 - A3 * B1
    t3 = A2 * B3 - A3 * B2
    t4 = c1 * (x1 - xa) - c1 * (x1 - xb) + c2 * (x2 - ya) - c2 * (x2 - yb) + c3 * (x3 - za) - c3 * (x3 - zb)
    t5 = c4 * (x1 - xa) - c4 * (x1 - xb) + c5 * (x2 - ya) - c5 * (x2 - yb) + c6 * (x3 - za) - c6 * (x3 - zb)
    t6 = c7 * (x1 - xa) - c7 * (x1 - xb) + c8 * (x2 - ya) - c8 * (x2 - yb) + c9 * (x3 - za) - c9 * (x3 - zb)
    with np.errstate(all="ignore"):
        return float(np.sqrt(np.float64(t1 * t1 + t2 * t2 + t3 * t3) / np.float64(t4 * t4 + t5 * t5 + t6 * t6)))


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def mul(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


def norm(a):  # cv::norm(Point3d) = sqrt(x*x + y*y + z*z)
    return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def proj_pt(P, mid, drct):  # projPt3d2Ln3d
    A = mid
    B = add(mid, drct)
    AB = sub(B, A)
    AP = sub(P, A)
    return add(A, mul(AB, dot(AB, AP) / dot(AB, AB)))


def verify3d(pos_list, A, B):
    n_cells, ratio = 10, 0.7
    cells = [0] * n_cells
    minv, maxv, i1, i2 = 100.0, -100.0, 0, 0
    BA = sub(B, A)
    for i, p in enumerate(pos_list):
        v = dot(sub(p, A), BA)
        if v < minv:
            minv, i1 = v, i
        if v > maxv:
            maxv, i2 = v, i
    mid = mul(add(A, B), 0.5)
    C = proj_pt(pos_list[i1], mid, BA)
    D = proj_pt(pos_list[i2], mid, BA)
    cd = norm(sub(D, C))
    if cd < 0.0000000001:
        return False
    DC = sub(D, C)
    for p in pos_list:
        lam = abs(dot(sub(p, C), DC) / cd / cd)
        if lam >= 1:
            cells[n_cells - 1] += 1
        else:
            cells[int(math.floor(lam * 10))] += 1
    s = sum(1 for c in cells if c > 0)
    return s / n_cells > ratio


def line3d_svd(pos, idx):
    n = len(idx)
    mean = (0.0, 0.0, 0.0)
    for i in idx:
        mean = add(mean, pos[i])
    mean = mul(mean, 1.0 / n)
    P = np.array([[pos[i][0] - mean[0], pos[i][1] - mean[1], pos[i][2] - mean[2]] for i in idx], np.float64)  # = P.t(): n x 3
    _, _, vt = cv2.SVDecomp(P, flags=cv2.SVD_MODIFY_A)
    return mean, (float(vt[0, 0]), float(vt[0, 1]), float(vt[0, 2]))


def extract3dline(pos, DUs, rnd):
    n = len(pos)
    max_iter = min(10, int(n * (n - 1) * 0.5))
    thresh = 3.0
    idxs = list(range(n))
    best_set, bestA, bestB = [], None, None
    for _ in range(max_iter):
        left = n
        for b in range(2):  # random_unique(begin, end, 2)
            r = b + rnd() % left
            idxs[b], idxs[r] = idxs[r], idxs[b]
            left -= 1
        A, B = pos[idxs[0]], pos[idxs[1]]
        if norm(sub(B, A)) < 0.0000000001:
            continue
        inl = [i for i in range(n) if mah_dist(pos[i], DUs[i], A, B) < thresh]
        if len(inl) > len(best_set):
            if verify3d([pos[i] for i in inl], A, B):
                best_set, bestA, bestB = inl, A, B
        if len(best_set) > n * 0.6:
            break
    if len(best_set) >= 2:
        m, d = mul(add(bestA, bestB), 0.5), sub(bestB, bestA)
        while True:
            tm, td = line3d_svd(pos, best_set)
            tmp = [i for i in range(n) if mah_dist(pos[i], DUs[i], tm, add(tm, td)) < thresh]
            if len(tmp) > len(best_set):
                best_set, m, d = tmp, tm, td
            else:
                break
        minv, maxv, e1, e2 = 100.0, -100.0, 0, 0
        for k, i in enumerate(best_set):
            v = dot(sub(pos[i], m), d)
            if v < minv:
                minv, e1 = v, k
            if v > maxv:
                maxv, e2 = v, k
        return pos[best_set[e1]], pos[best_set[e2]]
    return (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)   # RandomLine3d's default points


def is_line_good(kl, depth, fx, fy, cx, cy, seed):
    """kl: KEYLINE_DTYPE; depth: float32 [h, w] metres.  Returns (lines3d [n,6] f64, line_eq [n,3] f32)."""
    h, w = depth.shape
    n = len(kl)
    fx, fy, cx, cy = F32(fx), F32(fy), F32(cx), F32(cy)
    invfx, invfy = F32(F32(1.0) / fx), F32(F32(1.0) / fy)
    f = float(fx)
    lines3d = np.zeros((n, 6), np.float64)
    eq = np.full((n, 3), -1.0, np.float32)
    for i in range(n):
        sx, sy, ex, ey = (F32(kl[i][k]) for k in ("start_x", "start_y", "end_x", "end_y"))
        dx, dy = F32(sx - ex), F32(sy - ey)
        ln = math.sqrt(float(dx) * float(dx) + float(dy) * float(dy))
        num = float(min(int(ln), 20))
        if num < 1:
            continue  # 0 / 0 sample positions in the reference: undefined, no line
        pos = []
        for j in range(int(num) + 1):
            a, b = 1 - j / num, j / num
            px = float(F32(F32(float(sx) * a) + F32(float(ex) * b)))
            py = float(F32(F32(float(sy) * a) + F32(float(ey) * b)))
            if px < 0 or py < 0 or px >= w or py >= h:
                continue
            if math.floor(px) == px and math.floor(py) == py:
                col, row = max(int(px - 1), 0), max(int(py - 1), 0)
            else:
                col, row = int(px), int(py)
            d = depth[row, col]
            if float(d) <= 0.01:
                continue
            z = float(d)
            x = float(F32(F32(col) - cx)) * z * float(invfx)
            y = float(F32(F32(row) - cy)) * z * float(invfy)
            pos.append((x, y, z))
        if len(pos) < 5:
            continue
        DUs = [comp_pt3d_cov(p, f) for p in pos]
        A, B = extract3dline(pos, DUs, Rand(seed * 1000003 + i + 1))
        if norm(sub(A, B)) > 0.02:
            lines3d[i] = [*A, *B]
            le = np.array([F32(B[0] - A[0]), F32(B[1] - A[1]), F32(B[2] - A[2])], F32)
            magn = F32(np.sqrt(F32(F32(F32(le[0] * le[0]) + F32(le[1] * le[1])) + F32(le[2] * le[2]))))
            eq[i] = [F32(le[0] / magn), F32(le[1] / magn), F32(le[2] / magn)]
    return lines3d, eq
