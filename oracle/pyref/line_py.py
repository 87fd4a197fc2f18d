"""TEST INFRASTRUCTURE — golden-vector generator for the line front end.

The real cv2 4.13 supplies what OpenCV supplies to the reference: cv2.createLineSegmentDetector().detect
(LSD), cv2.GaussianBlur(5x5, sigma 1) + cv2.Sobel (LBD gradients) and cv2.clipLine (LineIterator count).
The first-party / contrib parts are restated independently in Python float32, written from the reference
source and not from oracle/c: the long-line merge (add_src/uselongline.cpp:17-351, 411-485), the top-N
filter and line equations (add_src/LineExtractor.cpp:342-363) and the LBD descriptor
(Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:219-261, 402-413, 1027-1373).
"""
from __future__ import annotations

import math

import cv2
import numpy as np

F = np.float32
PI = math.pi


def lsd_cv2(img):
    l = cv2.createLineSegmentDetector().detect(img)[0]
    return np.zeros((0, 4), F) if l is None else l.reshape(-1, 4).astype(F)


def clamp_segments(lines, w, h):
    """checkLineExtremes, LSDDetector_custom.cpp:112-138"""
    out = lines.copy()
    for e in out:
        for k, lim in ((0, w), (2, w), (1, h), (3, h)):
            if e[k] < 0:
                e[k] = 0
            if e[k] >= lim:
                e[k] = F(lim) - F(1.0)
    return out


def _pld(l, x0, y0):
    x1, y1, x2, y2 = (F(v) for v in l)
    num = abs(F(F(F(F(y2 - y1) * x0) + F(F(x1 - x2) * y0)) + F(F(x2 * y1) - F(x1 * y2))))
    den = math.sqrt(float(F(y2 - y1)) ** 2 + float(F(x1 - x2)) ** 2)
    return F(float(num) / den)


def _angle_diff(a1, a2):
    c1 = abs(F(a2 - a1))
    c2 = F(PI + float(min(a1, a2)) - float(max(a1, a2)))
    return min(c1, c2)


def merge_two(l1, l2):
    ax, ay, bx, by = (F(v) for v in l1)
    cx, cy, dx, dy = (F(v) for v in l2)
    dlix, dliy, dljx, dljy = F(bx - ax), F(by - ay), F(dx - cx), F(dy - cy)
    li = math.sqrt(float(F(dlix * dlix)) + float(F(dliy * dliy)))
    lj = math.sqrt(float(F(dljx * dljx)) + float(F(dljy * dljy)))
    xg = (li * float(F(ax + bx)) + lj * float(F(cx + dx))) / (2.0 * (li + lj))
    yg = (li * float(F(ay + by)) + lj * float(F(cy + dy))) / (2.0 * (li + lj))
    thi = PI / 2.0 if dlix == 0 else math.atan(float(F(dliy / dlix)))
    thj = PI / 2.0 if dljx == 0 else math.atan(float(F(dljy / dljx)))
    if abs(thi - thj) <= PI / 2.0:
        thr = (li * thi + lj * thj) / (li + lj)
    else:
        tmp = thj - PI * (thj / abs(thj))
        thr = (li * thi + lj * tmp) / (li + lj)
    s, c = math.sin(thr), math.cos(thr)
    g = [(float(py) - yg) * s + (float(px) - xg) * c for px, py in ((ax, ay), (bx, by), (cx, cy), (dx, dy))]
    d1, d2 = min(g), max(g)
    return np.array([d1 * c + xg, d1 * s + yg, d2 * c + xg, d2 * s + yg], F)


def merge_lines(src, angle_thr, dist_thr, ep_threshold):
    n = len(src)
    if n == 0:
        return np.zeros((0, 4), F)
    angle_thr, dist_thr = F(angle_thr), F(dist_thr)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (src[:, 3] - src[:, 1]) / (src[:, 2] - src[:, 0])
    angles = np.array([F(math.atan(float(v))) for v in q], F)
    d = src[:, 2:] - src[:, :2]
    length = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(F)).astype(F)
    order = sorted(range(n), key=lambda i: angles[i])  # stable
    ep_thr = F(F(ep_threshold) * F(ep_threshold))
    qpi = F(PI / 4.0)
    nb = [[] for _ in range(n)]
    for i in range(n):
        i1 = order[i]
        x11, y11, x12, y12 = src[i1]
        a1 = angles[i1]
        sx = abs(a1) < qpi
        if (sx and x12 < x11) or ((not sx) and y12 < y11):
            x11, x12, y11, y12 = x12, x11, y12, y11
        for j in range(i + 1, n):
            i2 = order[j]
            x21, y21, x22, y22 = src[i2]
            if (sx and x22 < x21) or ((not sx) and y22 < y21):
                x21, x22, y21, y22 = x22, x21, y22, y21
            if _angle_diff(a1, angles[i2]) > angle_thr:
                if float(abs(a1)) < (PI / 2 - float(angle_thr)):
                    break
                continue
            m1 = (F(0.5 * float(F(src[i1][0] + src[i1][2]))), F(0.5 * float(F(src[i1][1] + src[i1][3]))))
            m2 = (F(0.5 * float(F(src[i2][0] + src[i2][2]))), F(0.5 * float(F(src[i2][1] + src[i2][3]))))
            if _pld(src[i2], *m1) > dist_thr and _pld(src[i1], *m2) > dist_thr:
                continue
            if (sx and x12 > x22) or ((not sx) and y12 > y22):
                c12, c21 = (x22, y22), (x11, y11)
            else:
                c12, c21 = (x12, y12), (x21, y21)
            ok = (sx and c12[0] >= c21[0]) or ((not sx) and c12[1] >= c21[1])
            if not ok:
                ex, ey = F(c21[0] - c12[0]), F(c21[1] - c12[1])
                ok = F(F(ex * ex) + F(ey * ey)) < ep_thr
            if ok:
                nb[i1].append(i2)
                nb[i2].append(i1)
    code = [-1] * n
    clusters = []
    for i in range(n):
        if code[i] >= 0:
            continue
        cid = len(clusters)
        code[i] = cid
        cl, check = [i], list(nb[i])
        while check:
            tmp = set()
            for j in check:
                if code[j] < 0:
                    code[j] = cid
                    cl.append(j)
                for k in nb[j]:
                    if code[k] < 0:
                        tmp.add(k)
            check = sorted(tmp)
        clusters.append(cl)
    subs = []
    for cl in clusters:
        if len(cl) <= 2:
            subs.append(cl)
            continue
        cl = sorted(cl, key=lambda i: -float(length[i]))  # stable, descending
        loc = {v: k for k, v in enumerate(cl)}
        done = [False] * len(cl)
        for j, li in enumerate(cl):
            if done[j]:
                continue
            sub = [li]
            for k in nb[li]:
                done[loc[k]] = True
                sub.append(k)
            subs.append(sub)
    out = []
    for c in subs:
        nl = src[c[0]].copy()
        for k in c:
            nl = merge_two(nl, src[k])
        out.append(nl)
    return np.array(out, F).reshape(-1, 4)


def filter_short(lines, thr):
    if len(lines) == 0:
        return lines
    dx, dy = lines[:, 2] - lines[:, 0], lines[:, 3] - lines[:, 1]
    keep = (dx * dx + dy * dy).astype(F) > F(thr) * F(thr)
    return lines[keep]


def optimize_and_merge(lines):
    t = filter_short(merge_lines(lines, 0.05, 5, 15), 30)
    return filter_short(merge_lines(t, 0.03, 3, 30), 50)


def keylines(lines, w, h, nfeatures):
    out = []
    for i, l in enumerate(lines):
        sx, sy, ex, ey = (F(v) for v in l)
        length = F(math.sqrt(float(F(sx - ex)) ** 2 + float(F(sy - ey)) ** 2))
        ok, a, b = cv2.clipLine((0, 0, w, h), (int(np.rint(sx)), int(np.rint(sy))), (int(np.rint(ex)), int(np.rint(ey))))
        npx = max(abs(b[0] - a[0]), abs(b[1] - a[1])) + 1 if ok else 0
        out.append(dict(angle=F(math.atan2(float(F(ey - sy)), float(F(ex - sx)))), class_id=i, octave=0,
                        pt_x=F(F(ex + sx) / F(2)), pt_y=F(F(ey + sy) / F(2)), response=F(length / F(max(w, h))),
                        size=F(F(ex - sx) * F(ey - sy)), start_x=sx, start_y=sy, end_x=ex, end_y=ey, s_oct_x=sx,
                        s_oct_y=sy, e_oct_x=ex, e_oct_y=ey, line_length=length, num_pixels=npx))
    if len(out) > nfeatures:
        out = sorted(out, key=lambda k: -float(k["response"]))[:nfeatures]
        for i, k in enumerate(out):
            k["class_id"] = i
    return out


_COMB = [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (2, 3), (2, 4), (2, 5),
         (2, 6), (2, 7), (2, 8), (3, 4), (3, 5), (3, 6), (3, 7), (3, 8), (4, 5), (4, 6), (4, 7), (4, 8), (5, 6), (5, 7),
         (5, 8), (6, 7), (6, 8), (7, 8)]
_GL = [math.exp((i - 10.0) ** 2 * (-1 / (2 * 7.0 * 7.0))) for i in range(21)]   # u = 20//2, sigma = 15//2
_GG = [math.exp((i - 31.0) ** 2 * (-1 / (2 * 31.0 * 31.0))) for i in range(63)]  # u = sigma = 62//2


def _c_round(v):
    v = float(v)
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


def lbd(dx, dy, kl):
    h, w = dx.shape
    band = {k: [F(0)] * 9 for k in ("pL", "nL", "pL2", "nL2", "pO", "nO", "pO2", "nO2")}
    n = int(kl["num_pixels"])
    half_w, half_h = (n - 1) // 2, 31
    if n <= 0:
        half_w = int((n - 1) / 2)  # C division truncates towards zero
    mx = F(0.5 * float(F(kl["s_oct_x"] + kl["e_oct_x"])))
    my = F(0.5 * float(F(kl["s_oct_y"] + kl["e_oct_y"])))
    dl0, dl1 = F(math.cos(float(kl["angle"]))), F(math.sin(float(kl["angle"])))
    do0, do1 = F(-dl1), dl0
    x0 = F(F(F(F(-dl0) * F(half_w)) + F(dl1 * F(half_h))) + mx)
    y0 = F(F(F(F(-dl1) * F(half_w)) - F(dl0 * F(half_h))) + my)
    for hid in range(63):
        sx, sy = x0, y0
        pL = nL = pO = nO = F(0)
        for _ in range(n):
            xc = min(max(_c_round(sx), 0), w - 1)
            yc = min(max(_c_round(sy), 0), h - 1)
            gx, gy = F(dx[yc, xc]), F(dy[yc, xc])
            gdl = F(F(gx * dl0) + F(gy * dl1))
            gdo = F(F(gx * do0) + F(gy * do1))
            if gdl > 0:
                pL = F(pL + gdl)
            else:
                nL = F(nL - gdl)
            if gdo > 0:
                pO = F(pO + gdo)
            else:
                nO = F(nO - gdo)
            sx, sy = F(sx + dl0), F(sy + dl1)
        x0, y0 = F(x0 - dl1), F(y0 + dl0)
        c = F(_GG[hid])
        pL, nL, pO, nO = F(c * pL), F(c * nL), F(c * pO), F(c * nO)
        sq = dict(pL2=F(pL * pL), nL2=F(nL * nL), pO2=F(pO * pO), nO2=F(nO * nO))
        lin = dict(pL=pL, nL=nL, pO=pO, nO=nO)

        def add(b, cc):
            cc = F(cc)
            for k, v in lin.items():
                band[k][b] = F(band[k][b] + F(cc * v))
            for k, v in sq.items():
                band[k][b] = F(band[k][b] + F(F(cc * cc) * v))
        b = hid // 7
        add(b, _GL[hid % 7 + 7])
        if b - 1 >= 0:
            add(b - 1, _GL[hid % 7 + 14])
        if b + 1 < 9:
            add(b + 1, _GL[hid % 7])
    des = np.zeros(72, F)
    inv2, inv3 = F(1.0 / 14.0), F(1.0 / 21.0)
    for b in range(9):
        inv = inv2 if b in (0, 8) else inv3
        for k, (m, s2) in enumerate((("pL", "pL2"), ("nL", "nL2"), ("pO", "pO2"), ("nO", "nO2"))):
            t = F(band[m][b] * inv)
            des[8 * b + k] = t
            v = float(F(F(band[s2][b] * inv) - F(t * t)))
            des[8 * b + 4 + k] = F(math.sqrt(v)) if v >= 0 else F(np.nan)
    tm = ts = F(0)
    for b in range(9):
        for k in range(4):
            tm = F(tm + F(des[8 * b + k] * des[8 * b + k]))
        for k in range(4, 8):
            ts = F(ts + F(des[8 * b + k] * des[8 * b + k]))
    tm, ts = F(1 / math.sqrt(float(tm))), F(1 / math.sqrt(float(ts)))
    for b in range(9):
        des[8 * b:8 * b + 4] = (des[8 * b:8 * b + 4] * tm).astype(F)
        des[8 * b + 4:8 * b + 8] = (des[8 * b + 4:8 * b + 8] * ts).astype(F)
    des[des.astype(np.float64) > 0.4] = F(0.4)
    t = F(0)
    for v in des:
        t = F(t + F(v * v))
    t = F(1 / math.sqrt(float(t)))
    des = (des * t).astype(F)
    bits = np.zeros(32, np.uint8)
    for c, (a, b) in enumerate(_COMB):
        r = 0
        for i in range(8):
            if des[8 * a + i] > des[8 * b + i]:
                r += 1 << i
        bits[c] = r
    return des, bits


def line_extract(img, nfeatures=200):
    """LINEextractor::operator() (LineExtractor.cpp:325-366) over cv2 primitives."""
    h, w = img.shape
    raw = lsd_cv2(img)
    merged = optimize_and_merge(clamp_segments(raw, w, h))
    kls = keylines(merged, w, h, nfeatures)
    bl = cv2.GaussianBlur(img, (5, 5), 1)
    dx = cv2.Sobel(bl, cv2.CV_16S, 1, 0, ksize=3)
    dy = cv2.Sobel(bl, cv2.CV_16S, 0, 1, ksize=3)
    des, bits, eq = [], [], []
    for k in kls:
        d, b = lbd(dx, dy, k)
        des.append(d)
        bits.append(b)
        sx, sy, ex, ey = (float(k[f]) for f in ("start_x", "start_y", "end_x", "end_y"))
        l = np.array([sy - ey, ex - sx, sx * ey - sy * ex])
        eq.append(l / math.sqrt(l[0] * l[0] + l[1] * l[1]))
    return raw, merged, kls, np.array(des, F).reshape(-1, 72), np.array(bits, np.uint8).reshape(-1, 32), \
        np.array(eq, np.float64).reshape(-1, 3)
