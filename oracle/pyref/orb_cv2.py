"""TEST INFRASTRUCTURE — golden-vector generator, never imported by the product.

Literal restatement of the reference's ORB extractor control flow
(/root/reference/src/ORBextractor.cc) on top of the *real* OpenCV primitives
it calls, as exposed by cv2 4.13 in the build container: cv2.resize
(ORBextractor.cc:1120), cv2.FAST per cell (:809,814), cv2.GaussianBlur (:1086),
cv2.fastAtan2 (:103).  Only the Python control flow is ours.  The outputs are
committed under tests/golden/ and pin oracle/c (the self-contained restatement)
and the CUDA path.

Pinned choices: H1 octree tie-break = node creation sequence,
H2 cos/sin evaluated in double then rounded to float, H3 no FMA contraction,
H4 cv2 4.13 arithmetic.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

F32 = np.float32
EDGE_THRESHOLD = 19
HALF_PATCH = 15
PATCH_SIZE = 31


def cv_round(x) -> int:
    """cvRound: round-half-to-even."""
    return int(np.rint(x))


def load_pattern() -> np.ndarray:
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    inc = os.path.join(here, "..", "..", "psl_slam_b200", "csrc", "orb_pattern.inc")
    txt = "".join(l for l in open(inc) if not l.startswith("//"))
    v = np.array([int(t) for t in txt.replace("\n", "").split(",") if t.strip()], np.int32)
    assert v.size == 1024
    return v.reshape(512, 2)


class OrbParams:
    """ORBextractor ctor tables, ORBextractor.cc:410-470."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf_d = float(F32(scale_factor))  # float ctor arg stored in a double member
        self.scale = [F32(1.0)]
        for _ in range(1, nlevels):
            self.scale.append(F32(float(self.scale[-1]) * sf_d))  # :421 float*double -> float
        self.inv_scale = [F32(1.0) / s for s in self.scale]
        self.sigma2 = [s * s for s in self.scale]
        self.inv_sigma2 = [F32(1.0) / s for s in self.sigma2]
        factor = F32(1.0 / sf_d)  # :437
        nd = F32(nfeatures) * (F32(1) - factor) / (F32(1) - F32(math.pow(float(factor), float(nlevels))))
        self.quota, tot = [], 0
        for _ in range(nlevels - 1):
            q = cv_round(nd)
            self.quota.append(q)
            tot += q
            nd = F32(nd * factor)
        self.quota.append(max(nfeatures - tot, 0))
        # umax :454-469
        umax = [0] * (HALF_PATCH + 2)
        vmax = int(math.floor(F32(HALF_PATCH) * F32(math.sqrt(F32(2.0))) / 2 + 1))
        vmin = int(math.ceil(F32(HALF_PATCH) * F32(math.sqrt(F32(2.0))) / 2))
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(HALF_PATCH * HALF_PATCH - v * v))
        v0 = 0
        for v in range(HALF_PATCH, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax[: HALF_PATCH + 1]
        self.pattern = load_pattern()

    def level_sizes(self, w, h):
        return [(cv_round(F32(w) * s), cv_round(F32(h) * s)) for s in self.inv_scale]  # :1111-1112


def compute_pyramid(img: np.ndarray, P: OrbParams):
    """ORBextractor.cc:1107-1132.  Borders are never read on this path, so they are not built."""
    pyr = [img]
    for (w, h) in P.level_sizes(img.shape[1], img.shape[0])[1:]:
        pyr.append(cv2.resize(pyr[-1], (w, h), interpolation=cv2.INTER_LINEAR))
    return pyr


def fast_cells(level_img: np.ndarray, P: OrbParams):
    """Per-cell FAST with threshold fallback, ORBextractor.cc:765-829.
    Returns list of (x, y, response) with coords relative to minBorder, in reference order."""
    rows, cols = level_img.shape
    min_bx = min_by = EDGE_THRESHOLD - 3
    max_bx = cols - EDGE_THRESHOLD + 3
    max_by = rows - EDGE_THRESHOLD + 3
    width = F32(max_bx - min_bx)
    height = F32(max_by - min_by)
    n_cols = int(width / F32(30))
    n_rows = int(height / F32(30))
    w_cell = int(math.ceil(width / F32(n_cols)))
    h_cell = int(math.ceil(height / F32(n_rows)))
    det_ini = cv2.FastFeatureDetector_create(P.ini_th, True)
    det_min = cv2.FastFeatureDetector_create(P.min_th, True)
    out = []
    for i in range(n_rows):
        ini_y = min_by + i * h_cell
        max_y = ini_y + h_cell + 6
        if ini_y >= max_by - 3:
            continue
        max_y = min(max_y, max_by)
        for j in range(n_cols):
            ini_x = min_bx + j * w_cell
            max_x = ini_x + w_cell + 6
            if ini_x >= max_bx - 6:
                continue
            max_x = min(max_x, max_bx)
            sub = np.ascontiguousarray(level_img[ini_y:max_y, ini_x:max_x])
            kps = det_ini.detect(sub)
            if len(kps) == 0:
                kps = det_min.detect(sub)
            for kp in kps:
                out.append((F32(kp.pt[0]) + F32(j * w_cell), F32(kp.pt[1]) + F32(i * h_cell), F32(kp.response)))
    return out


class _Node:
    __slots__ = ("ULx", "ULy", "URx", "BRy", "keys", "no_more", "seq")

    def __init__(self):
        self.keys = []
        self.no_more = False
        self.seq = -1


def _divide(n: _Node):
    """ExtractorNode::DivideNode, ORBextractor.cc:481-537 (key lists keep candidate order)."""
    half_x = int(math.ceil(F32(n.URx - n.ULx) / F32(2)))
    half_y = int(math.ceil(F32(n.BRy - n.ULy) / F32(2)))
    c = [_Node() for _ in range(4)]
    mx, my = n.ULx + half_x, n.ULy + half_y
    c[0].ULx, c[0].URx, c[0].ULy, c[0].BRy = n.ULx, mx, n.ULy, my
    c[1].ULx, c[1].URx, c[1].ULy, c[1].BRy = mx, n.URx, n.ULy, my
    c[2].ULx, c[2].URx, c[2].ULy, c[2].BRy = n.ULx, mx, my, n.BRy
    c[3].ULx, c[3].URx, c[3].ULy, c[3].BRy = mx, n.URx, my, n.BRy
    for k in n.keys:
        x, y = k[0], k[1]
        if x < mx:
            (c[0] if y < my else c[2]).keys.append(k)
        elif y < my:
            c[1].keys.append(k)
        else:
            c[3].keys.append(k)
    for ch in c:
        if len(ch.keys) == 1:
            ch.no_more = True
    return c


def distribute_octtree(cands, min_x, max_x, min_y, max_y, N):
    """ORBextractor::DistributeOctTree, ORBextractor.cc:539-763.  H1: pointer order := creation order."""
    n_ini = int(math.floor(float(F32(max_x - min_x) / F32(max_y - min_y)) + 0.5))  # round() half away, positive
    hX = F32(max_x - min_x) / F32(n_ini)
    seq = [0]

    def stamp(n):
        n.seq = seq[0]
        seq[0] += 1
        return n

    nodes = []  # list order front->back
    ini = []
    for i in range(n_ini):
        n = _Node()
        n.ULx = int(hX * F32(i))
        n.URx = int(hX * F32(i + 1))
        n.ULy, n.BRy = 0, max_y - min_y
        nodes.append(stamp(n))
        ini.append(n)
    for k in cands:
        ini[int(k[0] / hX)].keys.append(k)
    keep = []
    for n in nodes:
        if len(n.keys) == 1:
            n.no_more = True
            keep.append(n)
        elif len(n.keys) > 1:
            keep.append(n)
    nodes = keep

    finish = False
    while not finish:
        prev = len(nodes)
        to_expand = 0
        last_children = []
        front = []  # newly pushed-front nodes, newest first
        rest = []
        for n in nodes:
            if n.no_more:
                rest.append(n)
                continue
            for ch in _divide(n):
                if ch.keys:
                    stamp(ch)
                    front.insert(0, ch)
                    if len(ch.keys) > 1:
                        to_expand += 1
                        last_children.append(ch)
        nodes = front + rest
        if len(nodes) >= N or len(nodes) == prev:
            finish = True
        elif len(nodes) + to_expand * 3 > N:
            while not finish:
                prev = len(nodes)
                prev_children = sorted(last_children, key=lambda n: (len(n.keys), n.seq))
                last_children = []
                for n in reversed(prev_children):
                    for ch in _divide(n):
                        if ch.keys:
                            stamp(ch)
                            nodes.insert(0, ch)
                            if len(ch.keys) > 1:
                                last_children.append(ch)
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    finish = True
    res = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if k[2] > best[2]:
                best = k
        res.append(best)
    return res


def ic_angle(img: np.ndarray, x: int, y: int, umax) -> float:
    """IC_Angle, ORBextractor.cc:77-104."""
    m01 = m10 = 0
    for u in range(-HALF_PATCH, HALF_PATCH + 1):
        m10 += u * int(img[y, x + u])
    for v in range(1, HALF_PATCH + 1):
        d = umax[v]
        vs = 0
        for u in range(-d, d + 1):
            p, m = int(img[y + v, x + u]), int(img[y - v, x + u])
            vs += p - m
            m10 += u * (p + m)
        m01 += v * vs
    return cv2.fastAtan2(float(m01), float(m10))


_FACTOR_PI = F32(math.pi / float(F32(180.0)))  # :107


def orb_descriptors(blur: np.ndarray, kps, P: OrbParams) -> np.ndarray:
    """computeOrbDescriptor, ORBextractor.cc:108-147 (H2/H3 pinned)."""
    out = np.zeros((len(kps), 32), np.uint8)
    px = P.pattern[:, 0].astype(F32)
    py = P.pattern[:, 1].astype(F32)
    for i, (x, y, ang) in enumerate(kps):
        a_ = F32(ang) * _FACTOR_PI
        a = F32(math.cos(float(a_)))
        b = F32(math.sin(float(a_)))
        ry = np.rint((px * b).astype(F32) + (py * a).astype(F32)).astype(np.int64)
        rx = np.rint((px * a).astype(F32) - (py * b).astype(F32)).astype(np.int64)
        vals = blur[cv_round(y) + ry, cv_round(x) + rx].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        out[i] = np.packbits(bits.reshape(32, 8), axis=1, bitorder="little")[:, 0]
    return out


def orb_extract(img: np.ndarray, P: OrbParams, stages: dict | None = None):
    """ORBextractor::operator(), ORBextractor.cc:1043-1105.
    Returns (kps float32 [n,5] = x,y,size,angle,response ; octave int32 [n] ; desc u8 [n,32])."""
    assert img.dtype == np.uint8 and img.ndim == 2
    pyr = compute_pyramid(img, P)
    rows_out, oct_out, desc_out = [], [], []
    for lvl, im in enumerate(pyr):
        cands = fast_cells(im, P)
        min_b = EDGE_THRESHOLD - 3
        sel = distribute_octtree(cands, min_b, im.shape[1] - EDGE_THRESHOLD + 3, min_b,
                                 im.shape[0] - EDGE_THRESHOLD + 3, P.quota[lvl]) if cands else []
        size = F32(int(F32(PATCH_SIZE) * P.scale[lvl]))
        kl = []
        for (x, y, r) in sel:
            xx, yy = F32(x + F32(min_b)), F32(y + F32(min_b))
            ang = F32(ic_angle(im, cv_round(xx), cv_round(yy), P.umax))
            kl.append((xx, yy, ang, r))
        if stages is not None:
            stages.setdefault("level_img", []).append(im)
            stages.setdefault("cands", []).append(np.array(cands, F32).reshape(-1, 3))
            stages.setdefault("selected", []).append(np.array([(k[0], k[1], k[3]) for k in kl], F32).reshape(-1, 3))
        if not kl:
            continue
        blur = cv2.GaussianBlur(im.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        if stages is not None:
            stages.setdefault("blur", {})[lvl] = blur
        desc_out.append(orb_descriptors(blur, [(k[0], k[1], k[2]) for k in kl], P))
        sc = P.scale[lvl]
        for (xx, yy, ang, r) in kl:
            if lvl != 0:
                xx, yy = F32(xx * sc), F32(yy * sc)
            rows_out.append((xx, yy, size, ang, r))
            oct_out.append(lvl)
    kps = np.array(rows_out, F32).reshape(-1, 5)
    octv = np.array(oct_out, np.int32)
    desc = np.concatenate(desc_out, 0) if desc_out else np.zeros((0, 32), np.uint8)
    return kps, octv, desc
