"""TEST INFRASTRUCTURE — float32 model of the Frame bookkeeping between extraction and matching
(RGB-D path): depth scaling (src/Tracking.cc:234-235), Frame::ComputeStereoFromRGBD and
UnprojectStereo (src/Frame.cc:1342-1381), and the projection loop at the top of
ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (src/ORBmatcher.cc:1339-1393).
cv::Mat float products accumulate in double and round once (OpenCV gemm for CV_32F).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"),
                        ("u_right", "<f4"), ("angle", "<f4"), ("flags", "<u4")])
assert QUERY_DTYPE.itemsize == 32


def gemm_affine(R, x, t):
    """float( sum_k double(R[i,k])*double(x[k]) + double(t[i]) ) — cv::gemm(R, x, 1, t, 1) on CV_32F."""
    R64, x64, t64 = R.astype(np.float64), x.astype(np.float64), t.astype(np.float64)
    out = np.empty(3, F32)
    for i in range(3):
        s = R64[i, 0] * x64[0] + R64[i, 1] * x64[1] + R64[i, 2] * x64[2]
        out[i] = F32(s + t64[i])
    return out


def depth_to_float(depth_u16, factor=5000.0):
    f = F32(1.0) / F32(factor)  # Tracking.cc:142-145
    return (depth_u16.astype(F32) * f).astype(F32)


def stereo_from_rgbd(kps_xy, depth_f, bf):
    """Frame::ComputeStereoFromRGBD: (mvuRight, mvDepth); undistorted == raw keypoints (zero distortion)."""
    n = len(kps_xy)
    u_right = np.full(n, -1, F32)
    dep = np.full(n, -1, F32)
    for i in range(n):
        d = depth_f[int(kps_xy[i, 1]), int(kps_xy[i, 0])]
        if d > 0:
            dep[i] = d
            u_right[i] = F32(F32(kps_xy[i, 0]) - F32(F32(bf) / d))
    return u_right, dep


def unproject(kps_xy, dep, K, Twc):
    """Frame::UnprojectStereo for every keypoint with depth; world points (n,3) float32, NaN where none."""
    fx, fy, cx, cy = (F32(K[k]) for k in ("fx", "fy", "cx", "cy"))
    invfx, invfy = F32(1.0) / fx, F32(1.0) / fy
    Rwc, Ow = Twc[:3, :3].astype(F32), Twc[:3, 3].astype(F32)
    out = np.full((len(dep), 3), np.nan, F32)
    for i, z in enumerate(dep):
        if z > 0:
            x = F32(F32(F32(kps_xy[i, 0]) - cx) * z) * invfx
            y = F32(F32(F32(kps_xy[i, 1]) - cy) * z) * invfy
            out[i] = gemm_affine(Rwc, np.array([x, y, z], F32), Ow)
    return out


def projection_queries(pw, last_octave, last_angle, valid, claims, Tcw_cur, Tcw_last, K, scale_factors, th, bounds,
                       mono=False):
    """The per-MapPoint prologue of SearchByProjection(CurrentFrame, LastFrame, th, bMono)."""
    fx, fy, cx, cy, bf = (F32(K[k]) for k in ("fx", "fy", "cx", "cy", "bf"))
    mb = bf / fx
    Rcw, tcw = Tcw_cur[:3, :3].astype(F32), Tcw_cur[:3, 3].astype(F32)
    Rlw, tlw = Tcw_last[:3, :3].astype(F32), Tcw_last[:3, 3].astype(F32)
    twc = gemm_affine(-Rcw.T, tcw, np.zeros(3, F32))
    tlc = gemm_affine(Rlw, twc, tlw)
    fwd = (tlc[2] > mb) and not mono
    bwd = (-tlc[2] > mb) and not mono
    min_x, min_y, max_x, max_y = bounds
    q = np.zeros(len(pw), QUERY_DTYPE)
    for i in range(len(pw)):
        if not valid[i] or np.isnan(pw[i, 0]):
            continue
        xc = gemm_affine(Rcw, pw[i], tcw)
        invz = F32(1.0 / float(xc[2]))
        if invz < 0:
            continue
        u = F32(F32(F32(fx * xc[0]) * invz) + cx)
        v = F32(F32(F32(fy * xc[1]) * invz) + cy)
        if u < min_x or u > max_x or v < min_y or v > max_y:
            continue
        o = int(last_octave[i])
        if fwd:
            lo, hi = o, -1
        elif bwd:
            lo, hi = 0, o
        else:
            lo, hi = o - 1, o + 1
        q[i] = (u, v, F32(F32(th) * scale_factors[o]), lo, hi, F32(u - F32(bf * invz)), last_angle[i],
                1 | (2 if claims[i] else 0))
    return q
