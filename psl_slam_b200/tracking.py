"""Batched point front end of one tracking step (psl_track_orb_batch): what Frame::Frame (ExtractORB +
ComputeStereoFromRGBD, src/Frame.cc:133-210) and the SearchByProjection call of
Tracking::TrackWithMotionModel (src/Tracking.cc:1193) do per frame, for a batch of consecutive RGB-D frames."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KEYLINE_DTYPE, KP_DTYPE, Camera, FrontendOut, TrackParams
from .orb import ORBextractor, _ptr


def make_camera(fx, fy, cx, cy, bf, depth_map_factor=5000.0) -> Camera:
    f = np.float32(1.0) / np.float32(depth_map_factor)  # Tracking.cc:142-145
    return Camera(fx, fy, cx, cy, bf, float(f))


def make_track_params(th=15.0, nn_ratio=0.9, check_orientation=True, th_dist=100) -> TrackParams:
    return TrackParams(th, nn_ratio, int(check_orientation), th_dist)


def pose_rows(Tcw: np.ndarray) -> np.ndarray:
    """[B,4,4] or [B,3,4] poses -> contiguous float32 [B,12]."""
    T = np.asarray(Tcw, np.float32)
    return np.ascontiguousarray(T[:, :3, :4].reshape(len(T), 12))


def track_orb_batch(ex: ORBextractor, gray: np.ndarray, depth: np.ndarray, Tcw: np.ndarray, cam: Camera,
                    prm: TrackParams | None = None):
    """HOST arrays in, host arrays out.  Returns dict(kps, desc, n, u_right, z, assign, nmatches)."""
    prm = prm or make_track_params()
    gray = np.ascontiguousarray(gray, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    B, H, W = gray.shape
    T = pose_rows(Tcw)
    cap = ex.cap
    out = dict(kps=np.zeros((B, cap), KP_DTYPE), desc=np.zeros((B, cap, 32), np.uint8), n=np.zeros(B, np.int32),
               u_right=np.zeros((B, cap), np.float32), z=np.zeros((B, cap), np.float32),
               assign=np.full((B, cap), -1, np.int32), nmatches=np.zeros(B, np.int32))
    ex.ctx.check(_lib.lib().psl_track_orb_batch(ex.ctx.handle, _ptr(gray), _ptr(depth), B, W, H, _ptr(T),
                                                C.addressof(cam), C.addressof(prm), _ptr(out["kps"]),
                                                _ptr(out["desc"]), _ptr(out["n"]), _ptr(out["u_right"]),
                                                _ptr(out["z"]), _ptr(out["assign"]), _ptr(out["nmatches"]), cap))
    return out


def track_orb_batch_dev(ex: ORBextractor, d_gray: int, d_depth: int, B: int, W: int, H: int, d_Tcw: int, cam: Camera,
                        prm: TrackParams, d_kps: int, d_desc: int, d_n: int, d_u_right: int, d_z: int, d_assign: int,
                        d_nmatches: int, cap: int):
    """DEVICE pointers (tightly packed frames), asynchronous on the ctx stream."""
    ex.ctx.check(_lib.lib().psl_track_orb_batch_dev(ex.ctx.handle, d_gray, W, W * H, d_depth, W, W * H, B, W, H,
                                                    d_Tcw, C.addressof(cam), C.addressof(prm), d_kps, d_desc, d_n,
                                                    d_u_right, d_z, d_assign, d_nmatches, cap))


def track_frontend_batch(ex: ORBextractor, gray: np.ndarray, depth: np.ndarray, Tcw: np.ndarray, cam: Camera,
                         prm: TrackParams | None = None, line_desc_th: float = 0.95):
    """Combined front end (psl_track_frontend_batch): points as track_orb_batch plus LINEextractor per frame and
    LSDmatcher::SearchByGeomNApearance against the previous frame (Tracking.cc:1182-1183).  HOST arrays in/out.
    `ex.ctx` must be configured for the line path too (default config is)."""
    prm = prm or make_track_params()
    gray = np.ascontiguousarray(gray, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    B, H, W = gray.shape
    T = pose_rows(Tcw)
    cap, lc = ex.cap, int(ex.ctx.cfg.line_nfeatures)
    out = dict(kps=np.zeros((B, cap), KP_DTYPE), desc=np.zeros((B, cap, 32), np.uint8), n=np.zeros(B, np.int32),
               u_right=np.zeros((B, cap), np.float32), z=np.zeros((B, cap), np.float32),
               assign=np.full((B, cap), -1, np.int32), nmatches=np.zeros(B, np.int32),
               kl=np.zeros((B, lc), KEYLINE_DTYPE), ldesc=np.zeros((B, lc, 32), np.uint8),
               lineeq=np.zeros((B, lc, 3), np.float64), nl=np.zeros(B, np.int32),
               line_assign=np.full((B, lc), -1, np.int32), line_nmatches=np.zeros(B, np.int32))
    fo = FrontendOut(*[out[k].ctypes.data for k in ("kps", "desc", "n", "u_right", "z", "assign", "nmatches")], cap, lc,
                     *[out[k].ctypes.data for k in ("kl", "ldesc", "lineeq", "nl", "line_assign", "line_nmatches")])
    ex.ctx.check(_lib.lib().psl_track_frontend_batch(ex.ctx.handle, _ptr(gray), _ptr(depth), B, W, H, _ptr(T),
                                                     C.addressof(cam), C.addressof(prm), C.c_float(line_desc_th),
                                                     C.byref(fo)))
    return out


def track_frontend_batch_dev(ex: ORBextractor, d_gray: int, d_depth: int, B: int, W: int, H: int, d_Tcw: int,
                             cam: Camera, prm: TrackParams, line_desc_th: float, fo: FrontendOut):
    """DEVICE pointers (tightly packed frames; every FrontendOut pointer in HBM), asynchronous on the ctx stream."""
    ex.ctx.check(_lib.lib().psl_track_frontend_batch_dev(ex.ctx.handle, d_gray, W, W * H, d_depth, W, W * H, B, W, H,
                                                         d_Tcw, C.addressof(cam), C.addressof(prm),
                                                         C.c_float(line_desc_th), C.byref(fo)))


def convert_rgbd(ctx, color: np.ndarray | None, rgb_order: bool, depth: np.ndarray | None, depth_map_factor: float = 5000.0):
    """Input conversion of Tracking::GrabImageRGBD (src/Tracking.cc:219-235): colour [B,H,W,3|4] u8 -> gray [B,H,W] u8
    (cvtColor RGB/BGR/RGBA/BGRA2GRAY) and depth [B,H,W] u16 -> float32 * (1/DepthMapFactor).  HOST arrays."""
    gray = dep = None
    B = H = W = ch = 0
    if color is not None:
        color = np.ascontiguousarray(color, np.uint8)
        B, H, W, ch = color.shape
        gray = np.empty((B, H, W), np.uint8)
    if depth is not None:
        depth = np.ascontiguousarray(depth, np.uint16)
        B, H, W = depth.shape
        dep = np.empty((B, H, W), np.float32)
    f = np.float32(1.0) / np.float32(depth_map_factor)
    ctx.check(_lib.lib().psl_convert_rgbd(ctx.handle, None if color is None else _ptr(color), ch or 3, int(rgb_order),
                                          None if gray is None else _ptr(gray), None if depth is None else _ptr(depth),
                                          C.c_float(f), None if dep is None else _ptr(dep), B, W, H))
    return gray, dep


def make_distortion(fx, fy, cx, cy, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0):
    """mK and mDistCoef of Tracking's settings (Tracking.cc:52-77)."""
    return _lib.Distortion(fx, fy, cx, cy, k1, k2, p1, p2, k3)


def undistort_keypoints(ctx, kps: np.ndarray, cam) -> np.ndarray:
    """Frame::UndistortKeyPoints (Frame.cc:1062-1092): KP_DTYPE [n] -> mvKeysUn."""
    kps = np.ascontiguousarray(kps, _lib.KP_DTYPE)
    out = np.empty_like(kps)
    ctx.check(_lib.lib().psl_undistort_keypoints(ctx.handle, _ptr(kps), len(kps), C.byref(cam), _ptr(out)))
    return out


def image_bounds(ctx, cols: int, rows: int, cam):
    """Frame::ComputeImageBounds (Frame.cc:1135-1163): (mnMinX, mnMinY, mnMaxX, mnMaxY)."""
    b = np.zeros(4, np.float32)
    ctx.check(_lib.lib().psl_image_bounds(ctx.handle, int(cols), int(rows), C.byref(cam), _ptr(b)))
    return tuple(float(v) for v in b)


def extract_lsd_batch_dev(lex, d_gray: int, d_depth_u16: int, B: int, W: int, H: int, cam: Camera, seed: int = 0,
                          gray_stride: int | None = None, junction_cap: int = 2048):
    """Frame::ExtractLSD up to the plane hypotheses (src/Frame.cc:489-511) for B frames resident in HBM, chained on the
    context's stream without a round trip to the host:

        (*mpLSDextractorLeft)(im, mask, mvKeylinesUn, mLdesc, mvKeyLineFunctions)       psl_line_extract_batch_dev
        imDepth.convertTo(CV_32F, mDepthMapFactor)           (Tracking.cc:232-235)      psl_convert_rgbd_dev
        isLineGood(imGray, imDepth, K)                       (Frame.cc:662-750)         psl_lines_3d_dev
        CPartiallyRecoverConnectivity + convertFansToKeyLines (Frame.cc:504-507)        psl_line_junctions_dev

    d_gray: u8 [B][H][gray_stride], d_depth_u16: u16 [B][H][W] (device addresses).  Returns torch tensors on the device
    (rows past the per-frame counts are unspecified): kl [B,cap] bytes of KEYLINE_DTYPE, ldesc [B,cap,32], lineeq
    [B,cap,3] f64, n [B], depth [B,H,W] f32, lines3d [B,cap,6] f64, line_eq3 [B,cap,3] f32, fans [B,jcap,4] f32,
    junctions [B,jcap] bytes of JUNCTION_DTYPE, n_fans [B], n_junctions [B].  psl_plane_hypotheses (host pointers, one
    frame) takes rows of these arrays as they are."""
    import torch

    from ._lib import JUNCTION_DTYPE
    ctx, L = lex.ctx, _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    cap, gs = lex.cap, gray_stride or W
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
    out = {"kl": z((B, cap, KEYLINE_DTYPE.itemsize), torch.uint8), "ldesc": z((B, cap, 32), torch.uint8),
           "lineeq": z((B, cap, 3), torch.float64), "n": z((B,), torch.int32), "depth": z((B, H, W), torch.float32),
           "lines3d": z((B, cap, 6), torch.float64), "line_eq3": z((B, cap, 3), torch.float32),
           "fans": z((B, junction_cap, 4), torch.float32),
           "junctions": z((B, junction_cap, JUNCTION_DTYPE.itemsize), torch.uint8), "n_fans": z((B,), torch.int32),
           "n_junctions": z((B,), torch.int32)}
    p = {k: v.data_ptr() for k, v in out.items()}
    lex.extract_batch_dev(d_gray, B, W, H, gs, gs * H, p["kl"], p["ldesc"], p["lineeq"], None, p["n"])
    ctx.check(L.psl_convert_rgbd_dev(ctx.handle, None, 3, 1, 0, 0, None, 0, 0, d_depth_u16, W, W * H,
                                     C.c_float(cam.depth_factor), p["depth"], B, W, H))
    ctx.check(L.psl_lines_3d_dev(ctx.handle, p["kl"], p["n"], cap, B, p["depth"], W, H, W, W * H, C.c_float(cam.fx),
                                 C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy), C.c_uint32(seed), p["lines3d"],
                                 p["line_eq3"]))
    ctx.check(L.psl_line_junctions_dev(ctx.handle, p["kl"], p["n"], cap, B, p["lines3d"], W, H, C.c_float(20.0),
                                       C.c_float(float(np.float32(0.25 * np.pi))), p["fans"], p["junctions"],
                                       junction_cap, p["n_fans"], p["n_junctions"]))
    return out
