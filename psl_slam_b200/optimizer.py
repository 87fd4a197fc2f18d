"""Host-side mirror of ORB_SLAM2::Optimizer::PoseOptimization (src/Optimizer.cc:239-1023 in the reference) over the
C-ABI ("next" row N4): point edges and, with `lils`, the structural-line edges EdgeLILSE3ProjectXYZ (Optimizer.cc:619-693,
add_inc/EdgeLIL.h:210-374).  The reference takes a Frame*; here the caller passes what it reads from it."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import POSE_LIL_DTYPE, POSE_POINT_DTYPE
from .orb import Context, _ptr


def PoseOptimization(ctx: Context, Tcw, points, fx, fy, cx, cy, bf, lils=None):
    """points: POSE_POINT_DTYPE [N] (one per keypoint; flags & 1 = it holds a MapPoint); lils: POSE_LIL_DTYPE [N_LJL] (one
    per structural line of the frame; flags & 1 = it holds a map InsectLine that is not bad) or None for a frame
    without structural lines.  Returns (Tcw [4,4] f32 = what pFrame->SetPose receives, mvbOutlier [N] u8,
    nInitialCorrespondences - nBad) and, with lils, additionally mvbOutlier_Insec [N_LJL] u8."""
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
    p = np.ascontiguousarray(points, POSE_POINT_DTYPE)
    out = np.zeros(16, np.float32)
    bad = np.zeros(max(len(p), 1), np.uint8)
    n = C.c_int32()
    if lils is not None:
        l = np.ascontiguousarray(lils, POSE_LIL_DTYPE)
        lbad = np.zeros(max(len(l), 1), np.uint8)
        ctx.check(_lib.lib().psl_pose_optimization_lil(ctx.handle, _ptr(T), _ptr(p), len(p), _ptr(l), len(l), C.c_float(fx),
                                                       C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(bf), _ptr(out),
                                                       _ptr(bad), _ptr(lbad), C.byref(n)))
        return out.reshape(4, 4), bad[: len(p)], n.value, lbad[: len(l)]
    ctx.check(_lib.lib().psl_pose_optimization(ctx.handle, _ptr(T), _ptr(p), len(p), C.c_float(fx), C.c_float(fy),
                                               C.c_float(cx), C.c_float(cy), C.c_float(bf), _ptr(out), _ptr(bad),
                                               C.byref(n)))
    return out.reshape(4, 4), bad[: len(p)], n.value
