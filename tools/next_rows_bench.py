#!/usr/bin/env python
"""Timing of the widened ("next") rows on one B200 next to the CPU oracle on the box's host (development tool; the
numbers go to DESIGN.md §6).  Batched device forms are timed with CUDA events on device-resident data; the single-pair
host-pointer matchers are timed per call (they include their own H2D / D2H)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    from conftest import load_golden
    from oracle import orc
    from psl_slam_b200 import (KEYLINE_DTYPE, Context, FrameData, LSDmatcher, ORBmatcher, _lib, default_config,
                               make_distortion, plane_hypotheses)
    from psl_slam_b200._lib import KP_DTYPE
    ctx = Context(default_config())
    L = _lib.lib()
    out = {}

    def ev_time(fn, reps=5):
        fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(ctx.stream())
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st)
        ctx.sync()
        return e0.elapsed_time(e1) / reps

    def cpu_time(fn, min_s=0.5):
        fn()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < min_s:
            fn()
            n += 1
        return (time.perf_counter() - t0) / n * 1e3

    # --- N2: Frame::isLineGood, 1024 frames x 124 lines, device resident --------------------------------------------
    g = load_golden("lines3d_noisy")
    B, kl = 1024, np.ascontiguousarray(g["kl"], KEYLINE_DTYPE)
    n, (h, w) = len(kl), g["depth"].shape
    cam = [float(v) for v in g["cam"]]
    d_kl = torch.from_numpy(np.tile(kl.view(np.uint8).reshape(1, n, 68), (B, 1, 1))).cuda()
    d_n = torch.full((B,), n, dtype=torch.int32, device="cuda")
    d_dep = torch.from_numpy(g["depth"]).cuda().unsqueeze(0).repeat(B, 1, 1).contiguous()
    d_l3 = torch.zeros((B, n, 6), dtype=torch.float64, device="cuda")
    d_eq = torch.zeros((B, n, 3), dtype=torch.float32, device="cuda")
    ms = ev_time(lambda: ctx.check(L.psl_lines_3d_dev(ctx.handle, d_kl.data_ptr(), d_n.data_ptr(), n, B, d_dep.data_ptr(), w, h, w,
                                                       h * w, C.c_float(cam[0]), C.c_float(cam[1]), C.c_float(cam[2]),
                                                       C.c_float(cam[3]), C.c_uint32(3), d_l3.data_ptr(), d_eq.data_ptr())))
    cpu = cpu_time(lambda: orc.lines_3d(kl, g["depth"], *cam, 3))
    out["lines_3d (isLineGood), 124 lines/frame"] = {"gpu_us_per_frame": ms * 1e3 / B, "cpu_oracle_us_per_frame_1thread": cpu * 1e3}

    # --- N3: UndistortKeyPoints, 1024 frames x 1000 keypoints ----------------------------------------------------------
    gu = load_golden("undistort")
    d = make_distortion(*[float(v) for v in gu["cam_tum1"]])
    k = np.zeros(1000, KP_DTYPE)
    k["x"], k["y"] = gu["pts_rnd"][:1000, 0], gu["pts_rnd"][:1000, 1]
    dk = torch.from_numpy(np.tile(k.view(np.uint8).reshape(1, 1000, 28), (B, 1, 1))).cuda()
    dn = torch.full((B,), 1000, dtype=torch.int32, device="cuda")
    do = torch.zeros_like(dk)
    ms = ev_time(lambda: ctx.check(L.psl_undistort_keypoints_dev(ctx.handle, dk.data_ptr(), dn.data_ptr(), 1000, B, C.byref(d),
                                                                  do.data_ptr())))
    cpu = cpu_time(lambda: orc.undistort_keypoints(k, d))
    out["undistort_keypoints, 1000 kps/frame"] = {"gpu_us_per_frame": ms * 1e3 / B, "cpu_oracle_us_per_frame_1thread": cpu * 1e3}

    # --- single-pair host-pointer calls (latency per call, copies included) -------------------------------------------
    def call_time(fn, reps=30):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    gf = load_golden("linefuse_pair0")
    m = LSDmatcher(0.75)
    out["line_fuse, 124 lines x 124 queries"] = {
        "gpu_ms_per_call": call_time(lambda: m.Fuse(gf["kl"], gf["kf_desc"], gf["queries"], gf["qdesc"])),
        "cpu_oracle_ms_per_call": cpu_time(lambda: orc.line_fuse(gf["kl"], gf["kf_desc"], gf["queries"], gf["qdesc"]))}
    gp = load_golden("planes_case1")
    out["plane_hypotheses, 118 junctions"] = {
        "gpu_ms_per_call": call_time(lambda: plane_hypotheses(ctx, gp["kl"], gp["line_eq"], gp["lines3d"], gp["junctions"])),
        "cpu_oracle_ms_per_call": cpu_time(lambda: orc.plane_hypotheses(gp["kl"], gp["line_eq"], gp["lines3d"], gp["junctions"]))}
    gr = load_golden("reloc_pair0a")
    cur = FrameData(gr["kps_cur"], gr["desc_cur"], None, tuple(gr["bounds"]))
    om = ORBmatcher(0.9, True)
    q = gr["queries"].copy()
    q["flags"] = np.where(q["flags"] & 1, 3, 0).astype(np.uint32)
    out["SearchByProjection (relocalization), 767 queries x 1000 kps"] = {
        "gpu_ms_per_call": call_time(lambda: om.SearchByProjectionKeyFrame(cur, gr["queries"], gr["desc_kf"], gr["held"], 100)),
        "cpu_oracle_ms_per_call": cpu_time(lambda: orc.match_projection(gr["kps_cur"], None, gr["desc_cur"], tuple(gr["bounds"]), q,
                                                                        gr["desc_kf"], gr["held"], 0, 100, 0.9, True))}
    # --- N2: junction detection, 1024 frames x 124 lines, device resident ---------------------------------------------
    from psl_slam_b200._lib import JUNCTION_DTYPE
    gj = load_golden("junctions_case1")
    klj, l3j = np.ascontiguousarray(gj["kl"], KEYLINE_DTYPE), np.ascontiguousarray(gj["lines3d"], np.float64)
    nj, capj = len(klj), 512
    d_klj = torch.from_numpy(np.tile(klj.view(np.uint8).reshape(1, nj, 68), (B, 1, 1))).cuda()
    d_l3j = torch.from_numpy(np.tile(l3j.reshape(1, nj, 6), (B, 1, 1))).cuda()
    d_nj = torch.full((B,), nj, dtype=torch.int32, device="cuda")
    d_f = torch.zeros((B, capj, 4), dtype=torch.float32, device="cuda")
    d_j = torch.zeros((B, capj * JUNCTION_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    d_c = torch.zeros((2, B), dtype=torch.int32, device="cuda")
    ms = ev_time(lambda: ctx.check(L.psl_line_junctions_dev(ctx.handle, d_klj.data_ptr(), d_nj.data_ptr(), nj, B, d_l3j.data_ptr(),
                                                             640, 480, C.c_float(20.0), C.c_float(float(gj["fan_thr"])),
                                                             d_f.data_ptr(), d_j.data_ptr(), capj, d_c.data_ptr(),
                                                             d_c.data_ptr() + 4 * B)))
    cpu = cpu_time(lambda: orc.line_junctions(klj, l3j, 640, 480, 20.0, float(gj["fan_thr"])))
    out["line_junctions (CPartiallyRecoverConnectivity + cross points), 124 lines/frame"] = {
        "gpu_us_per_frame": ms * 1e3 / B, "cpu_oracle_us_per_frame_1thread": cpu * 1e3}

    # --- N1, second batch: loop-closing / initialisation matchers, single pair, host pointers ---------------------------
    gl = load_golden("loop_pair0")
    bnd = tuple(gl["bounds"])
    k1, k2 = FrameData(gl["kps1"], gl["desc1"], None, bnd), FrameData(gl["kps2"], gl["desc2"], None, bnd)
    c1, c2 = (gl["nodes1"], gl["offs1"], gl["idx1"]), (gl["nodes2"], gl["offs2"], gl["idx2"])
    a1, a2 = gl["kps1"]["angle"], gl["kps2"]["angle"]
    mb = ORBmatcher(0.75, True)
    out["SearchByBoW(KF, KF), 1006 x 1008 kps"] = {
        "gpu_ms_per_call": call_time(lambda: mb.SearchByBoWKeyFrames(gl["desc1"], a1, gl["valid1"], c1, gl["desc2"], a2, gl["valid2"], c2)),
        "cpu_oracle_ms_per_call": cpu_time(lambda: orc.match_bow_kf(gl["desc1"], a1, gl["valid1"], c1, gl["desc2"], a2, gl["valid2"], c2))}
    out["SearchBySim3, 1006 + 1008 queries"] = {
        "gpu_ms_per_call": call_time(lambda: mb.SearchBySim3(k1, k2, gl["q12"], gl["desc1"], gl["q21"], gl["desc2"])),
        "cpu_oracle_ms_per_call": cpu_time(lambda: orc.match_sim3((gl["kps1"], gl["desc1"]), (gl["kps2"], gl["desc2"]), bnd,
                                                                  gl["q12"], gl["desc1"], gl["q21"], gl["desc2"], 100))}
    mi = ORBmatcher(0.9, True)
    out["SearchForInitialization, window 100"] = {
        "gpu_ms_per_call": call_time(lambda: mi.SearchForInitialization(gl["kps1"], gl["desc1"], k2, gl["prev_matched"], 100)),
        "cpu_oracle_ms_per_call": cpu_time(lambda: orc.match_initialization(gl["kps1"], gl["desc1"], gl["prev_matched"],
                                                                            (gl["kps2"], gl["desc2"]), bnd, 100, 0.9, 50, True))}
    # --- N4: pose-only optimisation (point edges), 1024 frames x 600 points, device resident ----------------------------
    from psl_slam_b200._lib import POSE_POINT_DTYPE
    gp4 = load_golden("pose_case0")
    npt = len(gp4["pts"])
    fxp, fyp, cxp, cyp, bfp = (float(v) for v in gp4["cam"])
    d_pp = torch.from_numpy(np.tile(np.ascontiguousarray(gp4["pts"], POSE_POINT_DTYPE).view(np.uint8).reshape(1, -1), (B, 1))).cuda()
    d_T0 = torch.from_numpy(np.tile(gp4["Tcw0"].reshape(1, 16), (B, 1))).cuda()
    d_np = torch.full((B,), npt, dtype=torch.int32, device="cuda")
    d_T1 = torch.zeros_like(d_T0)
    d_ol = torch.zeros(B * npt, dtype=torch.uint8, device="cuda")
    d_ci = torch.zeros(B, dtype=torch.int32, device="cuda")
    ms = ev_time(lambda: ctx.check(L.psl_pose_optimization_dev(ctx.handle, d_T0.data_ptr(), d_pp.data_ptr(), d_np.data_ptr(), npt, B,
                                                                C.c_float(fxp), C.c_float(fyp), C.c_float(cxp), C.c_float(cyp),
                                                                C.c_float(bfp), d_T1.data_ptr(), d_ol.data_ptr(), d_ci.data_ptr())))
    cpu = cpu_time(lambda: orc.pose_optimization(gp4["Tcw0"], gp4["pts"], fxp, fyp, cxp, cyp, bfp))
    out["PoseOptimization (point edges), 600 points/frame"] = {"gpu_us_per_frame": ms * 1e3 / B,
                                                              "cpu_oracle_us_per_frame_1thread": cpu * 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
