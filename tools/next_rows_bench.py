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
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
