#!/usr/bin/env python
"""Benchmark of the PSL-SLAM feature front end on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU path on the host cores

    python bench.py --config cfg1|cfg2|cfg3|cfg5              # the other BASELINE.json configs (default: cfg4)

One *step* = one pass of the hot path over a batch of synthetic frames (per GPU).  cfg4 (the configuration the
metric is quoted on): 4096 RGB-D frames of 640x480 per step, rendered on the device from a long synthetic
trajectory (512 distinct frames), through Tracking::GrabImageRGBD's conversion, both extractors and both
frame-to-frame matchers.
`value`   : whole-job frames/s with the RGB-D frames already resident in HBM (device-pointer C-ABI).
`e2e`     : the same through the host-pointer C-ABI (pinned host RGB + depth in, all results out; H2D and D2H
            inside the timed region; the two-phase begin/end form uploads batch k+1 while batch k is computed).
`roofline`: the dominant kernel stage, algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (a port of the reference's algorithm) timed on this box's cores.
Frames are sharded across GPUs with no collective: scaling is "weak".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 640, 480
ORB = dict(nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7)  # Examples/RGB-D/TUM1.yaml:42-55


def level_sizes(w, h, nlevels, scale):
    s = [np.float32(1.0)]
    for _ in range(1, nlevels):
        s.append(np.float32(float(s[-1]) * float(np.float32(scale))))
    return [(int(np.rint(np.float32(w) * (np.float32(1) / x))), int(np.rint(np.float32(h) * (np.float32(1) / x))))
            for x in s]


def algorithmic_bytes(w, h, nlevels, scale, nfeat):
    """Compulsory bytes per frame and stage, unfused, stage-wise."""
    P = [a * b for a, b in level_sizes(w, h, nlevels, scale)]
    return {
        "pyramid": sum(P[:-1]) + sum(P[1:]),          # read level l-1, write level l
        "fast": sum(P) + 60_000,                      # read every level once (+ candidates)
        "octree": 2 * 4 * 10_000,                     # candidate keys in, selected keys out (L2)
        "blur": 2 * sum(P),                           # read + write every level
        "describe": 2 * nfeat * 961 + nfeat * 60,     # 31x31 patches (x2 images) + records
        "match_single": 0,
        "stereo_queries": nfeat * (28 + 2 + 8 + 28 + 4 + 32),  # kp + depth px + (uRight,z) ; kp + z -> query
        "grid": nfeat * (28 + 2) + 3073 * 4,          # keypoints in, CSR out
        "candidates": nfeat * 20 * 32 + nfeat * 64,   # ~20 candidate descriptors per query
        "resolve": nfeat * 20 * 4 + nfeat * 4,        # candidate lists in, assignment out
        # line stages (LSD works on the 0.8x image, Ws x Hs = lrint(.8 w) x lrint(.8 h))
        "lsd_prologue": 2 * w * h + (w * h + LS(w, h)) + LS(w, h) * (1 + 4 + 4 + 1) + LS(w, h) * (8 + 6),
        #                 blur R+W    resize R + W        gradient R u8, W deg/n2/used   seed keys R, W (key, idx)
        "lsd_order": 2 * int(0.3 * LS(w, h)) * 6 * 2,  # two radix passes over the (u16 key, u32 idx) pairs of the seed-capable
        #                                                pixels (~30 % of the 0.8x image on the textured frames), R + W
        "lsd_grow": LS(w, h) * (4 + 4 + 1 + 1) + LS(w, h) * 4,  # deg, n2, used R+W once each + the seed list
        "line_merge": 4096 * 16 * 4,                  # raw segments through two merge passes (bound by the raw cap)
        "lbd": 2 * w * h + w * h + w * h * 4 + 200 * 63 * 120 * 4,  # blur R+W, Sobel R + W short2, 63 x len gathers/line
        "line_match": 2 * 200 * 32 + 200 * 68 * 2,    # two descriptor sets + keylines
    }


def LS(w, h):
    return int(round(w * 0.8)) * int(round(h * 0.8))


def synth_frames_rgb(n_distinct, seed=4):
    """n consecutive synthetic RGB-D frames rendered on the host (rgb u8 [n,H,W,3], depth u16) + poses [n,12] f32."""
    from psl_slam_b200 import synth
    poster = synth.make_poster(seed, 2048)
    T = synth.trajectory(n_distinct, seed)
    rgb, dep = zip(*[synth.render(poster, T[i], W, H, noise_seed=seed * 100003 + i) for i in range(n_distinct)])
    return np.stack(rgb), np.stack(dep), np.ascontiguousarray(T.astype(np.float32)[:, :3, :4].reshape(n_distinct, 12))


def synth_frames(n_distinct, seed=4):
    """n consecutive synthetic RGB-D frames (gray u8, depth u16) + world->camera poses [n,12] f32."""
    from psl_slam_b200 import synth
    gray, depth, T = synth.sequence(seed, n_distinct, W, H)
    return gray, depth, np.ascontiguousarray(T.astype(np.float32)[:, :3, :4].reshape(n_distinct, 12))


def render_sequence_cuda(seed, n, w, h, device, poster_size=2048, chunk=32):
    """n consecutive RGB-D views of the synthetic poster scene, rendered on the GPU (psl_slam_b200/synth.py: render, same
    geometry and intrinsics; bench input only).  Returns (rgb u8 [n,h,w,3], depth u16-bits int16 [n,h,w], Tcw [n,12] f32)."""
    import torch

    from psl_slam_b200 import synth
    K = synth.ICL
    poster = torch.from_numpy(synth.make_poster(seed, poster_size)).to(device).float()
    S = poster.shape[0]
    T = synth.trajectory(n, seed)
    sx, sy = w / 640.0, h / 480.0
    fx, fy, cx, cy = K["fx"] * sx, K["fy"] * sy, (K["cx"] + 0.5) * sx - 0.5, (K["cy"] + 0.5) * sy - 0.5
    v, u = torch.meshgrid(torch.arange(h, device=device, dtype=torch.float64),
                          torch.arange(w, device=device, dtype=torch.float64), indexing="ij")
    d = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], -1)          # [h,w,3]
    rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, device=device)
    depth = torch.empty((n, h, w), dtype=torch.int16, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 100003)
    mpp = 0.0022   # metres per poster pixel (synth.render)
    for i in range(n):
        Tc = torch.from_numpy(T[i]).to(device)
        Rwc = Tc[:3, :3].T
        twc = -Rwc @ Tc[:3, 3]
        dw = d @ Rwc.T
        t = -twc[2] / dw[..., 2]
        px = ((twc[0] + t * dw[..., 0]) / mpp + S / 2).clamp(0, S - 1.001)
        py = ((twc[1] + t * dw[..., 1]) / mpp + S / 2).clamp(0, S - 1.001)
        x0, y0 = px.floor().long(), py.floor().long()
        ax, ay = (px - x0).float()[..., None], (py - y0).float()[..., None]
        val = ((1 - ay) * ((1 - ax) * poster[y0, x0] + ax * poster[y0, x0 + 1])
               + ay * ((1 - ax) * poster[y0 + 1, x0] + ax * poster[y0 + 1, x0 + 1]))
        val = val + torch.randn(val.shape, generator=gen, device=device)
        rgb[i] = val.round().clamp(0, 255).to(torch.uint8)
        dq = (t * K["depth_factor"]).round().clamp(0, 65535).to(torch.int32)
        depth[i] = torch.where(dq > 32767, dq - 65536, dq).to(torch.int16)
    T12 = np.ascontiguousarray(T.astype(np.float32)[:, :3, :4].reshape(n, 12))
    return rgb, depth, T12


def gray_cuda(rgb):
    """cv2-4.x RGB -> Y in Q15 with torch ops, for the configs whose boundary takes gray frames."""
    import torch
    c = rgb.to(torch.int32)
    return ((c[..., 0] * 9798 + c[..., 1] * 19235 + c[..., 2] * 3735 + 16384) >> 15).to(torch.uint8)


def cam6():
    from psl_slam_b200 import synth
    K = synth.ICL
    return np.array([K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], np.float32(1.0) / np.float32(K["depth_factor"])],
                    np.float32)


def ping_pong(n_distinct, total):
    """Index pattern 0,1,..,n-1,n-2,..,1,0,1,.. : every consecutive pair of the long batch is a real
    consecutive pair of the short synthetic sequence (so SearchByProjection always has a valid prior)."""
    period = list(range(n_distinct)) + list(range(n_distinct - 2, 0, -1))
    return np.array([period[i % len(period)] for i in range(total)], np.int64)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.rows.append([t.strip() for t in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower() == "active"
                                                         for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_run(rgb, depth, T12, nthreads):
    """GrabImageRGBD for a batch on the host: cvtColor, both extractors, stereo, both frame-to-frame matchers."""
    from oracle import orc
    p = orc.params(ORB["nfeatures"], ORB["scale"], ORB["nlevels"], ORB["ini"], ORB["mn"])
    return orc.rgbd_frontend_batch_mt(rgb, True, depth, T12, cam6(), p, TRACK["th"], TRACK["nn_ratio"], TRACK["ori"],
                                      LINE["nfeatures"], LINE["desc_th"], nthreads)


def native_oracle():
    """The timed CPU arm uses an -O3 -march=native build made on this host; must be called before the
    oracle library is first loaded."""
    os.environ["PSL_ORACLE_NATIVE"] = "1"
    from oracle import orc
    orc.build()
    return os.path.basename(orc.lib().path)


def cpu_baseline(gray, depth, T12, budget_s: float = 14.0):
    """The oracle (CPU port of the reference path: ORB + LSD/LBD extraction + stereo + both frame-to-frame
    matchers) on a bounded sample of the same workload, 1 thread (how the reference runs, Frame.cc:179-180)
    and all threads."""
    build = native_oracle()
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    cpu_run(gray[:2], depth[:2], T12[:2], 1)
    t1 = (time.perf_counter() - t0) / 2
    n1 = max(3, min(len(gray), int(budget_s * 0.3 / t1)))
    t0 = time.perf_counter()
    cpu_run(gray[:n1], depth[:n1], T12[:n1], 1)
    fps1 = n1 / (time.perf_counter() - t0)
    total = max(len(gray), int(budget_s * 0.7 * fps1 * cores * 0.6))
    idx = ping_pong(len(gray), total)
    g, d, t = gray[idx], depth[idx], T12[idx]
    t0 = time.perf_counter()
    cpu_run(g, d, t, cores)
    fpsN = total / (time.perf_counter() - t0)
    return {"value": fpsN, "unit": "frames/s", "cores": cores, "kind": "port", "value_1_thread": fps1, "build": build,
            "sample": f"{total} RGB-D frames (cvtColor + ORB + LSD/LBD extract + stereo + SearchByProjection + "
                      f"SearchByGeomNApearance vs previous frame) on {cores} threads, one frame per task; {n1} frames on "
                      f"1 thread"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (oracle port)."""
    if rank != 0:
        return
    build = native_oracle()
    if args.config != "cfg4":
        print(json.dumps({"impl": "reference", "unavailable": f"the reference arm is defined for cfg4; {args.config} "
                          f"reports its CPU timing in cpu_baseline of the regular run"}))
        return
    gray, depth, T12 = synth_frames_rgb(24)
    cores = os.cpu_count() or 1
    per_step = 8 * cores   # enough tasks per thread that the step does not end on one straggler
    idx = ping_pong(len(gray), per_step)
    g, d, t = gray[idx], depth[idx], T12[idx]
    for _ in range(args.warmup):
        cpu_run(g, d, t, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(g, d, t, cores)
    dt = time.perf_counter() - t0
    fps = per_step * args.steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": workload_config(per_step, 24),
           "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "build": build,
                            "sample": f"{per_step} RGB-D frames per step (24 distinct), one frame per task on {cores} threads"},
           "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


METRIC = "frames/s ORB+LSD/LBD extract+match @640x480"
TRACK = dict(th=15.0, nn_ratio=0.9, ori=True)  # Tracking.cc:1166,1189
LINE = dict(nfeatures=200, desc_th=0.95)        # TUM1.yaml:60-63, Tracking.cc:1182


def workload_config(frames_per_gpu, distinct):
    return {"workload": "cfg4: batched combined front end per RGB-D frame — cvtColor RGB->GRAY (GrabImageRGBD) + ORBextractor "
                        "(TUM1.yaml: 1000 features, 8 levels, 1.2, FAST 20/7) + LINEextractor (LSD, long-line merge, top 200, "
                        "LBD) + ComputeStereoFromRGBD + SearchByProjection(Cur, Last, th=15) + "
                        "SearchByGeomNApearance(Cur, Last, 0.95) between consecutive synthetic textured 640x480 RGB-D "
                        "frames (ICL intrinsics)",
            "frames_per_step_per_gpu": frames_per_gpu, "distinct_frames": distinct, "width": W, "height": H,
            "l2_policy": "inputs larger than L2 (frames_per_step x 1.5 MB >> 126 MB), no flush",
            "parallelism": "one sequence of frames_per_step_per_gpu x n_gpus frames in contiguous shards (one-frame "
                           "halo re-extracted per shard), no collective"}


def load_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return peak, ("measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)")


def stage_table(ctx, alg, frames, steps, peak):
    st_ms, st_launch = ctx.profile_read()
    names = ctx.STAGES
    tot_ms = float(sum(st_ms[: len(names)]))
    stages = []
    for i, nme in enumerate(names):
        t_ms = float(st_ms[i])
        gbs = alg[nme] * frames * steps / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        stages.append({"stage": nme, "ms_per_step": t_ms / steps, "share": t_ms / tot_ms if tot_ms else 0,
                       "launches_per_step": int(st_launch[i]) // steps, "alg_bytes_per_frame": alg[nme],
                       "achieved_gbs": gbs, "frac": gbs / peak})
    return stages


def roofline_of(stages, frames, peak, peak_src):
    dom = max(stages, key=lambda s: s["ms_per_step"])
    nl = max(dom["launches_per_step"], 1)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom["stage"])
        if tj and tj.get("dram_bytes_per_frame"):
            traffic = tj["dram_bytes_per_frame"] * frames / nl
            traffic_src = f'{tj["kernel"]}: {tj["source"]}'
    except Exception:
        pass
    return {"bound": "hbm", "kernel": dom["stage"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "alg_bytes_per_launch": dom["alg_bytes_per_frame"] * frames / nl, "avg_launch_ms": dom["ms_per_step"] / nl,
            "note": "achieved = algorithmic bytes of the stage x frames / CUDA-event time of the stage"}


def extended_chain(ex, L, step_dev, stream, steps, F, W, H, cap, lcap, cam, d_depth, d_T, d_kps, d_ur, d_z, d_assign, d_n,
                   d_kl, d_nl, base_ms, jcap=1024):
    """One tracking step continued on the device arrays of the front end, everything on the context's stream:
    Frame::ExtractLSD's depth conversion, 3-D lines and junctions (Frame.cc:496-511) and the second half of
    TrackWithMotionModel (UnprojectStereo of the matched points + Optimizer::PoseOptimization, Tracking.cc:1193-1240).
    Timed like `value` (device-resident, CUDA events)."""
    import ctypes as C

    import torch
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
    depth_f = z((F, H, W), torch.float32)
    l3, eq3 = z((F, lcap, 6), torch.float64), z((F, lcap, 3), torch.float32)
    fans, junc = z((F, jcap, 4), torch.float32), z((F, jcap, 40), torch.uint8)
    n_fans, n_junc = z((F,), torch.int32), z((F,), torch.int32)
    T_out, outl, n_in = z((F, 16), torch.float32), z((F, cap), torch.uint8), z((F,), torch.int32)
    h = ex.ctx.handle

    def step():
        step_dev()
        ex.ctx.check(L.psl_convert_rgbd_dev(h, None, 3, 1, 0, 0, None, 0, 0, d_depth.data_ptr(), W, W * H,
                                            C.c_float(cam.depth_factor), depth_f.data_ptr(), F, W, H))
        ex.ctx.check(L.psl_lines_3d_dev(h, d_kl.data_ptr(), d_nl.data_ptr(), lcap, F, depth_f.data_ptr(), W, H, W, W * H,
                                        C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy),
                                        C.c_uint32(0), l3.data_ptr(), eq3.data_ptr()))
        ex.ctx.check(L.psl_line_junctions_dev(h, d_kl.data_ptr(), d_nl.data_ptr(), lcap, F, l3.data_ptr(), W, H,
                                              C.c_float(20.0), C.c_float(float(np.float32(0.25 * np.pi))), fans.data_ptr(),
                                              junc.data_ptr(), jcap, n_fans.data_ptr(), n_junc.data_ptr()))
        ex.ctx.check(L.psl_track_pose_batch_dev(h, d_kps.data_ptr(), d_ur.data_ptr(), d_z.data_ptr(), d_assign.data_ptr(),
                                                d_n.data_ptr(), cap, F, d_T.data_ptr(), C.addressof(cam), T_out.data_ptr(),
                                                outl.data_ptr(), n_in.data_ptr()))

    for _ in range(2):
        step()
    ex.ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    ex.ctx.sync()
    ms = e0.elapsed_time(e1) / steps
    return {"chain": "front end (value) + depth convertTo + Frame::isLineGood (3-D lines) + junctions + UnprojectStereo + "
                     "Optimizer::PoseOptimization on the same device arrays",
            "value": F / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "added_ms_per_step": ms - base_ms,
            "lines3d_per_frame": float((l3[:, :, 0] != 0).sum().item()) / F,
            "junctions_per_frame": float(n_junc.sum().item()) / F, "fans_over_cap": int((n_fans > jcap).sum().item()),
            "pose_inliers_per_frame": float(n_in.sum().item()) / max(F - 1, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (cfg4: 4096, cfg5: 512)")
    ap.add_argument("--distinct", type=int, default=512, help="distinct synthetic frames behind a cfg4 step")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--line-chunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    json_fd = 1
    if world > 1:
        # stdout carries exactly one JSON line.  NCCL writes its version banner to the C-level stdout when the
        # communicator comes up (NCCL_DEBUG=VERSION and above), so file descriptor 1 points at stderr for the rest of
        # the run and the JSON line goes to a duplicate of the original stdout.
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.config != "cfg4":
        from bench_small import run_small
        out = run_small(args, rank, world, local)
        if rank == 0 and out is not None:
            os.write(json_fd, (json.dumps(out) + "\n").encode())
        if world > 1:
            dist.destroy_process_group()
        return

    from psl_slam_b200 import Context, ORBextractor, default_config
    from psl_slam_b200._lib import FrontendOut

    F = args.frames or 4096
    cfg = default_config()
    cfg.device, cfg.max_width, cfg.max_height, cfg.max_batch = local, W, H, F + 1
    cfg.orb_nfeatures, cfg.orb_scale_factor, cfg.orb_nlevels = ORB["nfeatures"], ORB["scale"], ORB["nlevels"]
    cfg.orb_ini_th_fast, cfg.orb_min_th_fast = ORB["ini"], ORB["mn"]
    cfg.chunk_frames, cfg.line_nfeatures = args.chunk, LINE["nfeatures"]
    cfg.line_chunk_frames = args.line_chunk or min(F + 1, 4100)
    ex = ORBextractor(ctx=Context(cfg))
    cap, lcap = ex.cap, LINE["nfeatures"]
    from psl_slam_b200 import make_camera, make_track_params, synth
    K = synth.ICL
    cam = make_camera(K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], K["depth_factor"])
    tprm = make_track_params(TRACK["th"], TRACK["nn_ratio"], TRACK["ori"])
    # cfg 4: one long sequence (F frames per GPU, F x N in total) in contiguous shards; ranks > 0 re-extract the
    # frame before their range (one-frame halo) so that every owned frame is matched against its predecessor.
    # The sequence is D distinct consecutive views of the scene (rendered here, on the device) walked back and forth.
    from psl_slam_b200.shard import shard_with_halo
    D = max(2, min(args.distinct, F))
    base_rgb, base_d, base_T = render_sequence_cuda(4, D, W, H, torch.device("cuda", local))
    s0, s1, halo = shard_with_halo(F * world, rank, world)
    idx = torch.from_numpy(ping_pong(D, F * world)[s0:s1]).cuda()
    F_own, F = F, s1 - s0                                         # F now counts the halo frame too
    d_rgb = base_rgb[idx].contiguous()                            # [F,H,W,3] u8 resident in HBM
    d_depth = base_d[idx].contiguous()                            # u16 bits
    d_T = torch.from_numpy(base_T).cuda()[idx].contiguous()
    d_kps = torch.empty((F, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.empty((F, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_ur = torch.empty((F, cap), dtype=torch.float32, device="cuda")
    d_z = torch.empty((F, cap), dtype=torch.float32, device="cuda")
    d_assign = torch.empty((F, cap), dtype=torch.int32, device="cuda")
    d_nm = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_kl = torch.empty((F, lcap, 68), dtype=torch.uint8, device="cuda")
    d_ld = torch.empty((F, lcap, 32), dtype=torch.uint8, device="cuda")
    d_eq = torch.empty((F, lcap, 3), dtype=torch.float64, device="cuda")
    d_nl = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_la = torch.empty((F, lcap), dtype=torch.int32, device="cuda")
    d_lnm = torch.zeros(F, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ex.ctx.stream(), device=local)
    fo_dev = FrontendOut(d_kps.data_ptr(), d_desc.data_ptr(), d_n.data_ptr(), d_ur.data_ptr(), d_z.data_ptr(),
                         d_assign.data_ptr(), d_nm.data_ptr(), cap, lcap, d_kl.data_ptr(), d_ld.data_ptr(),
                         d_eq.data_ptr(), d_nl.data_ptr(), d_la.data_ptr(), d_lnm.data_ptr())
    import ctypes as C

    from psl_slam_b200 import _lib
    L = _lib.lib()

    def step_dev():
        ex.ctx.check(L.psl_track_rgbd_batch_dev(ex.ctx.handle, d_rgb.data_ptr(), 3, 1, W * 3, W * H * 3, d_depth.data_ptr(),
                                                W, W * H, F, W, H, d_T.data_ptr(), C.addressof(cam), C.addressof(tprm),
                                                C.c_float(LINE["desc_th"]), C.byref(fo_dev)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_dev()
    ex.ctx.sync()
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(2):  # keep the GPU under load while nvidia-smi starts sampling
        step_dev()
    barrier()
    l0 = ex.ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    ex.ctx.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ex.ctx.launch_count() - l0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_kp = int(d_n.sum().item())
    n_match = int(d_nm.sum().item())
    n_lines, n_lmatch = int(d_nl.sum().item()), int(d_lnm.sum().item())
    value = world * F_own * args.steps / (ms * 1e-3)

    # ---- per-stage pass (same steps, events between stages) -> roofline ----------------------
    ex.ctx.profile(True)
    ex.ctx.profile_read()
    for _ in range(args.steps):
        step_dev()
    alg = algorithmic_bytes(W, H, ORB["nlevels"], ORB["scale"], ORB["nfeatures"])
    alg["stereo_queries"] += W * H * 4                # + the colour conversion: RGB read, gray written
    peak, peak_src = load_peak()
    stages = stage_table(ex.ctx, alg, F, args.steps, peak)
    ex.ctx.profile(False)
    roofline = roofline_of(stages, F, peak, peak_src)

    # ---- end to end through the host-pointer C-ABI ---------------------------------------------
    # pinned host inputs (every batch of the loop is the same one, so one set serves); all results come back to the host
    pin = dict(pin_memory=True)
    h_rgb = torch.empty((F, H, W, 3), dtype=torch.uint8, **pin)
    h_depth = torch.empty((F, H, W), dtype=torch.int16, **pin)
    h_T = torch.empty((F, 12), dtype=torch.float32, **pin)
    h_rgb.copy_(d_rgb)
    h_depth.copy_(d_depth)
    h_T.copy_(d_T)
    torch.cuda.synchronize()
    h_kps = torch.empty((F, cap, 28), dtype=torch.uint8, **pin)
    h_desc = torch.empty((F, cap, 32), dtype=torch.uint8, **pin)
    h_n = torch.empty(F, dtype=torch.int32, **pin)
    h_nm = torch.empty(F, dtype=torch.int32, **pin)
    h_ur = torch.empty((F, cap), dtype=torch.float32, **pin)
    h_z = torch.empty((F, cap), dtype=torch.float32, **pin)
    h_assign = torch.empty((F, cap), dtype=torch.int32, **pin)
    h_kl = torch.empty((F, lcap, 68), dtype=torch.uint8, **pin)
    h_ld = torch.empty((F, lcap, 32), dtype=torch.uint8, **pin)
    h_eq = torch.empty((F, lcap, 3), dtype=torch.float64, **pin)
    h_nl = torch.empty(F, dtype=torch.int32, **pin)
    h_la = torch.empty((F, lcap), dtype=torch.int32, **pin)
    h_lnm = torch.empty(F, dtype=torch.int32, **pin)
    fo_host = FrontendOut(h_kps.data_ptr(), h_desc.data_ptr(), h_n.data_ptr(), h_ur.data_ptr(), h_z.data_ptr(),
                          h_assign.data_ptr(), h_nm.data_ptr(), cap, lcap, h_kl.data_ptr(), h_ld.data_ptr(),
                          h_eq.data_ptr(), h_nl.data_ptr(), h_la.data_ptr(), h_lnm.data_ptr())

    def begin(k):
        ex.ctx.check(L.psl_track_rgbd_batch_begin(ex.ctx.handle, h_rgb.data_ptr(), 3, 1, h_depth.data_ptr(), F, W, H,
                                                  h_T.data_ptr()))

    def end():
        ex.ctx.check(L.psl_track_rgbd_batch_end(ex.ctx.handle, C.addressof(cam), C.addressof(tprm),
                                                C.c_float(LINE["desc_th"]), C.byref(fo_host)))

    begin(0)
    end()                       # warm-up (allocates the staging sets)
    # Steady state of the double-buffered feed: batch 0 is already on its way when the clock starts, and every timed
    # iteration uploads one whole batch (k+1) while it computes and returns another (k); the batch left in flight at
    # the end is finished outside the timed region.  K uploads, K computes and K result downloads are timed.
    e2e_steps = max(3, args.steps // 2)
    begin(0)
    barrier()                   # (the device is idle and batch 0 resident: nothing of it is left to overlap with)
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        begin(k + 1)            # batch k+1 goes up while batch k is computed
        end()
    dt = time.perf_counter() - t0
    end()
    torch.cuda.synchronize()
    # the upload alone (what bounds the end-to-end number when the host cannot feed all GPUs at full PCIe speed)
    barrier()
    t1 = time.perf_counter()
    for k in range(2):
        begin(k)
    torch.cuda.synchronize()
    dt_up = time.perf_counter() - t1
    for k in range(2):
        end()
    h2d_bytes = F * (W * H * 5 + 48)
    if world > 1:
        t = torch.tensor([dt, dt_up], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_up = float(t[0].item()), float(t[1].item())
    e2e = {"value": world * F_own * e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": F * (cap * (28 + 32 + 12) + lcap * (68 + 32 + 24 + 4) + 16), "steps": e2e_steps,
           "input": "pinned host RGB u8 [F,H,W,3] + depth u16 [F,H,W] + poses; psl_track_rgbd_batch_begin / _end, "
                    "upload of batch k+1 overlapped with the kernels of batch k",
           "h2d_only_gbs_per_gpu": 2 * h2d_bytes / dt_up / 1e9,
           "h2d_only_frames_per_s": world * 2 * F_own / dt_up,
           "results_equal_device_path": int(h_n.sum().item()) == n_kp and int(h_nm.sum().item()) == n_match and
           int(h_nl.sum().item()) == n_lines and int(h_lnm.sum().item()) == n_lmatch}

    # ---- the rest of the frame on the same device arrays (not part of BASELINE's metric; reported beside it) ---------
    extended = None
    if rank == 0 and world == 1:
        try:
            extended = extended_chain(ex, L, step_dev, stream, args.steps, F, W, H, cap, lcap, cam, d_depth, d_T, d_kps, d_ur,
                                      d_z, d_assign, d_n, d_kl, d_nl, ms / args.steps)
        except Exception as e:  # the headline numbers above do not depend on it
            extended = {"error": f"{type(e).__name__}: {e}"[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nb = min(D, 24)
        cpu = cpu_baseline(base_rgb[:nb].cpu().numpy(), base_d[:nb].cpu().numpy().view(np.uint16), base_T[:nb])

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
               "config": workload_config(F_own, D), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": roofline, "stages": stages, "cpu_baseline": cpu, "extended_chain": extended,
               "keypoints_per_frame": n_kp / F, "matches_per_frame": n_match / max(F - 1, 1),
               "lines_per_frame": n_lines / F, "line_matches_per_frame": n_lmatch / max(F - 1, 1)}
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
