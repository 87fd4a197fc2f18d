#!/usr/bin/env python
"""Benchmark of the PSL-SLAM feature front end on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU path on the host cores

One *step* = one pass of the hot path over a batch of synthetic 640x480 frames (per GPU).
`value`   : whole-job frames/s with the frames already resident in HBM (device-pointer C-ABI).
`e2e`     : the same through the host-pointer C-ABI call (pinned host frames in, keypoints +
            descriptors out; H2D and D2H inside the timed region).
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


def synth_frames(n_distinct, seed=4):
    """n consecutive synthetic RGB-D frames (gray u8, depth u16) + world->camera poses [n,12] f32."""
    from psl_slam_b200 import synth
    gray, depth, T = synth.sequence(seed, n_distinct, W, H)
    return gray, depth, np.ascontiguousarray(T.astype(np.float32)[:, :3, :4].reshape(n_distinct, 12))


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


def cpu_run(gray, depth, T12, nthreads):
    from oracle import orc
    p = orc.params(ORB["nfeatures"], ORB["scale"], ORB["nlevels"], ORB["ini"], ORB["mn"])
    return orc.frontend_batch_mt(gray, depth, T12, cam6(), p, TRACK["th"], TRACK["nn_ratio"], TRACK["ori"],
                                 LINE["nfeatures"], LINE["desc_th"], nthreads)


def cpu_baseline(gray, depth, T12, budget_s: float = 14.0):
    """The oracle (CPU port of the reference path: ORB + LSD/LBD extraction + stereo + both frame-to-frame
    matchers) on a bounded sample of the same workload, 1 thread (how the reference runs, Frame.cc:179-180)
    and all threads."""
    from oracle import orc
    orc.build()
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
    return {"value": fpsN, "unit": "frames/s", "cores": cores, "kind": "port", "value_1_thread": fps1,
            "sample": f"{total} frames (ORB + LSD/LBD extract + stereo + SearchByProjection + SearchByGeomNApearance vs "
                      f"previous frame) on {cores} threads, one frame per task; {n1} frames on 1 thread"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (oracle port)."""
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    gray, depth, T12 = synth_frames(8)
    cores = os.cpu_count() or 1
    per_step = 2 * cores
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
           "config": workload_config(per_step),
           "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                            "sample": f"{per_step} frames per step, one frame per task on {cores} threads"},
           "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


METRIC = "frames/s ORB+LSD/LBD extract+match @640x480"
TRACK = dict(th=15.0, nn_ratio=0.9, ori=True)  # Tracking.cc:1166,1189
LINE = dict(nfeatures=200, desc_th=0.95)        # TUM1.yaml:60-63, Tracking.cc:1182


def workload_config(frames_per_gpu):
    return {"workload": "cfg4: batched combined front end per frame — ORBextractor (TUM1.yaml: 1000 features, 8 levels, "
                        "1.2, FAST 20/7) + LINEextractor (LSD, long-line merge, top 200, LBD) + ComputeStereoFromRGBD + "
                        "SearchByProjection(Cur, Last, th=15) + SearchByGeomNApearance(Cur, Last, 0.95) between "
                        "consecutive synthetic textured 640x480 RGB-D frames (ICL intrinsics)",
            "frames_per_step_per_gpu": frames_per_gpu, "width": W, "height": H,
            "l2_policy": "inputs larger than L2 (frames_per_step x 307 KB >> 126 MB), no flush",
            "parallelism": "one sequence of frames_per_step_per_gpu x n_gpus frames in contiguous shards (one-frame "
                           "halo re-extracted per shard), no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per step per GPU")
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

    from psl_slam_b200 import Context, ORBextractor, default_config
    from psl_slam_b200._lib import FrontendOut

    F = args.frames
    cfg = default_config()
    cfg.device, cfg.max_width, cfg.max_height, cfg.max_batch = local, W, H, F + 1
    cfg.orb_nfeatures, cfg.orb_scale_factor, cfg.orb_nlevels = ORB["nfeatures"], ORB["scale"], ORB["nlevels"]
    cfg.orb_ini_th_fast, cfg.orb_min_th_fast = ORB["ini"], ORB["mn"]
    cfg.chunk_frames, cfg.line_nfeatures = args.chunk, LINE["nfeatures"]
    cfg.line_chunk_frames = args.line_chunk or min(F + 1, 4100)
    ex = ORBextractor(ctx=Context(cfg))
    cap, lcap = ex.cap, LINE["nfeatures"]
    from psl_slam_b200 import make_camera, make_track_params, synth, track_frontend_batch_dev
    K = synth.ICL
    cam = make_camera(K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], K["depth_factor"])
    tprm = make_track_params(TRACK["th"], TRACK["nn_ratio"], TRACK["ori"])
    # cfg 4: one long sequence (F frames per GPU, F x N in total) in contiguous shards; ranks > 0 re-extract the
    # frame before their range (one-frame halo) so that every owned frame is matched against its predecessor
    from psl_slam_b200.shard import shard_with_halo
    base_g, base_d, base_T = synth_frames(16, seed=4)
    s0, s1, halo = shard_with_halo(F * world, rank, world)
    idx = torch.from_numpy(ping_pong(16, F * world)[s0:s1]).cuda()
    F_own, F = F, s1 - s0                                         # F now counts the halo frame too
    d_gray = torch.from_numpy(base_g).cuda()[idx].contiguous()    # [F,H,W] u8 resident in HBM
    d_depth = torch.from_numpy(base_d.view(np.int16)).cuda()[idx].contiguous()  # u16 bits
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

    def step_dev():
        track_frontend_batch_dev(ex, d_gray.data_ptr(), d_depth.data_ptr(), F, W, H, d_T.data_ptr(), cam, tprm,
                                 LINE["desc_th"], fo_dev)

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
    st_ms, st_launch = ex.ctx.profile_read()
    ex.ctx.profile(False)
    names = ex.ctx.STAGES
    alg = algorithmic_bytes(W, H, ORB["nlevels"], ORB["scale"], ORB["nfeatures"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tot_ms = float(sum(st_ms[: len(names)]))
    stages = []
    for i, nme in enumerate(names):
        t_ms = float(st_ms[i])
        gbs = alg[nme] * F * args.steps / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        stages.append({"stage": nme, "ms_per_step": t_ms / args.steps, "share": t_ms / tot_ms if tot_ms else 0,
                       "launches_per_step": int(st_launch[i]) // args.steps, "alg_bytes_per_frame": alg[nme],
                       "achieved_gbs": gbs, "frac": gbs / peak})
    dom = max(stages, key=lambda s: s["ms_per_step"])
    nl = max(dom["launches_per_step"], 1)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom["stage"])
        if tj and tj.get("dram_bytes_per_frame"):
            traffic = tj["dram_bytes_per_frame"] * F / nl
            traffic_src = f'{tj["kernel"]}: {tj["source"]}'
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom["stage"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "alg_bytes_per_launch": dom["alg_bytes_per_frame"] * F / nl,
                "avg_launch_ms": dom["ms_per_step"] / nl,
                "note": "achieved = algorithmic bytes of the stage x frames / CUDA-event time of the stage"}

    # ---- end to end through the host-pointer C-ABI ---------------------------------------------
    import ctypes as C

    from psl_slam_b200 import _lib
    pin = dict(pin_memory=True)
    h_gray = torch.empty((F, H, W), dtype=torch.uint8, **pin)
    h_gray.copy_(d_gray.cpu())
    h_depth = torch.empty((F, H, W), dtype=torch.int16, **pin)
    h_depth.copy_(d_depth.cpu())
    h_T = torch.empty((F, 12), dtype=torch.float32, **pin)
    h_T.copy_(d_T.cpu())
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

    def step_host():
        ex.ctx.check(_lib.lib().psl_track_frontend_batch(ex.ctx.handle, h_gray.data_ptr(), h_depth.data_ptr(), F, W, H,
                                                         h_T.data_ptr(), C.addressof(cam), C.addressof(tprm),
                                                         C.c_float(LINE["desc_th"]), C.byref(fo_host)))

    step_host()
    barrier()
    e2e_steps = max(2, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": world * F_own * e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": F * (W * H * 3 + 48),
           "d2h_bytes_per_step": F * (cap * (28 + 32 + 12) + lcap * (68 + 32 + 24 + 4) + 16), "steps": e2e_steps,
           "results_equal_device_path": int(h_n.sum().item()) == n_kp and int(h_nm.sum().item()) == n_match and
           int(h_nl.sum().item()) == n_lines and int(h_lnm.sum().item()) == n_lmatch}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(base_g, base_d, base_T)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
               "config": workload_config(F_own), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": roofline, "stages": stages, "cpu_baseline": cpu,
               "keypoints_per_frame": n_kp / F, "matches_per_frame": n_match / max(F - 1, 1),
               "lines_per_frame": n_lines / F, "line_matches_per_frame": n_lmatch / max(F - 1, 1)}
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
