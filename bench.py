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
    }


def synth_frames(n_distinct, seed=4):
    from psl_slam_b200 import synth
    gray, depth, T = synth.sequence(seed, n_distinct, W, H)
    return gray


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


def cpu_baseline(frames_u8: np.ndarray, budget_s: float = 12.0):
    """The oracle (CPU port of the reference path) on a bounded sample, 1 thread and all threads."""
    from oracle import orc
    orc.build()
    p = orc.params(ORB["nfeatures"], ORB["scale"], ORB["nlevels"], ORB["ini"], ORB["mn"])
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    orc.orb_extract_batch_mt(frames_u8[:2], p, 1)
    t1 = (time.perf_counter() - t0) / 2
    n1 = max(2, min(len(frames_u8), int(budget_s * 0.3 / t1)))
    t0 = time.perf_counter()
    orc.orb_extract_batch_mt(frames_u8[:n1], p, 1)
    fps1 = n1 / (time.perf_counter() - t0)
    reps = max(1, int(budget_s * 0.7 * fps1 * cores * 0.7 / len(frames_u8)))
    sample = np.concatenate([frames_u8] * reps) if reps > 1 else frames_u8
    t0 = time.perf_counter()
    orc.orb_extract_batch_mt(sample, p, cores)
    fpsN = len(sample) / (time.perf_counter() - t0)
    return {"value": fpsN, "unit": "frames/s", "cores": cores, "kind": "port",
            "value_1_thread": fps1,
            "sample": f"{len(sample)} frames ORB extract on {cores} threads (one frame per task); "
                      f"{n1} frames on 1 thread"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (oracle port)."""
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    frames = synth_frames(8)
    p = orc.params(ORB["nfeatures"], ORB["scale"], ORB["nlevels"], ORB["ini"], ORB["mn"])
    cores = os.cpu_count() or 1
    per_step = max(cores, 2 * cores)
    batch = np.concatenate([frames] * ((per_step + len(frames) - 1) // len(frames)))[:per_step]
    for _ in range(args.warmup):
        orc.orb_extract_batch_mt(batch, p, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.orb_extract_batch_mt(batch, p, cores)
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


def workload_config(frames_per_gpu):
    return {"workload": "cfg1-batched: ORBextractor on synthetic 640x480 frames, TUM1.yaml settings "
                        "(1000 features, 8 levels, 1.2, FAST 20/7); matching and line stages join as they land",
            "frames_per_step_per_gpu": frames_per_gpu, "width": W, "height": H,
            "l2_policy": "inputs larger than L2 (frames_per_step x 307 KB >> 126 MB), no flush",
            "parallelism": "frames sharded across GPUs, no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=2048, help="frames per step per GPU")
    ap.add_argument("--chunk", type=int, default=0)
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from psl_slam_b200 import ORBextractor

    F = args.frames
    ex = ORBextractor(ORB["nfeatures"], ORB["scale"], ORB["nlevels"], ORB["ini"], ORB["mn"], device=local,
                      max_width=W, max_height=H, chunk_frames=args.chunk)
    cap = ex.cap
    base = synth_frames(16, seed=4 + rank)                       # distinct synthetic frames per rank
    d_base = torch.from_numpy(base).cuda()
    idx = torch.arange(F, device="cuda") % d_base.shape[0]
    d_gray = d_base[idx].contiguous()                            # [F,H,W] u8 resident in HBM
    d_kps = torch.empty((F, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.empty((F, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ex.ctx.stream(), device=local)

    def step_dev():
        ex.extract_batch_dev(d_gray.data_ptr(), F, W, H, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(),
                             d_n.data_ptr(), cap)

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
    value = world * F * args.steps / (ms * 1e-3)

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
    roofline = {"bound": "hbm", "kernel": dom["stage"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_launch": dom["alg_bytes_per_frame"] * F / nl,
                "avg_launch_ms": dom["ms_per_step"] / nl,
                "note": "achieved = algorithmic bytes of the stage x frames / CUDA-event time of the stage"}

    # ---- end to end through the host-pointer C-ABI ---------------------------------------------
    import ctypes as C

    from psl_slam_b200 import _lib
    h_gray = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
    h_gray.copy_(d_gray.cpu())
    h_kps = torch.empty((F, cap, 28), dtype=torch.uint8, pin_memory=True)
    h_desc = torch.empty((F, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_n = torch.empty(F, dtype=torch.int32, pin_memory=True)

    def step_host():
        ex.ctx.check(_lib.lib().psl_orb_extract_batch(ex.ctx.handle, h_gray.data_ptr(), F, W, H, W, W * H,
                                                      h_kps.data_ptr(), h_desc.data_ptr(), cap, h_n.data_ptr()))

    for _ in range(2):
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
    e2e = {"value": world * F * e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": F * W * H,
           "d2h_bytes_per_step": F * (cap * 60 + 4), "steps": e2e_steps,
           "keypoints_checked": int(h_n.sum().item()) == n_kp}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(base)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
               "config": workload_config(F), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": roofline, "stages": stages, "cpu_baseline": cpu,
               "keypoints_per_frame": n_kp / F}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
