"""bench.py --config cfg1|cfg2|cfg3|cfg5: the BASELINE.json configurations other than the one the metric is quoted on.

cfg1  ORBextractor alone, ONE 640x480 gray frame per call (TUM1.yaml settings)              -> latency of operator()
cfg2  point front end of one tracking step: ORB on two consecutive frames + stereo + SearchByProjection(Cur, Last)
cfg3  LINEextractor alone, ONE low-texture 640x480 frame per call (LSD + merge + top-200 + LBD + line equations)
cfg5  high-resolution stress: 1920x1080, 4000 features, 12 levels, combined ORB + line front end, batched per GPU

Same JSON contract as cfg4: `value` = frames/s with the input resident in HBM (device-pointer C-ABI, CUDA events),
`e2e` = the host-pointer call a drop-in user makes (wall clock, copies inside), `stages` / `roofline` from the per-stage
event pass, `cpu_baseline` = the oracle port on this host (one thread: how the reference runs a frame).  For the
single-frame configs the number that matters is `latency_ms`.
"""
from __future__ import annotations

import os
import time

import numpy as np


def _time_dev(fn, ctx, stream, iters, torch):
    for _ in range(3):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    ctx.sync()
    return e0.elapsed_time(e1) / iters


def _time_host(fn, iters):
    for _ in range(2):
        fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters * 1e3


def _cpu(fn, budget_s=8.0):
    fn()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < budget_s and n < 200:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n * 1e3, n


def run_small(args, rank, world, local):
    import torch
    import torch.distributed as dist

    import bench as B
    from psl_slam_b200 import (Context, LINEextractor, ORBextractor, default_config, make_camera, make_track_params, synth,
                               track_frontend_batch_dev, track_orb_batch, track_orb_batch_dev)
    from psl_slam_b200._lib import FrontendOut

    cfgname = args.config
    dev = torch.device("cuda", local)
    hd = cfgname == "cfg5"
    W, H = (1920, 1080) if hd else (640, 480)
    orb = dict(nfeatures=4000, scale=1.2, nlevels=12, ini=20, mn=7) if hd else B.ORB
    F = (args.frames or 512) if hd else (2 if cfgname == "cfg2" else 1)
    cfg = default_config()
    cfg.device, cfg.max_width, cfg.max_height, cfg.max_batch = local, W, H, F + 1
    cfg.orb_nfeatures, cfg.orb_scale_factor, cfg.orb_nlevels = orb["nfeatures"], orb["scale"], orb["nlevels"]
    cfg.orb_ini_th_fast, cfg.orb_min_th_fast = orb["ini"], orb["mn"]
    cfg.chunk_frames = args.chunk or min(F, 128 if hd else 512)
    cfg.line_nfeatures = B.LINE["nfeatures"]
    cfg.line_chunk_frames = args.line_chunk or F
    if hd:
        cfg.line_max_raw = 16384
    ctx = Context(cfg)
    ex = ORBextractor(ctx=ctx)
    lx = LINEextractor(ctx=ctx)
    cap, lcap = ex.cap, B.LINE["nfeatures"]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)
    K = synth.ICL
    sx, sy = W / 640.0, H / 480.0
    cam = make_camera(K["fx"] * sx, K["fy"] * sy, (K["cx"] + 0.5) * sx - 0.5, (K["cy"] + 0.5) * sy - 0.5, K["bf"] * sx,
                      K["depth_factor"])
    cam6 = np.array([K["fx"] * sx, K["fy"] * sy, (K["cx"] + 0.5) * sx - 0.5, (K["cy"] + 0.5) * sy - 0.5, K["bf"] * sx,
                     np.float32(1.0) / np.float32(K["depth_factor"])], np.float32)
    tprm = make_track_params(B.TRACK["th"], B.TRACK["nn_ratio"], B.TRACK["ori"])
    seed = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg5": 5}[cfgname]
    iters = max(args.steps, 1) * (1 if hd else 20)
    peak, peak_src = B.load_peak()
    alg = B.algorithmic_bytes(W, H, orb["nlevels"], orb["scale"], orb["nfeatures"])
    from oracle import orc
    B.native_oracle()
    p = orc.params(orb["nfeatures"], orb["scale"], orb["nlevels"], orb["ini"], orb["mn"])
    extra = {}

    if cfgname == "cfg3":
        frames = np.stack([synth.make_lowtex(300 + i) for i in range(2)])
        d_gray = torch.from_numpy(frames).to(dev)
    else:
        D = 2 if not hd else min(F, 64)
        rgb, depth, T12 = B.render_sequence_cuda(seed, D, W, H, dev, poster_size=4096 if hd else 2048)
        gray_all = B.gray_cuda(rgb)
        if hd:
            idx = torch.from_numpy(B.ping_pong(D, F)).to(dev)
            d_gray, d_depth = gray_all[idx].contiguous(), depth[idx].contiguous()
            d_T = torch.from_numpy(T12).to(dev)[idx].contiguous()
        else:
            d_gray, d_depth, d_T = gray_all[:F].contiguous(), depth[:F].contiguous(), torch.from_numpy(T12[:F]).to(dev)
        frames = d_gray[: min(F, 8)].cpu().numpy()
    d_kps = torch.empty((F, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((F, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(F, dtype=torch.int32, device=dev)
    d_kl = torch.empty((F, lcap, 68), dtype=torch.uint8, device=dev)
    d_ld = torch.empty((F, lcap, 32), dtype=torch.uint8, device=dev)
    d_eq = torch.empty((F, lcap, 3), dtype=torch.float64, device=dev)
    d_nl = torch.zeros(F, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    if cfgname == "cfg1":
        what = ("cfg1: ORBextractor::operator() on one synthetic 640x480 gray frame (TUM1.yaml: 1000 features, 8 levels, 1.2, "
                "FAST 20/7)")
        step_dev = lambda: ex.extract_batch_dev(d_gray.data_ptr(), 1, W, H, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(),
                                                d_n.data_ptr())
        step_host = lambda: ex(frames[0])
        cpu_fn = lambda: orc.orb_extract(frames[0], p)
        check = lambda: int(d_n.sum().item()) == len(orc.orb_extract(frames[0], p)[0])
        used = ("pyramid", "fast", "octree", "blur", "describe")
    elif cfgname == "cfg2":
        what = ("cfg2: point front end of one tracking step — ORBextractor on two consecutive synthetic 640x480 RGB-D frames + "
                "ComputeStereoFromRGBD + SearchByProjection(Cur, Last, th=15) (ICL intrinsics)")
        d_ur = torch.empty((F, cap), dtype=torch.float32, device=dev)
        d_z = torch.empty((F, cap), dtype=torch.float32, device=dev)
        d_as = torch.empty((F, cap), dtype=torch.int32, device=dev)
        d_nm = torch.zeros(F, dtype=torch.int32, device=dev)
        step_dev = lambda: track_orb_batch_dev(ex, d_gray.data_ptr(), d_depth.data_ptr(), F, W, H, d_T.data_ptr(), cam, tprm,
                                               d_kps.data_ptr(), d_desc.data_ptr(), d_n.data_ptr(), d_ur.data_ptr(),
                                               d_z.data_ptr(), d_as.data_ptr(), d_nm.data_ptr(), cap)
        h_depth = d_depth.cpu().numpy().view(np.uint16)
        T44 = np.tile(np.eye(4, dtype=np.float32), (F, 1, 1))
        T44[:, :3, :4] = T12[:F].reshape(F, 3, 4)
        step_host = lambda: track_orb_batch(ex, frames, h_depth, T44, cam, tprm)
        cpu_fn = lambda: orc.track_batch_mt(frames, h_depth, T12[:F], cam6, p, B.TRACK["th"], B.TRACK["nn_ratio"],
                                            B.TRACK["ori"], 1)
        def check():
            n, nm = orc.track_batch_mt(frames, h_depth, T12[:F], cam6, p, B.TRACK["th"], B.TRACK["nn_ratio"], B.TRACK["ori"], 1)
            extra["matches"] = int(d_nm[1].item())
            return np.array_equal(d_n.cpu().numpy(), n) and np.array_equal(d_nm.cpu().numpy(), nm)
        used = ("pyramid", "fast", "octree", "blur", "describe", "stereo_queries", "grid", "candidates", "resolve")
    elif cfgname == "cfg3":
        what = ("cfg3: LINEextractor::operator() on one synthetic low-texture 640x480 frame — LSD, long-line merge, top 200, "
                "LBD descriptors, 2-D line equations")
        step_dev = lambda: lx.extract_batch_dev(d_gray.data_ptr(), 1, W, H, W, W * H, d_kl.data_ptr(), d_ld.data_ptr(),
                                                d_eq.data_ptr(), None, d_nl.data_ptr())
        step_host = lambda: lx(frames[0])
        cpu_fn = lambda: orc.line_extract(frames[0], lcap)
        def check():
            extra["lines"] = int(d_nl[0].item())
            return int(d_nl[0].item()) == len(orc.line_extract(frames[0], lcap)[0])
        used = ("lsd_prologue", "lsd_order", "lsd_grow", "line_merge", "lbd")
    else:
        what = (f"cfg5: high-resolution stress — combined ORB + line front end on {F} synthetic 1920x1080 RGB-D frames per GPU "
                "(4000 features, 12 levels, 1.2, FAST 20/7; LSD + merge + top 200 + LBD; stereo, SearchByProjection, "
                "SearchByGeomNApearance against the previous frame)")
        d_ur = torch.empty((F, cap), dtype=torch.float32, device=dev)
        d_z = torch.empty((F, cap), dtype=torch.float32, device=dev)
        d_as = torch.empty((F, cap), dtype=torch.int32, device=dev)
        d_nm = torch.zeros(F, dtype=torch.int32, device=dev)
        d_la = torch.empty((F, lcap), dtype=torch.int32, device=dev)
        d_lnm = torch.zeros(F, dtype=torch.int32, device=dev)
        fo = FrontendOut(d_kps.data_ptr(), d_desc.data_ptr(), d_n.data_ptr(), d_ur.data_ptr(), d_z.data_ptr(),
                         d_as.data_ptr(), d_nm.data_ptr(), cap, lcap, d_kl.data_ptr(), d_ld.data_ptr(), d_eq.data_ptr(),
                         d_nl.data_ptr(), d_la.data_ptr(), d_lnm.data_ptr())
        step_dev = lambda: track_frontend_batch_dev(ex, d_gray.data_ptr(), d_depth.data_ptr(), F, W, H, d_T.data_ptr(), cam,
                                                    tprm, B.LINE["desc_th"], fo)
        step_host = None
        h_depth = d_depth[:3].cpu().numpy().view(np.uint16)
        T12h = d_T[:3].cpu().numpy()
        cpu_fn = lambda: orc.frontend_batch_mt(frames[:3], h_depth, T12h, cam6, p, B.TRACK["th"], B.TRACK["nn_ratio"],
                                               B.TRACK["ori"], lcap, B.LINE["desc_th"], 1)
        def check():
            n, nm, nl, lnm = cpu_fn()
            extra.update(keypoints_per_frame=float(d_n.float().mean().item()), lines_per_frame=float(d_nl.float().mean().item()),
                         matches_per_frame=float(d_nm[1:].float().mean().item()))
            return (np.array_equal(d_n[:3].cpu().numpy(), n) and np.array_equal(d_nm[:3].cpu().numpy(), nm) and
                    np.array_equal(d_nl[:3].cpu().numpy(), nl) and np.array_equal(d_lnm[:3].cpu().numpy(), lnm))
        used = tuple(ctx.STAGES)

    sampler = B.ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    ms = _time_dev(step_dev, ctx, stream, iters, torch)
    launches = (ctx.launch_count() - l0) * iters // (iters + 3)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ok = bool(check())
    ctx.profile(True)
    ctx.profile_read()
    for _ in range(iters):
        step_dev()
    stages = [s for s in B.stage_table(ctx, alg, F, iters, peak) if s["stage"] in used]
    ctx.profile(False)
    roofline = B.roofline_of(stages, F, peak, peak_src)
    e2e = None
    if step_host is not None:
        hms = _time_host(step_host, max(iters // 4, 5))
        e2e = {"value": world * F / (hms * 1e-3), "unit": "frames/s", "latency_ms": hms,
               "h2d_bytes_per_step": F * W * H * (3 if cfgname == "cfg2" else 1),
               "d2h_bytes_per_step": F * (cap * 60 if cfgname != "cfg3" else lcap * 124),
               "input": "host numpy frames through the host-pointer C-ABI call, one synchronous call per step"}
    else:
        e2e = {"value": None, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "input": "not measured for cfg5: the host-pointer path is the one cfg4 measures"}
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cms, n = _cpu(cpu_fn, 20.0 if hd else 8.0)
        nf = 3 if hd else F
        cpu = {"value": nf / (cms * 1e-3), "unit": "frames/s", "cores": 1, "kind": "port", "latency_ms_per_frame": cms / nf,
               "build": os.path.basename(orc.lib().path),
               "sample": f"{n} repeats of the same {nf} frame(s) on one thread (how the reference runs a frame, Frame.cc:179-180)"}
    if rank != 0:
        return None
    out = {"metric": B.METRIC.replace("@640x480", f"@{W}x{H}"), "value": world * F / (ms * 1e-3), "unit": "frames/s",
           "n_gpus": world, "steps": iters, "warmup": 3, "ms_per_step": ms, "latency_ms": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": what, "frames_per_step_per_gpu": F, "width": W, "height": H,
                      "l2_policy": "inputs larger than L2, no flush" if hd else
                      "single-frame latency: the frame (0.3 MB) is L2-resident by construction, as in online tracking",
                      "parallelism": "independent frames per GPU, no collective" if hd else "one frame per call"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "stages": stages,
           "cpu_baseline": cpu, "results_equal_oracle": ok}
    out.update(extra)
    return out
