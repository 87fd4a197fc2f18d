#!/usr/bin/env python
"""Stage timing of the CUDA line extractor on device-resident frames (development tool, not the bench)."""
import argparse
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--distinct", type=int, default=8)
    ap.add_argument("--lowtex", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import torch
    from psl_slam_b200 import LINEextractor, synth, KEYLINE_DTYPE
    idx = torch.from_numpy(np.arange(a.frames) % a.distinct).cuda()
    if a.lowtex:
        gray = np.stack([synth.make_lowtex(100 + i) for i in range(a.distinct)])
        d = torch.from_numpy(gray).cuda()[idx].contiguous()
    elif a.distinct > 32:   # the bench's frames: rendered on the device
        import bench
        rgb, _, _ = bench.render_sequence_cuda(4, a.distinct, 640, 480, torch.device("cuda"))
        d = bench.gray_cuda(rgb)[idx].contiguous()
    else:
        gray, _, _ = synth.sequence(4, a.distinct)
        d = torch.from_numpy(gray).cuda()[idx].contiguous()
    ex = LINEextractor(chunk_frames=a.chunk)
    cap = ex.cap
    kl = torch.zeros(a.frames * cap * 68, dtype=torch.uint8, device="cuda")
    ld = torch.zeros(a.frames * cap * 32, dtype=torch.uint8, device="cuda")
    eq = torch.zeros(a.frames * cap * 3, dtype=torch.float64, device="cuda")
    n = torch.zeros(a.frames, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.ctx.profile(True)
    for r in range(a.reps + 1):
        ex.extract_batch_dev(d.data_ptr(), a.frames, 640, 480, 640, 640 * 480, kl.data_ptr(), ld.data_ptr(), eq.data_ptr(),
                             None, n.data_ptr())
        ex.ctx.sync()
        ms, ln = ex.ctx.profile_read()
        tot = float(ms[10:15].sum())
        print(f"rep {r}: total {tot:.1f} ms  {a.frames / tot * 1e3:.0f} frames/s  prologue={ms[10]:.2f} order={ms[11]:.2f} "
              f"core={ms[12]:.2f} post={ms[13]:.2f} lbd={ms[14]:.2f}  lines/frame={n.float().mean().item():.1f}", flush=True)
        _stats()


def _stats():
    import ctypes as C
    from psl_slam_b200 import _lib
    L = _lib.lib()
    if hasattr(L, "psl_lsd_stats"):
        z = (C.c_ulonglong * 16)()
        L.psl_lsd_stats(z)
        if hasattr(L, "psl_post_stats"):
            z2 = (C.c_ulonglong * 16)()
            L.psl_post_stats(z2)
            print("post stats cyc [sort, scan, cluster, fold, total]:", list(z2)[:5])
        print("lsd stats [0 steps, 1 region points visited, 2 regions>=min, 3 their pixels, 4 seeds unused at ballot, 5 of those used at their turn, 6 singleton regions, 7 sequential steps, 9 pixels of small regions | cyc: 10 grow, 11 rect, 12 refine, 13 regions grown, 14 total, 15 slowest frame]:", list(z)[:16])


if __name__ == "__main__":
    main()
