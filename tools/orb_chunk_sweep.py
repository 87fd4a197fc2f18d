import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from psl_slam_b200 import ORBextractor, synth
F = 2048
gray, _, _ = synth.sequence(4, 16)
d = torch.from_numpy(gray).cuda()[torch.arange(F).cuda() % 16].contiguous()
for chunk in ([int(a) for a in sys.argv[1:]] or (32, 64, 128, 256, 512, 1024)):
    ex = ORBextractor(chunk_frames=chunk)
    cap = ex.cap
    kps = torch.empty((F, cap, 28), dtype=torch.uint8, device='cuda'); desc = torch.empty((F, cap, 32), dtype=torch.uint8, device='cuda'); n = torch.zeros(F, dtype=torch.int32, device='cuda')
    ex.ctx.profile(True)
    for r in range(3):
        ex.extract_batch_dev(d.data_ptr(), F, 640, 480, 640, 640*480, kps.data_ptr(), desc.data_ptr(), n.data_ptr())
        ex.ctx.sync()
        ms, ln = ex.ctx.profile_read()
    print(chunk, "total %.2f" % ms[:5].sum(), " ".join(f"{nm}={v:.2f}" for nm, v in zip(ex.ctx.STAGES[:5], ms[:5])), flush=True)
    del ex
