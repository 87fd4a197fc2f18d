#!/usr/bin/env python
"""Turn what tools/profile_round.sh brought back in gpurun_out/ into the tracked files under profiles/.

    python tools/profiles_from_round.py r02 r02
      gpurun_out/bench_<tag>.json, _reference.json, _cfgN.json      -> profiles/<out>_bench_cfg4.json, _reference_arm.json, _bench_cfgN.json
      gpurun_out/launches_<tag>.csv                                  -> profiles/<out>_launches_cfg4.md + .csv.gz
      gpurun_out/<tag>_kernels_{core,main,stream}_ncu.md (+ raw csv) -> profiles/<out>_kernels_ncu.md, traffic.json
      gpurun_out/<tag>_lsd_core_source.csv.gz                        -> profiles/<out>_lsd_core_phases.md
"""
import csv
import gzip
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

tag, out = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, f"bench_{tag}.json"), os.path.join(P, f"{out}_bench_cfg4.json"))
shutil.copy(os.path.join(G, f"bench_{tag}_reference.json"), os.path.join(P, f"{out}_bench_reference_arm.json"))
for c in ("cfg1", "cfg2", "cfg3", "cfg5"):
    src = os.path.join(G, f"bench_{tag}_{c}.json")
    if os.path.exists(src) and os.path.getsize(src):
        shutil.copy(src, os.path.join(P, f"{out}_bench_{c}.json"))
ncu_summary.launches(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{out}_launches_cfg4.md"))
with open(os.path.join(G, f"launches_{tag}.csv"), "rb") as f, gzip.open(os.path.join(P, f"{out}_launches_cfg4.csv.gz"), "wb") as g:
    g.write(f.read())
with open(os.path.join(P, f"{out}_kernels_ncu.md"), "w") as f:
    f.write("One `ncu --set full --clock-control none` capture per kernel (first launch of each), summarised on the GPU box by\n"
            "`tools/profile_round.sh`.  `lsd_core_kernel` is captured at the bench's own launch size (4096 frames per launch); the ORB\n"
            "kernels run 512 frames per launch in the bench as well; the other line kernels are captured at 512 frames per launch.\n\n")
    for part in ("core", "main", "stream"):
        f.write(open(os.path.join(G, f"{tag}_kernels_{part}_ncu.md")).read() + "\n")

# DRAM traffic per frame of each stage's dominant kernel (bench.py's roofline.traffic)
per_kernel = {}
for part, frames in (("core", 4096), ("main", 512), ("stream", 512)):
    rows = list(csv.reader(open(os.path.join(G, f"{tag}_kernels_{part}_raw.csv"))))
    hdr, units = rows[0], rows[1]
    ni, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ni]).split("::")[-1]
        name = re.sub(r"^void ", "", name)
        if name not in per_kernel:
            per_kernel[name] = ((float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]) / frames, frames)
stage_kernel = {"lsd_grow": "lsd_core_kernel", "line_merge": "line_scan_kernel", "lbd": "lbd_kernel",
                "fast": "fast_rows_kernel<1>", "describe": "describe_kernel", "blur": "gauss7_kernel<1>",
                "pyramid": "resize_words_kernel", "candidates": "proj_candidates_kernel", "lsd_order": "lsd_seed_order_kernel"}
traffic = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per frame of the stage's dominant kernel, from one `ncu --set full` "
                    "capture (bytes of the launch / frames of the launch); bench.py multiplies by the frames one launch processes. "
                    "For the pyramid and blur stages the captured launch is the largest level only."}
for stage, k in stage_kernel.items():
    if k in per_kernel:
        traffic[stage] = {"kernel": k, "dram_bytes_per_frame": int(per_kernel[k][0]),
                          "source": f"profiles/{out}_kernels_ncu.md ({per_kernel[k][1]} frames per launch)"}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "lsd_phase_table.py"),
                       os.path.join(G, f"{tag}_lsd_core_source.csv.gz"), "4096", os.path.join(P, f"{out}_lsd_core_phases.md")])
