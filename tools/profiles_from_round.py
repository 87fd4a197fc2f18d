#!/usr/bin/env python
"""Turn what tools/profile_round.sh brought back in gpurun_out/ into the tracked files under profiles/.

    python tools/profiles_from_round.py r01b r01
      gpurun_out/bench_<tag>.json, bench_<tag>_reference.json      -> profiles/<out>_bench_combined.json, _reference_arm.json
      gpurun_out/launches_<tag>.csv                                -> profiles/<out>_launches_combined.md + .csv.gz
      gpurun_out/<tag>_kernels_{main,stream}_ncu.md (+ raw csv)    -> profiles/<out>_kernels_ncu.md, traffic.json
"""
import csv
import gzip
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

tag, out = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, f"bench_{tag}.json"), os.path.join(P, f"{out}_bench_combined.json"))
shutil.copy(os.path.join(G, f"bench_{tag}_reference.json"), os.path.join(P, f"{out}_bench_reference_arm.json"))
ncu_summary.launches(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{out}_launches_combined.md"))
with open(os.path.join(G, f"launches_{tag}.csv"), "rb") as f, gzip.open(os.path.join(P, f"{out}_launches_combined.csv.gz"), "wb") as g:
    g.write(f.read())
with open(os.path.join(P, f"{out}_kernels_ncu.md"), "w") as f:
    f.write("One `ncu --set full --clock-control none` capture per kernel (first launch of each; `bench.py --frames 512`,\n"
            "so one launch covers 512 frames), summarised on the GPU box by `tools/profile_round.sh`.\n\n")
    for part in ("main", "stream"):
        f.write(open(os.path.join(G, f"{tag}_kernels_{part}_ncu.md")).read() + "\n")

# DRAM traffic per frame of each stage's dominant kernel (bench.py's roofline.traffic)
frames = 512
per_kernel = {}
for part in ("main", "stream"):
    rows = list(csv.reader(open(os.path.join(G, f"{tag}_kernels_{part}_raw.csv"))))
    hdr, units = rows[0], rows[1]
    ni, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ni]).split("::")[-1]
        name = re.sub(r"^void ", "", name)
        if name not in per_kernel:
            per_kernel[name] = (float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]) / frames
stage_kernel = {"lsd_grow": "lsd_core_kernel", "line_merge": "line_post_kernel", "lbd": "lbd_kernel",
                "fast": "fast_score_tiles_kernel<1>", "describe": "describe_kernel", "blur": "gauss7_kernel<1>",
                "pyramid": "resize_words_kernel", "candidates": "proj_candidates_kernel"}
traffic = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per frame of the stage's dominant kernel, from one `ncu --set full` "
                    "capture (bytes of the launch / frames of the launch); bench.py multiplies by the frames one launch processes. "
                    "For the pyramid and blur stages the captured launch is the largest level only."}
for stage, k in stage_kernel.items():
    if k in per_kernel:
        traffic[stage] = {"kernel": k, "dram_bytes_per_frame": int(per_kernel[k]),
                          "source": f"profiles/{out}_kernels_ncu.md ({frames} frames per launch)"}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
