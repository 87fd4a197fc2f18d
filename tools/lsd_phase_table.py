#!/usr/bin/env python
"""Per-phase table of lsd_core_kernel from an ncu source-page export (`ncu -i rep --page source --print-source cuda,sass
--csv`, gzipped by tools/profile_round.sh): warp instructions and stall samples per device function and per marked part
of region_grow, plus the hottest source lines.

    python tools/lsd_phase_table.py gpurun_out/r02_lsd_core_source.csv.gz 4096 profiles/r02_lsd_core_phases.md
"""
import csv
import gzip
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_csv, frames, dst = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(gzip.open(src_csv, "rt")))
cur = hdr = None
data = []
stalls = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1])
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iW = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if r[0] and hdr and r[0] != "Function Name":
        try:
            ln = int(r[0])
        except ValueError:
            continue
        num = lambda x: int(x) if x.isdigit() else 0
        data.append((cur, ln, r[1], num(r[iI]), num(r[iW])))
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "(Not Issued)" not in h and i < len(r) and r[i].isdigit():
                stalls[h] = stalls.get(h, 0) + int(r[i])
tot = sum(d[3] for d in data)
tots = sum(d[4] for d in data) or 1

# source line -> enclosing function of lsd_kernels.cu, and -> marked part inside region_grow
src = open(os.path.join(ROOT, "psl_slam_b200", "csrc", "lsd_kernels.cu")).read().split("\n")
func_of, part_of = {}, {}
fname, part = "(file scope)", None
for i, line in enumerate(src, 1):
    m = re.match(r"^(?:template.*\n)?(?:__device__|__global__|Sums )[^;]*?\b([a-z_0-9]+)\(", line)
    if line.startswith(("__device__", "__global__", "Sums step_sequential", "    lsd_core_kernel(")):
        m2 = re.search(r"\b([a-z_0-9]+)\(", line.replace("__launch_bounds__(", ""))
        if m2:
            fname, part = m2.group(1), None
    if fname == "region_grow":
        for key, name in (("Nbr cur = load_nbr", "setup (seed entry, first neighbourhood)"), ("for (;;) {", "test + guess (dot products, ballots, match.any)"),
                          ("if (inA) {   // speculative commit", "speculative commit + next loads"),
                          ("// verification, under the latency of those loads", "verification (prefix sums, re-test)"),
                          ("if (proven) {", "step epilogue / fallback call"), ("if (s.n >= min_n)", "region angle")):
            if key in line:
                part = name
    func_of[i] = fname
    part_of[i] = part

agg, parts = {}, {}
for f, ln, text, ins, smp in data:
    key = func_of.get(ln, "?") if f == "lsd_kernels.cu" else f
    a = agg.setdefault(key, [0, 0])
    a[0] += ins
    a[1] += smp
    if f == "lsd_kernels.cu" and func_of.get(ln) == "region_grow":
        b = parts.setdefault(part_of.get(ln) or "setup (seed entry, first neighbourhood)", [0, 0])
        b[0] += ins
        b[1] += smp
with open(dst, "w") as out:
    out.write(f"# lsd_core_kernel, where the instructions and the stall samples go ({frames} frames per launch)\n\n")
    out.write(f"`ncu --set full --clock-control none --import-source on -k regex:lsd_core -c 1 python tools/line_bench.py --frames {frames} "
              f"--distinct 512 --reps 0` (tools/profile_round.sh); {tot / frames / 1e6:.2f} M warp instructions per frame, "
              f"{tots} stall samples.  Inlined helpers are attributed to the line that calls them only where ncu does so; `sm_*_intrinsics.hpp` "
              f"rows are shuffles / votes / match.\n\n")
    out.write("| function (lsd_kernels.cu) / header | k warp-instr per frame | share | stall samples |\n|---|---:|---:|---:|\n")
    for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.write(f"| `{k}` | {i / frames / 1e3:.1f} | {100 * i / tot:.1f} % | {100 * s / tots:.1f} % |\n")
    out.write("\n## inside region_grow\n\n| part | k warp-instr per frame | share of the kernel | stall samples |\n|---|---:|---:|---:|\n")
    for k, (i, s) in sorted(parts.items(), key=lambda kv: -kv[1][0]):
        out.write(f"| {k} | {i / frames / 1e3:.1f} | {100 * i / tot:.1f} % | {100 * s / tots:.1f} % |\n")
    st = sum(stalls.values()) or 1
    out.write("\n## stall reasons (all samples)\n\n" + ", ".join(f"{k[6:]} {100 * v / st:.1f} %" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]) + "\n")
    out.write("\n## hottest source lines (by stall samples)\n\n| line | k instr / frame | instr share | samples | source |\n|---|---:|---:|---:|---|\n")
    for f, ln, text, ins, smp in sorted(data, key=lambda d: -d[4])[:28]:
        out.write(f"| {f}:{ln} | {ins / frames / 1e3:.1f} | {100 * ins / tot:.1f} % | {100 * smp / tots:.1f} % | `{text.strip()[:110].replace('|', '/')}` |\n")
print("written", dst)
