"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
    python tools/ncu_summary.py kernel   gpurun_out/prof_fast_r01.ncu-rep profiles/r01_fast_cells_ncu.md
    python tools/ncu_summary.py kernels  gpurun_out/prof_line_r01.ncu-rep profiles/r01_line_kernels_ncu.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1][:48]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({len(rows)} launches; gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `stages`.\n\n")
        f.write("| kernel | launches | total µs | share | avg µs |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / tot:.3f} | {t / n:.1f} |\n")


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full: {src}\n\n")
        name_i = hdr.index("Kernel Name")
        f.write("kernel: `" + re.sub(r"\(.*", "", data[0][name_i]) + f"`, {len(data)} captured launches\n\n")
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---:|" * len(data) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {units[i]} | " + " | ".join(d[i] for d in data) + " |\n")


def kernels(src, dst):
    """One column per distinct kernel of the report (its first captured launch: the largest pyramid level)."""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    last = {}
    for d in data:
        last.setdefault(re.sub(r"\(.*", "", d[name_i]).split("::")[-1], d)
    names = list(last)
    extra = ["lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
             "smsp__thread_inst_executed_per_inst_executed.ratio"]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {src}\n\n")
        f.write("| metric | unit | " + " | ".join(f"`{n}`" for n in names) + " |\n")
        f.write("|---|---|" + "---:|" * len(names) + "\n")
        for k in KEYS + extra:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {units[i]} | " + " | ".join(last[n][i] for n in names) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "kernels": kernels}[sys.argv[1]](sys.argv[2], sys.argv[3])
