#!/bin/bash
# Round profile on the GPU box (run under gpurun from the repo root): the plain bench first, then the ncu passes of
# B200_PROFILING.md.  The .ncu-rep files are summarised on the box (tools/ncu_summary.py) and only the summaries
# and the raw-page CSVs come back in gpurun_out/ (the reports themselves exceed the 64 MiB return limit).
set -u
TAG=${1:-r01}
O=gpurun_out
T=/tmp/psl_prof
mkdir -p $T
python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err || { tail -5 $O/bench_${TAG}.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err
# launch list of the same command (capped: warm-up + timed steps of the device-resident loop)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches_${TAG}.log 2>&1
# full sets (512 frames per launch keeps ncu's save/restore between replay passes small)
ncu --set full --clock-control none \
    -k regex:"lsd_core|line_post|lbd_kernel|sobel|resize_exact|lsd_gradient|fast_|octree|describe|proj_candidates|proj_resolve" \
    -c 20 -o $T/prof_main -f python bench.py --frames 512 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_main_${TAG}.log 2>&1
ncu --set full --clock-control none -k regex:"resize_words|gauss7" -c 3 -o $T/prof_stream -f \
    python bench.py --frames 512 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_stream_${TAG}.log 2>&1
for r in main stream; do
  python tools/ncu_summary.py kernels $T/prof_$r.ncu-rep $O/${TAG}_kernels_${r}_ncu.md
  ncu -i $T/prof_$r.ncu-rep --page raw --csv > $O/${TAG}_kernels_${r}_raw.csv 2>/dev/null
done
sed -i "s#$T/#gpurun_out/#" $O/${TAG}_kernels_*_ncu.md
ls -la $O | tail -12; du -sh $O
