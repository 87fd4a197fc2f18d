#!/bin/bash
# Round profile on the GPU box (run under gpurun from the repo root): the plain bench lines first, then the ncu passes
# of B200_PROFILING.md.  The .ncu-rep files are summarised on the box (tools/ncu_summary.py) and only the summaries,
# the raw-page CSVs and the source-page CSV of the dominant kernel come back in gpurun_out/.
set -u
TAG=${1:-r02}
O=gpurun_out
T=/tmp/psl_prof
mkdir -p $T
python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err || { tail -5 $O/bench_${TAG}.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err
for c in cfg1 cfg2 cfg3 cfg5; do
  python bench.py --config $c --steps 5 > $O/bench_${TAG}_$c.json 2> $O/bench_${TAG}_$c.err || tail -3 $O/bench_${TAG}_$c.err
done
# launch list of the same command (capped: warm-up + timed steps of the device-resident loop)
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:psl:: -c 2600 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches_${TAG}.log 2>&1
# the dominant kernel at the bench's own launch size (4096 frames per launch): full set + source counters
ncu --set full --clock-control none --import-source on -k regex:lsd_core -c 1 -o $T/prof_core -f \
    python tools/line_bench.py --frames 4096 --distinct 512 --reps 0 > $O/ncu_core_${TAG}.log 2>&1
python tools/ncu_summary.py kernels $T/prof_core.ncu-rep $O/${TAG}_kernels_core_ncu.md
ncu -i $T/prof_core.ncu-rep --page raw --csv > $O/${TAG}_kernels_core_raw.csv 2>/dev/null
ncu -i $T/prof_core.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > $O/${TAG}_lsd_core_source.csv.gz
# full sets of the other kernels (the ORB kernels run 512 frames per launch in the bench as well; the line kernels are
# captured at 512 frames per launch to keep ncu's save / restore between replay passes small)
ncu --set full --clock-control none \
    -k regex:"line_post|line_scan|lbd_kernel|sobel|resize_exact|lsd_gradient|lsd_seed_order|fast_|octree|describe|proj_candidates|proj_resolve|color_to_gray" \
    -c 30 -o $T/prof_main -f python bench.py --frames 512 --distinct 64 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_main_${TAG}.log 2>&1
ncu --set full --clock-control none -k regex:"resize_words|gauss7" -c 3 -o $T/prof_stream -f \
    python bench.py --frames 512 --distinct 64 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_stream_${TAG}.log 2>&1
for r in main stream; do
  python tools/ncu_summary.py kernels $T/prof_$r.ncu-rep $O/${TAG}_kernels_${r}_ncu.md
  ncu -i $T/prof_$r.ncu-rep --page raw --csv > $O/${TAG}_kernels_${r}_raw.csv 2>/dev/null
done
sed -i "s#$T/#gpurun_out/#" $O/${TAG}_kernels_*_ncu.md
ls -la $O | tail -12; du -sh $O
