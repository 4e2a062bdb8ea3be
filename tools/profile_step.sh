#!/usr/bin/env bash
# The ncu captures behind profiles/ (run on a B200 box through gpurun; one GPU, never a multi-rank command).
#   tools/profile_step.sh [targets]            # default 10000 (BASELINE.json configs[1])
# Produces gpurun_out/launches_<N>.csv (every launch with its device time) and
# gpurun_out/prof_<N>.ncu-rep (--set full of the step kernels); summarise them here with
#   python tools/ncu_raw_summary.py gpurun_out/prof_<N>.ncu-rep profiles/<name>_metrics.csv
#   ncu -i gpurun_out/prof_<N>.ncu-rep --page source --csv --kernel-name regex:<k> > src.csv
#   python tools/ncu_src_summary.py src.csv ; python tools/ncu_lines.py src.csv clockpunch_b200/libpunch_b200.so <k>
set -euo pipefail
N=${1:-10000}
CMD="python bench.py --targets $N --steps 5 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$N.log 2>&1        # must exit 0 without ncu first
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$N.csv $CMD > gpurun_out/ncu_launches_$N.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"propagate_sigma|ukf_finish|pre_step" -s 9 -c 6 -o gpurun_out/prof_$N $CMD > gpurun_out/ncu_full_$N.log 2>&1
