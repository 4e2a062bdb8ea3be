#!/usr/bin/env bash
# The ncu captures behind profiles/ (run on a B200 box through gpurun; one GPU, never a multi-rank command).
#   tools/profile_step.sh [targets] [tag] [launches|full]   # default 10000 (BASELINE.json configs[1]), tag r2, launches
# One profiler pass per call (run it twice, once per mode), each after the same command has exited 0 without ncu.
# Produces gpurun_out/<tag>_launches_<N>.csv (every launch with its device time) and
# gpurun_out/<tag>_prof_<N>.ncu-rep (--set full of the step kernels); summarise them here with
#   python tools/ncu_raw_summary.py gpurun_out/<tag>_prof_<N>.ncu-rep profiles/<tag>_ncu_full_step_<N>_metrics.csv
#   python tools/kernel_profile_summary.py gpurun_out/<tag>_launches_<N>.csv profiles/<tag>_ncu_full_step_<N>_metrics.csv <N> 100
#   ncu -i gpurun_out/<tag>_prof_<N>.ncu-rep --page source --csv --kernel-name regex:<k> > src.csv
#   python tools/ncu_src_summary.py src.csv ; python tools/ncu_lines.py src.csv clockpunch_b200/libpunch_b200.so <k>
set -euo pipefail
N=${1:-10000}
TAG=${2:-r2}
MODE=${3:-launches}
CMD="python bench.py --targets $N --steps 5 --warmup 3 --no-cpu-baseline --no-extra-configs"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain_$N.log 2>&1        # must exit 0 without ncu first
if [ "$MODE" = launches ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/${TAG}_launches_$N.csv $CMD > gpurun_out/${TAG}_ncu_launches_$N.log 2>&1
else
  ncu --set full --clock-control none --import-source on \
      -k regex:"propagate_sigma|ukf_finish|pre_step" -s 9 -c 6 -o gpurun_out/${TAG}_prof_$N $CMD > gpurun_out/${TAG}_ncu_full_$N.log 2>&1
fi
