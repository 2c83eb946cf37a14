#!/bin/bash
# ncu launch list of ONE alternating step in single-stream mode (per-launch durations must not include a co-running
# kernel; graphs off so that every launch is listed). Run on the GPU box:   bash tests/tools/ncu_step.sh
# Then:  python tests/tools/launch_summary.py gpurun_out/step_launches.csv   (or the per-step slicing in DESIGN.md section 6)
# Keep the launch count bounded (-c): bench.py also runs e2e / eval / instrumented passes, and an unbounded ncu run of all of
# them takes >10 minutes.
set -e
export MLA_OVERLAP=0 MLA_OVERLAP_WGRAD=0 MLA_GRAPHS=0
timeout 200 python bench.py --steps 1 --warmup 3 --no-sweep --no-cpu-baseline --no-extra > gpurun_out/step_plain_ss.json 2> gpurun_out/step_plain_ss.err
echo plain rc=$?
# 3 warm-up steps + 1 timed step ~ 2100 launches including model set-up: -c 1800 keeps three complete steps
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/step_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-sweep --no-cpu-baseline --no-extra > gpurun_out/step_ncu.log 2>&1 || true
wc -l gpurun_out/step_launches.csv
