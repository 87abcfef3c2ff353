#!/bin/bash
# ncu capture of one steady-state step_kernel launch (after the 256-step burn-in) at 16384 envs.
# usage: tools/prof_step.sh <tag> [full]
tag=${1:-x}
if [ "$2" = "full" ]; then SET="--set full"; else SET="--section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis"; fi
ncu $SET --import-source on --clock-control none --kernel-name regex:"event_kernel|feature_kernel|pair_kernel" --launch-skip 1032 --launch-count 4 \
    -o gpurun_out/prof_$tag -f python bench.py --envs 16384 --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/prof_$tag.log 2>&1
tail -2 gpurun_out/prof_$tag.log | cut -c1-300
