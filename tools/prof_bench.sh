#!/bin/bash
# Round artefacts for profiles/: (1) plain bench run, (2) ncu launch list of the same command,
# (3) one --set full capture of the two step kernels at full bench size.
tag=${1:-r01}
CMD="python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 2 --no-configs --late-burnin 0"
$CMD > gpurun_out/${tag}_bench_plain.json 2> gpurun_out/${tag}_bench_plain.err || exit 1
[ -n "$SKIP_LAUNCH_LIST" ] || ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:"event_kernel|feature_kernel|pair_kernel" --launch-skip 1060 --launch-count 4 \
    -o gpurun_out/${tag}_full -f $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
tail -1 gpurun_out/${tag}_bench_plain.json | cut -c1-200
