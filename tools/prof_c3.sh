#!/bin/bash
# ncu artefacts for the policy path at config C3: launch list of the rollout bench (eager launches) and one
# --set full capture of the tcgen05 linear kernel (the GRU input product, M=16384 K=352 N=192).
tag=${1:-r01}
CMD="python bench.py --workload c3 --steps 10 --warmup 3 --burnin 32 --no-graph --no-cpu --e2e-steps 2"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_c3_launches.csv $CMD > gpurun_out/${tag}_c3_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:"linear_tc_kernel" --launch-skip 300 --launch-count 2 \
    -o gpurun_out/${tag}_c3_tc -f $CMD > gpurun_out/${tag}_c3_ncu_tc.log 2>&1
tail -2 gpurun_out/${tag}_c3_ncu_tc.log | cut -c1-200
