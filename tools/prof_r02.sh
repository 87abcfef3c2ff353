#!/bin/bash
# Round-2 artefacts for profiles/: plain bench run (timed), ncu launch list of the same c5 command, ncu launch lists
# of one SAC (C4) / one QMIX (C3) update, one --set full capture of the update GEMM kernels.
tag=${1:-r02}
( time python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err ) 2> gpurun_out/${tag}_bench_default.time
CMD="python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 2 --no-configs --late-burnin 0"
$CMD > gpurun_out/${tag}_bench_plain.json 2> gpurun_out/${tag}_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/sac_update_prof.py > gpurun_out/${tag}_sac_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_sac_update_launches.csv python tools/sac_update_prof.py > gpurun_out/${tag}_sac_ncu.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name regex:"gemm_tc_kernel|gemm_tc_reduce|gemm_skinny" --launch-count 12 \
    -o gpurun_out/${tag}_sac_gemm_full -f python tools/sac_update_prof.py > gpurun_out/${tag}_sac_ncu_full.log 2>&1
tail -c 300 gpurun_out/${tag}_bench_default.time
