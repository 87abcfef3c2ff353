#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sac_launches.csv python tools/policy_bench.py > gpurun_out/sac_ncu.log 2>&1
tail -3 gpurun_out/sac_ncu.log
