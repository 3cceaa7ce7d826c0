#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t4.log; tail -n 6 gpurun_out/r2_t4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches.csv python tools/profile_step.py > gpurun_out/r2_prof_step.log 2>&1
tail -n 2 gpurun_out/r2_prof_step.log
