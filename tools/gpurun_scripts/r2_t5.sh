#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/r2_dbg_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_dbg_smoke.log
tail -n 2 gpurun_out/r2_dbg_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_goldens.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t2.log; tail -n 4 gpurun_out/r2_t2.log
PRECONDS=3 KRYLOV=mixed timeout 300 python tools/iter_experiment.py
PPRTS_B200_CG_UNFUSED=1 PRECONDS=3 KRYLOV=mixed timeout 300 python tools/iter_experiment.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches.csv python tools/profile_step.py > gpurun_out/r2_prof_step.log 2>&1
tail -n 2 gpurun_out/r2_prof_step.log
