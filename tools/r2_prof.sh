#!/bin/bash
# parity of the changed kernels, then ncu: launch list of one step + full capture of the lookup / column kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_goldens.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t3.log; tail -n 6 gpurun_out/r2_t3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches.csv python tools/profile_step.py > gpurun_out/r2_prof_step.log 2>&1
PROF_MAXIT=2 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_lookup_tile|k_prec_down|k_prec_up|k_edir_pass' -c 12 -f -o gpurun_out/r2_prof_lookup_prec \
    python tools/profile_step.py > gpurun_out/r2_prof_full.log 2>&1
tail -n 3 gpurun_out/r2_prof_full.log
