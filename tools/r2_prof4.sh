#!/bin/bash
# ncu --set full of the Krylov kernels at the final code (first Krylov iteration of one step); the raw page is exported on
# the box because the report itself is larger than what gpurun copies back
mkdir -p gpurun_out
PROF_MAXIT=2 timeout 1200 ncu --set full --clock-control none --profile-from-start off \
    -k regex:'k_matmult_tma|k_prec_down|k_prec_up|k_bicg32_update|k_cg_' -c 26 -f -o /tmp/r2_prof_hot \
    python tools/profile_step.py > gpurun_out/r2_prof_full4.log 2>&1
tail -n 2 gpurun_out/r2_prof_full4.log
ncu -i /tmp/r2_prof_hot.ncu-rep --page raw --csv > gpurun_out/r2_prof_hot.csv
ls -la gpurun_out/r2_prof_hot.csv
