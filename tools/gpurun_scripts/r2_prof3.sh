#!/bin/bash
mkdir -p gpurun_out
PROF_MAXIT=2 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_prec_down|k_prec_up|k_prec_pack' -c 6 -f -o gpurun_out/r2_prof_prec_packed \
    python tools/profile_step.py > gpurun_out/r2_prof_full3.log 2>&1
tail -n 2 gpurun_out/r2_prof_full3.log
