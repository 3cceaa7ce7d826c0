#!/bin/bash
# ncu --set full of the two Galerkin accumulation kernels (coalesced rows vs the per-aggregate walk) on one C2 step
mkdir -p gpurun_out
for v in rows walk; do
  if [ $v = walk ]; then export PPRTS_B200_CG_ACC_WALK=1; else unset PPRTS_B200_CG_ACC_WALK; fi
  PROF_MAXIT=1 timeout 600 ncu --set full --clock-control none --profile-from-start off \
      -k regex:'k_cg_accumulate' -c 2 -f -o /tmp/r2_prof_acc_$v python tools/profile_step.py > gpurun_out/r2_prof_acc_$v.log 2>&1
  tail -n 1 gpurun_out/r2_prof_acc_$v.log
  ncu -i /tmp/r2_prof_acc_$v.ncu-rep --page raw --csv > gpurun_out/r2_prof_acc_$v.csv
done
ls -la gpurun_out/r2_prof_acc_*.csv
