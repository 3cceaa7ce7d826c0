#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/r2_dbg_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_dbg_smoke.log
tail -n 2 gpurun_out/r2_dbg_smoke.log
{ PRECONDS=3,1 KRYLOV=mixed timeout 300 python tools/iter_experiment.py
  PRECONDS=3,1 KRYLOV=fp64 timeout 300 python tools/iter_experiment.py
  PPRTS_B200_CG_AGG=2 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  PPRTS_B200_CG_AGG=8 PRECONDS=3 timeout 300 python tools/iter_experiment.py ; } > gpurun_out/r2_cg_iter.log 2>&1
cat gpurun_out/r2_cg_iter.log
