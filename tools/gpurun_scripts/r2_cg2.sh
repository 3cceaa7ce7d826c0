#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/r2_dbg_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_dbg_smoke.log
tail -n 2 gpurun_out/r2_dbg_smoke.log
{ echo "flush 4"; PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "flush 1000"; PPRTS_B200_FLUSH_EVERY=1000 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "flush 2"; PPRTS_B200_FLUSH_EVERY=2 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 4,16"; PPRTS_B200_CG_AGG=4,16 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 4,8"; PPRTS_B200_CG_AGG=4,8 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 2,8"; PPRTS_B200_CG_AGG=2,8 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 3"; PPRTS_B200_CG_AGG=3 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 4,16 fp64"; KRYLOV=fp64 PPRTS_B200_CG_AGG=4,16 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  } > gpurun_out/r2_cg_iter2.log 2>&1
cat gpurun_out/r2_cg_iter2.log
