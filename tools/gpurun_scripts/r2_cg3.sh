#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/r2_dbg_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_dbg_smoke.log
tail -n 2 gpurun_out/r2_dbg_smoke.log
{ echo "default (4,8)"; PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 4"; PPRTS_B200_CG_AGG=4 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 4,8,16"; PPRTS_B200_CG_AGG=4,8,16 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 4,16"; PPRTS_B200_CG_AGG=4,16 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "agg 2,4,8"; PPRTS_B200_CG_AGG=2,4,8 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  echo "fp64 default"; KRYLOV=fp64 PRECONDS=3 timeout 300 python tools/iter_experiment.py
  } > gpurun_out/r2_cg_iter3.log 2>&1
cat gpurun_out/r2_cg_iter3.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_goldens.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t2.log; tail -n 4 gpurun_out/r2_t2.log
