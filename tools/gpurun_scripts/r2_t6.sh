#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/r2_dbg_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_dbg_smoke.log
tail -n 2 gpurun_out/r2_dbg_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_goldens.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t2.log; tail -n 4 gpurun_out/r2_t2.log
PRECONDS=3,1 KRYLOV=mixed timeout 300 python tools/iter_experiment.py
PPRTS_B200_CG_SEPARATE_RESTRICT=1 PRECONDS=3 KRYLOV=mixed timeout 300 python tools/iter_experiment.py
timeout 600 python bench.py --steps 5 --warmup 3 --legs c1,c5 --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms_per_step'], d['c1'], d['c5'])"; tail -n 3 gpurun_out/r2_bench2.err
