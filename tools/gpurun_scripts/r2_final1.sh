#!/bin/bash
# one-GPU record of the round: full parity suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_pytest_gpu.log; tail -n 4 gpurun_out/r2_pytest_gpu.log
( timeout 120 python __graft_entry__.py --smoke ) 2>&1 | tail -n 1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -n 2 gpurun_out/r2_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['step']['frac'], d['iterations'], d['parity']['default_tolerances']['max_rel_err'], d['parity'].get('parity_tol'), d['c1'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err; cut -c1-400 gpurun_out/r2_bench_reference.json
