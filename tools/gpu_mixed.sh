#!/bin/bash
# mixed-precision Krylov: parity tests + bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed" 2>&1 | tail -15 > gpurun_out/pytest_mixed.log
cat gpurun_out/pytest_mixed.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed_err.log
cat gpurun_out/bench_mixed.json; tail -5 gpurun_out/bench_mixed_err.log
