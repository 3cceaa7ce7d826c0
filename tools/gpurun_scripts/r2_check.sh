#!/bin/bash
# Round-2 GPU-box visit: full parity suite, the default bench line (all single-GPU legs), lookup kernel comparison.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json
tail -5 gpurun_out/r2_bench.err
PPRTS_B200_LOOKUP_PER_THREAD=1 python bench.py --steps 3 --warmup 3 --legs "" --no-cpu-baseline > gpurun_out/r2_bench_oldlookup.json 2>> gpurun_out/r2_bench.err
