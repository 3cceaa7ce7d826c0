#!/bin/bash
# Multi-GPU visit (N = $1): decomposition / band-sharding parity tests and the bench line with its c3 + decomp legs.
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_decomp.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_pytest_multi_${N}gpu.log
tail -5 gpurun_out/r2_pytest_multi_${N}gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
cat gpurun_out/r2_bench_${N}gpu.json
tail -5 gpurun_out/r2_bench_${N}gpu.err
