#!/bin/bash
# 8-GPU visit: the bench line with its c3 + decomp legs only (the parity tests ran on 2 and 4 GPUs)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
cat gpurun_out/r2_bench_8gpu.json | cut -c1-300
tail -n 5 gpurun_out/r2_bench_8gpu.err
