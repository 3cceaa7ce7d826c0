#!/bin/bash
# strong scaling of one domain under x/y decomposition: bash tools/scale_decomp.sh NGPU NX NY NZ
N=${1:-2}; NX=${2:-512}; NY=${3:-512}; NZ=${4:-80}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 2 --warmup 3 --nx $NX --ny $NY --nz $NZ --decomp --no-cpu-baseline 2>gpurun_out/decomp_err_$N.log | tee gpurun_out/decomp_${N}_${NX}x${NY}x${NZ}.json
