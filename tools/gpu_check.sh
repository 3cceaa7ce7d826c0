#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one ncu --set full capture of the hot kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench_err.log
cat gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
PROF_MAXIT=3 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_matmult_tma|k_prec_down|k_prec_up|k_bicg' -c 16 -f -o gpurun_out/prof_hot \
    python tools/profile_step.py > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/prof_full.log
