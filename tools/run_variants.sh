python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1 2; do
PPRTS_B200_MM_VARIANT=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | python -c "import sys,json; d=json.load(sys.stdin); print('variant $v', d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['kernel_ms_per_step'])"
done
