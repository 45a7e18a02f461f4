#!/bin/bash
# first GPU trip: smoke, gpu tests, small + full bench. Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( timeout 300 python bench.py --genome 3e8 --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_small.log 2>&1; echo "rc=$?" >> gpurun_out/bench_small.log
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_small.log gpurun_out/bench_full.log
