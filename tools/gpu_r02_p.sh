#!/bin/bash
# round 2, trip P (1 GPU): parity after the in-line cut kernels and the scatter-flag change; the device arm of the bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_bgzf_gpu.py -q -n 8 ) > gpurun_out/p_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest_gpu.log
tail -n 6 gpurun_out/p_pytest_gpu.log | cut -c1-300
( timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline ) > gpurun_out/p_bench.log 2>&1; tail -n 1 gpurun_out/p_bench.log | cut -c1-1700
