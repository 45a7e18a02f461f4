#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( timeout 300 python tools/pcie_probe.py ) > gpurun_out/pcie.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
for f in smoke.log pytest_gpu.log pcie.log bench_full.log; do echo "== $f"; tail -n 6 gpurun_out/$f; done
