#!/bin/bash
# short trip r01c (the round's last GPU minutes): the scan kernel without the delta re-zero (lazy clear in the engine),
# CRAM input through the CLI, then the delta-path bench and a full capture of k_scan_add.
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_engine_gpu.py tests/test_cli_gpu.py -m gpu -q -n 6 -k "not random_bams" --durations=5 ) > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -n 12 gpurun_out/c_pytest.log
( timeout 90 python tools/bench_delta_path.py 2>&1 | tail -1 ) > gpurun_out/c_delta_path.log 2>&1
cat gpurun_out/c_delta_path.log
( timeout 60 python __graft_entry__.py smoke ) > gpurun_out/c_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c_smoke.log
tail -n 2 gpurun_out/c_smoke.log
( timeout 90 ncu --set full --clock-control none --import-source on -k regex:k_scan_add -s 3 -c 1 -f -o gpurun_out/prof_r01c_scan python tools/bench_delta_path.py ) > gpurun_out/c_ncu_scan.log 2>&1
tail -n 3 gpurun_out/c_ncu_scan.log
( timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/c_bench.log 2>&1; echo "rc=$?" >> gpurun_out/c_bench.log
tail -n 2 gpurun_out/c_bench.log
