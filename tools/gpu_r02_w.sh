#!/bin/bash
# round 2, trip W (1 GPU): bucketed unsorted path after the device-side run count: parity, unsorted-path bench with and without it
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_engine_gpu.py -q -n 8 -k "bucketed or striped or random_mixed or fixtures or range_sharded" ) > gpurun_out/x_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest_gpu.log
tail -n 12 gpurun_out/x_pytest_gpu.log | cut -c1-400
( timeout 300 python tools/bench_delta_path.py ) > gpurun_out/x_delta_bucket.log 2>&1; echo "bucketed: $(tail -n 1 gpurun_out/x_delta_bucket.log | cut -c1-1200)"
( PANDEPTH_BUCKET_MIN_READS=0 timeout 300 python tools/bench_delta_path.py ) > gpurun_out/x_delta_scatter.log 2>&1; echo "scatter:  $(tail -n 1 gpurun_out/x_delta_scatter.log | cut -c1-700)"
