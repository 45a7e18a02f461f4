#!/bin/bash
# round 2, trip Z2 (1 GPU): bucket passes staged in shared memory: parity of the bucketed path, unsorted-path bench, per-kernel times
mkdir -p gpurun_out
( PANDEPTH_BUCKET_MIN_READS=1 timeout 1500 python -m pytest tests/test_engine_gpu.py -q -n 8 -k "bucketed or fixtures or random_mixed or range_sharded or striped" ) > gpurun_out/z2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z2_pytest_gpu.log
tail -n 6 gpurun_out/z2_pytest_gpu.log | cut -c1-300
( PANDEPTH_BUCKET_MIN_READS=1 timeout 300 python tools/bench_delta_path.py ) > gpurun_out/z2_delta_bucket.log 2>&1; echo "bucketed: $(tail -n 1 gpurun_out/z2_delta_bucket.log | cut -c1-400)"
( PANDEPTH_BUCKET_MIN_READS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_(bucket|expand|tile)" -s 30 -c 7 --csv --log-file gpurun_out/z2_bucket_launches.csv python tools/bench_delta_path.py ) > gpurun_out/z2_ncu_bucket.log 2>&1
grep -h "k_" gpurun_out/z2_bucket_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,150-200 | head -8
