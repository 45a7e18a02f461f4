#!/bin/bash
# round 2, trip Y (1 GPU): the whole GPU suite on the final code; unsorted-path bench: default (scatter + scan) and with the bucket path forced;
# per-kernel times of the bucket passes
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -n 8 --durations=3 ) > gpurun_out/y_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest_gpu.log
tail -n 8 gpurun_out/y_pytest_gpu.log | cut -c1-300
( timeout 300 python tools/bench_delta_path.py ) > gpurun_out/y_delta_default.log 2>&1; echo "default:  $(tail -n 1 gpurun_out/y_delta_default.log | cut -c1-400)"
( PANDEPTH_BUCKET_MIN_READS=1 timeout 300 python tools/bench_delta_path.py ) > gpurun_out/y_delta_bucket.log 2>&1; echo "bucketed: $(tail -n 1 gpurun_out/y_delta_bucket.log | cut -c1-400)"
( PANDEPTH_BUCKET_MIN_READS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_(bucket|expand|tile)" -s 30 -c 14 --csv --log-file gpurun_out/y_bucket_launches.csv python tools/bench_delta_path.py ) > gpurun_out/y_ncu_bucket.log 2>&1
grep -h "k_" gpurun_out/y_bucket_launches.csv | awk -F'","' '{print $5, $NF}' | head -14
( timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -n 2
