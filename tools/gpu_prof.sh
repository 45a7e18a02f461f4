#!/bin/bash
# tests + ncu launch list + full capture of the top kernel + bench
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --genome 9e8 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ) > gpurun_out/ncu_launch.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 9 -c 2 -f -o gpurun_out/prof_tile \
    python bench.py --genome 9e8 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ) > gpurun_out/ncu_full.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
for f in pytest_gpu.log ncu_launch.log ncu_full.log bench_full.log; do echo "== $f"; tail -n 4 gpurun_out/$f; done
ls -la gpurun_out
