#!/bin/bash
# evidence run: launch list + full captures of the three hot kernels; summaries go to profiles/ (copied by hand)
mkdir -p gpurun_out
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ) > gpurun_out/ncu_launch.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile_accumulate|k_reduce_units|k_cut" -s 30 -c 4 -f -o gpurun_out/prof_r01_tile \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ) > gpurun_out/ncu_tile.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scan_add|k_delta_scatter" -s 4 -c 2 -f -o gpurun_out/prof_r01_scan \
    python tools/bench_delta_path.py 3e8 ) > gpurun_out/ncu_scan.log 2>&1
( timeout 600 python tools/bench_delta_path.py ) > gpurun_out/delta_path.log 2>&1
tail -1 gpurun_out/delta_path.log; tail -3 gpurun_out/ncu_tile.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
