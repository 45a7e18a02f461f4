#!/bin/bash
# evidence run: launch list + full captures of the hot kernels + the bench line; tools/summarize_profiles.py turns them into profiles/
TAG=${1:-r01}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $B ) > gpurun_out/ncu_launch.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 9 -c 1 -f -o gpurun_out/prof_${TAG}_tile $B ) > gpurun_out/ncu_tile.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reduce_units -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_reduce $B ) > gpurun_out/ncu_reduce.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_cut$' -s 9 -c 1 -f -o gpurun_out/prof_${TAG}_cut $B ) > gpurun_out/ncu_cut.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scan_add|k_delta_scatter" -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_scan python tools/bench_delta_path.py 3e8 ) > gpurun_out/ncu_scan.log 2>&1
( timeout 600 python tools/bench_delta_path.py ) > gpurun_out/delta_path.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_full.log 2>&1
tail -1 gpurun_out/delta_path.log; tail -1 gpurun_out/bench_full.log; ls -la gpurun_out/*.ncu-rep
