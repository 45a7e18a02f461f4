#!/bin/bash
# final trip of the round: full GPU suite (xdist), the bench line, launch list + full capture of the dominant kernel,
# windows / delta-path side benches.  tools/summarize_profiles.py r01b turns gpurun_out into profiles/.
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
( timeout 900 python -m pytest tests -m gpu -q -n 6 --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $B ) > gpurun_out/ncu_launch.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 9 -c 1 -f -o gpurun_out/prof_r01b_tile $B ) > gpurun_out/ncu_tile.log 2>&1
( timeout 200 python tools/bench_windows.py 3e9 100 2>&1 | tail -1; timeout 200 python tools/bench_windows.py 3e9 1000 2>&1 | tail -1 ) > gpurun_out/windows_lines.json 2>&1
( timeout 300 python tools/bench_delta_path.py 2>&1 | tail -1 ) > gpurun_out/delta_path.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_reduce_windows -s 11 -c 1 -f -o gpurun_out/prof_r01b_windows python tools/bench_windows.py 3e9 100 ) > gpurun_out/ncu_windows.log 2>&1
for f in smoke.log pytest_gpu.log bench_full.log windows_lines.json delta_path.log; do echo "== $f"; tail -n 8 gpurun_out/$f; done
