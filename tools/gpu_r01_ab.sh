#!/bin/bash
# A/B trip: full GPU suite on the current default build, then the tuning variants (base = previous commit's kernels)
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
( timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/ab.log
for so in build/variants/*.so; do
  name=$(basename $so .so)
  export PANDEPTH_B200_LIB=$PWD/$so
  line=$(timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$name $(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_ms=%.3f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})" 2>&1 | tail -1)" | tee -a gpurun_out/ab.log
  w=$(timeout 300 python tools/bench_windows.py 3e9 100 2>&1 | tail -1)
  echo "$name windows100 $w" | tee -a gpurun_out/ab.log
done
unset PANDEPTH_B200_LIB
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_reduce_windows -s 2 -c 1 -f -o gpurun_out/prof_r01_windows python tools/bench_windows.py 3e8 100 ) > gpurun_out/ncu_windows.log 2>&1
for f in smoke.log pytest_gpu.log; do echo "== $f"; tail -n 12 gpurun_out/$f; done
cat gpurun_out/ab.log
