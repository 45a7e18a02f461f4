#!/bin/bash
# second A/B trip: which of the tile-kernel changes pay (trim = default build; sink / halo / lean = compile-time options)
mkdir -p gpurun_out
: > gpurun_out/ab2.log
for name in trim sink halo base; do
  export PANDEPTH_B200_LIB=$PWD/build/variants/$name.so
  if [ $name != base ]; then
    if timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "cfg1 or random_mixed or other_read or fixtures or unsorted_batch or reduce_hint" > gpurun_out/ab2_$name.test 2>&1; then ok=PASS; else ok=FAIL; fi
  else ok=-; fi
  line=$(timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$name $ok $(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_ms=%.3f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})" 2>&1 | tail -1)" | tee -a gpurun_out/ab2.log
done
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
for name in trim lean; do
  export PANDEPTH_B200_LIB=$PWD/build/variants/$name.so
  ( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 9 -c 1 -f -o gpurun_out/prof_r01_tile_$name $B ) > gpurun_out/ncu_tile_$name.log 2>&1
done
cat gpurun_out/ab2.log; tail -2 gpurun_out/ab2_*.test; ls -la gpurun_out/*.ncu-rep
