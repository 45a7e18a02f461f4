#!/bin/bash
# third A/B trip: LSU-pipe reductions in the tile kernel (paired scan + redux totals, 16-bit packed scan, st.bulk clear)
mkdir -p gpurun_out
: > gpurun_out/ab4.log
for name in pf vec pfvec; do
  export PANDEPTH_B200_LIB=$PWD/build/variants/$name.so
  if timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "cfg1 or random_mixed or other_read or fixtures or unsorted_batch or reduce_hint or 18bit" > gpurun_out/ab4_$name.test 2>&1; then ok=PASS; else ok=FAIL; fi
  line=$(timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$name $ok $(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_ms=%.3f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})" 2>&1 | tail -1)" | tee -a gpurun_out/ab4.log
done
unset PANDEPTH_B200_LIB
cat gpurun_out/ab4.log; tail -2 gpurun_out/ab4_*.test
