#!/bin/bash
# run the correctness subset + short bench for every tuning variant in build/variants
mkdir -p gpurun_out
: > gpurun_out/tune.log
for so in build/variants/*.so; do
  name=$(basename $so .so)
  export PANDEPTH_B200_LIB=$PWD/$so
  if timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "cfg1 or random_mixed or other_read or unsorted_batch or fixtures" > gpurun_out/tune_$name.test 2>&1; then ok=PASS; else ok=FAIL; fi
  line=$(timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$name $ok $(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_ms=%.2f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], {k:round(v,2) for k,v in d['kernels_ms_per_step'].items()})" 2>&1 | tail -1)" | tee -a gpurun_out/tune.log
done
