#!/bin/bash
# trip r01d: A/B of k_scan_add variants (tools/build_scan_variants.sh): delta-path bench per variant, then the scan-path
# parity tests of every variant in parallel.
mkdir -p gpurun_out
: > gpurun_out/d_scan_ab.log
for name in base ld ld_i32 ld_t128 ld_t512 ld_c8; do
  export PANDEPTH_B200_LIB=$PWD/build/variants/scan_$name.so
  line=$(timeout 60 python tools/bench_delta_path.py 2>&1 | tail -1)
  echo "$name $(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('scan_ms=%.4f frac=%.3f scatter_ms=%.3f' % (d['k_scan_add']['ms'], d['k_scan_add']['frac_of_measured_peak'], d['k_delta_scatter']['ms']))" 2>&1 | tail -1)" | tee -a gpurun_out/d_scan_ab.log
done
for name in ld ld_i32 ld_t128 ld_t512 ld_c8; do
  ( export PANDEPTH_B200_LIB=$PWD/build/variants/scan_$name.so
    if timeout 150 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "after_a_scan or fixtures or unsorted or 18bit or empty or mixed_paths or spliced" > gpurun_out/d_$name.test 2>&1; then echo "$name PASS"; else echo "$name FAIL"; fi ) >> gpurun_out/d_scan_tests.log &
done
wait
cat gpurun_out/d_scan_tests.log; tail -n 2 gpurun_out/d_*.test
