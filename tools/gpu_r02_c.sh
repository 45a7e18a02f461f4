#!/bin/bash
# round 2, trip C: tile-kernel A/B (bulk flush, SEG/unroll variants), cut overlap in the bench step, faster decode loop,
# start-up breakdown of the CLI.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_bgzf_gpu.py tests/test_engine_gpu.py -q -n 8 -x ) > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
for v in base bulk bulkadd seg1024u8 u8c3; do
  if [ $v = base ]; then L=""; else L="PANDEPTH_B200_LIB=$PWD/build/variants/libpd_$v.so"; fi
  ( env $L timeout 300 python tools/bench_tile.py --steps 10 ) >> gpurun_out/c_tile_ab.log 2>&1
done
for v in base bulk seg1024u8; do
  if [ $v = base ]; then L=""; else L="PANDEPTH_B200_LIB=$PWD/build/variants/libpd_$v.so"; fi
  ( env $L timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline ) >> gpurun_out/c_bench_variants.log 2>&1
done
# decode loop: same files as trip B
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
( timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 3 ) > gpurun_out/c_bgzf_A.log 2>&1; echo "rc=$?" >> gpurun_out/c_bgzf_A.log
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 3 ) > gpurun_out/c_bgzf_B.log 2>&1; echo "rc=$?" >> gpurun_out/c_bgzf_B.log
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 3 --piece-mb 96 ) > gpurun_out/c_bgzf_B96.log 2>&1; echo "rc=$?" >> gpurun_out/c_bgzf_B96.log
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 2 -c 1 -f -o gpurun_out/prof_r02c_bgzf_B python tools/bench_bgzf.py /tmp/cB.bam --steps 1 --no-resident ) > gpurun_out/c_ncu_bgzf_B.log 2>&1
# CLI start-up breakdown
for i in 1 2; do ( PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/cA.bam -o /tmp/cA_out -t 16 ) >> gpurun_out/c_cli_timing.log 2>&1; done
( /usr/bin/env time -v pandepth_b200/bin/pandepth -i /tmp/cA.bam -o /tmp/cA_out -t 16 ) >> gpurun_out/c_cli_timing.log 2>&1
for f in c_pytest c_tile_ab c_bench_variants c_bgzf_A c_bgzf_B c_bgzf_B96 c_cli_timing; do echo "== $f"; tail -n 30 gpurun_out/$f.log | cut -c1-1800; done
