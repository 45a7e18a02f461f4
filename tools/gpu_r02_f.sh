#!/bin/bash
# round 2, trip F (1 GPU; trip E repeated after the shared-window alignment fix, plus decoders-per-warp A/B): parity suite, the bench contract (both arms, flavour A full size; flavour B sample), decode throughput,
# ncu evidence for the round (launch list of the bench step, full captures of k_tile_accumulate and k_bgzf_decode), CLI times.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -n 8 --durations=5 ) > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest_gpu.log
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/f_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/f_bench_ref.log
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/f_bench.log 2>&1; echo "rc=$?" >> gpurun_out/f_bench.log
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/f_bench_B.log 2>&1; echo "rc=$?" >> gpurun_out/f_bench_B.log
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
( timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 3 ) > gpurun_out/f_bgzf_A.log 2>&1; echo "rc=$?" >> gpurun_out/f_bgzf_A.log
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 3 ) > gpurun_out/f_bgzf_B.log 2>&1; echo "rc=$?" >> gpurun_out/f_bgzf_B.log
for K in 1 2 4; do ( PANDEPTH_BGZF_K=$K timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 3 ) > gpurun_out/f_bgzf_B_K$K.log 2>&1; done
for K in 1 2; do ( PANDEPTH_BGZF_K=$K timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 3 ) > gpurun_out/f_bgzf_A_K$K.log 2>&1; done
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_launches.csv $B ) > gpurun_out/f_ncu_launch.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 9 -c 1 -f -o gpurun_out/prof_r02f_tile $B ) > gpurun_out/f_ncu_tile.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 4 -c 1 -f -o gpurun_out/prof_r02f_bgzf_B python tools/bench_bgzf.py /tmp/cB.bam --steps 1 ) > gpurun_out/f_ncu_bgzf_B.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 4 -c 1 -f -o gpurun_out/prof_r02f_bgzf_A python tools/bench_bgzf.py /tmp/cA.bam --steps 1 ) > gpurun_out/f_ncu_bgzf_A.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_bgzf_B.csv python tools/bench_bgzf.py /tmp/cB.bam --steps 1 --no-resident ) > gpurun_out/f_ncu_launch_B.log 2>&1
( REPS=2 timeout 900 bash tools/bench_cli.sh 3000 A ) > gpurun_out/f_cli_3000_A.log 2>&1
( REPS=2 THREADS=16 timeout 900 bash tools/bench_cli.sh 3000 A -w 100 ) > gpurun_out/f_cli_3000_A_w100.log 2>&1
for f in f_pytest_gpu f_bench_ref f_bench f_bench_B f_bgzf_A f_bgzf_B f_bgzf_B_K1 f_bgzf_B_K2 f_bgzf_B_K4 f_bgzf_A_K1 f_bgzf_A_K2 f_cli_3000_A f_cli_3000_A_w100; do echo "== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-2500; done
