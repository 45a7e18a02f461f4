#!/bin/bash
# round 2, trip B: full GPU suite (incl. range sharding), decode-kernel throughput (resident + feeder) with ncu evidence,
# CLI at cfg2 size, the new bench.py (both arms, flavour A full size; flavour B sample).
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -n 8 --durations=6 ) > gpurun_out/b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest_gpu.log
( REPS=2 timeout 900 bash tools/bench_cli.sh 300 A ) > gpurun_out/b_cli_300_A.log 2>&1
( REPS=1 timeout 900 bash tools/bench_cli.sh 60 B ) > gpurun_out/b_cli_60_B.log 2>&1
( timeout 600 python tools/bench_bgzf.py /tmp/pdcliP_300_A/s.bam --steps 3 ) > gpurun_out/b_bgzf_A.log 2>&1; echo "rc=$?" >> gpurun_out/b_bgzf_A.log
( timeout 600 python tools/bench_bgzf.py /tmp/pdcliP_60_B/s.bam --steps 3 ) > gpurun_out/b_bgzf_B.log 2>&1; echo "rc=$?" >> gpurun_out/b_bgzf_B.log
B="python tools/bench_bgzf.py /tmp/pdcliP_300_A/s.bam --steps 1"
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/b_launches_A.csv $B ) > gpurun_out/b_ncu_launch_A.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 2 -c 1 -f -o gpurun_out/prof_r02b_bgzf_A $B ) > gpurun_out/b_ncu_bgzf_A.log 2>&1
B="python tools/bench_bgzf.py /tmp/pdcliP_60_B/s.bam --steps 1"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 2 -c 1 -f -o gpurun_out/prof_r02b_bgzf_B $B ) > gpurun_out/b_ncu_bgzf_B.log 2>&1
# the bench contract, both arms, on the full cfg2 file (flavour A), then flavour B on a 100 Mb sample
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/b_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/b_bench_ref.log
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/b_bench.log 2>&1; echo "rc=$?" >> gpurun_out/b_bench.log
( REPS=1 timeout 900 bash tools/bench_cli.sh 3000 A ) > gpurun_out/b_cli_3000_A.log 2>&1
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/b_bench_B.log 2>&1; echo "rc=$?" >> gpurun_out/b_bench_B.log
for f in b_pytest_gpu b_cli_300_A b_cli_60_B b_bgzf_A b_bgzf_B b_bench_ref b_bench b_cli_3000_A b_bench_B; do echo "== $f"; tail -n 14 gpurun_out/$f.log | cut -c1-1500; done
