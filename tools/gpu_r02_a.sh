#!/bin/bash
# round 2, trip A: first run of the device-side BGZF decode.  Parity first, then throughput (resident + feeder + CLI),
# then sanitizer and ncu evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_env.log 2>&1; nproc >> gpurun_out/a_env.log; cat /sys/fs/cgroup/cpu.max >> gpurun_out/a_env.log 2>&1; free -g >> gpurun_out/a_env.log; df -h /tmp >> gpurun_out/a_env.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
( timeout 900 python -m pytest tests/test_bgzf_gpu.py -x -q --durations=5 ) > gpurun_out/a_pytest_bgzf.log 2>&1; echo "pytest bgzf rc=$?" >> gpurun_out/a_pytest_bgzf.log
( timeout 1500 python -m pytest tests -m gpu -q -n 8 --durations=8 ) > gpurun_out/a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest_gpu.log
# CLI whole-process, 300 Mb x 50x flavour A (100 M reads) and 30 Mb flavour B (10 M reads, SEQ+QUAL)
( REPS=2 timeout 900 bash tools/bench_cli.sh 300 A ) > gpurun_out/a_cli_300_A.log 2>&1
( REPS=2 timeout 900 bash tools/bench_cli.sh 30 B ) > gpurun_out/a_cli_30_B.log 2>&1
( timeout 600 python tools/bench_bgzf.py /tmp/pdcli_300_A/s.bam --steps 3 ) > gpurun_out/a_bgzf_A.log 2>&1; echo "rc=$?" >> gpurun_out/a_bgzf_A.log
( timeout 600 python tools/bench_bgzf.py /tmp/pdcli_30_B/s.bam --steps 3 ) > gpurun_out/a_bgzf_B.log 2>&1; echo "rc=$?" >> gpurun_out/a_bgzf_B.log
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/a_bench.log 2>&1; echo "rc=$?" >> gpurun_out/a_bench.log
# ncu: launch list of one resident pass per flavour, full capture of one k_bgzf_decode launch
B="python tools/bench_bgzf.py /tmp/pdcli_300_A/s.bam --steps 1"
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/a_launches_A.csv $B ) > gpurun_out/a_ncu_launch_A.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 2 -c 1 -f -o gpurun_out/prof_r02a_bgzf_A $B ) > gpurun_out/a_ncu_bgzf_A.log 2>&1
B="python tools/bench_bgzf.py /tmp/pdcli_30_B/s.bam --steps 1"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 2 -c 1 -f -o gpurun_out/prof_r02a_bgzf_B $B ) > gpurun_out/a_ncu_bgzf_B.log 2>&1
# compute-sanitizer: memcheck + racecheck over the smoke run and the device-decode parity tests (small inputs)
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke ) > gpurun_out/a_san_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/a_san_memcheck_smoke.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke ) > gpurun_out/a_san_racecheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/a_san_racecheck_smoke.log
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bgzf_gpu.py -x -q -k "fixtures or tiny or straddling or corrupt" ) > gpurun_out/a_san_memcheck_bgzf.log 2>&1; echo "rc=$?" >> gpurun_out/a_san_memcheck_bgzf.log
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_bgzf_gpu.py -x -q -k "fixtures" ) > gpurun_out/a_san_racecheck_bgzf.log 2>&1; echo "rc=$?" >> gpurun_out/a_san_racecheck_bgzf.log
for f in a_smoke a_pytest_bgzf a_pytest_gpu a_cli_300_A a_cli_30_B a_bgzf_A a_bgzf_B a_bench a_san_memcheck_smoke a_san_racecheck_smoke a_san_memcheck_bgzf a_san_racecheck_bgzf; do echo "== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-600; done
