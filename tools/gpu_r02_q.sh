#!/bin/bash
# round 2, trip Q (1 GPU): after the device-side CRC32 (PD_CRC) and the bounded pinned pool of the command line: BGZF parity incl. the CRC
# tests, both bench arms again (the e2e numbers of record), decode bench + ncu captures of the decode kernels, CLI on the flavour-B file.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_bgzf_gpu.py tests/test_cli_gpu.py -q -n 8 ) > gpurun_out/q_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest_gpu.log
tail -n 5 gpurun_out/q_pytest_gpu.log | cut -c1-300
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/q_bench.log 2>&1; echo "rc=$?" >> gpurun_out/q_bench.log
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/q_bench_B.log 2>&1; echo "rc=$?" >> gpurun_out/q_bench_B.log
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
( timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 5 ) > gpurun_out/q_bgzf_A.log 2>&1
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 5 ) > gpurun_out/q_bgzf_B.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bgzf_A.csv python tools/bench_bgzf.py /tmp/cA.bam --steps 1 --no-feed ) > gpurun_out/q_ncu_launch_A.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bgzf_B.csv python tools/bench_bgzf.py /tmp/cB.bam --steps 1 --no-feed ) > gpurun_out/q_ncu_launch_B.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 12 -c 1 -f -o gpurun_out/prof_r02_bgzf_A python tools/bench_bgzf.py /tmp/cA.bam --steps 1 --no-feed ) > gpurun_out/q_ncu_bgzf_A.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 12 -c 1 -f -o gpurun_out/prof_r02_bgzf_B python tools/bench_bgzf.py /tmp/cB.bam --steps 1 --no-feed ) > gpurun_out/q_ncu_bgzf_B.log 2>&1
( REPS=2 THREADS=16 timeout 900 bash tools/bench_cli.sh 100 B ) > gpurun_out/q_cli_100_B.log 2>&1
( REPS=2 timeout 900 bash tools/bench_cli.sh 3000 A ) > gpurun_out/q_cli_3000_A.log 2>&1
head -n 7 gpurun_out/q_cli_100_B.log; head -n 6 gpurun_out/q_cli_3000_A.log
python - <<'PY'
import json
for f in ("q_bench", "q_bench_B"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l); e = d.get("e2e", {})
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", (d.get("roofline") or {}).get("frac"), "e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_vs_reference_rows"), "ref", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
for f in ("q_bgzf_A", "q_bgzf_B"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1]); r = d["resident"]; fd = d["feed"]
        print(f, "resident", round(r["ms_per_step"], 2), "ms", round(r["Mreads_per_s"], 1), "Mreads/s", round(r["inflated_GBps"], 1), "GB/s | feed", round(fd["s_best"] * 1e3, 2), "ms", round(fd["Mreads_per_s"], 1))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
