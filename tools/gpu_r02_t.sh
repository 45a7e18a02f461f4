#!/bin/bash
# round 2, trip T (1 GPU): CRC32 with the x^(2^k) table: parity, decode bench, bench lines of record, launch list of the engine's kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_bgzf_gpu.py -q -n 8 ) > gpurun_out/t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest_gpu.log
tail -n 3 gpurun_out/t_pytest_gpu.log | cut -c1-300
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
( timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 5 ) > gpurun_out/t_bgzf_A.log 2>&1
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 5 ) > gpurun_out/t_bgzf_B.log 2>&1
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/t_bench_ref.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/t_bench.log 2>&1
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/t_bench_B.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_(cut|tile|fill|reduce|scan|delta|bgzf|filter)" -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ) > gpurun_out/t_ncu_launch.log 2>&1
python - <<'PY'
import json
for f in ("t_bench_ref", "t_bench", "t_bench_B"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l); e = d.get("e2e", {})
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", (d.get("roofline") or {}).get("frac"), "e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_vs_reference_rows"), "ref", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
for f in ("t_bgzf_A", "t_bgzf_B"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1]); r = d["resident"]; fd = d["feed"]
        print(f, "resident", round(r["ms_per_step"], 2), "ms", round(r["Mreads_per_s"], 1), "Mreads/s | feed (CRC)", round(fd["s_best"] * 1e3, 2), "ms", round(fd["Mreads_per_s"], 1))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
grep -c pdk gpurun_out/launches.csv
