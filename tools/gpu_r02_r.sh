#!/bin/bash
# round 2, trip R (1 GPU): CRC32 with sector-wise loads: parity, cost on flavour A / B (feed = with PD_CRC, resident = without), bench e2e
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_bgzf_gpu.py -q -n 8 ) > gpurun_out/r_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest_gpu.log
tail -n 4 gpurun_out/r_pytest_gpu.log | cut -c1-300
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
( timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 5 ) > gpurun_out/r_bgzf_A.log 2>&1
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 5 ) > gpurun_out/r_bgzf_B.log 2>&1
( timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e-arrays --no-cpu-baseline ) > gpurun_out/r_bench.log 2>&1
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays --no-cpu-baseline ) > gpurun_out/r_bench_B.log 2>&1
python - <<'PY'
import json
for f in ("r_bench", "r_bench_B"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l); e = d.get("e2e", {})
        print(f, "value", round(d["value"], 1), "e2e", e.get("value"), e.get("ms_per_step"))
    except Exception as ex:
        print(f, "FAILED", ex)
for f in ("r_bgzf_A", "r_bgzf_B"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1]); r = d["resident"]; fd = d["feed"]
        print(f, "resident", round(r["ms_per_step"], 2), "ms", round(r["Mreads_per_s"], 1), "Mreads/s | feed (CRC)", round(fd["s_best"] * 1e3, 2), "ms", round(fd["Mreads_per_s"], 1))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
