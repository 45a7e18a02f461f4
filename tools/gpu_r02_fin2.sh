#!/bin/bash
# round 2, last trip (1 GPU): the whole GPU suite, smoke, unsorted-path bench and the bench lines on the final code
mkdir -p gpurun_out
( timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/f2_smoke.log 2>&1; tail -n 1 gpurun_out/f2_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q -n 8 ) > gpurun_out/f2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f2_pytest_gpu.log
tail -n 4 gpurun_out/f2_pytest_gpu.log | cut -c1-300
( timeout 300 python tools/bench_delta_path.py ) > gpurun_out/f2_delta.log 2>&1; echo "unsorted path: $(tail -n 1 gpurun_out/f2_delta.log | cut -c1-330)"
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/f2_bench_ref.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/f2_bench.log 2>&1
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/f2_bench_B.log 2>&1
python - <<'PY'
import json
for f in ("f2_bench_ref", "f2_bench", "f2_bench_B"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l); e = d.get("e2e", {})
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", (d.get("roofline") or {}).get("frac"), "e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_vs_reference_rows"), "ref", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
