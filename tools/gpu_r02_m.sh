#!/bin/bash
# round 2, trip M (1 GPU): tile kernel with one drain site + 32-bit op walk: engine / BGZF / CLI parity, variants bench, ncu of the mixed variant
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_bgzf_gpu.py tests/test_seam_gpu.py -q -n 8 ) > gpurun_out/m2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m2_pytest_gpu.log
tail -n 8 gpurun_out/m2_pytest_gpu.log | cut -c1-300
( timeout 600 python tools/bench_variants.py --steps 10 ) > gpurun_out/m2_variants.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/m2_variants.log").read().strip().splitlines()[-1])
    for k, v in d["variants"].items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
except Exception as e:
    print("variants FAILED", e); print(open("gpurun_out/m2_variants.log").read()[-2000:])
PY
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 3 -c 1 -f -o gpurun_out/prof_r02m2_tile_mixed python tools/bench_variants.py --steps 1 --only mixed ) > gpurun_out/m2_ncu_mixed.log 2>&1
