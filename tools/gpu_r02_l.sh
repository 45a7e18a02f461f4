#!/bin/bash
# round 2, trip L (1 GPU): full GPU suite after the tile-kernel changes (filter bitmap, queued multi-op reads), the kernel on
# filter / mixed inputs, ncu captures of those two variants.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -n 8 --durations=5 ) > gpurun_out/l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest_gpu.log
tail -n 12 gpurun_out/l_pytest_gpu.log | cut -c1-300
( timeout 600 python tools/bench_variants.py --steps 10 ) > gpurun_out/l_variants.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/l_variants.log").read().strip().splitlines()[-1])
    for k, v in d["variants"].items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
except Exception as e:
    print("variants FAILED", e); print(open("gpurun_out/l_variants.log").read()[-2000:])
PY
for v in mixed filter; do
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 3 -c 1 -f -o gpurun_out/prof_r02l_tile_$v python tools/bench_variants.py --steps 1 --only $v ) > gpurun_out/l_ncu_$v.log 2>&1
done
ls -la gpurun_out/prof_r02l_*
