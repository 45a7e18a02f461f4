#!/bin/bash
# round 2, trip O (1 GPU): adjacent-duplicate aggregation in the tile kernel: parity, then A/B against the build without it at 50x and 200x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_bgzf_gpu.py -q -n 8 ) > gpurun_out/o_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest_gpu.log
tail -n 6 gpurun_out/o_pytest_gpu.log | cut -c1-300
for lib in pandepth_b200/libpandepth_b200.so build/variants/libpd_noagg.so; do
  for reads in 111111111 444444444; do
    echo "$(basename $lib) reads=$reads: $(PANDEPTH_B200_LIB=$lib timeout 300 python tools/bench_tile.py --steps 10 --reads $reads 2>&1 | tail -n 1 | cut -c1-400)" | tee -a gpurun_out/o_tile_ab.log
  done
done
( timeout 600 python tools/bench_variants.py --steps 10 ) > gpurun_out/o_variants.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/o_variants.log").read().strip().splitlines()[-1])
    for k, v in d["variants"].items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
except Exception as e:
    print("variants FAILED", e); print(open("gpurun_out/o_variants.log").read()[-2000:])
PY
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline ) > gpurun_out/o_bench.log 2>&1; tail -n 1 gpurun_out/o_bench.log | cut -c1-1500
