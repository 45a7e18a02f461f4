#!/bin/bash
# correctness subset + short device-resident bench with the default build
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step_ms=%.2f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], {k:round(v,2) for k,v in d['kernels_ms_per_step'].items()})"
