#!/bin/bash
mkdir -p gpurun_out
LIBV=${1:-}
[ -n "$LIBV" ] && export PANDEPTH_B200_LIB=$PWD/build/variants/$LIBV.so
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 9 -c 1 -f -o gpurun_out/prof_tile2 \
    python bench.py --genome 9e8 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ) > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log | cut -c1-300
