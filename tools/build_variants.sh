#!/bin/bash
# tuning variants of the CUDA library -> build/variants/*.so (git-ignored; they travel with gpurun)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
b() { name=$1; shift; /usr/local/cuda/bin/nvcc $F "$@" -o build/variants/$name.so pandepth_b200/csrc/pd_engine.cu & }
b c4u8
b c5u8 -DPD_TILE_MIN_CTAS=5
b c4u16 -DPD_TILE_UNROLL=16
b c6u4 -DPD_TILE_MIN_CTAS=6 -DPD_TILE_UNROLL=4
b s2048c3u8 -DPD_SEG=2048 -DPD_TILE_MIN_CTAS=3
b s2048c3u16 -DPD_SEG=2048 -DPD_TILE_MIN_CTAS=3 -DPD_TILE_UNROLL=16
wait
ls -la build/variants
