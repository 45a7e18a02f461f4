#!/bin/bash
# scan-kernel tuning variants -> build/variants/scan_*.so (git-ignored; they travel with gpurun)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v"
b() { name=$1; shift; /usr/local/cuda/bin/nvcc $F "$@" -o build/variants/scan_$name.so pandepth_b200/csrc/pd_engine.cu 2> build/variants/scan_$name.ptxas & }
b base
b ld -DPD_SCAN_POLL_LD=1
b ld_i32 -DPD_SCAN_POLL_LD=1 -DPD_SCAN_ITEMS=32
b ld_t128 -DPD_SCAN_POLL_LD=1 -DPD_SCAN_THREADS=128
b ld_t512 -DPD_SCAN_POLL_LD=1 -DPD_SCAN_THREADS=512
b ld_c8 -DPD_SCAN_POLL_LD=1 -DPD_SCAN_MIN_CTAS=8
wait
for f in build/variants/scan_*.ptxas; do echo "$f: $(grep -A1 'k_scan_add' $f | grep -o 'Used [0-9]* registers' | head -1) $(grep -c 'error' $f) errors"; done
