#!/bin/bash
# round 2, trip H (1 GPU): decode pipeline depth (3 vs 6 chunks in flight) and 9-bit first-level tables for the K-decoder kernels
# (more resident warps), resident mode of tools/bench_bgzf.py on flavour B and A.
mkdir -p gpurun_out
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
run() { # tag lib K bam extra
  ( PANDEPTH_B200_LIB=$2 PANDEPTH_BGZF_K=$3 timeout 300 python tools/bench_bgzf.py $4 --steps 3 --no-feed $5 ) > gpurun_out/h_$1.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/h_$1.log").read().strip().splitlines()[-1]); r = d["resident"]
    print("%-22s %8.2f ms  %7.1f Mreads/s  %6.1f GB/s inflated  pieces %d x %.0f MB  decode-kernel sum %.1f ms" % ("$1", r["ms_per_step"], r["Mreads_per_s"], r["inflated_GBps"], r["pieces"], r["piece_mb"], r["kernels_ms_per_step"].get("k_bgzf_decode", 0)))
except Exception as e:
    print("$1 FAILED", e)
PY
}
L=pandepth_b200/libpandepth_b200.so; V=build/variants
for K in 1 2; do run B_d3_K$K $L $K /tmp/cB.bam; done
for K in 1 2 4; do run B_d6_K$K $V/libpd_d6.so $K /tmp/cB.bam; done
for K in 2 4; do run B_d6l9_K$K $V/libpd_d6l9.so $K /tmp/cB.bam; done
for K in 2 4; do run B_d6l9c8_K$K $V/libpd_d6l9c8.so $K /tmp/cB.bam; done
run B_d6_K1_p48 $V/libpd_d6.so 1 /tmp/cB.bam "--piece-mb 48"
run B_d6_K2_p192 $V/libpd_d6l9.so 2 /tmp/cB.bam "--piece-mb 192"
run A_d3_K1 $L 1 /tmp/cA.bam
run A_d6_K1 $V/libpd_d6.so 1 /tmp/cA.bam
run A_d6_K1_p9 $V/libpd_d6.so 1 /tmp/cA.bam "--piece-mb 9"
run A_d6_K1_p36 $V/libpd_d6.so 1 /tmp/cA.bam "--piece-mb 36"
