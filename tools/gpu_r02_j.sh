#!/bin/bash
# round 2, trip J (1 GPU): striped scatter (parity + stripe size), the tile kernel on filter / mixed-CIGAR inputs, register cap of the
# one-decoder-per-warp kernel again (two warm-up steps, 5 timed).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_engine_gpu.py -q -n 8 -k "striped or fixtures or random_mixed" ) > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
for mb in off 32 64 128; do
  if [ $mb = off ]; then export PANDEPTH_SCATTER_STRIPE_MIN_READS=4000000000; else unset PANDEPTH_SCATTER_STRIPE_MIN_READS; export PANDEPTH_SCATTER_STRIPE_MB=$mb; fi
  ( timeout 300 python tools/bench_delta_path.py ) > gpurun_out/j_delta_$mb.log 2>&1
  echo "delta stripe=$mb: $(tail -n 1 gpurun_out/j_delta_$mb.log | cut -c1-400)"
done
unset PANDEPTH_SCATTER_STRIPE_MIN_READS PANDEPTH_SCATTER_STRIPE_MB
( timeout 600 python tools/bench_variants.py --steps 10 ) > gpurun_out/j_variants.log 2>&1; tail -n 3 gpurun_out/j_variants.log | cut -c1-1800
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
run() { # tag lib K bam extra
  ( PANDEPTH_B200_LIB=$2 PANDEPTH_BGZF_K=$3 timeout 300 python tools/bench_bgzf.py $4 --steps 5 $5 ) > gpurun_out/j_$1.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/j_$1.log").read().strip().splitlines()[-1]); r = d["resident"]; f = d.get("feed")
    print("%-14s resident %8.2f ms %7.1f Mreads/s %6.1f GB/s inflated (%d x %.0f MB)" % ("$1", r["ms_per_step"], r["Mreads_per_s"], r["inflated_GBps"], r["pieces"], r["piece_mb"]),
          ("| feed %8.2f ms %7.1f Mreads/s %5.1f GB/s file (%d pieces)" % (f["s_best"] * 1e3, f["Mreads_per_s"], f["file_GBps"], f["pieces"])) if f else "")
except Exception as e:
    print("$1 FAILED", e)
PY
}
L=pandepth_b200/libpandepth_b200.so; V=build/variants
for rep in 1 2; do
run A_c4_$rep $L "" /tmp/cA.bam --no-feed
run A_c5_$rep $V/libpd_k1c5.so "" /tmp/cA.bam --no-feed
run A_c6_$rep $V/libpd_k1c6.so "" /tmp/cA.bam --no-feed
done
run A_c5_p18 $V/libpd_k1c5.so "" /tmp/cA.bam "--no-feed --piece-mb 18"
run A_c5_p9 $V/libpd_k1c5.so "" /tmp/cA.bam "--no-feed --piece-mb 9"
run B_K1_c4 $L 1 /tmp/cB.bam --no-feed
run B_K1_c5 $V/libpd_k1c5.so 1 /tmp/cB.bam --no-feed
run B_K1_c6 $V/libpd_k1c6.so 1 /tmp/cB.bam --no-feed
run B_K2 $L "" /tmp/cB.bam --no-feed
tail -n 5 gpurun_out/j_pytest.log
