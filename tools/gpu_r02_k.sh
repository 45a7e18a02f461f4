#!/bin/bash
# round 2, trip K (1 GPU): the bench contract with the new decode defaults, then piece-size / register-cap sweeps of the product feeder
# (tools/bench_bgzf.py feed mode = pdh_feed_bam) on the full-size cfg2 file and on the flavour-B bench file.
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e-arrays ) > gpurun_out/k_bench_A.log 2>&1; echo "rc=$?" >> gpurun_out/k_bench_A.log
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays --no-cpu-baseline ) > gpurun_out/k_bench_B.log 2>&1; echo "rc=$?" >> gpurun_out/k_bench_B.log
python - <<'PY'
import json
for f in ("k_bench_A", "k_bench_B"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l)
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 1), "Mreads/s", round(d["e2e"]["ms_per_step"], 1), "ms; ref", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
run() { # tag lib bam piece_mb mode
  ( PANDEPTH_B200_LIB=$2 timeout 300 python tools/bench_bgzf.py $3 --steps 4 $5 ${4:+--piece-mb $4} ) > gpurun_out/k_$1.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/k_$1.log").read().strip().splitlines()[-1]); r = d.get("resident"); f = d.get("feed")
    print("%-16s" % "$1", ("resident %8.2f ms %7.1f Mreads/s %6.1f GB/s inflated (%d x %.0f MB)" % (r["ms_per_step"], r["Mreads_per_s"], r["inflated_GBps"], r["pieces"], r["piece_mb"])) if r else "",
          ("| feed %8.2f ms %7.1f Mreads/s %5.1f GB/s file (%d pieces) all %s" % (f["s_best"] * 1e3, f["Mreads_per_s"], f["file_GBps"], f["pieces"], [round(x * 1e3, 1) for x in f["s_all"]])) if f else "")
except Exception as e:
    print("$1 FAILED", e)
PY
}
L=pandepth_b200/libpandepth_b200.so; V=build/variants
A=/tmp/pdb200_bam_3000mb_50x_A/s.bam; B=/tmp/pdb200_bam_100mb_50x_B/s.bam
for p in 9 18 37 74; do run Afeed_c6_p$p $L $A $p --no-resident; done
for p in 18 37; do run Afeed_c4_p$p $V/libpd_k1c4.so $A $p --no-resident; done
for p in 48 96 192; do run Bfeed_p$p $L $B $p --no-resident; done
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
for p in 9 18 37; do run Ares_c6_p$p $L /tmp/cA.bam $p --no-feed; run Ares_c4_p$p $V/libpd_k1c4.so /tmp/cA.bam $p --no-feed; done
