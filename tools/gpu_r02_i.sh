#!/bin/bash
# round 2, trip I (1 GPU): new decode defaults (6 chunks in flight, K=2 with 9-bit tables for literal-dominated chunks, bigger pieces),
# register cap of the one-decoder-per-warp kernel (4/5/6 CTAs per SM), BGZF parity tests, both bench arms on flavour B.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_bgzf_gpu.py -q -n 8 ) > gpurun_out/i_pytest_bgzf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest_bgzf.log
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
run() { # tag lib K bam extra
  ( PANDEPTH_B200_LIB=$2 PANDEPTH_BGZF_K=$3 timeout 300 python tools/bench_bgzf.py $4 --steps 3 $5 ) > gpurun_out/i_$1.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/i_$1.log").read().strip().splitlines()[-1]); r = d["resident"]; f = d.get("feed")
    print("%-14s resident %8.2f ms %7.1f Mreads/s %6.1f GB/s inflated (%d x %.0f MB)" % ("$1", r["ms_per_step"], r["Mreads_per_s"], r["inflated_GBps"], r["pieces"], r["piece_mb"]),
          ("| feed %8.2f ms %7.1f Mreads/s %5.1f GB/s file (%d pieces)" % (f["s_best"] * 1e3, f["Mreads_per_s"], f["file_GBps"], f["pieces"])) if f else "")
except Exception as e:
    print("$1 FAILED", e)
PY
}
L=pandepth_b200/libpandepth_b200.so; V=build/variants
run A_c4 $L "" /tmp/cA.bam
run A_c5 $V/libpd_k1c5.so "" /tmp/cA.bam --no-feed
run A_c6 $V/libpd_k1c6.so "" /tmp/cA.bam --no-feed
run B_K2 $L "" /tmp/cB.bam
run B_K1_c4 $L 1 /tmp/cB.bam --no-feed
run B_K1_c5 $V/libpd_k1c5.so 1 /tmp/cB.bam --no-feed
run B_K1_c6 $V/libpd_k1c6.so 1 /tmp/cB.bam --no-feed
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/i_bench_B.log 2>&1; echo "rc=$?" >> gpurun_out/i_bench_B.log
( timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e-arrays --no-cpu-baseline ) > gpurun_out/i_bench_A.log 2>&1; echo "rc=$?" >> gpurun_out/i_bench_A.log
for f in i_pytest_bgzf; do echo "== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-400; done
python - <<'PY'
import json
for f in ("i_bench_B", "i_bench_A"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l)
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "Mreads/s", round(d["e2e"]["ms_per_step"], 1), "ms; ref", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
