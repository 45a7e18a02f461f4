#!/bin/bash
# round 2, trip U (1 GPU): speculative record chain + table entry reuse in the K=1 inflate: parity, decode bench, the bench lines of record,
# and the command line on a 50 000-contig BAM (the reference needs 664-1276 s there: not run)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_bgzf_gpu.py tests/test_cli_gpu.py -q -n 8 ) > gpurun_out/u_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest_gpu.log
tail -n 3 gpurun_out/u_pytest_gpu.log | cut -c1-300
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
( timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 5 ) > gpurun_out/u_bgzf_A.log 2>&1
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 5 ) > gpurun_out/u_bgzf_B.log 2>&1
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/u_bench_ref.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/u_bench.log 2>&1
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/u_bench_B.log 2>&1
# many contigs: 50 000 x 20 kb, 2 M reads (uniform generator writes sorted reads over all contigs)
( python tools/make_many_contigs.py /tmp/many.sam 50000 20000 2000000 && oracle/_ref/bamgen -I /tmp/many.sam -o /tmp/many.bam ) > gpurun_out/u_many_gen.log 2>&1
for rep in 1 2; do ( /usr/bin/time -f "wall %e s" env PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/many.bam -o /tmp/many_out -t 16 ) >> gpurun_out/u_many_cli.log 2>&1; done
( /usr/bin/time -f "wall %e s" env PANDEPTH_GPU_DECODE=0 PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/many.bam -o /tmp/many_out_h -t 16 ) >> gpurun_out/u_many_cli_host.log 2>&1
cmp <(zcat /tmp/many_out.chr.stat.gz) <(zcat /tmp/many_out_h.chr.stat.gz) && echo "many contigs: device decode == host decode" | tee -a gpurun_out/u_many_cli.log
grep -h "wall\|decode+push\|bgzf feed\|CUDA context" gpurun_out/u_many_cli.log gpurun_out/u_many_cli_host.log | cut -c1-250
python - <<'PY'
import json
for f in ("u_bench_ref", "u_bench", "u_bench_B"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l); e = d.get("e2e", {})
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", (d.get("roofline") or {}).get("frac"), "e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_vs_reference_rows"), "ref", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
for f in ("u_bgzf_A", "u_bgzf_B"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1]); r = d["resident"]; fd = d["feed"]
        print(f, "resident", round(r["ms_per_step"], 2), "ms", round(r["Mreads_per_s"], 1), "Mreads/s | feed (CRC)", round(fd["s_best"] * 1e3, 2), "ms", round(fd["Mreads_per_s"], 1))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
