#!/bin/bash
# round 2, trip V (1 GPU): the bucketed unsorted path (expand -> bucket -> tile): parity, tools/bench_delta_path.py with and without it,
# the command line on a 50 000-contig BAM, and the bench lines of record for the final code
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_bgzf_gpu.py -q -n 8 ) > gpurun_out/v_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest_gpu.log
tail -n 12 gpurun_out/v_pytest_gpu.log | cut -c1-400
( timeout 300 python tools/bench_delta_path.py ) > gpurun_out/v_delta_bucket.log 2>&1; echo "bucketed: $(tail -n 1 gpurun_out/v_delta_bucket.log | cut -c1-700)"
( PANDEPTH_BUCKET_MIN_READS=0 timeout 300 python tools/bench_delta_path.py ) > gpurun_out/v_delta_scatter.log 2>&1; echo "scatter:  $(tail -n 1 gpurun_out/v_delta_scatter.log | cut -c1-500)"
( python tools/make_many_contigs.py /tmp/many.sam 50000 20000 2000000 && oracle/_ref/bamgen -I /tmp/many.sam -o /tmp/many.bam && ls -la /tmp/many.bam ) > gpurun_out/v_many_gen.log 2>&1
for mode in dev dev host; do
  t0=$(date +%s.%N)
  if [ $mode = host ]; then PANDEPTH_GPU_DECODE=0 PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/many.bam -o /tmp/many_h -t 16 > gpurun_out/v_many_cli_host.log 2>&1
  else PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i /tmp/many.bam -o /tmp/many_d -t 16 > gpurun_out/v_many_cli_dev.log 2>&1; fi
  t1=$(date +%s.%N); echo "many contigs ($mode decode): wall $(python -c "print('%.3f' % ($t1-$t0))") s" | tee -a gpurun_out/v_many_wall.log
done
if cmp -s <(zcat /tmp/many_d.chr.stat.gz) <(zcat /tmp/many_h.chr.stat.gz) && [ -s /tmp/many_d.chr.stat.gz ]; then echo "many contigs: device decode == host decode ($(zcat /tmp/many_d.chr.stat.gz | wc -l) lines)"; else echo "many contigs: outputs DIFFER or missing"; fi | tee -a gpurun_out/v_many_wall.log
grep -h "decode+push\|bgzf feed\|CUDA context\|reduce+report" gpurun_out/v_many_cli_dev.log gpurun_out/v_many_cli_host.log | cut -c1-250
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/v_bench_ref.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/v_bench.log 2>&1
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/v_bench_B.log 2>&1
python - <<'PY'
import json
for f in ("v_bench_ref", "v_bench", "v_bench_B"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l); e = d.get("e2e", {})
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", (d.get("roofline") or {}).get("frac"), "e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_vs_reference_rows"), "ref", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
