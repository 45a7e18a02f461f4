#!/bin/bash
# round 2, final trip (1 GPU): smoke, the whole GPU suite, both bench arms (flavour A full size, flavour B sample), kernel benches, the
# ncu evidence (launch lists + full captures of the dominant kernels), CLI wall times.  Outputs feed tools/summarize_profiles.py r02.
mkdir -p gpurun_out
( timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/z_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/z_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q -n 8 --durations=5 ) > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest_gpu.log
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/z_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/z_bench_ref.log
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/z_bench.log 2>&1; echo "rc=$?" >> gpurun_out/z_bench.log
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/z_bench_B.log 2>&1; echo "rc=$?" >> gpurun_out/z_bench_B.log
( timeout 600 python tools/bench_variants.py --steps 10 ) > gpurun_out/z_variants.log 2>&1
( timeout 300 python tools/bench_delta_path.py ) > gpurun_out/z_delta_path.log 2>&1
( timeout 300 python tools/bench_windows.py ) > gpurun_out/z_windows.log 2>&1
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
( timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 5 ) > gpurun_out/z_bgzf_A.log 2>&1
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 5 ) > gpurun_out/z_bgzf_B.log 2>&1
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_(cut|tile|fill|reduce|scan|delta|bgzf|filter)" -c 200 --csv --log-file gpurun_out/launches.csv $B ) > gpurun_out/z_ncu_launch.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bgzf_A.csv python tools/bench_bgzf.py /tmp/cA.bam --steps 1 --no-feed ) > gpurun_out/z_ncu_launch_A.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bgzf_B.csv python tools/bench_bgzf.py /tmp/cB.bam --steps 1 --no-feed ) > gpurun_out/z_ncu_launch_B.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 9 -c 1 -f -o gpurun_out/prof_r02_tile $B ) > gpurun_out/z_ncu_tile.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 12 -c 1 -f -o gpurun_out/prof_r02_bgzf_A python tools/bench_bgzf.py /tmp/cA.bam --steps 1 --no-feed ) > gpurun_out/z_ncu_bgzf_A.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 12 -c 1 -f -o gpurun_out/prof_r02_bgzf_B python tools/bench_bgzf.py /tmp/cB.bam --steps 1 --no-feed ) > gpurun_out/z_ncu_bgzf_B.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_accumulate -s 3 -c 1 -f -o gpurun_out/prof_r02_tile_mixed python tools/bench_variants.py --steps 1 --only mixed ) > gpurun_out/z_ncu_tile_mixed.log 2>&1
# compute-sanitizer over the kernels this round changed: tile kernel with the multi-op queue and the filter bitmap (smoke: mixed CIGARs,
# flag + mapq arrays), the K-decoder inflate kernels, the striped scatter
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke ) > gpurun_out/z_san_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/z_san_memcheck_smoke.log
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke ) > gpurun_out/z_san_racecheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/z_san_racecheck_smoke.log
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bgzf_gpu.py -x -q -k "fixtures or literal_dominated or few_blocks" ) > gpurun_out/z_san_memcheck_bgzf.log 2>&1; echo "rc=$?" >> gpurun_out/z_san_memcheck_bgzf.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_bgzf_gpu.py -x -q -k "literal_dominated" ) > gpurun_out/z_san_racecheck_bgzf.log 2>&1; echo "rc=$?" >> gpurun_out/z_san_racecheck_bgzf.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -x -q -k "striped or (fixtures and mix)" ) > gpurun_out/z_san_racecheck_engine.log 2>&1; echo "rc=$?" >> gpurun_out/z_san_racecheck_engine.log
( REPS=2 timeout 900 bash tools/bench_cli.sh 3000 A ) > gpurun_out/z_cli_3000_A.log 2>&1
( REPS=2 THREADS=16 timeout 900 bash tools/bench_cli.sh 3000 A -w 100 ) > gpurun_out/z_cli_3000_A_w100.log 2>&1
( REPS=2 THREADS=16 timeout 900 bash tools/bench_cli.sh 100 B ) > gpurun_out/z_cli_100_B.log 2>&1
for f in z_smoke z_pytest_gpu z_san_memcheck_smoke z_san_racecheck_smoke z_san_memcheck_bgzf z_san_racecheck_bgzf z_san_racecheck_engine z_variants z_delta_path z_windows z_cli_3000_A z_cli_3000_A_w100 z_cli_100_B; do echo "== $f"; tail -n 14 gpurun_out/$f.log | cut -c1-900; done
python - <<'PY'
import json
for f in ("z_bench_ref", "z_bench", "z_bench_B"):
    try:
        l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1]; d = json.loads(l); e = d.get("e2e", {})
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "frac", (d.get("roofline") or {}).get("frac"), "e2e", e.get("value"), e.get("ms_per_step"), e.get("parity_vs_reference_rows"), "ref", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
