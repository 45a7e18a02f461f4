#!/bin/bash
# round 2, trip G (1 GPU): the warp-uniform K-decoders-per-warp inflate (K = 1/2/4/8) -- parity on the device, decode throughput on
# flavour A and B, one ncu capture of the K=4 kernel on B; process start-up probe (CUDA context).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_bgzf_gpu.py -q -n 8 -x ) > gpurun_out/g_pytest_bgzf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest_bgzf.log
oracle/_ref/bamgen -P -o /tmp/cA.bam -c $(python -c "print(','.join(f'chr{i+1}:{33333333}' for i in range(9)))") -n 100000000 -s 7 -t 16
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
for K in 1 2 4 8; do ( PANDEPTH_BGZF_K=$K timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 3 ) > gpurun_out/g_bgzf_B_K$K.log 2>&1; done
for K in 1 4; do ( PANDEPTH_BGZF_K=$K timeout 600 python tools/bench_bgzf.py /tmp/cA.bam --steps 3 ) > gpurun_out/g_bgzf_A_K$K.log 2>&1; done
( PANDEPTH_BGZF_K=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 4 -c 1 -f -o gpurun_out/prof_r02g_bgzf_B_K4 python tools/bench_bgzf.py /tmp/cB.bam --steps 1 ) > gpurun_out/g_ncu_bgzf_B_K4.log 2>&1
( PANDEPTH_BGZF_K=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_decode -s 4 -c 1 -f -o gpurun_out/prof_r02g_bgzf_B_K1 python tools/bench_bgzf.py /tmp/cB.bam --steps 1 ) > gpurun_out/g_ncu_bgzf_B_K1.log 2>&1
( nvidia-smi -q | grep -i -m3 "persistence"; for i in 1 2 3; do echo "-- all GPUs visible"; build/ctx_probe; echo "-- CUDA_VISIBLE_DEVICES=0"; CUDA_VISIBLE_DEVICES=0 build/ctx_probe; done ) > gpurun_out/g_ctx.log 2>&1
( timeout 900 python bench.py --flavour B --bam-mb 100 --steps 3 --warmup 3 --no-e2e-arrays ) > gpurun_out/g_bench_B.log 2>&1; echo "rc=$?" >> gpurun_out/g_bench_B.log
for f in g_pytest_bgzf g_bgzf_B_K1 g_bgzf_B_K2 g_bgzf_B_K4 g_bgzf_B_K8 g_bgzf_A_K1 g_bgzf_A_K4 g_ctx g_bench_B; do echo "== $f"; tail -n 30 gpurun_out/$f.log | cut -c1-1500; done
