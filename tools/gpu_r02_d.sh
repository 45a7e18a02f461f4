#!/bin/bash
# round 2, trip D (2 GPUs): cfg5 strong scaling over NCCL (range sharding, all-reduce in the timed region), the CLI on two
# devices, and -- on one GPU -- the step with small cut CTAs / bulk reduce-add default and the decode loop.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/d_env.log 2>&1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/d_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/d_bench_n2.log
B=/tmp/pdb200_bam_3000mb_200x_A/s.bam
if [ -f $B ]; then
  for rep in 1 2; do
    t0=$(date +%s.%N); PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i $B -o /tmp/d_one -t 16 2> gpurun_out/d_cli_one.err; t1=$(date +%s.%N)
    t2=$(date +%s.%N); PANDEPTH_TIMING=1 PANDEPTH_DEVICES=0,1 pandepth_b200/bin/pandepth -i $B -o /tmp/d_two -t 16 2> gpurun_out/d_cli_two.err; t3=$(date +%s.%N)
    python -c "print('cfg5 BAM (4 G reads): one device %.3f s, two devices %.3f s' % ($t1-$t0, $t3-$t2))" >> gpurun_out/d_cli.log
  done
  cmp <(zcat /tmp/d_one.chr.stat.gz) <(zcat /tmp/d_two.chr.stat.gz) && echo "identical: one vs two devices" >> gpurun_out/d_cli.log
  zcat /tmp/d_two.chr.stat.gz | tail -3 >> gpurun_out/d_cli.log
  grep timing gpurun_out/d_cli_two.err >> gpurun_out/d_cli.log
fi
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline ) > gpurun_out/d_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/d_bench_n1.log
oracle/_ref/bamgen -P -q -o /tmp/cB.bam -c $(python -c "print(','.join(f'chr{i+1}:{6666666}' for i in range(9)))") -n 20000000 -s 7 -t 16
( timeout 600 python tools/bench_bgzf.py /tmp/cB.bam --steps 3 ) > gpurun_out/d_bgzf_B.log 2>&1; echo "rc=$?" >> gpurun_out/d_bgzf_B.log
for f in d_bench_n2 d_cli d_bench_n1 d_bgzf_B; do echo "== $f"; tail -n 16 gpurun_out/$f.log | cut -c1-3000; done
