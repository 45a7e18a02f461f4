#!/bin/bash
# round 2, multi-GPU trip (gpurun --gpus N): cfg5 (3 Gb x 200x, 4e9 reads) strong-scaled over N ranks by genomic range, NCCL all-reduce inside
# the timed region; at N = 2 also the command line on two GPUs in one process (PANDEPTH_DEVICES) against its one-GPU output.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8 > gpurun_out/n${N}_env.log
nproc >> gpurun_out/n${N}_env.log; cat /sys/fs/cgroup/cpu.max >> gpurun_out/n${N}_env.log 2>/dev/null
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/n${N}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/n${N}_bench.log
python - <<PY
import json as _j
try:
    _d = _j.loads([x for x in open("gpurun_out/n${N}_bench.log") if x.startswith("{")][-1]); print("host phases", _d.get("host_phase_ms_per_step"))
except Exception as _e: print(_e)
import json
try:
    l = [x for x in open("gpurun_out/n${N}_bench.log") if x.startswith("{")][-1]; d = json.loads(l); e = d.get("e2e", {})
    print("N=${N} value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "kernels", {k: round(v, 2) for k, v in d["kernels_ms_per_step"].items()}, "e2e", e.get("value"), e.get("ms_per_step"), e.get("h2d_bytes_per_step"))
    print(d["config"]["workload"][:200]); print(e.get("input"), e.get("path"))
except Exception as ex:
    print("N=${N} FAILED", ex); print(open("gpurun_out/n${N}_bench.log").read()[-3000:])
PY
if [ "$N" = 2 ]; then
  BAM=/tmp/pdb200_bam_3000mb_200x_A/s.bam
  if [ -f $BAM ]; then
    ( PANDEPTH_TIMING=1 pandepth_b200/bin/pandepth -i $BAM -o /tmp/one -t 16 ) > gpurun_out/n2_cli_one.log 2>&1
    ( PANDEPTH_TIMING=1 PANDEPTH_DEVICES=0,1 pandepth_b200/bin/pandepth -i $BAM -o /tmp/two -t 16 ) > gpurun_out/n2_cli_two.log 2>&1
    if cmp -s <(zcat /tmp/one.chr.stat.gz) <(zcat /tmp/two.chr.stat.gz); then echo "CLI 2 GPUs: identical to 1 GPU"; else echo "CLI 2 GPUs: DIFFERENT"; fi | tee -a gpurun_out/n2_cli_two.log
    grep -h "decode+push\|CUDA context" gpurun_out/n2_cli_one.log gpurun_out/n2_cli_two.log
  fi
fi
