#!/usr/bin/env python
"""Latency of the summary all-reduce (a few hundred int64) as bench.py issues it: torchrun --nproc-per-node N tools/nccl_probe.py"""
import os, time, numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
x = np.arange(618, dtype=np.int64)
def once(sync_before):
    t = torch.from_numpy(x.copy()).to(dev)
    if sync_before: dist.barrier()
    t0 = time.perf_counter(); dist.all_reduce(t); r = t.cpu(); return time.perf_counter() - t0
for _ in range(20): once(False)
a = [once(False) for _ in range(200)]; b = [once(True) for _ in range(50)]
d = torch.ones(618, dtype=torch.int64, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): dist.all_reduce(d)
torch.cuda.synchronize(); c = (time.perf_counter() - t0) / 200
if rank == 0:
    print("all_reduce(618 x i64) + .cpu(): median %.1f us, p90 %.1f us; after a barrier: median %.1f us; back-to-back on device: %.1f us" %
          (1e6 * np.median(a), 1e6 * np.percentile(a, 90), 1e6 * np.median(b), 1e6 * c))
dist.destroy_process_group()
