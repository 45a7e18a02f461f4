"""H2D bandwidth probe: torch pinned copy vs the engine's host push path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pandepth_b200 as pd
dev = torch.device("cuda", 0)
_stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(_stream)   # non-default stream shared with the engine
n = 256 << 20
h = torch.empty(n, dtype=torch.int32, pin_memory=True); h.fill_(3)
d = torch.empty(n, dtype=torch.int32, device=dev)
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"torch pinned H2D 1 GiB: {n*4/dt/1e9:.1f} GB/s")
hp = torch.empty(n, dtype=torch.int32); hp.fill_(3)
t0 = time.perf_counter(); d.copy_(hp); torch.cuda.synchronize(); print(f"torch pageable H2D: {n*4/(time.perf_counter()-t0)/1e9:.1f} GB/s")
L = 300_000_000; nr = 100_000_000
pos = torch.sort(torch.randint(0, L - 150, (nr,), device=dev, dtype=torch.int32)).values
hpos = torch.empty(nr, dtype=torch.int32, pin_memory=True); hpos.copy_(pos)
hcig = torch.full((nr,), 150 << 4, dtype=torch.int32).pin_memory()
eng = pd.Engine([L])
for rep in range(3):
    eng.reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.push_contig(0, hpos.numpy(), hcig.numpy(), flags=pd.PD_SORTED | pd.PD_ASYNC); eng.sync()
    dt = time.perf_counter() - t0
    print(f"engine host push 100M reads: {dt*1e3:.1f} ms  {nr*8/dt/1e9:.1f} GB/s  {nr/dt/1e6:.0f} Mreads/s")
