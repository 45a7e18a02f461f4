#!/usr/bin/env python
"""k_tile_accumulate A/B (GPU box): one cfg2 contig (333,333,333 bases, 111,111,111 sorted 150M reads, inputs resident in HBM).
   PANDEPTH_B200_LIB=<variant .so> python tools/bench_tile.py [--steps K]
Prints one JSON line: ms of the tile kernel for a FIRST pass (every segment written for the first time: plain / bulk store) and
for a SECOND pass over the same reads without a reset (every segment already holds depth: load+add+store vs bulk reduce-add),
achieved GB/s on the algorithmic bytes (8 B/read + 4 B/base; the second pass moves 8 B/base when it re-reads depth)."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=10); ap.add_argument("--reads", type=int, default=111_111_111)
    ap.add_argument("--len", type=int, default=333_333_333); ap.add_argument("--hint", type=int, default=1)
    a = ap.parse_args()
    import torch
    import pandepth_b200 as pd
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(7)
    pos = torch.sort(torch.randint(0, a.len - 149, (a.reads,), device=dev, dtype=torch.int32, generator=g)).values.contiguous()
    cig = torch.full((a.reads,), (150 << 4), device=dev, dtype=torch.int32)
    eng = pd.Engine(np.array([a.len], np.uint32))
    if a.hint:
        eng.set_reduce_hint(1, 0xFFFFFFFF)
    F = pd.PD_SORTED | pd.PD_DEVICE | pd.PD_READY
    res = {}
    for name, passes in (("first_write", 1), ("add_pass", 2)):
        for _ in range(3):
            eng.reset(); [eng.push_contig(0, pos.data_ptr(), cig.data_ptr(), flags=F, n=a.reads) for _ in range(passes)]; eng.sync()
        tot = 0.0; n = 0
        for _ in range(a.steps):
            eng.reset()
            if passes == 2:
                eng.push_contig(0, pos.data_ptr(), cig.data_ptr(), flags=F, n=a.reads); eng.sync()
            eng.profile_enable(True)
            eng.push_contig(0, pos.data_ptr(), cig.data_ptr(), flags=F, n=a.reads); eng.sync()
            p = eng.profile(); eng.profile_enable(False)
            tot += p["k_tile_accumulate"]["ms"]; n += p["k_tile_accumulate"]["launches"]
        ms = tot / n
        alg = 8 * a.reads + 4 * a.len
        res[name] = {"ms": ms, "alg_GBps": alg / ms / 1e6, "frac_of_6550": alg / ms / 1e6 / 6550.1}
    cov, tot = eng.reduce([0], [0], [a.len])
    assert int(tot[0]) == 2 * 150 * a.reads
    print(json.dumps({"lib": os.path.basename(pd.lib_path()), "hint": a.hint, **res}))


if __name__ == "__main__":
    main()
