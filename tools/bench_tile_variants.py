#!/usr/bin/env python
"""k_tile_accumulate on the batches real BAMs produce (GPU box): one cfg2 contig (333,333,333 bases, 111,111,111 sorted reads,
inputs resident in HBM), three variants, one JSON line each with its own roofline entry:
  uniform      every read 150M, no offsets, no filter arrays          <false,false>   alg bytes 8 R + 4 G
  mixed        30 % multi-op CIGARs (soft clip, insertion, deletion)   <true,false>    4 R (pos) + 4 (R+1) (offsets) + 4 W (words) + 4 G
  mixed+filter + flag (u16) / mapq (u8) arrays, 10 % of the reads fail <true,true>     ... + 3 R
Depth is checked against a torch restatement on a sample window."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import pandepth_b200 as pd
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    R, G = 111_111_111, 333_333_333
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(11)
    pos = torch.sort(torch.randint(0, G - 160, (R,), device=dev, dtype=torch.int32, generator=g)).values.contiguous()
    kind = torch.randint(0, 10, (R,), device=dev, dtype=torch.int32, generator=g)
    # words per read: 150M | 20S130M | 75M1I74M | 70M2D80M
    nw = torch.ones(R, device=dev, dtype=torch.int32); nw[kind == 7] = 2; nw[kind >= 8] = 3
    off = torch.zeros(R + 1, device=dev, dtype=torch.int32); off[1:] = torch.cumsum(nw, 0)
    W = int(off[-1].item())
    cig = torch.zeros(W, device=dev, dtype=torch.int32)
    o = off[:-1].long()
    M = lambda n: (n << 4) | 0
    cig[o[kind < 7]] = M(150)
    i7 = o[kind == 7]; cig[i7] = (20 << 4) | 4; cig[i7 + 1] = M(130)
    i8 = o[kind == 8]; cig[i8] = M(75); cig[i8 + 1] = (1 << 4) | 1; cig[i8 + 2] = M(74)
    i9 = o[kind == 9]; cig[i9] = M(70); cig[i9 + 1] = (2 << 4) | 2; cig[i9 + 2] = M(80)
    flag = torch.zeros(R, device=dev, dtype=torch.int16); flag[torch.randint(0, 20, (R,), device=dev, generator=g) == 0] = 1024
    mapq = torch.full((R,), 60, device=dev, dtype=torch.uint8); mapq[torch.randint(0, 20, (R,), device=dev, generator=g) == 0] = 3
    cig_u = torch.full((R,), M(150), device=dev, dtype=torch.int32)
    torch.cuda.synchronize()
    peak = 6550.1
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    F = pd.PD_SORTED | pd.PD_DEVICE
    for name, kw, alg in (("uniform", dict(cigar=cig_u.data_ptr()), 8 * R + 4 * G),
                          ("mixed", dict(cigar=cig.data_ptr(), cigar_off=off.data_ptr()), 4 * R + 4 * (R + 1) + 4 * W + 4 * G),
                          ("mixed+filter", dict(cigar=cig.data_ptr(), cigar_off=off.data_ptr(), flag=flag.data_ptr(), mapq=mapq.data_ptr()), 4 * R + 4 * (R + 1) + 4 * W + 3 * R + 4 * G)):
        eng = pd.Engine(np.array([G], np.uint32), xflags=1796, min_mapq=10)
        eng.set_reduce_hint(1, 0xFFFFFFFF)
        tot = 0.0; n = 0
        for it in range(steps + 3):
            eng.reset()
            eng.profile_enable(True)
            eng.push_contig(0, pos.data_ptr(), kw["cigar"], flag=kw.get("flag"), mapq=kw.get("mapq"), cigar_off=kw.get("cigar_off"), flags=F, n=R)
            eng.sync()
            p = eng.profile(); eng.profile_enable(False)
            if it >= 3:
                tot += p["k_tile_accumulate"]["ms"]; n += 1
        ms = tot / n
        cov, td = eng.reduce([0], [0], [G])
        # expected total depth
        if name == "uniform":
            want = 150 * R
        else:
            bases = torch.where(kind < 7, 150, torch.where(kind == 7, 130, torch.where(kind == 8, 149, 150))).long()
            keep = torch.ones(R, device=dev, dtype=torch.bool) if name == "mixed" else ((flag.int() & 1796) == 0) & (mapq.int() >= 10)
            want = int(bases[keep].sum().item())
        assert int(td[0]) == want, (name, int(td[0]), want)
        print(json.dumps({"variant": name, "kernel": "k_tile_accumulate", "ms": ms, "reads": R, "cigar_words": W if name != "uniform" else R,
                          "algorithmic_bytes": alg, "achieved_GBps": alg / ms / 1e6, "peak_GBps": peak, "frac": alg / ms / 1e6 / peak,
                          "total_depth_checked": want}))
        eng.close()


if __name__ == "__main__":
    main()
