#!/usr/bin/env python
"""The accumulate kernel on the inputs real BAMs bring (GPU box): one cfg2 contig (333,333,333 bases, 111,111,111 sorted reads,
inputs resident in HBM), three variants of k_tile_accumulate, one JSON line:

  single_op   every read 150M, no filter arrays                 -> k_tile_accumulate<false,false>  (bench.py's `value` kernel)
  filter      150M + flag[] / mapq[] on the device, ~12 % of the reads fail `-x 1796 -q 20`         -> <false,true>
  mixed       ~30 % multi-op CIGARs (soft clips, insertions, deletions, short N skips, 2-5 words) with cigar_off[]  -> <true,false>
  mixed_filter  both                                                                                 -> <true,true>

Algorithmic bytes per launch (SURVEY 8(d) extended): 4 R (pos) + 4 W (CIGAR words) [+ 4 (R+1) offsets] [+ 3 R flag+mapq] + 4 G.
Each variant is checked: sum of depth == aligned bases of the reads that pass, computed with torch from the same arrays.
"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PEAK = 6550.1


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=10); ap.add_argument("--reads", type=int, default=111_111_111)
    ap.add_argument("--len", type=int, default=333_333_333); ap.add_argument("--only", default="", help="run one variant (for ncu captures)")
    a = ap.parse_args()
    import torch
    import pandepth_b200 as pd
    try:
        PK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", PEAK))
    except Exception:
        PK = PEAK
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(7)
    R, G = a.reads, a.len
    pos = torch.sort(torch.randint(0, G - 400, (R,), device=dev, dtype=torch.int32, generator=g)).values.contiguous()
    # ---- filter fields: 6 % duplicates (0x400), 3 % secondary (0x100), 4 % mapq < 20
    u = torch.rand(R, device=dev, generator=g)
    flag = torch.zeros(R, device=dev, dtype=torch.int16)
    flag[u < 0.06] = 0x400; flag[(u >= 0.06) & (u < 0.09)] = 0x100; flag[u >= 0.55] |= 16
    mapq = torch.full((R,), 60, device=dev, dtype=torch.uint8); mapq[(u >= 0.09) & (u < 0.13)] = 3
    passes = ((flag.to(torch.int32) & 1796) == 0) & (mapq >= 20)
    # ---- CIGARs: class per read; words built class by class
    c = torch.rand(R, device=dev, generator=g)
    M, I, D, N, S = 0, 1, 2, 3, 4
    def w(n, op): return (n << 4) | op
    classes = [  # (upper bound of c, words, aligned bases)
        (0.70, [w(150, M)], 150),
        (0.80, [w(20, S), w(130, M)], 130),
        (0.88, [w(75, M), w(1, D), w(75, M)], 150),
        (0.94, [w(60, M), w(2, I), w(88, M)], 148),
        (0.98, [w(50, M), w(180, N), w(100, M)], 150),
        (1.01, [w(10, S), w(40, M), w(1, I), w(50, M), w(3, D), w(49, M)], 139),
    ]
    nw = torch.zeros(R, device=dev, dtype=torch.int32); al = torch.zeros(R, device=dev, dtype=torch.int32); cls = torch.zeros(R, device=dev, dtype=torch.int8)
    lo = 0.0
    for k, (hi, words, bases) in enumerate(classes):
        m = (c >= lo) & (c < hi); nw[m] = len(words); al[m] = bases; cls[m] = k; lo = hi
    off = torch.zeros(R + 1, device=dev, dtype=torch.int32); off[1:] = torch.cumsum(nw, 0)
    W = int(off[-1].item())
    cig_mixed = torch.zeros(W, device=dev, dtype=torch.int32)
    for k, (hi, words, bases) in enumerate(classes):
        idx = off[:-1][cls == k].to(torch.int64)
        for j, word in enumerate(words):
            cig_mixed[idx + j] = word
    cig_single = torch.full((R,), w(150, M), device=dev, dtype=torch.int32)
    multi_frac = float((nw > 1).float().mean().item())
    torch.cuda.synchronize()

    F = pd.PD_SORTED | pd.PD_DEVICE | pd.PD_READY
    out = {"reads": R, "bases": G, "cigar_words_mixed": W, "multi_op_fraction": multi_frac, "peak_GBps": PK, "variants": {}}
    for name, use_mixed, use_filter in (("single_op", False, False), ("filter", False, True), ("mixed", True, False), ("mixed_filter", True, True)):
        if a.only and name != a.only:
            continue
        eng = pd.Engine(np.array([G], np.uint32), xflags=1796, min_mapq=20 if use_filter else -1)
        eng.set_reduce_hint(1, 0xFFFFFFFF)
        kw = dict(flags=F, n=R)
        if use_filter:
            kw.update(flag=flag.data_ptr(), mapq=mapq.data_ptr())
        cg = cig_mixed if use_mixed else cig_single
        if use_mixed:
            kw.update(cigar_off=off.data_ptr())
        for _ in range(3):
            eng.reset(); eng.push_contig(0, pos.data_ptr(), cg.data_ptr(), **kw); eng.sync()
        tot = 0.0; n = 0
        for _ in range(a.steps):
            eng.reset(); eng.profile_enable(True)
            eng.push_contig(0, pos.data_ptr(), cg.data_ptr(), **kw); eng.sync()
            p = eng.profile(); eng.profile_enable(False)
            tot += p["k_tile_accumulate"]["ms"]; n += p["k_tile_accumulate"]["launches"]
        ms = tot / n
        cov, td = eng.reduce([0], [0], [G])
        albases = (al if use_mixed else torch.full_like(al, 150)).to(torch.int64)
        want = int((albases * (passes if use_filter else torch.ones_like(passes))).sum().item())
        assert int(td[0]) == want, (name, int(td[0]), want)
        st = eng.stats()
        alg = 4 * R + 4 * (W if use_mixed else R) + (4 * (R + 1) if use_mixed else 0) + (3 * R if use_filter else 0) + 4 * G
        out["variants"][name] = {"ms": ms, "alg_bytes": alg, "alg_GBps": alg / ms / 1e6, "frac_of_peak": alg / ms / 1e6 / PK, "Mreads_per_s": R / ms / 1e3,
                                 "total_depth": int(td[0]), "delta_scan_needed": bool(st.get("scan_passes", 0))}
        eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
