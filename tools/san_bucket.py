#!/usr/bin/env python
"""One unsorted batch through the bucketed path (expand -> bucket -> tile), checked against the oracle: small enough to run under
compute-sanitizer (racecheck / memcheck of the shared-memory staging kernels).  PANDEPTH_BUCKET_MIN_READS=1 python tools/san_bucket.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PANDEPTH_BUCKET_MIN_READS", "1")
import pandepth_b200 as pd
from oracle import model
rng = np.random.default_rng(3)
L = 1_200_000                                   # three coarse buckets
rows = []
for p in rng.integers(0, L, 60_000).tolist():
    k = rng.integers(0, 6)
    cs = ("150M", "150M", f"{rng.integers(1, 80)}M{rng.integers(1, 3000)}N{rng.integers(1, 80)}M", f"{rng.integers(1, 40)}S{rng.integers(1, 120)}M",
          f"{rng.integers(1, 60)}={rng.integers(1, 9)}I{rng.integers(1, 60)}X{rng.integers(1, 30)}D{rng.integers(1, 60)}M", f"{rng.integers(1, 700)}M")[k]
    rows.append((int(rng.choice([0, 16, 1024])), 0, p + 1, int(rng.integers(0, 61)), cs))
rec = model.records_from_tuples(rows, ["c"])
tlen = np.array([L], np.uint32)
want = model.accumulate(tlen, rec, 1796, 10)
with pd.Engine(tlen, xflags=1796, min_mapq=10) as eng:
    eng.push_contig(0, rec.pos, rec.cigar, flag=rec.flag, mapq=rec.mapq, cigar_off=rec.cigar_off, flags=0)
    eng.push_contig(0, rec.pos[:20000].copy(), rec.cigar[:int(rec.cigar_off[20000])].copy(), flag=rec.flag[:20000].copy(), mapq=rec.mapq[:20000].copy(), cigar_off=rec.cigar_off[:20001].copy(), flags=0)
    eng.finalize()
    got = eng.depth(0)
    st = eng.stats()
sub = model.Records(rec.tid[:20000], rec.pos[:20000], rec.flag[:20000], rec.mapq[:20000], rec.cigar_off[:20001], rec.cigar[:int(rec.cigar_off[20000])])
want2 = want[0] + model.accumulate(tlen, sub, 1796, 10)[0]
assert np.array_equal(got, want2), "bucketed path differs from the oracle"
print("bucketed path ok:", rec.n, "+ 20000 reads,", st["kernel_launches"], "launches")
