#!/usr/bin/env python
"""A pan-genome-shaped test input: N contigs of L bases, R sorted reads of 100M spread over them, written as SAM text
(bamgen -I turns it into BAM + .bai).  usage: make_many_contigs.py out.sam [N=50000] [L=20000] [R=2000000]"""
import sys
import numpy as np
out = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 50000; L = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
R = int(sys.argv[4]) if len(sys.argv) > 4 else 2000000
rng = np.random.default_rng(5)
tid = np.sort(rng.integers(0, N, R)); pos = rng.integers(1, L - 100, R)
order = np.lexsort((pos, tid)); tid = tid[order]; pos = pos[order]
with open(out, "w") as fh:
    fh.write("@HD\tVN:1.6\tSO:coordinate\n")
    fh.write("".join(f"@SQ\tSN:s{i}\tLN:{L}\n" for i in range(N)))
    step = 200000
    for a in range(0, R, step):
        fh.write("".join(f"r\t0\ts{t}\t{p}\t60\t100M\t*\t0\t0\t*\t*\n" for t, p in zip(tid[a:a + step].tolist(), pos[a:a + step].tolist())))
