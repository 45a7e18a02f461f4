"""Synthetic GFF3 for cfg4: 33,688 transcripts / 175,274 non-overlapping CDS lines (sizes of README.md:128), exon length
U[50,500], distributed in proportion to contig length, seed 11.  usage: make_gff.py out.gff name:len[,name:len...]"""
import sys, random
out, spec = sys.argv[1], sys.argv[2]
contigs = [(c.split(":")[0], int(c.split(":")[1])) for c in spec.split(",")]
G = sum(l for _, l in contigs)
rng = random.Random(11)
NT, NC = 33688, 175274
with open(out, "w") as fh:
    fh.write("##gff-version 3\n")
    tid = 0; cds = 0
    for ci, (nm, ln) in enumerate(contigs):
        nt = NT * ln // G if ci + 1 < len(contigs) else NT - tid
        if nt <= 0: continue
        span = ln // nt                      # each transcript gets its own slot of the contig
        for k in range(nt):
            base = k * span + 1
            ne = max(1, min(12, round(rng.gauss(NC / NT, 2))))
            p = base + rng.randint(0, max(0, span // 10))
            for e in range(ne):
                L = rng.randint(50, 500)
                if p + L >= base + span - 1 or p + L > ln: break
                fh.write(f"{nm}\tsyn\tCDS\t{p}\t{p + L - 1}\t.\t+\t0\tID=cds{cds};Parent=t{tid}\n")
                cds += 1
                p += L + rng.randint(80, 2000)
            tid += 1
print(tid, "transcripts", cds, "CDS")
