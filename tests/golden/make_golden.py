#!/usr/bin/env python
"""
Regenerates tests/golden/ from the REFERENCE BINARY (oracle/_ref/pandepth_ref, built from
/root/reference by `make -C oracle`).  Run in the build container only:

    python tests/golden/make_golden.py

Writes the input fixtures (SAM text, BED/GFF/GTF) and `expected.json`, which maps every case to the
exact decompressed bytes the reference printed.  The SAM fixtures are run twice where stated: as SAM
text (no index -> the reference's 18-bit SiteInfo path) and converted to BAM+.bai with
oracle/_ref/bamgen -I (indexed -> uint32 windowed path, PanDepth.cpp:618-801).
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "pandepth_ref")
BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")

K_SAM = (
    "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:k1\tLN:301\n@SQ\tSN:k2\tLN:1\n@SQ\tSN:k3\tLN:250\n"
    "a\t0\tk1\t1\t60\t50M\t*\t0\t0\t*\t*\n"
    "b\t16\tk1\t11\t0\t5S20M10D20M\t*\t0\t0\t*\t*\n"
    "c\t0\tk1\t21\t30\t10=5X5I10M100N10M\t*\t0\t0\t*\t*\n"
    "d\t1024\tk1\t30\t60\t50M\t*\t0\t0\t*\t*\n"
    "e\t2048\tk1\t252\t60\t50M\t*\t0\t0\t*\t*\n"
    "f\t0\tk3\t100\t5\t30M2D30M\t*\t0\t0\t*\t*\n"
)
K_BED3 = "k1\t1\t60\nk1\t40\t301\nk3\t90\t170\n#comment\nk9\t1\t5\nk1\t70\t60\n"
K_BED4 = "k1\t1\t60\tx\nk1\t40\t301\ty\nk3\t90\t170\tx\nk1\t100\t120\tx\n"
# region that starts on the last base of a contig: exercises the MeMEnd clip in PanDepth.cpp:299-308
K_BEDEDGE = "k1\t10\t20\nk1\t301\t301\nk3\t250\t250\n"
K_GFF = (
    "##gff-version 3\n"
    "k1\tsrc\tgene\t1\t300\t.\t+\t.\tID=g1\n"
    "k1\tsrc\tmRNA\t1\t300\t.\t+\t.\tID=t1;Parent=g1\n"
    "k1\tsrc\tCDS\t5\t60\t.\t+\t0\tID=c1;Parent=t1\n"
    "k1\tsrc\tCDS\t100\t180\t.\t+\t0\tID=c2;Parent=t1\n"
    "k1\tsrc\tCDS\t20\t40\t.\t+\t0\tID=c3;Parent=t2,t9\n"
    "k1\tsrc\texon\t1\t300\t.\t+\t.\tID=e1;Parent=t1\n"
    "k3\tsrc\tCDS\t80\t170\t.\t-\t0\tID=c4;Parent=t3\n"
    "k3\tsrc\tCDS\t5\t10\t.\t-\t0\tParent=t1\n"
    "zz\tsrc\tCDS\t5\t10\t.\t-\t0\tID=c5;Parent=t4\n"
)
K_GTF = (
    'k1\tsrc\tCDS\t5\t60\t.\t+\t0\tgene_id "g1"; transcript_id "t1";\n'
    'k1\tsrc\tCDS\t100\t180\t.\t+\t0\tgene_id "g1"; transcript_id "t1";\n'
    'k1\tsrc\texon\t1\t300\t.\t+\t.\tgene_id "g1"; transcript_id "t1";\n'
    'k3\tsrc\tCDS\t80\t170\t.\t-\t0\tgene_id "g2"; transcript_id "t3";\n'
)

# PAF fixtures (paf_main, PanDepth.cpp:852-2024): cg:Z: walks, tag-less lines ([start-1, end)), tp:A:S lines, swapped
# start/end, space-separated tags, a 1-bp target (vanishes from the report), an op the walk ignores (S, I)
K_PAF = (
    "r1\t500\t0\t100\t+\tp1\t400\t10\t110\t100\t100\t60\ttp:A:P\tcg:Z:100M\n"
    "r2\t500\t0\t80\t-\tp1\t400\t50\t150\t70\t90\t30\ttp:A:P\tcg:Z:20M5I10M10D30=5X20N5M\n"
    "r3\t500\t0\t50\t+\tp1\t400\t300\t350\t50\t50\t60\ttp:A:S\tcg:Z:50M\n"
    "r4\t300\t0\t100\t+\tp2\t300\t21\t120\t99\t99\t5\ttp:A:P\n"
    "r5\t300\t0\t50\t-\tp2\t300\t200\t150\t50\t50\t40\ttp:A:P\n"
    "r6\t100\t0\t60\t+\tp4\t250\t99\t161\t60\t62\t0\ttp:A:P\tcg:Z:30M2D30M\n"
    "r7\t100\t0\t10\t+\tp4\t250\t5\t15\t10\t10\t60 tp:A:P cg:Z:5S10M\n"
    "r8\t10\t0\t1\t+\tp3\t1\t0\t1\t1\t1\t60\ttp:A:P\tcg:Z:1M\n"
    "\n"
    "r9\t500\t0\t100\t+\tp1\t400\t10\t110\t100\t100\t12\ttp:A:P\tcg:Z:100M\n"
)
K2_PAF = (
    "s1\t500\t0\t100\t+\tp2\t300\t100\t220\t100\t120\t60\ttp:A:P\tcg:Z:60M20D40M\n"
    "s2\t500\t0\t45\t+\tzz\t100\t5\t50\t45\t45\t60\ttp:A:P\tcg:Z:45M\n"
    "s3\t500\t0\t100\t+\tp1\t400\t301\t400\t99\t99\t60\ttp:A:P\n"
)
K_PAF_BED = "p1\t1\t60\tx\np1\t40\t400\ty\np2\t90\t170\tx\np4\t100\t120\tz\n"



def rand_paf(seed=5, n=600):
    """Seeded PAF soup; every walk stays within len+400 (the reference's arrays carry +500 of slack, :1520)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    names = [("t%d" % i, int(rng.integers(3000, 40000))) for i in range(5)] + [("tiny", 1)]
    lines = []
    for i in range(n):
        nm, ln = names[int(rng.integers(0, len(names) - 1))] if i % 97 else names[-1]
        s = int(rng.integers(1, ln)) if ln > 1 else 0
        e = int(rng.integers(s, ln + 1))
        tags = ["tp:A:" + "PSI"[int(rng.integers(0, 3))]]
        if rng.random() < 0.7 or ln == 1:          # (a tag-less line at start 0 makes the reference write depth[-1])
            budget, ops = ln + 400 - s, ""
            for _ in range(int(rng.integers(1, 9))):
                k = int(rng.integers(1, 400)); op = "M=XDNISH"[int(rng.integers(0, 8))]
                if op in "M=XDN":
                    if k > budget:
                        continue
                    budget -= k
                ops += "%d%s" % (k, op)
            tags.append("cg:Z:" + (ops or "1I"))
        if rng.random() < 0.3:
            tags.insert(0, "NM:i:3")
        if rng.random() < 0.1:
            s, e = e, s
        sep = " " if rng.random() < 0.1 else "\t"
        lines.append("\t".join(["q%d" % i, "1000", "0", "500", "+-"[i & 1], nm, str(ln), str(s), str(e), "10", "20",
                                str(int(rng.integers(0, 61)))]) + sep + sep.join(tags))
    return "\n".join(lines) + "\n"


MIX_CONTIGS = "cA:100001,cB:1001,cC:1,cD:2,cE:300"
BIG_CONTIGS = "cF:10000001,cG:150"
MIX_BED = "cA\t1\t5000\ncA\t4000\t9000\ncA\t50000\t50100\ncA\t99000\t100001\ncB\t1001\t1001\ncB\t5\t900\ncE\t10\t290\ncD\t1\t2\n"
BIG_BED = "cF\t1\t150\ncF\t5000100\t5000400\ncF\t9999000\t10000001\ncG\t1\t150\n"


def run_ref(inp, args, workdir, tag):
    pre = os.path.join(workdir, tag)
    cmd = [REF, "-i", inp, "-o", pre] + args
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    res = {}
    for fn in sorted(os.listdir(workdir)):
        if fn.startswith(tag + ".") and fn.endswith(".gz"):
            with gzip.open(os.path.join(workdir, fn), "rt") as fh:
                res[fn[len(tag):]] = fh.read()
    return res


def compact(text, limit=65536):
    """Outputs above `limit` bytes are stored as sha256 + size + head/tail instead of verbatim."""
    if len(text) <= limit:
        return text
    import hashlib
    return {"sha256": hashlib.sha256(text.encode()).hexdigest(), "n_bytes": len(text), "head": text[:2048], "tail": text[-512:]}


def main():
    if not (os.path.exists(REF) and os.path.exists(BAMGEN)):
        sys.exit("build oracle/_ref first: make -C oracle")
    w = lambda name, text: open(os.path.join(HERE, name), "w").write(text)
    w("k.sam", K_SAM); w("k.bed3", K_BED3); w("k.bed4", K_BED4); w("k.bededge", K_BEDEDGE)
    w("k.paf", K_PAF); w("k2.paf", K2_PAF); w("k.pafbed", K_PAF_BED); w("rand.paf", rand_paf())
    w("k.gff", K_GFF); w("k.gtf", K_GTF); w("mix.bed", MIX_BED); w("big.bed", BIG_BED)
    subprocess.run([BAMGEN, "-o", os.path.join(HERE, "mix.sam"), "-c", MIX_CONTIGS, "-n", "3000", "-m", "mixed",
                    "-s", "20260925", "-S"], check=True)
    subprocess.run([BAMGEN, "-o", os.path.join(HERE, "big.sam"), "-c", BIG_CONTIGS, "-n", "1500", "-m", "mixed",
                    "-s", "7", "-S"], check=True)

    cases = []
    def add(name, fixture, args, kinds=("sam", "bam")):
        for k in kinds:
            cases.append({"name": f"{name}.{k}", "fixture": fixture, "kind": k, "args": args})

    g = lambda f: os.path.join(HERE, f)
    add("k.default", "k.sam", [])
    add("k.w100", "k.sam", ["-w", "100"])
    add("k.w150", "k.sam", ["-w", "150"])
    add("k.w7", "k.sam", ["-w", "7"])
    add("k.q10d2", "k.sam", ["-q", "10", "-d", "2"])
    add("k.x0", "k.sam", ["-x", "0"])
    add("k.x16", "k.sam", ["-x", "16", "-q", "1"])
    add("k.bed3", "k.sam", ["-b", "@k.bed3"])
    add("k.bed4", "k.sam", ["-b", "@k.bed4"])
    add("k.bededge", "k.sam", ["-b", "@k.bededge"])
    add("k.gff", "k.sam", ["-g", "@k.gff"])
    add("k.gffexon", "k.sam", ["-g", "@k.gff", "-f", "exon"])
    add("k.gtf", "k.sam", ["-g", "@k.gtf"])
    add("k.a", "k.sam", ["-a"])
    add("k.a.w100", "k.sam", ["-a", "-w", "100"])
    add("k.s", "k.sam", ["-s"], kinds=("bam",))
    # -a together with regions: the per-site dump covers the contigs that carry regions, from the reads the region path keeps
    add("k.a.bed4", "k.sam", ["-a", "-b", "@k.bed4"])
    add("k.a.gff", "k.sam", ["-a", "-g", "@k.gff"])
    add("mix.a.bed", "mix.sam", ["-a", "-b", "@mix.bed"])
    # mode-6 widths that take the five-load lanes of k_reduce_windows (129..149) and the first mode-5 width
    add("k.w129", "k.sam", ["-w", "129"])
    add("mix.w129", "mix.sam", ["-w", "129"])
    add("mix.w149d2", "mix.sam", ["-w", "149", "-d", "2"])
    add("mix.default", "mix.sam", [])
    add("mix.w100", "mix.sam", ["-w", "100"])
    add("mix.w1000", "mix.sam", ["-w", "1000"])
    add("mix.q20d3x3844", "mix.sam", ["-q", "20", "-d", "3", "-x", "3844"])
    add("mix.bed", "mix.sam", ["-b", "@mix.bed"])
    add("mix.bed.d2", "mix.sam", ["-b", "@mix.bed", "-d", "2"])
    add("mix.w151", "mix.sam", ["-w", "151", "-d", "2"])
    add("big.default", "big.sam", [])
    add("big.w1000", "big.sam", ["-w", "1000"])
    add("big.bed", "big.sam", ["-b", "@big.bed"])
    # multi-file mode (-i x.list, BamList_main PanDepth.cpp:2029-3410): depth of all files summed in one 18-bit array
    add("list2k.default", "list:k.sam,k.sam", [])
    add("list2k.w100", "list:k.sam,k.sam", ["-w", "100"])
    add("list2k.bed4", "list:k.sam,k.sam", ["-b", "@k.bed4"])
    add("list2k.gff", "list:k.sam,k.sam", ["-g", "@k.gff"])
    add("list2k.a", "list:k.sam,k.sam", ["-a"])
    # (the reference's multi-file arrays carry only +100 bases of slack (:2689) instead of +500, so fixtures whose reads overhang
    # a contig end by more -- mix.sam, big.sam -- corrupt its heap: it segfaults, aborts or hangs.  Lists use k.sam only.)
    add("list3k.default", "list:k.sam,k.sam,k.sam", ["-q", "5", "-d", "2"])
    add("list3k.w7", "list:k.sam,k.sam,k.sam", ["-w", "7"])
    add("list3k.w150", "list:k.sam,k.sam,k.sam", ["-w", "150"])
    add("list2stack.default", "list:@stack,@stack", [])          # 2 x 262149 reads: the 18-bit sum wraps across files
    # PAF input (paf_main :852-2024): text lines, 18-bit arrays; .paf.gz is looked through (:3469-3473)
    add("paf.default", "k.paf", [], kinds=("paf", "pafgz"))
    add("paf.w100", "k.paf", ["-w", "100"], kinds=("paf",))
    add("paf.w150", "k.paf", ["-w", "150"], kinds=("paf",))
    add("paf.q10d2", "k.paf", ["-q", "10", "-d", "2"], kinds=("paf",))
    add("paf.x0", "k.paf", ["-x", "0"], kinds=("paf",))
    add("paf.bed4", "k.paf", ["-b", "@k.pafbed"], kinds=("paf",))
    add("paf.a", "k.paf", ["-a"], kinds=("paf",))
    add("pafrand.default", "rand.paf", [], kinds=("paf",))
    add("pafrand.w100d2", "rand.paf", ["-w", "100", "-d", "2"], kinds=("paf",))
    add("pafrand.x0q20", "rand.paf", ["-x", "0", "-q", "20"], kinds=("paf",))
    add("pafrand.w1000", "rand.paf", ["-w", "1000", "-x", "0"], kinds=("paf",))
    add("paflist.default", "list:k.paf,k2.paf", [], kinds=("paf",))
    add("paflist.w7", "list:k.paf,k2.paf", ["-w", "7", "-x", "0"], kinds=("paf",))
    # 18-bit wrap vector (SURVEY App. C `w.sam`): 262149 stacked 10M reads at pos 100 of a 1 kb contig
    add("stack.default", "@stack", [])
    add("stack.w100", "@stack", ["-w", "100"])
    add("stack.w200", "@stack", ["-w", "200"])

    out = {}
    tmp = tempfile.mkdtemp(prefix="pdgold")
    try:
        stack = os.path.join(tmp, "stack.sam")
        with open(stack, "w") as fh:
            fh.write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:c\tLN:1000\n")
            fh.write("r\t0\tc\t100\t60\t10M\t*\t0\t0\t*\t*\n" * 262149)
        bams = {}
        def to_bam(sam):
            if sam not in bams:
                b = os.path.join(tmp, os.path.basename(sam) + ".bam")
                subprocess.run([BAMGEN, "-I", sam, "-o", b], check=True)
                bams[sam] = b
            return bams[sam]
        for c in cases:
            if c["fixture"].startswith("list:"):
                members = [stack if x == "@stack" else g(x) for x in c["fixture"][5:].split(",")]
                if c["kind"] == "bam":
                    members = [to_bam(m) for m in members]
                inp = os.path.join(tmp, c["name"] + ".list")
                with open(inp, "w") as fh:
                    fh.write("\n".join(members) + "\n")
                args = [g(a[1:]) if a.startswith("@") else a for a in c["args"]]
                res = run_ref(inp, args, tmp, c["name"])
                res = {k: compact(v) for k, v in res.items()}
                out[c["name"]] = {"fixture": c["fixture"], "kind": c["kind"], "args": c["args"], "outputs": res}
                print(c["name"], {k: (len(v) if isinstance(v, str) else v["n_bytes"]) for k, v in res.items()})
                continue
            sam = stack if c["fixture"] == "@stack" else g(c["fixture"])
            if c["kind"] == "pafgz":
                inp = os.path.join(tmp, os.path.basename(sam) + ".gz")
                with open(sam, "rb") as fi, gzip.open(inp, "wb") as fo:
                    fo.write(fi.read())
            elif c["kind"] == "bam":
                if sam not in bams:
                    b = os.path.join(tmp, os.path.basename(sam) + ".bam")
                    subprocess.run([BAMGEN, "-I", sam, "-o", b], check=True)
                    bams[sam] = b
                inp = bams[sam]
            else:
                inp = sam
            args = [g(a[1:]) if a.startswith("@") else a for a in c["args"]]
            res = run_ref(inp, args, tmp, c["name"])
            res = {k: compact(v) for k, v in res.items()}
            out[c["name"]] = {"fixture": c["fixture"], "kind": c["kind"], "args": c["args"], "outputs": res}
            print(c["name"], {k: (len(v) if isinstance(v, str) else v["n_bytes"]) for k, v in res.items()})
    finally:
        shutil.rmtree(tmp)
    with open(os.path.join(HERE, "expected.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
