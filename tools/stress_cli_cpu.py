#!/usr/bin/env python
"""Randomised differential campaign for the HOST program (no GPU): pandepth_b200/bin/pandepth, with the CUDA library replaced by
the test stand-in (tests/stand_in/pd_standin.cpp, built into a temp dir), against the reference binary oracle/_ref/pandepth_ref
on random inputs (BAM / CRAM / SAM, single file or list), random region files and random option sets.  Any difference in the
decompressed outputs is printed with the command line that reproduces it.
PDSTRESS_REAL_ENGINE=1 (on a GPU box) runs the same campaign against the real CUDA library instead of the stand-in.
usage: python tools/stress_cli_cpu.py [seconds | n=<seeds>] [first_seed]"""
import gzip, os, random, subprocess, sys, tempfile, time, shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "pandepth_b200", "bin", "pandepth")
REF = os.path.join(ROOT, "oracle", "_ref", "pandepth_ref")
GEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")
ORA = os.path.join(ROOT, "oracle", "_ref")
n_seeds = int(sys.argv[1][2:]) if len(sys.argv) > 1 and sys.argv[1].startswith("n=") else None
budget = 1e9 if n_seeds else (float(sys.argv[1]) if len(sys.argv) > 1 else 300.0)
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000

work = tempfile.mkdtemp(prefix="pdstress")
so_dir = os.path.join(work, "standin"); os.makedirs(so_dir)
subprocess.run(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", os.path.join(ROOT, "tests", "stand_in", "pd_standin.cpp"), "-L" + ORA, "-lpd_oracle",
                "-Wl,-rpath," + ORA, "-Wl,-soname,libpandepth_b200.so", "-o", os.path.join(so_dir, "libpandepth_b200.so")], check=True)
ENV = dict(os.environ) if os.environ.get("PDSTRESS_REAL_ENGINE") else dict(os.environ, LD_LIBRARY_PATH=so_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


def run(binary, inp, args, tag, env=None):
    pre = os.path.join(work, tag)
    try:
        r = subprocess.run([binary, "-i", inp, "-o", pre] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, env=env)
        rc = r.returncode
    except subprocess.TimeoutExpired:
        rc = -999
    outs = {}
    for fn in sorted(os.listdir(work)):
        if fn.startswith(tag + ".") and fn.endswith(".gz"):
            try:
                with gzip.open(os.path.join(work, fn), "rt") as fh:
                    outs[fn[len(tag):]] = fh.read()
            except (EOFError, OSError):
                outs[fn[len(tag):]] = "<truncated>"
            if not os.environ.get("PDSTRESS_KEEP"): os.unlink(os.path.join(work, fn))
    return rc, outs


t_end = time.time() + budget
seed, n_cmp, n_bad = seed0, 0, 0
n_empty, by_kind = 0, {}
while time.time() < t_end and (n_seeds is None or seed < seed0 + n_seeds):
    rng = random.Random(seed)
    n_ctg = rng.choice([1, 2, 3, 6, 12])
    lens = []
    for _ in range(n_ctg):
        k = rng.random()
        lens.append(rng.choice([1, 2, 3, 149, 150, 151, 2048, 4095, 4096, 4097]) if k < 0.25 else
                    rng.randint(200, 400000) if k < 0.9 else rng.randint(10_000_000, 10_400_000))
    names = [f"c{i}" for i in range(n_ctg)]
    spec = ",".join(f"{n}:{l}" for n, l in zip(names, lens))
    reads = rng.choice([1, 50, 5000, 60000])
    kind = rng.choice(["bam", "bam", "csi", "cram", "sam", "list", "paf"])
    seq = ["-q"] if rng.random() < 0.3 else []
    inputs = []
    if kind == "paf":
        # PAF lines (paf_main, :1536-1623): target columns 6-9, mapq column 12, optional cg:Z: walk.  Kept inside the contig: the
        # reference walks past the end of its arrays otherwise (undefined there, clipped here).
        lens = [max(l, 200) for l in lens]
        path = os.path.join(work, f"in{seed}.paf" + (".gz" if rng.random() < 0.3 else ""))
        lines = []
        for i in range(min(reads, 5000)):
            t = rng.randrange(n_ctg); ln = lens[t]
            a = rng.randint(1, ln - 1); z = rng.randint(a, ln)
            tags = ["tp:A:" + rng.choice("PSI")]
            if rng.random() < 0.7:
                ops, room = [], z - a
                for _ in range(rng.randint(1, 8)):
                    op = rng.choice("M=XDNISH"); n = rng.randint(1, 300)
                    if op in "M=XDN":
                        if n > room: continue
                        room -= n
                    ops.append(f"{n}{op}")
                if ops: tags.append("cg:Z:" + "".join(ops))
            if rng.random() < 0.3: tags.insert(0, "NM:i:3")
            lines.append("\t".join([f"q{i}", "1000", "0", "500", "+-"[i & 1], names[t], str(ln), str(a), str(z), "10", "20", str(rng.randint(0, 60))] + tags))
        # every target appears (the reference takes the target list from the file itself, :918-941)
        text = "\n".join(lines) + "\n"
        with (gzip.open(path, "wt") if path.endswith(".gz") else open(path, "w")) as fh:
            fh.write(text)
        inputs.append(path)
    for part in range(0 if kind == "paf" else (2 if kind == "list" else 1)):
        ext = {"bam": "bam", "csi": "bam", "cram": "cram", "sam": "sam", "list": rng.choice(["bam", "sam"])}[kind]
        path = os.path.join(work, f"in{seed}_{part}.{ext}")
        cmd = [GEN, "-o", path, "-c", spec, "-n", str(reads), "-m", rng.choice(["mixed", "mixed", "uniform"]), "-s", str(seed * 7 + part)] + seq
        if ext == "sam":
            cmd.append("-S")
        if kind == "csi":
            cmd.append("-X")      # .csi index only: htslib iterators instead of the raw scanner
        if kind == "list":
            cmd.append("-O")      # BamList_main has no slack behind its arrays: an overhanging read lands in another contig's depth
        subprocess.run(cmd, check=True)
        if ext == "sam" and kind == "sam" and rng.random() < 0.25:
            # a header that says SO:coordinate over records that are not in order: the reference's no-index cursor then decides
            # which reads count; the host restates it (Feeder::RegionCursor)
            lines = open(path).read().splitlines(keepends=True)
            body = [x for x in lines if not x.startswith("@")]
            rng.shuffle(body)
            if rng.random() < 0.5: body.sort(key=lambda x: x.split("\t")[2])
            with open(path, "w") as fh:
                fh.writelines([x for x in lines if x.startswith("@")] + body)
        inputs.append(path)
    if kind == "list":
        inp = os.path.join(work, f"in{seed}.list")
        with open(inp, "w") as fh:
            fh.write("\n".join(inputs) + "\n")
    else:
        inp = inputs[0]
    bed = os.path.join(work, f"r{seed}.bed"); gff = os.path.join(work, f"r{seed}.gff")
    bed4 = rng.random() < 0.5
    gtf = os.path.join(work, f"r{seed}.gtf")
    with open(bed, "w") as fb, open(gff, "w") as fg, open(gtf, "w") as ft:
        fg.write("##gff-version 3\n")
        for k in range(rng.choice([1, 5, 60])):
            t = rng.randrange(n_ctg + 1)
            nm = names[t] if t < n_ctg else "nowhere"
            ln = lens[t] if t < n_ctg else 1000
            s = rng.randint(1, max(1, ln)); e = min(ln, s + rng.randint(0, 3000))     # inside the contig: past its end the reference reads its slack (undefined)
            fb.write(f"{nm}\t{s}\t{e}" + (f"\tg{k // 2}" if bed4 else "") + "\n")
            feat = "CDS" if k % 4 else "exon"
            fg.write(f"{nm}\tsrc\t{feat}\t{s}\t{e}\t.\t+\t0\tID=x{k};Parent=t{k // 3}\n")
            ft.write(f'{nm}\tsrc\t{feat}\t{s}\t{e}\t.\t+\t0\tgene_id "g{k // 3}"; transcript_id "t{k // 3}";\n')
    if rng.random() < 0.3:                                   # region files may be gzip-compressed (igzstream in the reference)
        with open(bed, "rb") as fi, gzip.open(bed + ".gz", "wb") as fo: fo.write(fi.read())
        bed = bed + ".gz"
    if rng.random() < 0.3: gff = gtf
    big = sum(lens) > 2_000_000
    for trial in range(6):
        opts = []
        r = rng.random()
        if r < 0.3: opts += ["-w", str(rng.choice([1, 2, 7, 33, 100, 128, 129, 149, 150, 151, 160, 161, 500, 1000, 2000]))]
        elif r < 0.45: opts += ["-b", bed]
        elif r < 0.6: opts += ["-g", gff] + (["-f", "exon"] if rng.random() < 0.3 else [])
        if rng.random() < 0.3: opts += ["-d", str(rng.choice([1, 2, 5]))]
        if rng.random() < 0.3: opts += ["-q", str(rng.choice([0, 1, 20, 60]))]
        if rng.random() < 0.3: opts += ["-x", str(rng.choice([0, 4, 1796, 3844, 256]))]
        if rng.random() < 0.25: opts += ["-s"]
        # -a with regions: the reference's indexed path and its no-index path disagree with each other; the no-index
        # path (region cursor) is restated exactly, so there the per-site dump is compared too
        noindex = "-s" in opts or kind == "sam" or bool(os.environ.get("PDSTRESS_A_ALL"))
        if rng.random() < 0.2 and not big and (noindex or not ("-b" in opts or "-g" in opts)): opts += ["-a"]
        if "-w" in opts and int(opts[opts.index("-w") + 1]) < 7 and sum(lens) > 300_000: continue
        opts += ["-t", str(rng.choice([1, 3, 8]))]
        rc_r, ref = run(REF, inp, opts, f"ref{seed}_{trial}")
        rc_o, our = run(CLI, inp, opts, f"our{seed}_{trial}", ENV)
        if rc_r < 0:
            print(f"note: the reference crashed or hung (rc {rc_r}) seed={seed} kind={kind} opts={opts}; ours rc {rc_o}", flush=True)
            continue
        n_cmp += 1
        by_kind[kind] = by_kind.get(kind, 0) + 1
        if not ref: n_empty += 1
        if ref != our or (rc_r == 0) != (rc_o == 0):
            n_bad += 1
            diff = [k for k in set(ref) | set(our) if ref.get(k) != our.get(k)]
            print(f"DIFF seed={seed} kind={kind} spec={spec} reads={reads} seq={seq} opts={opts} rc={rc_r}/{rc_o} files={diff}", flush=True)
    if not os.environ.get("PDSTRESS_KEEP"):
        for p in inputs:
            for q in (p, p + ".bai", p + ".crai"):
                if os.path.exists(q): os.unlink(q)
    seed += 1
    if os.environ.get("PDSTRESS_ONE"): break
print(f"seeds {seed0}..{seed - 1}: {n_cmp} comparisons, {n_bad} differences; by input kind {by_kind}; runs without any output file {n_empty}")
if os.environ.get("PDSTRESS_KEEP"): print("kept", work)
else: shutil.rmtree(work, ignore_errors=True)
sys.exit(1 if n_bad else 0)
