#!/usr/bin/env python
"""Randomised campaign over the warp-emulator checkers (tests/cpp/*_emu_check.cpp): every kernel of the library, run from its
source on the CPU, on random shapes biased to the edges (contig lengths and read counts around 1, 32, 256, 512, 2048, 4096, 8192,
2^19; ranges that start inside a contig; empty / one-read / piled-up batches; window widths around every branch of the window
reducer; random BAMs from oracle/_ref/bamgen through the inflate kernels with K = 1/2/4/8 decoders per warp).

    python tools/fuzz_emu_cpu.py [--seconds 600] [--seed 1] [--only tile|bucket|scatter|reduce|windows|bgzf|corrupt] [--sanitize]

--sanitize builds the checkers with -fsanitize=address,undefined; `--only corrupt` feeds the inflate / record-walk kernels bit-flipped BAMs
(any checker verdict is fine there; a sanitizer report, a crash or a hang is a failure).

Prints every failing command line (nothing is expected to fail) and a summary.  The pytest suite runs fixed cases of the same
checkers (tests/test_{tile,bucket,reduce,bgzf}_emu_cpu.py); this is the long-running version.
"""
import argparse, os, random, subprocess, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")


def build(tmp, sanitize=False):
    exe = {}
    san = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"] if sanitize else ["-O2"]
    for name, libs in (("tile", []), ("bucket", []), ("reduce", []), ("bgzf", ["-lz"])):
        exe[name] = os.path.join(tmp, name + "_emu_check")
        subprocess.run(["g++", *san, "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"), "-I" + os.path.join(ROOT, "tests", "cpp"),
                        os.path.join(ROOT, "tests", "cpp", name + "_emu_check.cpp"), *libs, "-o", exe[name]], check=True)
    return exe


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--seconds", type=float, default=600); ap.add_argument("--seed", type=int, default=1); ap.add_argument("--only", default="")
    ap.add_argument("--sanitize", action="store_true")
    a = ap.parse_args()
    rnd = random.Random(a.seed)
    tmp = tempfile.mkdtemp(prefix="pd_fuzz_emu_")
    exe = build(tmp, a.sanitize)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    kinds = ["tile", "tile", "bucket", "scatter", "reduce", "windows"] + (["bgzf", "bgzf"] if os.path.exists(BAMGEN) else [])
    if a.only:
        kinds = [a.only]
    t0 = time.time(); runs = {}; fails = 0
    rr = rnd.randrange
    while time.time() - t0 < a.seconds:
        kind = rnd.choice(kinds); seed = rr(1, 10**9)
        if kind == "tile":
            cmd = [exe["tile"], rnd.choice([1, 255, 256, 257, 2047, 2048, 2049, 4095, 4096, rr(1, 200000), rr(1, 30000)]),
                   rnd.choice([0, 1, 31, 32, 33, 511, 512, 513, rr(0, 40000)]), seed, rr(2), rr(2), rnd.choice([0, 0, rr(1, 10**6), 2048, 256, 255])]
        elif kind == "bucket":
            cmd = [exe["bucket"], rnd.choice([1, 255, 256, 257, 524287, 524288, 524289, rr(1, 1500000)]), rnd.choice([1, 2, 255, 256, 8191, 8192, 8193, rr(1, 60000)]), seed,
                   rnd.choice([0, 0, rr(1, 10**6)])]
        elif kind == "scatter":
            cmd = [exe["reduce"], "scatter", rnd.choice([1, 4095, 4096, 8191, 8192, rr(1, 120000)]), rr(1, 5000), seed, rnd.choice([0, rr(1, 50000)]), rr(3)]
        elif kind == "reduce":
            cmd = [exe["reduce"], "reduce", rnd.choice([1, 2047, 2048, 8191, 8192, 8193, rr(1, 150000)]), rr(1, 80), seed, rnd.choice([1, 1, 2, 5, 100]), rnd.choice(["0xffffffff", "0x3ffff"])]
        elif kind == "windows":
            ln = rnd.choice([1, 31, 32, 33, 127, 128, 640, rr(1, 40000)])
            cmd = [exe["reduce"], "windows", ln, rnd.choice([1, 2, 31, 32, 36, 100, 128, 132, 160, 161, 512, 513, rr(1, 3000)]), rnd.choice([0, 0, 4, 64, rr(0, ln + 2)]),
                   rnd.choice([1, 1, 2, 7]), rnd.choice(["0xffffffff", "0x3ffff"]), seed]
        elif kind == "corrupt":
            bam = os.path.join(tmp, "f.bam"); src = os.path.join(tmp, "src%d.bam" % rr(3))
            if not os.path.exists(src):
                gen = [BAMGEN, "-o", src, "-c", "c1:200000,c2:1,c3:150000,c4:900,c5:30000", "-n", str(rnd.choice([80, 500, 12000])), "-s", str(rr(1, 10**6)), "-l", "100"]
                gen += rnd.choice([["-m", "mixed"], ["-q"], []])
                subprocess.run(gen, check=True, capture_output=True)
            d = bytearray(open(src, "rb").read())
            for _ in range(rnd.choice([1, 1, 2, 5])):
                d[rr(200, len(d) - 28)] ^= 1 << rr(8)               # behind the BAM header block, in front of the EOF block
            open(bam, "wb").write(d); gen = ["bit flips in", src]
            cmd = [exe["bgzf"], bam, rnd.choice([1, 2, 3]), 1796, -1, rnd.choice([1, 2]), 0, rnd.choice([1, 2, 4, 8])]
        else:
            bam = os.path.join(tmp, "f.bam")
            contigs = ",".join(f"c{i}:{rnd.choice([1, 2, 100, 900, 30000, 200000, 1000000])}" for i in range(rnd.choice([1, 1, 2, 5, 40])))
            gen = [BAMGEN, "-o", bam, "-c", contigs, "-n", str(rnd.choice([1, 10, 300, 3000, 20000])), "-s", str(rr(1, 10**6)), "-l", str(rnd.choice([1, 30, 100, 150, 250, 1000, 5000]))]
            gen += ["-m", "mixed"] if rnd.random() < 0.6 else []
            gen += ["-q"] if rnd.random() < 0.4 else []
            if subprocess.run(gen, capture_output=True).returncode != 0:
                continue
            cmd = [exe["bgzf"], bam, rnd.choice([1, 2, 5, 9]), rnd.choice([0, 1796, 4, 1024]), rnd.choice([-1, 0, 20, 60]), rnd.choice([1, 2, 3, 4]), rnd.choice([0, 0, 1, 3]), rnd.choice([1, 1, 2, 4, 8])]
        cmd = [str(c) for c in cmd]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
            out = [l for l in r.stdout.strip().splitlines() if "checking" not in l]
            clean = "Sanitizer" not in r.stderr and "runtime error" not in r.stderr and r.returncode >= 0
            ok = clean and (kind == "corrupt" or (r.returncode == 0 and out and all(l.endswith(": ok") or "nothing to check" in l for l in out)))
            tail = r.stdout[-600:] + r.stderr[-300:]
        except subprocess.TimeoutExpired:
            ok, tail = False, "timeout"
        runs[kind] = runs.get(kind, 0) + 1
        if not ok:
            fails += 1
            if kind in ("bgzf", "corrupt"):
                keep = os.path.join(tmp, f"fail{fails}.bam"); os.replace(cmd[1], keep); cmd[1] = keep; print("generated by:", " ".join(gen))
            print("FAIL", " ".join(cmd), "\n", tail, flush=True)
    print(f"{sum(runs.values())} runs {runs}, {fails} failures, {time.time() - t0:.0f} s, seed {a.seed}")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
