"""
GPU: the drop-in command line (pandepth_b200/bin/pandepth) against
  (1) the golden outputs the reference binary printed for every fixture/option set (tests/golden/expected.json), and
  (2) the reference binary itself (oracle/_ref/pandepth_ref, prebuilt, travels with the snapshot) on freshly generated
      mixed-CIGAR BAMs -- `diff <(zcat ref) <(zcat ours)` must be empty.
Nothing here reads /root/reference.
"""
import gzip
import json
import os
import shutil
import subprocess
import tempfile

import pytest

from tests.helpers import GOLD, ROOT, assert_matches

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "pandepth_b200", "bin", "pandepth")
REF = os.path.join(ROOT, "oracle", "_ref", "pandepth_ref")
BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")

with open(os.path.join(GOLD, "expected.json")) as _fh:
    _CASES = json.load(_fh)


@pytest.fixture(scope="module")
def workdir():
    d = tempfile.mkdtemp(prefix="pdcli")
    yield d
    shutil.rmtree(d, ignore_errors=True)


@pytest.fixture(scope="module")
def inputs(workdir):
    """SAM fixtures as-is, plus BAM+.bai conversions made with the prebuilt bamgen."""
    res = {}
    stack = os.path.join(workdir, "stack.sam")
    with open(stack, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:c\tLN:1000\n")
        fh.write("r\t0\tc\t100\t60\t10M\t*\t0\t0\t*\t*\n" * 262149)
    for fx in ("k.sam", "mix.sam", "big.sam", "@stack"):
        sam = stack if fx == "@stack" else os.path.join(GOLD, fx)
        bam = os.path.join(workdir, os.path.basename(sam) + ".bam")
        subprocess.run([BAMGEN, "-I", sam, "-o", bam], check=True)
        res[fx] = {"sam": sam, "bam": bam}
    return res


def _run(binary, inp, args, workdir, tag, env=None):
    pre = os.path.join(workdir, tag)
    r = subprocess.run([binary, "-i", inp, "-o", pre] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout, r.stderr)
    outs = {}
    for fn in sorted(os.listdir(workdir)):
        if fn.startswith(tag + ".") and fn.endswith(".gz"):
            with gzip.open(os.path.join(workdir, fn), "rt") as fh:
                outs[fn[len(tag):]] = fh.read()
    return outs, r.stdout


@pytest.mark.parametrize("case", sorted(_CASES))
def test_cli_matches_golden(case, inputs, workdir):
    c = _CASES[case]
    args = [os.path.join(GOLD, a[1:]) if a.startswith("@") else a for a in c["args"]]
    def member(m):
        if m.endswith(".paf"):                   # PAF text is read as it lies (paf_main, PanDepth.cpp:852)
            if c["kind"] != "pafgz":
                return os.path.join(GOLD, m)
            z = os.path.join(workdir, m + ".gz")
            with open(os.path.join(GOLD, m), "rb") as fi, gzip.open(z, "wb") as fo:
                fo.write(fi.read())
            return z
        return inputs[m][c["kind"]]
    if c["fixture"].startswith("list:"):
        inp = os.path.join(workdir, case + ".list")
        with open(inp, "w") as fh:
            fh.write("\n".join(member(m) for m in c["fixture"][5:].split(",")) + "\n")
    else:
        inp = member(c["fixture"])
    outs, _ = _run(CLI, inp, args + ["-t", "4"], workdir, "our_" + case)
    assert sorted(outs) == sorted(c["outputs"]), (sorted(outs), sorted(c["outputs"]))
    for k, want in c["outputs"].items():
        assert_matches(want, outs[k], f"{case}{k}")


OPTION_SETS = [
    [], ["-w", "100"], ["-w", "1000"], ["-w", "33"], ["-q", "20", "-d", "2"], ["-x", "0"], ["-x", "3844", "-q", "1"],
    ["-s"], ["-a"], ["-t", "1"], ["-t", "16", "-d", "5"], ["-w", "129"], ["-w", "149", "-d", "2"], ["-w", "150"],
]


@pytest.mark.parametrize("seed,contigs,reads", [
    (11, "a:300001,b:1,c:5000,d:70000,e:2,f:1001", 60000),
    (12, "chrA:2000001,chrB:999999", 120000),
    (13, "x:10000001,y:150", 40000),
])
def test_cli_matches_reference_binary_on_random_bams(seed, contigs, reads, workdir):
    bam = os.path.join(workdir, f"rand{seed}.bam")
    subprocess.run([BAMGEN, "-o", bam, "-c", contigs, "-n", str(reads), "-m", "mixed", "-s", str(seed)], check=True)
    names = [c.split(":")[0] for c in contigs.split(",")]
    lens = [int(c.split(":")[1]) for c in contigs.split(",")]
    # region files derived from the contig table (deterministic)
    bed = os.path.join(workdir, f"rand{seed}.bed")
    gff = os.path.join(workdir, f"rand{seed}.gff")
    with open(bed, "w") as fb, open(gff, "w") as fg:
        fg.write("##gff-version 3\n")
        g = 0
        for nm, ln in zip(names, lens):
            if ln < 200:
                fb.write(f"{nm}\t1\t{ln}\n")
                continue
            step = max(ln // 40, 150)
            for s in range(1, ln - 100, step):
                e = min(ln, s + step // 3)
                fb.write(f"{nm}\t{s}\t{e}\n")
                fg.write(f"{nm}\tsrc\tCDS\t{s}\t{e}\t.\t+\t0\tID=c{g};Parent=t{g // 3}\n")
                g += 1
            fb.write(f"{nm}\t{ln}\t{ln}\n")          # region on the last base: the MeMEnd clip quirk (PanDepth.cpp:299-308)
    sets = OPTION_SETS + [["-b", bed], ["-b", bed, "-d", "3"], ["-g", gff], ["-g", gff, "-s"], ["-b", bed, "-s"]]
    if seed != 11:
        # every option set runs against the reference on the CPU as well (tests/test_cli_cpu.py, host program + stand-in); on
        # the GPU the larger genomes keep the sets that differ in what they ask of the engine (each run pays a CUDA start-up)
        sets = [[], ["-w", "100"], ["-w", "1000"], ["-q", "20", "-d", "2"], ["-s"], ["-t", "16", "-d", "5"], ["-b", bed], ["-g", gff, "-s"], ["-w", "129"]]
    for k, opts in enumerate(sets):
        if "-a" in opts and sum(lens) > 5_000_000:
            continue                     # keep the suite quick: the per-site dump is one line per base
        # whole-genome modes decode the BAM on the device by default; every other option set also runs the host decode path
        env = dict(os.environ, PANDEPTH_GPU_DECODE="0") if k % 2 else None
        ref, _ = _run(REF, bam, opts + (["-t", "4"] if "-t" not in opts else []), workdir, f"ref{seed}_{k}")
        our, _ = _run(CLI, bam, opts + (["-t", "4"] if "-t" not in opts else []), workdir, f"our{seed}_{k}", env)
        assert sorted(ref) == sorted(our), (opts, sorted(ref), sorted(our))
        for name in ref:
            assert ref[name] == our[name], f"seed {seed} opts {opts}: {name} differs from the reference binary"


def test_cli_matches_reference_binary_on_cram(workdir):
    """CRAM + .crai input (htslib iterators, `hts_set_fai_filename` / required-fields options of PanDepth.cpp:637-643):
    same bytes as the reference binary, indexed and -s, per-chr / window / BED modes."""
    contigs = "a:300001,b:1,c:5000,d:70000,e:2,f:1001"
    cram = os.path.join(workdir, "rand21.cram")
    subprocess.run([BAMGEN, "-o", cram, "-c", contigs, "-n", "60000", "-m", "mixed", "-s", "21"], check=True)
    assert os.path.exists(cram + ".crai")
    bed = os.path.join(workdir, "rand21.bed")
    with open(bed, "w") as fb:
        for c in contigs.split(","):
            nm, ln = c.split(":")[0], int(c.split(":")[1])
            if ln < 200:
                fb.write(f"{nm}\t1\t{ln}\n")
                continue
            step = max(ln // 40, 150)
            for s in range(1, ln - 100, step):
                fb.write(f"{nm}\t{s}\t{min(ln, s + step // 3)}\n")
    for k, opts in enumerate([[], ["-w", "100"], ["-w", "1000"], ["-s"], ["-q", "20", "-d", "2"], ["-b", bed], ["-t", "1"], ["-t", "9", "-x", "0"]]):
        extra = [] if "-t" in opts else ["-t", "4"]
        ref, _ = _run(REF, cram, opts + extra, workdir, f"refcram_{k}")
        our, _ = _run(CLI, cram, opts + extra, workdir, f"ourcram_{k}")
        assert sorted(ref) == sorted(our), (opts, sorted(ref), sorted(our))
        for name in ref:
            assert ref[name] == our[name], f"cram opts {opts}: {name} differs from the reference binary"


def test_cli_cfg1_one_contig_one_million_reads(workdir):
    """BASELINE config 1: one 10 Mb contig, 1 M sorted 150 bp reads, per-chr mode, reference `-t 1` -- byte-identical output,
    with the device decoder (default) and with the host decoder."""
    bam = os.path.join(workdir, "cfg1.bam")
    subprocess.run([BAMGEN, "-o", bam, "-c", "chr1:10000000", "-n", "1000000", "-s", "12345", "-l", "150"], check=True)
    ref, _ = _run(REF, bam, ["-t", "1"], workdir, "cfg1_ref")
    our, _ = _run(CLI, bam, ["-t", "1"], workdir, "cfg1_our")
    host, _ = _run(CLI, bam, ["-t", "4"], workdir, "cfg1_host", dict(os.environ, PANDEPTH_GPU_DECODE="0"))
    assert sorted(ref) == sorted(our) == sorted(host) == [".chr.stat.gz"]
    assert ref[".chr.stat.gz"] == our[".chr.stat.gz"] == host[".chr.stat.gz"]
    row = ref[".chr.stat.gz"].splitlines()[1].split("\t")
    assert row[0] == "chr1" and row[1] == "10000000" and row[3] == "150000000"      # every aligned base counted once


def test_cli_device_decode_is_the_default_path(workdir):
    """PANDEPTH_TIMING names the decode path: whole-genome BAM input goes through pd_push_bgzf unless it is switched off."""
    bam = os.path.join(workdir, "cfg1.bam")
    if not os.path.exists(bam):
        subprocess.run([BAMGEN, "-o", bam, "-c", "chr1:10000000", "-n", "1000000", "-s", "12345", "-l", "150"], check=True)
    r = subprocess.run([CLI, "-i", bam, "-o", os.path.join(workdir, "dd"), "-t", "4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, PANDEPTH_TIMING="1"))
    assert r.returncode == 0 and "device-side BGZF decode" in r.stderr, r.stderr
    r = subprocess.run([CLI, "-i", bam, "-o", os.path.join(workdir, "dh"), "-t", "4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, PANDEPTH_TIMING="1", PANDEPTH_GPU_DECODE="0"))
    assert r.returncode == 0 and "device-side BGZF decode" not in r.stderr, r.stderr


def test_cli_usage_errors_exit_zero_like_the_reference(workdir):
    r = subprocess.run([CLI], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "Usage: pandepth" in r.stdout
    r = subprocess.run([CLI, "stray"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "Command option error" in r.stderr
    r = subprocess.run([CLI, "-i", "x.bam"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "lack argument -i or -o" in r.stderr
