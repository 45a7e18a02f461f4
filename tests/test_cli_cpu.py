"""
CPU: the host program (pandepth_b200/bin/pandepth: arguments, region model, SAM / BAM / CRAM / PAF / list feeders, batching,
report writers) end to end against
  (1) the golden outputs the reference binary printed (tests/golden/expected.json), and
  (2) the reference binary itself on freshly generated mixed-CIGAR BAMs and CRAMs,
with the CUDA engine replaced, FOR THESE TESTS ONLY, by tests/stand_in/pd_standin.cpp -- an adapter that implements the C ABI
over the oracle.  It is built into a temporary directory and injected with LD_LIBRARY_PATH; nothing in the product can load
it.  The same cases run against the real engine in tests/test_cli_gpu.py (-m gpu): this file is about the host logic, which
does not depend on where depth is computed, so that it can be checked on every CPU run.
"""
import gzip
import os
import subprocess

import pytest

from tests.helpers import GOLD, ROOT, assert_matches
from tests.test_cli_gpu import _CASES, OPTION_SETS

CLI = os.path.join(ROOT, "pandepth_b200", "bin", "pandepth")
REF = os.path.join(ROOT, "oracle", "_ref", "pandepth_ref")
BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")
ORACLE_DIR = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.skipif(not all(os.path.exists(p) for p in (CLI, REF, BAMGEN, os.path.join(ORACLE_DIR, "libpd_oracle.so"))),
                                reason="host binary / oracle checkers not built")


@pytest.fixture(scope="module")
def standin_env(tmp_path_factory):
    d = tmp_path_factory.mktemp("standin")
    so = str(d / "libpandepth_b200.so")
    subprocess.run(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", os.path.join(ROOT, "tests", "stand_in", "pd_standin.cpp"),
                    "-L" + ORACLE_DIR, "-lpd_oracle", "-Wl,-rpath," + ORACLE_DIR, "-Wl,-soname,libpandepth_b200.so", "-o", so], check=True)
    return dict(os.environ, LD_LIBRARY_PATH=str(d) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("clicpu"))


@pytest.fixture(scope="module")
def inputs(workdir):
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


@pytest.mark.parametrize("decode", ["device", "host"])
@pytest.mark.parametrize("case", sorted(_CASES))
def test_host_program_matches_golden(case, decode, inputs, workdir, standin_env):
    """decode = device: whole-genome BAM inputs go through pd_push_bgzf (answered here by the stand-in's block-parallel CPU
    decoder); decode = host: PANDEPTH_GPU_DECODE=0, the host feeders decode (the path -g / -b / SAM / CRAM always take)."""
    c = _CASES[case]
    if decode == "host":
        if c["kind"] != "bam":
            pytest.skip("only BAM input has two decode paths")
        standin_env = dict(standin_env, PANDEPTH_GPU_DECODE="0")
    args = [os.path.join(GOLD, a[1:]) if a.startswith("@") else a for a in c["args"]]

    def member(m):
        if m.endswith(".paf"):
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
    outs, _ = _run(CLI, inp, args + ["-t", "4"], workdir, "our_" + decode + case, standin_env)
    assert sorted(outs) == sorted(c["outputs"]), (sorted(outs), sorted(c["outputs"]))
    for k, want in c["outputs"].items():
        assert_matches(want, outs[k], f"{case}{k}")


def _region_files(workdir, tag, contigs):
    names = [c.split(":")[0] for c in contigs.split(",")]
    lens = [int(c.split(":")[1]) for c in contigs.split(",")]
    bed = os.path.join(workdir, tag + ".bed")
    gff = os.path.join(workdir, tag + ".gff")
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
    return bed, gff, lens


@pytest.mark.parametrize("seed,contigs,reads,ext", [
    (31, "a:300001,b:1,c:5000,d:70000,e:2,f:1001", 30000, "bam"),
    (32, "chrA:1000001,chrB:151,chrC:299999", 40000, "bam"),
    (33, "a:300001,b:1,c:5000,d:70000,e:2,f:1001", 30000, "cram"),
    (34, "big:20000123,s:149,t:10000001", 30000, "bam"),      # contigs longer than the reference's 10 Mb windows (:695-784)
    (35, "a:300001,b:1,c:5000,d:70000,e:2,f:1001", 30000, "csi"),  # BAM with a .csi index only: htslib iterators
])
def test_host_program_matches_reference_binary_on_random_inputs(seed, contigs, reads, ext, workdir, standin_env):
    inp = os.path.join(workdir, f"rand{seed}.{'bam' if ext == 'csi' else ext}")
    subprocess.run([BAMGEN, "-o", inp, "-c", contigs, "-n", str(reads), "-m", "mixed", "-s", str(seed)] + (["-X"] if ext == "csi" else []), check=True)
    assert os.path.exists(inp + {"bam": ".bai", "cram": ".crai", "csi": ".csi"}[ext])
    bed, gff, lens = _region_files(workdir, f"rand{seed}", contigs)
    sets = OPTION_SETS + [["-w", "150"], ["-w", "151", "-d", "2"], ["-b", bed], ["-b", bed, "-d", "3"], ["-g", gff], ["-g", gff, "-s"], ["-b", bed, "-s"],
                          ["-a", "-b", bed], ["-a", "-g", gff, "-s"], ["-a", "-b", bed, "-s", "-x", "0"]]
    for k, opts in enumerate(sets):
        if "-a" in opts and sum(lens) > 500_000:
            continue                     # the per-site dump is one line per base
        extra = [] if "-t" in opts else ["-t", "4"]
        ref, _ = _run(REF, inp, opts + extra, workdir, f"ref{seed}_{k}")
        our, _ = _run(CLI, inp, opts + extra, workdir, f"our{seed}_{k}", standin_env)
        assert sorted(ref) == sorted(our), (opts, sorted(ref), sorted(our))
        for name in ref:
            assert ref[name] == our[name], f"seed {seed} {ext} opts {opts}: {name} differs from the reference binary"


def test_header_says_sorted_but_records_are_not(workdir, standin_env):
    """`@HD SO:coordinate` over shuffled records: the reference does not care (its no-index loop just walks them); the host
    must not claim PD_SORTED for batches that are not in order (the stand-in, like the engine, rejects such a claim)."""
    import random
    sam = os.path.join(workdir, "claims_sorted.sam")
    subprocess.run([BAMGEN, "-o", sam, "-c", "a:200000,b:3000", "-n", "20000", "-m", "mixed", "-s", "77", "-S"], check=True)
    lines = open(sam).read().splitlines(keepends=True)
    head = [x for x in lines if x.startswith("@")]
    body = [x for x in lines if not x.startswith("@")]
    random.Random(5).shuffle(body)
    body.sort(key=lambda x: x.split("\t")[2])            # keep contigs together, positions shuffled inside
    with open(sam, "w") as fh:
        fh.writelines(head + body)
    assert "SO:coordinate" in head[0]
    for k, opts in enumerate([[], ["-w", "100"], ["-q", "20"], ["-w", "1000"], ["-w", "151", "-d", "2"]]):
        ref, _ = _run(REF, sam, opts + ["-t", "2"], workdir, f"refclaim_{k}")
        our, _ = _run(CLI, sam, opts + ["-t", "2"], workdir, f"ourclaim_{k}", standin_env)
        assert ref == our, opts


def test_cigar_longer_than_65535_ops_in_the_cg_tag(workdir, standin_env):
    """BAM keeps CIGARs with more than 65535 operations in a CG:B,I tag behind a `<l_seq>S<ref_len>N` placeholder
    (SAMv1 4.2.2); htslib's reader swaps the real one in, and so do the raw scanners (fastbam.hpp)."""
    sam = os.path.join(workdir, "cg.sam")
    ops = "1M1D" * 35000 + "10M"
    with open(sam, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:a\tLN:200000\n")
        fh.write("r0\t0\ta\t100\t60\t50M\t*\t0\t0\t*\t*\n")
        fh.write(f"r1\t0\ta\t1000\t60\t{ops}\t*\t0\t0\t*\t*\n")
        fh.write("r2\t0\ta\t1500\t60\t5S20M\t*\t0\t0\t*\t*\n")
    bam = os.path.join(workdir, "cg.bam")
    subprocess.run([BAMGEN, "-I", sam, "-o", bam], check=True)
    for k, opts in enumerate([[], ["-s"], ["-w", "100"]]):
        ref, _ = _run(REF, bam, opts + ["-t", "2"], workdir, f"refcg_{k}")
        our, _ = _run(CLI, bam, opts + ["-t", "2"], workdir, f"ourcg_{k}", standin_env)
        assert ref == our, opts
        assert "35070" in ref[".chr.stat.gz" if "-w" not in opts else ".win.stat.gz"] or "-w" in opts


def test_degenerate_inputs_match_the_reference(workdir, standin_env):
    """A header without @SQ lines (no engine is created), a header-only file, a zero-length contig."""
    nosq = os.path.join(workdir, "nosq.sam")
    with open(nosq, "w") as fh:
        fh.write("@HD\tVN:1.6\n")
    empty = os.path.join(workdir, "empty.sam")
    with open(empty, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:a\tLN:5000\n@SQ\tSN:b\tLN:0\n")
    empty_bam = os.path.join(workdir, "empty.bam")
    subprocess.run([BAMGEN, "-I", empty, "-o", empty_bam], check=True)
    k = 0
    for inp in (nosq, empty, empty_bam):
        for opts in ([], ["-a"], ["-w", "100"], ["-w", "1000"], ["-s"]):
            ref, _ = _run(REF, inp, opts, workdir, f"refdeg_{k}")
            our, _ = _run(CLI, inp, opts, workdir, f"ourdeg_{k}", standin_env)
            assert ref == our, (inp, opts)
            k += 1


def test_device_decode_entry_point_through_the_host_program(workdir, standin_env):
    """PANDEPTH_GPU_DECODE=1 (experimental, off by default): the host hands BGZF chunks to pd_push_bgzf.  The stand-in answers
    with the block-parallel decoder of pandepth_b200/csrc/pd_decode.hpp on a host backend, so this runs the chunker, the header
    skip, the contig runs and the fallback (reads larger than a block -> PD_ERR_FALLBACK -> host feeders) end to end; the
    real engine of this build always answers PD_ERR_FALLBACK, which the same fallback handles."""
    trace = os.path.join(workdir, "standin_trace.txt")
    env = dict(standin_env, PANDEPTH_GPU_DECODE="1", PD_STANDIN_TRACE=trace)
    cases = [("gd_mixed.bam", ["-c", "a:300001,b:1,c:5000,d:70000,e:2,f:1001", "-n", "60000", "-m", "mixed", "-s", "41", "-q"]),
             ("gd_long.bam", ["-c", "a:3000000,b:300000", "-n", "120", "-l", "100000", "-m", "mixed", "-s", "42", "-q"]),
             ("gd_small.bam", ["-c", "a:2000", "-n", "3", "-s", "43"])]
    for name, gen in cases:
        bam = os.path.join(workdir, name)
        subprocess.run([BAMGEN, "-o", bam] + gen, check=True)
        for k, opts in enumerate([[], ["-w", "100"], ["-w", "1000"], ["-q", "20", "-x", "0"], ["-a"] if "small" in name or "mixed" in name else ["-d", "3"]]):
            ref, _ = _run(REF, bam, opts + ["-t", "3"], workdir, f"refgd_{name}_{k}")
            if os.path.exists(trace):
                os.unlink(trace)
            our, _ = _run(CLI, bam, opts + ["-t", "3"], workdir, f"ourgd_{name}_{k}", env)
            assert ref == our, (name, opts)
            calls = open(trace).read().split()[0::3] if os.path.exists(trace) else []
            assert calls == (["bgzf_fallback"] if "long" in name else ["bgzf_ok"]), (name, opts, calls)     # the decoder ran / the fallback ran


def test_failed_runs_leave_no_partial_outputs(workdir, standin_env, inputs):
    """A truncated BAM fails with exit status 1 and leaves no (truncated) .gz at the output path; so does a list with an
    unreadable member (partial totals are not a result); -c (GC content, not provided) exits 1 instead of a silent 0."""
    src = inputs["big.sam"]["bam"]
    data = open(src, "rb").read()
    cut = os.path.join(workdir, "cut.bam")
    with open(cut, "wb") as fh:
        fh.write(data[:len(data) * 2 // 3])
    subprocess.run([BAMGEN, "-I", inputs["big.sam"]["sam"], "-o", os.path.join(workdir, "cutidx.bam")], check=True)
    os.replace(os.path.join(workdir, "cutidx.bam.bai"), cut + ".bai")
    # with an index both decode paths report the truncation (the reference itself crashes on this file) ...
    for mode_env in ({}, {"PANDEPTH_GPU_DECODE": "0"}):
        pre = os.path.join(workdir, "cutout")
        r = subprocess.run([CLI, "-i", cut, "-o", pre, "-t", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(standin_env, **mode_env))
        assert r.returncode == 1, (r.stdout, r.stderr)
        assert not os.path.exists(pre + ".chr.stat.gz")
    lst = os.path.join(workdir, "broken.list")
    with open(lst, "w") as fh:
        fh.write(inputs["k.sam"]["bam"] + "\n" + os.path.join(workdir, "does_not_exist.bam") + "\n")
    pre = os.path.join(workdir, "listout")
    r = subprocess.run([CLI, "-i", lst, "-o", pre], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=standin_env)
    assert r.returncode == 1 and not os.path.exists(pre + ".chr.stat.gz"), (r.stdout, r.stderr)
    r = subprocess.run([CLI, "-i", inputs["k.sam"]["bam"], "-o", os.path.join(workdir, "gc"), "-c"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=standin_env)
    assert r.returncode == 1 and "GC content" in r.stderr


@pytest.mark.parametrize("case", ["k.default.bam", "k.w150.bam", "mix.default.bam", "mix.w1000.bam", "mix.w151.bam", "big.default.bam", "big.w1000.bam", "k.q10d2.bam", "mix.q20d3x3844.bam"])
@pytest.mark.parametrize("devices", ["0,1", "0,1,2,3,4"])
def test_host_program_multi_device_range_sharding(case, devices, inputs, workdir, standin_env):
    """PANDEPTH_DEVICES: one engine per device, each owning an equal range of the concatenated genome (pd_create_ranges), fed
    through the .bai with the part of the file that can touch its range, reduce results summed -- same bytes as one device.
    (The stand-in ignores the device numbers; the host-side routing, ranges and sums are what is under test here.)"""
    c = _CASES[case]
    args = [os.path.join(GOLD, a[1:]) if a.startswith("@") else a for a in c["args"]]
    env = dict(standin_env, PANDEPTH_DEVICES=devices, PANDEPTH_TIMING="1")
    pre = os.path.join(workdir, "md_" + devices.replace(",", "") + case)
    r = subprocess.run([CLI, "-i", inputs[c["fixture"]]["bam"], "-o", pre] + args + ["-t", "4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "decode+push (multi-GPU)" in r.stderr, r.stderr
    for k, want in c["outputs"].items():
        with gzip.open(pre + k, "rt") as fh:
            assert_matches(want, fh.read(), f"{case}{k} devices={devices}")
