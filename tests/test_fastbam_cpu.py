"""
CPU: the raw-record BAM scanner (pandepth_b200/host/fastbam.*) must hand the engine exactly the reads the htslib
iterator path does.  `PANDEPTH_DRY=1` makes the CLI decode without touching a GPU and print an order-independent digest
of every (tid, pos, cigar) it would push; `PANDEPTH_HTS_ITER=1` forces the htslib path.
"""
import os
import re
import subprocess

import pytest

from tests.helpers import GOLD, ROOT

CLI = os.path.join(ROOT, "pandepth_b200", "bin", "pandepth")
BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")

pytestmark = pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(BAMGEN)), reason="host binary / bamgen not built")


def _digest(bam, opts, hts):
    env = dict(os.environ, PANDEPTH_DRY="1")
    if hts:
        env["PANDEPTH_HTS_ITER"] = "1"
    r = subprocess.run([CLI, "-i", bam, "-o", bam + ".out"] + opts, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    m = re.search(r"dry: reads=(\d+) digest=(\d+) path=(\w+)", r.stdout)
    assert r.returncode == 0 and m, (r.stdout, r.stderr)
    return int(m.group(1)), int(m.group(2)), m.group(3)


@pytest.fixture(scope="module")
def bams(tmp_path_factory):
    d = tmp_path_factory.mktemp("fastbam")
    out = {}
    for fx in ("k", "mix", "big"):
        out[fx] = str(d / f"{fx}.bam")
        subprocess.run([BAMGEN, "-I", os.path.join(GOLD, fx + ".sam"), "-o", out[fx]], check=True)
    out["r1"] = str(d / "r1.bam")
    subprocess.run([BAMGEN, "-o", out["r1"], "-c", "a:300001,b:1,c:5000,d:70000,e:2,f:1001", "-n", "60000", "-m", "mixed", "-s", "11"], check=True)
    out["r2"] = str(d / "r2.bam")     # with SEQ/QUAL: records span BGZF blocks
    subprocess.run([BAMGEN, "-o", out["r2"], "-c", "chrA:2000001,chrB:999999", "-n", "150000", "-m", "mixed", "-s", "12", "-q"], check=True)
    out["long"] = str(d / "long.bam")    # 100 kb reads with SEQ/QUAL: every record is larger than a BGZF block
    subprocess.run([BAMGEN, "-o", out["long"], "-c", "a:5000000,b:300000", "-n", "300", "-l", "100000", "-m", "mixed", "-s", "9", "-q"], check=True)
    bed = str(d / "r.bed")
    with open(bed, "w") as fh:
        for nm, ln in (("a", 300001), ("c", 5000), ("d", 70000), ("f", 1001), ("chrA", 2000001), ("chrB", 999999), ("cA", 100001), ("cB", 1001), ("k1", 301)):
            step = max(ln // 50, 150)
            for s in range(1, ln - 100, step):
                fh.write(f"{nm}\t{s}\t{min(ln, s + step // 3)}\n")
            fh.write(f"{nm}\t{ln}\t{ln}\n")
    out["bed"] = bed
    return out


@pytest.mark.parametrize("name", ["k", "mix", "big", "r1", "r2", "long"])
def test_scanner_feeds_the_same_reads_as_htslib(name, bams):
    for opts in ([], ["-w", "100"], ["-b", bams["bed"]], ["-x", "0"], ["-q", "20"], ["-x", "3844", "-q", "3", "-t", "7"], ["-t", "1"]):
        fast = _digest(bams[name], opts, hts=False)
        slow = _digest(bams[name], opts, hts=True)
        assert fast[2] == "fastbam" and slow[2] == "htslib"
        assert fast[:2] == slow[:2], (name, opts, fast, slow)


def test_cram_indexed_ranges_feed_the_same_reads_as_one_sequential_pass(tmp_path):
    """CRAM + .crai goes through htslib iterators, one cram_fd and one index per worker (a .crai lives inside the cram_fd
    it was loaded through).  The range-parallel feed must hand over exactly the reads of a single sequential pass (-s)."""
    cram = str(tmp_path / "r.cram")
    subprocess.run([BAMGEN, "-o", cram, "-c", "a:300001,b:1,c:5000,d:70000,e:2,f:1001", "-n", "60000", "-m", "mixed", "-s", "21"], check=True)
    assert os.path.exists(cram + ".crai")
    seq = _digest(cram, ["-s"], hts=False)
    assert seq[2] == "htslib" and seq[0] > 10000
    for opts in ([], ["-t", "1"], ["-t", "7", "-w", "100"]):
        assert _digest(cram, opts, hts=False)[:2] == seq[:2], opts
    for opts in (["-q", "20"], ["-x", "0"]):
        assert _digest(cram, opts, hts=False)[:2] == _digest(cram, opts + ["-s"], hts=False)[:2], opts


def test_feeders_hand_over_exactly_the_oracle_parsers_records(bams):
    """An independent pin of the batch layout (pos / cigar / cigar offsets, implicit offsets of single-op batches): the digest
    of everything the CLI would push equals the digest of the records the oracle's SAM parser (oracle/model.py parse_sam)
    keeps after the record filter (PanDepth.cpp:732-733), for the SAM stream, the raw BAM scanner and the htslib iterators."""
    from oracle import model
    from tests.test_paf_host_cpu import _digest_of
    import numpy as np
    for fx in ("k", "mix", "big"):
        sam = os.path.join(GOLD, fx + ".sam")
        _, rec, _ = model.parse_sam(sam)
        for opts, x, q in (([], 1796, -1), (["-x", "0"], 0, -1), (["-x", "3844", "-q", "3"], 3844, 3)):
            keep = np.flatnonzero(((rec.flag & x) == 0) & (rec.mapq.astype(np.int64) >= q) & (rec.tid >= 0))
            want = rec.select(keep)
            want = (want.n, _digest_of(want))
            assert _digest(sam, opts, hts=False)[:2] == want, (fx, opts, "sam stream")
            assert _digest(bams[fx], opts, hts=False)[:2] == want, (fx, opts, "fastbam")
            assert _digest(bams[fx], opts + ["-t", "5"], hts=True)[:2] == want, (fx, opts, "htslib iterators")


@pytest.mark.parametrize("name", ["k", "mix", "big", "r1", "r2", "long"])
def test_stream_scanner_feeds_the_same_reads_as_sam_read1(name, bams):
    """No-index mode (-s): the ordered parallel BGZF reader + single record walker against htslib's sam_read1 loop."""
    for opts in (["-s"], ["-s", "-t", "1"], ["-s", "-x", "0", "-t", "7"], ["-s", "-q", "20", "-w", "100"]):
        fast = _digest(bams[name], opts, hts=False)
        slow = _digest(bams[name], opts, hts=True)
        assert fast[2] == "fastbam" and slow[2] == "htslib"
        assert fast[:2] == slow[:2], (name, opts, fast, slow)


def test_stream_scanner_on_large_and_damaged_files(tmp_path):
    """Records and the header straddling many blocks and groups; a file cut in the middle of a block or of a record ends
    quietly with a prefix of the reads (the reference's `while (sam_read1(...) >= 0)` does the same); garbage is an error."""
    big = str(tmp_path / "s.bam")        # SEQ+QUAL: ~100 MB inflated, groups of 4 MB -> records straddle every group boundary
    subprocess.run([BAMGEN, "-o", big, "-c", "a:3000001,b:999,c:2000000", "-n", "400000", "-s", "5", "-q"], check=True)
    want = _digest(big, [], hts=False)
    assert _digest(big, ["-s"], hts=False)[:2] == want[:2] == _digest(big, ["-s"], hts=True)[:2]
    data = open(big, "rb").read()
    for cut in (len(data) // 2, len(data) - 40, 70000):
        t = str(tmp_path / f"cut{cut}.bam")
        with open(t, "wb") as fh:
            fh.write(data[:cut])
        n, _, path = _digest(t, ["-s"], hts=False)
        assert path == "fastbam" and 0 <= n < want[0]
    g = str(tmp_path / "garbage.bam")
    with open(g, "wb") as fh:
        fh.write(data[:300000] + bytes(range(256)) * 400 + data[300000:])
    env = dict(os.environ, PANDEPTH_DRY="1")
    r = subprocess.run([CLI, "-i", g, "-o", g + ".out", "-s"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 1 and "malformed BAM" in r.stderr
