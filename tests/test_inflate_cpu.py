"""
CPU: pandepth_b200/csrc/pd_inflate.cuh -- the DEFLATE decoder and BAM record walker written as host+device functions
(groundwork for GPU-side BGZF decode) -- compiled for the host and checked against zlib:
every BGZF block of generated BAMs (SEQ-less, SEQ+QUAL, mixed CIGARs, many contigs), zlib streams at every level and strategy
(stored / fixed / dynamic blocks, empty to 64 KB inputs), truncated and bit-flipped streams (no out-of-bounds writes), and
the record walker's count against the number of reads generated.
"""
import os
import re
import subprocess

import pytest

from tests.helpers import GOLD, ROOT

BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("inflate") / "inflate_check")
    subprocess.run(["g++", "-O2", "-std=c++14", "-Wall", "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "inflate_check.cpp"), "-lz", "-o", exe], check=True)
    return exe


def _run(exe, args):
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    m = re.search(r"blocks (\d+), records (\d+), blocks ending inside a record (\d+), mismatches (\d+)", r.stdout)
    assert r.returncode == 0 and m, (r.stdout, r.stderr)
    return tuple(int(x) for x in m.groups())


def test_decoder_matches_zlib_on_synthetic_streams(checker):
    blocks, _, _, bad = _run(checker, ["--synth"])
    assert blocks == 1750 and bad == 0


@pytest.mark.skipif(not os.path.exists(BAMGEN), reason="bamgen not built")
def test_decoder_and_record_walker_on_generated_bams(checker, tmp_path):
    cases = [("a", ["-c", "chr1:3000000,chr2:1000000", "-n", "300000", "-s", "3"], 300000),
             ("b", ["-c", "chr1:3000000,chr2:1000000", "-n", "60000", "-s", "4", "-q"], 60000),
             ("m", ["-c", "a:300001,b:1,c:5000,d:70000,e:2,f:1001", "-n", "60000", "-m", "mixed", "-s", "11", "-q"], 60000)]
    for tag, args, reads in cases:
        bam = str(tmp_path / f"{tag}.bam")
        subprocess.run([BAMGEN, "-o", bam] + args, check=True)
        blocks, records, inside, bad = _run(checker, [bam])
        assert bad == 0 and blocks > 0
        assert records == reads
        assert inside == 0          # htslib's writer starts every block on a record boundary: blocks parse independently
    for fx, reads in (("k.sam", None), ("mix.sam", None), ("big.sam", None)):
        bam = str(tmp_path / (fx + ".bam"))
        subprocess.run([BAMGEN, "-I", os.path.join(GOLD, fx), "-o", bam], check=True)
        n_sam = sum(1 for line in open(os.path.join(GOLD, fx)) if not line.startswith("@"))
        blocks, records, inside, bad = _run(checker, [bam])
        assert bad == 0 and records == n_sam and inside == 0


@pytest.fixture(scope="module")
def decode_checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("decode") / "decode_check")
    subprocess.run(["g++", "-O2", "-std=c++14", "-Wall", "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "decode_check.cpp"), "-lz", "-o", exe], check=True)
    return exe


@pytest.mark.skipif(not os.path.exists(BAMGEN), reason="bamgen not built")
def test_block_parallel_decode_equals_the_sequential_walk(decode_checker, tmp_path):
    """pd_decode.hpp (inflate -> count -> prefix sums -> emit, every step per BGZF block, here on a host backend) against one
    sequential walk of the whole stream: same records in the same order, contig runs that cover them, no false claim of
    order; reads larger than a block are reported as such (the caller then uses the host feeders)."""
    def run(bam, *args):
        r = subprocess.run([decode_checker, bam] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        return r.returncode, r.stdout
    mixed = str(tmp_path / "m.bam")
    subprocess.run([BAMGEN, "-o", mixed, "-c", "a:300001,b:1,c:5000,d:70000,e:2,f:1001", "-n", "60000", "-m", "mixed", "-s", "11", "-q"], check=True)
    for args in ((), (0, -1, 1), (3844, 3, 5), (1796, 20, 1000)):
        rc, out = run(mixed, *args)
        assert rc == 0 and "digest equal" in out, (args, out)
    many = str(tmp_path / "many.sam")
    with open(many, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:c{i}\tLN:3000\n" for i in range(3000)))
        for i in range(3000):
            for p in (5, 100, 100, 2000):
                fh.write(f"r\t0\tc{i}\t{p}\t60\t150M\t*\t0\t0\t*\t*\n")
    many_bam = str(tmp_path / "many.bam")
    subprocess.run([BAMGEN, "-I", many, "-o", many_bam], check=True)
    rc, out = run(many_bam)
    assert rc == 0 and "records 12000 (want 12000)" in out and "contig runs 3000" in out, out
    # unsorted records under a sorted header: decoded all the same, and no run may claim order it does not have
    sam = str(tmp_path / "u.sam")
    subprocess.run([BAMGEN, "-o", sam, "-c", "a:200000,b:3000", "-n", "20000", "-m", "mixed", "-s", "77", "-S"], check=True)
    lines = open(sam).read().splitlines(keepends=True)
    import random
    body = [x for x in lines if not x.startswith("@")]
    random.Random(5).shuffle(body)
    with open(sam, "w") as fh:
        fh.writelines([x for x in lines if x.startswith("@")] + body)
    ubam = str(tmp_path / "u.bam")
    subprocess.run([BAMGEN, "-I", sam, "-o", ubam], stderr=subprocess.DEVNULL)        # (the index build fails on purpose: not sorted)
    rc, out = run(ubam, 0, -1, 3)
    assert rc == 0 and "records 20000 (want 20000)" in out and "run errors 0" in out, out
    # reads larger than a BGZF block
    long_bam = str(tmp_path / "long.bam")
    subprocess.run([BAMGEN, "-o", long_bam, "-c", "a:5000000", "-n", "50", "-l", "100000", "-s", "9", "-q"], check=True)
    rc, out = run(long_bam)
    assert rc == 3 and "returned -3" in out, out
