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
