"""
CPU: the device BGZF decoder's kernels (pandepth_b200/csrc/pd_bgzf.cuh: k_bgzf_decode, k_bgzf_offsets, k_bgzf_compact) run
through a warp emulator (tests/cpp/warp_emu.h: one host thread per lane, *_sync collectives as barriers) and are compared,
block by block, with zlib + a sequential record walk: inflated bytes, dense pos / tid / cigar_off / cigar arrays, per-block
summaries and flags.  This checks the kernels' index arithmetic and warp choreography without a GPU; the same kernels are
checked on the device against the oracle's depth in tests/test_bgzf_gpu.py (-m gpu).
"""
import os
import subprocess
import zlib

import numpy as np
import pytest

from tests import bgzf_util as bu
from tests.helpers import ROOT

BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")

pytestmark = pytest.mark.skipif(not os.path.exists(BAMGEN), reason="oracle/_ref/bamgen not built")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu")
    exe = str(d / "bgzf_emu_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"), "-I" + os.path.join(ROOT, "tests", "cpp"),
                    os.path.join(ROOT, "tests", "cpp", "bgzf_emu_check.cpp"), "-lz", "-o", exe], check=True)
    return exe


@pytest.fixture(scope="module")
def bams(tmp_path_factory):
    d = tmp_path_factory.mktemp("emubam")
    res = {}
    res["A"] = str(d / "A.bam")          # flavour A of SURVEY 8(d): SEQ-less 150M reads (match-dominated deflate streams)
    subprocess.run([BAMGEN, "-o", res["A"], "-c", "c1:300000,c2:200000", "-n", "30000", "-s", "7", "-l", "150"], check=True)
    res["B"] = str(d / "B.bam")          # flavour B: random SEQ + QUAL (literal-dominated)
    subprocess.run([BAMGEN, "-o", res["B"], "-c", "c1:300000", "-n", "1500", "-s", "7", "-l", "150", "-q"], check=True)
    res["M"] = str(d / "M.bam")          # mixed CIGARs / flags / mapq, five contigs (one of length 1)
    subprocess.run([BAMGEN, "-o", res["M"], "-c", "c1:200000,c2:1,c3:150000,c4:900,c5:30000", "-n", "12000", "-s", "3", "-l", "100", "-m", "mixed"], check=True)
    data = open(res["M"], "rb").read()
    raw, first, recs, _ = bu.bam_payload(data)
    variants = {"stored": dict(level=0), "fixed": dict(level=6, strategy=zlib.Z_FIXED), "huff": dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY),
                "rle": dict(level=6, strategy=zlib.Z_RLE), "l9": dict(level=9)}
    for name, kw in variants.items():
        cuts = bu.record_cuts(raw, first, recs, max_block=9000, header_own_block=name != "l9")     # l9: header shares its block with records
        res[name] = str(d / f"M_{name}.bam")
        open(res[name], "wb").write(bu.rewrite_bgzf(raw, cuts, **kw))
    n = 150
    small = raw[:recs[n]]
    res["tiny"] = str(d / "M_tiny.bam")                                                     # one record per block
    open(res["tiny"], "wb").write(bu.rewrite_bgzf(small, [0, first] + recs[1:n] + [len(small)], level=6))
    res["straddle"] = str(d / "M_straddle.bam")                                            # records cut by block boundaries
    open(res["straddle"], "wb").write(bu.rewrite_bgzf(raw, list(range(0, len(raw), 7000)) + [len(raw)], level=6))
    rng = np.random.default_rng(3)                                                          # records shuffled: unsorted blocks, mixed contigs
    perm = rng.permutation(len(recs))
    ends = recs[1:] + [len(raw)]
    shuf = raw[:first] + b"".join(raw[recs[i]:ends[i]] for i in perm)
    lens = np.array([ends[i] - recs[i] for i in perm], np.int64)
    starts = (first + np.concatenate([[0], np.cumsum(lens)[:-1]])).tolist()
    res["shuffled"] = str(d / "M_shuf.bam")
    open(res["shuffled"], "wb").write(bu.rewrite_bgzf(shuf, bu.record_cuts(shuf, first, starts, max_block=9000), level=4))
    return res


CASES = [
    # (file, max_blocks, xflags, min_mapq, warps)
    ("A", 3, 1796, -1, 2), ("B", 2, 1796, -1, 1), ("B", 7, 0, -1, 2), ("M", 4, 1796, -1, 2), ("M", 3, 0, 20, 1), ("stored", 6, 1796, -1, 2), ("fixed", 5, 3844, 5, 1),
    ("huff", 4, 1796, -1, 1), ("rle", 4, 0, 20, 2), ("l9", 5, 1796, -1, 1), ("tiny", 150, 1796, -1, 2), ("straddle", 6, 1796, -1, 1), ("shuffled", 5, 0, 20, 2),
]


@pytest.mark.parametrize("K", [1, 2, 4, 8])
@pytest.mark.parametrize("name,blocks,xflags,q,warps", CASES)
def test_emulated_decoder_matches_zlib_and_sequential_walk(name, blocks, xflags, q, warps, K, checker, bams):
    """K = 1: k_bgzf_decode; K = 2, 4, 8: k_bgzf_decode_multi<K> (K decoders per warp, warp-uniform symbol loop)."""
    r = subprocess.run([checker, bams[name], str(blocks), str(xflags), str(q), str(warps), "0", str(K)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith(": ok"), (r.stdout[-2000:], r.stderr[-2000:])
