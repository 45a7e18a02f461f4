"""
Pins the oracle (oracle/pd_oracle.c + oracle/model.py) to the reference: every golden case stored in
tests/golden/expected.json was printed by the reference binary itself (tests/golden/make_golden.py),
and includes the hand-verified known-answer vectors of SURVEY.md App. C.  CPU only.
"""
import json
import os

import numpy as np
import pytest

from oracle import model
from tests.helpers import GOLD, assert_matches, load_fixture, options_from_args

with open(os.path.join(GOLD, "expected.json")) as _fh:
    _CASES = json.load(_fh)


@pytest.mark.parametrize("case", sorted(_CASES))
def test_model_matches_reference(case):
    c = _CASES[case]
    targets, rec, sorted_hdr = load_fixture(c["fixture"], c["args"])
    opts = options_from_args(c["args"], c["kind"], c["fixture"])
    res = model.run_model(targets, rec, opts, sorted_hdr=sorted_hdr)
    outs = c["outputs"]
    assert res["suffix"] in outs, (res["suffix"], list(outs))
    assert_matches(outs[res["suffix"]], res["stat"], case)
    if ".SiteDepth.gz" in outs:
        assert_matches(outs[".SiteDepth.gz"], res["site"], case + " site")


def test_known_answer_vectors_appendix_c():
    """SURVEY.md App. C, hand-verified base by base: k.sam default / -q 10 -d 2."""
    targets, rec, _ = load_fixture("k.sam")
    r = model.run_model(targets, rec, model.Options(indexed=False))
    assert "k1\t301\t120\t175\t39.87\t0.58\n" in r["stat"]
    assert "k3\t250\t60\t60\t24.00\t0.24\n" in r["stat"]
    assert r["stat"].endswith("##RegionLength: 551\tCoveredSite: 180\tCoverage(%): 32.67\tMeanDepth: 0.43\n")
    r = model.run_model(targets, rec, model.Options(indexed=False, min_mapq=10, min_dep=2))
    assert "k1\t301\t25\t50\t8.31\t0.17\n" in r["stat"]


def test_tiling_quirks():
    """PanDepth.cpp:3997-4046: len<2 vanishes; len % W == 1 drops the last 1-bp tile."""
    assert model.tile_contig(1, 10_000_000)[0].size == 0
    a, b = model.tile_contig(2, 10_000_000)
    assert list(zip(a, b)) == [(1, 2)]
    a, b = model.tile_contig(20_000_001, 10_000_000)
    assert list(zip(a, b)) == [(1, 10_000_000), (10_000_001, 20_000_000)]
    a, b = model.tile_contig(20_000_002, 10_000_000)
    assert list(zip(a, b))[-1] == (20_000_001, 20_000_002)
    a, b = model.tile_contig(301, 150)
    assert list(zip(a, b)) == [(1, 150), (151, 300)]


def test_mode6_window_quirk():
    """PanDepth.cpp:4366: `for (j=1; j<CC; j+=W)` drops a window that would start at j == len."""
    a, b = model.mode6_windows(301, 100)
    assert list(zip(a, b)) == [(0, 100), (100, 200), (200, 300)]
    a, b = model.mode6_windows(250, 100)
    assert list(zip(a, b)) == [(0, 100), (100, 200), (200, 250)]
    assert model.mode6_windows(1, 100)[0].size == 0


def test_accumulate_matches_python_loop():
    """The C walk against a literal pure-Python restatement of PanDepth.cpp:735-760 on the mixed fixture."""
    targets, rec, _ = load_fixture("mix.sam")
    tlen = np.array([t[1] for t in targets], np.uint32)
    got = model.accumulate(tlen, rec, 1796, 5)
    want = [np.zeros(int(n), np.int64) for n in tlen]
    for i in range(rec.n):
        if rec.flag[i] & 1796 or rec.mapq[i] < 5 or rec.tid[i] < 0:
            continue
        p = int(rec.pos[i]); d = want[rec.tid[i]]
        for c in rec.cigar[rec.cigar_off[i]:rec.cigar_off[i + 1]]:
            op, n = int(c) & 15, int(c) >> 4
            if op in (0, 7, 8):
                d[max(p, 0):min(p + n, d.size)] += 1
                p += n
            elif op in (2, 3):
                p += n
    for g, w in zip(got, want):
        assert np.array_equal(g.astype(np.int64), w)


def test_18bit_wrap_vector():
    """SURVEY App. C `w.sam`: 262149 stacked reads -> 2621490 via uint32, 50 via the 18-bit SiteInfo field."""
    targets, rec, _ = load_fixture("@stack")
    d = model.accumulate(np.array([1000], np.uint32), rec)[0]
    cov, tot = model.reduce_intervals(d, [0], [1000], 1, 0xFFFFFFFF)
    assert (int(cov[0]), int(tot[0])) == (10, 2621490)
    cov, tot = model.reduce_intervals(d, [0], [1000], 1, 0x3FFFF)
    assert (int(cov[0]), int(tot[0])) == (10, 50)
