"""
CPU: the host PAF feeder (pandepth_b200/host/main.cpp paf_stream, role of paf_main PanDepth.cpp:1536-1623) must hand the
engine exactly the (tid, pos, cigar) records the oracle's parser (oracle/model.py parse_paf, pinned to the reference binary
by tests/test_oracle_golden.py) derives from the same text.  `PANDEPTH_DRY=1` makes the CLI parse without touching a GPU
and print an order-independent digest of everything it would push.
"""
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import model
from tests.helpers import GOLD, ROOT

CLI = os.path.join(ROOT, "pandepth_b200", "bin", "pandepth")
pytestmark = pytest.mark.skipif(not os.path.exists(CLI), reason="host binary not built")
M64 = (1 << 64) - 1


def _mix(x):
    x ^= x >> 33; x = (x * 0xff51afd7ed558ccd) & M64
    x ^= x >> 33; x = (x * 0xc4ceb9fe1a85ec53) & M64
    return x ^ (x >> 33)


def _digest_of(rec):
    """Batch::digest of main.cpp, summed over all records."""
    total = 0
    for i in range(rec.n):
        h = _mix(((int(rec.tid[i]) & 0xffffffff) << 32) | (int(rec.pos[i]) & 0xffffffff))
        for k in range(int(rec.cigar_off[i]), int(rec.cigar_off[i + 1])):
            h = _mix(h ^ int(rec.cigar[k]))
        total = (total + h) & M64
    return total


def _dry(inp, opts, tmp):
    env = dict(os.environ, PANDEPTH_DRY="1")
    r = subprocess.run([CLI, "-i", inp, "-o", os.path.join(tmp, "o")] + opts, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    m = re.search(r"dry: reads=(\d+) digest=(\d+)", r.stdout)
    assert r.returncode == 0 and m, (r.stdout, r.stderr)
    return int(m.group(1)), int(m.group(2)), r.stdout


@pytest.mark.parametrize("opts,x,q", [([], 1796, -1), (["-x", "0"], 0, -1), (["-q", "10"], 1796, 10), (["-x", "4", "-q", "31"], 4, 31)])
def test_paf_feeder_matches_oracle_parser(opts, x, q, tmp_path):
    paf = os.path.join(GOLD, "k.paf")
    _, rec, _ = model.parse_paf([paf], x, q)
    n, dig, out = _dry(paf, opts, str(tmp_path))
    assert "INFO: Run paf Format data" in out
    assert (n, dig) == (rec.n, _digest_of(rec))


def test_paf_list_and_gz(tmp_path):
    a, b = os.path.join(GOLD, "k.paf"), os.path.join(GOLD, "k2.paf")
    z = str(tmp_path / "k2.paf.gz")
    with open(b, "rb") as fi, gzip.open(z, "wb") as fo:
        fo.write(fi.read())
    lst = str(tmp_path / "two.list")
    with open(lst, "w") as fh:
        fh.write(a + "\n" + z + "\n")
    _, rec, _ = model.parse_paf([a, b], 1796, -1)
    # the unknown target name `zz` of k2.paf lands on target 0 (Chr2IntMap default-insert, PanDepth.cpp:1562)
    assert int(np.sum(rec.tid == 0)) == 5
    n, dig, out = _dry(lst, [], str(tmp_path))
    assert "INFO: Run multi-file data" in out and "INFO: Run PAF format data" in out
    assert (n, dig) == (rec.n, _digest_of(rec))


def test_paf_random_lines_against_oracle_parser(tmp_path):
    rng = np.random.default_rng(5)
    names = [("t%d" % i, int(rng.integers(2, 50000))) for i in range(6)]
    lines = []
    for i in range(4000):
        nm, ln = names[int(rng.integers(0, len(names)))]
        s = int(rng.integers(1, ln)); e = int(rng.integers(s, ln + 1))
        tags = ["tp:A:" + "PSI"[int(rng.integers(0, 3))]]
        if rng.random() < 0.7:
            ops = "".join("%d%s" % (int(rng.integers(1, 400)), "M=XDNISH"[int(rng.integers(0, 8))]) for _ in range(int(rng.integers(1, 9))))
            tags.append("cg:Z:" + ops)
        if rng.random() < 0.3:
            tags.insert(0, "NM:i:3")
        if rng.random() < 0.1:
            s, e = e, s
        sep = " " if rng.random() < 0.1 else "\t"
        lines.append("\t".join(["q%d" % i, "1000", "0", "500", "+-"[i & 1], nm, str(ln), str(s), str(e), "10", "20", str(int(rng.integers(0, 61)))]) + sep + sep.join(tags))
    paf = str(tmp_path / "rand.paf")
    with open(paf, "w") as fh:
        fh.write("\n".join(lines) + "\n\n")
    for opts, x, q in (([], 1796, -1), (["-x", "0", "-q", "20"], 0, 20)):
        _, rec, _ = model.parse_paf([paf], x, q)
        n, dig, _ = _dry(paf, opts, str(tmp_path))
        assert (n, dig) == (rec.n, _digest_of(rec))


def test_paf_golden_random_fixture(tmp_path):
    """rand.paf is also a golden fixture (the reference binary's outputs for it are in expected.json)."""
    paf = os.path.join(GOLD, "rand.paf")
    for opts, x, q in (([], 1796, -1), (["-x", "0", "-q", "20"], 0, 20)):
        _, rec, _ = model.parse_paf([paf], x, q)
        n, dig, _ = _dry(paf, opts, str(tmp_path))
        assert (n, dig) == (rec.n, _digest_of(rec))
