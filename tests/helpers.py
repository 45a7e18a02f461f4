"""Shared helpers for the parity tests (golden-case plumbing)."""
import hashlib
import os

import numpy as np

from oracle import model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

STACK_TARGETS = [("c", 1000)]


def _paf_filters(args):
    """(-x, -q) of a golden case: paf_main applies both while parsing the text (PanDepth.cpp:1552-1568)."""
    x, q = 1796, -1
    for i, a in enumerate(args or []):
        if a == "-x":
            x = int(args[i + 1])
        elif a == "-q":
            q = int(args[i + 1])
    return x, q


def load_fixture(name, args=None):
    """-> (targets, Records, sorted) for a golden fixture name ('k.sam', '@stack', 'k.paf', 'list:...')."""
    members = name[5:].split(",") if name.startswith("list:") else [name]
    if members[0].endswith(".paf"):
        return model.parse_paf([os.path.join(GOLD, m) for m in members], *_paf_filters(args))
    if name.startswith("list:"):
        parts = [load_fixture(x) for x in name[5:].split(",")]
        targets = parts[0][0]
        recs = [p[1] for p in parts]
        off = [0]
        for r in recs:
            off.append(off[-1] + int(r.cigar.size))
        cat = model.Records(np.concatenate([r.tid for r in recs]), np.concatenate([r.pos for r in recs]),
                            np.concatenate([r.flag for r in recs]), np.concatenate([r.mapq for r in recs]),
                            np.concatenate([[0]] + [r.cigar_off[1:].astype(np.int64) + o for r, o in zip(recs, off)]).astype(np.uint32),
                            np.concatenate([r.cigar for r in recs]))
        return targets, cat, False
    if name == "@stack":
        n = 262149
        rec = model.Records(np.zeros(n, np.int32), np.full(n, 99, np.int32), np.zeros(n, np.uint16),
                            np.full(n, 60, np.uint8), np.arange(n + 1, dtype=np.uint32),
                            np.full(n, (10 << 4) | 0, np.uint32))
        return STACK_TARGETS, rec, True
    return model.parse_sam(os.path.join(GOLD, name))


def options_from_args(args, kind, fixture=""):
    """Translate a golden case's CLI args into model.Options (what bamCov_help01 would set, PanDepth.cpp:84-293)."""
    o = model.Options()
    o.indexed = (kind == "bam") and not fixture.startswith("list:")     # multi-file mode (:2687) and PAF (:1520) are 18-bit paths
    it = iter(args)
    for a in it:
        if a == "-w":
            o.win = max(1, int(next(it)))
        elif a == "-q":
            o.min_mapq = int(next(it))
        elif a == "-d":
            o.min_dep = max(1, int(next(it)))
        elif a == "-x":
            o.xflags = int(next(it))
        elif a == "-a":
            o.site_output = True
        elif a == "-s":
            o.indexed = False
        elif a == "-f":
            o.feature = next(it)
        elif a == "-g":
            o.gff = os.path.join(GOLD, next(it)[1:])
        elif a == "-b":
            o.bed = os.path.join(GOLD, next(it)[1:])
        else:
            raise ValueError(a)
    return o


def assert_matches(expected_entry, text, what=""):
    """Compare text with a golden output (verbatim string or compacted {sha256,...})."""
    if isinstance(expected_entry, str):
        assert text == expected_entry, f"{what}: output differs from the reference"
    else:
        assert len(text) == expected_entry["n_bytes"], f"{what}: size differs"
        assert text[:2048] == expected_entry["head"], f"{what}: head differs"
        assert text[-512:] == expected_entry["tail"], f"{what}: tail differs"
        assert hashlib.sha256(text.encode()).hexdigest() == expected_entry["sha256"], f"{what}: sha256 differs"
