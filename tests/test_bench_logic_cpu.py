"""CPU checks of bench.py's workload logic (no GPU): the tile-defined read set is the same whichever way it is cut into
ranks, range-sharded ranks add up to the single-rank depth, and the BAM the two arms share is read back correctly."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import model
from tests.helpers import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _depth(tlen, t, pos):
    n = pos.size
    rec = model.Records(np.full(n, t, np.int32), pos.astype(np.int32), np.zeros(n, np.uint16), np.zeros(n, np.uint8),
                        np.arange(n + 1, dtype=np.uint32), np.full(n, (bench.READLEN << 4), np.uint32))
    return model.accumulate(tlen, rec)[t]


def test_tile_plan_reads_are_rank_independent_and_ranges_add_up():
    from pandepth_b200.shard import assign_ranges
    lens = [230_017, 90_000, 151]
    tlen = np.array(lens, np.uint32)
    plan = bench.tile_plan(lens, 40.0, tile=50_000)
    R_total = sum(p[3] for p in plan)
    whole = {t: bench.rank_reads(plan, t, 0, lens[t], "cpu", torch).numpy() for t in range(3)}
    assert sum(v.size for v in whole.values()) == R_total
    for t in range(3):
        assert (np.diff(whole[t]) >= 0).all() and whole[t].min() >= 0 and whole[t].max() + bench.READLEN <= lens[t]
    want = [_depth(tlen, t, whole[t]) for t in range(3)]
    assert sum(int(w.sum()) for w in want) == bench.READLEN * R_total
    for world in (2, 3, 5):
        beg, end = assign_ranges(lens, world)
        got = [np.zeros(n, np.int64) for n in lens]
        for r in range(world):
            for t in range(3):
                b, z = int(beg[r, t]), int(end[r, t])
                if z <= b:
                    continue
                pos = bench.rank_reads(plan, t, b, z, "cpu", torch).numpy()
                # every read of the global set that overlaps [b, z) is there (boundary reads included) ...
                need = whole[t][(whole[t] + bench.READLEN > b) & (whole[t] < z)]
                assert np.isin(need, pos).all()
                d = _depth(tlen, t, pos)
                got[t][b:z] += d[b:z]                     # ... and the engine keeps only its range
        for t in range(3):
            assert np.array_equal(got[t], want[t].astype(np.int64)), (world, t)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bamgen")), reason="oracle/_ref not built")
def test_shared_bam_and_reference_rows(tmp_path, monkeypatch):
    monkeypatch.setattr(bench.tempfile, "gettempdir", lambda: str(tmp_path))
    bam, reads, why = bench.make_sample_bam(3.0, 50.0, False, threads=2)
    assert bam and os.path.exists(bam + ".bai") and reads == 1_000_000
    times, rows = bench.time_reference(bam, reads, 2, 1, 0, keep_rows=True)
    assert len(times) == 1 and len(rows) == 9
    assert sum(r[2] for r in rows.values()) == bench.READLEN * reads          # the -P generator keeps every read inside its contig
    lens = bench.contig_lengths(3.0e6)
    assert [rows[f"chr{i + 1}"][0] for i in range(9)] == lens
    # the host ABI finds block offsets through the .bai written by bamgen -P
    import ctypes
    H = ctypes.CDLL(os.path.join(ROOT, "pandepth_b200", "libpandepth_host.so"))
    H.pdh_bai_block_offset.restype = ctypes.c_uint64
    H.pdh_bai_block_offset.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int64]
    offs = [H.pdh_bai_block_offset(bam.encode(), t, 0) for t in range(9)]
    assert offs == sorted(offs) and offs[0] > 0 and offs[-1] < os.path.getsize(bam)
    assert H.pdh_bai_block_offset(bam.encode(), 8, lens[8] + 20000) == 2 ** 64 - 1
