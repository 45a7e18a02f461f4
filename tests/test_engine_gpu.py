"""
GPU parity tests: the CUDA engine, called through its C ABI (include/pandepth_b200.h), against the oracle
(oracle/pd_oracle.c restatement, itself pinned to the reference binary) and against the reference's own printed
outputs (tests/golden/expected.json).  Integer work: every comparison is bit-exact.
"""
import json
import os

import numpy as np
import pytest

from oracle import model
from tests.helpers import GOLD, ROOT, assert_matches, load_fixture, options_from_args

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLD, "expected.json")) as _fh:
    _CASES = json.load(_fh)


def _engine(tlen, **kw):
    import pandepth_b200 as pd
    return pd.Engine(tlen, **kw)


def _sorted_order(rec):
    return np.lexsort((rec.pos, rec.tid))


def _push_all(eng, rec, sorted_flag=True, chunks=1):
    """Push a Records batch through pd_push in `chunks` consecutive pieces."""
    import pandepth_b200 as pd
    n = rec.n
    edges = np.linspace(0, n, chunks + 1).astype(np.int64)
    for a, b in zip(edges[:-1], edges[1:]):
        if b <= a:
            continue
        sub = rec.select(np.arange(a, b))
        eng.push(sub.tid, sub.pos, sub.cigar, flag=sub.flag, mapq=sub.mapq, cigar_off=sub.cigar_off,
                 flags=pd.PD_SORTED if sorted_flag else 0)


def _check_depth(eng, tlen, want):
    for t, w in enumerate(want):
        if tlen[t] == 0:
            continue
        got = eng.depth(t)
        assert got.shape == w.shape
        if not np.array_equal(got, w):
            bad = np.flatnonzero(got != w)
            raise AssertionError(f"contig {t}: {bad.size} bases differ, first at {bad[0]}: got {got[bad[0]]} want {w[bad[0]]}")


# ------------------------------------------------------------------------------------------------------
# golden fixtures: per-base depth, interval reduce and the final printed text
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fixture", ["k.sam", "mix.sam", "big.sam"])
@pytest.mark.parametrize("path", ["sorted", "sorted_chunked", "unsorted"])
def test_depth_matches_oracle_on_fixtures(fixture, path):
    targets, rec, _ = load_fixture(fixture)
    tlen = np.array([t[1] for t in targets], np.uint32)
    for xflags, q in ((1796, -1), (0, 20), (3844, 5)):
        want = model.accumulate(tlen, rec, xflags, q)
        with _engine(tlen, xflags=xflags, min_mapq=q) as eng:
            if path == "unsorted":
                perm = np.random.default_rng(1).permutation(rec.n)
                _push_all(eng, rec.select(perm), sorted_flag=False, chunks=3)
            else:
                r = rec.select(_sorted_order(rec))
                _push_all(eng, r, sorted_flag=True, chunks=1 if path == "sorted" else 7)
            eng.finalize()
            _check_depth(eng, tlen, want)
            st = eng.stats()
            assert st["kernel_launches"] > 0


@pytest.mark.parametrize("case", sorted(_CASES))
def test_printed_output_matches_reference(case):
    """Engine depth + engine reduce -> the exact text the reference binary printed for this case."""
    c = _CASES[case]
    targets, rec, sorted_hdr = load_fixture(c["fixture"], c["args"])
    opts = options_from_args(c["args"], c["kind"], c["fixture"])
    tlen = np.array([t[1] for t in targets], np.uint32)
    if opts.site_output:
        rec = rec.select(model.reads_reaching_depth(targets, rec, opts, sorted_hdr))    # -a with -g / -b: only the reads the region path keeps
    r = rec.select(_sorted_order(rec))
    with _engine(tlen, xflags=opts.xflags, min_mapq=opts.min_mapq) as eng:
        _push_all(eng, r, sorted_flag=True, chunks=2)
        eng.finalize()
        depth = [eng.depth(t) if tlen[t] else np.zeros(0, np.uint32) for t in range(len(targets))]
        res = model.run_model(targets, rec, opts, depth_override=depth)
        assert_matches(c["outputs"][res["suffix"]], res["stat"], case)
        if ".SiteDepth.gz" in c["outputs"]:
            assert_matches(c["outputs"][".SiteDepth.gz"], res["site"], case + " site")
        # and the device-side reducers against the oracle's on the same depth
        whole = (not opts.indexed) or opts.site_output or res["mode"] == 6
        mask = 0x3FFFF if whole else 0xFFFFFFFF
        for t, (nm, ln) in enumerate(targets):
            if ln < 2:
                continue
            a, b = model.tile_contig(ln, 97)
            cov_o, sum_o = model.reduce_intervals(depth[t], a - 1, b, opts.min_dep, mask)
            cov_g, sum_g = eng.reduce(np.full(a.size, t, np.int32), a - 1, b, opts.min_dep, mask)
            assert np.array_equal(cov_o, cov_g) and np.array_equal(sum_o, sum_g)
            for win in (7, 100, 1000):
                s0, e0 = model.mode6_windows(ln, win)
                cov_o, sum_o = model.reduce_intervals(depth[t], s0, e0, opts.min_dep, mask)
                cov_w, sum_w = eng.reduce_windows(t, 0, win, s0.size, opts.min_dep, mask)
                assert np.array_equal(cov_o, cov_w.astype(np.uint64)) and np.array_equal(sum_o, sum_w)


# ------------------------------------------------------------------------------------------------------
# seeded synthetic batches (SURVEY 8d shapes, scaled so the oracle finishes in seconds)
# ------------------------------------------------------------------------------------------------------
def _uniform_reads(rng, length, n, readlen=150):
    pos = np.sort(rng.integers(0, max(1, length - readlen + 1), n)).astype(np.int32)
    cig = np.full(n, (readlen << 4) | 0, np.uint32)
    return pos, cig


def test_cfg1_shape_sorted_150M():
    """BASELINE config 1 shape: one 10 Mb contig, 1 M sorted 150M reads (seed 12345)."""
    rng = np.random.default_rng(12345)
    L = 10_000_000
    pos, cig = _uniform_reads(rng, L, 1_000_000)
    tlen = np.array([L], np.uint32)
    rec = model.Records(np.zeros(pos.size, np.int32), pos, np.zeros(pos.size, np.uint16), np.full(pos.size, 60, np.uint8),
                        np.arange(pos.size + 1, dtype=np.uint32), cig)
    want = model.accumulate(tlen, rec)
    import pandepth_b200 as pd
    with _engine(tlen) as eng:
        eng.push_contig(0, pos, cig, flags=pd.PD_SORTED)
        eng.finalize()
        _check_depth(eng, tlen, want)
        a, b = model.tile_contig(L, 10_000_000)
        cov, tot = eng.reduce(np.zeros(a.size, np.int32), a - 1, b)
        assert int(tot.sum()) == 150 * pos.size            # every aligned base counted once
        co, so = model.reduce_intervals(want[0], a - 1, b)
        assert np.array_equal(cov, co) and np.array_equal(tot, so)
        st = eng.stats()
        assert st["tile_path_batches"] == 1 and st["delta_path_batches"] == 0 and st["scan_passes"] == 0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_mixed_cigars_all_paths(seed):
    """Random CIGARs over MIDNSHP=X with long N skips (far runs), random flags/mapq, duplicates, three push styles."""
    rng = np.random.default_rng(seed)
    lens = [300_001, 1, 5000, 70_000, 2]
    names = [f"c{i}" for i in range(len(lens))]
    rows = []
    for t, ln in enumerate(lens):
        n = 0 if ln < 10 else int(ln * (0.5 if t else 0.2))
        starts = np.sort(rng.integers(0, ln, n))
        if n:
            starts[rng.integers(0, n, n // 10)] = starts[n // 2]          # a pile-up at one position
            starts = np.sort(starts)
        for p in starts.tolist():
            k = rng.integers(0, 10)
            if k < 5:
                cs = "150M"
            elif k == 5:
                cs = f"{rng.integers(1, 40)}S{rng.integers(1, 120)}M"
            elif k == 6:
                cs = f"{rng.integers(1, 80)}M{rng.integers(1, 5000)}N{rng.integers(1, 80)}M"
            elif k == 7:
                cs = f"{rng.integers(1, 60)}={rng.integers(1, 9)}I{rng.integers(1, 60)}X{rng.integers(1, 30)}D{rng.integers(1, 60)}M2H"
            elif k == 8:
                cs = f"3D{rng.integers(1, 100)}M1P{rng.integers(1, 100)}M"
            else:
                cs = f"{rng.integers(1, 300)}M"
            fl = int(rng.choice([0, 0, 0, 16, 1024, 256, 4, 2048, 512]))
            rows.append((fl, t, p + 1, int(rng.integers(0, 61)), cs))
    rec = model.records_from_tuples(rows, names)
    tlen = np.array(lens, np.uint32)
    for xflags, q in ((1796, -1), (0, 30)):
        want = model.accumulate(tlen, rec, xflags, q)
        for style in ("sorted", "chunked", "unsorted", "half_half"):
            with _engine(tlen, xflags=xflags, min_mapq=q) as eng:
                if style == "sorted":
                    _push_all(eng, rec, True, 1)
                elif style == "chunked":
                    _push_all(eng, rec, True, 13)
                elif style == "unsorted":
                    _push_all(eng, rec.select(rng.permutation(rec.n)), False, 2)
                else:                                   # depth is additive across batches and paths
                    half = rec.n // 2
                    _push_all(eng, rec.select(np.arange(0, half)), True, 2)
                    _push_all(eng, rec.select(rng.permutation(np.arange(half, rec.n))), False, 1)
                eng.finalize()
                _check_depth(eng, tlen, want)


def test_striped_scatter_applies_every_run_end_once(monkeypatch):
    """Big unsorted batches scatter stripe by stripe (one launch per L2-sized stripe of the delta array, run ends outside the
    stripe skipped): forced here with a 4 KB stripe, so runs straddle stripes, end on stripe boundaries and on the contig's
    padding element.  Same depth as the oracle, multi-op CIGARs and the device-side filter included."""
    monkeypatch.setenv("PANDEPTH_SCATTER_STRIPE_MB", str(1024 * 4 / (1 << 20)))     # 1024 elements per stripe
    monkeypatch.setenv("PANDEPTH_SCATTER_STRIPE_MIN_READS", "64")
    rng = np.random.default_rng(21)
    L = 15_000                                                                       # 15 stripes (the cap is 16)
    rows = []
    for p in rng.integers(0, L, 20_000).tolist():
        k = rng.integers(0, 6)
        cs = ("150M", f"{rng.integers(1, 1100)}M", f"{rng.integers(1, 80)}M{rng.integers(1, 3000)}N{rng.integers(1, 80)}M",
              f"{rng.integers(1, 40)}S{rng.integers(1, 120)}M", "1024M", f"{rng.integers(1, 60)}={rng.integers(1, 9)}I{rng.integers(1, 60)}X{rng.integers(1, 30)}D{rng.integers(1, 60)}M")[k]
        rows.append((int(rng.choice([0, 16, 1024, 256])), 0, p + 1, int(rng.integers(0, 61)), cs))
    rows += [(0, 0, 1, 60, "1024M"), (0, 0, L - 1023, 60, "1024M"), (0, 0, L, 60, "5M"), (0, 0, 1025, 60, "1024M")]   # stripe-aligned ends, contig end
    rec = model.records_from_tuples(rows, ["c"])
    tlen = np.array([L], np.uint32)
    for xflags, q in ((1796, -1), (0, 30)):
        want = model.accumulate(tlen, rec, xflags, q)
        with _engine(tlen, xflags=xflags, min_mapq=q) as eng:
            _push_all(eng, rec, False, 1)
            eng.finalize()
            _check_depth(eng, tlen, want)
            assert eng.stats()["kernel_launches"] >= 15


@pytest.mark.parametrize("seed", [5])
def test_bucketed_unsorted_path_matches_oracle(seed, monkeypatch):
    """Unsorted batches at scale go expand -> bucket -> tile (pd_bucket.cuh) instead of global reductions; forced here for
    small batches.  Mixed CIGARs (far N skips, clips, I/D/=/X/P), the device-side filter, pile-ups, reads over the contig
    ends, several contigs (one of a single base), two unsorted batches on top of each other and on top of a sorted one,
    and range-sharded engines whose ranges start inside a contig."""
    import pandepth_b200 as pd
    from pandepth_b200.shard import assign_ranges
    monkeypatch.setenv("PANDEPTH_BUCKET_MIN_READS", "1")
    rng = np.random.default_rng(seed)
    lens = [700_001, 1, 90_000, 2_500]                                           # contig 0 spans two coarse buckets (2^19 bases each)
    names = [f"c{i}" for i in range(len(lens))]
    rows = []
    for t, ln in enumerate(lens):
        n = 0 if ln < 10 else ln // 6                                            # (every read is its own tiny batch below: keep the count moderate)
        starts = rng.integers(-100, ln, n)
        starts[rng.integers(0, max(1, n), n // 12)] = ln // 2                    # a pile-up
        for p in starts.tolist():
            k = rng.integers(0, 10)
            cs = ("150M", "150M", "150M", f"{rng.integers(1, 40)}S{rng.integers(1, 120)}M", f"{rng.integers(1, 80)}M{rng.integers(1, 6000)}N{rng.integers(1, 80)}M",
                  f"{rng.integers(1, 60)}={rng.integers(1, 9)}I{rng.integers(1, 60)}X{rng.integers(1, 30)}D{rng.integers(1, 60)}M2H", f"3D{rng.integers(1, 100)}M1P{rng.integers(1, 100)}M",
                  f"{rng.integers(1, 900)}M", "20M", "*")[k]
            if p + 1 < 1:
                continue
            rows.append((int(rng.choice([0, 0, 16, 1024, 256, 4])), t, p + 1, int(rng.integers(0, 61)), cs))
    rec = model.records_from_tuples(rows, names)
    tlen = np.array(lens, np.uint32)
    perm = rng.permutation(rec.n)
    for xflags, q in ((1796, -1), (0, 25)):
        want = model.accumulate(tlen, rec, xflags, q)
        with _engine(tlen, xflags=xflags, min_mapq=q) as eng:                    # one unsorted batch per contig
            _push_all(eng, rec.select(perm), False, 1)
            eng.finalize()
            _check_depth(eng, tlen, want)
        with _engine(tlen, xflags=xflags, min_mapq=q) as eng:                    # three unsorted pieces + nothing else: additive
            _push_all(eng, rec.select(perm), False, 3)
            eng.finalize()
            _check_depth(eng, tlen, want)
        half = rec.n // 2
        first_half = rec.select(perm[:half])
        with _engine(tlen, xflags=xflags, min_mapq=q) as eng:                    # a sorted batch first, unsorted on top of it
            eng.set_reduce_hint(1, 0xFFFFFFFF)
            _push_all(eng, first_half.select(_sorted_order(first_half)), True, 2)
            _push_all(eng, rec.select(perm[half:]), False, 2)
            eng.finalize()
            _check_depth(eng, tlen, want)
            it = np.arange(len(lens), dtype=np.int32)[tlen > 1]
            cov, tot = eng.reduce(it, np.zeros(it.size, np.int32), tlen[it].astype(np.int32), 1, 0xFFFFFFFF)
            assert [int(x) for x in tot] == [int(want[t].sum(dtype=np.uint64)) for t in it]
    # range sharding: three engines, every one fed everything unsorted
    beg, end = assign_ranges(lens, 3)
    want = model.accumulate(tlen, rec)
    total = [np.zeros(int(n), np.uint64) for n in lens]
    for r in range(3):
        with pd.Engine(tlen, own_beg=beg[r], own_end=end[r]) as eng:
            _push_all(eng, rec.select(perm), False, 2)
            eng.finalize()
            for t, ln in enumerate(lens):
                if ln:
                    total[t] += eng.depth(t)
    for t in range(len(lens)):
        assert np.array_equal(total[t].astype(np.uint32), want[t]), f"range-sharded bucketed path, contig {t}"


def test_other_read_lengths_and_no_common_span():
    """Read lengths above the modelled span (L > HALO) and batches with no dominant span fall back to corrections."""
    rng = np.random.default_rng(9)
    L = 400_000
    tlen = np.array([L], np.uint32)
    import pandepth_b200 as pd
    for readlen in (36, 250, 256, 257, 300, 1000, 5000):
        pos, cig = _uniform_reads(rng, L, 40_000, readlen)
        rec = model.Records(np.zeros(pos.size, np.int32), pos, np.zeros(pos.size, np.uint16), np.zeros(pos.size, np.uint8),
                            np.arange(pos.size + 1, dtype=np.uint32), cig)
        want = model.accumulate(tlen, rec)
        with _engine(tlen) as eng:
            eng.push_contig(0, pos, cig, flags=pd.PD_SORTED)
            eng.finalize()
            _check_depth(eng, tlen, want)
    # every read a different length
    pos = np.sort(rng.integers(0, L - 400, 50_000)).astype(np.int32)
    cig = ((rng.integers(1, 400, pos.size).astype(np.uint32)) << 4) | np.uint32(0)
    rec = model.Records(np.zeros(pos.size, np.int32), pos, np.zeros(pos.size, np.uint16), np.zeros(pos.size, np.uint8),
                        np.arange(pos.size + 1, dtype=np.uint32), cig)
    want = model.accumulate(tlen, rec)
    with _engine(tlen) as eng:
        eng.push_contig(0, pos, cig, flags=pd.PD_SORTED)
        eng.finalize()
        _check_depth(eng, tlen, want)


def test_18bit_wrap_and_uint32_paths():
    """SURVEY App. C `w.sam`: 262149 stacked reads; mask 0x3FFFF reproduces the SiteInfo bit-field."""
    targets, rec, _ = load_fixture("@stack")
    tlen = np.array([1000], np.uint32)
    import pandepth_b200 as pd
    with _engine(tlen) as eng:
        eng.push_contig(0, rec.pos, rec.cigar, flags=pd.PD_SORTED)
        cov, tot = eng.reduce([0], [0], [1000], 1, 0xFFFFFFFF)
        assert (int(cov[0]), int(tot[0])) == (10, 2621490)
        cov, tot = eng.reduce([0], [0], [1000], 1, 0x3FFFF)
        assert (int(cov[0]), int(tot[0])) == (10, 50)
        cov, tot = eng.reduce_windows(0, 0, 100, 10, 1, 0x3FFFF)
        assert cov.tolist()[:2] == [1, 9] and tot.tolist()[:2] == [5, 45]


@pytest.mark.parametrize("mask", [0xFFFFFFFF, 0x3FFFF])
@pytest.mark.parametrize("min_dep", [1, 3])
def test_reduce_windows_every_width_branch(mask, min_dep):
    """k_reduce_windows on the device at the widths each of its branches covers: lane-per-window (< 32), the 4-window groups
    with 128-bit loads (multiples of 4 up to 128), the five-load lanes (129..160, reachable from the CLI with -w 129..149),
    the 32-bit lane-strided branch (161..512) and the generic 64-bit one (> 512) -- against the oracle's interval reduce,
    on a contig whose length leaves a ragged last window and a depth profile that crosses the 18-bit wrap."""
    import pandepth_b200 as pd
    rng = np.random.default_rng(77)
    L = 300_007
    tlen = np.array([L, 1000], np.uint32)
    pos, cig = _uniform_reads(rng, L, 30_000, 150)
    # a pile of 262150 stacked reads: depth there exceeds 2^18, so the two masks disagree
    pile_pos = np.full(262_150, 150_000, np.int32); pile_cig = np.full(262_150, (40 << 4), np.uint32)
    with _engine(tlen) as eng:
        eng.push_contig(0, pos, cig, flags=pd.PD_SORTED)
        eng.push_contig(0, pile_pos, pile_cig, flags=pd.PD_SORTED)
        eng.push_contig(1, np.array([5, 900], np.int32), np.array([(200 << 4), (200 << 4)], np.uint32), flags=pd.PD_SORTED)
        eng.finalize()
        depth = [eng.depth(0), eng.depth(1)]
        rec = model.Records(np.zeros(pos.size, np.int32), pos, np.zeros(pos.size, np.uint16), np.zeros(pos.size, np.uint8),
                            np.arange(pos.size + 1, dtype=np.uint32), cig)
        want0 = model.accumulate(tlen, rec)[0]
        want0[150_000:150_040] += 262_150
        assert np.array_equal(depth[0], want0)
        for win in (1, 7, 31, 32, 33, 64, 100, 128, 129, 130, 149, 150, 159, 160, 161, 200, 511, 512, 513, 1000, 4096, 70_001):
            for t, ln in ((0, L), (1, 1000)):
                s0, e0 = model.mode6_windows(ln, win)
                cov_o, sum_o = model.reduce_intervals(depth[t], s0, e0, min_dep, mask)
                cov_w, sum_w = eng.reduce_windows(t, 0, win, s0.size, min_dep, mask)
                assert np.array_equal(cov_o, cov_w.astype(np.uint64)), (win, t, "cov")
                assert np.array_equal(sum_o, sum_w), (win, t, "sum")
            # a start that is not 16-byte aligned takes the scalar lanes of the same branches
            n = (L - 3) // win
            s0 = 3 + np.arange(n, dtype=np.int64) * win
            cov_o, sum_o = model.reduce_intervals(depth[0], s0.astype(np.int32), np.minimum(s0 + win, L).astype(np.int32), min_dep, mask)
            cov_w, sum_w = eng.reduce_windows(0, 3, win, n, min_dep, mask)
            assert np.array_equal(cov_o, cov_w.astype(np.uint64)) and np.array_equal(sum_o, sum_w), (win, "offset start")


def test_reset_after_an_unsorted_error_clears_far_run_deltas():
    """A PD_SORTED batch whose tile kernel wrote far-run deltas, then a batch that trips PD_ERR_UNSORTED, then pd_reset:
    the deltas of the first batch must not leak into the next input (the device flag is folded before the error returns)."""
    import pandepth_b200 as pd
    rng = np.random.default_rng(41)
    L = 200_000
    tlen = np.array([L], np.uint32)
    rows = [(0, 0, int(p) + 1, 60, f"50M{rng.integers(3000, 9000)}N50M") for p in np.sort(rng.integers(0, L - 20_000, 5000))]
    far = model.records_from_tuples(rows, ["a"])
    with _engine(tlen) as eng:
        _push_all(eng, far, True, 1)
        bad_pos = np.array([5000, 100, 90_000, 7], np.int32)
        with pytest.raises(pd.PDError) as ei:
            eng.push_contig(0, bad_pos, np.full(4, (50 << 4), np.uint32), flags=pd.PD_SORTED)
            eng.finalize()
        assert ei.value.code == 4
        eng.reset()
        pos, cig = _uniform_reads(rng, L, 10_000)
        rec = model.Records(np.zeros(pos.size, np.int32), pos, np.zeros(pos.size, np.uint16), np.zeros(pos.size, np.uint8),
                            np.arange(pos.size + 1, dtype=np.uint32), cig)
        # an unsorted batch forces the scan, which would fold any stale deltas into depth
        perm = rng.permutation(pos.size)
        eng.push_contig(0, pos[perm], cig[perm], flags=0)
        eng.finalize()
        _check_depth(eng, tlen, model.accumulate(tlen, rec))


def test_empty_ragged_and_reset():
    import pandepth_b200 as pd
    tlen = np.array([5000, 0, 1, 123456], np.uint32)
    with _engine(tlen) as eng:
        eng.finalize()                                            # no input at all
        assert not eng.depth(0).any() and not eng.depth(3).any()
        cov, tot = eng.reduce([0, 3, 2], [0, -5, 0], [5000, 999999, 1])
        assert cov.tolist() == [0, 0, 0] and tot.tolist() == [0, 0, 0]
        eng.push_contig(3, np.array([], np.int32), np.array([], np.uint32), flags=pd.PD_SORTED)   # empty batch
        eng.push_contig(3, np.array([10, 10, 123400], np.int32), np.array([(100 << 4), (100 << 4) | 7, (100 << 4) | 8], np.uint32),
                        flags=pd.PD_SORTED)
        d = eng.depth(3)
        assert d[10:110].tolist() == [2] * 100 and d[110] == 0 and d[123400:].tolist() == [1] * 56   # clipped at the contig end
        eng.reset()
        assert not eng.depth(3).any()
        eng.push_contig(0, np.array([4990], np.int32), np.array([(5 << 4) | 4], np.uint32), flags=0)   # soft clip only
        assert not eng.depth(0).any()


def test_unsorted_batch_declared_sorted_is_reported():
    import pandepth_b200 as pd
    tlen = np.array([100_000], np.uint32)
    pos = np.array([5000, 100, 90_000, 7], np.int32)
    cig = np.full(4, (50 << 4), np.uint32)
    with _engine(tlen) as eng:
        with pytest.raises(pd.PDError) as ei:
            eng.push_contig(0, pos, cig, flags=pd.PD_SORTED)
            eng.finalize()
        assert ei.value.code == 4
    # sorted except for two swapped reads in the middle: first <= last, so only the device can see it
    rng = np.random.default_rng(3)
    pos = np.sort(rng.integers(0, 99_000, 20_000)).astype(np.int32)
    pos[5000], pos[15000] = pos[15000], pos[5000]
    cig = np.full(pos.size, (150 << 4), np.uint32)
    with _engine(tlen) as eng:
        with pytest.raises(pd.PDError) as ei:
            eng.push_contig(0, pos, cig, flags=pd.PD_SORTED)
            eng.finalize()
        assert ei.value.code == 4


def test_contig_sharding_two_engines():
    """Contig sharding: two engines that each own a subset reproduce the single-engine triples."""
    rng = np.random.default_rng(5)
    lens = [200_000, 150_000, 90_000, 10_000]
    tlen = np.array(lens, np.uint32)
    import pandepth_b200 as pd
    batches = [_uniform_reads(rng, ln, ln // 5) for ln in lens]
    tid = np.arange(4, dtype=np.int32)
    s0 = np.zeros(4, np.int32); e0 = np.array(lens, np.int32)
    with _engine(tlen) as eng:
        for t, (p, c) in enumerate(batches):
            eng.push_contig(t, p, c, flags=pd.PD_SORTED)
        full = eng.reduce(tid, s0, e0)
    from pandepth_b200.shard import assign_contigs
    owner = assign_contigs(lens, 2)
    parts = []
    for r in range(2):
        own = (np.array(owner) == r).astype(np.uint8)
        with _engine(tlen, owned=own) as eng:
            for t, (p, c) in enumerate(batches):
                eng.push_contig(t, p, c, flags=pd.PD_SORTED)       # not-owned contigs are ignored by the engine
            parts.append(eng.reduce(tid, s0, e0))
    cov = parts[0][0] + parts[1][0]; tot = parts[0][1] + parts[1][1]
    assert np.array_equal(cov, full[0]) and np.array_equal(tot, full[1])


@pytest.mark.parametrize("style", ["one_push", "chunked", "unsorted", "mixed_paths", "spliced_far_runs"])
def test_reduce_hint_partials_give_identical_results(style):
    """pd_set_reduce_hint is purely an optimisation: whole written-once segments come from partials recorded by the tile
    kernel, everything else (edges, re-written segments, scan path, other (min_dep, mask) pairs) still reads depth."""
    import pandepth_b200 as pd
    rng = np.random.default_rng(21)
    lens = [1_000_003, 40_000, 4096, 2048, 300]
    tlen = np.array(lens, np.uint32)
    rows = []
    for t, ln in enumerate(lens):
        n = ln // 3
        starts = np.sort(rng.integers(0, max(1, ln - 150), n))
        for p in starts.tolist():
            if style == "spliced_far_runs" and rng.integers(0, 4) == 0:
                cs = f"{rng.integers(20, 80)}M{rng.integers(300, 20000)}N{rng.integers(20, 80)}M"
            else:
                cs = "150M" if rng.integers(0, 5) else f"{rng.integers(1, 149)}M"
            rows.append((0, t, p + 1, 60, cs))
    rec = model.records_from_tuples(rows, [f"c{i}" for i in range(len(lens))])
    want = model.accumulate(tlen, rec)
    # intervals: whole contigs, 10 kb tiles, odd-aligned pieces, tiny pieces
    iv = []
    for t, ln in enumerate(lens):
        iv.append((t, 0, ln))
        for s in range(0, ln, 10_000):
            iv.append((t, s, min(ln, s + 10_000)))
        iv += [(t, 1, ln - 1), (t, 2047, min(ln, 2049 + 4096)), (t, ln // 2, ln // 2 + 1), (t, -5, ln + 99)]
    it = np.array([x[0] for x in iv], np.int32); s0 = np.array([x[1] for x in iv], np.int32); e0 = np.array([x[2] for x in iv], np.int32)
    for (hd, hm), (md, mk) in (((1, 0xFFFFFFFF), (1, 0xFFFFFFFF)), ((3, 0x3FFFF), (3, 0x3FFFF)), ((1, 0xFFFFFFFF), (2, 0xFFFFFFFF)), ((0, 0xFFFFFFFF), (0, 0xFFFFFFFF))):
        with _engine(tlen) as eng:
            eng.set_reduce_hint(hd, hm)
            if style in ("one_push", "spliced_far_runs"):
                _push_all(eng, rec, True, 1)
            elif style == "chunked":
                _push_all(eng, rec, True, 9)
            elif style == "unsorted":
                _push_all(eng, rec.select(rng.permutation(rec.n)), False, 2)
            else:
                half = rec.n // 2
                _push_all(eng, rec.select(np.arange(0, half)), True, 1)
                _push_all(eng, rec.select(rng.permutation(np.arange(half, rec.n))), False, 1)
            cov, tot = eng.reduce(it, s0, e0, md, mk)
            for k, (t, a, b) in enumerate(iv):
                co, so = model.reduce_intervals(want[t], [a], [b], md, mk)
                assert (int(cov[k]), int(tot[k])) == (int(co[0]), int(so[0])), (style, (hd, hm), (md, mk), iv[k])
            _check_depth(eng, tlen, want)
            with pytest.raises(pd.PDError):                   # the hint must precede the first push
                eng.set_reduce_hint(1, 0xFFFFFFFF)
            eng.reset()
            eng.set_reduce_hint(1, 0xFFFFFFFF)


@pytest.mark.parametrize("second", ["unsorted", "sorted_far", "sorted_near", "reset"])
def test_batches_after_a_scan_see_a_clean_delta_array(second):
    """k_scan_add leaves the consumed deltas in place; the engine must clear them before anything writes deltas again:
    a later scatter batch, a later tile batch with far runs, a tile batch followed by a second finalize, and pd_reset."""
    import pandepth_b200 as pd
    rng = np.random.default_rng(33)
    lens = [250_000, 9_000]
    tlen = np.array(lens, np.uint32)

    def make(n_frac, far):
        rows = []
        for t, ln in enumerate(lens):
            for p in np.sort(rng.integers(0, ln - 150, int(ln * n_frac))).tolist():
                cs = f"{rng.integers(10, 70)}M{rng.integers(400, 9000)}N{rng.integers(10, 70)}M" if far and rng.integers(0, 3) == 0 else "150M"
                rows.append((0, t, p + 1, 60, cs))
        return model.records_from_tuples(rows, ["a", "b"])

    first = make(0.2, True)
    with _engine(tlen) as eng:
        _push_all(eng, first.select(rng.permutation(first.n)), False, 2)        # scatter path -> finalize runs the scan
        eng.finalize()
        assert eng.stats()["scan_passes"] == 1
        _check_depth(eng, tlen, model.accumulate(tlen, first))
        if second == "reset":
            eng.reset()
            nxt = make(0.1, True)
            _push_all(eng, nxt.select(rng.permutation(nxt.n)), False, 1)
            eng.finalize()
            _check_depth(eng, tlen, model.accumulate(tlen, nxt))
            return
        nxt = make(0.15, second != "sorted_near")
        if second == "unsorted":
            _push_all(eng, nxt.select(rng.permutation(nxt.n)), False, 1)
        else:
            _push_all(eng, nxt, True, 3)
        eng.finalize()
        both = model.accumulate(tlen, first)
        add = model.accumulate(tlen, nxt)
        _check_depth(eng, tlen, [a + b for a, b in zip(both, add)])
        # and once more on top: a third generation after the second scan / fill
        third = make(0.05, True)
        _push_all(eng, third, True, 1)
        eng.finalize()
        _check_depth(eng, tlen, [a + b + c for a, b, c in zip(both, add, model.accumulate(tlen, third))])


@pytest.mark.parametrize("style", ["sorted", "chunked", "unsorted", "device_decode"])
def test_range_sharded_engines_add_up_to_one_engine(style, tmp_path):
    """Range sharding (pd_create_ranges): three engines own disjoint thirds of the concatenated genome, every engine is
    fed ALL reads -- boundary reads included, also spliced reads that start thousands of bases left of a range and reach
    into it -- clips to its range, and the depth pieces / reduce results add up to the single-engine result."""
    import pandepth_b200 as pd
    from pandepth_b200.shard import assign_ranges
    rng = np.random.default_rng(17)
    lens = [200_003, 50_000, 1, 120_000]
    names = [f"c{i}" for i in range(len(lens))]
    rows = []
    for t, ln in enumerate(lens):
        n = 0 if ln < 10 else ln // 4
        for p in np.sort(rng.integers(0, ln, n)).tolist():
            k = rng.integers(0, 8)
            cs = "150M" if k < 5 else (f"{rng.integers(10, 80)}M{rng.integers(300, 9000)}N{rng.integers(10, 80)}M" if k < 7 else f"{rng.integers(1, 3000)}M")
            rows.append((0, t, p + 1, 60, cs))
    rec = model.records_from_tuples(rows, names)
    tlen = np.array(lens, np.uint32)
    want = model.accumulate(tlen, rec)
    world = 3
    beg, end = assign_ranges(lens, world)
    iv = []
    for t, ln in enumerate(lens):
        if ln < 2:
            continue
        iv.append((t, 0, ln))
        for s in range(0, ln, 9_000):
            iv.append((t, s, min(ln, s + 9_000)))
    it = np.array([x[0] for x in iv], np.int32); s0 = np.array([x[1] for x in iv], np.int32); e0 = np.array([x[2] for x in iv], np.int32)
    bam = None
    if style == "device_decode":
        import subprocess
        from tests import bgzf_util as bu
        sam = str(tmp_path / "r.sam")
        with open(sam, "w") as fh:
            fh.write("@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in zip(names, lens)))
            for fl, t, p1, mq, cs in rows:
                fh.write(f"r\t{fl}\t{names[t]}\t{p1}\t{mq}\t{cs}\t*\t0\t0\t*\t*\n")
        bampath = str(tmp_path / "r.bam")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "bamgen"), "-I", sam, "-o", bampath], check=True)
        bam = open(bampath, "rb").read()
    cov_sum = np.zeros(len(iv), np.uint64); tot_sum = np.zeros(len(iv), np.uint64)
    depth_sum = [np.zeros(int(n), np.uint64) for n in lens]
    for r in range(world):
        with pd.Engine(tlen, own_beg=beg[r], own_end=end[r]) as eng:
            eng.set_reduce_hint(1, 0xFFFFFFFF)
            if style == "sorted":
                _push_all(eng, rec, True, 1)
            elif style == "chunked":
                _push_all(eng, rec, True, 11)
            elif style == "unsorted":
                _push_all(eng, rec.select(rng.permutation(rec.n)), False, 2)
            else:
                from tests import bgzf_util as bu
                raw, first, recs, _ = bu.bam_payload(bam)
                tab = bu.block_table(bam)
                acc = 0; b0 = 0
                while b0 < tab.size and acc + int(tab["out_len"][b0]) <= first:
                    acc += int(tab["out_len"][b0]); b0 += 1
                buf = np.frombuffer(bam, np.uint8)
                for a in range(b0, tab.size, 5):
                    eng.push_bgzf(buf, tab[a:a + 5].copy(), first_skip=(first - acc) if a == b0 else 0, flags=pd.PD_ASYNC)
            cov, tot = eng.reduce(it, s0, e0, 1, 0xFFFFFFFF)
            cov_sum += cov; tot_sum += tot
            for t, ln in enumerate(lens):
                if ln == 0:
                    continue
                d = eng.depth(t)
                lo, hi = int(beg[r, t]), int(end[r, t])
                assert not d[:lo].any() and not d[max(hi, lo):].any(), (style, r, t, "depth outside the owned range")
                depth_sum[t] += d
            with pytest.raises(pd.PDError):
                t_part = next(t for t in range(len(lens)) if 0 < int(end[r, t]) - int(beg[r, t]) < lens[t])
                eng.reduce_windows(t_part, 0, 100, 10)
    for t in range(len(lens)):
        assert np.array_equal(depth_sum[t], want[t].astype(np.uint64)), (style, t)
    for k, (t, a, b) in enumerate(iv):
        co, so = model.reduce_intervals(want[t], [a], [b], 1, 0xFFFFFFFF)
        assert (int(cov_sum[k]), int(tot_sum[k])) == (int(co[0]), int(so[0])), (style, iv[k])
