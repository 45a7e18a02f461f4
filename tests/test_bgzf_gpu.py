"""
GPU parity of the device-side BGZF decode (pd_push_bgzf: k_bgzf_decode / k_bgzf_offsets / k_bgzf_compact feeding the tile and
delta-scatter kernels), through the C ABI: compressed BAM bytes in, per-base depth out, compared bit for bit with the oracle
(oracle/pd_oracle.c restatement of PanDepth.cpp:732-760) on the records of the same file.

Covers: the golden SAM fixtures as BAM; zlib strategies that produce stored / fixed-Huffman / Huffman-only / RLE blocks; tiny
blocks; a header that shares its block with records; chunked + PD_ASYNC pushes; chunks in reverse order; unsorted files and
contig changes inside blocks (per-record-tid scatter path); records straddling blocks (PD_ERR_FALLBACK); corrupt blocks.
"""
import os
import subprocess
import zlib

import numpy as np
import pytest

from oracle import model
from tests import bgzf_util as bu
from tests.helpers import GOLD, ROOT

pytestmark = pytest.mark.gpu

BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")
FILTERS = ((1796, -1), (0, 20), (3844, 5))


def _bam_of_sam(sam, tmp_path, name="x.bam"):
    bam = str(tmp_path / name)
    subprocess.run([BAMGEN, "-I", sam, "-o", bam], check=True)
    return open(bam, "rb").read()


def _chunks(data, per_chunk, first=None):
    """-> list of (block sub-table, first_skip) covering every block from the first record on."""
    if first is None:
        first = bu.bam_payload(data)[1]
    tab = bu.block_table(data)
    acc = 0
    b0 = 0
    while b0 < tab.size and acc + int(tab["out_len"][b0]) <= first:
        acc += int(tab["out_len"][b0])
        b0 += 1
    skip = first - acc
    out = []
    for a in range(b0, tab.size, per_chunk):
        out.append((tab[a:a + per_chunk].copy(), skip if a == b0 else 0))
    return out


def _decode_depth(data, tlen, xflags, q, per_chunk=1 << 30, flags=0, order=None, hint=None, crc=True):
    import pandepth_b200 as pd
    if crc:
        flags |= pd.PD_CRC                      # the buffers are whole files: every payload has its trailer behind it
    buf = np.frombuffer(data, np.uint8)
    ch = _chunks(data, per_chunk)
    if order == "reverse":
        ch = ch[::-1]
    with pd.Engine(tlen, xflags=xflags, min_mapq=q) as eng:
        if hint:
            eng.set_reduce_hint(*hint)
        for tab, skip in ch:
            eng.push_bgzf(buf, tab, first_skip=skip, flags=flags)
        eng.finalize()
        depth = [eng.depth(t) if tlen[t] else np.zeros(0, np.uint32) for t in range(len(tlen))]
        st = eng.stats()
    return depth, st


def _assert_depth(got, want, what):
    for t, (g, w) in enumerate(zip(got, want)):
        if w.size == 0:
            continue
        if not np.array_equal(g, w):
            bad = np.flatnonzero(g != w)
            raise AssertionError(f"{what}: contig {t}: {bad.size} bases differ, first at {bad[0]}: got {g[bad[0]]} want {w[bad[0]]}")


@pytest.mark.parametrize("fixture", ["k.sam", "mix.sam", "big.sam"])
def test_device_decode_matches_oracle_on_fixtures(fixture, tmp_path):
    import pandepth_b200 as pd
    targets, rec, _ = model.parse_sam(os.path.join(GOLD, fixture))
    tlen = np.array([t[1] for t in targets], np.uint32)
    data = _bam_of_sam(os.path.join(GOLD, fixture), tmp_path)
    for xflags, q in FILTERS:
        want = model.accumulate(tlen, rec, xflags, q)
        got, st = _decode_depth(data, tlen, xflags, q)
        _assert_depth(got, want, f"{fixture} x={xflags} q={q}")
        assert st["kernel_launches"] >= 3
        got, _ = _decode_depth(data, tlen, xflags, q, per_chunk=1, flags=pd.PD_ASYNC)
        _assert_depth(got, want, f"{fixture} x={xflags} q={q} one block per chunk, async")


@pytest.fixture(scope="module")
def mixed(tmp_path_factory):
    """A random mixed-CIGAR BAM (bamgen -m mixed) with its records parsed from the SAM text of the same generator run."""
    d = tmp_path_factory.mktemp("bz")
    contigs = "c1:400000,c2:1,c3:150000,c4:2100,c5:70000"
    bam, sam = str(d / "m.bam"), str(d / "m.sam")
    subprocess.run([BAMGEN, "-o", sam, "-c", contigs, "-n", "120000", "-s", "5", "-l", "100", "-m", "mixed", "-S"], check=True)
    subprocess.run([BAMGEN, "-I", sam, "-o", bam], check=True)
    targets, rec, _ = model.parse_sam(sam)
    return open(bam, "rb").read(), np.array([t[1] for t in targets], np.uint32), rec


VARIANTS = {
    "stored": dict(level=0),
    "fixed": dict(level=6, strategy=zlib.Z_FIXED),
    "huffman_only": dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY),
    "rle": dict(level=6, strategy=zlib.Z_RLE),
    "level9": dict(level=9),
    "level1": dict(level=1),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_device_decode_every_deflate_block_type(variant, mixed):
    import pandepth_b200 as pd
    data, tlen, rec = mixed
    raw, first, recs, _ = bu.bam_payload(data)
    # level9: the header shares its block with the first records (first_skip > 0)
    cuts = bu.record_cuts(raw, first, recs, max_block=20000 if variant != "level1" else 64000, header_own_block=variant != "level9")
    re = bu.rewrite_bgzf(raw, cuts, **VARIANTS[variant])
    xflags, q = FILTERS[hash(variant) % 3] if False else FILTERS[sorted(VARIANTS).index(variant) % 3]
    want = model.accumulate(tlen, rec, xflags, q)
    got, st = _decode_depth(re, tlen, xflags, q, per_chunk=7, flags=pd.PD_ASYNC)
    _assert_depth(got, want, variant)
    got, _ = _decode_depth(re, tlen, xflags, q, per_chunk=11, order="reverse")
    _assert_depth(got, want, variant + " reverse chunk order")


def test_device_decode_tiny_blocks_and_hint(mixed):
    """One record per BGZF block (every block boundary is a run boundary) + the reduce hint on."""
    data, tlen, rec = mixed
    raw, first, recs, _ = bu.bam_payload(data)
    n = 3000
    small = raw[:recs[n]]
    cuts = [0, first] + recs[1:n] + [len(small)]
    re = bu.rewrite_bgzf(small, cuts, level=6)
    sub = rec.select(np.arange(n))
    want = model.accumulate(tlen, sub, 1796, -1)
    got, _ = _decode_depth(re, tlen, 1796, -1, per_chunk=500, hint=(1, 0xFFFFFFFF))
    _assert_depth(got, want, "tiny blocks")


def test_device_decode_unsorted_file(mixed, tmp_path):
    """Shuffled records: blocks are unsorted inside and mix contigs -> the per-record-tid scatter path."""
    data, tlen, rec = mixed
    raw, first, recs, _ = bu.bam_payload(data)
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(recs))
    ends = recs[1:] + [len(raw)]
    body = b"".join(raw[recs[i]:ends[i]] for i in perm)
    shuf = raw[:first] + body
    # record starts of the shuffled stream
    lens = np.array([ends[i] - recs[i] for i in perm], np.int64)
    starts = (first + np.concatenate([[0], np.cumsum(lens)[:-1]])).tolist()
    cuts = bu.record_cuts(shuf, first, starts, max_block=30000)
    re = bu.rewrite_bgzf(shuf, cuts, level=4)
    for xflags, q in FILTERS[:2]:
        want = model.accumulate(tlen, rec, xflags, q)
        got, st = _decode_depth(re, tlen, xflags, q, per_chunk=9)
        _assert_depth(got, want, "shuffled")
        assert st["delta_path_batches"] > 0 and st["scan_passes"] == 1


def test_device_decode_then_host_pushes_add_up(mixed):
    """Depth is additive across entry points: a file decoded on the device plus the same records pushed from the host."""
    import pandepth_b200 as pd
    data, tlen, rec = mixed
    want = [2 * w for w in model.accumulate(tlen, rec, 1796, -1)]
    buf = np.frombuffer(data, np.uint8)
    with pd.Engine(tlen) as eng:
        for tab, skip in _chunks(data, 13):
            eng.push_bgzf(buf, tab, first_skip=skip, flags=pd.PD_ASYNC)
        order = np.lexsort((rec.pos, rec.tid))
        r = rec.select(order)
        eng.push(r.tid, r.pos, r.cigar, flag=r.flag, mapq=r.mapq, cigar_off=r.cigar_off, flags=pd.PD_SORTED)
        eng.finalize()
        got = [eng.depth(t) if tlen[t] else np.zeros(0, np.uint32) for t in range(len(tlen))]
    _assert_depth(got, want, "device decode + host push")


def test_device_decode_many_small_contigs(tmp_path):
    """Thousands of contigs, a few reads each: contig changes inside every block."""
    n_contigs = 3000
    contigs = ",".join(f"s{i}:{2000 + (i % 7) * 300}" for i in range(n_contigs))
    bam, sam = str(tmp_path / "many.bam"), str(tmp_path / "many.sam")
    subprocess.run([BAMGEN, "-o", sam, "-c", contigs, "-n", "60000", "-s", "9", "-l", "100", "-m", "mixed", "-S"], check=True)
    subprocess.run([BAMGEN, "-I", sam, "-o", bam], check=True)
    targets, rec, _ = model.parse_sam(sam)
    tlen = np.array([t[1] for t in targets], np.uint32)
    want = model.accumulate(tlen, rec, 1796, -1)
    got, st = _decode_depth(open(bam, "rb").read(), tlen, 1796, -1, per_chunk=40)
    _assert_depth(got, want, "many contigs")
    assert st["kernel_launches"] < 400          # launches per chunk, not per contig


def test_device_decode_contigs_of_a_few_blocks_each(tmp_path):
    """Hundreds of contigs spanning two or three BGZF blocks each: more than 32 single-contig runs in a chunk, so the short
    runs join the per-record-tid scatter (one launch) instead of three launches each; long runs keep the tile path."""
    n_contigs = 300
    contigs = ",".join(f"m{i}:{40000 + (i % 5) * 9000}" for i in range(n_contigs)) + ",big:2500000"
    bam, sam = str(tmp_path / "mid.bam"), str(tmp_path / "mid.sam")
    subprocess.run([BAMGEN, "-o", sam, "-c", contigs, "-n", "700000", "-s", "4", "-l", "100", "-S"], check=True)
    subprocess.run([BAMGEN, "-I", sam, "-o", bam], check=True)
    targets, rec, _ = model.parse_sam(sam)
    tlen = np.array([t[1] for t in targets], np.uint32)
    want = model.accumulate(tlen, rec, 1796, -1)
    got, st = _decode_depth(open(bam, "rb").read(), tlen, 1796, -1)
    _assert_depth(got, want, "contigs of a few blocks")
    assert st["tile_path_batches"] >= 1 and st["delta_path_batches"] >= 1
    assert st["kernel_launches"] < 200


def test_records_straddling_blocks_ask_for_the_host_path(mixed):
    import pandepth_b200 as pd
    data, tlen, rec = mixed
    raw, first, recs, _ = bu.bam_payload(data)
    cuts = list(range(0, len(raw), 30000)) + [len(raw)]
    re = bu.rewrite_bgzf(raw, cuts, level=6)
    buf = np.frombuffer(re, np.uint8)
    with pd.Engine(tlen) as eng:
        with pytest.raises(pd.PDError) as ei:
            for tab, skip in _chunks(re, 1 << 30):
                eng.push_bgzf(buf, tab, first_skip=skip)
            eng.finalize()
        assert ei.value.code == 6               # PD_ERR_FALLBACK
        # the engine is usable again after a reset
        eng.reset()
        order = np.lexsort((rec.pos, rec.tid))
        r = rec.select(order)
        eng.push(r.tid, r.pos, r.cigar, flag=r.flag, mapq=r.mapq, cigar_off=r.cigar_off, flags=pd.PD_SORTED)
        eng.finalize()
        got = [eng.depth(t) if tlen[t] else np.zeros(0, np.uint32) for t in range(len(tlen))]
    _assert_depth(got, model.accumulate(tlen, rec, 1796, -1), "after fallback + reset")


def test_corrupt_block_is_an_error(mixed):
    import pandepth_b200 as pd
    data, tlen, _ = mixed
    first = bu.bam_payload(data)[1]
    bad = bytearray(data)
    tab = bu.block_table(data)
    mid = tab[tab.size // 2]
    for k in range(40, 60):                     # garble the middle of one payload
        bad[int(mid["in_off"]) + k] ^= 0x5A
    buf = np.frombuffer(bytes(bad), np.uint8)
    with pd.Engine(tlen) as eng:
        with pytest.raises(pd.PDError) as ei:
            for t, skip in _chunks(bytes(bad), 1 << 30, first):
                eng.push_bgzf(buf, t, first_skip=skip)
            eng.finalize()
        assert ei.value.code in (1, 6)          # does not inflate to ISIZE (PD_ERR_ARG) or garbage records (PD_ERR_FALLBACK)


@pytest.mark.parametrize("K", ["1", "2", "4", "8"])
def test_device_decode_with_k_decoders_per_warp(K, mixed, monkeypatch):
    """PANDEPTH_BGZF_K forces the number of decoders per warp (1: k_bgzf_decode; 2 / 4 / 8: k_bgzf_decode_multi, the lockstep
    decoder the engine picks by itself for streams that inflate by less than 4x): same depth from every variant, on the
    mixed BAM as written, re-compressed Huffman-only (literal runs only) and as stored blocks."""
    import pandepth_b200 as pd
    monkeypatch.setenv("PANDEPTH_BGZF_K", K)
    data, tlen, rec = mixed
    raw, first, recs, _ = bu.bam_payload(data)
    want = model.accumulate(tlen, rec, 1796, -1)
    got, _ = _decode_depth(data, tlen, 1796, -1, per_chunk=17, flags=pd.PD_ASYNC)
    _assert_depth(got, want, f"K={K} as written")
    for name in ("huffman_only", "stored", "fixed"):
        re = bu.rewrite_bgzf(raw, bu.record_cuts(raw, first, recs, max_block=15000), **VARIANTS[name])
        got, _ = _decode_depth(re, tlen, 1796, -1, per_chunk=23)
        _assert_depth(got, want, f"K={K} {name}")


@pytest.mark.parametrize("K", ["1", "2", "4", "8"])
def test_device_decode_literal_dominated_bam_with_k_decoders(K, tmp_path, monkeypatch):
    """Flavour B of SURVEY 8(d) (random SEQ + QUAL: literal runs, short matches the lane-group leaders copy themselves, 14-bit
    codes behind the first-level table): every decoder variant against depth computed from the generator's records."""
    import pandepth_b200 as pd
    monkeypatch.setenv("PANDEPTH_BGZF_K", K)
    bam = str(tmp_path / "b.bam")
    subprocess.run([BAMGEN, "-P", "-q", "-o", bam, "-c", "c1:400000,c2:250000", "-n", "12000", "-s", "11", "-t", "2"], check=True)
    data = open(bam, "rb").read()
    raw, first, recs, _ = bu.bam_payload(data)
    tlen = np.array([400000, 250000], np.uint32)
    want = [np.zeros(int(n) + 1, np.int64) for n in tlen]
    ends = recs[1:] + [len(raw)]
    for a in recs:
        tid, pos = np.frombuffer(raw, np.int32, 2, a + 4)
        w = int(np.frombuffer(raw, np.uint32, 1, a + 36 + raw[a + 12])[0])       # the single CIGAR word (150M)
        assert w & 15 == 0
        want[tid][pos] += 1; want[tid][min(pos + (w >> 4), int(tlen[tid]))] -= 1
    want = [np.cumsum(w)[:-1].astype(np.uint32) for w in want]
    got, _ = _decode_depth(data, tlen, 1796, -1, per_chunk=5, flags=pd.PD_ASYNC)
    _assert_depth(got, want, f"flavour B, K={K}")
    assert sum(int(g.sum(dtype=np.uint64)) for g in got) == 150 * len(recs)


def test_crc32_catches_a_flipped_byte_that_keeps_the_stream_valid(mixed):
    """A stored (level 0) block with one payload byte changed -- the low byte of a record's pos -- inflates to its ISIZE and
    still parses as records, so only the CRC32 of the trailer can tell.  With PD_CRC the engine reports the mismatch (the
    reference, through htslib's bgzf_read_block, fails on such a file too); without the flag the changed depth comes out."""
    import pandepth_b200 as pd
    data, tlen, rec = mixed
    raw, first, recs, _ = bu.bam_payload(data)
    good = bu.rewrite_bgzf(raw, bu.record_cuts(raw, first, recs, max_block=15000), level=0)
    want = model.accumulate(tlen, rec, 1796, -1)
    got, _ = _decode_depth(good, tlen, 1796, -1)
    _assert_depth(got, want, "stored blocks, CRC verified")
    # flip one byte of the compressed file: a stored payload byte IS the inflated byte
    tab = bu.block_table(good)
    blk = tab[tab.size // 2]
    r = next(x for x in recs if x > first + 40000)                 # some record well inside the file
    bad = bytearray(good)
    # find the record's position inside its block: blocks hold raw[cuts[i]:cuts[i+1]] back to back
    acc = 0
    for b in tab:
        if acc + int(b["out_len"]) > r + 8:
            off = int(b["in_off"]) + 5 + (r + 8 - acc)            # 5 = stored-block header (BFINAL/BTYPE byte, LEN, NLEN)
            break
        acc += int(b["out_len"])
    assert bad[off] == raw[r + 8]
    bad[off] ^= 0x01
    with pytest.raises(pd.PDError) as ei:
        _decode_depth(bytes(bad), tlen, 1796, -1)
    assert ei.value.code == 1 and "CRC" in str(ei.value)
    got2, _ = _decode_depth(bytes(bad), tlen, 1796, -1, crc=False)  # without the check: silently different depth
    assert any(not np.array_equal(g, w) for g, w in zip(got2, want) if w.size)
