"""Test helpers for BGZF / BAM byte streams (TEST INFRASTRUCTURE).

read_bgzf_blocks   parse the BGZF block table of a file (payload offset / length, inflated size)
bam_payload        inflate a BAM, return (inflated bytes, offset of the first record, list of record start offsets)
rewrite_bgzf       write a byte stream as BGZF blocks compressed with a chosen zlib level / strategy, cut at given offsets
                   (stored blocks, fixed-Huffman blocks, Huffman-only, tiny blocks, records straddling blocks ...)
"""
import struct
import zlib

import numpy as np

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def read_bgzf_blocks(data):
    """-> list of (in_off, in_len, out_len) for every non-empty block."""
    out = []
    p = 0
    data = memoryview(data).cast("B")           # plain Python ints whatever the caller holds (numpy uint8 scalars would wrap)
    n = len(data)
    while p + 18 <= n:
        if data[p] != 31 or data[p + 1] != 139:
            break
        xlen = data[p + 10] | data[p + 11] << 8
        bsize = 0
        x = 0
        while x + 4 <= xlen:
            e = p + 12 + x
            sl = data[e + 2] | data[e + 3] << 8
            if data[e] == 66 and data[e + 1] == 67 and sl == 2:
                bsize = (data[e + 4] | data[e + 5] << 8) + 1
                break
            x += 4 + sl
        if not bsize or p + bsize > n:
            break
        isize = struct.unpack_from("<I", data, p + bsize - 4)[0]
        if isize:
            out.append((p + 12 + xlen, bsize - xlen - 20, isize))
        p += bsize
    return out


def inflate_all(data):
    return b"".join(zlib.decompress(data[o:o + l], -15) for o, l, _ in read_bgzf_blocks(data))


def bam_payload(data):
    """-> (inflated bytes, first record offset, record start offsets, n_ref)."""
    raw = inflate_all(data)
    assert raw[:4] == b"BAM\1"
    q = 8 + struct.unpack_from("<I", raw, 4)[0]
    n_ref = struct.unpack_from("<I", raw, q)[0]
    q += 4
    for _ in range(n_ref):
        q += 8 + struct.unpack_from("<I", raw, q)[0]
    recs = []
    p = q
    while p + 4 <= len(raw):
        bs = struct.unpack_from("<I", raw, p)[0]
        recs.append(p)
        p += 4 + bs
    return raw, q, recs, n_ref


def bgzf_block(payload, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    assert len(payload) <= 65536
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    comp = co.compress(payload) + co.flush()
    if len(comp) + 26 > 65536:                      # does not fit a BGZF block: store instead
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        comp = co.compress(payload) + co.flush()
    bsize = len(comp) + 25
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return hdr + comp + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


def rewrite_bgzf(raw, cuts, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    """raw bytes -> BGZF file with blocks raw[cuts[i]:cuts[i+1]] (cuts must start at 0 and end at len(raw))."""
    out = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert 0 < b - a <= 65280
        out.append(bgzf_block(raw[a:b], level, strategy))
    out.append(BGZF_EOF)
    return b"".join(out)


def record_cuts(raw, first, recs, max_block=60000, header_own_block=True):
    """Block boundaries on record boundaries, blocks of at most max_block bytes."""
    cuts = [0]
    if header_own_block and first > 0:
        while first - cuts[-1] > 65280:
            cuts.append(cuts[-1] + 65280)
        cuts.append(first)
    last = cuts[-1]
    prev = None
    for r in recs + [len(raw)]:
        if r - last > max_block and prev is not None and prev > last:
            cuts.append(prev)
            last = prev
        while r - last > 65280:                   # one record larger than a block
            cuts.append(last + 65280)
            last += 65280
        prev = r
    if cuts[-1] != len(raw):
        cuts.append(len(raw))
    return cuts


def block_table(data):
    """numpy structured table for Engine.push_bgzf: (in_off u64, in_len u32, out_len u32)."""
    blk = read_bgzf_blocks(data)
    t = np.zeros(len(blk), dtype=np.dtype([("in_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4")]))
    for i, (o, l, n) in enumerate(blk):
        t[i] = (o, l, n)
    return t
