// pd_decode.hpp -- BGZF chunk -> per-contig (pos, cigar[, cigar_off]) arrays, block-parallel (groundwork, see pd_inflate.cuh).
//
// A chunk is a run of consecutive BGZF blocks that starts on a BAM record boundary.  Every step is a per-block function from
// pd_inflate.cuh applied to all blocks independently, so the same orchestration runs with a CUDA backend (one block per warp /
// thread) or, for verification, with a host backend that simply loops (tests/cpp/decode_check.cpp).  Steps:
//   1. inflate every block into one contiguous buffer                                   (inflate_raw)
//   2. count, per block, the records that pass the filter and their CIGAR words          (count_block)
//   3. host: check that every block ends on a record boundary, prefix-sum the counts, find the contig runs
//   4. emit pos / cigar / cigar_off / tid for every passing record at its final index    (emit_block)
// The result is what pd_push_contig(PD_DEVICE | PD_SORTED) takes, one call per contig run.
// Not handled here (the caller falls back to the host feeders): records that straddle blocks (reads larger than a block),
// CIGARs kept in the CG tag (> 65535 ops).
#pragma once
#include "pd_inflate.cuh"

#include <stddef.h>
#include <stdint.h>
#include <vector>

namespace pdk {

struct BlockCount {
    uint32_t n_reads;     // records passing the filter
    uint32_t n_cigar;     // their CIGAR words
    uint32_t used;        // bytes consumed by whole records (== block length when the block ends on a record boundary)
    int32_t first_tid, last_tid;   // refID of the first / last passing record (INT32_MIN when none)
    int32_t first_pos, last_pos;   // and their positions (order across block boundaries is checked on the host)
    uint32_t flags;       // bit 0: malformed record, bit 1: a CG-tag placeholder CIGAR was seen, bit 2: not sorted inside the block
};

// the record filter of PanDepth.cpp:732-733 plus "has a contig"
PD_HD bool decode_keep(int32_t tid, uint16_t flag, uint8_t mapq, int32_t n_targets, uint32_t xflags, int32_t min_mapq)
{
    return tid >= 0 && tid < n_targets && !(flag & xflags) && (int32_t)mapq >= min_mapq;
}

PD_HD BlockCount count_block(const uint8_t *buf, uint32_t len, int32_t n_targets, uint32_t xflags, int32_t min_mapq)
{
    BlockCount c; c.n_reads = 0; c.n_cigar = 0; c.first_tid = INT32_MIN; c.last_tid = INT32_MIN; c.first_pos = 0; c.last_pos = 0; c.flags = 0;
    int32_t prev_tid = INT32_MIN, prev_pos = INT32_MIN;
    const int64_t u = walk_bam_block(buf, len, [&](int32_t tid, int32_t pos, uint16_t flag, uint8_t mapq, uint32_t n_cigar, const uint8_t *cig) {
        if (!decode_keep(tid, flag, mapq, n_targets, xflags, min_mapq)) return;
        if (n_cigar == 2 && (cig[0] & 15u) == 4u && (cig[4] & 15u) == 3u) c.flags |= 2u;      // possible `<l_seq>S<ref_len>N` placeholder
        if (tid < prev_tid || (tid == prev_tid && pos < prev_pos)) c.flags |= 4u;
        prev_tid = tid; prev_pos = pos;
        if (c.first_tid == INT32_MIN) { c.first_tid = tid; c.first_pos = pos; }
        c.last_tid = tid; c.last_pos = pos;
        c.n_reads++; c.n_cigar += n_cigar;
    });
    if (u < 0) { c.flags |= 1u; c.used = 0; } else c.used = (uint32_t)u;
    return c;
}

// writes the block's passing records at read index r0.. and CIGAR word index c0..; cigar_off[i] is absolute into cigar[]
PD_HD void emit_block(const uint8_t *buf, uint32_t len, int32_t n_targets, uint32_t xflags, int32_t min_mapq,
                      uint32_t r0, uint32_t c0, int32_t *tid_out, int32_t *pos_out, uint32_t *cigar_off, uint32_t *cigar)
{
    uint32_t r = r0, c = c0;
    walk_bam_block(buf, len, [&](int32_t tid, int32_t pos, uint16_t flag, uint8_t mapq, uint32_t n_cigar, const uint8_t *cig) {
        if (!decode_keep(tid, flag, mapq, n_targets, xflags, min_mapq)) return;
        tid_out[r] = tid; pos_out[r] = pos; cigar_off[r] = c;
        for (uint32_t k = 0; k < n_cigar; k++) cigar[c + k] = inf_le32(cig + 4 * k);
        r++; c += n_cigar;
    });
}

struct ChunkBlock { uint64_t in_off; uint32_t in_len; uint32_t out_len; };      // one BGZF block of the chunk (payload only)

struct ContigRun { int32_t tid; uint32_t r0, r1; bool sorted; };                 // reads [r0, r1) of the chunk belong to tid

enum { DEC_OK = 0, DEC_INFLATE = -1, DEC_MALFORMED = -2, DEC_STRADDLE = -3, DEC_CG_TAG = -4 };

// Backend: where the per-block work runs.  Required members (all arrays are "device" arrays of the backend):
//   uint8_t  *inflated;  BlockCount *counts_host;                      // filled by inflate / count
//   int inflate(const uint8_t *in_host, size_t in_bytes, const ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint64_t out_total);
//   void count(const ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint32_t first_skip, int32_t n_targets, uint32_t xflags, int32_t min_mapq);
//   void emit(const ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint32_t first_skip, const uint32_t *r0, const uint32_t *c0,
//             uint32_t n_reads, uint32_t n_cigar, int32_t n_targets, uint32_t xflags, int32_t min_mapq);   // fills tid / pos / cigar_off / cigar
//   int32_t tid_at(uint32_t read_index);                                // one value back to the host (contig-run boundaries)
template <typename Backend>
int decode_chunk(Backend &be, const uint8_t *in_host, size_t in_bytes, const std::vector<ChunkBlock> &blocks, uint32_t first_skip,
                 int32_t n_targets, uint32_t xflags, int32_t min_mapq, std::vector<ContigRun> &runs, uint32_t &n_reads, uint32_t &n_cigar)
{
    runs.clear(); n_reads = 0; n_cigar = 0;
    const uint32_t n = (uint32_t)blocks.size();
    if (!n) return DEC_OK;
    std::vector<uint64_t> out_off(n);
    uint64_t out_total = 0;
    for (uint32_t b = 0; b < n; b++) { out_off[b] = out_total; out_total += blocks[b].out_len; }
    if (be.inflate(in_host, in_bytes, blocks.data(), out_off.data(), n, out_total) != 0) return DEC_INFLATE;
    be.count(blocks.data(), out_off.data(), n, first_skip, n_targets, xflags, min_mapq);
    const BlockCount *cnt = be.counts_host;
    std::vector<uint32_t> r0(n), c0(n);
    for (uint32_t b = 0; b < n; b++) {
        if (cnt[b].flags & 1u) return DEC_MALFORMED;
        if (cnt[b].flags & 2u) return DEC_CG_TAG;
        if (cnt[b].used + (b == 0 ? first_skip : 0) != blocks[b].out_len) return DEC_STRADDLE;
        r0[b] = n_reads; c0[b] = n_cigar;
        n_reads += cnt[b].n_reads; n_cigar += cnt[b].n_cigar;
    }
    if (!n_reads) return DEC_OK;
    be.emit(blocks.data(), out_off.data(), n, first_skip, r0.data(), c0.data(), n_reads, n_cigar, n_targets, xflags, min_mapq);
    // contig runs.  A block that is sorted inside and whose first and last passing records share a refID holds one contig and
    // needs no look inside; any other block (a contig change, or disorder) is walked through its (few) reads.  A run is
    // `sorted` when no block of it is unsorted inside and positions do not step back across its block boundaries.
    int32_t prev_tid = INT32_MIN, prev_pos = INT32_MIN;
    for (uint32_t b = 0; b < n; b++) {
        if (!cnt[b].n_reads) continue;
        const bool inside_ok = !(cnt[b].flags & 4u);
        const bool boundary_ok = !(cnt[b].first_tid == prev_tid && cnt[b].first_pos < prev_pos);
        if (inside_ok && cnt[b].first_tid == cnt[b].last_tid) {
            const int32_t t = cnt[b].first_tid;
            if (!runs.empty() && runs.back().tid == t) { runs.back().r1 = r0[b] + cnt[b].n_reads; runs.back().sorted = runs.back().sorted && boundary_ok; }
            else runs.push_back(ContigRun{t, r0[b], r0[b] + cnt[b].n_reads, true});
        } else {
            for (uint32_t r = r0[b]; r < r0[b] + cnt[b].n_reads; r++) {
                const int32_t t = be.tid_at(r);
                if (!runs.empty() && runs.back().tid == t) { runs.back().r1 = r + 1; runs.back().sorted = runs.back().sorted && inside_ok && (r > r0[b] || boundary_ok); }
                else runs.push_back(ContigRun{t, r, r + 1, inside_ok});
            }
        }
        prev_tid = cnt[b].last_tid; prev_pos = cnt[b].last_pos;
    }
    return DEC_OK;
}

}  // namespace pdk
