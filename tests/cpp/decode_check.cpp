// CPU checker for pandepth_b200/csrc/pd_decode.hpp with a HOST backend (the per-block functions are host+device; here
// they simply run in loops): a BAM is cut into chunks of consecutive BGZF blocks, every chunk goes through decode_chunk, and the
// emitted (tid, pos, CIGAR) records are compared -- count, order-independent digest and order -- with a plain sequential
// walk of the whole inflated stream under the same filter.  Also checks the contig runs (cover, refIDs, sortedness).
//   decode_check file.bam [xflags] [min_mapq] [blocks_per_chunk]
#include "pd_decode.hpp"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace pdk;

struct HostBackend {
    std::vector<uint8_t> inflated_v; std::vector<BlockCount> counts_v;
    std::vector<int32_t> tid, pos; std::vector<uint32_t> cigar_off, cigar;
    uint8_t *inflated = nullptr; BlockCount *counts_host = nullptr;
    int inflate(const uint8_t *in, size_t, const ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint64_t out_total)
    {
        inflated_v.assign(out_total + 8, 0); inflated = inflated_v.data();
        static InflateTables T;
        for (uint32_t b = 0; b < n; b++)
            if (blk[b].out_len && inflate_raw(in + blk[b].in_off, blk[b].in_len, inflated + out_off[b], blk[b].out_len, T) != (int)blk[b].out_len) return -1;
        return 0;
    }
    void count(const ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint32_t first_skip, int32_t n_targets, uint32_t xflags, int32_t min_mapq)
    {
        counts_v.resize(n); counts_host = counts_v.data();
        for (uint32_t b = 0; b < n; b++) { const uint32_t sk = b == 0 ? first_skip : 0; counts_v[b] = count_block(inflated + out_off[b] + sk, blk[b].out_len - sk, n_targets, xflags, min_mapq); }
    }
    void emit(const ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint32_t first_skip, const uint32_t *r0, const uint32_t *c0, uint32_t n_reads, uint32_t n_cigar,
              int32_t n_targets, uint32_t xflags, int32_t min_mapq)
    {
        tid.assign(n_reads, -7); pos.assign(n_reads, -7); cigar_off.assign(n_reads + 1, 0); cigar.assign(n_cigar, 0);
        for (uint32_t b = 0; b < n; b++) { const uint32_t sk = b == 0 ? first_skip : 0; emit_block(inflated + out_off[b] + sk, blk[b].out_len - sk, n_targets, xflags, min_mapq, r0[b], c0[b], tid.data(), pos.data(), cigar_off.data(), cigar.data()); }
        cigar_off[n_reads] = n_cigar;
    }
    int32_t tid_at(uint32_t r) { return tid[r]; }
};

static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const uint32_t xflags = argc > 2 ? (uint32_t)atoi(argv[2]) : 1796u; const int32_t min_mapq = argc > 3 ? atoi(argv[3]) : -1;
    const size_t per_chunk = argc > 4 ? (size_t)atoi(argv[4]) : 37;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<uint8_t> data; uint8_t tmp[1 << 16]; size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) data.insert(data.end(), tmp, tmp + k);
    fclose(f);
    // block table + the whole inflated stream through zlib (the reference walk)
    std::vector<ChunkBlock> blocks; std::vector<uint64_t> start; std::vector<uint8_t> all;
    for (size_t p = 0; p + 18 <= data.size();) {
        const uint8_t *h = data.data() + p;
        const uint32_t xlen = h[10] | h[11] << 8; uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) { const uint8_t *e = h + 12 + x; const uint32_t sl = e[2] | e[3] << 8; if (e[0] == 'B' && e[1] == 'C' && sl == 2) { bsize = (e[4] | e[5] << 8) + 1; break; } x += 4 + sl; }
        if (!bsize || p + bsize > data.size()) break;
        const uint32_t isize = inf_le32(h + bsize - 4);
        if (isize) {
            blocks.push_back(ChunkBlock{p + 12 + xlen, bsize - xlen - 20, isize}); start.push_back(all.size());
            const size_t o = all.size(); all.resize(o + isize);
            z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15);
            zs.next_in = (Bytef *)(h + 12 + xlen); zs.avail_in = bsize - xlen - 20; zs.next_out = all.data() + o; zs.avail_out = isize;
            inflate(&zs, Z_FINISH); inflateEnd(&zs);
        }
        p += bsize;
    }
    // header
    size_t q = 8 + inf_le32(all.data() + 4); const uint32_t n_ref = inf_le32(all.data() + q); q += 4;
    for (uint32_t r = 0; r < n_ref; r++) q += 8 + inf_le32(all.data() + q);
    // sequential reference walk
    uint64_t want_n = 0, want_digest = 0; std::vector<int32_t> want_tid, want_pos;
    {
        const int64_t u = walk_bam_block(all.data() + q, (uint32_t)(all.size() - q), [&](int32_t tid, int32_t pos, uint16_t flag, uint8_t mapq, uint32_t n_cigar, const uint8_t *cig) {
            if (!decode_keep(tid, flag, mapq, (int32_t)n_ref, xflags, min_mapq)) return;
            uint64_t h = mix(((uint64_t)(uint32_t)tid << 32) | (uint32_t)pos);
            for (uint32_t c = 0; c < n_cigar; c++) h = mix(h ^ inf_le32(cig + 4 * c));
            want_digest += h; want_n++; want_tid.push_back(tid); want_pos.push_back(pos);
        });
        if (u < 0 || (size_t)u != all.size() - q) { fprintf(stderr, "reference walk failed\n"); return 2; }
    }
    // first block holding a record
    size_t b0 = 0; while (b0 + 1 < blocks.size() && start[b0 + 1] <= q) b0++;
    uint64_t got_n = 0, got_digest = 0; size_t bad = 0, n_runs = 0, order_bad = 0;
    HostBackend be;
    for (size_t b = b0; b < blocks.size(); b += per_chunk) {
        std::vector<ChunkBlock> chunk(blocks.begin() + (long)b, blocks.begin() + (long)std::min(blocks.size(), b + per_chunk));
        const uint32_t skip = b == b0 ? (uint32_t)(q - start[b0]) : 0;
        std::vector<ContigRun> runs; uint32_t n_reads = 0, n_cigar = 0;
        const int rc = decode_chunk(be, data.data(), data.size(), chunk, skip, (int32_t)n_ref, xflags, min_mapq, runs, n_reads, n_cigar);
        if (rc != DEC_OK) { printf("decode_chunk returned %d at block %zu\n", rc, b); return rc == DEC_STRADDLE || rc == DEC_CG_TAG ? 3 : 1; }
        uint32_t cover = 0;
        for (const ContigRun &r : runs) {
            n_runs++;
            if (r.r0 != cover || r.r1 <= r.r0) bad++;
            cover = r.r1;
            bool sorted = true;
            for (uint32_t i = r.r0; i < r.r1; i++) { if (be.tid[i] != r.tid) bad++; if (i > r.r0 && be.pos[i] < be.pos[i - 1]) sorted = false; }
            if (r.sorted && !sorted) bad++;                  // claiming order that is not there is the dangerous direction
        }
        if (cover != n_reads) bad++;
        for (uint32_t i = 0; i < n_reads; i++) {
            uint64_t h = mix(((uint64_t)(uint32_t)be.tid[i] << 32) | (uint32_t)be.pos[i]);
            for (uint32_t c = be.cigar_off[i]; c < be.cigar_off[i + 1]; c++) h = mix(h ^ be.cigar[c]);
            got_digest += h;
            if (got_n + i < want_n && (be.tid[i] != want_tid[got_n + i] || be.pos[i] != want_pos[got_n + i])) order_bad++;
        }
        got_n += n_reads;
    }
    printf("records %llu (want %llu), digest %s, order mismatches %zu, run errors %zu, contig runs %zu\n", (unsigned long long)got_n, (unsigned long long)want_n,
           got_digest == want_digest ? "equal" : "DIFFERENT", order_bad, bad, n_runs);
    return (got_n == want_n && got_digest == want_digest && !order_bad && !bad) ? 0 : 1;
}
