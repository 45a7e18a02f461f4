// CPU check of the device BGZF decoder (pandepth_b200/csrc/pd_bgzf.cuh) through the warp emulator (warp_emu.h):
//   bgzf_emu_check file.bam [max_blocks] [xflags] [min_mapq] [warps] [first_block] [K]
//     K = 1: k_bgzf_decode (one decoder per warp); K = 2 / 4 / 8: k_bgzf_decode_multi<K> (K decoders per warp, symbol loop in lockstep)
// Every block: inflated bytes == zlib; dense pos / tid / cigar_off / cigar arrays, per-block summaries and totals == a plain
// sequential walk (walk_bam_block of pd_inflate.cuh, itself pinned to htslib by tests/test_inflate_cpu.py).
#define PD_WARP_EMU 1
#include "warp_emu.h"
#include "pd_bgzf.cuh"
#include "pd_inflate.cuh"
#include <zlib.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <tuple>
using namespace pdk;

static bool zlib_raw(const uint8_t *in, size_t n, uint8_t *out, size_t isize)
{
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = (Bytef *)in; zs.avail_in = (uInt)n; zs.next_out = out; zs.avail_out = (uInt)isize;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
    inflateEnd(&zs);
    return ok;
}

using OffArgs = std::tuple<const BgzfSummary *, uint32_t, uint32_t *, uint32_t *, uint32_t *>;
static void offsets_entry(OffArgs t) { k_bgzf_offsets(std::get<0>(t), std::get<1>(t), std::get<2>(t), std::get<3>(t), std::get<4>(t)); }

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: bgzf_emu_check file.bam [max_blocks] [xflags] [min_mapq] [warps] [first_block]\n"); return 2; }
    const size_t max_blocks = argc > 2 ? (size_t)atoll(argv[2]) : 8;
    const uint32_t xflags = argc > 3 ? (uint32_t)atoi(argv[3]) : 1796u;
    const int32_t min_mapq = argc > 4 ? atoi(argv[4]) : -1;
    const unsigned warps = argc > 5 ? (unsigned)atoi(argv[5]) : 2;
    const size_t first_block = argc > 6 ? (size_t)atoll(argv[6]) : 0;      // blocks to skip after the header (0 = start at the first record)
    const int K = argc > 7 ? atoi(argv[7]) : 1;
    FILE *f = fopen(argv[1], "rb"); if (!f) { perror(argv[1]); return 2; }
    std::vector<uint8_t> data; uint8_t tmp[1 << 16]; size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) data.insert(data.end(), tmp, tmp + k);
    fclose(f);
    const size_t file_len = data.size();
    data.resize(file_len + 64, 0);                           // padding the device buffers have too
    struct Blk { uint64_t in_off; uint32_t in_len, out_len; };
    std::vector<Blk> blocks;
    for (size_t p = 0; p + 18 <= file_len;) {
        const uint8_t *h = data.data() + p;
        if (h[0] != 31 || h[1] != 139) break;
        const uint32_t xlen = h[10] | h[11] << 8; uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) { const uint8_t *e = h + 12 + x; const uint32_t sl = e[2] | e[3] << 8; if (e[0] == 'B' && e[1] == 'C' && sl == 2) { bsize = (e[4] | e[5] << 8) + 1; break; } x += 4 + sl; }
        if (!bsize || p + bsize > file_len) break;
        const uint32_t isize = inf_le32(h + bsize - 4);
        if (isize) blocks.push_back(Blk{p + 12 + xlen, bsize - xlen - 20, isize});
        p += bsize;
    }
    // BAM header: first record offset
    std::vector<uint8_t> head; size_t hb = 0, q = 0; uint32_t n_ref = 0;
    for (;;) {
        if (hb >= blocks.size()) { fprintf(stderr, "no records\n"); return 2; }
        const size_t o = head.size(); head.resize(o + blocks[hb].out_len);
        if (!zlib_raw(data.data() + blocks[hb].in_off, blocks[hb].in_len, head.data() + o, blocks[hb].out_len)) { fprintf(stderr, "zlib rejects block %zu\n", hb); return 2; }
        hb++;
        if (head.size() < 12) continue;
        q = 8 + inf_le32(head.data() + 4); if (head.size() < q + 4) continue;
        n_ref = inf_le32(head.data() + q); q += 4; bool ok = true;
        for (uint32_t r = 0; r < n_ref; r++) { if (head.size() < q + 4) { ok = false; break; } q += 8 + inf_le32(head.data() + q); }
        if (ok && head.size() >= q) break;
    }
    size_t b0 = 0; uint64_t acc = 0; while (b0 < blocks.size() && acc + blocks[b0].out_len <= q) { acc += blocks[b0].out_len; b0++; }
    uint32_t skip0 = (uint32_t)(q - acc);
    if (first_block) { b0 += first_block; skip0 = 0; }
    if (b0 >= blocks.size()) { printf("%s: nothing to check\n", argv[1]); return 0; }
    const size_t nb = std::min(max_blocks, blocks.size() - b0);
    printf("%s: %zu blocks, %u contigs, first block %zu at +%u; checking %zu blocks\n", argv[1], blocks.size(), n_ref, b0, skip0, nb);

    // ---- device structures ----
    std::vector<BgzfBlockDesc> blk(nb); uint64_t oo = 0;
    for (size_t b = 0; b < nb; b++) { blk[b] = BgzfBlockDesc{blocks[b0 + b].in_off, blocks[b0 + b].in_len, blocks[b0 + b].out_len, oo}; oo += (blocks[b0 + b].out_len + 31u) & ~31u; }
    std::vector<uint8_t> out_store(oo + 64 + 32, 0xAB);
    uint8_t *const out = out_store.data() + ((32 - ((uintptr_t)out_store.data() & 31)) & 31);     // 32-byte aligned like the engine's buffer
    const size_t st_rec = oo / 32 + 2 * nb + 64, st_cig = oo / 4 + nb + 64, n_rec = oo / 36 + 64;
    std::vector<int32_t> st_pos(st_rec), st_tid(st_rec), pos(n_rec + 1), tid(n_rec + 1);
    std::vector<uint32_t> st_coff(st_rec), st_cigv(st_cig), coff(n_rec + 1), cig(oo / 4 + 64), r0(nb), c0(nb), totals(4);
    std::vector<BgzfSummary> sum(nb);
    uint32_t ticket = 0;
    BzArgs a{data.data(), out, blk.data(), (uint32_t)nb, skip0, (int32_t)n_ref, xflags, min_mapq,
             st_pos.data(), st_tid.data(), st_coff.data(), st_cigv.data(), sum.data(), &ticket, 1u /* verify the CRC32 behind every payload */};
    if (K == 2) emu::launch(1, 32 * std::min(warps, 4u), k_bgzf_decode_multi<2, 4>, a);
    else if (K == 4) emu::launch(1, 32 * std::min(warps, 2u), k_bgzf_decode_multi<4, 2>, a);
    else if (K == 8) emu::launch(1, 32, k_bgzf_decode_multi<8, 1>, a);
    else emu::launch(1, 32 * warps, k_bgzf_decode, a);
    emu::launch(1, 1024, offsets_entry, OffArgs((const BgzfSummary *)sum.data(), (uint32_t)nb, r0.data(), c0.data(), totals.data()));
    BzCompactArgs c{blk.data(), sum.data(), r0.data(), c0.data(), (uint32_t)nb, st_pos.data(), st_tid.data(), st_coff.data(), st_cigv.data(),
                    pos.data(), tid.data(), coff.data(), cig.data()};
    {   // no collectives in this kernel: lanes run one after another
        const unsigned grid = (unsigned)((nb + 7) / 8);
        for (unsigned b = 0; b < grid; b++) for (unsigned t = 0; t < 256; t++) { threadIdx.x = t; blockIdx.x = b; blockDim.x = 256; gridDim.x = grid; k_bgzf_compact(c); }
    }

    // ---- reference: zlib + sequential walk ----
    int bad = 0;
    std::vector<int32_t> w_pos, w_tid; std::vector<uint32_t> w_coff, w_cig;
    uint32_t flags_or = 0;
    for (size_t b = 0; b < nb; b++) {
        std::vector<uint8_t> want(blk[b].out_len);
        if (!zlib_raw(data.data() + blk[b].in_off, blk[b].in_len, want.data(), blk[b].out_len)) { printf("block %zu: zlib rejects\n", b); bad++; continue; }
        if (memcmp(want.data(), out + blk[b].out_off, blk[b].out_len) != 0) {
            size_t d = 0; while (want[d] == out[blk[b].out_off + d]) d++;
            printf("block %zu: inflated bytes differ at %zu of %u\n", b, d, blk[b].out_len); bad++;
        }
        for (size_t x = blk[b].out_len; x < ((blk[b].out_len + 31u) & ~31u); x++) if (out[blk[b].out_off + x] != 0xAB) { printf("block %zu: wrote past its output\n", b); bad++; break; }
        const uint32_t sk = b == 0 ? skip0 : 0;
        uint32_t n = 0, nc = 0, fl = 0; int32_t ft = INT32_MIN, fp = 0, lt = INT32_MIN, lp = INT32_MIN;
        const int64_t used = walk_bam_block(want.data() + sk, blk[b].out_len - sk, [&](int32_t t, int32_t p, uint16_t flag, uint8_t mapq, uint32_t ncg, const uint8_t *cg) {
            if (!(t >= 0 && t < (int32_t)n_ref && !(flag & xflags) && (int32_t)mapq >= min_mapq)) return;
            if (ncg == 2 && (cg[0] & 15u) == 4u && (cg[4] & 15u) == 3u) fl |= BZ_CG_TAG;
            if (ncg != 1) fl |= BZ_MULTI_OP;
            if (t < lt || (t == lt && p < lp)) fl |= BZ_UNSORTED;
            if (ft == INT32_MIN) { ft = t; fp = p; }
            lt = t; lp = p;
            w_pos.push_back(p); w_tid.push_back(t); w_coff.push_back((uint32_t)w_cig.size());
            for (uint32_t j = 0; j < ncg; j++) w_cig.push_back(inf_le32(cg + 4 * j));
            n++; nc += ncg;
        });
        if (used < 0) fl |= BZ_MALFORMED; else if ((uint32_t)used + sk != blk[b].out_len) fl |= BZ_STRADDLE;
        flags_or |= fl;
        const BgzfSummary &s = sum[b];
        if (used >= 0 && (s.n_reads != n || s.n_cigar != nc || s.flags != fl || s.first_tid != ft || s.last_tid != lt || (n && (s.first_pos != fp || s.last_pos != lp)) ||
                          s.used != (uint32_t)used + sk)) {
            printf("block %zu: summary differs: reads %u/%u cigar %u/%u flags %x/%x first %d:%d/%d:%d last %d:%d/%d:%d used %u/%lld\n", b, s.n_reads, n, s.n_cigar, nc, s.flags, fl,
                   s.first_tid, s.first_pos, ft, fp, s.last_tid, s.last_pos, lt, lp, s.used, (long long)used + sk);
            bad++;
        }
        if (used < 0 && !(s.flags & BZ_MALFORMED)) { printf("block %zu: malformed record not flagged\n", b); bad++; }
    }
    w_coff.push_back((uint32_t)w_cig.size());
    if (!(flags_or & BZ_MALFORMED)) {
        if (totals[0] != w_pos.size() || totals[1] != w_cig.size() || totals[2] != flags_or) { printf("totals differ: %u/%zu %u/%zu %x/%x\n", totals[0], w_pos.size(), totals[1], w_cig.size(), totals[2], flags_or); bad++; }
        else {
            if (!w_pos.empty() && memcmp(pos.data(), w_pos.data(), w_pos.size() * 4)) { printf("pos[] differs\n"); bad++; }
            if (!w_tid.empty() && memcmp(tid.data(), w_tid.data(), w_tid.size() * 4)) { printf("tid[] differs\n"); bad++; }
            if (memcmp(coff.data(), w_coff.data(), w_coff.size() * 4)) { printf("cigar_off[] differs\n"); bad++; }
            if (!w_cig.empty() && memcmp(cig.data(), w_cig.data(), w_cig.size() * 4)) { printf("cigar[] differs\n"); bad++; }
        }
    }
    // ---- k_delta_scatter_multi over the dense arrays (the path of blocks with a contig change or records out of order): contig
    // table from the BAM header; contig 0 is owned in part (range sharding), the last of three or more contigs not at all ----
    uint64_t multi_bases = 0;
    if (!bad && !(flags_or & BZ_MALFORMED) && n_ref) {
        std::vector<uint32_t> tlen(n_ref), org(n_ref, 0); std::vector<uint64_t> toff(n_ref); std::vector<uint8_t> owned(n_ref, 1);
        size_t hq = 8 + inf_le32(head.data() + 4) + 4; uint64_t total = 0;
        for (uint32_t r = 0; r < n_ref; r++) { const uint32_t ln = inf_le32(head.data() + hq); tlen[r] = inf_le32(head.data() + hq + 4 + ln); hq += 8 + ln; }
        if (tlen[0] >= 8) { org[0] = tlen[0] / 4; tlen[0] /= 2; }
        if (n_ref >= 3) owned[n_ref - 1] = 0;
        for (uint32_t r = 0; r < n_ref; r++) { toff[r] = total; total += ((uint64_t)tlen[r] + 1 + 8191) / 8192 * 8192; }
        if (total <= (64u << 20)) {
            std::vector<uint32_t> delta(total, 0), want(total, 0);
            const uint32_t n = totals[0];
            ScatterMultiArgs m{tid.data(), pos.data(), cig.data(), coff.data(), n, toff.data(), tlen.data(), org.data(), owned.data(), (int32_t)n_ref, delta.data()};
            const unsigned grid = (n + 255) / 256;
            for (unsigned b = 0; b < grid; b++) for (unsigned t = 0; t < 256; t++) { threadIdx.x = t; blockIdx.x = b; blockDim.x = 256; gridDim.x = grid; k_delta_scatter_multi(m); }
            for (size_t i = 0; i < w_pos.size(); i++) {                              // PanDepth.cpp:735-760 on the walked records
                const int32_t t = w_tid[i];
                if (!owned[t] || w_pos[i] < 0) continue;
                int64_t x = (int64_t)w_pos[i] - org[t];
                for (uint32_t k = w_coff[i]; k < w_coff[i + 1]; k++) {
                    const uint32_t op = w_cig[k] & 15u; const int64_t len = w_cig[k] >> 4;
                    if (op == 0 || op == 7 || op == 8) { for (int64_t y = std::max<int64_t>(x, 0); y < std::min<int64_t>(x + len, tlen[t]); y++) want[toff[t] + (uint64_t)y]++; x += len; }
                    else if (op == 2 || op == 3) x += len;
                }
            }
            uint32_t run = 0; size_t nd = 0; multi_bases = total;
            for (uint64_t i = 0; i < total; i++) { run += delta[i]; if (run != want[i]) nd++; }
            if (nd || run) { printf("k_delta_scatter_multi: %zu bases differ from the walk (residue %u)\n", nd, run); bad++; }
        }
    }
    printf("%s: %zu blocks, %zu records kept, %zu cigar words, flags %x, scatter_multi over %llu bases: %s\n", argv[1], nb, w_pos.size(), w_cig.size(), flags_or,
           (unsigned long long)multi_bases, bad ? "MISMATCH" : "ok");
    return bad ? 1 : 0;
}
