// pd_bgzf.cuh -- device-side BGZF decode: DEFLATE inflate + BAM record walk, one warp per BGZF block (sm_100a).
//
// Role in the reference: the decode side of the `sam_read1` / `sam_itr_next` loops (PanDepth.cpp:4608, :730, inside
// lib/libhts.a: bgzf_read_block + bam_read1) and the record filter :732-733.  A BGZF file is a sequence of independent
// <= 64 KB deflate streams and htslib's writer starts every block on a record boundary (bam_write1 flushes first) unless a
// single record is larger than a block, so blocks inflate AND parse independently: tens of thousands of decoders in flight.
//
//   k_bgzf_decode   warp per block (dynamic tickets).  Lane 0 runs the Huffman symbol loop (RFC 1951; same arithmetic as the
//                   CPU-verified pd_inflate.cuh); the 32 lanes build the decoding tables together, copy matches / stored
//                   blocks together (coalesced), and then walk the block's records together: lane 0 follows the
//                   block_size chain 32 records at a time, every lane parses one record, applies the filter and writes
//                   (tid, pos, cigar offset, CIGAR words) of the records that pass into the block's own staging region,
//                   plus a 32-byte summary per block.
//   k_bgzf_offsets  exclusive prefix sums of the per-block (reads, CIGAR words) counts -> final positions (one CTA).
//   k_bgzf_compact  staging regions -> dense pos[] / tid[] / cigar_off[] / cigar[] arrays (warp per block, coalesced),
//                   i.e. exactly what k_cut / k_tile_accumulate / k_delta_scatter take.
//
// Bounds: this kernel is latency / issue bound (a serial bit stream per block), not HBM bound.  Measured numbers live in
// DESIGN.md; algorithmic bytes per block = compressed bytes in + inflated bytes out.
#pragma once
#ifndef PD_WARP_EMU              // tests/cpp/warp_emu.h runs these kernels on the CPU (one host thread per lane)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace pdk {

struct BgzfBlockDesc { uint64_t in_off; uint32_t in_len; uint32_t out_len; uint64_t out_off; };   // payload inside the chunk; out_off = prefix of out_len

enum { BZ_MALFORMED = 1, BZ_CG_TAG = 2, BZ_UNSORTED = 4, BZ_MULTI_OP = 8, BZ_INFLATE = 16, BZ_STRADDLE = 32, BZ_CRC = 64 };

struct BgzfSummary {                 // per block, 32 bytes
    uint32_t n_reads;                // records passing the filter
    uint32_t n_cigar;                // their CIGAR words
    uint32_t used;                   // inflated bytes consumed by whole records (incl. the first block's skip)
    uint32_t flags;                  // BZ_*
    int32_t first_tid, last_tid;     // refID of the first / last passing record (INT32_MIN when none)
    int32_t first_pos, last_pos;
};

// staging geometry: a block with L inflated bytes holds at most L/36 records and fewer than L/4 CIGAR words
__host__ __device__ __forceinline__ uint64_t bz_rec_base(uint64_t out_off, uint32_t b) { return out_off / 32 + 2ull * b; }
__host__ __device__ __forceinline__ uint64_t bz_cig_base(uint64_t out_off, uint32_t b) { return out_off / 4 + b; }

constexpr int BZ_LIT_BITS = 10, BZ_DIST_BITS = 8;
constexpr int BZ_WPC = 8;            // warps (decoders) per CTA
// First-level table entry (16 bits): symbol in the LOW byte (a literal is stored to memory straight from the entry), code
// length in bits 8..11, bit 12 = "symbol is 256 + low byte" (end of block / length codes), 0xFFFF = code longer than the table.
// A literal hit is therefore e < 0x1000.
constexpr uint16_t BZ_MISS = 0xffffu;
__host__ __device__ __forceinline__ uint16_t bz_entry(uint32_t sym, uint32_t len) { return (uint16_t)((sym & 0xffu) | (len << 8) | (sym >= 256u ? 0x1000u : 0u)); }
__host__ __device__ __forceinline__ int bz_entry_sym(uint32_t e) { return (int)((e & 0xffu) + ((e & 0x1000u) >> 4)); }
__host__ __device__ __forceinline__ uint32_t bz_entry_len(uint32_t e) { return (e >> 8) & 15u; }

struct BzTables {                    // per warp, shared memory
    uint16_t dist_fast[1 << BZ_DIST_BITS];
    uint16_t lit_sym[288];                   // symbols in canonical code order
    uint16_t dist_sym[32];
    uint32_t lit_cnt[16], dist_cnt[16];      // codes per length
    uint32_t aux[48];                        // table build: running offsets / first code / first index per length
    uint8_t lens[320 + 24];                  // code-length code (19, padded to 24) then literal/length + distance lengths
};

struct BzArgs {
    const uint8_t *in;               // compressed bytes of the chunk (padded by >= 8 readable bytes)
    uint8_t *out;                    // inflated bytes (padded by >= 8 bytes)
    const BgzfBlockDesc *blk; uint32_t n_blocks; uint32_t first_skip;
    int32_t n_targets; uint32_t xflags; int32_t min_mapq;
    int32_t *st_pos; int32_t *st_tid; uint32_t *st_coff; uint32_t *st_cig;   // staging, indexed by bz_rec_base / bz_cig_base
    BgzfSummary *sum;
    uint32_t *ticket;                // dynamic block ids
    uint32_t check_crc;              // PD_CRC: the 4 bytes behind every payload are the block's CRC32 (BGZF trailer): verify it
};

// ---- bit reader over aligned 32-bit words (lane 0 only) -----------------------------------------------------------
// A 64-bit window (lo = word k, hi = word k+1) and a bit offset sh < 32 into it: peek32() is ONE funnel shift and always
// returns 32 valid bits, dropping bits is one add, and a refill (every >= 3 symbols) slides the window by a word.
// (The first version kept a 64-bit buffer: two 64-bit shifts + a counter update per symbol, 36 warp instructions per
// literal on SEQ+QUAL data -- the kernel is issue-bound, so the instruction count IS the throughput.)
struct BzBits {
    const uint32_t *w; uint32_t wi, wlimit; uint32_t lo, hi, nx; uint32_t sh; int mis;
    // beyond the stream every read returns the word just behind it (the chunk buffer is padded): garbage bits there only ever
    // feed a decode that the consumed_bits() checks reject, and the clamp keeps a corrupt stream from running off the buffer
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return w[i < wlimit ? i : wlimit]; }
    // nx = word wi is loaded one refill AHEAD of its use, so the load's latency is off the symbol chain
    __device__ __forceinline__ void init(const uint8_t *p, uint32_t n)
    {
        mis = (int)((uintptr_t)p & 3);
        w = reinterpret_cast<const uint32_t *>(p - mis);
        wlimit = (uint32_t)((mis + n + 3) >> 2);                 // words holding stream bytes
        lo = word(0); hi = word(1); nx = word(2); wi = 2; sh = 8u * (uint32_t)mis;
    }
    __device__ __forceinline__ void norm() { if (sh >= 32u) { lo = hi; hi = nx; wi++; nx = word(wi); sh -= 32u; } }   // sh < 64 on entry
    __device__ __forceinline__ uint32_t peek32() const { return __funnelshift_r(lo, hi, sh); }               // needs sh < 32
    __device__ __forceinline__ void drop(uint32_t n) { sh += n; }                                            // then norm() before the next peek
    __device__ __forceinline__ uint32_t bits(int n) { norm(); const uint32_t v = peek32() & ((1u << n) - 1u); sh += (uint32_t)n; return v; }   // n <= 16
    __device__ __forceinline__ int64_t consumed_bits() const { return ((int64_t)wi - 2) * 32 + (int64_t)sh - 8 * mis; }
    __device__ __forceinline__ void align_byte() { sh = (sh + 7u) & ~7u; }
    __device__ __forceinline__ void seek_bytes(uint32_t byte_pos)  // continue at stream byte `byte_pos`
    {
        const uint32_t a = (uint32_t)mis + byte_pos;
        const uint32_t idx = a >> 2;
        lo = word(idx); hi = word(idx + 1); nx = word(idx + 2); wi = idx + 2; sh = (a & 3u) * 8u;
    }
};

// canonical Huffman tables from code lengths, built by the whole warp.  Returns (to every lane) `left` as in
// inf_build(): < 0 over-subscribed, 0 complete, > 0 incomplete.
__device__ __forceinline__ int bz_build(const uint8_t *lens, int n, uint32_t *cnt, uint16_t *sym, uint16_t *fast, int fast_bits,
                                        uint32_t *aux, int lane)
{
    if (lane < 16) cnt[lane] = 0;
    for (int i = lane; i < (1 << fast_bits); i += 32) fast[i] = BZ_MISS;
    __syncwarp();
    for (int s = lane; s < n; s += 32) atomicAdd(&cnt[lens[s]], 1u);
    __syncwarp();
    if (cnt[0] == (uint32_t)n) return 0;                          // no codes at all
    int left = 1;
    for (int l = 1; l < 16; l++) { left <<= 1; left -= (int)cnt[l]; if (left < 0) return left; }
    // aux[l] = running index (symbols of length l placed so far), aux[16+l] = first code of length l, aux[32+l] = first index
    if (lane < 16) {
        uint32_t off = 0, code = 0;
        for (int l = 1; l < lane; l++) { off += cnt[l]; code = (code + cnt[l]) << 1; }
        if (lane >= 1) { aux[lane] = off; aux[16 + lane] = code; aux[32 + lane] = off; }
    }
    __syncwarp();
    // symbols in canonical order: by length, then by symbol value.  32 symbols at a time; equal lengths inside a group are
    // ranked with match_any.
    for (int g = 0; g < n; g += 32) {
        const int s = g + lane;
        const uint32_t l = s < n ? lens[s] : 0u;
        const unsigned same = __match_any_sync(0xffffffffu, l);
        if (l) {
            const uint32_t rank = __popc(same & ((1u << lane) - 1u));
            sym[aux[l] + rank] = (uint16_t)s;
        }
        __syncwarp();
        if (l && (same >> lane) == 1u) aux[l] += __popc(same);   // the highest lane of each length group advances its cursor
        __syncwarp();
    }
    // first-level table: symbol at canonical index i with length l <= fast_bits has code first[l] + (i - first_index[l]);
    // every fast_bits-bit pattern whose low l bits are the reversed code decodes to it
    const int coded = n - (int)cnt[0];
    for (int i = lane; i < coded; i += 32) {
        const uint32_t s = sym[i], l = lens[s];
        if (l <= (uint32_t)fast_bits) {
            const uint32_t code = aux[16 + l] + ((uint32_t)i - aux[32 + l]);
            const uint32_t rev = __brev(code) >> (32 - l);
            const uint16_t e = bz_entry(s, l);
            for (uint32_t p = rev; p < (1u << fast_bits); p += 1u << l) fast[p] = e;
        }
    }
    __syncwarp();
    return left;
}

// one symbol (lane 0): first-level hit, else canonical bit-by-bit decode (RFC 1951 3.2.2)
__device__ __forceinline__ int bz_symbol(BzBits &br, const uint32_t *cnt, const uint16_t *sym, const uint16_t *fast, int fast_bits)
{
    br.norm();
    const uint32_t b32 = br.peek32();
    const uint16_t e = fast[b32 & ((1u << fast_bits) - 1u)];
    if (e != BZ_MISS) { br.drop(bz_entry_len(e)); return bz_entry_sym(e); }
    int code = 0, first = 0, index = 0;
    uint32_t b = b32;
    for (int l = 1; l < 16; l++) {
        code |= (int)(b & 1u); b >>= 1;
        const int c = (int)cnt[l];
        if (code - c < first) { br.drop((uint32_t)l); return sym[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

__device__ __forceinline__ uint32_t bz_ld32(const uint8_t *p)     // little-endian u32 at any alignment (reads <= 7 bytes ahead)
{
    const uint32_t *a = reinterpret_cast<const uint32_t *>((uintptr_t)p & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)((uintptr_t)p & 3) * 8;
    const uint32_t w0 = a[0], w1 = a[1];
    return __funnelshift_r(w0, w1, sh);
}

// Inflate one raw deflate stream with the whole warp.  Returns (to every lane) bytes written, or < 0 on a malformed stream.
__device__ __forceinline__ uint32_t bz_shared_addr(const void *p)
{
#ifndef PD_WARP_EMU
    return (uint32_t)__cvta_generic_to_shared(p);
#else
    return emu_shared_addr(p);
#endif
}

__device__ __forceinline__ void *bz_shared_ptr(uint32_t sa)
{
#ifndef PD_WARP_EMU
    return __cvta_shared_to_generic(sa);
#else
    return const_cast<void *>(emu_shared_ptr(sa));
#endif
}

// First-level literal tables live in a raw shared array that is 2 KB larger than needed and are placed at the first 2 KB
// boundary OF THE SHARED WINDOW inside it, at run time: `__align__(2048)` on a static shared variable only aligns it inside
// the kernel's own section, and that section starts behind the 1 KB the system reserves at the bottom of the window
// (sm_90 and later) -- found the hard way: lookups formed with OR read the wrong entries on the device.
__device__ __forceinline__ uint32_t bz_align_tables(const void *raw) { return (bz_shared_addr(raw) + 2047u) & ~2047u; }

__device__ __forceinline__ uint32_t bz_lds_u16(uint32_t saddr)
{
#ifndef PD_WARP_EMU
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
    return v;
#else
    return *reinterpret_cast<const uint16_t *>(emu_shared_ptr(saddr));
#endif
}

// lit_fast: this warp's 1024-entry literal/length table, 2 KB aligned in shared memory (lit_sa = its shared-window address):
// a lookup is ONE logic op -- lit_sa | (bits << 1 & 0x7fe) -- and the load.
__device__ __forceinline__ int bz_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, BzTables &T, uint16_t *lit_fast, uint32_t lit_sa, int lane)
{
    constexpr unsigned FULLM = 0xffffffffu;
    BzBits br; br.init(in, in_len);
    uint32_t o = 0;
    const int64_t in_bits = (int64_t)in_len * 8;
    for (;;) {
        // ---- block header (lane 0), broadcast ----
        uint32_t hdr = 0;
        if (lane == 0) { hdr = br.bits(3); if (br.consumed_bits() > in_bits + 64) hdr |= 8u; }
        hdr = __shfl_sync(FULLM, hdr, 0);
        if (hdr & 8u) return -1;
        const uint32_t last = hdr & 1u, type = (hdr >> 1) & 3u;
        if (type == 0) {                                         // stored block: copied by all lanes
            uint32_t n = 0, p = 0; int bad = 0;
            if (lane == 0) {
                br.align_byte();
                n = br.bits(16); const uint32_t nn = br.bits(16);
                if ((n ^ nn) != 0xffffu) bad = 1;
                p = (uint32_t)(br.consumed_bits() >> 3);         // byte position of the payload in the stream
                if ((uint64_t)p + n > in_len || o + n > out_cap) bad = 1;
            }
            bad = __shfl_sync(FULLM, bad, 0);
            if (bad) return -1;
            n = __shfl_sync(FULLM, n, 0); p = __shfl_sync(FULLM, p, 0); o = __shfl_sync(FULLM, o, 0);
            for (uint32_t k = lane; k < n; k += 32) out[o + k] = in[p + k];
            o += n;
            if (lane == 0) br.seek_bytes(p + n);
            __syncwarp();
        } else if (type == 1 || type == 2) {
            int nlen = 288, ndist = 30;
            if (type == 1) {                                     // fixed code (RFC 1951 3.2.6)
                for (int i = lane; i < 288; i += 32) T.lens[24 + i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                if (lane < 30) T.lens[24 + 288 + lane] = 5;
                __syncwarp();
            } else {                                             // dynamic code (3.2.7)
                uint32_t h = 0;
                if (lane == 0) h = br.bits(14);
                h = __shfl_sync(FULLM, h, 0);
                nlen = (int)(h & 31u) + 257; ndist = (int)((h >> 5) & 31u) + 1;
                const int ncode = (int)(h >> 10) + 4;
                if (nlen > 286 || ndist > 30) return -1;
                if (lane < 24) T.lens[lane] = 0;
                __syncwarp();
                if (lane == 0) {
                    const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                    for (int i = 0; i < ncode; i++) T.lens[ORDER[i]] = (uint8_t)br.bits(3);
                }
                __syncwarp();
                // the code-length code is decoded through the distance tables' storage
                if (bz_build(T.lens, 19, T.dist_cnt, T.dist_sym, T.dist_fast, BZ_DIST_BITS, T.aux, lane) != 0) return -1;
                int bad = 0;
                if (lane == 0) {
                    int idx = 0;
                    while (idx < nlen + ndist) {
                        const int s = bz_symbol(br, T.dist_cnt, T.dist_sym, T.dist_fast, BZ_DIST_BITS);
                        if (s < 0) { bad = 1; break; }
                        if (s < 16) { T.lens[24 + idx++] = (uint8_t)s; continue; }
                        int rep, val = 0;
                        if (s == 16) { if (idx == 0) { bad = 1; break; } val = T.lens[24 + idx - 1]; rep = 3 + (int)br.bits(2); }
                        else if (s == 17) rep = 3 + (int)br.bits(3);
                        else rep = 11 + (int)br.bits(7);
                        if (idx + rep > nlen + ndist) { bad = 1; break; }
                        while (rep--) T.lens[24 + idx++] = (uint8_t)val;
                    }
                    if (!bad && T.lens[24 + 256] == 0) bad = 1;   // no end-of-block code
                    if (br.consumed_bits() > in_bits + 64) bad = 1;
                }
                bad = __shfl_sync(FULLM, bad, 0);
                if (bad) return -1;
                __syncwarp();
            }
            int r = bz_build(T.lens + 24, nlen, T.lit_cnt, T.lit_sym, lit_fast, BZ_LIT_BITS, T.aux, lane);
            if (r < 0 || (r > 0 && (type == 1 ? false : nlen - (int)T.lit_cnt[0] != 1))) return -1;
            r = bz_build(T.lens + 24 + nlen, ndist, T.dist_cnt, T.dist_sym, T.dist_fast, BZ_DIST_BITS, T.aux, lane);
            if (r < 0 || (r > 0 && (type == 1 ? false : ndist - (int)T.dist_cnt[0] != 1))) return -1;
            // ---- symbols: lane 0 decodes and stores literals; a match is copied by all lanes ----
            for (;;) {
                int cmd = 0;                                     // 1 = match, 2 = end of block, 3 = error
                uint32_t mlen = 0, dist = 0;
                if (lane == 0) {
                    for (;;) {
                        // 32 window bits hold three first-level literal hits (<= 10 bits each): one funnel shift, then 32-bit
                        // shifts; the output bound is checked once per group (a group writes at most 3 bytes)
                        br.norm();
                        uint32_t wnd1 = br.peek32() << 1;             // kept shifted left by one: (wnd1 & 0x7fe) is the byte offset of the entry
                        uint32_t e = bz_lds_u16(lit_sa | (wnd1 & 0x7feu));
                        if (e < 0x1000u && o + 3u <= out_cap) {
                            uint8_t *op = out + o;                       // one address per group, immediate offsets for the stores
                            uint32_t used = e >> 8;
                            op[0] = (uint8_t)e; wnd1 >>= used;
                            e = bz_lds_u16(lit_sa | (wnd1 & 0x7feu));
                            if (e < 0x1000u) {
                                op[1] = (uint8_t)e; wnd1 >>= (e >> 8); used += e >> 8;
                                e = bz_lds_u16(lit_sa | (wnd1 & 0x7feu));
                                if (e < 0x1000u) { op[2] = (uint8_t)e; used += e >> 8; o += 3; br.drop(used); continue; }
                                o += 2;
                            } else o += 1;
                            br.drop(used);
                            // e is the entry of the symbol BEHIND the literals (at most 20 of the window's 32 bits are used, so its
                            // 10 lookup bits were valid): a length code or the end of the block -- taken from here instead of being
                            // looked up twice more (match-dominated streams alternate literal / match: a fifth of their instructions)
                        }
                        int s;
                        if (e == BZ_MISS) s = bz_symbol(br, T.lit_cnt, T.lit_sym, lit_fast, BZ_LIT_BITS);     // code longer than the table
                        else { br.drop(bz_entry_len(e)); s = bz_entry_sym(e); }
                        if (s < 0) { cmd = 3; break; }
                        if (s < 256) { if (o >= out_cap) { cmd = 3; break; } out[o++] = (uint8_t)s; continue; }
                        if (s == 256) { cmd = 2; break; }
                        s -= 257;
                        if (s >= 29) { cmd = 3; break; }
                        // length / distance codes (3.2.5) are arithmetic: code s has e = (s-4)>>2 extra bits and base
                        // 3 + ((4 + (s&3)) << e) (258 for the last one); distance code d: e = (d-2)>>1, base 1 + ((2 + (d&1)) << e)
                        if (s < 8) mlen = 3u + (uint32_t)s;
                        else if (s == 28) mlen = 258u;
                        else { const int x = (s - 4) >> 2; mlen = 3u + ((4u + ((uint32_t)s & 3u)) << x) + br.bits(x); }
                        const int d = bz_symbol(br, T.dist_cnt, T.dist_sym, T.dist_fast, BZ_DIST_BITS);
                        if (d < 0 || d >= 30) { cmd = 3; break; }
                        if (d < 4) dist = 1u + (uint32_t)d;
                        else { const int x = (d - 2) >> 1; dist = 1u + ((2u + ((uint32_t)d & 1u)) << x) + br.bits(x); }
                        if (dist > o || o + mlen > out_cap) { cmd = 3; break; }
                        cmd = 1;
                        break;
                    }
                }
                cmd = __shfl_sync(FULLM, cmd, 0);                // (also orders lane 0's literal stores before the copy below)
                if (cmd == 3) return -1;
                if (cmd == 2) break;
                mlen = __shfl_sync(FULLM, mlen, 0); dist = __shfl_sync(FULLM, dist, 0); o = __shfl_sync(FULLM, o, 0);
                __syncwarp();
                // byte o+k of a match equals byte o-dist+(k mod dist): every source lies before o, so all lanes copy at once
                const uint8_t *src = out + o - dist;
                if (dist >= mlen) { for (uint32_t k = lane; k < mlen; k += 32) out[o + k] = src[k]; }
                else if (mlen <= 2u * dist) { for (uint32_t k = lane; k < mlen; k += 32) out[o + k] = src[k >= dist ? k - dist : k]; }   // one wrap at most
                else { for (uint32_t k = lane; k < mlen; k += 32) out[o + k] = src[k % dist]; }
                o += mlen;
                __syncwarp();
            }
        } else return -1;
        int over = 0;
        if (lane == 0) over = br.consumed_bits() > in_bits + 7 ? 1 : 0;      // consumed more bytes than the stream has
        over = __shfl_sync(FULLM, over, 0);
        o = __shfl_sync(FULLM, o, 0);
        if (over) return -1;
        if (last) break;
    }
    __syncwarp();
    return (int)o;
}

// ---------------------------------------------------------------------------------------------------------------
// CRC32 of an inflated block (zlib's: reflected 0xEDB88320), by the whole warp -- what bgzf_read_block verifies on the host.
// The block is cut into 64 stripes of S bytes (the last ones may be short or empty); lane i computes the CRC32 of its two stripes
// byte by byte through a 1 KB table in shared memory, and the stripes' CRCs are combined with the linear-algebra identity
// behind zlib's crc32_combine:
//     crc(A || B) = x^(8 |B|) * crc(A)  (mod P)  xor  crc(B)
// so  crc(block) = XOR_i  x^(8 * bytes behind stripe i) * crc(stripe i); the powers of x are built by square-and-multiply.
// About 20 K warp instructions per 64 KB block.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t BZ_CRC_POLY = 0xedb88320u;

__host__ __device__ __forceinline__ uint32_t bz_crc_mulmod(uint32_t a, uint32_t b)   // a * b mod P, reflected representation (bit 31 = x^0)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1u)) == 0) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ BZ_CRC_POLY : b >> 1;
        if (m == 0) break;
    }
    return p;
}

__host__ __device__ __forceinline__ uint32_t bz_crc_xpow8n(uint32_t n)               // x^(8 n) mod P
{
    uint32_t p = 1u << 31;                                     // x^0
    uint32_t sq = 0x00800000u;                                 // x^8 (x^1 = 0x40000000, ..., x^8 = bit 23)
    while (n) {
        if (n & 1u) p = bz_crc_mulmod(sq, p);
        n >>= 1;
        if (n) sq = bz_crc_mulmod(sq, sq);
    }
    return p;
}

// tab[0..255]: the byte table; tab[256 + k] = x^(2^k) mod P, k < 32 (the powers every stripe weight is a product of: with them
// x^(8 n) costs popcount(n) multiplications instead of a square-and-multiply chain per block -- those chains were half of
// the check's instructions).  By the whole CTA, then a barrier.
constexpr int BZ_CRC_TAB_WORDS = 256 + 32;
__device__ __forceinline__ void bz_crc_table_init(uint32_t *tab, int tid, int nthreads)
{
    for (int i = tid; i < 256; i += nthreads) {
        uint32_t c = (uint32_t)i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ BZ_CRC_POLY : c >> 1;
        tab[i] = c;
    }
    if (tid == 0) {
        uint32_t p = 0x40000000u;                              // x^1
        for (int k = 0; k < 32; k++) { tab[256 + k] = p; p = bz_crc_mulmod(p, p); }
    }
}

__device__ __forceinline__ uint32_t bz_crc_xpow8n_tab(uint32_t n, const uint32_t *tab)     // x^(8 n) mod P, n < 2^28
{
    uint32_t p = 1u << 31;
    for (int k = 3; n; n >>= 1, k++) if (n & 1u) p = bz_crc_mulmod(tab[256 + k], p);
    return p;
}

__device__ __forceinline__ uint32_t bz_crc32_warp(const uint8_t *out, uint32_t len, const uint32_t *tab, int lane)
{
    // 64 stripes of S bytes, S a multiple of 32, left-aligned: stripe j = [j S, min((j + 1) S, len)); lane i owns stripes 2i and
    // 2i + 1 and walks them TOGETHER -- two independent table chains per lane: a single byte-serial chain (lookup -> xor ->
    // lookup) left the warp waiting on shared-memory latency for 2 K steps per block.  `out` is 32-byte aligned (the engine
    // places every block's output on a 32-byte boundary), so a lane reads whole 32-byte sectors, each exactly once -- word
    // loads walked every sector eight times through an L1 that 48 resident warps leave no room in (L2-bandwidth bound).
    const uint32_t S = (((len + 63u) >> 6) + 31u) & ~31u;
    const uint32_t a0 = 2u * (uint32_t)lane * S, a1 = min(a0 + S, len), b0 = a0 + S, b1 = min(b0 + S, len);
    uint32_t ca = 0xffffffffu, cb = 0xffffffffu;
    auto step2 = [&](uint32_t wa, uint32_t wb) {               // one word of either stripe
#pragma unroll
        for (int k = 0; k < 4; k++) {
            ca = tab[(ca ^ wa) & 0xffu] ^ (ca >> 8); wa >>= 8;
            cb = tab[(cb ^ wb) & 0xffu] ^ (cb >> 8); wb >>= 8;
        }
    };
    uint32_t i = a0, j = b0;
    if ((((uintptr_t)out) & 31u) == 0) {
        for (; j + 32 <= b1; i += 32, j += 32) {               // stripe 2i + 1 is never longer than stripe 2i
            const uint4 u = *reinterpret_cast<const uint4 *>(out + i), x = *reinterpret_cast<const uint4 *>(out + j);
            step2(u.x, x.x); step2(u.y, x.y); step2(u.z, x.z); step2(u.w, x.w);
            const uint4 v = *reinterpret_cast<const uint4 *>(out + i + 16), y = *reinterpret_cast<const uint4 *>(out + j + 16);
            step2(v.x, y.x); step2(v.y, y.y); step2(v.z, y.z); step2(v.w, y.w);
        }
    }
    for (; i < a1; i++) ca = tab[(ca ^ out[i]) & 0xffu] ^ (ca >> 8);        // what is left of either stripe (everything, for an unaligned buffer)
    for (; j < b1; j++) cb = tab[(cb ^ out[j]) & 0xffu] ^ (cb >> 8);
    ca = a0 < len ? ~ca : 0u;                                  // crc32 of an empty stripe is 0 (and contributes nothing)
    cb = b0 < len ? ~cb : 0u;
    // weight of a stripe: x^(8 * bytes behind it) = x^(8 rem) * (x^(8 S))^(nfull - 1 - j) for the full stripes, 1 for the last one
    const uint32_t nfull = len / S, rem = len - nfull * S;
    auto weight = [&](uint32_t stripe) {                       // bytes behind a full stripe: rem + (nfull - 1 - stripe) S
        return stripe >= nfull ? 1u << 31 : bz_crc_xpow8n_tab(rem + (nfull - 1u - stripe) * S, tab);
    };
    const uint32_t wb_ = weight(2u * (uint32_t)lane + 1u);
    const uint32_t wa_ = (2u * (uint32_t)lane + 1u < nfull) ? bz_crc_mulmod(bz_crc_xpow8n_tab(S, tab), wb_) : weight(2u * (uint32_t)lane);   // one stripe further from the end
    uint32_t t = bz_crc_mulmod(wa_, ca) ^ bz_crc_mulmod(wb_, cb);
#pragma unroll
    for (int o = 16; o; o >>= 1) t ^= __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// BAM records of one inflated block (SAMv1 4.2: block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, ...), walked by
// the whole warp: lane 0 follows the block_size chain 32 records at a time, every lane parses one record, applies the filter
// (PanDepth.cpp:732-733) and writes the records that pass into the block's staging region; one summary per block.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bz_walk_block(const BzArgs &a, uint32_t b, const BgzfBlockDesc &k, uint8_t *out, int got, uint32_t *rec_off, int lane)
{
    constexpr unsigned FULLM = 0xffffffffu;
    uint32_t flags = 0;
    if (got == -2) flags |= BZ_CRC;                            // inflated to ISIZE bytes, but not to the bytes the trailer's CRC32 describes
    else if (got != (int)k.out_len) flags |= BZ_INFLATE;
    // ---- BAM records of the block (SAMv1 4.2): block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, ... ----
    const uint64_t rb = bz_rec_base(k.out_off, b), cb = bz_cig_base(k.out_off, b);
    const uint32_t len = k.out_len;
    uint32_t p = b == 0 ? a.first_skip : 0u;                 // lane 0's chain position
    uint32_t n_reads = 0, n_cig = 0;
    int32_t first_tid = INT32_MIN, first_pos = 0, last_tid = INT32_MIN, last_pos = INT32_MIN;
    bool chain_end = (flags != 0) || p > len;
    if (p > len) flags |= BZ_MALFORMED;
    while (!chain_end) {
        uint32_t nrec = 0;
        p = __shfl_sync(FULLM, p, 0);
        // Speculation: records of one size follow each other in most blocks (always in files of fixed-length reads).  Lane i
        // guesses that record i starts i * (4 + block_size of the first record) behind it and reads the size there; the leading
        // lanes that all found that same size have their records' true offsets -- two dependent loads for up to 32 records where
        // following the chain costs one L2 round trip per record (17 % of this kernel's stall samples on SEQ-less files).
        bool spec = false;
        if (p + 4 <= len) {
            const uint32_t bs0 = bz_ld32(out + p);
            if (bs0 >= 32 && bs0 <= (1u << 30)) {
                const uint64_t step = 4ull + bs0, pi = (uint64_t)p + (uint64_t)lane * step;
                const bool fits = pi + step <= (uint64_t)len;
                const uint32_t bsi = fits ? bz_ld32(out + pi) : 0u;
                const unsigned same = __ballot_sync(FULLM, fits && bsi == bs0);
                const uint32_t nsame = same == FULLM ? 32u : (uint32_t)(__ffs((int)~same) - 1);
                if (nsame >= 2u) {
                    if ((uint32_t)lane < nsame) rec_off[lane] = (uint32_t)pi;
                    nrec = nsame;
                    p += (uint32_t)(nsame * step);
                    if (p + 4 > len) nrec |= 0x400u;                                  // end of the block reached
                    spec = true;
                }
            }
        }
        // otherwise lane 0 follows the block_size chain for up to 32 records
        if (!spec && lane == 0) {
            while (nrec < 32 && p + 4 <= len) {
                const uint32_t bs = bz_ld32(out + p);
                if (bs < 32 || bs > (1u << 30)) { nrec |= 0x100u; break; }       // malformed
                if ((uint64_t)p + 4 + bs > len) { nrec |= 0x200u; break; }       // continues in the next block
                rec_off[nrec++] = p;
                p += 4 + bs;
            }
            if (!(nrec & 0x300u) && p + 4 > len) nrec |= 0x400u;                 // end of the block reached
        }
        if (!spec) nrec = __shfl_sync(FULLM, nrec, 0);                            // (warp-uniform branch)
        if (nrec & 0x100u) flags |= BZ_MALFORMED;
        if (nrec & 0x700u) chain_end = true;
        nrec &= 0xffu;
        __syncwarp();
        // every lane parses one record
        bool keep = false; int32_t tid = 0, pos = 0; uint32_t ncg = 0; const uint8_t *cig = nullptr;
        if ((uint32_t)lane < nrec) {
            const uint8_t *rec = out + rec_off[lane] + 4;
            const uint32_t bs = bz_ld32(rec - 4);
            tid = (int32_t)bz_ld32(rec); pos = (int32_t)bz_ld32(rec + 4);
            const uint32_t w8 = bz_ld32(rec + 8), w12 = bz_ld32(rec + 12);
            const uint32_t l_name = w8 & 0xffu, mapq = (w8 >> 8) & 0xffu;
            ncg = w12 & 0xffffu;
            const uint32_t flag = w12 >> 16;
            if (32 + l_name + 4 * ncg > bs) flags |= BZ_MALFORMED;
            else {
                cig = rec + 32 + l_name;
                keep = tid >= 0 && tid < a.n_targets && !(flag & a.xflags) && (int32_t)mapq >= a.min_mapq;   // :732-733
                if (keep && ncg == 2 && (bz_ld32(cig) & 15u) == 4u && (bz_ld32(cig + 4) & 15u) == 3u) flags |= BZ_CG_TAG;   // `<l_seq>S<ref_len>N`: CIGAR may be in the CG tag
                if (keep && ncg != 1) flags |= BZ_MULTI_OP;
            }
        }
        const unsigned km = __ballot_sync(FULLM, keep);
        if (km) {
            // order among the records that pass: each compares with the previous passing one
            const unsigned before = km & ((1u << lane) - 1u);
            const int src = before ? 31 - __clz(before) : lane;
            const int32_t ptid = __shfl_sync(FULLM, tid, src), ppos = __shfl_sync(FULLM, pos, src);
            if (keep) {
                const int32_t qt = before ? ptid : last_tid, qp = before ? ppos : last_pos;
                if (tid < qt || (tid == qt && pos < qp)) flags |= BZ_UNSORTED;
            }
            // staging: record rank and CIGAR word offset by warp prefix sums
            const uint32_t rank = n_reads + __popc(before);
            uint32_t c = keep ? ncg : 0u, incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULLM, incl, o); if (lane >= o) incl += y; }
            const uint32_t coff = n_cig + incl - c;
            if (keep) {
                a.st_pos[rb + rank] = pos; a.st_tid[rb + rank] = tid; a.st_coff[rb + rank] = coff;
                for (uint32_t j = 0; j < ncg; j++) a.st_cig[cb + coff + j] = bz_ld32(cig + 4 * j);
            }
            const int lo_lane = __ffs(km) - 1, hi_lane = 31 - __clz(km);
            if (first_tid == INT32_MIN) { first_tid = __shfl_sync(FULLM, tid, lo_lane); first_pos = __shfl_sync(FULLM, pos, lo_lane); }
            last_tid = __shfl_sync(FULLM, tid, hi_lane); last_pos = __shfl_sync(FULLM, pos, hi_lane);
            n_reads += __popc(km);
            n_cig += __shfl_sync(FULLM, incl, 31);
        }
        __syncwarp();
    }
    p = __shfl_sync(FULLM, p, 0);
    // flags raised by single lanes
    flags = __reduce_or_sync(FULLM, flags);
    if (!(flags & (BZ_INFLATE | BZ_MALFORMED | BZ_CRC)) && p != len) flags |= BZ_STRADDLE;
    if (lane == 0) {
        BgzfSummary s;
        s.n_reads = n_reads; s.n_cigar = n_cig; s.used = p; s.flags = flags;
        s.first_tid = first_tid; s.last_tid = last_tid; s.first_pos = first_pos; s.last_pos = last_tid == INT32_MIN ? 0 : last_pos;
        a.sum[b] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_bgzf_decode: one decoder per warp (match-dominated streams: SEQ-less BAMs, ratio >= 4)
// ---------------------------------------------------------------------------------------------------------------
#ifndef PD_BZ_K1_CTAS
#define PD_BZ_K1_CTAS 6               // resident CTAs per SM k_bgzf_decode is compiled for (register cap: 4 -> 64, 5 -> 48, 6 -> 40 registers and
                                      // 36 bytes of spills).  Measured, 20 M flavour-B reads at K=1: 117.6 / 101.4 / 98.5 ms for 4 / 5 / 6
#endif
__global__ void __launch_bounds__(BZ_WPC * 32, PD_BZ_K1_CTAS) k_bgzf_decode(BzArgs a)
{
    constexpr unsigned FULLM = 0xffffffffu;
    __shared__ uint16_t lit_raw[(BZ_WPC + 1) << BZ_LIT_BITS];
    __shared__ BzTables tabs[BZ_WPC];
    __shared__ uint32_t rec_offs[BZ_WPC][32];
    __shared__ uint32_t crc_tab[BZ_CRC_TAB_WORDS];
    bz_crc_table_init(crc_tab, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BzTables &T = tabs[warp];
    const uint32_t lit_sa = bz_align_tables(lit_raw) + (uint32_t)warp * (2u << BZ_LIT_BITS);
    uint16_t *lit_fast = static_cast<uint16_t *>(bz_shared_ptr(lit_sa));
    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(a.ticket, 1u);
        b = __shfl_sync(FULLM, b, 0);
        if (b >= a.n_blocks) return;
        const BgzfBlockDesc k = a.blk[b];
        uint8_t *out = a.out + k.out_off;
        int got = k.out_len ? bz_inflate(a.in + k.in_off, k.in_len, out, k.out_len, T, lit_fast, lit_sa, lane) : 0;
        __syncwarp();
        if (a.check_crc && k.out_len && got == (int)k.out_len) {
            const uint8_t *tr = a.in + k.in_off + k.in_len;                       // BGZF trailer: CRC32, ISIZE
            const uint32_t want = (uint32_t)tr[0] | (uint32_t)tr[1] << 8 | (uint32_t)tr[2] << 16 | (uint32_t)tr[3] << 24;
            if (bz_crc32_warp(out, k.out_len, crc_tab, lane) != want) got = -2;
        }
        bz_walk_block(a, b, k, out, got, rec_offs[warp], lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K decoders per warp (literal-dominated streams: BAMs with SEQ + QUAL).
//
// With one decoder per warp the symbol loop runs on ONE lane: every instruction spends a whole issue slot on one literal, and
// the kernel is issue-bound (about 13 warp instructions per literal).  Here a warp inflates K blocks at once.  The K lane-group
// leaders run the symbol loop TOGETHER: the loop itself is warp-uniform (every lane iterates, a warp vote closes each
// iteration, so the leaders re-converge after every step -- a loop that only the leaders execute does not: after their first
// divergence the hardware schedules them one after the other until the loop's exit, which is what the first version of this
// kernel did and why it was slower than one decoder per warp), and one step decodes up to three literals per leader.  A leader
// also serves most of its own matches (short, source before the destination: random SEQ coincidences and the record header
// fields -- 86 % of the matches of such a stream) and the rare codes longer than the first-level table.  What needs the whole
// warp -- long or overlapping matches, block headers and table builds -- stops the loop at the first such event and is served
// one decoder at a time with all 32 lanes.  Per-decoder progress lives in its leader's registers.
// ---------------------------------------------------------------------------------------------------------------
enum { BZS_HEADER = 0, BZS_SYMBOLS = 1, BZS_DONE = 2, BZS_ERROR = 3 };

__device__ __forceinline__ bool bz_any(bool p)
{
    return __ballot_sync(0xffffffffu, p) != 0u;
}

template <typename P>
__device__ __forceinline__ P *bz_shfl_ptr(P *p, int src)
{
    return reinterpret_cast<P *>((uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)p, src));
}

constexpr uint32_t BZ_SELF_MAX = 64;     // longest match a leader copies itself (source entirely before the destination)

// returns, to every lane, the result of the decoder its lane group describes: bytes written (== the block's ISIZE) or -1
template <int K, int LB>
__device__ __forceinline__ int bz_inflate_multi(const BzArgs &a, uint32_t b0, BzTables *T, uint32_t lit_sa0, int lane)
{
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int STRIDE = 32 / K;
    constexpr uint32_t LMASK = ((1u << LB) - 1u) << 1;          // byte offset of a first-level entry: (window << 1) & LMASK
    const int my = lane / STRIDE;                                // the decoder this lane's registers describe
    const bool leader = (lane % STRIDE) == 0;
    const bool valid = b0 + (uint32_t)my < a.n_blocks;
    const BgzfBlockDesc d = a.blk[valid ? b0 + (uint32_t)my : b0];
    const uint8_t *in_my = a.in + d.in_off; uint8_t *out_my = a.out + d.out_off;
    const uint32_t in_len = valid ? d.in_len : 0u, cap = valid ? d.out_len : 0u;
    BzBits br; br.init(in_my, in_len);
    uint32_t o = 0, last = 0; int state = (valid && cap) ? BZS_HEADER : BZS_DONE;
    const int64_t in_bits = (int64_t)in_len * 8;
    BzTables &Tm = T[my];
    const uint32_t lit_sa = lit_sa0 + (uint32_t)my * (2u << LB);          // decoder `my`'s table: 2 KB aligned in the shared window
    const uint16_t *lit_m = static_cast<const uint16_t *>(bz_shared_ptr(lit_sa));
    for (;;) {
        // ---- block headers and tables, one decoder at a time, all lanes ----
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int L = j * STRIDE;
            if (__shfl_sync(FULLM, state, L) != BZS_HEADER) continue;
            BzTables &Tj = T[j];
            uint16_t *lit_j = static_cast<uint16_t *>(bz_shared_ptr(lit_sa0 + (uint32_t)j * (2u << LB)));
            uint32_t hdr = 0;
            if (lane == L) { hdr = br.bits(3); if (br.consumed_bits() > in_bits + 64) hdr |= 8u; }
            hdr = __shfl_sync(FULLM, hdr, L);
            if (hdr & 8u) { if (lane == L) state = BZS_ERROR; continue; }
            const uint32_t lastj = hdr & 1u, type = (hdr >> 1) & 3u;
            if (lane == L) last = lastj;
            if (type == 0) {                                     // stored block: copied by all lanes
                uint32_t n = 0, p = 0, oo = 0; int bad = 0;
                if (lane == L) {
                    br.align_byte();
                    n = br.bits(16); const uint32_t nn = br.bits(16);
                    if ((n ^ nn) != 0xffffu) bad = 1;
                    p = (uint32_t)(br.consumed_bits() >> 3);
                    if ((uint64_t)p + n > in_len || o + n > cap) bad = 1;
                    oo = o;
                }
                bad = __shfl_sync(FULLM, bad, L);
                if (bad) { if (lane == L) state = BZS_ERROR; continue; }
                n = __shfl_sync(FULLM, n, L); p = __shfl_sync(FULLM, p, L); oo = __shfl_sync(FULLM, oo, L);
                const uint8_t *in_j = bz_shfl_ptr(in_my, L); uint8_t *out_j = bz_shfl_ptr(out_my, L);
                for (uint32_t q = lane; q < n; q += 32) out_j[oo + q] = in_j[p + q];
                if (lane == L) { o += n; br.seek_bytes(p + n); state = lastj ? BZS_DONE : BZS_HEADER; }
                __syncwarp();
                continue;
            }
            if (type == 3) { if (lane == L) state = BZS_ERROR; continue; }
            int nlen = 288, ndist = 30;
            if (type == 1) {                                     // fixed code (RFC 1951 3.2.6)
                for (int i = lane; i < 288; i += 32) Tj.lens[24 + i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                if (lane < 30) Tj.lens[24 + 288 + lane] = 5;
                __syncwarp();
            } else {                                             // dynamic code (3.2.7)
                uint32_t h = 0;
                if (lane == L) h = br.bits(14);
                h = __shfl_sync(FULLM, h, L);
                nlen = (int)(h & 31u) + 257; ndist = (int)((h >> 5) & 31u) + 1;
                const int ncode = (int)(h >> 10) + 4;
                if (nlen > 286 || ndist > 30) { if (lane == L) state = BZS_ERROR; continue; }
                if (lane < 24) Tj.lens[lane] = 0;
                __syncwarp();
                if (lane == L) {
                    const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                    for (int i = 0; i < ncode; i++) Tj.lens[ORDER[i]] = (uint8_t)br.bits(3);
                }
                __syncwarp();
                if (bz_build(Tj.lens, 19, Tj.dist_cnt, Tj.dist_sym, Tj.dist_fast, BZ_DIST_BITS, Tj.aux, lane) != 0) { if (lane == L) state = BZS_ERROR; continue; }
                int bad = 0;
                if (lane == L) {
                    int idx = 0;
                    while (idx < nlen + ndist) {
                        const int sy = bz_symbol(br, Tj.dist_cnt, Tj.dist_sym, Tj.dist_fast, BZ_DIST_BITS);
                        if (sy < 0) { bad = 1; break; }
                        if (sy < 16) { Tj.lens[24 + idx++] = (uint8_t)sy; continue; }
                        int rep, val = 0;
                        if (sy == 16) { if (idx == 0) { bad = 1; break; } val = Tj.lens[24 + idx - 1]; rep = 3 + (int)br.bits(2); }
                        else if (sy == 17) rep = 3 + (int)br.bits(3);
                        else rep = 11 + (int)br.bits(7);
                        if (idx + rep > nlen + ndist) { bad = 1; break; }
                        while (rep--) Tj.lens[24 + idx++] = (uint8_t)val;
                    }
                    if (!bad && Tj.lens[24 + 256] == 0) bad = 1;
                    if (br.consumed_bits() > in_bits + 64) bad = 1;
                }
                bad = __shfl_sync(FULLM, bad, L);
                if (bad) { if (lane == L) state = BZS_ERROR; continue; }
                __syncwarp();
            }
            int r = bz_build(Tj.lens + 24, nlen, Tj.lit_cnt, Tj.lit_sym, lit_j, LB, Tj.aux, lane);
            bool okb = !(r < 0 || (r > 0 && type == 2 && nlen - (int)Tj.lit_cnt[0] != 1));
            if (okb) {
                r = bz_build(Tj.lens + 24 + nlen, ndist, Tj.dist_cnt, Tj.dist_sym, Tj.dist_fast, BZ_DIST_BITS, Tj.aux, lane);
                okb = !(r < 0 || (r > 0 && type == 2 && ndist - (int)Tj.dist_cnt[0] != 1));
            }
            if (lane == L) state = okb ? BZS_SYMBOLS : BZS_ERROR;
        }
        // ---- symbols: a warp-uniform loop; the leaders inside a block advance together until one of them needs the warp ----
        int cmd = 0;                                             // 1 = match for the warp, 2 = end of block, 3 = error
        uint32_t mlen = 0, dist = 0;
        const bool act = leader && state == BZS_SYMBOLS;
        if (bz_any(act)) {
            do {
                if (act) {
                    br.norm();
                    uint32_t wnd1 = br.peek32() << 1;
                    uint32_t e = bz_lds_u16(lit_sa | (wnd1 & LMASK));
                    if (e < 0x1000u && o + 3u <= cap) {              // up to three first-level literal hits out of one 32-bit window
                        uint8_t *op = out_my + o;
                        uint32_t used = e >> 8;
                        op[0] = (uint8_t)e; wnd1 >>= used;
                        e = bz_lds_u16(lit_sa | (wnd1 & LMASK));
                        if (e < 0x1000u) {
                            op[1] = (uint8_t)e; wnd1 >>= (e >> 8); used += e >> 8;
                            e = bz_lds_u16(lit_sa | (wnd1 & LMASK));
                            if (e < 0x1000u) { op[2] = (uint8_t)e; used += e >> 8; o += 3; }
                            else o += 2;
                        } else o += 1;
                        br.drop(used);
                    } else {
                        int sy = bz_symbol(br, Tm.lit_cnt, Tm.lit_sym, lit_m, LB);
                        if (sy < 0) cmd = 3;
                        else if (sy < 256) { if (o >= cap) cmd = 3; else out_my[o++] = (uint8_t)sy; }
                        else if (sy == 256) cmd = 2;
                        else if (sy >= 257 + 29) cmd = 3;
                        else {
                            sy -= 257;
                            if (sy < 8) mlen = 3u + (uint32_t)sy;
                            else if (sy == 28) mlen = 258u;
                            else { const int x = (sy - 4) >> 2; mlen = 3u + ((4u + ((uint32_t)sy & 3u)) << x) + br.bits(x); }
                            const int dc = bz_symbol(br, Tm.dist_cnt, Tm.dist_sym, Tm.dist_fast, BZ_DIST_BITS);
                            if (dc < 0 || dc >= 30) cmd = 3;
                            else {
                                if (dc < 4) dist = 1u + (uint32_t)dc;
                                else { const int x = (dc - 2) >> 1; dist = 1u + ((2u + ((uint32_t)dc & 1u)) << x) + br.bits(x); }
                                if (dist > o || o + mlen > cap) cmd = 3;
                                else if (dist >= mlen && mlen <= BZ_SELF_MAX) {
                                    // the leader's own copy: all loads of a piece before its stores (one memory round trip per piece).
                                    // Away from the block's end a piece is stored whole -- the bytes behind the match are garbage that
                                    // the following symbols overwrite (output is strictly sequential) -- so no per-byte predicates;
                                    // the loads may read up to 19 bytes past the match, inside the padded buffer.
                                    const uint8_t *src = out_my + o - dist; uint8_t *dst = out_my + o;
                                    if (mlen <= 8u && o + 8u <= cap) {
                                        const uint32_t *al = reinterpret_cast<const uint32_t *>((uintptr_t)src & ~(uintptr_t)3);
                                        const uint32_t s8 = (uint32_t)((uintptr_t)src & 3) * 8;
                                        const uint32_t a0 = al[0], a1 = al[1], a2 = al[2];
                                        const uint32_t w0 = __funnelshift_r(a0, a1, s8), w1 = __funnelshift_r(a1, a2, s8);
                                        dst[0] = (uint8_t)w0; dst[1] = (uint8_t)(w0 >> 8); dst[2] = (uint8_t)(w0 >> 16); dst[3] = (uint8_t)(w0 >> 24);
                                        dst[4] = (uint8_t)w1; dst[5] = (uint8_t)(w1 >> 8); dst[6] = (uint8_t)(w1 >> 16); dst[7] = (uint8_t)(w1 >> 24);
                                    } else {
                                        uint32_t q = 0;
                                        do {
                                            const uint32_t *al = reinterpret_cast<const uint32_t *>((uintptr_t)(src + q) & ~(uintptr_t)3);
                                            const uint32_t s8 = (uint32_t)((uintptr_t)(src + q) & 3) * 8;
                                            const uint32_t a0 = al[0], a1 = al[1], a2 = al[2], a3 = al[3], a4 = al[4];
                                            uint32_t wv[4] = {__funnelshift_r(a0, a1, s8), __funnelshift_r(a1, a2, s8), __funnelshift_r(a2, a3, s8), __funnelshift_r(a3, a4, s8)};
                                            uint8_t *dp = dst + q;
                                            if (o + q + 16u <= cap) {
#pragma unroll
                                                for (int k = 0; k < 16; k++) dp[k] = (uint8_t)(wv[k >> 2] >> (8 * (k & 3)));
                                            } else {
                                                const uint32_t n = mlen - q;
#pragma unroll
                                                for (int k = 0; k < 16; k++) if ((uint32_t)k < n) dp[k] = (uint8_t)(wv[k >> 2] >> (8 * (k & 3)));
                                            }
                                            q += 16;
                                        } while (q < mlen);
                                    }
                                    o += mlen;
                                } else cmd = 1;
                            }
                        }
                    }
                }
            } while (!bz_any(cmd != 0));
        }
        // ---- events, one decoder at a time, all lanes ----
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int L = j * STRIDE;
            const int cj = __shfl_sync(FULLM, cmd, L);           // (also orders the leader's literal stores before the copy below)
            if (cj == 0) continue;
            if (cj == 1) {
                const uint32_t m = __shfl_sync(FULLM, mlen, L), dd = __shfl_sync(FULLM, dist, L), oo = __shfl_sync(FULLM, o, L);
                uint8_t *dst = bz_shfl_ptr(out_my, L) + oo;
                __syncwarp();
                // byte oo+q of a match equals byte oo-dd+(q mod dd): every source lies before oo, so all lanes copy at once
                const uint8_t *src = dst - dd;
                if (dd >= m) { for (uint32_t q = lane; q < m; q += 32) dst[q] = src[q]; }
                else if (m <= 2u * dd) { for (uint32_t q = lane; q < m; q += 32) dst[q] = src[q >= dd ? q - dd : q]; }
                else { for (uint32_t q = lane; q < m; q += 32) dst[q] = src[q % dd]; }
                if (lane == L) o += m;
                __syncwarp();
            } else if (cj == 2) {
                if (lane == L) state = br.consumed_bits() > in_bits + 7 ? BZS_ERROR : (last ? BZS_DONE : BZS_HEADER);
            } else if (lane == L) state = BZS_ERROR;
        }
        if (!bz_any(leader && (state == BZS_HEADER || state == BZS_SYMBOLS))) break;
    }
    __syncwarp();
    return __shfl_sync(FULLM, (state == BZS_DONE && o == cap) ? (int)o : -1, my * STRIDE);
}

#ifndef PD_BZ_MULTI_LIT_BITS
#define PD_BZ_MULTI_LIT_BITS 9       // first-level literal table of the K-decoder kernels: 1 KB per decoder.  Measured (B200, flavour B, K = 2,
                                     // 6 chunks in flight): 10 bits / 6 CTAs per SM 103.8 ms, 9 bits / 9 CTAs 89.7 ms per 20 M reads -- resident
                                     // decoders matter more than the few 10-bit codes that now take the slow path
#endif
#ifndef PD_BZ_MULTI_CTAS
#define PD_BZ_MULTI_CTAS 8           // (9 fitted before the 1 KB CRC table; 8 vs 9 measured 91.0 vs 89.7 ms) resident CTAs per SM the K-decoder kernels are compiled for (register cap)
#endif
template <int K, int WPC>
__global__ void __launch_bounds__(WPC * 32, PD_BZ_MULTI_CTAS) k_bgzf_decode_multi(BzArgs a)
{
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int STRIDE = 32 / K;
    constexpr int LB = PD_BZ_MULTI_LIT_BITS;
    __shared__ uint16_t lit_raw[(WPC * K << LB) + (1 << BZ_LIT_BITS)];   // + 2 KB: the tables start at a 2 KB boundary of the shared window
    __shared__ BzTables tabs[WPC][K];
    __shared__ uint32_t rec_offs[WPC][32];
    __shared__ uint32_t crc_tab[BZ_CRC_TAB_WORDS];
    bz_crc_table_init(crc_tab, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lit_sa0 = bz_align_tables(lit_raw) + (uint32_t)(warp * K) * (2u << LB);
    for (;;) {
        uint32_t b0 = 0;
        if (lane == 0) b0 = atomicAdd(a.ticket, (uint32_t)K);
        b0 = __shfl_sync(FULLM, b0, 0);
        if (b0 >= a.n_blocks) return;
        const int got_my = bz_inflate_multi<K, LB>(a, b0, tabs[warp], lit_sa0, lane);
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (b0 + j >= a.n_blocks) break;
            int got = __shfl_sync(FULLM, got_my, j * STRIDE);
            const BgzfBlockDesc k = a.blk[b0 + j];
            if (a.check_crc && k.out_len && got == (int)k.out_len) {
                const uint8_t *tr = a.in + k.in_off + k.in_len;                   // BGZF trailer: CRC32, ISIZE
                const uint32_t want = (uint32_t)tr[0] | (uint32_t)tr[1] << 8 | (uint32_t)tr[2] << 16 | (uint32_t)tr[3] << 24;
                if (bz_crc32_warp(a.out + k.out_off, k.out_len, crc_tab, lane) != want) got = -2;
            }
            bz_walk_block(a, b0 + j, k, a.out + k.out_off, got, rec_offs[warp], lane);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_bgzf_offsets: r0[b], c0[b] = exclusive prefix sums of n_reads / n_cigar; totals[0..2] = reads, CIGAR words, OR of flags.
// One CTA of 1024 threads; a chunk has at most a few hundred thousand blocks.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_bgzf_offsets(const BgzfSummary *sum, uint32_t n_blocks, uint32_t *r0, uint32_t *c0, uint32_t *totals)
{
    __shared__ uint32_t w_r[32], w_c[32], w_f[32];
    __shared__ uint32_t carry_r, carry_c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_r = 0; carry_c = 0; }
    uint32_t fl = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024) {
        const uint32_t b = base + threadIdx.x;
        uint32_t r = 0, c = 0;
        if (b < n_blocks) { r = sum[b].n_reads; c = sum[b].n_cigar; fl |= sum[b].flags; }
        uint32_t ir = r, ic = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t yr = __shfl_up_sync(0xffffffffu, ir, o), yc = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) { ir += yr; ic += yc; }
        }
        if (lane == 31) { w_r[warp] = ir; w_c[warp] = ic; }
        __syncthreads();
        uint32_t pr = carry_r, pc = carry_c;
        for (int w = 0; w < warp; w++) { pr += w_r[w]; pc += w_c[w]; }
        if (b < n_blocks) { r0[b] = pr + ir - r; c0[b] = pc + ic - c; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_r = pr + ir; carry_c = pc + ic; }
        __syncthreads();
    }
    fl = __reduce_or_sync(0xffffffffu, fl);
    if (lane == 0) w_f[warp] = fl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t f = 0;
        for (int w = 0; w < 32; w++) f |= w_f[w];
        totals[0] = carry_r; totals[1] = carry_c; totals[2] = f; totals[3] = n_blocks;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_bgzf_compact: block-local staging -> dense arrays.  cigar_off[i] is absolute into cigar[]; cigar_off[n_total] is written
// by the last block's warp.
// ---------------------------------------------------------------------------------------------------------------
struct BzCompactArgs {
    const BgzfBlockDesc *blk; const BgzfSummary *sum; const uint32_t *r0; const uint32_t *c0; uint32_t n_blocks;
    const int32_t *st_pos; const int32_t *st_tid; const uint32_t *st_coff; const uint32_t *st_cig;
    int32_t *pos; int32_t *tid; uint32_t *cigar_off; uint32_t *cigar;
};

__global__ void __launch_bounds__(256) k_bgzf_compact(BzCompactArgs a)
{
    const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= a.n_blocks) return;
    const int lane = threadIdx.x & 31;
    const uint64_t rb = bz_rec_base(a.blk[b].out_off, b), cb = bz_cig_base(a.blk[b].out_off, b);
    const uint32_t n = a.sum[b].n_reads, nc = a.sum[b].n_cigar, r0 = a.r0[b], c0 = a.c0[b];
    for (uint32_t i = lane; i < n; i += 32) {
        a.pos[r0 + i] = a.st_pos[rb + i];
        a.tid[r0 + i] = a.st_tid[rb + i];
        a.cigar_off[r0 + i] = c0 + a.st_coff[rb + i];
    }
    for (uint32_t j = lane; j < nc; j += 32) a.cigar[c0 + j] = a.st_cig[cb + j];
    if (b == a.n_blocks - 1 && lane == 0) a.cigar_off[r0 + n] = c0 + nc;
}

// ---------------------------------------------------------------------------------------------------------------
// k_delta_scatter_multi: the delta-scatter walk (PanDepth.cpp:735-760) for records of ANY contig (per-record tid): blocks of
// a decoded chunk that hold a contig change or records out of order.  The device decoder already applied the filter.
// ---------------------------------------------------------------------------------------------------------------
struct ScatterMultiArgs {
    const int32_t *tid; const int32_t *pos; const uint32_t *cigar; const uint32_t *cigar_off;
    uint32_t n;
    const uint64_t *contig_off; const uint32_t *contig_len; const uint32_t *contig_org; const uint8_t *owned; int32_t n_targets;
    uint32_t *delta;   // array base
};                     // contig_len / contig_org: owned length and first owned base per contig (range sharding)

__global__ void __launch_bounds__(256) k_delta_scatter_multi(ScatterMultiArgs a)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int32_t t = a.tid[i];
    if (t < 0 || t >= a.n_targets || !a.owned[t]) return;
    const int32_t p = a.pos[i];
    if (p < 0) return;
    const uint32_t c0 = a.cigar_off[i], nops = a.cigar_off[i + 1] - c0;
    const int64_t clen = a.contig_len[t];
    uint32_t *delta = a.delta + a.contig_off[t];
    int64_t q = (int64_t)p - (int64_t)a.contig_org[t], rs = 0, re = 0; bool open = false;
    auto flush = [&]() {
        if (!open) return;
        const int64_t s = rs < 0 ? 0 : rs, e = re > clen ? clen : re;
        if (s >= e) return;
#ifndef PD_WARP_EMU
        asm volatile("red.global.add.u32 [%0], %1;" ::"l"(delta + s), "r"(1u) : "memory");
        asm volatile("red.global.add.u32 [%0], %1;" ::"l"(delta + e), "r"(0xffffffffu) : "memory");
#else
        atomicAdd(delta + s, 1u); atomicAdd(delta + e, 0xffffffffu);
#endif
    };
    for (uint32_t k = 0; k < nops; ++k) {
        const uint32_t w = a.cigar[c0 + k];
        const uint32_t op = w & 15u;
        const int64_t n = w >> 4;
        if ((0x181u >> op) & 1u) {                               // M = X
            if (open && re == q) re = q + n; else { flush(); rs = q; re = q + n; open = true; }
            q += n;
        } else if ((0x00cu >> op) & 1u) q += n;                  // D N
    }
    flush();
}

}  // namespace pdk
