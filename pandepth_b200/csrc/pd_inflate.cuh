// pd_inflate.cuh -- DEFLATE (RFC 1951) decoder and BAM record walker written as host+device functions.
//
// The sequential restatement of what pd_bgzf.cuh does warp-cooperatively on the device: one function inflates a raw deflate
// stream, one walks the BAM records of an inflated block.  A BGZF file is a sequence of independent <= 64 KB deflate streams,
// and htslib's writer starts every block on a record boundary unless a single record is larger than a block (bam_write1
// flushes first), so blocks can be inflated AND parsed independently.
//
// TEST INFRASTRUCTURE: used by tests/cpp (inflate_check, decode_check, bgzf_emu_check: the reference the emulated device
// kernels are compared with) and by the CPU stand-in engine of tests/stand_in.  Verified against zlib on every BGZF block of
// generated BAMs and on zlib streams at all levels / strategies.  The product path (pd_engine.cu) does not include this file.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PD_HD __host__ __device__ __forceinline__
#else
#define PD_HD inline
#endif

namespace pdk {

// Decoder tables of one block.  First-level lookup of LIT_BITS / DIST_BITS bits: entry = (symbol << 4) | code length, 0 =
// code longer than the table (resolved canonically from count[] / symbol order, like RFC 1951 3.2.2 describes the code).
constexpr int INF_LIT_BITS = 10, INF_DIST_BITS = 8;
struct InflateTables {
    uint16_t lit_fast[1 << INF_LIT_BITS];
    uint16_t dist_fast[1 << INF_DIST_BITS];
    uint16_t lit_count[16], lit_sym[288];
    uint16_t dist_count[16], dist_sym[32];
    uint8_t lens[19 + 286 + 30 + 5];                           // code-length code (19) followed by the literal/length and distance lengths
};                                                           // 3.5 KB: fits a warp's slice of shared memory

enum { INF_OK = 0, INF_ERR_INPUT = -1, INF_ERR_OUTPUT = -2, INF_ERR_CODE = -3, INF_ERR_DIST = -4, INF_ERR_BLOCK = -5 };

struct BitReader {
    const uint8_t *in; uint32_t len, pos; uint64_t buf; int cnt;
    PD_HD void init(const uint8_t *p, uint32_t n) { in = p; len = n; pos = 0; buf = 0; cnt = 0; }
    // at least 32 valid bits in buf while input lasts (bytes past the end read as zero; over-consumption is detected at the end)
    PD_HD void refill()
    {
        if (cnt > 32) return;
        if (pos + 4 <= len) {                                // four bytes at once (byte loads: the stream is not aligned)
            const uint32_t w = (uint32_t)in[pos] | (uint32_t)in[pos + 1] << 8 | (uint32_t)in[pos + 2] << 16 | (uint32_t)in[pos + 3] << 24;
            buf |= (uint64_t)w << cnt; pos += 4; cnt += 32;
        } else while (cnt <= 56) { buf |= (uint64_t)(pos < len ? in[pos] : 0) << cnt; pos++; cnt += 8; }
    }
    PD_HD uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    PD_HD void drop(int n) { buf >>= n; cnt -= n; }
    PD_HD uint32_t bits(int n) { refill(); const uint32_t v = peek(n); drop(n); return v; }
    PD_HD bool overrun() const { return (int64_t)pos - (cnt >> 3) > (int64_t)len; }       // consumed more bytes than exist
};

PD_HD uint32_t inf_reverse(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; } return r; }

// canonical Huffman code from code lengths (RFC 1951 3.2.2): count[], symbols in code order, and the first-level table
PD_HD int inf_build(const uint8_t *lens, int n, uint16_t *count, uint16_t *sym, uint16_t *fast, int fast_bits)
{
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[lens[i]]++;
    for (int i = 0; i < (1 << fast_bits); i++) fast[i] = 0;
    if (count[0] == n) return INF_OK;                        // no codes at all (legal for the distance code of a literal-only block)
    int left = 1;
    for (int l = 1; l < 16; l++) { left <<= 1; left -= count[l]; if (left < 0) return INF_ERR_CODE; }   // over-subscribed
    uint16_t offs[16]; offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int s = 0; s < n; s++) if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
    // first-level table: every fast_bits-bit pattern whose low `l` bits (LSB first = reversed code) are a code of length l
    uint32_t code = 0; int idx = 0;
    for (int l = 1; l <= fast_bits; l++) {
        for (int k = 0; k < count[l]; k++, idx++, code++) {
            const uint32_t rev = inf_reverse(code, l);
            for (uint32_t p = rev; p < (1u << fast_bits); p += 1u << l) fast[p] = (uint16_t)((sym[idx] << 4) | l);
        }
        code <<= 1;
    }
    return left;                                             // > 0: incomplete code (only legal with a single code of length 1)
}

// one symbol: table hit, else bit-by-bit canonical decode of the longer codes
PD_HD int inf_symbol(BitReader &br, const uint16_t *count, const uint16_t *sym, const uint16_t *fast, int fast_bits)
{
    br.refill();
    const uint16_t e = fast[br.peek(fast_bits)];
    if (e) { br.drop(e & 15); return e >> 4; }
    int code = 0, first = 0, index = 0;
    uint64_t b = br.buf;
    for (int l = 1; l < 16; l++) {
        code |= (int)(b & 1); b >>= 1;
        const int c = count[l];
        if (code - c < first) { br.drop(l); return sym[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return INF_ERR_CODE;
}

// Inflate one raw deflate stream.  Returns the number of bytes written (>= 0) or an INF_ERR_* code.  out_cap is the BGZF
// ISIZE, so a well-formed block fills it exactly.
PD_HD int inflate_raw(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_cap, InflateTables &T)
{
    const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    BitReader br; br.init(in, in_len);
    uint32_t o = 0;
    for (;;) {
        const uint32_t last = br.bits(1), type = br.bits(2);
        if (type == 0) {                                     // stored
            br.drop(br.cnt & 7);                             // to the byte boundary
            const uint32_t n = br.bits(16), nn = br.bits(16);
            if ((n ^ nn) != 0xffffu) return INF_ERR_BLOCK;
            // bytes still buffered belong to the stream: rewind the reader to the byte position
            uint32_t p = br.pos - (uint32_t)(br.cnt >> 3);
            if (p + n > in_len) return INF_ERR_INPUT;
            if (o + n > out_cap) return INF_ERR_OUTPUT;
            for (uint32_t k = 0; k < n; k++) out[o + k] = in[p + k];
            o += n; p += n;
            br.pos = p; br.buf = 0; br.cnt = 0;
        } else if (type == 1 || type == 2) {
            if (type == 1) {                                 // fixed code (RFC 1951 3.2.6)
                for (int i = 0; i < 144; i++) T.lens[i] = 8;
                for (int i = 144; i < 256; i++) T.lens[i] = 9;
                for (int i = 256; i < 280; i++) T.lens[i] = 7;
                for (int i = 280; i < 288; i++) T.lens[i] = 8;
                inf_build(T.lens, 288, T.lit_count, T.lit_sym, T.lit_fast, INF_LIT_BITS);
                for (int i = 0; i < 30; i++) T.lens[i] = 5;
                inf_build(T.lens, 30, T.dist_count, T.dist_sym, T.dist_fast, INF_DIST_BITS);
            } else {                                         // dynamic code (3.2.7)
                const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
                if (nlen > 286 || ndist > 30) return INF_ERR_CODE;
                for (int i = 0; i < 19; i++) T.lens[i] = 0;
                for (int i = 0; i < ncode; i++) T.lens[ORDER[i]] = (uint8_t)br.bits(3);
                // the code-length code is small: decode it through the distance tables' storage
                if (inf_build(T.lens, 19, T.dist_count, T.dist_sym, T.dist_fast, INF_DIST_BITS) != 0) return INF_ERR_CODE;
                int idx = 0;
                while (idx < nlen + ndist) {
                    const int s = inf_symbol(br, T.dist_count, T.dist_sym, T.dist_fast, INF_DIST_BITS);
                    if (s < 0) return INF_ERR_CODE;
                    if (s < 16) { T.lens[19 + idx++] = (uint8_t)s; continue; }
                    int rep, val = 0;
                    if (s == 16) { if (idx == 0) return INF_ERR_CODE; val = T.lens[19 + idx - 1]; rep = 3 + (int)br.bits(2); }
                    else if (s == 17) rep = 3 + (int)br.bits(3);
                    else rep = 11 + (int)br.bits(7);
                    if (idx + rep > nlen + ndist) return INF_ERR_CODE;
                    while (rep--) T.lens[19 + idx++] = (uint8_t)val;
                }
                if (T.lens[19 + 256] == 0) return INF_ERR_CODE;                       // no end-of-block code
                int r = inf_build(T.lens + 19, nlen, T.lit_count, T.lit_sym, T.lit_fast, INF_LIT_BITS);
                if (r < 0 || (r > 0 && nlen - T.lit_count[0] != 1)) return INF_ERR_CODE;
                r = inf_build(T.lens + 19 + nlen, ndist, T.dist_count, T.dist_sym, T.dist_fast, INF_DIST_BITS);
                if (r < 0 || (r > 0 && ndist - T.dist_count[0] != 1)) return INF_ERR_CODE;
            }
            for (;;) {
                // literal fast path: one refill leaves >= 33 bits, enough for two table hits (<= 10 bits each) in a row
                br.refill();
                uint16_t e = T.lit_fast[br.peek(INF_LIT_BITS)];
                if (e && e < (256u << 4)) {
                    if (o >= out_cap) return INF_ERR_OUTPUT;
                    br.drop(e & 15); out[o++] = (uint8_t)(e >> 4);
                    e = T.lit_fast[br.peek(INF_LIT_BITS)];
                    if (e && e < (256u << 4)) {
                        if (o >= out_cap) return INF_ERR_OUTPUT;
                        br.drop(e & 15); out[o++] = (uint8_t)(e >> 4);
                        e = T.lit_fast[br.peek(INF_LIT_BITS)];
                        if (e && e < (256u << 4)) {
                            if (o >= out_cap) return INF_ERR_OUTPUT;
                            br.drop(e & 15); out[o++] = (uint8_t)(e >> 4);
                        }
                    }
                    continue;
                }
                int s = inf_symbol(br, T.lit_count, T.lit_sym, T.lit_fast, INF_LIT_BITS);
                if (s < 0) return INF_ERR_CODE;
                if (s < 256) { if (o >= out_cap) return INF_ERR_OUTPUT; out[o++] = (uint8_t)s; continue; }
                if (s == 256) break;
                s -= 257;
                if (s >= 29) return INF_ERR_CODE;
                // length and distance codes (RFC 1951 3.2.5) are arithmetic: code s has e = (s-4)>>2 extra bits and base
                // 3 + ((4 + (s&3)) << e) (258 for the last one); distance code d has e = (d-2)>>1 and base 1 + ((2 + (d&1)) << e)
                uint32_t mlen;
                if (s < 8) mlen = 3u + (uint32_t)s;
                else if (s == 28) mlen = 258u;
                else { const int e = (s - 4) >> 2; mlen = 3u + ((4u + ((uint32_t)s & 3u)) << e) + br.bits(e); }
                const int d = inf_symbol(br, T.dist_count, T.dist_sym, T.dist_fast, INF_DIST_BITS);
                if (d < 0 || d >= 30) return INF_ERR_CODE;
                uint32_t dist;
                if (d < 4) dist = 1u + (uint32_t)d;
                else { const int e = (d - 2) >> 1; dist = 1u + ((2u + ((uint32_t)d & 1u)) << e) + br.bits(e); }
                if (dist > o) return INF_ERR_DIST;
                if (o + mlen > out_cap) return INF_ERR_OUTPUT;
                for (uint32_t k = 0; k < mlen; k++) out[o + k] = out[o + k - dist];   // byte order matters when the match overlaps itself
                o += mlen;
            }
        } else return INF_ERR_BLOCK;
        if (br.overrun()) return INF_ERR_INPUT;
        if (last) break;
    }
    return (int)o;
}

// ---------------------------------------------------------------------------------------------------------------
// BAM records of one inflated block that starts on a record boundary (SAMv1 4.2).  For every record the depth path needs
// refID, pos, flag, mapq and the CIGAR words; `emit(refID, pos, flag, mapq, n_cigar, cigar_bytes)` is called per record.
// Returns the offset of the first byte not consumed (== len when the block ends on a record boundary; smaller when the last
// record continues in the next block -- the caller stitches those on the host), or -1 if a record is malformed.
// ---------------------------------------------------------------------------------------------------------------
PD_HD uint32_t inf_le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

template <typename Emit>
PD_HD int64_t walk_bam_block(const uint8_t *buf, uint32_t len, Emit &&emit)
{
    uint32_t p = 0;
    while (p + 4 <= len) {
        const uint32_t bs = inf_le32(buf + p);
        if (bs < 32 || bs > (1u << 30)) return -1;
        if (p + 4 + bs > len) break;                         // continues in the next block
        const uint8_t *rec = buf + p + 4;
        const uint32_t l_name = rec[8], n_cigar = (uint32_t)rec[12] | (uint32_t)rec[13] << 8;
        if (32 + l_name + 4 * n_cigar > bs) return -1;
        emit((int32_t)inf_le32(rec), (int32_t)inf_le32(rec + 4), (uint16_t)((uint32_t)rec[14] | (uint32_t)rec[15] << 8), rec[9], n_cigar, rec + 32 + l_name);
        p += 4 + bs;
    }
    return p;
}

}  // namespace pdk

