// CPU checker for pandepth_b200/csrc/pd_inflate.cuh (host build of the host+device decoder):
//   inflate_check <file.bam|file.gz ...>   every BGZF block of each file: inflate_raw == zlib, byte for byte; BAM files are
//                                          also walked block by block (walk_bam_block) and the records counted
//   inflate_check --synth                  zlib-made raw deflate streams: levels 0-9, strategies default / filtered /
//                                          huffman-only / RLE / fixed, inputs from empty to 64 KB, random / text / runs
//   plus corrupt streams: truncated, bit-flipped (must return an error or a wrong-size result, never touch memory outside)
#include "pd_inflate.cuh"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
using namespace pdk;

static long long n_blocks = 0, n_bad = 0, n_records = 0, n_straddle = 0;

static bool zlib_raw(const uint8_t *in, size_t n, std::vector<uint8_t> &out, size_t isize)
{
    out.assign(isize + 1, 0);
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = (Bytef *)in; zs.avail_in = (uInt)n; zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    const int rc = inflate(&zs, Z_FINISH);
    const size_t got = out.size() - zs.avail_out;
    inflateEnd(&zs);
    out.resize(got);
    return rc == Z_STREAM_END && got == isize;
}

static void check_stream(const uint8_t *in, size_t n, size_t isize, const char *what)
{
    std::vector<uint8_t> want;
    if (!zlib_raw(in, n, want, isize)) { fprintf(stderr, "%s: zlib itself rejects the stream\n", what); n_bad++; return; }
    static InflateTables T;
    std::vector<uint8_t> got(isize + 64, 0xAB);
    const int r = inflate_raw(in, (uint32_t)n, got.data(), (uint32_t)isize, T);
    n_blocks++;
    if (r != (int)isize || memcmp(got.data(), want.data(), isize) != 0) { n_bad++; if (n_bad < 10) fprintf(stderr, "%s: mismatch (ret %d, isize %zu)\n", what, r, isize); }
    for (size_t k = isize; k < got.size(); k++) if (got[k] != 0xAB) { n_bad++; fprintf(stderr, "%s: wrote past the output\n", what); break; }
}

static int check_file(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return 1; }
    std::vector<uint8_t> data; uint8_t tmp[1 << 16]; size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) data.insert(data.end(), tmp, tmp + k);
    fclose(f);
    const bool is_bam = strlen(path) > 4 && strcmp(path + strlen(path) - 4, ".bam") == 0;
    size_t p = 0; bool header_done = !is_bam; std::vector<uint8_t> stream;          // BAM: skip the header on the re-assembled stream
    std::vector<uint8_t> carry;
    while (p + 18 <= data.size()) {
        const uint8_t *h = data.data() + p;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { fprintf(stderr, "%s: not BGZF at %zu\n", path, p); return 1; }
        const uint32_t xlen = h[10] | h[11] << 8;
        uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) { const uint8_t *e = h + 12 + x; const uint32_t sl = e[2] | e[3] << 8; if (e[0] == 'B' && e[1] == 'C' && sl == 2) { bsize = (e[4] | e[5] << 8) + 1; break; } x += 4 + sl; }
        if (!bsize || p + bsize > data.size()) break;
        const uint8_t *cd = h + 12 + xlen; const size_t cl = bsize - xlen - 12 - 8;
        const uint32_t isize = inf_le32(h + bsize - 4);
        if (isize) {
            check_stream(cd, cl, isize, path);
            if (is_bam) {
                static InflateTables T;
                std::vector<uint8_t> blk(isize);
                if (inflate_raw(cd, (uint32_t)cl, blk.data(), isize, T) == (int)isize) {
                    carry.insert(carry.end(), blk.begin(), blk.end());
                    if (!header_done && carry.size() >= 12) {                       // magic, l_text, text, n_ref, refs
                        size_t q = 8 + inf_le32(carry.data() + 4);
                        if (carry.size() >= q + 4) {
                            uint32_t n_ref = inf_le32(carry.data() + q); q += 4; bool ok = true;
                            for (uint32_t r = 0; r < n_ref && ok; r++) { if (carry.size() < q + 4) { ok = false; break; } q += 4 + inf_le32(carry.data() + q) + 4; }
                            if (ok && carry.size() >= q) { carry.erase(carry.begin(), carry.begin() + (long)q); header_done = true; }
                        }
                    }
                    if (header_done) {
                        const int64_t used = walk_bam_block(carry.data(), (uint32_t)carry.size(), [&](int32_t, int32_t, uint16_t, uint8_t, uint32_t, const uint8_t *) { n_records++; });
                        if (used < 0) { n_bad++; fprintf(stderr, "%s: malformed record\n", path); carry.clear(); }
                        else { carry.erase(carry.begin(), carry.begin() + (long)used); if (!carry.empty()) n_straddle++; }
                    }
                }
            }
        }
        p += bsize;
    }
    return 0;
}

static void synth()
{
    std::mt19937_64 rng(99);
    const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
    for (int kind = 0; kind < 5; kind++)
        for (size_t n : {(size_t)0, (size_t)1, (size_t)2, (size_t)257, (size_t)4096, (size_t)65280, (size_t)65535}) {
            std::vector<uint8_t> src(n);
            for (size_t i = 0; i < n; i++) {
                switch (kind) {
                case 0: src[i] = (uint8_t)rng(); break;                                          // incompressible
                case 1: src[i] = (uint8_t)("ACGT"[rng() & 3]); break;                            // 2 bits of entropy
                case 2: src[i] = (uint8_t)(i / 300);  break;                                     // long runs (overlapping matches)
                case 3: src[i] = (uint8_t)((i % 46 < 36) ? (i % 46) * 7 : rng() % 3); break;    // record-like period 46
                default: src[i] = (uint8_t)(rng() % 40 + 33); break;                            // quality-like
                }
            }
            for (int level = 0; level <= 9; level++)
                for (int st : strategies) {
                    z_stream zs; memset(&zs, 0, sizeof zs);
                    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, st) != Z_OK) continue;
                    std::vector<uint8_t> z(deflateBound(&zs, (uLong)n) + 64);
                    zs.next_in = src.data(); zs.avail_in = (uInt)n; zs.next_out = z.data(); zs.avail_out = (uInt)z.size();
                    deflate(&zs, Z_FINISH);
                    const size_t zn = z.size() - zs.avail_out;
                    deflateEnd(&zs);
                    check_stream(z.data(), zn, n, "synthetic");
                    // corrupt copies: must not crash or write outside; result must not claim success with wrong bytes
                    if (n >= 257 && level % 3 == 0) {
                        static InflateTables T;
                        std::vector<uint8_t> out(n + 64, 0xCD);
                        for (int t = 0; t < 6; t++) {
                            std::vector<uint8_t> bad(z.begin(), z.begin() + (long)zn);
                            if (t < 3) bad.resize(zn * (size_t)(t + 1) / 4);                     // truncated
                            else bad[rng() % zn] ^= (uint8_t)(1u << (rng() & 7));                // one bit flipped
                            const int r = inflate_raw(bad.data(), (uint32_t)bad.size(), out.data(), (uint32_t)n, T);
                            for (size_t k = n; k < out.size(); k++) if (out[k] != 0xCD) { n_bad++; fprintf(stderr, "corrupt stream wrote past the output\n"); break; }
                            if (t < 3 && r == (int)n && n > 0 && bad.size() < zn - 8) { /* a truncated stream cannot produce everything */ n_bad++; fprintf(stderr, "truncated stream accepted\n"); }
                        }
                    }
                }
        }
}

int main(int argc, char **argv)
{
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--synth")) synth();
        else if (check_file(argv[i])) return 2;
    }
    printf("blocks %lld, records %lld, blocks ending inside a record %lld, mismatches %lld\n", n_blocks, n_records, n_straddle, n_bad);
    return n_bad ? 1 : 0;
}
