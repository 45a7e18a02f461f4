// fastbam.cpp -- see fastbam.hpp.  BAM / BGZF / BAI layouts follow the SAM/BAM specification (SAMv1.pdf sections 4.1, 4.2, 5.2).
#include "fastbam.hpp"
#include <algorithm>

#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdio>
#include <cstring>

#ifdef PD_HAVE_LIBDEFLATE
// libdeflate's public C API (the reference vendors lib/libdeflate.a without its header)
extern "C" {
struct libdeflate_decompressor;
struct libdeflate_decompressor *libdeflate_alloc_decompressor(void);
int libdeflate_deflate_decompress(struct libdeflate_decompressor *d, const void *in, size_t in_nbytes, void *out,
                                  size_t out_nbytes_avail, size_t *actual_out_nbytes_ret);
void libdeflate_free_decompressor(struct libdeflate_decompressor *d);
}
#endif

namespace pdh {

// ---------------------------------------------------------------------------------------------------------------
// .bai linear index
// ---------------------------------------------------------------------------------------------------------------
bool BaiLinear::load(const std::string &path)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    auto rd = [&](void *p, size_t n) { return fread(p, 1, n, fp) == n; };
    char magic[4]; int32_t n_ref = 0;
    bool ok = rd(magic, 4) && memcmp(magic, "BAI\1", 4) == 0 && rd(&n_ref, 4) && n_ref >= 0;
    if (ok) refs.assign((size_t)n_ref, Ref());
    for (int32_t r = 0; ok && r < n_ref; r++) {
        int32_t n_bin = 0;
        ok = rd(&n_bin, 4) && n_bin >= 0;
        for (int32_t b = 0; ok && b < n_bin; b++) {
            uint32_t bin; int32_t n_chunk;
            ok = rd(&bin, 4) && rd(&n_chunk, 4) && n_chunk >= 0;
            if (!ok) break;
            if (bin == 37450 && n_chunk >= 1) {             // pseudo-bin: chunk 0 = (first, last) virtual offsets of this contig's data
                uint64_t c[2]; ok = rd(c, 16);
                refs[(size_t)r].data_beg = c[0]; refs[(size_t)r].data_end = c[1];
                if (ok && n_chunk > 1) ok = fseek(fp, (long)(n_chunk - 1) * 16, SEEK_CUR) == 0;
            } else ok = fseek(fp, (long)n_chunk * 16, SEEK_CUR) == 0;
        }
        int32_t n_intv = 0;
        ok = ok && rd(&n_intv, 4) && n_intv >= 0;
        if (ok && n_intv) { refs[(size_t)r].ioff.resize((size_t)n_intv); ok = rd(refs[(size_t)r].ioff.data(), (size_t)n_intv * 8); }
    }
    fclose(fp);
    if (!ok) refs.clear();
    return ok;
}

uint64_t BaiLinear::start_voffset(int tid, int64_t beg) const
{
    if (tid < 0 || (size_t)tid >= refs.size()) return 0;
    const Ref &r = refs[(size_t)tid];
    if (beg < 0) beg = 0;
    size_t w = (size_t)(beg >> 14);
    // window w and later: the first non-zero entry is the first read overlapping that window; zeros (older writers leave
    // them for empty windows) mean "no read overlaps", so later windows' entries are still upper bounds of nothing missed
    for (; w < r.ioff.size(); w++) if (r.ioff[w]) return r.ioff[w];
    return 0;                                                // no read overlaps beg or anything after it
}

// ---------------------------------------------------------------------------------------------------------------
// scanner
// ---------------------------------------------------------------------------------------------------------------
BamRangeScanner::BamRangeScanner()
{
#ifdef PD_HAVE_LIBDEFLATE
    dec_ = libdeflate_alloc_decompressor();
#endif
    raw_.resize(2u << 20);
    buf_.resize(1u << 20);
}

BamRangeScanner::~BamRangeScanner()
{
    if (fd_ >= 0) close(fd_);
#ifdef PD_HAVE_LIBDEFLATE
    if (dec_) libdeflate_free_decompressor((libdeflate_decompressor *)dec_);
#endif
}

bool BamRangeScanner::open(const std::string &path)
{
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) { err_ = "cannot open " + path; return false; }
    return true;
}

bool BamRangeScanner::inflate_next_block()
{
    for (;;) {
        // need the fixed 18-byte BGZF header (12 gzip + 6 extra) before anything else
        if (raw_len_ - raw_pos_ < (1u << 16) + 64 && !eof_) {
            if (raw_pos_) { memmove(raw_.data(), raw_.data() + raw_pos_, raw_len_ - raw_pos_); raw_len_ -= raw_pos_; raw_pos_ = 0; }
            // read-ahead grows with the scan (128 KB, 256 KB, ... up to the 2 MB buffer): a range that ends after a few
            // records (tens of thousands of small contigs) must not pay for 2 MB of page-cache copies
            while (raw_len_ < raw_.size()) {
                const size_t want = std::min(raw_.size() - raw_len_, read_ahead_);
                ssize_t k = pread(fd_, raw_.data() + raw_len_, want, (off_t)file_pos_);
                if (k < 0) { err_ = "read error"; return false; }
                if (k == 0) { eof_ = true; break; }
                raw_len_ += (size_t)k; file_pos_ += (uint64_t)k;
                if (raw_len_ - raw_pos_ >= (1u << 16) + 64) break;
            }
            if (read_ahead_ < raw_.size()) read_ahead_ *= 2;
        }
        const size_t avail = raw_len_ - raw_pos_;
        if (avail == 0) return false;                                         // clean EOF
        if (avail < 18) { err_ = "truncated BGZF header"; return false; }
        const uint8_t *h = raw_.data() + raw_pos_;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { err_ = "not a BGZF block"; return false; }
        const uint32_t xlen = le16(h + 10);
        if (avail < 12 + xlen) { err_ = "truncated BGZF extra field"; return false; }
        uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const uint8_t *f = h + 12 + x;
            const uint32_t slen = le16(f + 2);
            if (f[0] == 'B' && f[1] == 'C' && slen == 2) { bsize = (uint32_t)le16(f + 4) + 1; break; }
            x += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 || bsize > avail) { err_ = "bad BGZF block size"; return false; }
        const uint8_t *cdata = h + 12 + xlen;
        const size_t clen = bsize - xlen - 12 - 8;
        const uint32_t isize = le32(h + bsize - 4);
        raw_pos_ += bsize;
        if (isize == 0) continue;                                             // empty block (EOF marker)
        // keep the unconsumed tail, append this block
        if (bpos_ && bpos_ == blen_) { bpos_ = blen_ = 0; }
        if (blen_ + isize > buf_.size()) {
            if (bpos_) { memmove(buf_.data(), buf_.data() + bpos_, blen_ - bpos_); blen_ -= bpos_; bpos_ = 0; }
            if (blen_ + isize > buf_.size()) buf_.resize((blen_ + isize) * 2);
        }
#ifdef PD_HAVE_LIBDEFLATE
        size_t got = 0;
        if (libdeflate_deflate_decompress((libdeflate_decompressor *)dec_, cdata, clen, buf_.data() + blen_, isize, &got) != 0 || got != isize) {
            err_ = "inflate failed"; return false;
        }
#else
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) { err_ = "inflateInit2 failed"; return false; }
        zs.next_in = (Bytef *)cdata; zs.avail_in = (uInt)clen; zs.next_out = buf_.data() + blen_; zs.avail_out = isize;
        const int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END || zs.avail_out != 0) { err_ = "inflate failed"; return false; }
#endif
        blen_ += isize;
        return true;
    }
}

bool BamRangeScanner::fill_slow(size_t need)
{
    while (blen_ - bpos_ < need) if (!inflate_next_block()) return false;
    return true;
}

// CIGARs with more than 65535 ops are stored in the CG:B,I tag behind a placeholder `<l_seq>S<ref_len>N` (SAMv1 4.2.2);
// htslib's reader swaps the real one in, and so must we.
bool BamRangeScanner::resolve_long_cigar(const uint8_t *rec, uint32_t block_size, uint32_t l_read_name, uint32_t n_cigar, int32_t l_seq, RawRead &r)
{
    const uint8_t *cg = rec + 32 + l_read_name;
    const uint32_t c0 = le32(cg), c1 = le32(cg + 4);
    if ((c0 & 15u) != 4u || (int32_t)(c0 >> 4) != l_seq || (c1 & 15u) != 3u) return false;
    size_t p = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + (size_t)((l_seq + 1) / 2) + (size_t)l_seq;
    while (p + 3 <= block_size) {
        const uint8_t t0 = rec[p], t1 = rec[p + 1], ty = rec[p + 2];
        p += 3;
        size_t sz = 0;
        switch (ty) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'd': sz = 8; break;
        case 'Z': case 'H': while (p + sz < block_size && rec[p + sz]) sz++; sz++; break;
        case 'B': {
            if (p + 5 > block_size) return false;
            const uint8_t sub = rec[p]; const uint32_t cnt = le32(rec + p + 1);
            const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            if (t0 == 'C' && t1 == 'G' && sub == 'I') {
                if (p + 5 + (size_t)cnt * 4 > block_size) return false;
                cg_.resize(cnt);
                memcpy(cg_.data(), rec + p + 5, (size_t)cnt * 4);
                r.n_cigar = cnt; r.cigar = reinterpret_cast<const uint8_t *>(cg_.data());
                return true;
            }
            sz = 5 + (size_t)cnt * es;
            break;
        }
        default: return false;
        }
        p += sz;
    }
    return false;
}

}  // namespace pdh
