// fastbam.cpp -- see fastbam.hpp.  BAM / BGZF / BAI layouts follow the SAM/BAM specification (SAMv1.pdf sections 4.1, 4.2, 5.2).
#include "fastbam.hpp"
#include <algorithm>

#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

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
        if (isize > (1u << 16)) { err_ = "BGZF block claims more than 64 KB"; return false; }
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
        if ((uint32_t)crc32(0L, buf_.data() + blen_, isize) != le32(h + bsize - 8)) { err_ = "BGZF block CRC mismatch"; return false; }   // htslib checks it too
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
bool resolve_long_cigar_tag(const uint8_t *rec, uint32_t block_size, uint32_t l_read_name, uint32_t n_cigar, int32_t l_seq,
                            std::vector<uint32_t> &cg_, RawRead &r)
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

// ---------------------------------------------------------------------------------------------------------------
// whole-file stream: ordered parallel inflate
// ---------------------------------------------------------------------------------------------------------------
namespace {

// one deflate payload -> out (isize bytes); thread-safe (own decompressor per call site)
struct Inflater {
#ifdef PD_HAVE_LIBDEFLATE
    libdeflate_decompressor *d = libdeflate_alloc_decompressor();
    ~Inflater() { if (d) libdeflate_free_decompressor(d); }
    bool run(const uint8_t *in, size_t n, uint8_t *out, size_t isize)
    {
        size_t got = 0;
        return d && libdeflate_deflate_decompress(d, in, n, out, isize, &got) == 0 && got == isize;
    }
#else
    bool run(const uint8_t *in, size_t n, uint8_t *out, size_t isize)
    {
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) return false;
        zs.next_in = (Bytef *)in; zs.avail_in = (uInt)n; zs.next_out = out; zs.avail_out = (uInt)isize;
        const int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        return rc == Z_STREAM_END && zs.avail_out == 0;
    }
#endif
};

int64_t bgzf_block(const uint8_t *h, size_t avail, const uint8_t **cdata, size_t *clen, uint32_t *isize) { return bgzf_block_at(h, avail, cdata, clen, isize); }

}  // namespace

int64_t bgzf_block_at(const uint8_t *h, size_t avail, const uint8_t **cdata, size_t *clen, uint32_t *isize)
{
    if (avail < 18) return 0;
    if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) return -1;
    const uint32_t xlen = le16(h + 10);
    if (avail < 12 + (size_t)xlen) return 0;
    uint32_t bsize = 0;
    for (uint32_t x = 0; x + 4 <= xlen;) {
        const uint8_t *f = h + 12 + x;
        const uint32_t slen = le16(f + 2);
        if (f[0] == 'B' && f[1] == 'C' && slen == 2) { bsize = (uint32_t)le16(f + 4) + 1; break; }
        x += 4 + slen;
    }
    if (bsize < 12 + xlen + 8) return -1;
    if (bsize > avail) return 0;
    *cdata = h + 12 + xlen; *clen = bsize - xlen - 12 - 8; *isize = le32(h + bsize - 4);
    if (*isize > (1u << 16)) return -1;                      // BGZF blocks inflate to at most 64 KB: a larger ISIZE is corruption, not a size to allocate
    return bsize;
}

// Groups of consecutive BGZF blocks (up to ~4 MB inflated / 2 MB compressed) are cut by whoever asks for work next: the cut
// needs the block headers, a sequential walk, so it happens under the lock on a shared read-ahead buffer and copies only the
// group's compressed bytes out.  The worker inflates outside the lock and hands the group to the consumer strictly in
// sequence through a small ring of slots.
struct BamStreamScanner::Impl {
    // group size: measured on 8 cores, 100 M SEQ-less reads / 10 M reads with SEQ+QUAL: 0.5 MB groups 1.9 / 1.6 s, 4 MB groups
    // 1.2 / 1.2 s, 16 MB groups 1.2 / 1.1 s (every hand-over is a lock + a wake-up of the consumer)
    static constexpr size_t GROUP_IN = 2u << 20, GROUP_OUT = 4u << 20, AHEAD = 8u << 20;
    struct Slot { std::vector<uint8_t> out; size_t len = 0; uint64_t seq = ~0ull; bool ready = false, bad = false, last = false; };
    struct Blk { size_t off, clen; uint32_t isize; size_t out_off; };
    int fd = -1;
    std::vector<std::thread> workers;
    std::vector<Slot> ring;
    std::mutex mu;
    std::condition_variable cv_ready, cv_free;
    std::vector<uint8_t> ahead; size_t a_len = 0, a_pos = 0; uint64_t file_pos = 0; bool file_end = false;   // shared read-ahead
    uint64_t next_seq = 0, consumed = 0;                 // groups [consumed, consumed + ring.size()) may be in flight
    bool eof = false, stop = false;
    uint64_t held = ~0ull;                               // the group the consumer is reading (its slot is not reused yet)

    // under the lock: make at least `need` bytes available at ahead[a_pos] if the file has them
    void fill(size_t need)
    {
        if (a_len - a_pos >= need || file_end) return;
        if (a_pos) { memmove(ahead.data(), ahead.data() + a_pos, a_len - a_pos); a_len -= a_pos; a_pos = 0; }
        while (a_len < ahead.size()) {
            const ssize_t k = pread(fd, ahead.data() + a_len, ahead.size() - a_len, (off_t)file_pos);
            if (k <= 0) { file_end = true; break; }
            a_len += (size_t)k; file_pos += (uint64_t)k;
        }
    }

    // under the lock: cut the next group into raw/blks; returns false when nothing is left (bad = not BGZF)
    bool cut(std::vector<uint8_t> &raw, std::vector<Blk> &blks, size_t &out_len, bool &bad)
    {
        blks.clear(); out_len = 0; bad = false;
        size_t in_len = 0;
        while (in_len < GROUP_IN && out_len < GROUP_OUT) {
            fill((1u << 16) + 64);
            const uint8_t *cd; size_t cl; uint32_t is;
            const int64_t b = bgzf_block(ahead.data() + a_pos, a_len - a_pos, &cd, &cl, &is);
            if (b < 0) { bad = true; break; }
            if (b == 0) break;                                                 // end of file (a truncated last block ends the stream quietly)
            if (raw.size() < in_len + (size_t)b) raw.resize(in_len + (size_t)b + (1u << 16));
            memcpy(raw.data() + in_len, ahead.data() + a_pos, (size_t)b);
            blks.push_back(Blk{in_len + (size_t)(cd - (ahead.data() + a_pos)), cl, is, out_len});
            out_len += is; in_len += (size_t)b; a_pos += (size_t)b;
        }
        return !blks.empty();
    }

    void worker()
    {
        Inflater inf;
        std::vector<uint8_t> raw(GROUP_IN + (1u << 17));
        std::vector<Blk> blks;
        for (;;) {
            uint64_t seq; size_t out_len = 0; bool last = false, bad = false;
            {
                std::unique_lock<std::mutex> g(mu);
                cv_free.wait(g, [&] { return stop || eof || next_seq < consumed + ring.size(); });
                if (stop || eof) return;
                seq = next_seq++;
                if (!cut(raw, blks, out_len, bad)) { last = true; eof = true; }
            }
            // the slot is free: seq < consumed + ring.size() means its previous tenant (seq - ring.size()) was released
            Slot &s = ring[seq % ring.size()];
            if (s.out.size() < out_len) s.out.resize(out_len);
            // inflate, then the CRC32 that follows every payload (htslib verifies it too)
            if (!bad) for (const Blk &b : blks) if (b.isize && (!inf.run(raw.data() + b.off, b.clen, s.out.data() + b.out_off, b.isize) ||
                                                            (uint32_t)crc32(0L, s.out.data() + b.out_off, b.isize) != le32(raw.data() + b.off + b.clen))) { bad = true; break; }
            {
                std::lock_guard<std::mutex> g(mu);
                s.len = bad ? 0 : out_len; s.seq = seq; s.bad = bad; s.last = last; s.ready = true;
            }
            cv_ready.notify_all();
            if (last) { cv_free.notify_all(); return; }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// BgzfChunker
// ---------------------------------------------------------------------------------------------------------------
BgzfChunker::~BgzfChunker() { if (fd_ >= 0) close(fd_); }

bool BgzfChunker::open(const std::string &path, uint32_t &first_skip)
{
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) { err_ = "cannot open " + path; return false; }
    // inflate leading blocks until the BAM header (magic, text, reference table) is complete; the first record starts at
    // inflated offset q, inside the block that covers it
    Inflater inf;
    std::vector<uint8_t> head, raw((1u << 16) + 64), out(1u << 16);
    std::vector<std::pair<uint64_t, uint64_t>> starts;       // (file offset of the block, inflated offset of its first byte)
    uint64_t fpos = 0;
    for (;;) {
        const ssize_t k = pread(fd_, raw.data(), raw.size(), (off_t)fpos);
        const uint8_t *cd; size_t cl; uint32_t is;
        const int64_t b = k > 0 ? bgzf_block(raw.data(), (size_t)k, &cd, &cl, &is) : 0;
        if (b <= 0) return false;
        if (is) {
            if (!inf.run(cd, cl, out.data(), is)) return false;
            starts.emplace_back(fpos, (uint64_t)head.size());
            head.insert(head.end(), out.begin(), out.begin() + is);
        }
        fpos += (uint64_t)b;
        if (head.size() < 12) continue;
        if (memcmp(head.data(), "BAM\1", 4) != 0) return false;
        size_t q = 8 + (size_t)le32(head.data() + 4);
        if (head.size() < q + 4) continue;
        const uint32_t n_ref = le32(head.data() + q); q += 4;
        bool complete = true;
        for (uint32_t r = 0; r < n_ref; r++) { if (head.size() < q + 4) { complete = false; break; } q += 8 + (size_t)le32(head.data() + q); }
        if (!complete || head.size() < q) continue;
        // the block holding inflated offset q (q == head.size() means: the next block, at fpos)
        pos_ = fpos; first_skip = 0;
        for (size_t i = starts.size(); i-- > 0;) if (starts[i].second <= q && q < head.size()) { pos_ = starts[i].first; first_skip = (uint32_t)(q - starts[i].second); break; }
        return true;
    }
}

bool BgzfChunker::next(size_t max_bytes, std::vector<uint8_t> &bytes, std::vector<Block> &blocks)
{
    blocks.clear();
    if (bytes.size() < max_bytes + (1u << 16) + 64) bytes.resize(max_bytes + (1u << 16) + 64);
    const ssize_t k = pread(fd_, bytes.data(), bytes.size(), (off_t)pos_);
    if (k <= 0) return false;
    size_t p = 0;
    while (p < (size_t)k && p < max_bytes) {
        const uint8_t *cd; size_t cl; uint32_t is;
        const int64_t b = bgzf_block(bytes.data() + p, (size_t)k - p, &cd, &cl, &is);
        if (b < 0) { err_ = "not a BGZF block"; return false; }
        if (b == 0) break;                                   // truncated last block: the stream ends here
        if (is) blocks.push_back(Block{(uint64_t)(cd - bytes.data()), (uint32_t)cl, is});
        p += (size_t)b;
    }
    if (p == 0) return false;
    pos_ += p;
    return true;                                             // (a chunk of empty blocks only -- the EOF marker -- has no blocks)
}

BamStreamScanner::BamStreamScanner() : im_(new Impl) {}

BamStreamScanner::~BamStreamScanner()
{
    { std::lock_guard<std::mutex> g(im_->mu); im_->stop = true; }
    im_->cv_free.notify_all();
    for (auto &t : im_->workers) t.join();
    if (im_->fd >= 0) close(im_->fd);
    delete im_;
}

bool BamStreamScanner::open(const std::string &path, int n_threads)
{
    im_->fd = ::open(path.c_str(), O_RDONLY);
    if (im_->fd < 0) { err_ = "cannot open " + path; return false; }
    // must start with a BGZF block whose payload starts with "BAM\1"
    uint8_t head[1u << 16];
    const ssize_t k = pread(im_->fd, head, sizeof head, 0);
    const uint8_t *cd; size_t cl; uint32_t is;
    if (k < 28 || bgzf_block(head, (size_t)k, &cd, &cl, &is) <= 0 || is < 4) return false;
    std::vector<uint8_t> first(is);
    Inflater inf;
    if (!inf.run(cd, cl, first.data(), is) || memcmp(first.data(), "BAM\1", 4) != 0) return false;
    const int nt = std::max(1, n_threads);
    im_->ring.resize((size_t)nt * 2 + 2);
    im_->ahead.resize(Impl::AHEAD);
    for (int i = 0; i < nt; i++) im_->workers.emplace_back([this] { im_->worker(); });
    return true;
}

bool BamStreamScanner::next_group()
{
    Impl &m = *im_;
    std::unique_lock<std::mutex> g(m.mu);
    if (m.held != ~0ull) { m.held = ~0ull; m.consumed++; m.cv_free.notify_one(); }      // release the group just finished: one slot, one worker
    const uint64_t want = m.consumed;
    Impl::Slot &s = m.ring[want % m.ring.size()];
    m.cv_ready.wait(g, [&] { return s.ready && s.seq == want; });
    s.ready = false;
    if (s.bad) { err_ = "BGZF stream is corrupt"; return false; }
    m.held = want;
    cur_ = s.out.data(); cur_len_ = s.len; cur_pos_ = 0;
    if (s.last && s.len == 0) return false;
    return true;
}

const uint8_t *BamStreamScanner::take(size_t n)
{
    if (cur_len_ - cur_pos_ >= n) { const uint8_t *p = cur_ + cur_pos_; cur_pos_ += n; return p; }
    // straddles groups: assemble in carry_
    carry_.clear();
    while (carry_.size() < n) {
        const size_t have = cur_len_ - cur_pos_, need = n - carry_.size(), k = have < need ? have : need;
        carry_.insert(carry_.end(), cur_ + cur_pos_, cur_ + cur_pos_ + k);
        cur_pos_ += k;
        if (carry_.size() < n && !next_group()) return nullptr;
    }
    return carry_.data();
}

bool BamStreamScanner::skip(size_t n)
{
    while (n) {
        const size_t have = cur_len_ - cur_pos_, k = have < n ? have : n;
        cur_pos_ += k; n -= k;
        if (n && !next_group()) return false;
    }
    return true;
}

bool BamStreamScanner::read_header()
{
    const uint8_t *p = take(8);
    if (!p || memcmp(p, "BAM\1", 4) != 0) { if (err_.empty()) err_ = "not a BAM stream"; return false; }
    const uint32_t l_text = le32(p + 4);
    if (!skip(l_text)) return false;
    p = take(4);
    if (!p) return false;
    const uint32_t n_ref = le32(p);
    for (uint32_t i = 0; i < n_ref; i++) {
        p = take(4);
        if (!p) return false;
        const uint32_t l_name = le32(p);
        if (!skip((size_t)l_name + 4)) return false;
    }
    return true;
}

}  // namespace pdh
