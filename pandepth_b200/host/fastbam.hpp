// fastbam.hpp -- range scanner for coordinate-sorted BAM + .bai that touches only what the depth path needs.
//
// The reference (and our fallback path) goes through htslib's iterator: BGZF block cache, bam1_t copies of the whole record
// (name, SEQ, QUAL, aux), per-record function calls -- 50-470 ns per read per core.  The depth path
// needs four fields (refID, pos, flag/mapq, CIGAR words), so this scanner
//   * takes the start offset of a genomic range from the .bai LINEAR index (first read overlapping the 16 kb window;
//     in a sorted file every read with pos >= beg follows it),
//   * preads compressed bytes, inflates BGZF blocks itself (libdeflate), and
//   * walks the raw records in the inflated stream, handing (pos, flag, mapq, n_cigar, cigar*) to a callback until
//     pos >= end or the contig changes.
// Role in the reference: the `sam_itr_next` loop of ProDealChrBambai, PanDepth.cpp:726-734 (decode side only).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pdh {

struct BaiLinear {
    struct Ref { std::vector<uint64_t> ioff; uint64_t data_beg = 0, data_end = 0; };
    std::vector<Ref> refs;
    bool load(const std::string &bai_path);
    // virtual offset to start scanning for reads overlapping position `beg` (0-based) of `tid`; 0 = nothing to read
    uint64_t start_voffset(int tid, int64_t beg) const;
};

struct RawRead {
    int32_t pos; uint16_t flag; uint8_t mapq; uint32_t n_cigar;
    const uint8_t *cigar;              // n_cigar little-endian words, NOT necessarily 4-byte aligned; valid until the callback returns
    int32_t l_seq;
    uint32_t word(uint32_t i) const { const uint8_t *p = cigar + 4 * (size_t)i; return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
};

class BamRangeScanner {
public:
    BamRangeScanner();
    ~BamRangeScanner();
    bool open(const std::string &bam_path);
    // Calls fn(const RawRead&) for every record of contig `tid` from virtual offset `voff` on, in file order, until a record
    // with pos >= end_pos or another refID is met.  Returns false on a malformed file / IO error (see error()).
    template <typename F> bool scan(uint64_t voff, int tid, int64_t end_pos, F &&fn);
    const std::string &error() const { return err_; }

private:
    // make at least `need` bytes available at buf_[bpos_]; false at EOF/error (the common case -- already there -- is inline)
    bool fill(size_t need) { return blen_ - bpos_ >= need ? true : fill_slow(need); }
    bool fill_slow(size_t need);
    bool inflate_next_block();
    int fd_ = -1;
    void *dec_ = nullptr;
    std::vector<uint8_t> raw_; size_t raw_len_ = 0, raw_pos_ = 0; uint64_t file_pos_ = 0; bool eof_ = false;
    size_t read_ahead_ = 128u << 10;   // bytes per pread; reset by scan(), doubled as the scan goes on
    std::vector<uint8_t> buf_; size_t blen_ = 0, bpos_ = 0;
    std::vector<uint32_t> cg_;         // long CIGAR recovered from the CG tag
    std::string err_;
};

// The whole-file counterpart for BAMs without an index (the reference's "No Index mode", PanDepth.cpp:4604-4713, reads them
// through `sam_read1` on one thread): BGZF blocks are independent deflate streams, so a pool of workers inflates groups of
// consecutive blocks in parallel, and ONE consumer walks the records of the re-assembled stream in file order -- the order
// matters (batches are cut at contig changes and sorted input keeps the tile path).  The walk costs ~7 ns per record, the
// inflate 23-860 ns, so the consumer is not the bottleneck until ~8 inflating cores on SEQ-less files.
class BamStreamScanner {
public:
    struct Read : RawRead { int32_t tid; };
    BamStreamScanner();
    ~BamStreamScanner();
    // false if the file cannot be opened or does not start with a BGZF block holding the BAM magic (caller falls back to htslib)
    bool open(const std::string &bam_path, int n_threads);
    // Calls fn(const Read&) -> bool for every record of the file in order until it returns false.  Returns false on a malformed stream (see error());
    // a truncated final record ends the scan quietly, as `while (sam_read1(...) >= 0)` does.
    template <typename F> bool scan_all(F &&fn);
    const std::string &error() const { return err_; }

private:
    struct Impl;
    Impl *im_;
    std::string err_;
    std::vector<uint8_t> carry_;        // bytes of a record (or header field) that straddles two groups
    const uint8_t *cur_ = nullptr; size_t cur_len_ = 0, cur_pos_ = 0;
    std::vector<uint32_t> cg_;
    bool next_group();                                   // advance cur_ to the next inflated group; false at EOF / error
    const uint8_t *take(size_t n);                       // n contiguous bytes of the stream (stitched through carry_), or nullptr at EOF
    bool skip(size_t n);
    bool read_header();
};

// Sequential reader of whole BGZF blocks for callers that hand the compressed bytes on (pd_push_bgzf): finds the first BAM
// record (the header may span blocks) and then returns runs of consecutive blocks with their payload offsets.
class BgzfChunker {
public:
    struct Block { uint64_t in_off; uint32_t in_len; uint32_t out_len; };        // payload offset inside the chunk's bytes
    ~BgzfChunker();
    // false if the file is not a BAM in BGZF blocks; first_skip = offset of the first record inside the first block returned
    bool open(const std::string &bam_path, uint32_t &first_skip);
    // next chunk of up to max_bytes compressed bytes (whole blocks, at least one); false at the end of the file
    bool next(size_t max_bytes, std::vector<uint8_t> &bytes, std::vector<Block> &blocks);
    const std::string &error() const { return err_; }
    uint64_t position() const { return pos_; }          // file offset of the next block next() would return
private:
    int fd_ = -1; uint64_t pos_ = 0; std::string err_;
};

// BGZF block at h (avail bytes): > 0 total size (payload in cdata / clen, inflated size in isize); 0 = more bytes needed
// (header or block not complete within avail); -1 = not a BGZF block (bad magic, size fields, or ISIZE > 64 KB)
int64_t bgzf_block_at(const uint8_t *h, size_t avail, const uint8_t **cdata, size_t *clen, uint32_t *isize);

// long CIGAR (> 65535 ops) kept in the CG:B,I tag behind a `<l_seq>S<ref_len>N` placeholder (SAMv1 4.2.2): fills r from it
bool resolve_long_cigar_tag(const uint8_t *rec, uint32_t block_size, uint32_t l_read_name, uint32_t n_cigar, int32_t l_seq,
                            std::vector<uint32_t> &store, RawRead &r);

inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint16_t le16(const uint8_t *p) { return (uint16_t)((uint16_t)p[0] | (uint16_t)p[1] << 8); }

template <typename F>
bool BamRangeScanner::scan(uint64_t voff, int tid, int64_t end_pos, F &&fn)
{
    // position the compressed stream at the block, the inflated stream at the in-block offset
    file_pos_ = voff >> 16; raw_len_ = raw_pos_ = 0; eof_ = false; read_ahead_ = 128u << 10;
    blen_ = bpos_ = 0;
    if (!inflate_next_block()) return err_.empty();
    bpos_ = (size_t)(voff & 0xffff);
    if (bpos_ > blen_) { err_ = "virtual offset beyond its BGZF block"; return false; }
    for (;;) {
        if (!fill(4)) return err_.empty();
        const uint32_t block_size = le32(&buf_[bpos_]);
        if (block_size < 32 || block_size > (1u << 30)) { err_ = "implausible BAM record size"; return false; }
        if (!fill(4 + (size_t)block_size)) { if (err_.empty()) err_ = "truncated BAM record"; return false; }
        const uint8_t *rec = &buf_[bpos_ + 4];
        const int32_t ref = (int32_t)le32(rec);
        const int32_t pos = (int32_t)le32(rec + 4);
        if (ref != tid || (int64_t)pos >= end_pos) return true;
        const uint32_t l_read_name = rec[8];
        const uint32_t n_cigar = le16(rec + 12);
        RawRead r;
        r.pos = pos; r.mapq = rec[9]; r.flag = le16(rec + 14); r.l_seq = (int32_t)le32(rec + 16);
        r.n_cigar = n_cigar;
        if (32 + l_read_name + 4 * (size_t)n_cigar > block_size) { err_ = "BAM record shorter than its CIGAR"; return false; }
        r.cigar = rec + 32 + l_read_name;
        if (n_cigar == 2) resolve_long_cigar_tag(rec, block_size, l_read_name, n_cigar, r.l_seq, cg_, r);
        fn(r);
        bpos_ += 4 + (size_t)block_size;
    }
}

template <typename F>
bool BamStreamScanner::scan_all(F &&fn)
{
    if (!read_header()) return err_.empty();
    for (;;) {
        const uint8_t *h = take(4);
        if (!h) return err_.empty();
        const uint32_t block_size = le32(h);
        if (block_size < 32 || block_size > (1u << 30)) { err_ = "implausible BAM record size"; return false; }
        const uint8_t *rec = take(block_size);
        if (!rec) return err_.empty();                                        // truncated final record: stop quietly
        const uint32_t l_read_name = rec[8];
        const uint32_t n_cigar = le16(rec + 12);
        Read r;
        r.tid = (int32_t)le32(rec); r.pos = (int32_t)le32(rec + 4);
        r.mapq = rec[9]; r.flag = le16(rec + 14); r.l_seq = (int32_t)le32(rec + 16);
        r.n_cigar = n_cigar;
        if (32 + l_read_name + 4 * (size_t)n_cigar > block_size) { err_ = "BAM record shorter than its CIGAR"; return false; }
        r.cigar = rec + 32 + l_read_name;
        if (n_cigar == 2) resolve_long_cigar_tag(rec, block_size, l_read_name, n_cigar, r.l_seq, cg_, r);
        if (!fn(r)) return true;
    }
}

}  // namespace pdh
