/*
 * bamgen.cpp -- deterministic synthetic BAM generator (TEST / BENCH INFRASTRUCTURE ONLY).
 *
 * Builds against the htslib the reference vendors (headers under /root/reference/include,
 * lib/libhts.a), so it lives under oracle/ and its binary goes to oracle/_ref/bamgen.
 * Shapes follow SURVEY.md section 8(d): coordinate-sorted, @HD SO:coordinate, .bai via sam_index_build3,
 * BGZF level 1, starts uniform per contig, std::mt19937_64 seeded.
 *
 *   bamgen -o out.bam -c name:len[,name:len...] -n reads [-s seed] [-l readlen] [-m uniform|mixed]
 *   bamgen -I in.sam -o out.bam      (convert a SAM fixture to BAM + .bai)
 *          [-q] (emit random SEQ+QUAL, flavour B; default SEQ='*', flavour A) [-t threads] [-S] (write SAM text)
 *   an output name ending in .cram is written as CRAM 3.x without a reference (no_ref) + .crai
 *          [-X] build a .csi index (min_shift 14) instead of the .bai
 *          [-O] mixed mode: no read may run past its contig's end (the reference's multi-file path has no slack behind its
 *               arrays: overhanging reads corrupt its heap)
 *          [-P] uniform mode, parallel writer for bench-sized files (10^8 .. 4x10^9 reads): the genome is cut into slices, -t
 *               threads generate, sort and BGZF-compress (libdeflate level 1, whole records per block like htslib's writer)
 *               their slices independently, the slices are appended in order and the .bai is written directly from the
 *               virtual offsets known while writing -- no second pass over the file.  Same record layout as the serial
 *               mode; the random stream differs (one generator per slice).
 *   bamgen -R file.bam        rebuild file.bam.bai with htslib (to cross-check the -P index)
 *
 * mode uniform : every read flag 0, mapq 60, CIGAR <readlen>M.
 * mode mixed   : random CIGARs over MIDNSHP=X, random flags (incl. UNMAP/SECONDARY/QCFAIL/DUP/SUPP bits),
 *                random mapq -- for differential testing against the reference binary.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <random>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include "sam.h"
#include "hts.h"

// libdeflate's public C API (the reference vendors lib/libdeflate.a without its header)
extern "C" {
struct libdeflate_compressor;
struct libdeflate_compressor *libdeflate_alloc_compressor(int level);
size_t libdeflate_deflate_compress(struct libdeflate_compressor *c, const void *in, size_t in_nbytes, void *out, size_t out_nbytes_avail);
void libdeflate_free_compressor(struct libdeflate_compressor *c);
uint32_t libdeflate_crc32(uint32_t crc, const void *buffer, size_t len);
}

struct Contig { std::string name; int64_t len; };

static void die(const char *m) { fprintf(stderr, "bamgen: %s\n", m); exit(2); }

static bool is_cram_name(const std::string &p) { return p.size() > 5 && p.compare(p.size() - 5, 5, ".cram") == 0; }

static htsFile *open_out(const std::string &path, bool samtext)
{
    const bool cram = is_cram_name(path);
    htsFile *fp = hts_open(path.c_str(), samtext ? "w" : (cram ? "wc" : "wb1"));
    if (fp && cram && hts_set_opt(fp, CRAM_OPT_NO_REF, 1) != 0) die("cannot set no_ref");
    return fp;
}

static void make_mixed_cigar(std::mt19937_64 &rng, int readlen, std::vector<uint32_t> &cig, int &qlen, int &rlen)
{
    // a few ops; query-consuming ops sum to ~readlen
    cig.clear(); qlen = 0; rlen = 0;
    auto push = [&](int op, int n) { if (n > 0) cig.push_back((uint32_t)n << 4 | (uint32_t)op); };
    int r = rng() % 100;
    if (r < 8) { push(BAM_CHARD_CLIP, 1 + rng() % 30); }
    if (r % 5 == 0) { int n = 1 + rng() % 40; push(BAM_CSOFT_CLIP, n); qlen += n; }
    int nblocks = 1 + rng() % 4;
    for (int b = 0; b < nblocks; b++) {
        int n = 1 + rng() % (readlen / nblocks + 1);
        int mop = (rng() % 10 == 0) ? BAM_CEQUAL : (rng() % 10 == 1 ? BAM_CDIFF : BAM_CMATCH);
        push(mop, n); qlen += n; rlen += n;
        if (b + 1 < nblocks) {
            int k = rng() % 6;
            if (k == 0) { int d = 1 + rng() % 20; push(BAM_CDEL, d); rlen += d; }
            else if (k == 1) { int d = 1 + rng() % 400; push(BAM_CREF_SKIP, d); rlen += d; }
            else if (k == 2) { int d = 1 + rng() % 10; push(BAM_CINS, d); qlen += d; }
            else if (k == 3) { push(BAM_CPAD, 1 + rng() % 3); }
        }
    }
    if (r % 7 == 0) { int n = 1 + rng() % 40; push(BAM_CSOFT_CLIP, n); qlen += n; }
}

// ---------------------------------------------------------------------------------------------------------------
// -P: parallel uniform generator with a directly written .bai (SAMv1 section 5.2)
// ---------------------------------------------------------------------------------------------------------------
namespace par {

struct Rng {                                                     // xorshift64* (one per slice)
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull) { if (!s) s = 1; next(); next(); }
    uint64_t next() { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return s * 0x2545F4914F6CDD1Dull; }
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
};

static void put32(std::vector<uint8_t> &v, uint32_t x) { for (int k = 0; k < 4; k++) v.push_back((uint8_t)(x >> (8 * k))); }

// one BGZF block (<= 65280 payload bytes)
static void bgzf_block(libdeflate_compressor *c, const uint8_t *p, size_t n, std::vector<uint8_t> &out)
{
    const size_t at = out.size();
    out.resize(at + 18 + 65536 + 1024 + 8);
    uint8_t *h = out.data() + at;
    static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};
    memcpy(h, hdr, 16);
    size_t k = n ? libdeflate_deflate_compress(c, p, n, h + 18, 65536 + 1024) : 0;
    if (n == 0) { h[18] = 3; h[19] = 0; k = 2; }
    if (k == 0 || k + 26 > 65536) die("block does not compress into a BGZF block");
    const uint32_t bsize = (uint32_t)(k + 25);
    h[16] = (uint8_t)bsize; h[17] = (uint8_t)(bsize >> 8);
    const uint32_t crc = libdeflate_crc32(0, p, n), isz = (uint32_t)n;
    for (int j = 0; j < 4; j++) { h[18 + k + j] = (uint8_t)(crc >> (8 * j)); h[22 + k + j] = (uint8_t)(isz >> (8 * j)); }
    out.resize(at + 18 + k + 8);
}

struct Slice {
    int ci; int64_t lo, hi, n;                                   // positions uniform in [lo, hi), n reads
    std::vector<uint8_t> bytes;                                  // finished BGZF blocks
    std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> bins;   // bin -> (first begin, last end), offsets relative to the slice
    std::vector<std::pair<int64_t, uint64_t>> lin;               // (16 kb window, smallest relative voffset of a read overlapping it)
    uint64_t first_v = 0, last_v = 0;
    bool done = false;
};

static int run(const std::string &out, const std::vector<Contig> &contigs, int64_t nreads, uint64_t seed, int readlen, bool withseq, int threads, const std::string &text)
{
    int64_t G = 0; for (auto &c : contigs) G += c.len;
    const int64_t slice_reads = withseq ? 200000 : 2000000;
    std::vector<Slice> slices;
    int64_t assigned = 0;
    for (size_t ci = 0; ci < contigs.size(); ci++) {
        const int64_t len = contigs[ci].len;
        const int64_t n = (ci + 1 == contigs.size()) ? nreads - assigned : (int64_t)((long double)nreads * len / G);
        assigned += n;
        if (len < 1 || n <= 0) continue;
        const int64_t maxstart = std::max<int64_t>(0, len - readlen) + 1;                 // starts in [0, maxstart)
        const int64_t k = std::max<int64_t>(1, (n + slice_reads - 1) / slice_reads);
        int64_t given = 0;
        for (int64_t j = 0; j < k; j++) {
            Slice s; s.ci = (int)ci; s.lo = maxstart * j / k; s.hi = maxstart * (j + 1) / k;
            s.n = (j + 1 == k) ? n - given : n / k; given += s.n;
            if (s.hi > s.lo && s.n > 0) slices.push_back(std::move(s));
        }
    }
    FILE *fo = fopen(out.c_str(), "wb");
    if (!fo) die("cannot open output");
    // BAM header in its own block(s)
    uint64_t file_pos = 0;
    {
        std::vector<uint8_t> h; h.insert(h.end(), {'B', 'A', 'M', 1});
        put32(h, (uint32_t)text.size()); h.insert(h.end(), text.begin(), text.end());
        put32(h, (uint32_t)contigs.size());
        for (auto &c : contigs) { put32(h, (uint32_t)c.name.size() + 1); h.insert(h.end(), c.name.begin(), c.name.end()); h.push_back(0); put32(h, (uint32_t)c.len); }
        libdeflate_compressor *c = libdeflate_alloc_compressor(1);
        std::vector<uint8_t> z;
        for (size_t p = 0; p < h.size(); p += 65280) bgzf_block(c, h.data() + p, std::min<size_t>(65280, h.size() - p), z);
        libdeflate_free_compressor(c);
        if (fwrite(z.data(), 1, z.size(), fo) != z.size()) die("write failed");
        file_pos = z.size();
    }
    std::mutex mu; std::condition_variable cv_done, cv_room;
    std::atomic<size_t> next{0};
    size_t written = 0;
    const size_t window = (size_t)std::max(2, threads) * 2;       // slices that may be finished but not yet written
    auto worker = [&] {
        libdeflate_compressor *c = libdeflate_alloc_compressor(1);
        std::vector<int32_t> starts; std::vector<uint8_t> raw;
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= slices.size()) break;
            { std::unique_lock<std::mutex> g(mu); cv_room.wait(g, [&] { return k < written + window; }); }
            Slice &s = slices[k];
            Rng rng(seed * 1000003ull + (uint64_t)k);
            starts.resize((size_t)s.n);
            for (auto &x : starts) x = (int32_t)(s.lo + (int64_t)rng.below((uint64_t)(s.hi - s.lo)));
            std::sort(starts.begin(), starts.end());
            const int l_seq = withseq ? readlen : 0;
            const size_t rec = 4 + 32 + 4 + 4 + (size_t)(l_seq + 1) / 2 + (size_t)l_seq;
            raw.clear(); raw.reserve(65280 + rec);
            uint64_t cur_block = 0;                               // relative compressed offset of the block being filled
            std::unordered_map<uint32_t, std::pair<uint64_t, uint64_t>> bins;
            int64_t last_win = -1;
            auto flush = [&]() { if (raw.empty()) return; bgzf_block(c, raw.data(), raw.size(), s.bytes); raw.clear(); cur_block = s.bytes.size(); };
            for (int64_t i = 0; i < s.n; i++) {
                if (raw.size() + rec > 65280) flush();            // whole records per block, like bam_write1's bgzf_flush_try
                const uint64_t v0 = cur_block << 16 | raw.size();
                const int32_t pos = starts[(size_t)i], end = pos + readlen;
                const uint32_t bin = (uint32_t)hts_reg2bin(pos, end, 14, 5);
                const size_t at = raw.size(); raw.resize(at + rec);
                uint8_t *r = raw.data() + at;
                auto w32 = [&](size_t o, uint32_t x) { memcpy(r + o, &x, 4); };
                w32(0, (uint32_t)(rec - 4)); w32(4, (uint32_t)s.ci); w32(8, (uint32_t)pos);
                w32(12, 4u | 60u << 8 | bin << 16);                // l_read_name 4, mapq 60, bin
                w32(16, 1u | 0u << 16);                           // n_cigar_op 1, flag 0
                w32(20, (uint32_t)l_seq); w32(24, 0xffffffffu); w32(28, 0xffffffffu); w32(32, 0u);
                r[36] = 'r'; r[37] = r[38] = r[39] = 0;
                w32(40, (uint32_t)readlen << 4 | BAM_CMATCH);
                uint8_t *sq = r + 44;
                for (int q = 0; q < (l_seq + 1) / 2; q += 8) {    // 16 bases per random word
                    uint64_t x = rng.next();
                    for (int b = 0; b < 8 && q + b < (l_seq + 1) / 2; b++, x >>= 4) sq[q + b] = (uint8_t)((1u << (x & 3)) << 4 | (1u << ((x >> 2) & 3)));
                }
                uint8_t *ql = sq + (l_seq + 1) / 2;
                for (int q = 0; q < l_seq; q += 4) {              // 4 qualities per random word, 2..40
                    uint64_t x = rng.next();
                    for (int b = 0; b < 4 && q + b < l_seq; b++, x >>= 16) ql[q + b] = (uint8_t)(2 + (((x & 0xffff) * 39) >> 16));
                }
                const uint64_t v1 = cur_block << 16 | raw.size();
                auto it = bins.find(bin);
                if (it == bins.end()) bins.emplace(bin, std::make_pair(v0, v1)); else it->second.second = v1;
                for (int64_t w = std::max<int64_t>(last_win + 1, pos >> 14); w <= (end - 1) >> 14; w++) { s.lin.emplace_back(w, v0); last_win = w; }
                if (i == 0) s.first_v = v0;
                s.last_v = v1;
            }
            flush();
            s.bins.assign(bins.begin(), bins.end());
            { std::lock_guard<std::mutex> g(mu); s.done = true; }
            cv_done.notify_all();
        }
        libdeflate_free_compressor(c);
    };
    std::vector<std::thread> th;
    for (int i = 0; i < std::max(1, threads); i++) th.emplace_back(worker);
    // writer: slices in order; the index is merged as they pass
    struct RefIdx { std::unordered_map<uint32_t, std::pair<uint64_t, uint64_t>> bins; std::vector<uint64_t> lin; uint64_t beg = 0, end = 0, n = 0; bool any = false; };
    std::vector<RefIdx> idx(contigs.size());
    for (size_t k = 0; k < slices.size(); k++) {
        Slice &s = slices[k];
        { std::unique_lock<std::mutex> g(mu); cv_done.wait(g, [&] { return s.done; }); }
        if (fwrite(s.bytes.data(), 1, s.bytes.size(), fo) != s.bytes.size()) die("write failed");
        const uint64_t base = file_pos << 16;
        RefIdx &r = idx[(size_t)s.ci];
        for (auto &b : s.bins) {
            auto it = r.bins.find(b.first);
            if (it == r.bins.end()) r.bins.emplace(b.first, std::make_pair(b.second.first + base, b.second.second + base));
            else it->second.second = b.second.second + base;
        }
        for (auto &l : s.lin) {
            if ((size_t)l.first >= r.lin.size()) r.lin.resize((size_t)l.first + 1, 0);
            if (r.lin[(size_t)l.first] == 0) r.lin[(size_t)l.first] = l.second + base;
        }
        if (!r.any) { r.beg = s.first_v + base; r.any = true; }
        r.end = s.last_v + base; r.n += (uint64_t)s.n;
        file_pos += s.bytes.size();
        std::vector<uint8_t>().swap(s.bytes); std::vector<std::pair<int64_t, uint64_t>>().swap(s.lin); s.bins.clear(); s.bins.shrink_to_fit();
        { std::lock_guard<std::mutex> g(mu); written = k + 1; }
        cv_room.notify_all();
    }
    for (auto &t : th) t.join();
    static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (fwrite(eof, 1, 28, fo) != 28 || fclose(fo) != 0) die("write failed");
    // .bai
    FILE *fi = fopen((out + ".bai").c_str(), "wb");
    if (!fi) die("cannot write index");
    auto w32 = [&](uint32_t x) { fwrite(&x, 4, 1, fi); };
    auto w64 = [&](uint64_t x) { fwrite(&x, 8, 1, fi); };
    fwrite("BAI\1", 1, 4, fi); w32((uint32_t)contigs.size());
    for (auto &r : idx) {
        if (!r.any) { w32(0); w32(0); continue; }
        w32((uint32_t)r.bins.size() + 1);
        std::map<uint32_t, std::pair<uint64_t, uint64_t>> ordered(r.bins.begin(), r.bins.end());
        for (auto &b : ordered) { w32(b.first); w32(1); w64(b.second.first); w64(b.second.second); }
        w32(37450); w32(2); w64(r.beg); w64(r.end); w64(r.n); w64(0);
        for (size_t w = r.lin.size(); w-- > 1;) if (r.lin[w - 1] == 0) r.lin[w - 1] = r.lin[w];   // empty windows take the next window's offset
        w32((uint32_t)r.lin.size());
        for (uint64_t v : r.lin) w64(v);
    }
    w64(0);
    if (fclose(fi) != 0) die("index write failed");
    return 0;
}

}  // namespace par

int main(int argc, char **argv)
{
    std::string out, cspec, mode = "uniform", insam;
    int64_t nreads = 0; uint64_t seed = 12345; int readlen = 150; bool withseq = false; int threads = 4; bool samtext = false; int min_shift = 0; bool no_overhang = false;
    bool parallel = false; std::string reindex;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&]() -> const char * { if (i + 1 >= argc) die("missing argument"); return argv[++i]; };
        if (a == "-o") out = need();
        else if (a == "-c") cspec = need();
        else if (a == "-n") nreads = atoll(need());
        else if (a == "-s") seed = strtoull(need(), 0, 10);
        else if (a == "-l") readlen = atoi(need());
        else if (a == "-m") mode = need();
        else if (a == "-q") withseq = true;
        else if (a == "-t") threads = atoi(need());
        else if (a == "-S") samtext = true;
        else if (a == "-I") insam = need();
        else if (a == "-X") min_shift = 14;
        else if (a == "-O") no_overhang = true;
        else if (a == "-P") parallel = true;
        else if (a == "-R") reindex = need();
        else die("unknown option");
    }
    if (!reindex.empty()) { if (sam_index_build3(reindex.c_str(), NULL, 0, threads) < 0) die("index build failed"); return 0; }
    if (!insam.empty()) {                                   // SAM text -> BAM (+ .bai): fixture conversion
        if (out.empty()) die("-I needs -o");
        htsFile *in = hts_open(insam.c_str(), "r");
        if (!in) die("cannot open input SAM");
        sam_hdr_t *h = sam_hdr_read(in);
        if (!h) die("cannot read SAM header");
        htsFile *o = open_out(out, false);
        if (!o || sam_hdr_write(o, h) < 0) die("cannot write BAM");
        bam1_t *r = bam_init1();
        while (sam_read1(in, h, r) >= 0) if (sam_write1(o, h, r) < 0) die("write failed");
        bam_destroy1(r); hts_close(in);
        if (hts_close(o) < 0) die("close failed");
        sam_hdr_destroy(h);
        if (sam_index_build3(out.c_str(), NULL, min_shift, 1) < 0) die("index build failed");
        return 0;
    }
    if (out.empty() || cspec.empty() || nreads <= 0) die("usage: bamgen -o out.bam -c name:len,... -n reads [-s seed] [-l len] [-m uniform|mixed] [-q] [-t thr] [-S]");
    std::vector<Contig> contigs;
    {
        size_t p = 0;
        while (p < cspec.size()) {
            size_t q = cspec.find(',', p); if (q == std::string::npos) q = cspec.size();
            std::string tok = cspec.substr(p, q - p);
            size_t c = tok.rfind(':'); if (c == std::string::npos) die("bad contig spec");
            contigs.push_back({tok.substr(0, c), atoll(tok.c_str() + c + 1)});
            p = q + 1;
        }
    }
    int64_t G = 0; for (auto &c : contigs) G += c.len;
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (auto &c : contigs) text += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.len) + "\n";
    if (parallel) {
        if (mode != "uniform" || samtext || is_cram_name(out) || min_shift) die("-P writes uniform BAM + .bai only");
        return par::run(out, contigs, nreads, seed, readlen, withseq, threads, text);
    }
    sam_hdr_t *hdr = sam_hdr_parse(text.size(), text.c_str());
    if (!hdr) die("header parse failed");

    htsFile *fp = open_out(out, samtext);
    if (!fp) die("cannot open output");
    if (threads > 1 && !samtext) hts_set_threads(fp, threads);
    if (sam_hdr_write(fp, hdr) < 0) die("header write failed");

    std::mt19937_64 rng(seed);
    bam1_t *b = bam_init1();
    const bool mixed = (mode == "mixed");
    std::vector<uint32_t> cig;
    int64_t assigned = 0;
    for (size_t ci = 0; ci < contigs.size(); ci++) {
        int64_t len = contigs[ci].len;
        int64_t n = (ci + 1 == contigs.size()) ? nreads - assigned : (int64_t)((long double)nreads * len / G);
        assigned += n;
        if (len < 1) continue;
        int64_t maxstart = len - readlen; if (maxstart < 0) maxstart = 0;
        std::vector<int32_t> starts((size_t)n);
        for (auto &s : starts) s = (int32_t)(rng() % (uint64_t)(maxstart + 1));
        std::sort(starts.begin(), starts.end());
        for (int64_t i = 0; i < n; i++) {
            int qlen = readlen, rlen = readlen; uint16_t flag = 0; uint8_t mapq = 60;
            if (mixed) {
                make_mixed_cigar(rng, readlen, cig, qlen, rlen);
                if ((int64_t)starts[i] + rlen > len + (no_overhang ? 0 : 400)) {            // keep overhang inside the reference's +500 slack (:4132)
                    int m = (int)std::min<int64_t>(readlen, std::max<int64_t>(1, len - starts[i]));
                    cig.assign(1, (uint32_t)m << 4 | BAM_CMATCH); qlen = rlen = m;
                }
                uint64_t r = rng();
                static const uint16_t bits[] = {0x1, 0x2, 0x10, 0x20, 0x40, 0x80, 0x100, 0x200, 0x400, 0x800, 0x4};
                for (int k = 0; k < 11; k++) if (((r >> (k * 4)) & 0xf) == 0) flag |= bits[k];
                mapq = (uint8_t)((r >> 48) % 61);
            } else {
                cig.assign(1, (uint32_t)readlen << 4 | BAM_CMATCH);
            }
            int l_seq = withseq ? qlen : 0;
            const char qn[] = "r"; int l_qname = 4;                  // "r\0\0\0" padded to 4
            size_t dl = l_qname + cig.size() * 4 + (size_t)(l_seq + 1) / 2 + l_seq;
            if (b->m_data < dl) { b->m_data = (uint32_t)(dl + 64); b->data = (uint8_t *)realloc(b->data, b->m_data); }
            b->l_data = (int)dl;
            memset(b->data, 0, l_qname); memcpy(b->data, qn, 1);
            memcpy(b->data + l_qname, cig.data(), cig.size() * 4);
            uint8_t *sq = b->data + l_qname + cig.size() * 4;
            for (int k = 0; k < (l_seq + 1) / 2; k++) { static const uint8_t nt[4] = {1, 2, 4, 8}; uint64_t r = rng(); sq[k] = (uint8_t)(nt[r & 3] << 4 | nt[(r >> 2) & 3]); }
            uint8_t *ql = sq + (l_seq + 1) / 2;
            for (int k = 0; k < l_seq; k++) ql[k] = (uint8_t)(2 + (rng() % 39));
            b->core.tid = (int32_t)ci; b->core.pos = starts[i]; b->core.qual = mapq; b->core.flag = flag;
            b->core.l_qname = (uint16_t)l_qname; b->core.l_extranul = 2; b->core.n_cigar = (uint32_t)cig.size();
            b->core.l_qseq = l_seq; b->core.mtid = -1; b->core.mpos = -1; b->core.isize = 0;
            b->core.bin = (uint16_t)hts_reg2bin(starts[i], starts[i] + (rlen > 0 ? rlen : 1), 14, 5);
            if (sam_write1(fp, hdr, b) < 0) die("write failed");
        }
    }
    bam_destroy1(b);
    if (hts_close(fp) < 0) die("close failed");
    sam_hdr_destroy(hdr);
    if (!samtext) { if (sam_index_build3(out.c_str(), NULL, min_shift, threads) < 0) die("index build failed"); }
    return 0;
}
