// main.cpp -- `pandepth` on a B200: the reference's command line and output files, with the depth path on the GPU.
//
//   host threads (htslib decode)  --(tid, pos, cigar batches, pinned memory)-->  pd_push_contig   [role of ProDealChrBambai*]
//   pd_finalize / pd_reduce / pd_reduce_windows / pd_depth_read                                   [role of StatChrDepth*, mode-6 loop, -a]
//   report writer (same doubles, same text)                                                       [PanDepth.cpp:4879-5127]
//
// Decode parallelism is by genomic range through the BAI/CSI index (any number of threads on one contig), not by contig
// as in the reference (:4415-4523).
#include "pd_host.hpp"
#include "fastbam.hpp"
#include "fastfmt.hpp"
#include "../../include/pandepth_b200.h"

#include <zlib.h>
#include "sam.h"
#include "hts.h"

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <future>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>

using namespace pdh;

namespace {

#ifdef PD_HAVE_LIBDEFLATE
// libdeflate's public C API (the reference vendors lib/libdeflate.a without its header)
extern "C" {
struct libdeflate_compressor;
struct libdeflate_compressor *libdeflate_alloc_compressor(int compression_level);
size_t libdeflate_gzip_compress(struct libdeflate_compressor *c, const void *in, size_t in_nbytes, void *out, size_t out_nbytes_avail);
size_t libdeflate_gzip_compress_bound(struct libdeflate_compressor *c, size_t in_nbytes);
void libdeflate_free_compressor(struct libdeflate_compressor *c);
}
#endif

// gzip member of one text chunk (RFC 1952; concatenated members are one valid .gz stream -- parity is on decompressed bytes).
// libdeflate level 5 by default: on window rows it compresses slightly better than the zlib level 6 the reference's
// gzstream uses (ratio 7.0 vs 6.9) at a quarter of the CPU time (520 vs 1900-2000 ns per row); PANDEPTH_GZ_LEVEL=1..12
// overrides, PANDEPTH_GZ_ZLIB=1 forces the zlib path.
std::string gz_member(const std::string &in)
{
#ifdef PD_HAVE_LIBDEFLATE
    static const bool use_zlib = getenv("PANDEPTH_GZ_ZLIB") != nullptr;
    if (!use_zlib) {
        static const int level = [] { const char *e = getenv("PANDEPTH_GZ_LEVEL"); int l = e ? atoi(e) : 5; return l < 1 ? 1 : (l > 12 ? 12 : l); }();
        struct Holder { libdeflate_compressor *c = nullptr; ~Holder() { if (c) libdeflate_free_compressor(c); } };
        thread_local Holder h;
        if (!h.c) h.c = libdeflate_alloc_compressor(level);
        if (h.c) {
            std::string out;
            out.resize(libdeflate_gzip_compress_bound(h.c, in.size()));
            const size_t n = libdeflate_gzip_compress(h.c, in.data(), in.size(), &out[0], out.size());
            if (n) { out.resize(n); return out; }
        }
    }
#endif
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return std::string();
    std::string out;
    out.resize(deflateBound(&zs, (uLong)in.size()) + 64);
    zs.next_in = (Bytef *)in.data(); zs.avail_in = (uInt)in.size();
    zs.next_out = (Bytef *)&out[0]; zs.avail_out = (uInt)out.size();
    deflate(&zs, Z_FINISH);
    out.resize(out.size() - zs.avail_out);
    deflateEnd(&zs);
    return out;
}

// Ordered output file whose chunks are formatted and compressed by a pool of threads (the reference formats and
// compresses every row on one thread, PanDepth.cpp:4384 / :4280 -- that is where its -w / -a time goes).
struct GzWriter {
    FILE *fp = nullptr;
    int threads = 1;
    bool wrote = false;
    bool failed = false;                                     // a short write (disk full, quota): reported by close()
    std::string path;
    bool open(const std::string &p, int nthreads) { path = p; fp = fopen(p.c_str(), "wb"); threads = std::max(1, nthreads); return fp != nullptr; }
    void emit(const std::string &z) { if (z.empty()) return; if (fwrite(z.data(), 1, z.size(), fp) != z.size()) failed = true; wrote = true; }
    void put(const std::string &text) { if (text.empty()) return; emit(gz_member(text)); }
    // chunks 0..n-1, text of chunk k = produce(k); chunks must be < 2 GiB each
    template <typename F> void put_chunks(size_t n, F &&produce)
    {
        const size_t wave = (size_t)threads * 2;
        std::vector<std::string> z(wave);
        for (size_t base = 0; base < n; base += wave) {
            const size_t m = std::min(wave, n - base);
            std::atomic<size_t> next{0};
            auto work = [&] { for (;;) { size_t k = next.fetch_add(1); if (k >= m) break; std::string t = produce(base + k); z[k] = t.empty() ? std::string() : gz_member(t); } };
            std::vector<std::thread> th;
            const size_t nt = std::min<size_t>((size_t)threads, m);
            for (size_t i = 1; i < nt; i++) th.emplace_back(work);
            work();
            for (auto &t : th) t.join();
            for (size_t k = 0; k < m; k++) emit(z[k]);
        }
    }
    // false if anything could not be written; the caller reports it and exits non-zero
    bool close()
    {
        if (!fp) return !failed;
        if (!wrote) emit(gz_member(std::string()));
        if (fclose(fp) != 0) failed = true;
        fp = nullptr;
        if (failed) std::cerr << "Error: write failed (disk full?): " << path << std::endl;
        return !failed;
    }
};

// SoA batch of one contig's records in plain (pageable) host memory: workers can fill it before the CUDA context exists.
// The arrays are sized once to their capacity and filled through raw pointers (n reads, nc CIGAR words): the per-record
// cost of the decode loop is a handful of stores, not three vector::push_back calls.
struct Batch {
    size_t cap;                                              // reads (and, for single-op reads, CIGAR words) per batch
    std::vector<int32_t> pos; std::vector<uint32_t> cigar, off;   // allocated by the first add
    size_t n = 0, nc = 0;
    bool single_op = true;                                   // every read so far has exactly one op: off[] is the identity and
    size_t off_valid = 0;                                    //   is only written up to off_valid (filled in when it is needed)
    int32_t *pos_p = nullptr; uint32_t *cig_p = nullptr; size_t cig_cap = 0;   // raw views of pos / cigar for the hot path
    explicit Batch(size_t capacity = 1u << 20) : cap(capacity) {}
    void clear() { n = 0; nc = 0; single_op = true; off_valid = 0; }
    bool full() const { return n >= cap || nc >= cap; }      // callers flush a full batch before the next add
    void alloc() { pos.resize(cap); off.resize(cap + 1); off[0] = 0; cigar.resize(cap + (1u << 12)); views(); }
    void views() { pos_p = pos.data(); cig_p = cigar.data(); cig_cap = cigar.size(); }
    uint32_t *room(size_t k)
    {
        if (pos.empty()) alloc();
        if (nc + k > cigar.size()) { cigar.resize(std::max(cigar.size() * 2, nc + k)); views(); }
        return cigar.data() + nc;
    }
    // off[0..n] made explicit (single-op batches keep it implicit: off[i] == i)
    void materialize_off() { if (pos.empty()) alloc(); for (size_t i = off_valid; i <= n; i++) off[i] = (uint32_t)i; off_valid = n + 1; }
    void commit(int32_t p, size_t k)
    {
        if (pos.empty()) alloc();
        if (k != 1 && single_op) { materialize_off(); single_op = false; }
        pos[n] = p; nc += k; ++n;
        if (!single_op) { off[n] = (uint32_t)nc; off_valid = n + 1; }
    }
    void add(const bam1_t *b)
    {
        const uint32_t k = b->core.n_cigar;
        if (k) memcpy(room(k), bam_get_cigar(b), (size_t)k * 4);
        commit((int32_t)b->core.pos, k);                      // hts_pos_t truncated to int32 like :736
    }
    void add_raw(const RawRead &r)
    {
        if (__builtin_expect(r.n_cigar == 1 && single_op && nc < cig_cap, 1)) {     // the common read: two 4-byte stores
            memcpy(cig_p + nc, r.cigar, 4); pos_p[n] = r.pos; ++nc; ++n;
            return;
        }
        uint32_t *c = room(r.n_cigar);
        if (r.n_cigar) memcpy(c, r.cigar, (size_t)r.n_cigar * 4);
        commit(r.pos, r.n_cigar);
    }
    void add_op(uint32_t op, int64_t len)                     // one more op of the read being built (PAF); split above 2^28-1
    {
        while (len > 0) { const int64_t m = std::min<int64_t>(len, (1 << 28) - 1); *room(1) = ((uint32_t)m << 4) | op; nc++; len -= m; }
    }
    void commit_built(int32_t p, size_t nc_before)
    {
        if (pos.empty()) alloc();
        if (nc - nc_before != 1 && single_op) { materialize_off(); single_op = false; }
        pos[n] = p; ++n;
        if (!single_op) { off[n] = (uint32_t)nc; off_valid = n + 1; }
    }
    // order-independent digest of the batch (PANDEPTH_DRY diagnostics: compare decode paths without a GPU)
    uint64_t digest(int tid) const
    {
        auto mix = [](uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; };
        uint64_t sum = 0;
        for (size_t i = 0; i < n; i++) {
            uint64_t h = mix(((uint64_t)(uint32_t)tid << 32) | (uint32_t)pos[i]);
            const uint32_t k0 = single_op ? (uint32_t)i : off[i], k1 = single_op ? (uint32_t)i + 1 : off[i + 1];
            for (uint32_t k = k0; k < k1; k++) h = mix(h ^ cigar[k]);
            sum += h;
        }
        return sum;
    }
};

// A few pinned staging slots shared by all workers.  A worker copies its finished batch into a slot in parallel with the
// others (outside the engine lock); the locked section is then only a 55 GB/s H2D copy plus two kernel launches.
// (Pushing pageable memory directly measured 2.4 GB/s under the lock and serialised the whole feeder.)
struct PinnedPool {
    struct Slot { int32_t *pos = nullptr; uint32_t *cigar = nullptr, *off = nullptr; size_t cap_r = 0, cap_c = 0; };
    std::vector<Slot> slots; std::vector<int> free_ids;
    std::mutex mu; std::condition_variable cv;
    std::atomic<bool> ready{false};
    void init(int n, size_t reads, size_t cig)
    {
        slots.resize((size_t)n);
        for (int i = 0; i < n; i++) {
            Slot &s = slots[(size_t)i];
            s.pos = (int32_t *)pd_host_alloc(reads * 4); s.off = (uint32_t *)pd_host_alloc((reads + 1) * 4); s.cigar = (uint32_t *)pd_host_alloc(cig * 4);
            if (!s.pos || !s.off || !s.cigar) { slots.resize((size_t)i); break; }
            s.cap_r = reads; s.cap_c = cig;
        }
        std::lock_guard<std::mutex> g(mu);
        for (size_t i = 0; i < slots.size(); i++) free_ids.push_back((int)i);
        ready = !slots.empty();
        cv.notify_all();
    }
    int acquire() { std::unique_lock<std::mutex> g(mu); cv.wait(g, [&] { return !free_ids.empty(); }); int id = free_ids.back(); free_ids.pop_back(); return id; }
    void release(int id) { { std::lock_guard<std::mutex> g(mu); free_ids.push_back(id); } cv.notify_one(); }
    ~PinnedPool() { for (auto &s : slots) { pd_host_free(s.pos); pd_host_free(s.off); pd_host_free(s.cigar); } }
};

struct WorkItem { int tid; int64_t beg, end, prev_end; int r0, r1; };   // 0-based [beg,end); merged regions [r0,r1) of the contig (modes 1-4)

struct Feeder {
    const Options &o;
    pd_engine *eng = nullptr;
    std::shared_future<pd_engine *> eng_future;          // the CUDA context is created while the first ranges decode
    std::mutex eng_mu;
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    std::string err;
    std::mutex err_mu;
    std::vector<WorkItem> items;
    const std::map<int, std::vector<std::pair<int, int>>> &merged;   // per tid, 1-based closed, modes 1-4 only
    bool filter_by_region;
    std::atomic<uint64_t> reads_pushed{0};
    hts_idx_t *shared_idx = nullptr;                     // loaded once, read-only, shared by all workers
    // PANDEPTH_TIMING diagnostics (summed over workers, seconds); PANDEPTH_DRY=1 decodes without pushing
    std::atomic<uint64_t> ns_decode{0}, ns_wait{0}, ns_push{0}, n_push{0}, dry_digest{0};
    const BaiLinear *bai = nullptr;                      // non-null: scan raw BAM records directly (fastbam.hpp)
    std::string input;                                   // the file being fed (multi-file mode feeds several in turn)
    bool used_fast = false;
    PinnedPool *pool = nullptr;                          // pinned staging (exists once the CUDA context does)
    bool dry = getenv("PANDEPTH_DRY") != nullptr;
    bool dry_nodigest = dry && getenv("PANDEPTH_DRY")[0] == '2';     // PANDEPTH_DRY=2: decode timing without the digest
    struct Pending { int tid; bool sorted; Batch b; };
    std::vector<Pending> pending;                        // batches decoded before the CUDA context was ready
    std::mutex pend_mu;
    std::atomic<bool> eng_ok{false};

    Feeder(const Options &oo, std::shared_future<pd_engine *> e, const std::map<int, std::vector<std::pair<int, int>>> &m, bool f)
        : o(oo), eng_future(e), merged(m), filter_by_region(f) {}

    void fail(const std::string &m) { std::lock_guard<std::mutex> g(err_mu); if (!failed.exchange(true)) err = m; }

    bool engine_ready() { return eng_future.wait_for(std::chrono::seconds(0)) == std::future_status::ready; }

    bool push(int tid, Batch &b, bool sorted)
    {
        if (b.n == 0) return true;
        if (dry) { reads_pushed += b.n; if (!dry_nodigest) dry_digest += b.digest(tid); b.clear(); return true; }
        if (!eng_ok.load()) {
            // the CUDA context is still coming up: park the batch and keep decoding (bounded: ~3 GB of parked reads)
            if (!engine_ready()) {
                std::unique_lock<std::mutex> g(pend_mu);
                if (pending.size() < 256) {
                    pending.push_back(Pending{tid, sorted, std::move(b)});
                    b = Batch();
                    return true;
                }
                g.unlock();
                eng_future.wait();
            }
            std::lock_guard<std::mutex> g(eng_mu);
            if (!eng) { eng = eng_future.get(); if (!eng) { fail(std::string("Error: ") + pd_last_error(nullptr)); return false; } }
            eng_ok = true;
        }
        // stage through a pinned slot when the pool is up and the batch fits (parallel memcpy, outside the engine lock)
        int slot = -1;
        if (pool && pool->ready && b.n <= pool->slots[0].cap_r && b.nc <= pool->slots[0].cap_c) {
            slot = pool->acquire();
            PinnedPool::Slot &s = pool->slots[(size_t)slot];
            memcpy(s.pos, b.pos.data(), b.n * 4);
            memcpy(s.cigar, b.cigar.data(), b.nc * 4);
            if (!b.single_op) memcpy(s.off, b.off.data(), (b.n + 1) * 4);
        }
        struct Rel { PinnedPool *p; int id; ~Rel() { if (id >= 0) p->release(id); } } rel{pool, slot};
        const auto w0 = std::chrono::steady_clock::now();
        std::lock_guard<std::mutex> g(eng_mu);
        const auto w1 = std::chrono::steady_clock::now();
        ns_wait += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(w1 - w0).count();
        struct Done { Feeder *f; std::chrono::steady_clock::time_point t; ~Done() { f->ns_push += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t).count(); f->n_push++; } } done{this, w1};
        if (slot >= 0) {
            PinnedPool::Slot &s = pool->slots[(size_t)slot];
            int rc = pd_push_contig(eng, tid, (int64_t)b.n, s.pos, nullptr, nullptr, b.single_op ? nullptr : s.off, s.cigar, sorted ? PD_SORTED : 0u);
            if (rc != PD_OK) { fail(std::string("pd_push_contig: ") + pd_last_error(eng)); return false; }
            reads_pushed += b.n;
        } else if (!push_locked(tid, b, sorted)) return false;
        b.clear();
        return true;
    }

    bool push_locked(int tid, Batch &b, bool sorted)
    {
        int rc = pd_push_contig(eng, tid, (int64_t)b.n, b.pos.data(), nullptr, nullptr, b.single_op ? nullptr : b.off.data(), b.cigar.data(), sorted ? PD_SORTED : 0u);
        if (rc != PD_OK) { fail(std::string("pd_push_contig: ") + pd_last_error(eng)); return false; }
        reads_pushed += b.n;
        return true;
    }
    // after the workers are done: the batches parked during CUDA start-up go through the normal (pinned, parallel) path
    bool finish()
    {
        pd_engine *e = eng_future.get();
        if (!e) { fail(std::string("Error: ") + pd_last_error(nullptr)); return false; }
        { std::lock_guard<std::mutex> g(eng_mu); eng = e; eng_ok = true; }
        std::vector<Pending> todo;
        { std::lock_guard<std::mutex> g(pend_mu); todo.swap(pending); }
        std::atomic<size_t> nxt{0};
        auto work = [&] { for (;;) { size_t k = nxt.fetch_add(1); if (k >= todo.size() || failed) break; push(todo[k].tid, todo[k].b, todo[k].sorted); } };
        std::vector<std::thread> th;
        const size_t nt = std::min<size_t>(std::max(1, o.threads), std::min<size_t>(16, todo.size()));
        for (size_t i = 1; i < nt; i++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        return !failed;
    }

    htsFile *open_input(sam_hdr_t **hdr, int nthreads)
    {
        htsFile *fp = hts_open(input.c_str(), "r");
        if (!fp) return nullptr;
        hts_set_log_level(HTS_LOG_OFF);
        if (fp->format.format == cram) {                                   // :637-643
            hts_set_fai_filename(fp, o.reference.c_str());
            hts_set_opt(fp, CRAM_OPT_DECODE_MD, 0);
            hts_set_opt(fp, CRAM_OPT_REQUIRED_FIELDS, SAM_FLAG | SAM_RNAME | SAM_POS | SAM_MAPQ | SAM_CIGAR);
        }
        if (nthreads > 0) hts_set_opt(fp, HTS_OPT_NTHREADS, nthreads);
        *hdr = sam_hdr_read(fp);
        if (!*hdr) { hts_close(fp); return nullptr; }
        return fp;
    }

    // one worker of the indexed path: own file handle + index, ranges taken from the shared queue
    void indexed_worker()
    {
        sam_hdr_t *hdr = nullptr;
        htsFile *fp = open_input(&hdr, 0);
        hts_idx_t *idx = shared_idx;
        // a .crai lives inside the cram_fd it was loaded through (the hts_idx_t is only a handle to it), so every
        // worker loads its own; BAM indexes are plain read-only tables and are shared
        const bool own_idx = fp && fp->format.format == cram;
        if (own_idx) idx = sam_index_load(fp, input.c_str());
        if (!fp || !idx) { fail("Error: Failed to open the index file or BAM/CRAM file: " + input); if (fp) hts_close(fp); return; }
        bam1_t *b = bam_init1();
        Batch batch;
        const uint32_t xf = o.xflags;
        const auto t_begin = std::chrono::steady_clock::now();
        struct Tot { Feeder *f; std::chrono::steady_clock::time_point t; ~Tot() { f->ns_decode += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t).count(); } } tot{this, t_begin};
        while (!failed) {
            const size_t k = next.fetch_add(1);
            if (k >= items.size()) break;
            const WorkItem &w = items[k];
            hts_itr_t *it = sam_itr_queryi(idx, w.tid, w.beg, w.end);
            if (!it) continue;
            const std::vector<std::pair<int, int>> *regs = nullptr;
            if (filter_by_region) { auto f = merged.find(w.tid); if (f != merged.end()) regs = &f->second; }
            int rcur = w.r0;
            while (sam_itr_next(fp, it, b) >= 0) {
                if (b->core.flag & xf) continue;                              // :732
                if ((int)b->core.qual < o.min_mapq) continue;                 // :733
                const int64_t p = b->core.pos;
                if (p < w.prev_end) continue;                                 // an earlier range already took this read
                if (regs) {
                    // keep only reads the reference would fetch: overlapping a merged region's window chr:first-1..second+1 (:700-706)
                    const int64_t e = bam_endpos(b);
                    while (rcur < w.r1 && (int64_t)(*regs)[rcur].second + 1 <= p) rcur++;       // window end (0-based excl) = second+1
                    bool hit = false;
                    for (int r = rcur; r < w.r1; r++) {
                        const int64_t wb = std::max<int64_t>((int64_t)(*regs)[r].first - 1, 1) - 1;    // 0-based start of the window
                        if (wb >= e) break;
                        if ((int64_t)(*regs)[r].second + 1 > p) { hit = true; break; }
                    }
                    if (!hit) continue;
                }
                batch.add(b);
                if (batch.full()) { if (!push(w.tid, batch, true)) break; }
            }
            hts_itr_destroy(it);
            if (!push(w.tid, batch, true)) break;
        }
        bam_destroy1(b);
        if (own_idx) hts_idx_destroy(idx);
        sam_hdr_destroy(hdr);
        hts_close(fp);
    }

    // indexed path for sorted BAM + .bai: raw record scan, no htslib in the loop
    void fast_worker()
    {
        used_fast = true;
        BamRangeScanner sc;
        if (!sc.open(input)) { fail("Error: Failed to open the BAM/CRAM file: " + input); return; }
        Batch batch;
        const uint32_t xf = o.xflags;
        const auto t_begin = std::chrono::steady_clock::now();
        struct Tot { Feeder *f; std::chrono::steady_clock::time_point t; ~Tot() { f->ns_decode += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t).count(); } } tot{this, t_begin};
        while (!failed) {
            const size_t k = next.fetch_add(1);
            if (k >= items.size()) break;
            const WorkItem &w = items[k];
            const uint64_t voff = bai->start_voffset(w.tid, w.beg);
            if (!voff) continue;
            const std::vector<std::pair<int, int>> *regs = nullptr;
            if (filter_by_region) { auto f = merged.find(w.tid); if (f != merged.end()) regs = &f->second; }
            int rcur = w.r0;
            bool ok_push = true;
            const bool ok = sc.scan(voff, w.tid, w.end, [&](const RawRead &r) {
                if (!ok_push) return;
                if (r.flag & xf) return;                                      // :732
                if ((int)r.mapq < o.min_mapq) return;                         // :733
                const int64_t p = r.pos;
                if (p < w.prev_end) return;                                   // an earlier range already took this read
                if (regs) {
                    int64_t rlen = 0;                                         // bam_endpos(): unmapped or no reference-consuming op -> 1
                    if (!(r.flag & 4)) for (uint32_t c = 0; c < r.n_cigar; c++) { const uint32_t wd = r.word(c); if ((0x18du >> (wd & 15u)) & 1u) rlen += wd >> 4; }
                    const int64_t e = p + (rlen ? rlen : 1);
                    while (rcur < w.r1 && (int64_t)(*regs)[rcur].second + 1 <= p) rcur++;
                    bool hit = false;
                    for (int q = rcur; q < w.r1; q++) {
                        const int64_t wb = std::max<int64_t>((int64_t)(*regs)[q].first - 1, 1) - 1;
                        if (wb >= e) break;
                        if ((int64_t)(*regs)[q].second + 1 > p) { hit = true; break; }
                    }
                    if (!hit) return;
                }
                batch.add_raw(r);
                if (batch.full()) ok_push = push(w.tid, batch, true);
            });
            if (!ok) { fail("Error: malformed BAM while scanning " + input + ": " + sc.error()); break; }
            if (!ok_push || !push(w.tid, batch, true)) break;
        }
    }

    // no-index path for BAM input: ordered parallel BGZF inflate + one record walker (fastbam.hpp BamStreamScanner).
    // Returns false without having pushed anything if the file is not a BAM this scanner handles (caller uses htslib).
    bool stream_all_fast(bool sorted, size_t n_targets)
    {
        BamStreamScanner sc;
        if (!sc.open(input, std::max(1, o.threads))) return false;
        used_fast = true;
        Batch batch;
        int cur = -1;
        bool ok_push = true;
        const uint32_t xf = o.xflags; const int mq = o.min_mapq;
        const bool ok = sc.scan_all([&](const BamStreamScanner::Read &r) {
            if (!ok_push) return;
            if ((int)r.mapq < mq) return;
            if (r.flag & xf) return;
            const int t = r.tid;
            if (t < 0 || (size_t)t >= n_targets) return;
            if (t != cur) { ok_push = push(cur, batch, sorted); cur = t; if (!ok_push) return; }
            batch.add_raw(r);
            if (batch.full()) ok_push = push(cur, batch, sorted);
        });
        if (!ok) { fail("Error: malformed BAM while reading " + input + ": " + sc.error()); return true; }
        if (ok_push && !failed) push(cur, batch, sorted);
        return true;
    }

    // no-index path (:4604-4713): one pass over the file, batches split at contig changes
    bool stream_all(bool sorted)
    {
        sam_hdr_t *hdr = nullptr;
        htsFile *fp = open_input(&hdr, o.threads);
        if (!fp) { fail("Error: Failed to open the BAM/CRAM file: " + input); return false; }
        bam1_t *b = bam_init1();
        Batch batch;
        int cur = -1;
        while (sam_read1(fp, hdr, b) >= 0) {
            if ((int)b->core.qual < o.min_mapq) continue;
            if (b->core.flag & o.xflags) continue;
            const int t = b->core.tid;
            if (t < 0) continue;
            if (t != cur) { if (!push(cur, batch, sorted)) break; cur = t; }
            batch.add(b);
            if (batch.full()) { if (!push(cur, batch, sorted)) break; }
        }
        if (!failed) push(cur, batch, sorted);
        bam_destroy1(b);
        sam_hdr_destroy(hdr);
        hts_close(fp);
        return !failed;
    }

    // PAF input (paf_main, :1536-1623): one text line per alignment.  Column 6 = target name, 8/9 = target start/end,
    // 12 = mapping quality; a `cg:Z:` tag (token index > 1) is walked like a BAM CIGAR from the 0-based target start,
    // otherwise the line covers [start-1, end).  Lines carrying "tp:A:S" are dropped when -x has bit 0x100 (:1552-1558),
    // lines with mapq < -q are dropped (:1564-1568).  Unknown target names land on target 0 (`Chr2IntMap[inf[5]]`
    // default-inserts 0, :1562).  Lines are not sorted: batches go through the unsorted (delta scatter) path.
    bool paf_stream(std::map<std::string, int> &chr2int, size_t n_targets)
    {
        LineReader in(input);
        if (!in.ok()) { fail("open PAF file error: " + input); return false; }
        // per-target batches: bounded in total (many-contig assemblies), flushed through the unsorted path
        std::vector<Batch> batches(n_targets, Batch(std::min<size_t>(1u << 20, std::max<size_t>(4096, (64u << 20) / std::max<size_t>(1, n_targets)))));
        std::string line;
        std::vector<std::pair<const char *, size_t>> tok;
        const bool drop_secondary = (o.xflags & 0x100u) != 0;
        auto to_int = [](const char *p, size_t n) { char b[32]; if (n > 31) n = 31; memcpy(b, p, n); b[n] = 0; return atoi(b); };
        while (!failed && in.next(line)) {
            if (line.empty()) continue;
            if (drop_secondary && line.find("tp:A:S") != std::string::npos) continue;
            tok.clear();
            const char *s = line.data(), *e = s + line.size();
            while (s < e) {                                                    // split(line, inf, " \t")  (include/comm.h:98-109)
                while (s < e && (*s == ' ' || *s == '\t')) s++;
                const char *t = s;
                while (s < e && *s != ' ' && *s != '\t') s++;
                if (s > t) tok.emplace_back(t, (size_t)(s - t));
            }
            if (tok.size() < 12) continue;                                     // (the reference indexes inf[11] unchecked)
            int tid = 0;
            { auto it = chr2int.find(std::string(tok[5].first, tok[5].second)); if (it != chr2int.end()) tid = it->second; }
            if (to_int(tok[11].first, tok[11].second) < o.min_mapq) continue;
            int32_t a = to_int(tok[7].first, tok[7].second), z = to_int(tok[8].first, tok[8].second);
            if (a > z) std::swap(a, z);
            if ((size_t)tid >= n_targets) continue;
            Batch &b = batches[(size_t)tid];
            int cg = -1;
            for (size_t k = 0; k < tok.size(); k++) if (tok[k].second >= 5 && memcmp(tok[k].first, "cg:Z:", 5) == 0) { cg = (int)k; break; }
            const size_t c0 = b.nc;
            if (cg > 1) {                                                      // :1582-1613
                if (a < 0) continue;
                const char *q = tok[(size_t)cg].first + 5, *qe = tok[(size_t)cg].first + tok[(size_t)cg].second;
                while (q < qe) {
                    int64_t n = 0;
                    while (q < qe && *q >= '0' && *q <= '9') { n = n * 10 + (*q - '0'); if (n > INT32_MAX) n = INT32_MAX; q++; }
                    const char op = q < qe ? *q : 0;
                    q++;
                    switch (op) {
                    case 'M': b.add_op(0, n); break;
                    case '=': b.add_op(7, n); break;
                    case 'X': b.add_op(8, n); break;
                    case 'D': b.add_op(2, n); break;
                    case 'N': b.add_op(3, n); break;
                    default: break;                                            // I S H P ...: no depth, no advance
                    }
                }
                b.commit_built(a, c0);
            } else {                                                           // :1614-1620  for (i = StartRead-1; i < EndRead; ++i)
                const int32_t s0 = a > 0 ? a - 1 : 0;                          // (the reference's write at index -1 is dropped)
                if (z > s0) b.add_op(0, (int64_t)z - s0);
                b.commit_built(s0, c0);
            }
            if (b.full()) { if (!push(tid, b, false)) break; }
        }
        for (size_t t = 0; t < n_targets && !failed; t++) push((int)t, batches[t], false);
        return !failed;
    }
};

bool file_exists(const std::string &p) { return access(p.c_str(), 0) == 0; }

// CPUs this process may actually use: the cgroup v2 / v1 CPU quota if there is one, else the online CPU count.  Decode
// workers beyond that only get throttled (possibly while holding the engine lock), so -t is clamped to it.
int usable_cpus()
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    double quota = 0;
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char a[64]; long long period = 0;
        if (fscanf(f, "%63s %lld", a, &period) == 2 && strcmp(a, "max") != 0 && period > 0) quota = atof(a) / (double)period;
        fclose(f);
    } else {
        long long q = -1, p = 0;
        if (FILE *fq = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (fscanf(fq, "%lld", &q) != 1) q = -1; fclose(fq); }
        if (FILE *fp = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(fp, "%lld", &p) != 1) p = 0; fclose(fp); }
        if (q > 0 && p > 0) quota = (double)q / (double)p;
    }
    if (quota >= 1.0 && quota < (double)n) n = (long)(quota + 0.5);
    return (int)n;
}

// PANDEPTH_TIMING=1: phase wall times on stderr
struct Timer {
    bool on = getenv("PANDEPTH_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    void mark(const char *what)
    {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[timing] %-22s %8.3f s  (total %.3f s)\n", what, std::chrono::duration<double>(now - last).count(),
                std::chrono::duration<double>(now - t0).count());
        last = now;
    }
};

}  // namespace

int main(int argc, char **argv)
{
    Options o;
    if (parse_args(argc, argv, o) != 0) return 0;
    Timer tm;
    if (!getenv("PANDEPTH_NO_CLAMP")) {
        const int cpus = usable_cpus();
        if (o.threads > cpus) { if (tm.on) fprintf(stderr, "[timing] -t %d clamped to the %d usable CPUs (cgroup quota / online CPUs)\n", o.threads, cpus); o.threads = cpus; }
    }

    std::string ext = o.input.substr(o.input.rfind('.') == std::string::npos ? o.input.size() : o.input.rfind('.') + 1);
    if (ext == "gz") {
        std::string a = o.input.substr(0, o.input.rfind('.'));
        ext = a.substr(a.rfind('.') == std::string::npos ? a.size() : a.rfind('.') + 1);
    }
    const bool multi = o.inputs.size() > 1;                  // FileNum > 1 -> List_main / BamList_main (:3452-3458)
    if (multi) std::cout << "INFO: Run multi-file data " << std::endl;
    const bool is_paf = ext == "paf" || ext == "PAF";        // :3428 / :3474 (a trailing .gz is looked through)
    if (is_paf) std::cout << (multi ? "INFO: Run PAF format data " : "INFO: Run paf Format data ") << std::endl;

    std::vector<Target> targets;
    std::map<std::string, int> paf_chr2int;                  // Chr2IntMap of paf_main (:867)
    sam_hdr_t *hdr = nullptr;
    bool is_bam = false, is_cram = false, sorted_hdr = false;
    if (is_paf) {
        // targets: column 6/7 of the first PAF in first-seen order (:918-941).  With -r the reference takes them from the
        // FASTA and switches to its GC report (RefIn, :878-909 / :1488) -- GC content is outside the depth path.
        if (!o.reference.empty()) { std::cerr << "Error: -r with PAF input selects the GC report, which pandepth-b200 does not provide" << std::endl; return 1; }
        {
            LineReader in0(o.input);
            if (!in0.ok()) { std::cerr << "open PAF file error: " << o.input << std::endl; return 1; }
            std::string line;
            while (in0.next(line)) {
                if (line.empty()) continue;
                size_t k = 0, p = 0; std::string chr, len;
                while (k < 7) {                                                  // isone>>tmp1..tmp5>>chr>>chrlength
                    p = line.find_first_not_of(" \t\r\n\v\f", p);
                    if (p == std::string::npos) break;
                    const size_t q = line.find_first_of(" \t\r\n\v\f", p);
                    if (k == 5) chr = line.substr(p, q == std::string::npos ? std::string::npos : q - p);
                    if (k == 6) len = line.substr(p, q == std::string::npos ? std::string::npos : q - p);
                    k++;
                    if (q == std::string::npos) break;
                    p = q;
                }
                if (k < 7) continue;
                if (!paf_chr2int.count(chr)) { paf_chr2int[chr] = (int)targets.size(); targets.push_back(Target{chr, (uint32_t)atoi(len.c_str())}); }
            }
        }
    } else {
        // header (:3482-3507)
        htsFile *in = hts_open(o.input.c_str(), "r");
        hts_set_log_level(HTS_LOG_OFF);
        if (!in) { std::cerr << "Error: Failed to open the BAM/CRAM file: " << o.input << std::endl; return 1; }
        if (in->format.format == cram) hts_set_fai_filename(in, o.reference.c_str());
        is_bam = in->format.format == bam; is_cram = in->format.format == cram;
        hdr = sam_hdr_read(in);
        if (!hdr) { std::cerr << "Error: Failed to read the header for the BAM/CRAM file: " << o.input << std::endl; return 1; }
        targets.resize((size_t)hdr->n_targets);
        for (int i = 0; i < hdr->n_targets; i++) { targets[i].name = hdr->target_name[i]; targets[i].len = hdr->target_len[i]; }
        const char *txt = sam_hdr_str(hdr);
        std::string text = txt ? txt : "";
        size_t a = text.find("\tSO:");                                       // :4537-4549
        if (a != std::string::npos) { a += 4; size_t b = text.find_first_of("\n\t", a); sorted_hdr = text.substr(a, b - a) == "coordinate"; }
        hts_close(in);
    }

    GeneMap genes;
    std::string err;
    if (!build_regions(o, targets, genes, err)) { std::cerr << err << std::endl; return 1; }
    const int mode = o.region_mode;

    // the CUDA context (~0.5 s) is created in the background while the host threads already decode
    int device = 0;
    if (const char *d = getenv("PANDEPTH_DEVICE")) device = atoi(d);
    std::vector<uint32_t> tlen(targets.size());
    std::vector<uint8_t> owned(targets.size(), 0);
    for (size_t t = 0; t < targets.size(); t++) { tlen[t] = targets[t].len; owned[t] = genes.count((int)t) ? 1 : 0; }
    PinnedPool pool;
    const bool dry = getenv("PANDEPTH_DRY") != nullptr;      // decode only: no GPU is touched (diagnostics)
    // -d and the depth width are known before the first read: let the engine keep per-segment partials for pd_reduce
    const bool idx_early = (file_exists(o.input + ".bai") || file_exists(o.input + ".crai") || file_exists(o.input + ".csi")) && o.use_index;
    const uint32_t hint_mask = (is_paf || multi || !idx_early || o.site_output || mode == 6) ? 0x3FFFFu : 0xFFFFFFFFu;
    const bool hint_reduce = mode != 6;
    std::shared_future<pd_engine *> eng_future = std::async(std::launch::async, [&]() -> pd_engine * {
        if (dry) return nullptr;
        pd_engine *e = pd_create(device, (int)targets.size(), tlen.data(), owned.data(), o.xflags, o.min_mapq);
        if (e && hint_reduce) pd_set_reduce_hint(e, (uint32_t)o.min_dep, hint_mask);
        if (e) pool.init(8, (1u << 20) + 1024, (1u << 20) + (1u << 18));      // 8 x 13 MB pinned; bigger batches go pageable
        return e;
    }).share();

    const std::string prefix = output_prefix(o.out_arg);
    GzWriter out;
    const std::string stat_path = prefix + stat_suffix(mode);
    if (!out.open(stat_path, o.threads)) { std::cerr << "open OUT File error: " << stat_path << std::endl; eng_future.wait(); return 0; }

    const bool have_index = file_exists(o.input + ".bai") || file_exists(o.input + ".crai") || file_exists(o.input + ".csi");   // :4121-4124
    const bool indexed = have_index && o.use_index;
    // the reference's 18-bit SiteInfo paths: no index, -a, mode 6 (:4127, :4553) or several input files (BamList_main :2687);
    // everything else is the uint32 window path
    const bool whole = is_paf || multi || !indexed || o.site_output || mode == 6;
    const uint32_t mask = whole ? 0x3FFFFu : 0xFFFFFFFFu;

    // merged regions per contig (RegionMerger, :3912-3972)
    std::map<int, std::vector<std::pair<int, int>>> merged;
    for (const auto &kv : genes) merged[kv.first] = merge_regions(kv.second);
    tm.mark("header+regions");
    // PANDEPTH_DRY=1 PANDEPTH_DUMP_REGIONS=<file>: the region model as text (host-logic tests; no GPU is touched)
    if (const char *dump = getenv("PANDEPTH_DUMP_REGIONS")) {
        if (FILE *f = fopen(dump, "w")) {
            fprintf(f, "mode %d mask %x\n", mode, (unsigned)mask);
            if (mode == 5) {
                // implicit -w tiles (pd_host.hpp n_tiles), in the std::map<string> order of their ids; every tile is its own
                // merged region (adjacent tiles do not merge, :3919-3972) and is handed over exactly once
                for (const auto &kv : genes) {
                    const long long W = o.tile_win, len = tlen[kv.first], nt = n_tiles(len, W);
                    std::vector<std::pair<std::string, long long>> ids;
                    for (long long k = 0; k < nt; k++) ids.emplace_back(targets[(size_t)kv.first].name + std::to_string(1 + k * W), k);
                    std::sort(ids.begin(), ids.end());
                    for (const auto &id : ids) {
                        const long long S = 1 + id.second * W, E = std::min(S + W - 1, len);
                        fprintf(f, "G\t%d\t%s\t%lld\t%lld\t%lld\t%lld,%lld\n", kv.first, id.first.c_str(), S, E, E - S + 1, S, E);
                    }
                    fprintf(f, "M\t%d", kv.first);
                    for (long long k = 0; k < nt; k++) fprintf(f, "\t%lld,%lld", 1 + k * W, std::min(k * W + W, len));
                    fputc('\n', f);
                    for (const auto &id : ids) fprintf(f, "W\t%d\t%s\t1\n", kv.first, id.first.c_str());
                }
            } else
            for (const auto &kv : genes) {
                for (const auto &gk : kv.second) {
                    const Gene &g = gk.second;
                    fprintf(f, "G\t%d\t%s\t%d\t%d\t%llu", kv.first, gk.first.c_str(), g.start, g.end, (unsigned long long)g.length);
                    for (const auto &c : g.cds) fprintf(f, "\t%d,%d", c.first, c.second);
                    fputc('\n', f);
                }
                fprintf(f, "M\t%d", kv.first);
                for (const auto &r : merged[kv.first]) fprintf(f, "\t%d,%d", r.first, r.second);
                fputc('\n', f);
                for (const auto &wk : indexed_window_counts(merged[kv.first], kv.second, (int)tlen[kv.first])) fprintf(f, "W\t%d\t%s\t%d\n", kv.first, wk.first.c_str(), wk.second);
            }
            fclose(f);
        }
    }

    // PANDEPTH_DRY=1 PANDEPTH_SELFTEST_ROWS=<prefix>: the report writers (row formatters + parallel gzip members) on a
    // fixed synthetic table -> <prefix>.win.gz and <prefix>.site.gz (host-logic tests; no GPU is touched)
    if (const char *pre = getenv("PANDEPTH_SELFTEST_ROWS")) {
        const int64_t W = 100, CC = 10000003, n_win = (CC - 1 + W - 1) / W;
        std::vector<uint32_t> cov((size_t)n_win); std::vector<uint64_t> sum((size_t)n_win);
        for (int64_t k = 0; k < n_win; k++) {
            const int64_t len = std::min(W, CC - k * W);
            cov[(size_t)k] = (uint32_t)((k * 7919) % (len + 1));
            sum[(size_t)k] = (uint64_t)cov[(size_t)k] * (uint64_t)((k * 31) % 500);
        }
        GzWriter w;
        if (w.open(std::string(pre) + ".win.gz", o.threads)) {
            w.put(stat_header(6));
            const int64_t SUB = 1 << 14;
            w.put_chunks((size_t)((n_win + SUB - 1) / SUB), [&](size_t c) { const int64_t a = (int64_t)c * SUB, b = std::min(n_win, a + SUB); return format_window_rows("selftest", W, CC, cov.data(), sum.data(), a, b); });
            w.close();
        }
        const int32_t n = 600000, SUBS = 1 << 16;
        std::vector<uint32_t> d((size_t)n);
        for (int32_t j = 0; j < n; j++) d[(size_t)j] = (uint32_t)((uint64_t)j * 2654435761ull);
        GzWriter sw;
        if (sw.open(std::string(pre) + ".site.gz", o.threads)) {
            sw.put_chunks((size_t)((n + SUBS - 1) / SUBS), [&](size_t k) { const int32_t a = (int32_t)k * SUBS, b = std::min<int32_t>(n, a + SUBS); return format_site_rows("selftest", 5, d.data(), a, b); });
            sw.close();
        }
        GzWriter ew;                                          // a file nothing was written to is still a valid (empty) .gz
        if (ew.open(std::string(pre) + ".empty.gz", 1)) ew.close();
    }

    // ---- feed ---------------------------------------------------------------------------------------------------
    const bool region_modes = mode >= 1 && mode <= 4;
    Feeder feeder(o, eng_future, merged, false);
    feeder.pool = &pool;
    for (size_t fi = 0; fi < o.inputs.size() && !feeder.failed; fi++) {
        const std::string &path = o.inputs[fi];
        feeder.input = path;
        if (is_paf) { feeder.paf_stream(paf_chr2int, targets.size()); continue; }
        bool f_bam = is_bam, f_cram = is_cram, f_sorted = sorted_hdr;
        if (fi > 0) {                                        // later files: own format / sort order, the first file's targets (:2701-2862)
            htsFile *fp = hts_open(path.c_str(), "r");
            if (!fp) { std::cerr << "Error: Failed to open the BAM/CRAM file: " << path << std::endl; continue; }
            if (fp->format.format == cram) hts_set_fai_filename(fp, o.reference.c_str());
            f_bam = fp->format.format == bam; f_cram = fp->format.format == cram;
            sam_hdr_t *h2 = sam_hdr_read(fp);
            f_sorted = false;
            if (h2) {
                const char *txt = sam_hdr_str(h2);
                std::string text = txt ? txt : "";
                size_t a = text.find("\tSO:");
                if (a != std::string::npos) { a += 4; size_t b = text.find_first_of("\n\t", a); f_sorted = text.substr(a, b - a) == "coordinate"; }
                sam_hdr_destroy(h2);
            }
            hts_close(fp);
        }
        const bool f_indexed = (file_exists(path + ".bai") || file_exists(path + ".crai") || file_exists(path + ".csi")) && o.use_index;
        feeder.filter_by_region = f_indexed && region_modes;
        if (f_indexed) {
            if (feeder.items.empty()) {
                int64_t total_len = 0;
                for (const auto &kv : merged) total_len += tlen[kv.first];
                // bases per work item: ~6 items per thread, but never so small that BGZF blocks are decoded twice at every boundary
                const int64_t RANGE = std::max<int64_t>(1 << 20, std::min<int64_t>(8 << 20, total_len / (6 * (int64_t)o.threads) + 1));
                const int64_t GAP = 64 << 10;                // merged regions closer than this share one index query
                for (const auto &kv : merged) {
                    const int tid = kv.first;
                    const int64_t len = tlen[tid];
                    int64_t prev_end = 0;
                    if (!region_modes) {
                        for (int64_t b = 0; b < len; b += RANGE) { feeder.items.push_back({tid, b, std::min(len, b + RANGE), prev_end, 0, 0}); prev_end = std::min(len, b + RANGE); }
                    } else {
                        const auto &regs = kv.second;
                        size_t a = 0;
                        while (a < regs.size()) {
                            size_t b = a + 1;
                            int64_t beg = std::max<int64_t>((int64_t)regs[a].first - 1, 1) - 1;     // chr:first-1 .. second+1 (:700-706)
                            int64_t end = std::min<int64_t>((int64_t)regs[a].second + 1, len);
                            while (b < regs.size() && (int64_t)regs[b].first - 2 - end < GAP && end - beg < 2 * RANGE) { end = std::min<int64_t>((int64_t)regs[b].second + 1, len); b++; }
                            if (end > beg) feeder.items.push_back({tid, beg, end, prev_end, (int)a, (int)b});
                            prev_end = std::max(prev_end, end);
                            a = b;
                        }
                    }
                }
            }
            feeder.next = 0;
            // sorted BAM + .bai: raw-record scanner; anything else (CRAM, .csi only, PANDEPTH_HTS_ITER=1): htslib iterators
            BaiLinear bai;
            feeder.bai = nullptr; feeder.shared_idx = nullptr;
            if (f_bam && !getenv("PANDEPTH_HTS_ITER") && file_exists(path + ".bai") && bai.load(path + ".bai") && bai.refs.size() == targets.size())
                feeder.bai = &bai;
            if (!feeder.bai && !f_cram) {
                htsFile *ifp = hts_open(path.c_str(), "r");
                if (ifp) { feeder.shared_idx = sam_index_load(ifp, path.c_str()); hts_close(ifp); }
            }
            const int nthreads = (int)std::max<size_t>(1, std::min<size_t>((size_t)o.threads, feeder.items.size()));
            std::vector<std::thread> th;
            for (int i = 0; i < nthreads; i++) th.emplace_back([&] { if (feeder.bai) feeder.fast_worker(); else feeder.indexed_worker(); });
            for (auto &t : th) t.join();
            if (feeder.shared_idx) hts_idx_destroy(feeder.shared_idx);
            feeder.bai = nullptr; feeder.shared_idx = nullptr;
        } else {
            if (f_sorted) std::cout << "Warning: PanDepth will run in No Index mode: " << path << std::endl;                       // :4606
            else std::cout << "Warning: Can't find index file of input BAM/CRAM. PanDepth will run in No Index mode: " << path << std::endl;   // :4677
            // BAM: own parallel stream reader; SAM / CRAM / PANDEPTH_HTS_ITER=1: htslib's sam_read1
            if (!(f_bam && !getenv("PANDEPTH_HTS_ITER") && feeder.stream_all_fast(f_sorted, targets.size()))) feeder.stream_all(f_sorted);
        }
    }
    if (dry) {
        tm.mark("decode (dry)");
        std::cout << "dry: reads=" << feeder.reads_pushed.load() << " digest=" << feeder.dry_digest.load() << " path=" << (feeder.used_fast ? "fastbam" : "htslib") << std::endl;
        if (feeder.failed) { std::cerr << feeder.err << std::endl; return 1; }
        return 0;
    }
    feeder.finish();
    if (tm.on) fprintf(stderr, "[timing] workers: total %.3f thread-s, of which mutex wait %.3f s, inside push %.3f s (%llu pushes, %llu reads)\n",
                       feeder.ns_decode / 1e9, feeder.ns_wait / 1e9, feeder.ns_push / 1e9, (unsigned long long)feeder.n_push.load(), (unsigned long long)feeder.reads_pushed.load());
    pd_engine *eng = eng_future.get();
    if (!eng) { std::cerr << "Error: " << pd_last_error(nullptr) << std::endl; return 1; }
    tm.mark("decode+push");
    if (feeder.failed) { std::cerr << feeder.err << std::endl; pd_destroy(eng); return 1; }
    if (pd_finalize(eng) != PD_OK) { std::cerr << "Error: " << pd_last_error(eng) << std::endl; pd_destroy(eng); return 1; }
    std::cout << "INFO: Input data read done" << std::endl;                                                                        // :4879

    // ---- -a per-site dump (:4264-4284): D2H in 4 M-base pieces, rows formatted + compressed by the thread pool -------
    if (o.site_output) {
        GzWriter site;
        if (site.open(prefix + ".SiteDepth.gz", o.threads)) {
            const int32_t CH = 4 << 20, SUB = 1 << 18;
            std::vector<uint32_t> d((size_t)CH);
            for (size_t t = 0; t < targets.size(); t++) {
                if (!genes.count((int)t)) continue;
                const std::string &nm = targets[t].name;
                for (int64_t s = 0; s < (int64_t)tlen[t]; s += CH) {
                    const int32_t n = (int32_t)std::min<int64_t>(CH, (int64_t)tlen[t] - s);
                    if (pd_depth_read(eng, (int32_t)t, (int32_t)s, n, d.data(), 0) != PD_OK) { std::cerr << "Error: " << pd_last_error(eng) << std::endl; return 1; }
                    site.put_chunks((size_t)((n + SUB - 1) / SUB), [&](size_t k) {
                        const int32_t a = (int32_t)k * SUB, b = std::min<int32_t>(n, a + SUB);
                        return format_site_rows(nm, s, d.data(), a, b);
                    });
                }
            }
            if (!site.close()) return 1;
        }
    }
    tm.mark("finalize(+site dump)");

    // ---- reduce + report -----------------------------------------------------------------------------------------
    uint64_t SS_Len = 0, SS_Cov = 0, SS_TotalD = 0;
    out.put(stat_header(mode));
    char row[512];
    if (mode == 6) {                                                                         // :4353-4394
        std::vector<uint32_t> cov; std::vector<uint64_t> sum;
        for (size_t t = 0; t < targets.size(); t++) {
            if (!genes.count((int)t)) continue;
            const int64_t CC = tlen[t], W = o.win;
            const int64_t n_win = CC > 1 ? (CC - 1 + W - 1) / W : 0;                         // for (j=1; j<CC; j+=W)
            cov.assign((size_t)n_win, 0); sum.assign((size_t)n_win, 0);
            if (n_win && pd_reduce_windows(eng, (int32_t)t, 0, (int32_t)W, n_win, (uint32_t)o.min_dep, 0x3FFFFu, cov.data(), sum.data(), 0) != PD_OK) {
                std::cerr << "Error: " << pd_last_error(eng) << std::endl; return 1;
            }
            const char *nm = targets[t].name.c_str();
            const int64_t SUB = 1 << 16;
            out.put_chunks((size_t)((n_win + SUB - 1) / SUB), [&](size_t c) {
                const int64_t a = (int64_t)c * SUB, b = std::min(n_win, a + SUB);
                return format_window_rows(nm, W, CC, cov.data(), sum.data(), a, b);
            });
            for (int64_t k = 0; k < n_win; k++) {
                const int64_t j = 1 + k * W;
                const int64_t End = std::min(j - 1 + W, CC);
                SS_Cov += (uint64_t)(int64_t)(int)cov[(size_t)k]; SS_Len += (uint64_t)(int64_t)(int)(End - j + 1); SS_TotalD += (uint64_t)(int64_t)(int)sum[(size_t)k];
            }
        }
    } else if (mode == 5) {
        // -w >= 150: the pseudo-gene tiles are enumerated arithmetically (pd_host.hpp n_tiles) -- one pd_reduce over all of them,
        // rows formatted and compressed in parallel.  Same rows, same order (GeneStart) and same totals as the generic
        // branch below produces from materialised genes (:5080-5118); every tile counts exactly once (pd_host.hpp).
        const long long W = o.tile_win;
        std::vector<int> tids;
        std::vector<size_t> base;                                                            // first interval of each contig
        size_t n_iv = 0;
        for (auto &kv : genes) { tids.push_back(kv.first); base.push_back(n_iv); n_iv += (size_t)n_tiles(tlen[kv.first], W); }
        std::vector<int32_t> iv_t(n_iv), iv_s(n_iv), iv_e(n_iv);
        for (size_t ci = 0; ci < tids.size(); ci++) {
            const long long len = tlen[tids[ci]], nt = n_tiles(len, W);
            for (long long k = 0; k < nt; k++) {
                iv_t[base[ci] + (size_t)k] = tids[ci];
                iv_s[base[ci] + (size_t)k] = (int32_t)(k * W);
                iv_e[base[ci] + (size_t)k] = (int32_t)std::min(k * W + W, len);
            }
        }
        std::vector<uint64_t> cov(n_iv), sum(n_iv);
        if (pd_reduce(eng, (int64_t)n_iv, iv_t.data(), iv_s.data(), iv_e.data(), (uint32_t)o.min_dep, mask, cov.data(), sum.data(), 0) != PD_OK) {
            std::cerr << "Error: " << pd_last_error(eng) << std::endl; return 1;
        }
        for (size_t ci = 0; ci < tids.size(); ci++) {
            const long long len = tlen[tids[ci]], nt = n_tiles(len, W);
            const std::string &chr = targets[(size_t)tids[ci]].name;
            const uint64_t *c = cov.data() + base[ci], *d = sum.data() + base[ci];
            const long long SUB = 1 << 16;
            out.put_chunks((size_t)((nt + SUB - 1) / SUB), [&](size_t ch) {
                const long long a = (long long)ch * SUB, b = std::min(nt, a + SUB);
                return format_tile_rows(chr, W, len, c, d, a, b);
            });
            for (long long k = 0; k < nt; k++) {
                SS_Cov += (uint64_t)(int64_t)(int)(int64_t)c[k];
                SS_Len += (uint64_t)(std::min(k * W + W, len) - k * W);
                SS_TotalD += d[k];
            }
        }
    } else {
        // every CDS interval of every gene in one pd_reduce call (StatChrDepthWin / StatChrDepthLowMEM)
        std::vector<int32_t> iv_t, iv_s, iv_e;
        std::vector<Gene *> iv_g;
        for (auto &kv : genes) {
            std::map<std::string, int> times;
            if (!whole) times = indexed_window_counts(merged[kv.first], kv.second, (int)tlen[kv.first]);
            for (auto &gk : kv.second) {
                if (!whole && times[gk.first] == 0) continue;                                 // never handed to StatChrDepthWin (:299-308)
                for (const auto &c : gk.second.cds) { iv_t.push_back(kv.first); iv_s.push_back(c.first - 1); iv_e.push_back(c.second); iv_g.push_back(&gk.second); }
            }
        }
        std::vector<uint64_t> cov(iv_t.size()), sum(iv_t.size());
        if (pd_reduce(eng, (int64_t)iv_t.size(), iv_t.data(), iv_s.data(), iv_e.data(), (uint32_t)o.min_dep, mask, cov.data(), sum.data(), 0) != PD_OK) {
            std::cerr << "Error: " << pd_last_error(eng) << std::endl; return 1;
        }
        for (size_t i = 0; i < iv_g.size(); i++) { iv_g[i]->cover += (int64_t)cov[i]; iv_g[i]->depth += sum[i]; }

        std::vector<int> tids;
        for (auto &kv : genes) tids.push_back(kv.first);
        std::vector<uint64_t> pL(tids.size(), 0), pC(tids.size(), 0), pD(tids.size(), 0);
        out.put_chunks(tids.size(), [&](size_t ci) {                                         // one chunk per contig, rows in order
            auto &gm = genes[tids[ci]];
            const std::string &chr = targets[(size_t)tids[ci]].name;
            std::string text;
            char num[256];
            if (mode == 0) {                                                                   // :5051-5079
                uint64_t L = 0, C = 0, D = 0;
                for (auto &gk : gm) { L += gk.second.length; C += (uint64_t)(int64_t)(int)gk.second.cover; D += gk.second.depth; }
                pL[ci] = L; pC[ci] = C; pD[ci] = D;
                int m = snprintf(num, sizeof num, "%s\t%llu\t%llu\t%llu\t%s\t%s\n", chr.c_str(), (unsigned long long)L, (unsigned long long)C,
                                 (unsigned long long)D, fmt_ratio_pct((double)C, (double)L).c_str(), fmt_ratio((double)D, (double)L).c_str());
                text.append(num, (size_t)m);
            } else {                                                                           // :5011-5050 / :5080-5118
                std::map<int, std::string> by_start;                                           // rows by GeneStart; ties in GeneID order
                for (auto &gk : gm) {
                    const Gene &g = gk.second;
                    const int GeneCover = (int)g.cover;
                    pC[ci] += (uint64_t)(int64_t)GeneCover; pL[ci] += g.length; pD[ci] += g.depth;
                    const double Coverage = GeneCover * 100.0 / g.length, MeanDepth = g.depth * 1.0 / g.length;
                    std::string line = chr + "\t" + std::to_string(g.start) + "\t" + std::to_string(g.end) + "\t";
                    if (mode != 5) line += gk.first + "\t";
                    snprintf(num, sizeof num, "%llu\t%d\t%llu\t%.2f\t%.2f", (unsigned long long)g.length, GeneCover, (unsigned long long)g.depth, Coverage, MeanDepth);
                    line += num;
                    auto it = by_start.find(g.start);
                    if (it == by_start.end()) by_start.emplace(g.start, line);
                    else { it->second += "\n"; it->second += line; }
                }
                for (auto &r : by_start) { text += r.second; text += "\n"; }
            }
            return text;
        });
        for (size_t ci = 0; ci < tids.size(); ci++) { SS_Len += pL[ci]; SS_Cov += pC[ci]; SS_TotalD += pD[ci]; }
    }
    {
        const double Coverage = SS_Cov * 100.0 / SS_Len, MeanDepth = SS_TotalD * 1.0 / SS_Len;   // :5119-5122
        int m = snprintf(row, sizeof row, "##RegionLength: %llu\tCoveredSite: %llu\tCoverage(%%): %.2f\tMeanDepth: %.2f\n",
                         (unsigned long long)SS_Len, (unsigned long long)SS_Cov, Coverage, MeanDepth);
        out.put(std::string(row, (size_t)m));
    }
    if (!out.close()) return 1;
    tm.mark("reduce+report");
    // every output is flushed and closed: leave without the teardown of the CUDA context, 24 GB of device memory and the
    // parked/pinned host buffers -- the OS reclaims all of it, and the orderly teardown only costs wall time
    fflush(stdout); fflush(stderr);
    if (!getenv("PANDEPTH_ORDERLY_EXIT")) _exit(0);
    pd_destroy(eng);
    if (hdr) sam_hdr_destroy(hdr);
    tm.mark("pd_destroy");
    return 0;
}
