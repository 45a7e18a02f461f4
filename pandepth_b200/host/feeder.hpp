// feeder.hpp -- the decode side of the host: record batches, pinned staging, and the feeders that turn SAM / BAM / CRAM / PAF
// input into pd_push_contig calls (role of ProDealChrBambai and its clones, PanDepth.cpp:618-790, :4604-4713, :1536-1623).
#pragma once
#include "pd_host.hpp"
#include "fastbam.hpp"
#include "bgzf_feed.hpp"
#include "../../include/pandepth_b200.h"

#include "sam.h"
#include "hts.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace pdh {

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
    int32_t last_pos = INT32_MIN; bool in_order = true;      // positions seen so far are non-decreasing (PD_SORTED is only claimed then)
    explicit Batch(size_t capacity = 1u << 20) : cap(capacity) {}
    void clear() { n = 0; nc = 0; single_op = true; off_valid = 0; last_pos = INT32_MIN; in_order = true; }
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
        if (p < last_pos) in_order = false;
        last_pos = p;
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
            if (r.pos < last_pos) in_order = false;
            last_pos = r.pos;
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
        if (p < last_pos) in_order = false;
        last_pos = p;
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
    bool used_fast = false, used_device_decode = false;
    ChunkPool chunk_pool;                                // pinned buffers of the device-decode path (bgzf_feed.hpp)
    bool hint_set = false; uint32_t hint_min_dep = 1, hint_mask = 0xFFFFFFFFu;   // reduce hint to restore after a pd_reset
    PinnedPool *pool = nullptr;                          // pinned staging (exists once the CUDA context does)
    const std::vector<Target> *targets = nullptr;        // contig table (names, lengths)
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
        // a file whose header says SO:coordinate but whose records are not in order (the reference does not care) must not
        // reach the tile kernel as PD_SORTED: the claim is made per batch, from what the batch actually holds
        sorted = sorted && b.in_order;
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

    // The reference's no-index pass over a coordinate-sorted file (:4604-4675, list mode :2859-2945) does not accumulate every
    // read: per contig it keeps a cursor into the merged regions, skips reads that end before the cursor's region
    // (`bam_endpos`, which is pos+1 for a read flagged unmapped, whatever its CIGAR says), moves the cursor past regions the
    // read starts beyond, gives up on a contig after its last region and stops reading once every contig is finished.
    // Which reads reach the depth array therefore depends on that cursor; it is restated here so that the same reads do.
    struct RegionCursor {
        const std::map<int, std::vector<std::pair<int, int>>> &merged;
        std::vector<const std::vector<std::pair<int, int>> *> regs;   // per tid; nullptr = no regions (EndChr starts true)
        std::vector<size_t> at;
        std::vector<char> done;
        size_t open_contigs = 0;
        // mode 5: the reference's merged regions are the -w tiles themselves (adjacent tiles do not merge); they stay implicit
        // here (pd_host.hpp n_tiles): region k of contig t is [1 + k*W, min(k*W + W, len)]
        long long tile_win = 0; std::vector<long long> tlen;
        RegionCursor(const std::map<int, std::vector<std::pair<int, int>>> &m, size_t n_targets, long long tile_w, const std::vector<Target> *targets)
            : merged(m), regs(n_targets, nullptr), at(n_targets, 0), done(n_targets, 1), tile_win(tile_w)
        {
            for (const auto &kv : merged) if (kv.first >= 0 && (size_t)kv.first < n_targets && !kv.second.empty()) { regs[(size_t)kv.first] = &kv.second; done[(size_t)kv.first] = 0; open_contigs++; }
            if (tile_win && targets) for (const auto &t : *targets) tlen.push_back((long long)t.len);
        }
        bool contig_done(int t) const { return done[(size_t)t] != 0; }
        size_t n_regions(int t) const { return tile_win ? (size_t)n_tiles(tlen[(size_t)t], tile_win) : regs[(size_t)t]->size(); }
        long long first(int t, size_t k) const { return tile_win ? 1 + (long long)k * tile_win : (*regs[(size_t)t])[k].first; }
        long long second(int t, size_t k) const { return tile_win ? std::min<long long>((long long)k * tile_win + tile_win, tlen[(size_t)t]) : (*regs[(size_t)t])[k].second; }
        // 0 = skip the read, 1 = accumulate it, 2 = every contig is finished: stop reading (this read is not accumulated)
        int step(int t, int64_t pos, int64_t endpos)
        {
            size_t &k = at[(size_t)t];
            const size_t n = n_regions(t);
            if (endpos < first(t, k)) return 0;
            if (pos > second(t, k)) {
                for (++k; k < n; ++k) if (pos <= second(t, k)) break;
                if (k == n) { done[(size_t)t] = 1; if (--open_contigs == 0) return 2; }
            }
            return 1;
        }
    };

    // Whole-genome modes, BAM input: the compressed file goes to the engine piece by piece (bgzf_feed.hpp -> pd_push_bgzf:
    // inflate + record walk + accumulate on the device).  Returns 1 = the whole file was accumulated this way,
    // 0 = not applicable or the engine asked for the host path (nothing of this file is accumulated: the engine was reset if
    // needed), -1 = failed.
    int stream_bgzf(bool strict_eof)
    {
        if (dry) return 0;
        eng_future.wait();
        {
            std::lock_guard<std::mutex> g(eng_mu);
            if (!eng) { eng = eng_future.get(); if (!eng) { fail(std::string("Error: ") + pd_last_error(nullptr)); return -1; } }
            eng_ok = true;
        }
        const auto t0 = std::chrono::steady_clock::now();
        // one run per process: bound the pinned pool (PANDEPTH_PINNED_MB, default 512 MB; 0 = unbounded)
        { const char *pm = getenv("PANDEPTH_PINNED_MB"); chunk_pool.max_bytes = (size_t)((pm ? atof(pm) : 512.0) * 1048576.0); }
        const BgzfFeedResult r = bgzf_feed_file(eng, &eng_mu, input, std::max(1, o.threads), chunk_pool, strict_eof);
        ns_decode += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        n_push += r.pieces;
        if (getenv("PANDEPTH_TIMING"))
            fprintf(stderr, "[timing] bgzf feed: %llu pieces of %.1f MB, %d readers; thread-seconds: pinned alloc/wait %.3f, pread %.3f, block tables %.3f, pd_push_bgzf %.3f; final pd_sync %.3f s\n",
                    (unsigned long long)r.pieces, r.piece_bytes / 1048576.0, r.threads, r.s_alloc, r.s_read, r.s_chain, r.s_push, r.s_sync);
        if (r.status < 0) { fail(r.err.empty() ? "Error: device decode failed on " + input : r.err); return -1; }
        if (r.status == 0) {
            // (records that do not start on block boundaries, CG-tag CIGARs.)  Depth is additive but the pieces already
            // accumulated cannot be taken back one by one, so this file starts over on the host path from a clean engine;
            // only valid while this is the first input file
            std::lock_guard<std::mutex> g(eng_mu);
            if (r.pieces && pd_reset(eng) != PD_OK) { fail(std::string("pd_reset: ") + pd_last_error(eng)); return -1; }
            if (r.pieces && hint_set && pd_set_reduce_hint(eng, hint_min_dep, hint_mask) != PD_OK) { fail(std::string("pd_set_reduce_hint: ") + pd_last_error(eng)); return -1; }
            return 0;
        }
        pd_stats st;
        if (pd_get_stats(eng, &st) == PD_OK) reads_pushed = st.reads_pushed;
        used_fast = true; used_device_decode = true;
        return 1;
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
        RegionCursor cursor(merged, n_targets, o.tile_win, targets);
        const bool ok = sc.scan_all([&](const BamStreamScanner::Read &r) -> bool {
            if (!ok_push) return false;
            const int t = r.tid;
            if (t < 0 || (size_t)t >= n_targets) return true;
            if (sorted && cursor.contig_done(t)) return true;              // :4610
            if ((int)r.mapq < mq) return true;
            if (r.flag & xf) return true;
            if (sorted) {
                int64_t rlen = 0;                                         // bam_endpos(): unmapped or no reference-consuming op -> pos + 1
                if (!(r.flag & 4)) for (uint32_t c = 0; c < r.n_cigar; c++) { const uint32_t wd = r.word(c); if ((0x18du >> (wd & 15u)) & 1u) rlen += wd >> 4; }
                const int act = cursor.step(t, r.pos, (int64_t)r.pos + (rlen ? rlen : 1));
                if (act == 0) return true;
                if (act == 2) return false;
            }
            if (t != cur) { ok_push = push(cur, batch, sorted); cur = t; if (!ok_push) return false; }
            batch.add_raw(r);
            if (batch.full()) ok_push = push(cur, batch, sorted);
            return ok_push;
        });
        if (!ok) { fail("Error: malformed BAM while reading " + input + ": " + sc.error()); return true; }
        if (ok_push && !failed) push(cur, batch, sorted);
        return true;
    }

    // no-index path (:4604-4713): one pass over the file, batches split at contig changes
    bool stream_all(bool sorted, size_t n_targets)
    {
        sam_hdr_t *hdr = nullptr;
        htsFile *fp = open_input(&hdr, o.threads);
        if (!fp) { fail("Error: Failed to open the BAM/CRAM file: " + input); return false; }
        bam1_t *b = bam_init1();
        Batch batch;
        int cur = -1;
        RegionCursor cursor(merged, n_targets, o.tile_win, targets);
        while (sam_read1(fp, hdr, b) >= 0) {
            const int t = b->core.tid;
            if (t < 0 || (size_t)t >= n_targets) continue;
            if (sorted && cursor.contig_done(t)) continue;                 // :4610
            if ((int)b->core.qual < o.min_mapq) continue;
            if (b->core.flag & o.xflags) continue;
            if (sorted) {
                const int act = cursor.step(t, b->core.pos, bam_endpos(b));
                if (act == 0) continue;
                if (act == 2) break;
            }
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

}  // namespace pdh
