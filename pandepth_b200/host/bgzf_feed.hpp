// bgzf_feed.hpp -- whole-file feed of a BAM through the engine's device decoder (pd_push_bgzf).
//
// Role in the reference: the `sam_read1` / `sam_itr_next` loops of ProDealChrBambai and the no-index pass (PanDepth.cpp:725-734,
// :4608) -- here the host only moves compressed bytes: N threads pread fixed-size pieces of the file into pinned buffers in
// parallel, the BGZF block table of a piece is cut by hopping its block headers (a piece starts where the previous piece's
// last block ended: a short sequential hand-over between otherwise independent readers), and each piece goes to
// pd_push_bgzf(PD_ASYNC), where the GPU inflates the blocks, walks the records and accumulates them.
// Used by the command line (feeder.hpp) and by pdh_feed_bam (include/pandepth_host.h, what bench.py's e2e arm calls).
#pragma once
#include "fastbam.hpp"
#include "../../include/pandepth_b200.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace pdh {

// Pinned buffers for compressed pieces; kept for the life of the process (a second file or bench step reuses them).
struct ChunkPool {
    // Upper bound on the pinned bytes of this pool (0 = none).  Pinned memory costs ~0.3 ms per MB to allocate and about as much
    // again when the process exits -- paid on every run of the command line (measured: 23 x 81 MB for a 5 GB BAM with SEQ+QUAL
    // = 0.6 s + ~1.5 s of a 3.75 s run), and once in a long-lived caller (bench.py's end-to-end arm, a service).
    size_t max_bytes = 0;
    struct Buf { uint8_t *p = nullptr; size_t cap = 0; };
    std::vector<Buf> bufs; std::vector<int> free_ids;
    std::mutex mu; std::condition_variable cv;
    size_t max_bufs = 0, buf_bytes = 0;
    void configure(size_t n, size_t bytes)
    {
        std::lock_guard<std::mutex> g(mu);
        if (bytes > buf_bytes) {                             // bigger pieces than before: start over
            for (auto &b : bufs) pd_host_free(b.p);
            bufs.clear(); free_ids.clear(); buf_bytes = bytes;
        }
        max_bufs = std::max(max_bufs, n);
    }
    int acquire()                                            // a free buffer, a new one while the pool may grow, else wait
    {
        std::unique_lock<std::mutex> g(mu);
        for (;;) {
            if (!free_ids.empty()) { const int id = free_ids.back(); free_ids.pop_back(); return id; }
            if (bufs.size() < max_bufs) {
                Buf b; b.cap = buf_bytes;
                g.unlock();
                b.p = (uint8_t *)pd_host_alloc(b.cap);
                g.lock();
                if (!b.p) return -1;
                bufs.push_back(b);
                return (int)bufs.size() - 1;
            }
            cv.wait(g);
        }
    }
    void release(int id) { { std::lock_guard<std::mutex> g(mu); free_ids.push_back(id); } cv.notify_one(); }
    uint8_t *ptr(int id) { std::lock_guard<std::mutex> g(mu); return bufs[(size_t)id].p; }
    ~ChunkPool() { for (auto &b : bufs) pd_host_free(b.p); }
};

struct BgzfFeedResult {
    int status = 0; std::string err; uint64_t bytes = 0, pieces = 0, blocks = 0;   // status 1 = whole file fed, 0 = host path needed, -1 = failed
    double s_alloc = 0, s_read = 0, s_chain = 0, s_push = 0, s_sync = 0;           // thread-seconds: pinned buffers, pread, block-table hand-over + hop, pd_push_bgzf, final pd_sync
    size_t piece_bytes = 0; int threads = 0;
};

// Feeds every record of `path` (a BAM in BGZF blocks) to `eng`.  `eng_mu` (nullable) serialises engine calls with other users.
// strict_eof: a file that ends inside a BGZF block is an error (what the indexed host scanner reports); otherwise the stream
// ends quietly at the last complete block, like the reference's `while (sam_read1(...) >= 0)` loop (:4608).
// [range_beg, range_end): only the BGZF blocks that START in this byte range of the file are fed (range sharding over GPUs: the
// offsets come from the .bai, see pdh_bai_block_offset); range_beg = 0 means "from the first record".
// stop_tid / stop_pos (range sharding, coordinate-sorted files): stop pushing once the engine reports a block whose first
// record lies at or beyond (stop_tid, stop_pos) -- every read that starts before that position has been pushed by then.
inline BgzfFeedResult bgzf_feed_file(pd_engine *eng, std::mutex *eng_mu, const std::string &path, int threads, ChunkPool &pool, bool strict_eof = true,
                                     uint64_t range_beg = 0, uint64_t range_end = UINT64_MAX, int32_t stop_tid = INT32_MAX, int32_t stop_pos = INT32_MAX)
{
    BgzfFeedResult res;
    uint32_t first_skip = 0;
    BgzfChunker ch;
    if (!ch.open(path, first_skip)) { res.status = 0; return res; }            // not a BAM this path understands
    uint64_t data_start = ch.position();
    if (range_beg > data_start) { data_start = range_beg; first_skip = 0; }    // a block boundary inside the file: records start at offset 0
    const int fd = ::open(path.c_str(), O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) { if (fd >= 0) close(fd); res.status = -1; res.err = "cannot open " + path; return res; }
    const uint64_t true_size = (uint64_t)st.st_size;
    const uint64_t file_size = std::min<uint64_t>(true_size, range_end);      // blocks starting at or after this offset belong to someone else
    // piece size from the compression ratio of the first blocks: ~384 MB inflated, at most 96 MB compressed.  Measured on B200 through
    // this feeder (profiles/r02_bgzf_sweeps.txt): flavour A (10^9 reads, 2.1 GB) 9 / 18 / 37 / 74 MB pieces -> 294 / 288 / 298 / 311 ms;
    // flavour B (33 M reads, 5.4 GB) 48 / 96 / 192 MB -> 182 / 177 / 204 ms
    double ratio = 4.0;
    {
        std::vector<uint8_t> probe(1u << 20);
        const ssize_t k = pread(fd, probe.data(), probe.size(), (off_t)data_start);
        uint64_t in = 0, out = 0; size_t p = 0;
        while (k > 0 && p + 18 <= (size_t)k) {
            const uint8_t *cd; size_t cl; uint32_t is;
            const int64_t b = bgzf_block_at(probe.data() + p, (size_t)k - p, &cd, &cl, &is);
            if (b <= 0) break;
            in += (uint64_t)b; out += is; p += (size_t)b;
        }
        if (in && out) ratio = (double)out / (double)in;
    }
    size_t piece = (size_t)(384.0 * 1048576.0 / std::max(1.0, ratio));
    if (const char *e = getenv("PANDEPTH_BGZF_PIECE_MB")) piece = (size_t)(atof(e) * 1048576.0);
    piece = std::max<size_t>(1u << 20, std::min<size_t>(96u << 20, piece));
    const uint64_t span = file_size > data_start ? file_size - data_start : 0;
    // small inputs: more, smaller pieces (pinned memory costs ~0.3 ms per MB to allocate and as much again to release at exit)
    if (!getenv("PANDEPTH_BGZF_PIECE_MB")) piece = std::min<size_t>(piece, std::max<size_t>(8u << 20, (size_t)(span / (4 * (uint64_t)std::max(1, threads)))));
    if (pool.max_bytes && !getenv("PANDEPTH_BGZF_PIECE_MB"))   // the pool holds threads + PD_BGZF_DEPTH + 1 buffers of piece + 68 KB
        piece = std::min<size_t>(piece, std::max<size_t>(4u << 20, pool.max_bytes / (size_t)(std::max(1, threads) + PD_BGZF_DEPTH + 1)));
    piece = std::max<size_t>(piece, 1u << 17);
    const size_t n_pieces = (size_t)((span + piece - 1) / piece);
    if (n_pieces == 0) { close(fd); res.status = 1; return res; }
    const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), n_pieces));
    const size_t buf_bytes = piece + (1u << 16) + 4096;
    pool.configure((size_t)T + PD_BGZF_DEPTH + 1, buf_bytes);

    std::mutex own_mu;
    std::mutex &emu = eng_mu ? *eng_mu : own_mu;
    std::mutex chain_mu; std::condition_variable chain_cv;
    std::vector<uint64_t> start(n_pieces + 1, UINT64_MAX);                      // file offset of the first block that starts in piece j
    start[0] = data_start;
    std::atomic<size_t> next{0};
    std::atomic<int> state{1};                                                  // 1 running, 0 fallback, -1 error
    std::string err;
    std::deque<std::pair<uint64_t, int>> in_flight;                             // (push sequence, buffer) still read by the engine's copies
    uint64_t push_seq = 0;
    auto set_state = [&](int s, const std::string &m) {
        std::lock_guard<std::mutex> g(chain_mu);
        if (state.load() == 1) { state = s; err = m; }
        chain_cv.notify_all();
    };
    res.piece_bytes = piece; res.threads = T;
    std::mutex stat_mu;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    auto worker = [&] {
        std::vector<pd_bgzf_block> blocks;
        double t_alloc = 0, t_read = 0, t_chain = 0, t_push = 0;
        struct Fold { double &a, &r, &c, &p; BgzfFeedResult &res; std::mutex &mu; ~Fold() { std::lock_guard<std::mutex> g(mu); res.s_alloc += a; res.s_read += r; res.s_chain += c; res.s_push += p; } } fold{t_alloc, t_read, t_chain, t_push, res, stat_mu};
        for (;;) {
            const size_t j = next.fetch_add(1);
            if (j >= n_pieces || state.load() != 1) break;
            const uint64_t a = data_start + (uint64_t)j * piece, z = std::min<uint64_t>(file_size, a + piece);
            auto t0 = now();
            const int id = pool.acquire();
            t_alloc += secs(t0, now());
            if (id < 0) { set_state(-1, "out of pinned host memory"); break; }
            uint8_t *buf = pool.ptr(id);
            const size_t want = (size_t)std::min<uint64_t>(true_size - a, (uint64_t)piece + (1u << 16) + 64);   // the last block may end past the range
            size_t got = 0;
            t0 = now();
            while (got < want) {
                const ssize_t k = pread(fd, buf + got, want - got, (off_t)(a + got));
                if (k < 0) { set_state(-1, "read error on " + path); break; }
                if (k == 0) break;
                got += (size_t)k;
            }
            t_read += secs(t0, now());
            t0 = now();
            // my first block: handed over by the previous piece
            uint64_t s0;
            {
                std::unique_lock<std::mutex> g(chain_mu);
                chain_cv.wait(g, [&] { return start[j] != UINT64_MAX || state.load() != 1; });
                s0 = start[j];
            }
            if (state.load() != 1) { pool.release(id); break; }
            blocks.clear();
            uint64_t p = s0 - a;                                                // s0 >= a: a block is shorter than a piece
            bool bad = false, cut = false;
            while (a + p < z) {
                const uint8_t *cd; size_t cl; uint32_t is;
                const int64_t b = bgzf_block_at(buf + p, got > p ? got - (size_t)p : 0, &cd, &cl, &is);
                if (b < 0) { bad = true; break; }
                if (b == 0) { cut = true; p = z - a + (1u << 20); break; }      // the file ends inside this block: the stream ends here
                if (is) blocks.push_back(pd_bgzf_block{(uint64_t)(cd - buf), (uint32_t)cl, is});
                p += (uint64_t)b;
            }
            {
                std::lock_guard<std::mutex> g(chain_mu);
                start[j + 1] = a + p;
            }
            chain_cv.notify_all();
            if (bad) { set_state(-1, "Error: malformed BAM while reading " + path + ": not a BGZF block"); pool.release(id); break; }
            if (cut && strict_eof) { set_state(-1, "Error: malformed BAM while reading " + path + ": unexpected end of file inside a BGZF block"); pool.release(id); break; }
            if (blocks.empty()) { pool.release(id); continue; }
            t_chain += secs(t0, now());
            t0 = now();
            int rc;
            {
                std::lock_guard<std::mutex> g(emu);
                rc = state.load() == 1 ? pd_push_bgzf(eng, buf, got, blocks.data(), (int64_t)blocks.size(), j == 0 ? first_skip : 0u, PD_ASYNC | PD_CRC) : PD_OK;
                ++push_seq;
                in_flight.emplace_back(push_seq, id);
                // the engine has finished reading the bytes of a chunk once PD_BGZF_DEPTH later pushes have returned
                while (!in_flight.empty() && in_flight.front().first + PD_BGZF_DEPTH <= push_seq) { pool.release(in_flight.front().second); in_flight.pop_front(); }
                res.bytes += got; res.pieces++; res.blocks += blocks.size();
                if (rc == PD_OK && stop_tid != INT32_MAX) {
                    int32_t pt = -1, pp = -1;
                    pd_bgzf_progress(eng, &pt, &pp);
                    if (pt > stop_tid || (pt == stop_tid && pp >= stop_pos)) next.store(n_pieces);     // no further pieces: the rest belongs to others
                }
                if (rc != PD_OK) {
                    // this chunk or an earlier one: depth already accumulated cannot be taken back piecewise, the caller starts over
                    const std::string m = std::string("pd_push_bgzf: ") + pd_last_error(eng);
                    if (rc == PD_ERR_FALLBACK) set_state(0, m); else set_state(-1, m);
                }
            }
            t_push += secs(t0, now());
            if (rc != PD_OK) break;
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < T; i++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
    {
        std::lock_guard<std::mutex> g(emu);
        const auto t0 = now();
        struct S { double &d; std::chrono::steady_clock::time_point t; ~S() { d += std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); } } sy{res.s_sync, t0};
        if (state.load() == 1) {
            const int rc = pd_sync(eng);
            if (rc == PD_ERR_FALLBACK) set_state(0, std::string("pd_push_bgzf: ") + pd_last_error(eng));
            else if (rc != PD_OK) set_state(-1, std::string("pd_sync: ") + pd_last_error(eng));
        } else pd_sync(eng);                                                    // drain: the buffers go back to the pool
        for (auto &f : in_flight) pool.release(f.second);
        in_flight.clear();
    }
    close(fd);
    res.status = state.load(); res.err = err;
    return res;
}

}  // namespace pdh
