// pd_engine.cu -- the engine behind include/pandepth_b200.h (C ABI, CUDA runtime only; no torch types).
//
// HBM layout (one engine = one GPU):
//   depth   uint32[total]   all owned contigs back to back; contig t starts at off[t] (a multiple of 4096 elements) and is
//                           followed by >= 1 element of padding, so a run's closing -1 at `len` never lands in a neighbour
//   delta   uint32[total]   same geometry; +-1 run endpoints of the unsorted / far path; all-zero before the first batch;
//                           stale (consumed, not yet cleared) after a scan until the next batch or reset clears it
//   seg_state uint8[total/SEG]   per 2048-base segment: 0 = never written (first write is a plain store), 1 = written once and
//                           its CoveredSite/TotalDepth partial is valid (reduce hint), 2 = written
//   part_cov/part_sum       per-segment partials of the reduce hint (12 B per segment)
// 3 Gb genome: 12 GB + 12 GB + 19 MB of the 180 GB.
#include "../../include/pandepth_b200.h"
#include "pd_kernels.cuh"
#include "pd_bgzf.cuh"
#include "pd_bucket.cuh"

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges show up in nsys / ncu timelines, no-ops without a tool attached

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {

std::string g_create_error;

constexpr uint64_t PAD_ELEMS = pdk::SCAN_TILE > pdk::REDUCE_CHUNK ? pdk::SCAN_TILE : pdk::REDUCE_CHUNK;   // contig offsets / total are multiples of this
static_assert(PAD_ELEMS % pdk::SCAN_TILE == 0 && PAD_ELEMS % pdk::REDUCE_CHUNK == 0 && PAD_ELEMS % 4096 == 0,
              "padding must hold whole scan tiles, and a contig's reduce chunks must start on 4-segment boundaries of the whole array");
constexpr int64_t HOST_CHUNK_READS = 8 << 20;            // staging chunk for host-resident pushes
constexpr int N_SLOTS = 2;

enum KernelId { K_CUT = 0, K_TILE, K_SCATTER, K_SCAN, K_FILL, K_REDUCE, K_WINDOWS, K_BGZF, K_BGZF_PACK, K_FILTER, K_BUCKET, K_COUNT };
const char *const kKernelNames[K_COUNT] = {"k_cut", "k_tile_accumulate", "k_delta_scatter", "k_scan_add",
                                           "k_fill_untouched", "k_reduce_units", "k_reduce_windows", "k_bgzf_decode", "k_bgzf_pack", "k_filter_bits", "k_bucket"};

struct Slot {
    int32_t *pos = nullptr; uint32_t *cigar = nullptr; uint32_t *cigar_off = nullptr;
    uint16_t *flag = nullptr; uint8_t *mapq = nullptr;
    size_t cap_reads = 0, cap_cigar = 0;
    cudaEvent_t copied = nullptr, consumed = nullptr;
    bool used = false;
};

// One chunk of BGZF blocks in flight through the device decoder (pd_push_bgzf): compressed bytes in, dense record arrays out.
struct BzSlot {
    uint8_t *d_in = nullptr; size_t in_cap = 0;
    uint8_t *d_out = nullptr; size_t out_cap = 0;                 // inflated bytes
    pdk::BgzfBlockDesc *d_blk = nullptr; pdk::BgzfSummary *d_sum = nullptr; uint32_t *d_r0 = nullptr, *d_c0 = nullptr; size_t blk_cap = 0;
    int32_t *st_pos = nullptr, *st_tid = nullptr; uint32_t *st_coff = nullptr; size_t st_rec_cap = 0;
    uint32_t *st_cig = nullptr; size_t st_cig_cap = 0;
    int32_t *pos = nullptr, *tid = nullptr; uint32_t *coff = nullptr; size_t rec_cap = 0;
    uint32_t *cig = nullptr; size_t cig_cap = 0;
    uint32_t *d_totals = nullptr; uint32_t *d_ticket = nullptr;
    pdk::BgzfBlockDesc *h_blk = nullptr; pdk::BgzfSummary *h_sum = nullptr; uint32_t *h_totals = nullptr; size_t h_blk_cap = 0;   // pinned
    cudaStream_t stream = nullptr;                                // decode + pack of this slot (slots decode concurrently)
    cudaEvent_t copied = nullptr, in_free = nullptr, decoded = nullptr, consumed = nullptr;
    bool in_used = false, out_used = false, pending = false;
    uint32_t n_blocks = 0;
};
constexpr int N_BZ = PD_BGZF_DEPTH + 1;

}  // namespace

struct pd_engine {
    int device = 0;
    int n_targets = 0;
    std::vector<uint32_t> tlen;      // contig lengths (the header's)
    std::vector<uint32_t> org, elen; // owned range of contig t: [org[t], org[t] + elen[t]) (range sharding; whole contig: 0, tlen[t])
    bool has_ranges = false;         // some contig is owned only in part
    std::vector<uint8_t> owned;
    std::vector<uint64_t> off;       // element offset of contig t (valid when owned)
    uint64_t total = 0;              // elements in depth/delta
    uint64_t n_seg = 0;
    uint32_t xflags = 1796; int min_mapq = -1;

    uint32_t *d_depth = nullptr, *d_delta = nullptr;
    uint8_t *d_seg_state = nullptr;
    uint64_t *d_off = nullptr; uint32_t *d_tlen = nullptr, *d_org = nullptr;     // d_tlen holds the OWNED lengths
    // per-batch cut arrays, double-buffered: the cut kernels of batch k+1 run on cut_stream under batch k's tile kernel
    struct CutBuf { uint32_t *lo = nullptr, *hi = nullptr, *coarse = nullptr; size_t cap = 0; cudaEvent_t ready = nullptr, released = nullptr; bool used = false; };
    CutBuf cut[2]; int cut_next = 0;
    cudaStream_t cut_stream = nullptr; cudaEvent_t cut_input = nullptr;
    int *d_flags = nullptr;          // [0] err bits, [1] delta_used
    uint32_t *d_part_cov = nullptr; unsigned long long *d_part_sum = nullptr;   // per-segment reduce partials (hint)
    bool hint_on = false; uint32_t hint_min_dep = 1, hint_mask = 0xFFFFFFFFu;
    bool pushed_since_reset = false;
    int32_t tile_resident_warps = 148 * 3 * pdk::WPC;   // warps of k_tile_accumulate resident on the device (prefetch distance)
    unsigned long long *d_desc = nullptr; unsigned int *d_counter = nullptr;

    cudaStream_t own_stream = nullptr, copy_stream = nullptr, stream = nullptr;
    Slot slot[N_SLOTS]; int next_slot = 0;
    BzSlot bz[N_BZ]; int bz_next = 0;
    int32_t bz_prog_tid = -1, bz_prog_pos = -1;   // first record of the last block (with records) of the last chunk that was planned
    uint8_t *d_owned = nullptr;
    // unsorted path at scale (pd_bucket.cuh): runs in two ping-pong arrays, coarse / fine bucket tables
    int32_t *bk_pos[2] = {nullptr, nullptr}; uint32_t *bk_cig[2] = {nullptr, nullptr}; size_t bk_cap = 0;
    uint32_t *bk_tab = nullptr;      // [0] run cursor, [1] overflow flag (unused), then hist[4096], base[4097], cursor[4096]
    uint32_t *bk_fine = nullptr; size_t bk_fine_cap = 0;
    uint64_t bucket_min_reads = 1u << 20;   // unsorted batches of at least this many reads go expand -> bucket -> tile (PANDEPTH_BUCKET_MIN_READS;
                                     // 0 = never).  Measured on B200, 111 M shuffled reads over 333 Mb (profiles/r02_delta_path.txt): scatter + scan
                                     // 7.85 + 0.48 ms; bucket passes placing runs straight into global memory 10.8-11.5 ms (+ 0.45 ms tile);
                                     // staged in shared memory and written out in bursts 4.89 ms (+ 0.45): 1.56x faster than the scatter
    uint32_t *d_pass_bits = nullptr; size_t bits_cap = 0;   // record filter of the current batch, one bit per read (k_filter_bits)
    int bz_grid = 148 * 4, bz_grid_k2 = 148 * 6, bz_grid_k4 = 148 * 6, bz_grid_k8 = 148 * 6;
    uint64_t scatter_stripe = 32ull << 20;          // elements (128 MB of deltas) per stripe of a striped scatter (PANDEPTH_SCATTER_STRIPE_MB)
    uint64_t scatter_stripe_min_reads = 4ull << 20; // batches smaller than this scatter in one launch (PANDEPTH_SCATTER_STRIPE_MIN_READS)
    int bz_force_k = 0;              // PANDEPTH_BGZF_K=1|2|4|8: decoders per warp (default: by compression ratio)
    int bz_lit_k = 2;                // decoders per warp for literal-dominated chunks (PANDEPTH_BGZF_LIT_K); measured on flavour B: K=1 116.1, K=2 89.7,
                                     // K=4 108.4 ms per 20 M reads (K > 2: too few resident warps to hide the symbol chain's latency)

    // reduce scratch
    void *d_red = nullptr; size_t red_cap = 0;
    uint32_t *d_todo = nullptr; size_t todo_cap = 0;   // pd_reduce with the hint: units that still need depth after the partials pass

    bool finalized = false;
    bool host_delta_used = false;    // a scatter-path batch was queued since the last scan
    bool delta_stale = false;        // the last scan consumed the deltas but left them in place: clear before the next writer
    pd_stats stats{};
    bool profiling = false;
    struct Ev { int k; cudaEvent_t a, b; };
    std::vector<Ev> pending;
    double kms[K_COUNT] = {0}; uint64_t kcount[K_COUNT] = {0};
    std::string err;
};

namespace {

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            char _b[512];                                                                             \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            e->err = _b;                                                                              \
            return PD_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

int fail(pd_engine *e, int code, const char *msg) { if (e) e->err = msg; return code; }

struct NvtxRange { explicit NvtxRange(const char *name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

template <typename F>
int launch_on(pd_engine *e, cudaStream_t st, int k, F &&f)
{
    cudaEvent_t a = nullptr, b = nullptr;
    if (e->profiling) {
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        CK(cudaEventRecord(a, st));
    }
    f();
    CK(cudaGetLastError());
    if (e->profiling) { CK(cudaEventRecord(b, st)); e->pending.push_back({k, a, b}); }
    e->stats.kernel_launches++;
    return PD_OK;
}
template <typename F>
int launch(pd_engine *e, int k, F &&f) { return launch_on(e, e->stream, k, static_cast<F &&>(f)); }

int drain_profile(pd_engine *e)
{
    for (auto &p : e->pending) {
        CK(cudaEventSynchronize(p.b));
        float ms = 0; CK(cudaEventElapsedTime(&ms, p.a, p.b));
        e->kms[p.k] += ms; e->kcount[p.k]++;
        cudaEventDestroy(p.a); cudaEventDestroy(p.b);
    }
    e->pending.clear();
    return PD_OK;
}

int ensure_cut(pd_engine *e, pd_engine::CutBuf &c, size_t nseg)
{
    if (!c.ready) { CK(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c.released, cudaEventDisableTiming)); }
    if (nseg <= c.cap) return PD_OK;
    if (c.lo) { CK(cudaStreamSynchronize(e->stream)); CK(cudaStreamSynchronize(e->cut_stream)); cudaFree(c.lo); cudaFree(c.hi); cudaFree(c.coarse); c.lo = c.hi = c.coarse = nullptr; c.cap = 0; }
    size_t cap = nseg + nseg / 4 + 1024;
    CK(cudaMalloc(&c.lo, cap * 4)); CK(cudaMalloc(&c.hi, cap * 4)); CK(cudaMalloc(&c.coarse, (cap / pdk::CUT_COARSE + 2) * 4));
    c.cap = cap;
    return PD_OK;
}

// Queue the kernels for one contig batch whose arrays are on the device.  seg range [seg_a, seg_b].
// input_ready: the event after which the batch's arrays are complete (a copy / decode on another stream); nullptr = they are
// ordered by the main stream (a caller's device batch), unless ready_now says they are complete already (PD_READY).
int run_batch(pd_engine *e, int32_t tid, uint32_t n, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
              const uint32_t *cigar_off, const uint32_t *cigar, bool sorted, int64_t seg_a, int64_t seg_b,
              cudaEvent_t input_ready = nullptr, bool ready_now = false, uint64_t n_cigar_words = 0)
{
    const uint64_t off = e->off[tid];
    if (e->delta_stale) {            // a batch after a scan (rare): both paths may write deltas, so the consumed ones go first
        CK(cudaMemsetAsync(e->d_delta, 0, e->total * 4, e->stream));
        e->delta_stale = false;
    }
    if (sorted) {
        const int32_t nseg = (int32_t)(seg_b - seg_a + 1);
        // The cut kernels (binary searches: pure latency, a handful of CTAs) run on their own stream: they only need the batch's
        // positions (whatever the main stream had queued when this call was made) and a free pair of cut arrays, so they
        // overlap the previous batch's tile kernel instead of sitting between two tile kernels.
        pd_engine::CutBuf &cb = e->cut[e->cut_next]; e->cut_next ^= 1;
        int rc = ensure_cut(e, cb, (size_t)nseg + 1); if (rc) return rc;
        pdk::CutArgs c{pos, n, (int32_t)seg_a, nseg, cb.lo, cb.hi, (int64_t)e->org[tid]};
        const int32_t ncoarse = (nseg + pdk::CUT_COARSE - 1) / pdk::CUT_COARSE;
        // a batch whose arrays are ordered by the main stream cannot be cut before that stream gets there anyway: its cut
        // kernels run in line (two cross-stream hops per batch saved); the side stream is for batches that are complete already
        const bool side = input_ready != nullptr || ready_now;
        cudaStream_t cs = side ? e->cut_stream : e->stream;
        if (input_ready) CK(cudaStreamWaitEvent(e->cut_stream, input_ready, 0));
        if (side && cb.used) CK(cudaStreamWaitEvent(e->cut_stream, cb.released, 0));      // the tile kernel that last read these arrays
        rc = launch_on(e, cs, K_CUT, [&] {
            pdk::k_cut_coarse<<<(ncoarse + pdk::CUT_THREADS - 1) / pdk::CUT_THREADS, pdk::CUT_THREADS, 0, cs>>>(c, cb.coarse, ncoarse);
            pdk::k_cut<<<(nseg + 1 + pdk::CUT_THREADS - 1) / pdk::CUT_THREADS, pdk::CUT_THREADS, 0, cs>>>(c, cb.coarse, ncoarse);
        });
        if (rc) return rc;
        e->stats.kernel_launches++;
        if (side) { CK(cudaEventRecord(cb.ready, e->cut_stream)); CK(cudaStreamWaitEvent(e->stream, cb.ready, 0)); }
        if (e->org[tid] != 0 && seg_a == 0) {
            // range sharding: reads that start more than HALO bases left of the owned range are seen by no segment; the ones
            // among them that reach into the range are the first lo[0] reads of the batch (a device-side count, usually 0)
            pdk::ScatterArgs sp{pos, cigar, cigar_off, flag, mapq, n, (int32_t)e->elen[tid], e->d_delta + off, e->xflags, e->min_mapq,
                                (int32_t)e->org[tid], cb.lo, e->d_flags + 1, 0, INT64_MAX};
            rc = launch(e, K_SCATTER, [&] { pdk::k_delta_scatter<<<148, 256, 0, e->stream>>>(sp); }); if (rc) return rc;
        }
        const bool has_off = cigar_off != nullptr, has_filter = flag != nullptr || mapq != nullptr;
        if (has_filter) {
            // the record filter as one bit per read (the tile kernel reads a word per 32-read chunk)
            const size_t n_words = (size_t)n / 32 + 34;
            if (n_words > e->bits_cap) {
                if (e->d_pass_bits) { CK(cudaStreamSynchronize(e->stream)); cudaFree(e->d_pass_bits); e->d_pass_bits = nullptr; e->bits_cap = 0; }
                const size_t cap = n_words + n_words / 4 + 1024;
                CK(cudaMalloc(&e->d_pass_bits, cap * 4)); e->bits_cap = cap;
            }
            rc = launch(e, K_FILTER, [&] { pdk::k_filter_bits<<<(unsigned)((n_words * 32 + 255) / 256), 256, 0, e->stream>>>(flag, mapq, n, e->xflags, e->min_mapq, e->d_pass_bits, (uint32_t)n_words); });
            if (rc) return rc;
        }
        pdk::TileArgs t{pos, cigar, cigar_off, flag, mapq, e->d_pass_bits, cb.lo, cb.hi, n, (int32_t)seg_a, nseg, (int32_t)e->elen[tid],
                        e->d_depth + off, e->d_delta + off, e->d_seg_state + off / pdk::SEG, e->xflags, e->min_mapq,
                        e->d_flags, e->d_flags + 1,
                        e->hint_on ? 1u : 0u, e->hint_min_dep, e->hint_mask, e->d_part_cov + off / pdk::SEG, e->d_part_sum + off / pdk::SEG,
                        e->tile_resident_warps, (int32_t)e->org[tid], 0u, nullptr};
        const size_t smem = (size_t)pdk::WPC * (pdk::TILE_SMEM_WORDS + (has_off ? pdk::TILE_QUEUE_WORDS : 0)) * 4;
        const unsigned grid = (nseg + pdk::WPC - 1) / pdk::WPC, blk = pdk::WPC * 32;
        rc = launch(e, K_TILE, [&] {
            if (has_off) { if (has_filter) pdk::k_tile_accumulate<true, true><<<grid, blk, smem, e->stream>>>(t); else pdk::k_tile_accumulate<true, false><<<grid, blk, smem, e->stream>>>(t); }
            else { if (has_filter) pdk::k_tile_accumulate<false, true><<<grid, blk, smem, e->stream>>>(t); else pdk::k_tile_accumulate<false, false><<<grid, blk, smem, e->stream>>>(t); }
        });
        if (rc) return rc;
        CK(cudaEventRecord(cb.released, e->stream)); cb.used = true;
        e->stats.tile_path_batches++;
    } else if (e->bucket_min_reads && (uint64_t)n >= e->bucket_min_reads) {
        // ---- expand -> bucket -> tile (pd_bucket.cuh): the batch is brought into 256-base bucket order on the device and
        // accumulated by the tile kernel in shared memory instead of 2 n global reductions at random addresses ----
        uint64_t cap = n;
        if (cigar_off) {
            cap = n_cigar_words;
            if (!cap) {                                          // a device batch: the number of CIGAR words is cigar_off[n]
                uint32_t last = 0;
                CK(cudaMemcpyAsync(&last, cigar_off + n, 4, cudaMemcpyDeviceToHost, e->stream));
                CK(cudaStreamSynchronize(e->stream));
                uint32_t first = 0;
                CK(cudaMemcpyAsync(&first, cigar_off, 4, cudaMemcpyDeviceToHost, e->stream));
                CK(cudaStreamSynchronize(e->stream));
                cap = (uint64_t)last - first;
            }
        }
        if (cap == 0) { e->stats.delta_path_batches++; e->stats.reads_pushed += n; e->finalized = false; e->pushed_since_reset = true; return PD_OK; }   // no CIGAR word, no run
        if (cap > 0xffff0000ull) return fail(e, PD_ERR_ARG, "pd_push_contig: an unsorted batch with more than 2^32 - 2^16 CIGAR words; push smaller batches");   // (tile arithmetic of the bucket passes is 32-bit)
        if (cap > e->bk_cap) {
            CK(cudaStreamSynchronize(e->stream));
            for (int k = 0; k < 2; k++) { cudaFree(e->bk_pos[k]); cudaFree(e->bk_cig[k]); e->bk_pos[k] = nullptr; e->bk_cig[k] = nullptr; }
            e->bk_cap = 0;                                       // (a failed allocation below must not leave a capacity over null pointers)
            const size_t c = (size_t)cap + (size_t)cap / 8 + 4096;
            for (int k = 0; k < 2; k++) { CK(cudaMalloc(&e->bk_pos[k], c * 4)); CK(cudaMalloc(&e->bk_cig[k], c * 4)); }
            e->bk_cap = c; e->stats.hbm_bytes_allocated += c * 16;
        }
        const int n_coarse = (int)(((uint64_t)e->elen[tid] + (1u << pdk::BK_COARSE_SHIFT) - 1) >> pdk::BK_COARSE_SHIFT);
        const size_t n_fine = (size_t)n_coarse * pdk::BK_FINE;
        if (!e->bk_tab) CK(cudaMalloc(&e->bk_tab, (2 + 3 * (pdk::BK_MAX_COARSE + 1)) * 4));
        if (n_fine > e->bk_fine_cap) {
            if (e->bk_fine) { CK(cudaStreamSynchronize(e->stream)); cudaFree(e->bk_fine); e->bk_fine = nullptr; e->bk_fine_cap = 0; }
            CK(cudaMalloc(&e->bk_fine, n_fine * 4)); e->bk_fine_cap = n_fine;
        }
        uint32_t *d_hist = e->bk_tab + 2, *d_base = d_hist + (pdk::BK_MAX_COARSE + 1), *d_cursor = d_base + (pdk::BK_MAX_COARSE + 1);
        const uint32_t *d_nruns = d_base + n_coarse;             // total number of runs = base[n_coarse], known on the device only
        CK(cudaMemsetAsync(e->bk_tab, 0, (2 + (pdk::BK_MAX_COARSE + 1)) * 4, e->stream));      // hist
        const int32_t nseg = (int32_t)(e->elen[tid] / pdk::SEG) + 1;
        pd_engine::CutBuf &cb = e->cut[e->cut_next]; e->cut_next ^= 1;
        int rc = ensure_cut(e, cb, (size_t)nseg + 1); if (rc) return rc;
        if (cb.used) CK(cudaStreamWaitEvent(e->stream, cb.released, 0));
        pdk::ExpandArgs x{pos, cigar, cigar_off, flag, mapq, n, (int32_t)e->elen[tid], (int32_t)e->org[tid], e->xflags, e->min_mapq,
                          e->bk_pos[0], e->bk_cig[0], (uint32_t)cap, e->d_delta + off, e->d_flags + 1};
        const unsigned tiles = (unsigned)((cap + pdk::BK_TILE - 1) / pdk::BK_TILE);
        rc = launch(e, K_BUCKET, [&] {
            pdk::k_expand_runs<<<(n + 255) / 256, 256, 0, e->stream>>>(x);
            pdk::k_bucket_hist<<<tiles, 256, 0, e->stream>>>(e->bk_pos[0], (uint32_t)cap, d_hist, n_coarse);
            pdk::k_bucket_scan<<<1, 1024, 0, e->stream>>>(d_hist, d_base, d_cursor, n_coarse);
            pdk::k_bucket_scatter<<<(unsigned)((cap + pdk::BK_STAGE - 1) / pdk::BK_STAGE), 256, pdk::BK_SCATTER_SMEM, e->stream>>>(e->bk_pos[0], e->bk_cig[0], (uint32_t)cap, d_base, d_cursor, n_coarse, e->bk_pos[1], e->bk_cig[1]);
            pdk::k_bucket_fine<<<(unsigned)n_coarse, 512, pdk::BK_FINE_SMEM, e->stream>>>(e->bk_pos[1], e->bk_cig[1], d_base, e->bk_fine, e->bk_pos[0], e->bk_cig[0]);
            pdk::k_bucket_cuts<<<(unsigned)((nseg + 1 + 255) / 256), 256, 0, e->stream>>>(e->bk_fine, d_nruns, (uint32_t)n_fine, 0, nseg, cb.lo, cb.hi);
        });
        if (rc) return rc;
        e->stats.kernel_launches += 5;
        pdk::TileArgs t{e->bk_pos[0], e->bk_cig[0], nullptr, nullptr, nullptr, nullptr, cb.lo, cb.hi, (uint32_t)cap, 0, nseg, (int32_t)e->elen[tid],
                        e->d_depth + off, e->d_delta + off, e->d_seg_state + off / pdk::SEG, e->xflags, e->min_mapq,
                        e->d_flags, e->d_flags + 1,
                        e->hint_on ? 1u : 0u, e->hint_min_dep, e->hint_mask, e->d_part_cov + off / pdk::SEG, e->d_part_sum + off / pdk::SEG,
                        e->tile_resident_warps, 0, 1u, d_nruns};
        const size_t smem = (size_t)pdk::WPC * pdk::TILE_SMEM_WORDS * 4;
        rc = launch(e, K_TILE, [&] { pdk::k_tile_accumulate<false, false><<<(nseg + pdk::WPC - 1) / pdk::WPC, pdk::WPC * 32, smem, e->stream>>>(t); });
        if (rc) return rc;
        CK(cudaEventRecord(cb.released, e->stream)); cb.used = true;
        e->stats.delta_path_batches++;
    } else {
        pdk::ScatterArgs s{pos, cigar, cigar_off, flag, mapq, n, (int32_t)e->elen[tid], e->d_delta + off, e->xflags, e->min_mapq,
                           (int32_t)e->org[tid], nullptr, nullptr, 0, INT64_MAX};
        // A big unsorted batch scatters 2 n run ends over the whole contig: every one a 32-byte read-modify-write in DRAM (the
        // delta array is ten times the L2).  Striped: one launch per stripe of the array applies only the ends that fall into it,
        // so each delta line goes to DRAM once; the price is re-reading the batch (8 B per read, streaming) once per stripe.
        // Measured on B200 (111 M shuffled reads over 333 Mb, tools/bench_delta_path.py): unstriped 9.48 ms; stripes of 32 / 64 /
        // 128 MB: 25.9 / 13.5 / 7.74 ms -- the reductions run at ~25 G/s whether their lines sit in L2 or not (the L2's atomic
        // units are the limit, not DRAM), so only a few large stripes pay: -18 %.  A real fix needs fewer global reductions,
        // i.e. sorted input for the tile kernel.
        const uint64_t n_stripes = ((uint64_t)e->elen[tid] + e->scatter_stripe - 1) / e->scatter_stripe;
        const bool striped = n_stripes > 1 && n_stripes <= 16 && (uint64_t)n >= e->scatter_stripe_min_reads && (uint64_t)n >= n_stripes * std::max<uint64_t>(1, e->scatter_stripe_min_reads / 64);
        int rc = launch(e, K_SCATTER, [&] {
            if (!striped) { pdk::k_delta_scatter<<<(n + 255) / 256, 256, 0, e->stream>>>(s); return; }
            for (uint64_t k = 0; k < n_stripes; k++) {
                s.stripe_lo = (int64_t)(k * e->scatter_stripe); s.stripe_hi = k + 1 == n_stripes ? INT64_MAX : (int64_t)((k + 1) * e->scatter_stripe);
                pdk::k_delta_scatter<<<(n + 255) / 256, 256, 0, e->stream>>>(s);
            }
        }); if (rc) return rc;
        if (striped) e->stats.kernel_launches += n_stripes - 1;
        e->host_delta_used = true;
        e->stats.delta_path_batches++;
    }
    e->stats.reads_pushed += n;
    e->finalized = false;
    e->pushed_since_reset = true;
    return PD_OK;
}

int ensure_slot(pd_engine *e, Slot &s, size_t reads, size_t ncig, bool need_off, bool need_flag, bool need_mapq)
{
    if (!s.copied) { CK(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming)); }
    if (reads > s.cap_reads) {
        if (s.used) CK(cudaEventSynchronize(s.consumed));
        cudaFree(s.pos); cudaFree(s.cigar_off); cudaFree(s.flag); cudaFree(s.mapq);
        s.pos = nullptr; s.cigar_off = nullptr; s.flag = nullptr; s.mapq = nullptr; s.cap_reads = 0;
        size_t cap = reads + reads / 8 + 1024;
        CK(cudaMalloc(&s.pos, cap * 4)); CK(cudaMalloc(&s.cigar_off, (cap + 1) * 4));
        CK(cudaMalloc(&s.flag, cap * 2)); CK(cudaMalloc(&s.mapq, cap));
        s.cap_reads = cap;
        e->stats.hbm_bytes_allocated += cap * 11 + 4;
    }
    if (ncig > s.cap_cigar) {
        if (s.used) CK(cudaEventSynchronize(s.consumed));
        cudaFree(s.cigar); s.cigar = nullptr; s.cap_cigar = 0;
        size_t cap = ncig + ncig / 8 + 1024;
        CK(cudaMalloc(&s.cigar, cap * 4));
        s.cap_cigar = cap;
        e->stats.hbm_bytes_allocated += cap * 4;
    }
    (void)need_off; (void)need_flag; (void)need_mapq;
    return PD_OK;
}

int check_device_flags(pd_engine *e)
{
    int h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, e->d_flags, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (h[1]) e->host_delta_used = true;        // before the error returns: pd_reset must know the tile kernel wrote far-run deltas
    if (h[0] & pdk::ERR_UNSORTED) return fail(e, PD_ERR_UNSORTED, "a batch pushed with PD_SORTED was not sorted by pos");
    if (h[0] & pdk::ERR_RANGE) return fail(e, PD_ERR_UNSORTED, "a PD_SORTED batch had positions outside their segment (unsorted input)");
    return PD_OK;
}

// ---- device-side BGZF decode (pd_push_bgzf) ------------------------------------------------------------------------------
template <typename T>
int bz_grow(pd_engine *e, T *&p, size_t &cap, size_t need, size_t *cap_out = nullptr)
{
    (void)cap_out;
    if (need <= cap) return PD_OK;
    if (p) { CK(cudaDeviceSynchronize()); cudaFree(p); p = nullptr; cap = 0; }
    const size_t c = need + need / 4 + 1024;
    CK(cudaMalloc(&p, c * sizeof(T)));
    e->stats.hbm_bytes_allocated += c * sizeof(T);
    cap = c;
    return PD_OK;
}

int bz_ensure(pd_engine *e, BzSlot &s, size_t in_bytes, uint64_t out_total, uint32_t n_blocks, bool device_input)
{
    if (!s.copied) {
        CK(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s.in_free, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.decoded, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CK(cudaMalloc(&s.d_totals, 16)); CK(cudaMalloc(&s.d_ticket, 4));
        CK(cudaHostAlloc(&s.h_totals, 16, cudaHostAllocDefault));
    }
    int rc;
    if (!device_input) { rc = bz_grow(e, s.d_in, s.in_cap, in_bytes + 64); if (rc) return rc; }
    rc = bz_grow(e, s.d_out, s.out_cap, (size_t)out_total + 64); if (rc) return rc;
    if (n_blocks > s.blk_cap) {
        if (s.d_blk) CK(cudaDeviceSynchronize());
        cudaFree(s.d_blk); cudaFree(s.d_sum); cudaFree(s.d_r0); cudaFree(s.d_c0);
        s.d_blk = nullptr; s.d_sum = nullptr; s.d_r0 = s.d_c0 = nullptr; s.blk_cap = 0;
        const size_t c = (size_t)n_blocks + n_blocks / 4 + 256;
        CK(cudaMalloc(&s.d_blk, c * sizeof(pdk::BgzfBlockDesc))); CK(cudaMalloc(&s.d_sum, c * sizeof(pdk::BgzfSummary)));
        CK(cudaMalloc(&s.d_r0, c * 4)); CK(cudaMalloc(&s.d_c0, c * 4));
        s.blk_cap = c;
    }
    if (n_blocks > s.h_blk_cap) {
        if (s.h_blk) { cudaFreeHost(s.h_blk); cudaFreeHost(s.h_sum); s.h_blk = nullptr; s.h_sum = nullptr; s.h_blk_cap = 0; }
        const size_t c = (size_t)n_blocks + n_blocks / 4 + 256;
        CK(cudaHostAlloc(&s.h_blk, c * sizeof(pdk::BgzfBlockDesc), cudaHostAllocDefault));
        CK(cudaHostAlloc(&s.h_sum, c * sizeof(pdk::BgzfSummary), cudaHostAllocDefault));
        s.h_blk_cap = c;
    }
    // staging (block-local) and dense outputs: sized by what the inflated bytes can hold at most
    const size_t st_rec = (size_t)(out_total / 32) + 2ull * n_blocks + 64, st_cig = (size_t)(out_total / 4) + n_blocks + 64;
    const size_t n_rec = (size_t)(out_total / 36) + 64, n_cig = (size_t)(out_total / 4) + 64;
    if (st_rec > s.st_rec_cap) {
        size_t c0 = s.st_rec_cap, c1 = s.st_rec_cap, c2 = s.st_rec_cap; s.st_rec_cap = 0;      // (restored below only if all three grew)
        rc = bz_grow(e, s.st_pos, c0, st_rec); if (rc) return rc;
        rc = bz_grow(e, s.st_tid, c1, st_rec); if (rc) return rc;
        rc = bz_grow(e, s.st_coff, c2, st_rec); if (rc) return rc;
        s.st_rec_cap = c0;
    }
    rc = bz_grow(e, s.st_cig, s.st_cig_cap, st_cig); if (rc) return rc;
    if (n_rec + 1 > s.rec_cap) {
        size_t c0 = s.rec_cap, c1 = s.rec_cap, c2 = s.rec_cap; s.rec_cap = 0;
        rc = bz_grow(e, s.pos, c0, n_rec + 1); if (rc) return rc;
        rc = bz_grow(e, s.tid, c1, n_rec + 1); if (rc) return rc;
        rc = bz_grow(e, s.coff, c2, n_rec + 1); if (rc) return rc;
        s.rec_cap = c0;
    }
    rc = bz_grow(e, s.cig, s.cig_cap, n_cig); if (rc) return rc;
    return PD_OK;
}

// Drop whatever the decoder still has in flight (error paths, pd_reset).
int bz_cancel(pd_engine *e)
{
    bool any = false;
    for (auto &s : e->bz) { any = any || s.pending; s.pending = false; }
    if (any) { CK(cudaStreamSynchronize(e->copy_stream)); for (auto &s : e->bz) if (s.stream) CK(cudaStreamSynchronize(s.stream)); CK(cudaStreamSynchronize(e->stream)); }
    return PD_OK;
}

// Second half of a chunk: its per-block summaries are back on the host -> check them, cut the records into contig runs and
// queue the accumulate kernels.  Sorted single-contig runs take the tile path; blocks holding a contig change or records out
// of order go through the per-record-tid delta scatter.
int bz_finish(pd_engine *e, BzSlot &s)
{
    if (!s.pending) return PD_OK;
    s.pending = false;
    CK(cudaEventSynchronize(s.decoded));
    const uint32_t n_reads = s.h_totals[0], n_cigar = s.h_totals[1], fl = s.h_totals[2];
    if (fl & pdk::BZ_INFLATE) { bz_cancel(e); return fail(e, PD_ERR_ARG, "pd_push_bgzf: a BGZF block does not inflate to its ISIZE (corrupt or truncated block)"); }
    if (fl & pdk::BZ_CRC) { bz_cancel(e); return fail(e, PD_ERR_ARG, "pd_push_bgzf: BGZF block CRC mismatch (corrupt block)"); }
    if (fl & (pdk::BZ_MALFORMED | pdk::BZ_STRADDLE | pdk::BZ_CG_TAG)) {
        bz_cancel(e);
        return fail(e, PD_ERR_FALLBACK, (fl & pdk::BZ_CG_TAG) ? "pd_push_bgzf: a CIGAR may live in the CG tag (host path needed)"
                                                               : "pd_push_bgzf: records do not start on BGZF block boundaries (host path needed)");
    }
    if (!n_reads) return PD_OK;
    CK(cudaStreamWaitEvent(e->stream, s.decoded, 0));          // the accumulate kernels below read the slot's dense arrays
    const bool single_op = !(fl & pdk::BZ_MULTI_OP);           // every record has exactly one CIGAR word: cigar[i] belongs to read i
    if (e->delta_stale) { CK(cudaMemsetAsync(e->d_delta, 0, e->total * 4, e->stream)); e->delta_stale = false; }
    const pdk::BgzfSummary *sm = s.h_sum;
    // exclusive prefix of reads per block (the device computed the same; recomputed here to avoid another copy)
    uint32_t r = 0;
    struct Run { int32_t tid; uint32_t r0, r1; int32_t p0, p1; bool sorted; };
    Run cur{0, 0, 0, 0, 0, false}; bool open = false;
    uint32_t irr0 = 0, irr1 = 0; bool irr_open = false;
    int32_t prev_tid = INT32_MIN, prev_pos = INT32_MIN;
    int rc = PD_OK;
    auto flush_run = [&]() -> int {
        if (!open) return PD_OK;
        open = false;
        const int32_t t = cur.tid;
        if (t < 0 || t >= e->n_targets || !e->owned[t] || e->elen[t] == 0) return PD_OK;
        const int64_t max_seg = ((int64_t)e->elen[t]) / pdk::SEG;
        int64_t p0 = std::max<int64_t>((int64_t)cur.p0 - e->org[t], 0), p1 = std::max<int64_t>((int64_t)cur.p1 - e->org[t], 0);
        int64_t seg_a = std::min<int64_t>(p0 / pdk::SEG, max_seg), seg_b = std::min<int64_t>((p1 + pdk::HALO) / pdk::SEG, max_seg);
        if (seg_b < seg_a) seg_b = seg_a;
        return run_batch(e, t, cur.r1 - cur.r0, s.pos + cur.r0, nullptr, nullptr, single_op ? nullptr : s.coff + cur.r0,
                         single_op ? s.cig + cur.r0 : s.cig, true, seg_a, seg_b, s.decoded);
    };
    auto flush_irr = [&]() -> int {
        if (!irr_open) return PD_OK;
        irr_open = false;
        const uint32_t n = irr1 - irr0;
        pdk::ScatterMultiArgs m{s.tid + irr0, s.pos + irr0, s.cig, s.coff + irr0, n, e->d_off, e->d_tlen, e->d_org, e->d_owned, e->n_targets, e->d_delta};
        int rc2 = launch(e, K_SCATTER, [&] { pdk::k_delta_scatter_multi<<<(n + 255) / 256, 256, 0, e->stream>>>(m); });
        if (rc2) return rc2;
        e->host_delta_used = true; e->stats.delta_path_batches++; e->stats.reads_pushed += n;
        e->finalized = false; e->pushed_since_reset = true;
        return PD_OK;
    };
    // Many-contig inputs (pan-genome assemblies): a chunk may hold thousands of contigs of a few blocks each, and a tile-path
    // run costs three launches.  When a chunk has more than MANY_RUNS single-contig runs, the runs shorter than SHORT_RUN blocks
    // join the per-record-tid scatter (one launch for any number of contigs); long runs keep the tile path.
    constexpr uint32_t MANY_RUNS = 32, SHORT_RUN = 8;
    std::vector<uint8_t> to_scatter;
    {
        std::vector<std::pair<uint32_t, uint32_t>> runs;       // [first block, last block] of every would-be tile run
        bool o2 = false; int32_t ct = 0, pt = INT32_MIN, pp = INT32_MIN; uint32_t first = 0, lastb = 0;
        for (uint32_t b = 0; b < s.n_blocks; b++) {
            if (!sm[b].n_reads) continue;
            const bool regular = !(sm[b].flags & pdk::BZ_UNSORTED) && sm[b].first_tid == sm[b].last_tid;
            if (regular && o2 && ct == sm[b].first_tid && sm[b].first_pos >= pp && pt == ct) lastb = b;
            else {
                if (o2) runs.emplace_back(first, lastb);
                o2 = regular;
                if (regular) { ct = sm[b].first_tid; first = lastb = b; }
            }
            pt = sm[b].last_tid; pp = sm[b].last_pos;
        }
        if (o2) runs.emplace_back(first, lastb);
        if (runs.size() > MANY_RUNS) {
            to_scatter.assign(s.n_blocks, 0);
            for (auto &q : runs) {
                uint32_t nb = 0;
                for (uint32_t b = q.first; b <= q.second; b++) nb += sm[b].n_reads ? 1u : 0u;
                if (nb < SHORT_RUN) for (uint32_t b = q.first; b <= q.second; b++) to_scatter[b] = 1;
            }
        }
    }
    for (uint32_t b = 0; b < s.n_blocks; b++) {
        const uint32_t n = sm[b].n_reads;
        if (!n) continue;
        const bool regular = !(sm[b].flags & pdk::BZ_UNSORTED) && sm[b].first_tid == sm[b].last_tid && !(to_scatter.size() && to_scatter[b]);
        if (regular) {
            if ((rc = flush_irr())) return rc;
            const bool cont = open && cur.tid == sm[b].first_tid && sm[b].first_pos >= prev_pos && prev_tid == cur.tid;
            if (cont) { cur.r1 = r + n; cur.p1 = sm[b].last_pos; }
            else { if ((rc = flush_run())) return rc; cur = Run{sm[b].first_tid, r, r + n, sm[b].first_pos, sm[b].last_pos, true}; open = true; }
        } else {
            if ((rc = flush_run())) return rc;
            if (irr_open && irr1 == r) irr1 = r + n; else { if ((rc = flush_irr())) return rc; irr0 = r; irr1 = r + n; irr_open = true; }
        }
        prev_tid = sm[b].last_tid; prev_pos = sm[b].last_pos;
        r += n;
    }
    if ((rc = flush_run())) return rc;
    if ((rc = flush_irr())) return rc;
    for (uint32_t b = s.n_blocks; b-- > 0;) if (sm[b].n_reads) { e->bz_prog_tid = sm[b].first_tid; e->bz_prog_pos = sm[b].first_pos; break; }
    CK(cudaEventRecord(s.consumed, e->stream)); s.out_used = true;
    (void)n_cigar;
    return PD_OK;
}

}  // namespace

extern "C" {

const char *pd_last_error(pd_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

void *pd_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void pd_host_free(void *p) { if (p) cudaFreeHost(p); }

void pd_destroy(pd_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (auto &p : e->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    cudaFree(e->d_depth); cudaFree(e->d_delta); cudaFree(e->d_seg_state); cudaFree(e->d_off); cudaFree(e->d_tlen);
    for (auto &c : e->cut) { cudaFree(c.lo); cudaFree(c.hi); cudaFree(c.coarse); if (c.ready) cudaEventDestroy(c.ready); if (c.released) cudaEventDestroy(c.released); }
    if (e->cut_stream) cudaStreamDestroy(e->cut_stream);
    if (e->cut_input) cudaEventDestroy(e->cut_input);
    cudaFree(e->d_flags); cudaFree(e->d_desc); cudaFree(e->d_counter);
    cudaFree(e->d_red); cudaFree(e->d_todo); cudaFree(e->d_part_cov); cudaFree(e->d_part_sum);
    for (auto &s : e->slot) {
        cudaFree(s.pos); cudaFree(s.cigar); cudaFree(s.cigar_off); cudaFree(s.flag); cudaFree(s.mapq);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.consumed) cudaEventDestroy(s.consumed);
    }
    for (auto &s : e->bz) {
        cudaFree(s.d_in); cudaFree(s.d_out); cudaFree(s.d_blk); cudaFree(s.d_sum); cudaFree(s.d_r0); cudaFree(s.d_c0);
        cudaFree(s.st_pos); cudaFree(s.st_tid); cudaFree(s.st_coff); cudaFree(s.st_cig);
        cudaFree(s.pos); cudaFree(s.tid); cudaFree(s.coff); cudaFree(s.cig); cudaFree(s.d_totals); cudaFree(s.d_ticket);
        if (s.h_blk) cudaFreeHost(s.h_blk);
        if (s.h_sum) cudaFreeHost(s.h_sum);
        if (s.h_totals) cudaFreeHost(s.h_totals);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.in_free) cudaEventDestroy(s.in_free);
        if (s.decoded) cudaEventDestroy(s.decoded);
        if (s.consumed) cudaEventDestroy(s.consumed);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    cudaFree(e->d_owned); cudaFree(e->d_org); cudaFree(e->d_pass_bits);
    cudaFree(e->bk_pos[0]); cudaFree(e->bk_pos[1]); cudaFree(e->bk_cig[0]); cudaFree(e->bk_cig[1]); cudaFree(e->bk_tab); cudaFree(e->bk_fine);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    delete e;
}

pd_engine *pd_create(int device, int n_targets, const uint32_t *target_len, const uint8_t *owned, uint32_t xflags, int min_mapq)
{
    if (n_targets <= 0 || !target_len) { g_create_error = "pd_create: bad arguments"; return nullptr; }
    std::vector<uint32_t> beg((size_t)n_targets, 0), end(target_len, target_len + n_targets);
    if (owned) for (int t = 0; t < n_targets; ++t) if (!owned[t]) end[(size_t)t] = 0;
    return pd_create_ranges(device, n_targets, target_len, beg.data(), end.data(), xflags, min_mapq);
}

pd_engine *pd_create_ranges(int device, int n_targets, const uint32_t *target_len, const uint32_t *own_beg, const uint32_t *own_end,
                            uint32_t xflags, int min_mapq)
{
    if (n_targets <= 0 || !target_len || !own_beg || !own_end) { g_create_error = "pd_create: bad arguments"; return nullptr; }
    pd_engine *e = new pd_engine;
    auto bail = [&](const char *what, cudaError_t ce) -> pd_engine * {
        g_create_error = std::string("pd_create: ") + what + ": " + cudaGetErrorString(ce);
        pd_destroy(e);
        return nullptr;
    };
    e->device = device; e->n_targets = n_targets; e->xflags = xflags; e->min_mapq = min_mapq;
    e->tlen.assign(target_len, target_len + n_targets);
    e->owned.assign(n_targets, 0); e->org.assign(n_targets, 0); e->elen.assign(n_targets, 0);
    e->off.assign(n_targets, 0);
    uint64_t total = 0;
    for (int t = 0; t < n_targets; ++t) {
        // (segment arithmetic of the tile kernel is 32-bit up to one segment + halo beyond the contig end)
        if (e->tlen[t] > 0x7fffe000u) { g_create_error = "pd_create: contig longer than 2^31 - 8192 bases"; delete e; return nullptr; }
        const uint32_t b = std::min(own_beg[t], e->tlen[t]), z = std::min(own_end[t], e->tlen[t]);
        if (z <= b) continue;                                  // nothing of this contig is owned: pushes for it are ignored
        e->owned[t] = 1; e->org[t] = b; e->elen[t] = z - b;
        if (b != 0 || z != e->tlen[t]) e->has_ranges = true;
        e->off[t] = total;
        total += ((uint64_t)e->elen[t] + 1 + PAD_ELEMS - 1) / PAD_ELEMS * PAD_ELEMS;
    }
    if (total == 0) total = PAD_ELEMS;
    e->total = total; e->n_seg = total / pdk::SEG;
    cudaError_t ce;
    const bool timing = getenv("PANDEPTH_TIMING") != nullptr;           // phase times of the engine start-up on stderr
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (timing) fprintf(stderr, "[timing] pd_create: %-28s +%.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
    };
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", ce);
    if ((ce = cudaFree(0)) != cudaSuccess) return bail("context", ce);
    lap("CUDA context");
    if ((ce = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", ce);
    if ((ce = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", ce);
    if ((ce = cudaStreamCreateWithFlags(&e->cut_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", ce);
    if ((ce = cudaEventCreateWithFlags(&e->cut_input, cudaEventDisableTiming)) != cudaSuccess) return bail("event", ce);
    e->stream = e->own_stream;
    if ((ce = cudaMalloc(&e->d_depth, total * 4)) != cudaSuccess) return bail("cudaMalloc(depth)", ce);
    if ((ce = cudaMalloc(&e->d_delta, total * 4)) != cudaSuccess) return bail("cudaMalloc(delta)", ce);
    lap("cudaMalloc depth + delta");
    if ((ce = cudaMalloc(&e->d_seg_state, e->n_seg)) != cudaSuccess) return bail("cudaMalloc(seg_state)", ce);
    if ((ce = cudaMalloc(&e->d_off, (size_t)n_targets * 8)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMalloc(&e->d_tlen, (size_t)n_targets * 4)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMalloc(&e->d_flags, 8)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMalloc(&e->d_desc, (total / pdk::SCAN_TILE) * 8)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMalloc(&e->d_counter, 4)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMalloc(&e->d_part_cov, e->n_seg * 4)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMalloc(&e->d_part_sum, e->n_seg * 8)) != cudaSuccess) return bail("cudaMalloc", ce);
    e->stats.hbm_bytes_allocated = total * 8 + e->n_seg * 13 + (total / pdk::SCAN_TILE) * 8;
    if ((ce = cudaMemcpy(e->d_off, e->off.data(), (size_t)n_targets * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bail("memcpy", ce);
    if ((ce = cudaMemcpy(e->d_tlen, e->elen.data(), (size_t)n_targets * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bail("memcpy", ce);
    if ((ce = cudaMalloc(&e->d_org, (size_t)n_targets * 4)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMemcpy(e->d_org, e->org.data(), (size_t)n_targets * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bail("memcpy", ce);
    if ((ce = cudaMalloc(&e->d_owned, (size_t)n_targets)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMemcpy(e->d_owned, e->owned.data(), (size_t)n_targets, cudaMemcpyHostToDevice)) != cudaSuccess) return bail("memcpy", ce);
    if ((ce = cudaMemsetAsync(e->d_delta, 0, total * 4, e->stream)) != cudaSuccess) return bail("memset", ce);
    if ((ce = cudaMemsetAsync(e->d_seg_state, 0, e->n_seg, e->stream)) != cudaSuccess) return bail("memset", ce);
    if ((ce = cudaMemsetAsync(e->d_flags, 0, 8, e->stream)) != cudaSuccess) return bail("memset", ce);
    {
        const int smem = pdk::WPC * pdk::TILE_SMEM_WORDS * 4, smem_off = pdk::WPC * (pdk::TILE_SMEM_WORDS + pdk::TILE_QUEUE_WORDS) * 4;
        if ((ce = cudaFuncSetAttribute(pdk::k_tile_accumulate<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return bail("cudaFuncSetAttribute", ce);
        if ((ce = cudaFuncSetAttribute(pdk::k_tile_accumulate<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return bail("cudaFuncSetAttribute", ce);
        if ((ce = cudaFuncSetAttribute(pdk::k_tile_accumulate<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_off)) != cudaSuccess) return bail("cudaFuncSetAttribute", ce);
        if ((ce = cudaFuncSetAttribute(pdk::k_tile_accumulate<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_off)) != cudaSuccess) return bail("cudaFuncSetAttribute", ce);
        if ((ce = cudaFuncSetAttribute(pdk::k_bucket_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pdk::BK_SCATTER_SMEM)) != cudaSuccess) return bail("cudaFuncSetAttribute", ce);
        if ((ce = cudaFuncSetAttribute(pdk::k_bucket_fine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pdk::BK_FINE_SMEM)) != cudaSuccess) return bail("cudaFuncSetAttribute", ce);
    }
    {
        int n_sm = 0, per_sm = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pdk::k_tile_accumulate<false, false>, pdk::WPC * 32, pdk::WPC * pdk::TILE_SMEM_WORDS * 4) == cudaSuccess && n_sm > 0 && per_sm > 0)
            e->tile_resident_warps = n_sm * per_sm * pdk::WPC;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pdk::k_bgzf_decode, pdk::BZ_WPC * 32, 0) == cudaSuccess && n_sm > 0 && per_sm > 0)
            e->bz_grid = n_sm * per_sm;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pdk::k_bgzf_decode_multi<2, 4>, 4 * 32, 0) == cudaSuccess && n_sm > 0 && per_sm > 0)
            e->bz_grid_k2 = n_sm * per_sm;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pdk::k_bgzf_decode_multi<4, 2>, 2 * 32, 0) == cudaSuccess && n_sm > 0 && per_sm > 0)
            e->bz_grid_k4 = n_sm * per_sm;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pdk::k_bgzf_decode_multi<8, 1>, 1 * 32, 0) == cudaSuccess && n_sm > 0 && per_sm > 0)
            e->bz_grid_k8 = n_sm * per_sm;
        if (const char *fk = getenv("PANDEPTH_BGZF_K")) { const int v = atoi(fk); if (v == 1 || v == 2 || v == 4 || v == 8) e->bz_force_k = v; }
        if (const char *v = getenv("PANDEPTH_BUCKET_MIN_READS")) e->bucket_min_reads = (uint64_t)std::max<long long>(0, atoll(v));
        if (const char *v = getenv("PANDEPTH_SCATTER_STRIPE_MB")) { const double mb = atof(v); if (mb >= 0.001) e->scatter_stripe = std::max<uint64_t>(256, (uint64_t)(mb * 262144.0)); }
        if (const char *v = getenv("PANDEPTH_SCATTER_STRIPE_MIN_READS")) e->scatter_stripe_min_reads = (uint64_t)std::max<long long>(1, atoll(v));
        if (const char *fk = getenv("PANDEPTH_BGZF_LIT_K")) { const int v = atoi(fk); if (v == 1 || v == 2 || v == 4 || v == 8) e->bz_lit_k = v; }
    }
    if ((ce = cudaStreamSynchronize(e->stream)) != cudaSuccess) return bail("sync", ce);
    lap("tables, memsets, attributes");
    return e;
}

int pd_set_stream(pd_engine *e, void *cuda_stream)
{
    if (!e) return PD_ERR_ARG;
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return PD_OK;
}

int pd_set_reduce_hint(pd_engine *e, uint32_t min_dep, uint32_t depth_mask)
{
    if (!e) return PD_ERR_ARG;
    if (e->pushed_since_reset) return fail(e, PD_ERR_STATE, "pd_set_reduce_hint must be called before the first push (after pd_create or pd_reset)");
    e->hint_on = min_dep >= 1;          // with min_dep 0 every base counts, including never-written ones: no partials
    e->hint_min_dep = min_dep; e->hint_mask = depth_mask;
    return PD_OK;
}

int pd_reset(pd_engine *e)
{
    NvtxRange nvtx_("pd_reset");
    if (!e) return PD_ERR_ARG;
    CK(cudaSetDevice(e->device));
    { int rc = bz_cancel(e); if (rc) return rc; }
    if (!e->finalized) {                  // un-finalized pushes may have queued far-run deltas on the device
        CK(cudaStreamSynchronize(e->copy_stream));
        (void)check_device_flags(e);      // stale errors are dropped
    }
    e->err.clear();
    CK(cudaMemsetAsync(e->d_seg_state, 0, e->n_seg, e->stream));
    CK(cudaMemsetAsync(e->d_flags, 0, 8, e->stream));
    if (e->host_delta_used || e->delta_stale) CK(cudaMemsetAsync(e->d_delta, 0, e->total * 4, e->stream));
    e->host_delta_used = false;
    e->delta_stale = false;
    e->finalized = false;
    e->pushed_since_reset = false;
    e->bz_prog_tid = -1; e->bz_prog_pos = -1;
    uint64_t keep = e->stats.hbm_bytes_allocated;
    e->stats = pd_stats{};
    e->stats.hbm_bytes_allocated = keep;
    return PD_OK;
}

int pd_push_contig(pd_engine *e, int32_t tid, int64_t n_reads, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
                   const uint32_t *cigar_off, const uint32_t *cigar, uint32_t flags)
{
    NvtxRange nvtx_("pd_push_contig");
    if (!e) return PD_ERR_ARG;
    if (tid < 0 || tid >= e->n_targets || n_reads < 0 || (n_reads > 0 && (!pos || !cigar))) return fail(e, PD_ERR_ARG, "pd_push_contig: bad arguments");
    if (!e->owned[tid] || n_reads == 0 || e->elen[tid] == 0) return PD_OK;
    CK(cudaSetDevice(e->device));
    const bool sorted = (flags & PD_SORTED) != 0;
    const int64_t max_seg = ((int64_t)e->elen[tid]) / pdk::SEG;      // segment of the padding element
    const int64_t org = e->org[tid];

    if (flags & PD_DEVICE) {
        for (int64_t b = 0; b < n_reads; b += 0x7fff0000ll) {
            const uint32_t n = (uint32_t)std::min<int64_t>(0x7fff0000ll, n_reads - b);
            if (cigar_off && b) return fail(e, PD_ERR_ARG, "pd_push_contig: device batches with cigar_off must have < 2^31 reads");
            int rc = run_batch(e, tid, n, pos + b, flag ? flag + b : nullptr, mapq ? mapq + b : nullptr, cigar_off,
                               cigar_off ? cigar : cigar + b, sorted, 0, max_seg, nullptr, (flags & PD_READY) != 0);
            if (rc) return rc;
        }
        return PD_OK;
    }

    // host-resident batch: stage through device slots in chunks, copy stream overlapping the compute stream
    for (int64_t b = 0; b < n_reads; b += HOST_CHUNK_READS) {
        const int64_t n = std::min<int64_t>(HOST_CHUNK_READS, n_reads - b);
        const uint32_t cg0 = cigar_off ? cigar_off[b] : 0;
        const size_t ncig = cigar_off ? (size_t)(cigar_off[b + n] - cg0) : (size_t)n;
        Slot &s = e->slot[e->next_slot]; e->next_slot = (e->next_slot + 1) % N_SLOTS;
        int rc = ensure_slot(e, s, (size_t)n, ncig, cigar_off != nullptr, flag != nullptr, mapq != nullptr); if (rc) return rc;
        if (s.used) CK(cudaStreamWaitEvent(e->copy_stream, s.consumed, 0));
        CK(cudaMemcpyAsync(s.pos, pos + b, (size_t)n * 4, cudaMemcpyHostToDevice, e->copy_stream));
        CK(cudaMemcpyAsync(s.cigar, cigar_off ? cigar + cg0 : cigar + b, ncig * 4, cudaMemcpyHostToDevice, e->copy_stream));
        e->stats.h2d_bytes += (uint64_t)n * 4 + ncig * 4;
        if (cigar_off) { CK(cudaMemcpyAsync(s.cigar_off, cigar_off + b, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, e->copy_stream)); e->stats.h2d_bytes += (uint64_t)(n + 1) * 4; }
        if (flag) { CK(cudaMemcpyAsync(s.flag, flag + b, (size_t)n * 2, cudaMemcpyHostToDevice, e->copy_stream)); e->stats.h2d_bytes += (uint64_t)n * 2; }
        if (mapq) { CK(cudaMemcpyAsync(s.mapq, mapq + b, (size_t)n, cudaMemcpyHostToDevice, e->copy_stream)); e->stats.h2d_bytes += (uint64_t)n; }
        CK(cudaEventRecord(s.copied, e->copy_stream));
        CK(cudaStreamWaitEvent(e->stream, s.copied, 0));
        int64_t seg_a = 0, seg_b = max_seg;
        if (sorted) {
            // sorted => first/last positions bound the segments this chunk can touch (+HALO of look-ahead for the halo readers)
            int64_t p0 = pos[b], p1 = pos[b + n - 1];
            if (p1 < p0) return fail(e, PD_ERR_UNSORTED, "a batch pushed with PD_SORTED was not sorted by pos (last < first)");
            p0 -= org; p1 -= org;                               // positions relative to the owned range
            if (p0 < 0) p0 = 0;
            if (p1 < 0) p1 = 0;
            seg_a = std::min<int64_t>(p0 / pdk::SEG, max_seg);
            seg_b = std::min<int64_t>((p1 + pdk::HALO) / pdk::SEG, max_seg);
            if (seg_b < seg_a) seg_b = seg_a;
        }
        // cigar offsets on the device are absolute: rebase the cigar pointer instead of rewriting them
        const uint32_t *d_cig = cigar_off ? s.cigar - cg0 : s.cigar;
        rc = run_batch(e, tid, (uint32_t)n, s.pos, flag ? s.flag : nullptr, mapq ? s.mapq : nullptr,
                       cigar_off ? s.cigar_off : nullptr, d_cig, sorted, seg_a, seg_b, s.copied, false, cigar_off ? (uint64_t)ncig : 0);
        if (rc) return rc;
        CK(cudaEventRecord(s.consumed, e->stream));
        s.used = true;
    }
    if (!(flags & PD_ASYNC)) CK(cudaStreamSynchronize(e->copy_stream));
    return PD_OK;
}

int pd_push(pd_engine *e, int64_t n_reads, const int32_t *tid, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
            const uint32_t *cigar_off, const uint32_t *cigar, uint32_t flags)
{
    if (!e) return PD_ERR_ARG;
    if (flags & PD_DEVICE) return fail(e, PD_ERR_ARG, "pd_push: per-record tid arrays must be host pointers (use pd_push_contig for device batches)");
    if (n_reads < 0 || (n_reads > 0 && (!tid || !pos || !cigar))) return fail(e, PD_ERR_ARG, "pd_push: bad arguments");
    int64_t b = 0;
    while (b < n_reads) {
        const int32_t t = tid[b];
        int64_t end;
        if (flags & PD_SORTED) end = std::upper_bound(tid + b, tid + n_reads, t) - tid;     // tids are non-decreasing
        else { end = b + 1; while (end < n_reads && tid[end] == t) ++end; }
        if (t >= 0 && t < e->n_targets) {
            // cigar_off is absolute into `cigar`: pass the sub-array of offsets, the full cigar base
            int rc = pd_push_contig(e, t, end - b, pos + b, flag ? flag + b : nullptr, mapq ? mapq + b : nullptr,
                                    cigar_off ? cigar_off + b : nullptr, cigar_off ? cigar : cigar + b, flags);
            if (rc) return rc;
        }
        b = end;
    }
    return PD_OK;
}

int pd_push_bgzf(pd_engine *e, const uint8_t *bytes, size_t n_bytes, const pd_bgzf_block *blocks, int64_t n_blocks,
                 uint32_t first_skip, uint32_t flags)
{
    NvtxRange nvtx_("pd_push_bgzf");
    if (!e) return PD_ERR_ARG;
    if (n_blocks < 0 || (n_blocks > 0 && (!bytes || !blocks)) || n_blocks > 0x7fffffffll) return fail(e, PD_ERR_ARG, "pd_push_bgzf: bad arguments");
    CK(cudaSetDevice(e->device));
    const bool dev = (flags & PD_DEVICE) != 0, crc = (flags & PD_CRC) != 0;
    int rc = PD_OK;
    if (n_blocks > 0) {
        BzSlot &s = e->bz[e->bz_next];
        e->bz_next = (e->bz_next + 1) % N_BZ;
        if (s.pending) { rc = bz_finish(e, s); if (rc) return rc; }
        // block table with inflated offsets; the bytes that travel are [first payload, end of last payload)
        uint64_t lo = UINT64_MAX, hi = 0, out_total = 0;
        for (int64_t b = 0; b < n_blocks; b++) {
            const pd_bgzf_block &k = blocks[b];
            if (k.in_off > n_bytes || k.in_len > n_bytes - k.in_off || k.out_len > 65536u) return fail(e, PD_ERR_ARG, "pd_push_bgzf: block outside the buffer or ISIZE > 64 KB");
            lo = std::min<uint64_t>(lo, k.in_off); hi = std::max<uint64_t>(hi, k.in_off + k.in_len + (crc ? 8u : 0u));
            if (crc && (uint64_t)k.in_off + k.in_len + 8u > n_bytes) return fail(e, PD_ERR_ARG, "pd_push_bgzf: PD_CRC needs the 8-byte trailer behind every payload inside the buffer");
            out_total += k.out_len;
        }
        if (out_total > 0xf0000000ull) return fail(e, PD_ERR_ARG, "pd_push_bgzf: chunk inflates to more than 3.75 GB; pass smaller chunks");
        rc = bz_ensure(e, s, (size_t)(hi - lo), out_total + 32ull * (uint64_t)n_blocks, (uint32_t)n_blocks, dev); if (rc) return rc;
        uint64_t oo = 0;
        for (int64_t b = 0; b < n_blocks; b++) {
            s.h_blk[b] = pdk::BgzfBlockDesc{blocks[b].in_off - lo, blocks[b].in_len, blocks[b].out_len, oo};
            oo += (blocks[b].out_len + 31u) & ~31u;               // every block's output starts on a 32-byte boundary (whole sectors for the CRC)
        }
        s.n_blocks = (uint32_t)n_blocks;
        const uint8_t *d_in = dev ? bytes + lo : s.d_in;
        if (!dev) {
            if (s.in_used) CK(cudaStreamWaitEvent(e->copy_stream, s.in_free, 0));
            CK(cudaMemcpyAsync(s.d_in, bytes + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, e->copy_stream));
            e->stats.h2d_bytes += hi - lo;
        }
        CK(cudaMemcpyAsync(s.d_blk, s.h_blk, (size_t)n_blocks * sizeof(pdk::BgzfBlockDesc), cudaMemcpyHostToDevice, e->copy_stream));
        e->stats.h2d_bytes += (uint64_t)n_blocks * sizeof(pdk::BgzfBlockDesc);
        CK(cudaEventRecord(s.copied, e->copy_stream));
        CK(cudaStreamWaitEvent(s.stream, s.copied, 0));
        if (s.out_used) CK(cudaStreamWaitEvent(s.stream, s.consumed, 0));     // the slot's previous arrays have been accumulated
        CK(cudaMemsetAsync(s.d_ticket, 0, 4, s.stream));
        pdk::BzArgs a{d_in, s.d_out, s.d_blk, (uint32_t)n_blocks, first_skip, e->n_targets, e->xflags, e->min_mapq,
                      s.st_pos, s.st_tid, s.st_coff, s.st_cig, s.d_sum, s.d_ticket, crc ? 1u : 0u};
        // literal-dominated streams (BAMs with SEQ + QUAL inflate by less than 4x): K decoders per warp, symbol loop in lockstep
        int K = out_total < 4 * (hi - lo) ? e->bz_lit_k : 1;
        if (e->bz_force_k) K = e->bz_force_k;
        rc = launch_on(e, s.stream, K_BGZF, [&] {
            if (K == 2) { const unsigned g = (unsigned)std::min<int64_t>((n_blocks + 7) / 8, e->bz_grid_k2); pdk::k_bgzf_decode_multi<2, 4><<<g, 4 * 32, 0, s.stream>>>(a); }
            else if (K == 4) { const unsigned g = (unsigned)std::min<int64_t>((n_blocks + 7) / 8, e->bz_grid_k4); pdk::k_bgzf_decode_multi<4, 2><<<g, 2 * 32, 0, s.stream>>>(a); }
            else if (K == 8) { const unsigned g = (unsigned)std::min<int64_t>((n_blocks + 7) / 8, e->bz_grid_k8); pdk::k_bgzf_decode_multi<8, 1><<<g, 1 * 32, 0, s.stream>>>(a); }
            else { const unsigned g = (unsigned)std::min<int64_t>((n_blocks + pdk::BZ_WPC - 1) / pdk::BZ_WPC, e->bz_grid); pdk::k_bgzf_decode<<<g, pdk::BZ_WPC * 32, 0, s.stream>>>(a); }
        });
        if (rc) return rc;
        CK(cudaEventRecord(s.in_free, s.stream)); s.in_used = true;
        pdk::BzCompactArgs c{s.d_blk, s.d_sum, s.d_r0, s.d_c0, (uint32_t)n_blocks, s.st_pos, s.st_tid, s.st_coff, s.st_cig, s.pos, s.tid, s.coff, s.cig};
        rc = launch_on(e, s.stream, K_BGZF_PACK, [&] {
            pdk::k_bgzf_offsets<<<1, 1024, 0, s.stream>>>(s.d_sum, (uint32_t)n_blocks, s.d_r0, s.d_c0, s.d_totals);
            pdk::k_bgzf_compact<<<(unsigned)((n_blocks + 7) / 8), 256, 0, s.stream>>>(c);
        });
        if (rc) return rc;
        e->stats.kernel_launches++;
        CK(cudaMemcpyAsync(s.h_sum, s.d_sum, (size_t)n_blocks * sizeof(pdk::BgzfSummary), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(s.h_totals, s.d_totals, 16, cudaMemcpyDeviceToHost, s.stream));
        e->stats.d2h_bytes += (uint64_t)n_blocks * sizeof(pdk::BgzfSummary) + 16;
        CK(cudaEventRecord(s.decoded, s.stream));
        s.pending = true;
        e->finalized = false; e->pushed_since_reset = true;
        // the oldest chunk in flight has had whole decodes to finish: plan and queue its accumulate kernels now
        // (PD_BGZF_DEPTH chunks stay in flight; without PD_ASYNC everything is drained)
        if (!(flags & PD_ASYNC)) { for (int k = 0; k < N_BZ; k++) { rc = bz_finish(e, e->bz[(e->bz_next + k) % N_BZ]); if (rc) return rc; } }
        else { BzSlot &oldest = e->bz[e->bz_next]; if (oldest.pending) { rc = bz_finish(e, oldest); if (rc) return rc; } }
    }
    if (!(flags & PD_ASYNC)) CK(cudaStreamSynchronize(e->copy_stream));
    return PD_OK;
}

int pd_bgzf_progress(pd_engine *e, int32_t *tid, int32_t *pos)
{
    if (!e || !tid || !pos) return PD_ERR_ARG;
    *tid = e->bz_prog_tid; *pos = e->bz_prog_pos;
    return PD_OK;
}

int pd_sync(pd_engine *e)
{
    NvtxRange nvtx_("pd_sync");
    if (!e) return PD_ERR_ARG;
    CK(cudaSetDevice(e->device));
    for (int k = 0; k < N_BZ; k++) { int rc = bz_finish(e, e->bz[(e->bz_next + k) % N_BZ]); if (rc) return rc; }
    CK(cudaStreamSynchronize(e->copy_stream));
    CK(cudaStreamSynchronize(e->stream));
    return check_device_flags(e);
}

int pd_finalize(pd_engine *e)
{
    NvtxRange nvtx_("pd_finalize");
    if (!e) return PD_ERR_ARG;
    CK(cudaSetDevice(e->device));
    for (int k = 0; k < N_BZ; k++) { int rc = bz_finish(e, e->bz[(e->bz_next + k) % N_BZ]); if (rc) return rc; }
    if (e->finalized) return PD_OK;
    int rc = check_device_flags(e); if (rc) return rc;
    if (e->host_delta_used) {
        // the scan visits every element: it also zero-fills (plain store) the segments no tile batch wrote
        const uint64_t n_tiles = e->total / pdk::SCAN_TILE;
        CK(cudaMemsetAsync(e->d_desc, 0, n_tiles * 8, e->stream));
        CK(cudaMemsetAsync(e->d_counter, 0, 4, e->stream));
        rc = launch(e, K_SCAN, [&] { pdk::k_scan_add<<<(unsigned)n_tiles, pdk::SCAN_THREADS, 0, e->stream>>>(e->d_depth, e->d_delta, e->total, e->d_desc, e->d_counter, e->d_seg_state); });
        if (rc) return rc;
        CK(cudaMemsetAsync(e->d_seg_state, 2, e->n_seg, e->stream));     // written, partials (if any) no longer valid
        CK(cudaMemsetAsync(e->d_flags + 1, 0, 4, e->stream));
        e->host_delta_used = false;
        e->delta_stale = true;
        e->stats.scan_passes++;
    } else {
        rc = launch(e, K_FILL, [&] { pdk::k_fill_untouched<<<(unsigned)((e->n_seg + 255) / 256), 256, 0, e->stream>>>(e->d_depth, e->d_seg_state, e->n_seg, e->hint_on ? 1u : 0u, e->d_part_cov, e->d_part_sum); });
        if (rc) return rc;
    }
    e->finalized = true;
    return PD_OK;
}

static int ensure_todo(pd_engine *e, size_t n)
{
    if (n <= e->todo_cap) return PD_OK;
    if (e->d_todo) { CK(cudaStreamSynchronize(e->stream)); cudaFree(e->d_todo); e->d_todo = nullptr; e->todo_cap = 0; }
    const size_t cap = n + n / 4 + 1024;
    CK(cudaMalloc(&e->d_todo, cap * 4));
    e->todo_cap = cap;
    return PD_OK;
}

static int ensure_red(pd_engine *e, size_t bytes)
{
    if (bytes <= e->red_cap) return PD_OK;
    if (e->d_red) { CK(cudaStreamSynchronize(e->stream)); cudaFree(e->d_red); e->d_red = nullptr; e->red_cap = 0; }
    size_t cap = bytes + bytes / 4 + 4096;
    CK(cudaMalloc(&e->d_red, cap));
    e->red_cap = cap;
    return PD_OK;
}

int pd_reduce(pd_engine *e, int64_t n_iv, const int32_t *tid, const int32_t *start0, const int32_t *end0_excl,
              uint32_t min_dep, uint32_t depth_mask, uint64_t *cov, uint64_t *sum, uint32_t flags)
{
    NvtxRange nvtx_("pd_reduce");
    if (!e) return PD_ERR_ARG;
    if (n_iv < 0 || (n_iv > 0 && (!tid || !start0 || !end0_excl || !cov || !sum))) return fail(e, PD_ERR_ARG, "pd_reduce: bad arguments");
    if (flags & PD_DEVICE) return fail(e, PD_ERR_ARG, "pd_reduce: interval tables are host pointers");
    CK(cudaSetDevice(e->device));
    int rc = pd_finalize(e); if (rc) return rc;
    if (n_iv == 0) return PD_OK;
    // one warp per unit = (interval, 8192-base chunk); unit_prefix[i] = units before interval i.
    // intervals of contigs this engine does not own (or empty after clipping) have no units and yield 0
    std::vector<uint64_t> prefix((size_t)n_iv + 1);
    uint64_t n_units = 0;
    for (int64_t i = 0; i < n_iv; ++i) {
        prefix[i] = n_units;
        const int32_t t = tid[i];
        if (t < 0 || t >= e->n_targets) return fail(e, PD_ERR_ARG, "pd_reduce: tid out of range");
        if (!e->owned[t]) continue;
        int64_t s = std::max<int64_t>((int64_t)start0[i] - e->org[t], 0), en = std::min<int64_t>((int64_t)end0_excl[i] - e->org[t], e->elen[t]);
        if (s >= en) continue;
        n_units += (uint64_t)((en - 1) / pdk::REDUCE_CHUNK - s / pdk::REDUCE_CHUNK + 1);      // REDUCE_CHUNK-aligned pieces
    }
    prefix[n_iv] = n_units;
    const size_t iv_bytes = (size_t)n_iv * 4;
    const size_t a_tid = 0, a_s = a_tid + ((iv_bytes + 15) & ~15ull), a_e = a_s + ((iv_bytes + 15) & ~15ull);
    const size_t a_units = a_e + ((iv_bytes + 15) & ~15ull);
    const size_t a_cov = a_units + ((prefix.size() * 8 + 15) & ~15ull), a_sum = a_cov + (size_t)n_iv * 8;
    const size_t need = a_sum + (size_t)n_iv * 8;
    rc = ensure_red(e, need); if (rc) return rc;
    char *base = (char *)e->d_red;
    CK(cudaMemcpyAsync(base + a_tid, tid, iv_bytes, cudaMemcpyHostToDevice, e->stream));
    std::vector<int32_t> sh_s, sh_e;                         // range sharding: intervals relative to the owned ranges (clipped by the kernel)
    if (e->has_ranges) {
        sh_s.resize((size_t)n_iv); sh_e.resize((size_t)n_iv);
        for (int64_t i = 0; i < n_iv; ++i) {
            const int64_t o = e->org[tid[i]];
            sh_s[(size_t)i] = (int32_t)std::max<int64_t>((int64_t)start0[i] - o, INT32_MIN);
            sh_e[(size_t)i] = (int32_t)std::max<int64_t>((int64_t)end0_excl[i] - o, INT32_MIN);
        }
        start0 = sh_s.data(); end0_excl = sh_e.data();
    }
    CK(cudaMemcpyAsync(base + a_s, start0, iv_bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(base + a_e, end0_excl, iv_bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(base + a_units, prefix.data(), prefix.size() * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(base + a_cov, 0, (size_t)n_iv * 16, e->stream));
    e->stats.h2d_bytes += iv_bytes * 3 + prefix.size() * 8;
    if (n_units) {
        const bool use_part = e->hint_on && min_dep == e->hint_min_dep && depth_mask == e->hint_mask;
        uint32_t *d_todo = nullptr;
        if (use_part && pdk::SEG == 2048) {
            rc = ensure_todo(e, (size_t)n_units); if (rc) return rc;
            d_todo = e->d_todo;
            CK(cudaMemsetAsync(d_todo, 0, (size_t)n_units * 4, e->stream));
        }
        pdk::ReduceArgs r{e->d_depth, e->d_off, e->d_tlen, (const int32_t *)(base + a_tid), (const int32_t *)(base + a_s),
                          (const int32_t *)(base + a_e), (const uint64_t *)(base + a_units), (uint64_t)n_iv, n_units, min_dep, depth_mask,
                          (unsigned long long *)(base + a_cov), (unsigned long long *)(base + a_sum),
                          use_part ? 1u : 0u, e->d_seg_state, e->d_part_cov, e->d_part_sum, d_todo};
        rc = launch(e, K_REDUCE, [&] {
            if (d_todo) pdk::k_reduce_units_part<<<(unsigned)((n_units + 255) / 256), 256, 0, e->stream>>>(r, d_todo);
            pdk::k_reduce_units<<<(unsigned)((n_units + 7) / 8), 256, 0, e->stream>>>(r);
        });
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(cov, base + a_cov, (size_t)n_iv * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(sum, base + a_sum, (size_t)n_iv * 8, cudaMemcpyDeviceToHost, e->stream));
    e->stats.d2h_bytes += (uint64_t)n_iv * 16;
    CK(cudaStreamSynchronize(e->stream));
    return PD_OK;
}

int pd_reduce_windows(pd_engine *e, int32_t tid, int32_t start0, int32_t win, int64_t n_win, uint32_t min_dep, uint32_t depth_mask,
                      uint32_t *cov, uint64_t *sum, uint32_t flags)
{
    NvtxRange nvtx_("pd_reduce_windows");
    if (!e) return PD_ERR_ARG;
    if (tid < 0 || tid >= e->n_targets || win < 1 || start0 < 0 || n_win < 0 || (n_win > 0 && (!cov || !sum))) return fail(e, PD_ERR_ARG, "pd_reduce_windows: bad arguments");
    CK(cudaSetDevice(e->device));
    int rc = pd_finalize(e); if (rc) return rc;
    if (n_win == 0) return PD_OK;
    const bool dev = (flags & PD_DEVICE) != 0;
    if (!e->owned[tid]) {
        if (dev) { CK(cudaMemsetAsync(cov, 0, (size_t)n_win * 4, e->stream)); CK(cudaMemsetAsync(sum, 0, (size_t)n_win * 8, e->stream)); }
        else { memset(cov, 0, (size_t)n_win * 4); memset(sum, 0, (size_t)n_win * 8); }
        return PD_OK;
    }
    if (e->org[tid] != 0 || e->elen[tid] != e->tlen[tid])
        return fail(e, PD_ERR_STATE, "pd_reduce_windows: contig owned only in part (range sharding): use pd_reduce, whose results add up across engines");
    uint32_t *d_cov = cov; unsigned long long *d_sum = (unsigned long long *)sum;
    if (!dev) {
        const size_t a_sum = ((size_t)n_win * 4 + 15) & ~15ull;
        rc = ensure_red(e, a_sum + (size_t)n_win * 8); if (rc) return rc;
        d_cov = (uint32_t *)e->d_red; d_sum = (unsigned long long *)((char *)e->d_red + a_sum);
    }
    pdk::WindowArgs w{e->d_depth + e->off[tid], (int64_t)e->tlen[tid], start0, win, n_win, min_dep, depth_mask, d_cov, d_sum};
    const int64_t warps_needed = (win >= 32 && win <= 160 && min_dep >= 1) ? (n_win + 3) / 4 : (win >= 32 ? n_win : (n_win + 31) / 32);
    const unsigned grid = (unsigned)std::min<int64_t>((warps_needed + 7) / 8, 148 * 64);
    rc = launch(e, K_WINDOWS, [&] { pdk::k_reduce_windows<<<grid, 256, 0, e->stream>>>(w); }); if (rc) return rc;
    if (!dev) {
        CK(cudaMemcpyAsync(cov, d_cov, (size_t)n_win * 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(sum, d_sum, (size_t)n_win * 8, cudaMemcpyDeviceToHost, e->stream));
        e->stats.d2h_bytes += (uint64_t)n_win * 12;
        CK(cudaStreamSynchronize(e->stream));
    }
    return PD_OK;
}

int pd_depth_read(pd_engine *e, int32_t tid, int32_t start0, int32_t n, uint32_t *out, uint32_t flags)
{
    NvtxRange nvtx_("pd_depth_read");
    if (!e) return PD_ERR_ARG;
    if (tid < 0 || tid >= e->n_targets || start0 < 0 || n < 0 || (n > 0 && !out)) return fail(e, PD_ERR_ARG, "pd_depth_read: bad arguments");
    if ((int64_t)start0 + n > (int64_t)e->tlen[tid]) return fail(e, PD_ERR_ARG, "pd_depth_read: range beyond the contig");
    CK(cudaSetDevice(e->device));
    int rc = pd_finalize(e); if (rc) return rc;
    if (n == 0) return PD_OK;
    const bool dev = (flags & PD_DEVICE) != 0;
    if (!e->owned[tid]) {
        if (dev) CK(cudaMemsetAsync(out, 0, (size_t)n * 4, e->stream)); else memset(out, 0, (size_t)n * 4);
        return PD_OK;
    }
    // the part of [start0, start0 + n) inside the owned range; the rest reads as zero (depth adds up across engines)
    const int64_t a = std::max<int64_t>(start0, e->org[tid]), z = std::min<int64_t>((int64_t)start0 + n, (int64_t)e->org[tid] + e->elen[tid]);
    if (a > start0 || z < (int64_t)start0 + n) { if (dev) CK(cudaMemsetAsync(out, 0, (size_t)n * 4, e->stream)); else memset(out, 0, (size_t)n * 4); }
    if (z > a) CK(cudaMemcpyAsync(out + (a - start0), e->d_depth + e->off[tid] + (a - e->org[tid]), (size_t)(z - a) * 4, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
    if (!dev) { e->stats.d2h_bytes += (uint64_t)n * 4; CK(cudaStreamSynchronize(e->stream)); }
    return PD_OK;
}

int pd_get_stats(pd_engine *e, pd_stats *out)
{
    if (!e || !out) return PD_ERR_ARG;
    *out = e->stats;
    return PD_OK;
}

int pd_profile_enable(pd_engine *e, int on)
{
    if (!e) return PD_ERR_ARG;
    CK(cudaSetDevice(e->device));
    int rc = drain_profile(e); if (rc) return rc;
    e->profiling = on != 0;
    if (on) { for (int k = 0; k < K_COUNT; ++k) { e->kms[k] = 0; e->kcount[k] = 0; } }
    return PD_OK;
}

int pd_profile_get(pd_engine *e, pd_kernel_time *out, int cap)
{
    if (!e) return PD_ERR_ARG;
    CK(cudaSetDevice(e->device));
    int rc = drain_profile(e); if (rc) return -rc;
    for (int k = 0; k < K_COUNT && k < cap; ++k) {
        memset(out[k].name, 0, sizeof out[k].name);
        strncpy(out[k].name, kKernelNames[k], sizeof out[k].name - 1);
        out[k].launches = e->kcount[k]; out[k].ms = e->kms[k];
    }
    return K_COUNT;
}

}  // extern "C"
