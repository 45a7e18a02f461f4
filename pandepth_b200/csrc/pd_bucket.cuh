// pd_bucket.cuh -- the unsorted path at scale: expand -> bucket -> tile (sm_100a).
//
// Role in the reference: the same walk as everywhere (filter :732-733, CIGAR switch :735-760), for input that is not in
// coordinate order (unsorted SAM/BAM streams :4679-4710, out-of-order batches).  `k_delta_scatter` turns every run end into
// a global reduction at a random address; measured on B200 those run at ~25 G/s wherever their lines live (the L2's atomic
// units are the limit), i.e. 3.7 % of the HBM roofline for 111 M shuffled reads.  The fix is fewer global reductions: bring
// the batch into *bucket* order and let the tile kernel, which accumulates in shared memory, do the rest.
//
//   k_expand_runs     filter + CIGAR walk of every read -> aligned runs (start, len) relative to the owned range, clipped to it
//                     (one run per read for single-op input), written into the slots of the read's own CIGAR words
//   k_bucket_hist     histogram of the runs over coarse buckets of 2^19 bases (shared-memory pre-aggregation per CTA)
//   k_bucket_scan     exclusive scan of the <= 4096 coarse counts (one CTA)
//   k_bucket_scatter  runs -> coarse-bucket order (a CTA reserves, per coarse bucket, a range for its tile's runs)
//   k_bucket_fine     one CTA per coarse bucket: counts its runs over 2048 fine buckets of 256 bases (= the tile kernel's
//                     look-back), scans them in shared memory, publishes the fine offsets and scatters the runs -> fine order
//   k_bucket_cuts     lo[] / hi[] of every 2048-base segment straight from the fine offsets (no searches)
//
// Runs inside a fine bucket stay in arbitrary order: `k_tile_accumulate` only needs every run that starts in
// [s0 - 256, s0 + 2048) inside [lo[j], hi[j + 1]) -- its table updates are atomic, order-free -- so it is launched with
// TileArgs::unordered (the sortedness checks off).  Every pass moves whole arrays with coalesced reads; the scattered
// writes go out in short bursts per bucket.  Traffic: 16 B/read (expand) + 24 B/run (coarse) + 20 B/run (fine) + the
// tile kernel's 8 B/run + 4 B/base.
#pragma once
#ifndef PD_WARP_EMU              // tests/cpp/bucket_emu_check.cpp runs these kernels on the CPU (tests/cpp/warp_emu.h)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace pdk {

constexpr int BK_COARSE_SHIFT = 19;                 // coarse bucket = 2^19 bases
constexpr int BK_FINE_SHIFT = 8;                    // fine bucket = 256 bases = HALO
constexpr int BK_FINE = 1 << (BK_COARSE_SHIFT - BK_FINE_SHIFT);   // 2048 fine buckets per coarse bucket
constexpr int BK_MAX_COARSE = 4096;                 // 2^31 bases
constexpr int BK_TILE = 65536;                      // slots per CTA in the coarse histogram
static_assert(HALO == (1 << BK_FINE_SHIFT) && SEG == 8 * HALO, "fine buckets are the tile kernel's look-back; eight of them make a segment");

struct ExpandArgs {
    const int32_t *pos; const uint32_t *cigar; const uint32_t *cigar_off;
    const uint16_t *flag; const uint8_t *mapq;
    uint32_t n; int32_t contig_len; int32_t shift;
    uint32_t xflags; int32_t min_mapq;
    int32_t *run_pos; uint32_t *run_cig;            // out: start (>= 0, < contig_len) and (len << 4 | M); run_pos = -1: empty slot
    uint32_t cap;                                   // slots: n for single-op batches, the batch's CIGAR words otherwise
    uint32_t *delta; int *delta_used;               // contig base of the delta array: runs of 2^28 bases or more (no CIGAR word holds them) go there
};

// No cursor, no atomics: read i owns the slots of its own CIGAR words (slot i for a single-op batch) -- a run consumes at
// least one word, so its runs fit -- and marks the slots it does not fill as empty; the bucket passes skip those.
// (A first version appended runs through one global cursor: 3.5 M same-address atomics, most of an 11.5 ms pass.)
__global__ void __launch_bounds__(256) k_expand_runs(ExpandArgs a)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    bool live = true;
    if (a.flag && (a.flag[i] & a.xflags)) live = false;
    if (live && a.mapq && (int)a.mapq[i] < a.min_mapq) live = false;
    const int32_t p = __ldcs(a.pos + i);
    if (p < 0) live = false;
    uint32_t c0 = i, nops = 1, slot = i;
    if (a.cigar_off) { c0 = a.cigar_off[i]; nops = a.cigar_off[i + 1] - c0; slot = c0 - a.cigar_off[0]; }
    const uint32_t slot_end = slot + nops;
    const int64_t clen = a.contig_len;
    int64_t q = (int64_t)p - a.shift, rs = 0, re = 0; bool open = false;
    auto emit = [&]() {
        int64_t es = rs < 0 ? 0 : rs, ee = re > clen ? clen : re;
        if (es >= ee) return;
        if (ee - es >= (1ll << 28)) {                            // does not fit a (len << 4 | op) word: straight to the delta array
            red_add_u32(a.delta + es, 1u); red_add_u32(a.delta + ee, 0xffffffffu); raise_flag(a.delta_used);
            return;
        }
        if (slot < a.cap) { a.run_pos[slot] = (int32_t)es; a.run_cig[slot] = (uint32_t)(ee - es) << 4; }
        slot++;
    };
    if (live) {
        for (uint32_t k = 0; k < nops; ++k) {
            const uint32_t w = __ldcs(a.cigar + c0 + k);
            const uint32_t op = w & 15u; const int64_t n = w >> 4;
            if (is_match_op(op)) {                               // :745-752  adjacent M/=/X merge into one run
                if (open && re == q) re = q + n; else { if (open) emit(); rs = q; re = q + n; open = true; }
                q += n;
            } else if (is_skip_op(op)) q += n;                   // :754-757
        }
        if (open) emit();
    }
    for (; slot < slot_end && slot < a.cap; slot++) a.run_pos[slot] = -1;
}

// coarse histogram: hist[b] += runs with (pos >> 19) == b
__global__ void __launch_bounds__(256) k_bucket_hist(const int32_t *__restrict__ pos, uint32_t n, uint32_t *__restrict__ hist, int n_coarse)
{
    __shared__ uint32_t h[BK_MAX_COARSE];
    for (int i = threadIdx.x; i < n_coarse; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * BK_TILE + threadIdx.x, e = min(n, (blockIdx.x + 1u) * BK_TILE); i < e; i += blockDim.x) {
        const int32_t p = pos[i];
        if (p >= 0) atomicAdd(&h[(uint32_t)p >> BK_COARSE_SHIFT], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_coarse; i += blockDim.x) if (h[i]) atomicAdd(&hist[i], h[i]);
}

// base[b] = exclusive prefix of hist (base[n_coarse] = total); cursor[b] = 0.  One CTA of 1024 threads, <= 4096 bins.
__global__ void __launch_bounds__(1024) k_bucket_scan(const uint32_t *__restrict__ hist, uint32_t *__restrict__ base, uint32_t *__restrict__ cursor, int n_coarse)
{
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int i = t * 4 + k; v[k] = i < n_coarse ? hist[i] : 0u; s += v[k]; }
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { const uint32_t y = t >= o ? part[t - o] : 0u; __syncthreads(); part[t] += y; __syncthreads(); }
    uint32_t ex = part[t] - s;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int i = t * 4 + k; if (i < n_coarse) { base[i] = ex; cursor[i] = 0; } ex += v[k]; }
    if (t == 1023) base[n_coarse] = part[1023];
}

// Exclusive scan of cnt[0..n) (n <= 4096, a multiple of nothing in particular) by the whole CTA: start[i] = sum of cnt[< i];
// returns the total.  `part` holds blockDim.x words.  cnt is left untouched.
__device__ __forceinline__ uint32_t bk_cta_scan(const uint32_t *cnt, uint32_t *start, int n, uint32_t *part)
{
    const int t = threadIdx.x, T = blockDim.x;
    const int per = (n + T - 1) / T, a = t * per, z = min(n, a + per);
    uint32_t s = 0;
    for (int i = a; i < z; i++) s += cnt[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < T; o <<= 1) { const uint32_t y = t >= o ? part[t - o] : 0u; __syncthreads(); part[t] += y; __syncthreads(); }
    uint32_t ex = part[t] - s;
    for (int i = a; i < z; i++) { start[i] = ex; ex += cnt[i]; }
    const uint32_t total = part[T - 1];
    __syncthreads();
    return total;
}

constexpr int BK_STAGE = 8192;                      // runs staged in shared memory per step (64 KB of pos + cigar)

// runs -> coarse order, staged: a CTA sorts a tile of BK_STAGE slots by coarse bucket INSIDE shared memory, reserves one range per
// bucket with one global atomic, and writes the tile out bucket by bucket -- neighbouring threads write neighbouring addresses,
// so a sector is written at once.  (Placing every run straight into global memory -- 4-byte stores at 636 moving write fronts
// per CTA, more open sectors than the L2 holds -- took 5.2 ms per 111 M runs.)
// dynamic shared memory: cnt[4096] | start[4096] | gbase[4096] | part[256] | stage_pos[BK_STAGE] | stage_cig[BK_STAGE]
constexpr size_t BK_SCATTER_SMEM = (3 * BK_MAX_COARSE + 256 + 2 * BK_STAGE) * 4;
__global__ void __launch_bounds__(256) k_bucket_scatter(const int32_t *__restrict__ pos, const uint32_t *__restrict__ cig, uint32_t n,
                                                        const uint32_t *__restrict__ base, uint32_t *__restrict__ cursor, int n_coarse,
                                                        int32_t *__restrict__ out_pos, uint32_t *__restrict__ out_cig)
{
#ifndef PD_WARP_EMU
    extern __shared__ uint32_t bk_sm[];
#else
    uint32_t *bk_sm = emu::dyn_smem();
#endif
    uint32_t *cnt = bk_sm, *start = cnt + BK_MAX_COARSE, *gbase = start + BK_MAX_COARSE, *part = gbase + BK_MAX_COARSE;
    int32_t *st_pos = reinterpret_cast<int32_t *>(part + 256); uint32_t *st_cig = reinterpret_cast<uint32_t *>(st_pos + BK_STAGE);
    const uint32_t t0 = blockIdx.x * BK_STAGE, t1 = min(n, t0 + BK_STAGE);
    for (int i = threadIdx.x; i < n_coarse; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (uint32_t i = t0 + threadIdx.x; i < t1; i += blockDim.x) { const int32_t p = pos[i]; if (p >= 0) atomicAdd(&cnt[(uint32_t)p >> BK_COARSE_SHIFT], 1u); }
    __syncthreads();
    const uint32_t total = bk_cta_scan(cnt, start, n_coarse, part);
    for (int i = threadIdx.x; i < n_coarse; i += blockDim.x) { const uint32_t c = cnt[i]; gbase[i] = c ? base[i] + atomicAdd(&cursor[i], c) : 0u; cnt[i] = start[i]; }
    __syncthreads();
    for (uint32_t i = t0 + threadIdx.x; i < t1; i += blockDim.x) {
        const int32_t p = pos[i];
        if (p < 0) continue;                                     // an empty slot
        const uint32_t k = atomicAdd(&cnt[(uint32_t)p >> BK_COARSE_SHIFT], 1u);
        st_pos[k] = p; st_cig[k] = cig[i];
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < total; k += blockDim.x) {
        const int32_t p = st_pos[k];
        const uint32_t b = (uint32_t)p >> BK_COARSE_SHIFT, g = gbase[b] + (k - start[b]);
        out_pos[g] = p; out_cig[g] = st_cig[k];
    }
}

// one CTA per coarse bucket: fine order inside it + the fine offsets (fine_off[b * 2048 + f] = first run of fine bucket f).
// Pass A counts the bucket's runs per fine bucket (-> offsets); pass B goes through them BK_STAGE at a time, sorts each step by
// fine bucket in shared memory and appends it to the fine buckets' running cursors in whole bursts.
// dynamic shared memory: gcur[2048] | cnt[2048] | start[2048] | cur[2048] | part[512] | stage_pos[BK_STAGE] | stage_cig[BK_STAGE]
constexpr size_t BK_FINE_SMEM = (4 * BK_FINE + 512 + 2 * BK_STAGE) * 4;
__global__ void __launch_bounds__(512) k_bucket_fine(const int32_t *__restrict__ pos, const uint32_t *__restrict__ cig, const uint32_t *__restrict__ base,
                                                     uint32_t *__restrict__ fine_off, int32_t *__restrict__ out_pos, uint32_t *__restrict__ out_cig)
{
#ifndef PD_WARP_EMU
    extern __shared__ uint32_t bk_sm[];
#else
    uint32_t *bk_sm = emu::dyn_smem();
#endif
    uint32_t *gcur = bk_sm, *cnt = gcur + BK_FINE, *start = cnt + BK_FINE, *cur = start + BK_FINE, *part = cur + BK_FINE;
    int32_t *st_pos = reinterpret_cast<int32_t *>(part + 512); uint32_t *st_cig = reinterpret_cast<uint32_t *>(st_pos + BK_STAGE);
    const uint32_t b = blockIdx.x, r0 = base[b], r1 = base[b + 1];
    const int t = threadIdx.x;
    for (int i = t; i < BK_FINE; i += 512) cnt[i] = 0;
    __syncthreads();
    for (uint32_t i = r0 + t; i < r1; i += 512) atomicAdd(&cnt[((uint32_t)pos[i] >> BK_FINE_SHIFT) & (BK_FINE - 1)], 1u);
    __syncthreads();
    bk_cta_scan(cnt, gcur, BK_FINE, part);
    for (int i = t; i < BK_FINE; i += 512) { gcur[i] += r0; fine_off[(size_t)b * BK_FINE + i] = gcur[i]; }
    __syncthreads();
    for (uint32_t s0 = r0; s0 < r1; s0 += BK_STAGE) {
        const uint32_t s1 = min(r1, s0 + BK_STAGE);
        for (int i = t; i < BK_FINE; i += 512) cnt[i] = 0;
        __syncthreads();
        for (uint32_t i = s0 + t; i < s1; i += 512) atomicAdd(&cnt[((uint32_t)pos[i] >> BK_FINE_SHIFT) & (BK_FINE - 1)], 1u);
        __syncthreads();
        bk_cta_scan(cnt, start, BK_FINE, part);
        for (int i = t; i < BK_FINE; i += 512) cur[i] = start[i];
        __syncthreads();
        for (uint32_t i = s0 + t; i < s1; i += 512) {
            const int32_t p = pos[i];
            const uint32_t k = atomicAdd(&cur[((uint32_t)p >> BK_FINE_SHIFT) & (BK_FINE - 1)], 1u);
            st_pos[k] = p; st_cig[k] = cig[i];
        }
        __syncthreads();
        for (uint32_t k = t; k < s1 - s0; k += 512) {
            const int32_t p = st_pos[k];
            const uint32_t f = ((uint32_t)p >> BK_FINE_SHIFT) & (BK_FINE - 1), g = gcur[f] + (k - start[f]);
            out_pos[g] = p; out_cig[g] = st_cig[k];
        }
        __syncthreads();
        for (int i = t; i < BK_FINE; i += 512) gcur[i] += cnt[i];
        __syncthreads();
    }
}

// lo[j] = first run with pos >= s0_j - 256, hi[j] = first run with pos >= s0_j (nseg + 1 entries), from the fine offsets
__global__ void __launch_bounds__(256) k_bucket_cuts(const uint32_t *__restrict__ fine_off, const uint32_t *__restrict__ n_runs /* = &base[n_coarse] */, uint32_t n_fine_total,
                                                     int32_t seg0, int32_t nseg, uint32_t *__restrict__ lo, uint32_t *__restrict__ hi)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > nseg) return;
    const uint32_t total = *n_runs;
    const int64_t f = (int64_t)(seg0 + j) * 8;                   // first fine bucket of segment j
    auto off = [&](int64_t k) { return k <= 0 ? 0u : (k >= (int64_t)n_fine_total ? total : fine_off[k]); };
    if (j < nseg) lo[j] = off(f - 1);
    hi[j] = off(f);
}

}  // namespace pdk
