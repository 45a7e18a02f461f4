// pd_kernels.cuh -- sm_100a kernels of the per-base depth engine.
//
// Roles (reference = /root/reference/src/PanDepth.cpp):
//   k_cut              index a sorted batch by genomic segment (no CPU counterpart; replaces the BAI window queries :726)
//   k_tile_accumulate  filter + CIGAR walk -> per-base depth, one warp per 2048-base segment   (:732-760)
//   k_delta_scatter    same walk for unsorted / long-span batches: +-1 deltas with global reductions (:732-760)
//   k_scan_add         prefix-sum of the delta array folded into depth (the CPU increments every base instead)
//   k_fill_untouched   zero-fill segments no batch touched (the `depth[ccv]=0` loop :718-721, only where needed)
//   k_reduce_units     interval CoveredSite/TotalDepth                                         (:295-348)
//   k_reduce_windows   fixed-stride windows (mode 6 loop)                                      (:4366-4381)
//
// All arithmetic is integer (uint32 modular depth, uint64 sums); no tensor cores are involved.
#pragma once
#ifndef PD_WARP_EMU              // tests/cpp/{tile,reduce}_emu_check.cpp run these kernels on the CPU (tests/cpp/warp_emu.h)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace pdk {

#ifndef PD_SEG
#define PD_SEG 2048
#endif
#ifndef PD_TILE_UNROLL
#define PD_TILE_UNROLL 16
#endif
#ifndef PD_TILE_MIN_CTAS
#define PD_TILE_MIN_CTAS 3
#endif
constexpr int SEG = PD_SEG;   // bases owned by one warp of k_tile_accumulate (8 KB of depth at 2048)
constexpr int HALO = 256;   // look-back window in bases; also the largest "common span" L the fast path models
constexpr int WPC = 8;      // warps per CTA in k_tile_accumulate
#ifndef PD_TILE_AGG
#define PD_TILE_AGG 0       // 1: runs of equal adjacent reads of a chunk are applied once with a weight.  Measured on B200 (one cfg2 contig,
                            // profiles/r02_tile_agg.txt): 50x 0.465 -> 0.580 ms, 200x 1.420 -> 1.829 ms -- the shuffle + two ballots per chunk
                            // cost more than the 15 % / 45 % of the shared atomics they save; off
#endif
#ifndef PD_PREFETCH
#define PD_PREFETCH 1      // every warp prefetches (to L2) the reads of the segment one wave of resident warps ahead
#endif
constexpr int WIN = SEG + HALO;              // bases a warp sees reads from: [s0 - HALO, s0 + SEG); "window index" = pos - (s0 - HALO)
constexpr int TOFF = HALO;                   // window index of table entry 0: the table covers the segment itself
constexpr int TLEN = WIN - TOFF;             // table entries (= SEG)
constexpr int TILE_SMEM_WORDS = TLEN;        // per warp: 8 KB at SEG 2048
constexpr int TILE_QUEUE_WORDS = 3 * 64;     // per warp, HAS_OFF kernels only: the ring of queued multi-op reads (768 B)
constexpr int TILE_MIN_CTAS = PD_TILE_MIN_CTAS;   // resident CTAs per SM k_tile_accumulate is compiled for (register cap)
constexpr unsigned FULL = 0xffffffffu;

enum { ERR_UNSORTED = 1, ERR_RANGE = 2 };

__device__ __forceinline__ bool is_match_op(uint32_t op) { return (0x181u >> op) & 1u; }               // M = X  (:745-747)
__device__ __forceinline__ bool is_skip_op(uint32_t op) { return (0x00cu >> op) & 1u; }                // D N    (:754-755)

__device__ __forceinline__ int32_t ldg_i32(const int32_t *p) { return __ldg(p); }
__device__ __forceinline__ uint32_t ldg_u32(const uint32_t *p) { return __ldg(p); }

__device__ __forceinline__ void red_add_u32(uint32_t *addr, uint32_t v)
{
#ifndef PD_WARP_EMU
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
#else
    atomicAdd(addr, v);
#endif
}

// many threads may raise the same host-visible flag with the same value: a plain store on the device; the emulated build makes it
// an atomic store so that ThreadSanitizer (tests/test_emu_tsan_cpu.py) reports only races that matter
__device__ __forceinline__ void raise_flag(int *p)
{
#ifndef PD_WARP_EMU
    *p = 1;
#else
    __atomic_store_n(p, 1, __ATOMIC_RELAXED);
#endif
}

// one Kogge-Stone step of a warp inclusive scan: t += (lane >= o) ? shfl_up(t, o) : 0, using the predicate shfl.up
// itself produces (2 instructions instead of shuffle + compare + select + add)
__device__ __forceinline__ uint32_t scan_step_up(uint32_t t, int o)
{
#ifndef PD_WARP_EMU
    asm volatile("{ .reg .u32 y; .reg .pred p;\n\t"
                 "shfl.sync.up.b32 y|p, %0, %1, 0, 0xffffffff;\n\t"
                 "@p add.u32 %0, %0, y; }" : "+r"(t) : "r"(o));
#else
    const uint32_t y = __shfl_up_sync(0xffffffffu, t, (unsigned)o);
    if (emu::tl_lane >= o) t += y;
#endif
    return t;
}

// (x != 0) + (y != 0) + (z != 0) + (w != 0) as four min.u32 and three-input adds (the compiler's own form is a serial
// compare / add / conditional-move chain, three instructions per element)
__device__ __forceinline__ uint32_t nonzero4(const uint4 v)
{
#ifndef PD_WARP_EMU
    uint32_t r;
    asm("{ .reg .u32 a, b, c, d;\n\t"
        "min.u32 a, %1, 1; min.u32 b, %2, 1; min.u32 c, %3, 1; min.u32 d, %4, 1;\n\t"
        "add.u32 a, a, b; add.u32 c, c, d; add.u32 %0, a, c; }" : "=r"(r) : "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
    return r;
#else
    return (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// k_cut: for segment j (bases [s0, s0+SEG)) of a pos-sorted batch: lo[j] = first read with pos >= s0-HALO,
// hi[j] = first read with pos >= s0 (so segment j's reads end at hi[j+1]).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lower_bound_pos(const int32_t *pos, uint32_t n, int64_t key)
{
    uint32_t a = 0, b = n;
    while (a < b) {
        uint32_t m = a + ((b - a) >> 1);
        if ((int64_t)ldg_i32(pos + m) < key) a = m + 1; else b = m;
    }
    return a;
}

struct CutArgs {
    const int32_t *pos;
    uint32_t n; int32_t seg0; int32_t nseg;
    uint32_t *lo; uint32_t *hi;
    int64_t shift;      // contig coordinate of the engine's first owned base (range sharding): segment s covers [shift + s*SEG, ...)
};

constexpr int CUT_COARSE = 64;   // segments per coarse search key
constexpr int CUT_THREADS = 64;  // small CTAs: the cut kernels of batch k+1 run UNDER batch k's tile kernel (own stream) and must fit the
                                 // few registers three resident tile CTAs leave free on an SM instead of displacing one of them

// lower_bound(pos, n, key) found by galloping outwards from a guess, then bisecting the bracket: O(log distance) probes,
// all close to the answer (the plain bisection touches ~30 different 32-byte sectors per key).
__device__ __forceinline__ uint32_t lower_bound_near(const int32_t *pos, uint32_t n, int64_t key, uint32_t g)
{
    if (g > n) g = n;
    uint32_t a, b, step = 32;
    if (g < n && (int64_t)ldg_i32(pos + g) < key) {            // answer in (g, n]
        a = g + 1;
        while (true) {
            b = (n - a > step) ? a + step : n;
            if (b == n || (int64_t)ldg_i32(pos + b) >= key) break;
            a = b + 1; step <<= 1;
        }
    } else {                                                   // answer in [0, g]
        b = g;
        while (true) {
            if (b == 0) { a = 0; break; }
            a = (b > step) ? b - step : 0;
            if ((int64_t)ldg_i32(pos + a) < key) { a = a + 1; break; }
            b = a; step <<= 1;
        }
    }
    return a + lower_bound_pos(pos + a, b - a, key);
}

// pass A: one thread per CUT_COARSE segments: coarse[k] = lower_bound(pos, (seg0 + k*CUT_COARSE)*SEG - HALO), searched outwards
// from the position a uniform batch would have it at (a handful of dependent probes instead of the ~27 of a plain bisection
// over 10^8 reads; any guess gives the same answer -- lower_bound_near is exact)
__global__ void __launch_bounds__(CUT_THREADS) k_cut_coarse(CutArgs a, uint32_t *coarse, int32_t ncoarse)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncoarse) return;
    const int64_t key = (int64_t)(a.seg0 + (int64_t)k * CUT_COARSE) * SEG - HALO + a.shift;
    const uint32_t g = (uint32_t)(((uint64_t)a.n * (uint32_t)k) / (uint32_t)ncoarse);
    coarse[k] = lower_bound_near(a.pos, a.n, key, g);
}

// pass B: thread j < nseg: lo[j] = lower_bound(s0_j - HALO) from an interpolated guess inside its coarse bracket, and
// mid[j] = lower_bound(s0_j) a few reads to its right; thread nseg: mid[nseg].  Segment j's reads are [lo[j], mid[j+1]).
__global__ void __launch_bounds__(CUT_THREADS) k_cut(CutArgs a, const uint32_t *coarse, int32_t ncoarse)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > a.nseg) return;
    const int64_t s0 = (int64_t)(a.seg0 + j) * SEG + a.shift;
    int k = j / CUT_COARSE; if (k >= ncoarse) k = ncoarse - 1;
    const uint32_t ca = coarse[k];
    const uint32_t cb = (k + 1 < ncoarse) ? coarse[k + 1] : a.n;
    const uint32_t g = ca + (uint32_t)(((uint64_t)(cb - ca) * (uint32_t)(j - k * CUT_COARSE)) / CUT_COARSE);
    const uint32_t l = lower_bound_near(a.pos, a.n, s0 - HALO, g);
    if (j < a.nseg) a.lo[j] = l;
    a.hi[j] = lower_bound_near(a.pos, a.n, s0, l);             // "mid": nseg+1 entries
}

// ---------------------------------------------------------------------------------------------------------------
// k_tile_accumulate: one warp per SEG-base segment of one contig.
//
// The warp sees every read whose pos lies in [s0-HALO, s0+SEG) (lo/hi from k_cut) and builds, in its private 8 KB of
// shared memory, E[b] = (runs starting at b) - (runs ending at b) for the bases of its segment with shared-memory atomics
// (warp-private table, so the only contention is equal positions inside one 32-read chunk); every +-1 that lands left of
// s0 goes to a per-lane register carry instead.  Loads of TILE_UNROLL chunks are issued together before any is consumed.
// Runs that reach past the segments that see the read (pos + HALO) are "far": the owner segment scatters them as +-1
// deltas into the global delta array, folded in by k_scan_add at finalize.
// The segment is then prefix-summed (warp shuffles, two rows of 128 bases interleaved, row totals by redux.sync) and written
// with coalesced 16-byte stores: a plain streaming store if no earlier batch touched the segment (zero-initialisation is
// fused away), else load+add+store.  Each warp also prefetches, into L2, the reads of the segment one wave of resident
// warps ahead (exact range from the cut arrays).
// (Measured alternative, dropped: modelling reads as one common span and applying group counts with plain two-phase
// read-modify-writes was 15% slower than the atomics on B200 -- see DESIGN.md.)
// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// k_filter_bits: the record filter (:732-733) of a batch that brings flag[] / mapq[] arrays, as one bit per read.
// The tile kernel then reads one word per 32-read chunk instead of a 2-byte and a 1-byte load per read (those loads were
// consumed chunk by chunk and serialised sixteen DRAM round trips per group: 1.07 ms instead of 0.46 ms per cfg2 contig).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_filter_bits(const uint16_t *__restrict__ flag, const uint8_t *__restrict__ mapq, uint32_t n,
                                                     uint32_t xflags, int32_t min_mapq, uint32_t *__restrict__ bits, uint32_t n_words)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;     // whole warps: n_words * 32 threads
    if ((i >> 5) >= n_words) return;
    bool pass = false;
    if (i < n) {
        pass = true;
        if (flag && (__ldcs(flag + i) & xflags)) pass = false;                   // :732
        if (mapq && (int)__ldcs(mapq + i) < min_mapq) pass = false;              // :733
    }
    const uint32_t m = __ballot_sync(FULL, pass);
    if ((threadIdx.x & 31) == 0) bits[i >> 5] = m;
}

struct TileArgs {
    const int32_t *pos; const uint32_t *cigar; const uint32_t *cigar_off;
    const uint16_t *flag; const uint8_t *mapq;
    const uint32_t *pass_bits;   // HAS_FILTER: bit i = read i passes the record filter (k_filter_bits), n/32 + 2 words
    const uint32_t *lo; const uint32_t *hi;
    uint32_t n;
    int32_t seg0; int32_t nseg; int32_t contig_len;
    uint32_t *depth;      // contig base
    uint32_t *delta;      // contig base (global delta array)
    uint8_t *seg_state;   // contig base: 0 = never written, 1 = written once and part_* valid, 2 = written (read depth)
    uint32_t xflags; int32_t min_mapq;
    int *err; int *delta_used;
    // reduce hint (pd_set_reduce_hint): per-segment CoveredSite/TotalDepth partials for segments written exactly once
    uint32_t hint_on, hint_min_dep, hint_mask;
    uint32_t *part_cov; unsigned long long *part_sum;     // contig base, one entry per segment
    int32_t pf_dist;      // PD_PREFETCH: segments between a warp's own segment and the one whose reads it prefetches
    int32_t shift;        // contig coordinate of depth[0] (range sharding: this engine owns [shift, shift + contig_len) of the contig)
    uint32_t unordered;   // the reads of [lo[j], hi[j+1]) are those of segment j's window but in no particular order (bucketed
                          // unsorted batches, pd_bucket.cuh): the sortedness checks are off, the window check stays
    const uint32_t *n_dev;   // non-null: the number of reads lives on the device (`n` is only an upper bound)
};

constexpr int TILE_UNROLL = PD_TILE_UNROLL;   // 32-read chunks whose loads are issued together (memory-level parallelism)

#ifndef PD_TILE_BULK
#define PD_TILE_BULK 2       // TMA 1-D bulk flush of a finished segment from shared memory.  Measured on B200 (tools/bench_tile.py,
#endif                       // one cfg2 contig, profiles/r02_tile_ab.txt):
                             //   0  register stores; a segment that already holds depth is load+add+store   first write 0.455 ms, add pass 0.652 ms
                             //   1  cp.async.bulk for first writes too (8 KB shared -> global per segment)    first write 0.478 ms, add pass 0.550 ms
                             //   2  (default) first writes stay register stores; segments that already hold depth are ADDED by the copy
                             //      engine (cp.reduce.async.bulk .add.u32): no 4 B/base re-read of depth         first write 0.457 ms, add pass 0.550 ms

// TMA 1-D bulk copy shared::cta -> global of one segment (8 KB), issued by one lane; `add` folds into what is there
__device__ __forceinline__ void bulk_flush_segment(uint32_t *gdst, const uint32_t *ssrc, bool add)
{
#ifndef PD_WARP_EMU
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the generic-proxy stores of the scan are visible to the async proxy
    if (add) asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], %2;" :: "l"(gdst), "r"(sa), "n"(SEG * 4) : "memory");
    else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(sa), "n"(SEG * 4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory may be reused / the CTA may retire
#else
    for (int i = 0; i < SEG; i++) { if (add) gdst[i] += ssrc[i]; else gdst[i] = ssrc[i]; }
#endif
}

template <bool HAS_OFF, bool HAS_FILTER>
__global__ void __launch_bounds__(WPC * 32, TILE_MIN_CTAS) k_tile_accumulate(TileArgs a)
{
    constexpr int U = HAS_OFF ? (TILE_UNROLL > 8 ? 8 : TILE_UNROLL) : TILE_UNROLL;   // offsets cost two more registers per chunk
#ifndef PD_WARP_EMU
    extern __shared__ __align__(16) uint32_t smem[];
#else
    uint32_t *smem = emu::dyn_smem();
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * WPC + warp;
    if (j >= a.nseg) return;
    const uint32_t i_lo = a.lo[j], i_hi = a.hi[j + 1];        // hi[] holds lower_bound(s0_j): nseg+1 entries
    const uint32_t n_all = a.n_dev ? __ldg(a.n_dev) : a.n;
#if PD_PREFETCH
    // the reads of segment j + pf_dist (the segment that takes this warp's place when it retires) are pulled into L2 now,
    // so that warp finds them there instead of in DRAM: one 128-byte line (32 reads) per lane and instruction
    {
        const int jp = j + a.pf_dist;
        if (jp < a.nseg) {
            const uint32_t plo = a.lo[jp], phi = a.hi[jp + 1];
            if (plo < phi && phi <= n_all && phi - plo <= 32u * 32u * 8u) {
                for (uint32_t q = (plo & ~31u) + lane * 32; q < phi; q += 32 * 32) {
#ifndef PD_WARP_EMU
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(a.pos + q));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"((HAS_OFF ? a.cigar_off : a.cigar) + q));
#endif
                }
            }
        }
    }
#endif
    const int64_t s0 = (int64_t)(a.seg0 + j) * SEG;
    const int64_t hbase = s0 - HALO;
    const int64_t send = s0 + SEG;
    // PD_SORTED is verified on everything the kernel touches.  The cut must be consistent with its neighbours: reads left
    // of lo are before the window, the read at hi is beyond it, and the next window starts inside ours.  The loads are
    // issued here and checked at the very end, off the critical path.
    int32_t chk_l = INT32_MIN, chk_r = INT32_MAX; uint32_t chk_n = 0;
    bool bad = i_lo > i_hi || i_hi > n_all;
    if (lane == 0 && !bad) {
        if (i_lo > 0) chk_l = ldg_i32(a.pos + i_lo - 1);
        if (i_hi < n_all) chk_r = ldg_i32(a.pos + i_hi);
        if (j + 1 < a.nseg) chk_n = a.lo[j + 1];
    }
    const int64_t shift = a.shift;
    if (i_lo >= i_hi || bad) {
        if (lane == 0 && (bad || (int64_t)chk_l - shift >= hbase || (int64_t)chk_r - shift < send || chk_n > i_hi)) atomicOr(a.err, ERR_UNSORTED);
        return;
    }

    // E[w - TOFF] = (runs starting at window index w) - (runs ending there); window index w = pos - (s0 - HALO) in [0, WIN).
    // Deltas left of the table (w < TOFF: the look-back part) go to a per-lane register; their sum is the depth carried
    // into the segment.  (Keeping them in the table instead, and issuing the atomics unconditionally with per-lane sink
    // words, both measured slower: the LSU data pipe is this kernel's busiest unit -- profiles/r01b_tile_pipe_breakdown.txt.)
    uint32_t *E = smem + warp * TILE_SMEM_WORDS;
    const int64_t clen = a.contig_len;
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *E4 = reinterpret_cast<uint4 *>(E);
#pragma unroll
        for (int k = lane; k < TILE_SMEM_WORDS / 4; k += 32) E4[k] = z;
    }
    __syncwarp();

    uint32_t carry_reg = 0;        // per-lane part of "depth just left of the segment"
    int32_t carry_p = INT32_MIN;   // last position of the previous chunk (sortedness check)
    uint32_t idx_max = 0;          // largest window index seen: anything >= WIN means a read outside the window (inconsistent cut)
    const uint32_t win0 = (uint32_t)(hbase + shift);                      // contig position of window index 0 (modular: hbase >= -HALO)
    // reads that start left of an owned range (segment 0 of a range that does not begin at the contig start) have no owner
    // segment in this engine: segment 0 scatters their far runs itself
    const bool own_halo = (a.seg0 + j) == 0 && shift != 0;
    const int64_t clen_w64 = clen - hbase;                                // contig end as a window index
    const uint32_t clen_w = (uint32_t)(clen_w64 > 0x7fffffff ? 0x7fffffff : clen_w64);
    const int32_t pad_pos = (int32_t)(send - 1 + shift);                  // filler for lanes past the last read: in window, sorts last
    // delta v applies to every base >= window index x (anything at or beyond WIN belongs to a later segment)
    auto emit = [&](int64_t x, uint32_t v) {
        if (x < TOFF) carry_reg += v;
        else if (x < WIN) atomicAdd(&E[(int)x - TOFF], v);
    };
    // Multi-op reads (HAS_OFF) are not walked where they are met -- a per-read loop over the ops with a third of the lanes active
    // and the rest waiting cost 5.4x the single-op time on 30 % multi-op input -- but queued in the warp's ring (64 entries of
    // window index / first CIGAR word / op count behind the CTA's tables) and walked 32 at a time, one read per lane, all lanes busy.
    uint32_t *Q = smem + WPC * TILE_SMEM_WORDS + warp * TILE_QUEUE_WORDS;
    uint32_t qn = 0, qh = 0;
    bool far_flagged = false;                                          // this lane has raised *delta_used already
    // One read's ops, 32-bit window arithmetic: positions saturate at 0x7fffffff (every position at or beyond the contig end
    // -- clen_w <= 0x7fffffff -- is clipped away alike), op lengths are < 2^28.  One flush site: a run is closed when an M/=/X
    // op does not continue it (re != q: a D/N lay between) and by the terminator step k == nops.
    auto walk = [&](uint32_t w, uint32_t c0, uint32_t nops) {
        const bool owner = w >= (uint32_t)HALO || own_halo;
        const uint32_t far_w = (w >= (uint32_t)SEG) ? 2u * SEG + HALO : (uint32_t)(SEG + HALO);
        uint32_t q = w, rs = 0, re = 0; bool open = false;
        // the first four words are loaded together (independent loads: one memory round trip instead of one per op -- the
        // dependent per-op load was the top stall of this path); reads with more ops continue with one load per op
        const uint32_t *cp = a.cigar + c0;
        const uint32_t p0 = ldg_u32(cp), p1 = nops > 1u ? ldg_u32(cp + 1) : 0xfu, p2 = nops > 2u ? ldg_u32(cp + 2) : 0xfu, p3 = nops > 3u ? ldg_u32(cp + 3) : 0xfu;
        auto step = [&](uint32_t cw, bool term) {                                   // term: behind the read's last op
            const uint32_t op = cw & 15u, n = cw >> 4;
            const bool m = !term && is_match_op(op);                                // :745-752  adjacent M/=/X merge into one run
            if (open && (term || (m && re != q))) {
                const uint32_t e = min(re, clen_w);
                if (rs < e) {
                    if (rs < (uint32_t)TOFF) carry_reg += 1u; else if (rs < (uint32_t)WIN) atomicAdd(&E[rs - TOFF], 1u);
                    if (e < (uint32_t)TOFF) carry_reg -= 1u; else if (e < (uint32_t)WIN) atomicAdd(&E[e - TOFF], 0xffffffffu);
                    if (owner && e > far_w) {
                        red_add_u32(a.delta + hbase + max(rs, far_w), 1u);
                        red_add_u32(a.delta + hbase + e, 0xffffffffu);
                        if (!far_flagged) { raise_flag(a.delta_used); far_flagged = true; }
                    }
                }
                open = false;
            }
            if (m) { if (!open) { rs = q; open = true; } q = min(q + n, 0x7fffffffu); re = q; }
            else if (!term && is_skip_op(op)) q = min(q + n, 0x7fffffffu);          // :754-757
        };
        step(p0, nops < 1u);
#pragma unroll 1
        for (uint32_t k = 1; k <= nops; ++k) {
            const uint32_t cw = k == 1u ? p1 : k == 2u ? p2 : k == 3u ? p3 : (k < nops ? ldg_u32(cp + k) : 0xfu);
            step(cw, k >= nops);
        }
    };
    auto drain = [&](uint32_t cnt) {                                   // the oldest `cnt` (<= 32) queued reads, one per lane
        __syncwarp();
        if ((uint32_t)lane < cnt) { const uint32_t s = (qh + (uint32_t)lane) & 63u; walk(Q[s], Q[64 + s], Q[128 + s]); }
        __syncwarp();
        qh = (qh + cnt) & 63u; qn -= cnt;
    };
    for (uint32_t base = i_lo; base < i_hi; base += 32 * U) {
        // issue every load of TILE_UNROLL chunks before consuming any of them
        int32_t pv[U]; uint32_t cv[U]; uint32_t ov[U]; uint32_t nv[U];
        uint32_t failm = 0;        // bit u: fails the record filter
        const uint32_t left = i_hi - base;                    // reads left in this and later chunks (>= 1)
        const int32_t *ppos = a.pos + base + lane;
        const uint32_t *pcig = HAS_OFF ? a.cigar_off + base + lane : a.cigar + base + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool valid = (uint32_t)(u * 32 + lane) < left;
            pv[u] = valid ? ldg_i32(ppos + u * 32) : pad_pos;
            ov[u] = 0; nv[u] = 1; cv[u] = 0;                  // a filler lane carries a zero-length op: never applied
            if (HAS_OFF) {
                ov[u] = valid ? ldg_u32(pcig + u * 32) : 0u;
                nv[u] = valid ? ldg_u32(pcig + u * 32 + 1) : 0u;
            } else {
                cv[u] = valid ? ldg_u32(pcig + u * 32) : 0u;
            }
        }
        if (HAS_FILTER) {
            // filter bits of the group's reads [base, base + 32 U): lane l holds word (base >> 5) + l; chunk u's 32 bits straddle
            // words u and u + 1 (two shuffles and a funnel shift per chunk)
            const uint32_t wi = (base >> 5) + (uint32_t)lane;
            const uint32_t bw = (lane <= U) ? ldg_u32(a.pass_bits + wi) : 0u;     // the array has two words of slack
            const uint32_t sh = base & 31u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t w0 = __shfl_sync(FULL, bw, u), w1 = __shfl_sync(FULL, bw, u + 1);
                const uint32_t x = __funnelshift_r(w0, w1, sh);
                if (!((x >> lane) & 1u)) failm |= 1u << u;                          // :732-733 (filler lanes are never applied anyway)
            }
        }
        if (HAS_OFF) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                nv[u] -= ov[u];
                cv[u] = nv[u] ? ldg_u32(a.cigar + ov[u]) : 0u;
            }
        }
        // sortedness of everything just loaded: one rotate-shuffle per chunk (lane l gets lane l-1; lane 0 gets lane 31,
        // which it compares with the NEXT chunk's first read).  Filler lanes hold the window's last position and sort last.
        uint32_t eqm = 0;          // bit u: this lane's read of chunk u starts where its left neighbour's does (PD_TILE_AGG)
        {
            int32_t prev_last = carry_p;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int32_t r = __shfl_sync(FULL, pv[u], (lane + 31) & 31);
                bad |= !a.unordered && pv[u] < (lane == 0 ? prev_last : r);
                if (PD_TILE_AGG) eqm |= (uint32_t)(pv[u] == r && lane != 0) << u;
                prev_last = r;                                 // meaningful in lane 0 only
            }
            carry_p = prev_last;
        }
        // HAS_OFF: the unrolled chunk loop only queues multi-op reads; when 32 are waiting it is left (u0 = next chunk), the queue is
        // drained at ONE code location below, and the loop is re-entered past the chunks already consumed -- a drain inlined
        // behind every chunk made the kernel nine walks long and instruction-fetch bound (ncu: stall_no_instruction 10 per issue)
        int u0 = 0;
      consume_chunks:
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if ((uint32_t)(u * 32) >= left) break;            // warp-uniform
            if (HAS_OFF) {
                if (u < u0) continue;                         // consumed before the last drain
                if (qn >= 32u) { u0 = u; goto drain_queue; }
            }
            const int32_t p = pv[u];
            const uint32_t c = cv[u];
            const uint32_t nops = HAS_OFF ? nv[u] : 1u;
            const uint32_t w = (uint32_t)p - win0;             // window index: < WIN for a consistent cut
            const bool in_win = w < (uint32_t)WIN;
            idx_max = max(idx_max, w);                         // checked once after the loop
            bool pass = in_win && p >= 0;
            if (HAS_FILTER && ((failm >> u) & 1u)) pass = false;
            {
                // one CIGAR op (the common case): M/=/X is one run [p, p+len)                      :745-752
                const uint32_t e = min(w + (c >> 4), clen_w);                      // len < 2^28: no overflow
                const bool ok = pass && (!HAS_OFF || nops == 1u) && ((0x181u >> (c & 15u)) & 1u) && e > w;
                // PD_TILE_AGG (off, measured slower): sorted input puts equal reads into adjacent lanes (duplicates, pile-ups, deep
                // coverage: 15 % of the reads at 50x, 45 % at 200x); a run of lanes with the same (pos, op) is applied once, by its
                // first lane, with the run length as the weight.
                uint32_t cnt = 1u; bool head = ok;
                if (PD_TILE_AGG) {
                    const uint32_t cl = __shfl_up_sync(FULL, c, 1);
                    const unsigned oks = __ballot_sync(FULL, ok);
                    const bool cont = ok && ((eqm >> u) & 1u) && c == cl && ((oks >> ((lane + 31) & 31)) & 1u);   // continues its left neighbour's run
                    const unsigned conts = __ballot_sync(FULL, cont);
                    const unsigned above = lane == 31 ? 0u : (conts >> (lane + 1));
                    cnt = (uint32_t)__ffs((int)~above);                            // 1 + continuation lanes right behind this one
                    head = ok && !cont;
                }
                const bool s_tab = head && w >= (uint32_t)TOFF;                    // start / end land in the table
                const bool e_tab = head && (e - (uint32_t)TOFF) < (uint32_t)TLEN;
                carry_reg += ((head && w < (uint32_t)TOFF) ? cnt : 0u) - ((head && e < (uint32_t)TOFF) ? cnt : 0u);
                if (s_tab) atomicAdd(&E[w - TOFF], cnt);
                if (e_tab) atomicAdd(&E[e - TOFF], 0u - cnt);
                if (head && e > (uint32_t)WIN) {                                    // the run leaves the segment (7% of 150M reads)
                    // first base (window index) that no segment seeing this read owns: segments see reads up to HALO to their left
                    const uint32_t far_w = (w >= (uint32_t)SEG) ? 2 * SEG + HALO : SEG + HALO;
                    if ((w >= (uint32_t)HALO || own_halo) && e > far_w) {           // rare: run longer than HALO, applied by its owner
                        red_add_u32(a.delta + hbase + far_w, cnt);
                        red_add_u32(a.delta + hbase + e, 0u - cnt);
                        if (!far_flagged) { raise_flag(a.delta_used); far_flagged = true; }
                    }
                }
            }
            if (HAS_OFF) {
                const bool multi = pass && nops > 1u;
                const unsigned mm = __ballot_sync(FULL, multi);
                if (mm) {
                    if (multi) {
                        const uint32_t s = (qh + qn + (uint32_t)__popc(mm & ((1u << lane) - 1u))) & 63u;
                        Q[s] = w; Q[64 + s] = ov[u]; Q[128 + s] = nops;
                    }
                    qn += (uint32_t)__popc(mm);
                }
            }
        }
        u0 = U;                                               // every chunk of the group is consumed
      drain_queue:
        if (HAS_OFF) {
            const bool last_group = base + 32u * U >= i_hi;
            while (qn >= 32u || (qn && u0 == U && last_group)) drain(qn < 32u ? qn : 32u);
            if (u0 < U) goto consume_chunks;
        }
    }
    __syncwarp();

    // depth just left of the segment
    uint32_t carry = __reduce_add_sync(FULL, carry_reg);

    uint8_t *st_ptr = a.seg_state + (a.seg0 + j);
    const bool fresh = (*st_ptr == 0);
    uint32_t *out = a.depth + s0;
    // two rows (128 bases each, a lane holds 4 consecutive bases of either) per call: inclusive prefix of the table plus the
    // running carry; the two shuffle chains interleave, row totals come from redux.sync (no shuffle)
    auto scan_pair = [&](int b, uint4 &vA, uint4 &vB) {
        const uint4 dA = *reinterpret_cast<const uint4 *>(E + b);
        const uint4 dB = *reinterpret_cast<const uint4 *>(E + b + 128);
        const uint32_t a0 = dA.x, a1 = dA.y + a0, a2 = dA.z + a1, a3 = dA.w + a2;
        const uint32_t b0 = dB.x, b1 = dB.y + b0, b2 = dB.z + b1, b3 = dB.w + b2;
        uint32_t exA, exB;
        {
            uint32_t t = a3, u = b3;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { t = scan_step_up(t, o); u = scan_step_up(u, o); }
            exA = t - a3 + carry;
            carry += __reduce_add_sync(FULL, a3);
            exB = u - b3 + carry;
            carry += __reduce_add_sync(FULL, b3);
        }
        vA = make_uint4(a0 + exA, a1 + exA, a2 + exA, a3 + exA);
        vB = make_uint4(b0 + exB, b1 + exB, b2 + exB, b3 + exB);
    };
    // partials only for a first write of a segment that lies completely inside the contig
    const bool do_part = fresh && a.hint_on && send <= clen;
    if (do_part) {
        const uint32_t md = a.hint_min_dep, mk = a.hint_mask;
        uint32_t pc = 0; unsigned long long ps = 0;
        if (md == 1u && mk == 0xffffffffu) {
            // default -d 1 on the uint32 path: covered == non-zero, and zero depths add nothing to the sum
#pragma unroll 2
            for (int r = 0; r < SEG / 128; r += 2) {
                const int b = r * 128 + lane * 4;
                uint4 v, w;
                scan_pair(b, v, w);
#if PD_TILE_BULK == 1
                *reinterpret_cast<uint4 *>(E + b) = v; *reinterpret_cast<uint4 *>(E + b + 128) = w;
#else
                __stcs(reinterpret_cast<uint4 *>(out + b), v);
                __stcs(reinterpret_cast<uint4 *>(out + b + 128), w);
#endif
                pc += nonzero4(v) + nonzero4(w);
                ps += ((unsigned long long)v.x + v.y + v.z + v.w) + ((unsigned long long)w.x + w.y + w.z + w.w);
            }
        } else {
            auto acc = [&](const uint4 v) {                          // if (depth >= minDep) { Cover++; Depth += depth; }  :319-323
                uint32_t x;
                x = v.x & mk; if (x >= md) { pc++; ps += x; }
                x = v.y & mk; if (x >= md) { pc++; ps += x; }
                x = v.z & mk; if (x >= md) { pc++; ps += x; }
                x = v.w & mk; if (x >= md) { pc++; ps += x; }
            };
#pragma unroll 2
            for (int r = 0; r < SEG / 128; r += 2) {
                const int b = r * 128 + lane * 4;
                uint4 v, w;
                scan_pair(b, v, w);
#if PD_TILE_BULK == 1
                *reinterpret_cast<uint4 *>(E + b) = v; *reinterpret_cast<uint4 *>(E + b + 128) = w;
#else
                __stcs(reinterpret_cast<uint4 *>(out + b), v);
                __stcs(reinterpret_cast<uint4 *>(out + b + 128), w);
#endif
                acc(v); acc(w);
            }
        }
#if PD_TILE_BULK == 1
        __syncwarp();
        if (lane == 0) bulk_flush_segment(out, E, false);
#endif
        pc = __reduce_add_sync(FULL, pc);
#pragma unroll
        for (int o = 16; o; o >>= 1) ps += __shfl_xor_sync(FULL, ps, o);
        if (lane == 0) { a.part_cov[a.seg0 + j] = pc; a.part_sum[a.seg0 + j] = ps; *st_ptr = 1; }
    } else if (fresh) {
#pragma unroll 2
        for (int r = 0; r < SEG / 128; r += 2) {
            const int b = r * 128 + lane * 4;
            uint4 v, w;
            scan_pair(b, v, w);
#if PD_TILE_BULK == 1
            *reinterpret_cast<uint4 *>(E + b) = v; *reinterpret_cast<uint4 *>(E + b + 128) = w;
#else
            __stcs(reinterpret_cast<uint4 *>(out + b), v);
            __stcs(reinterpret_cast<uint4 *>(out + b + 128), w);
#endif
        }
#if PD_TILE_BULK == 1
        __syncwarp();
        if (lane == 0) bulk_flush_segment(out, E, false);
#endif
        if (lane == 0) *st_ptr = 2;
    } else {
#if PD_TILE_BULK >= 1
        // the segment already holds depth: the scanned rows are ADDED by the copy engine (no 4 B/base read of depth)
#pragma unroll 2
        for (int r = 0; r < SEG / 128; r += 2) {
            const int b = r * 128 + lane * 4;
            uint4 v, w;
            scan_pair(b, v, w);
            *reinterpret_cast<uint4 *>(E + b) = v; *reinterpret_cast<uint4 *>(E + b + 128) = w;
        }
        __syncwarp();
        if (lane == 0) bulk_flush_segment(out, E, true);
#else
#pragma unroll 1
        for (int r = 0; r < SEG / 128; r += 2) {
            const int b = r * 128 + lane * 4;
            uint4 *dst = reinterpret_cast<uint4 *>(out + b);
            const uint4 oA = dst[0], oB = dst[32];                   // +128 bases
            uint4 v, w;
            scan_pair(b, v, w);
            v.x += oA.x; v.y += oA.y; v.z += oA.z; v.w += oA.w;
            w.x += oB.x; w.y += oB.y; w.z += oB.z; w.w += oB.w;
            dst[0] = v; dst[32] = w;
        }
#endif
        if (lane == 0) *st_ptr = 2;
    }
    bad = __any_sync(FULL, bad || idx_max >= (uint32_t)WIN);
    if (lane == 0 && (bad || (int64_t)chk_l - shift >= hbase || (int64_t)chk_r - shift < send || chk_n > i_hi)) atomicOr(a.err, ERR_UNSORTED);
}

// ---------------------------------------------------------------------------------------------------------------
// k_delta_scatter: unsorted or long-span batches.  One thread per read; every aligned run becomes +1 at its start
// and -1 at its end in the global delta array (red.global.add.u32).
// ---------------------------------------------------------------------------------------------------------------
struct ScatterArgs {
    const int32_t *pos; const uint32_t *cigar; const uint32_t *cigar_off;
    const uint16_t *flag; const uint8_t *mapq;
    uint32_t n; int32_t contig_len;
    uint32_t *delta;   // contig base
    uint32_t xflags; int32_t min_mapq;
    int32_t shift;     // contig coordinate of delta[0] (range sharding); runs are clipped to [0, contig_len) after the shift
    const uint32_t *n_dev;   // non-null: only reads [0, *n_dev) (a device-side count: the reads left of a sorted batch's first cut)
    int *delta_used;         // non-null (n_dev mode): set when a run end was scattered (the engine then runs the scan at finalize)
    int64_t stripe_lo, stripe_hi;   // only run ends at delta indices in [stripe_lo, stripe_hi) are applied by this launch (striped
                                    // scatter: one launch per L2-sized stripe of the delta array; 0 / INT64_MAX = everything)
};

__device__ __forceinline__ void scatter_one(const ScatterArgs &a, uint32_t i)
{
    if (a.flag && (a.flag[i] & a.xflags)) return;
    if (a.mapq && (int)a.mapq[i] < a.min_mapq) return;
    const int32_t p = __ldcs(a.pos + i);                        // streaming loads: the reads pass through L2 once per stripe and
    if (p < 0) return;                                          // must not push the stripe's delta lines out of it
    uint32_t c0 = i, nops = 1;
    if (a.cigar_off) { c0 = a.cigar_off[i]; nops = a.cigar_off[i + 1] - c0; }
    const int64_t clen = a.contig_len;
    int64_t q = (int64_t)p - a.shift, rs = 0, re = 0; bool open = false;
    const int64_t slo = a.stripe_lo, shi = a.stripe_hi;
    auto flush = [&]() {
        if (!open) return;
        const int64_t s = rs < 0 ? 0 : rs, e = re > clen ? clen : re;
        if (s >= e) return;
        if (s >= slo && s < shi) red_add_u32(a.delta + s, 1u);
        if (e >= slo && e < shi) red_add_u32(a.delta + e, 0xffffffffu);
        if (a.delta_used) raise_flag(a.delta_used);                 // (n_dev mode) something was scattered: the scan has to run at finalize
    };
    for (uint32_t k = 0; k < nops; ++k) {
        const uint32_t w = __ldcs(a.cigar + c0 + k);
        const uint32_t op = w & 15u;
        const int64_t n = w >> 4;
        if (is_match_op(op)) {
            if (open && re == q) re = q + n; else { flush(); rs = q; re = q + n; open = true; }
            q += n;
        } else if (is_skip_op(op)) q += n;
    }
    flush();
}

__global__ void __launch_bounds__(256) k_delta_scatter(ScatterArgs a)
{
    if (a.n_dev) {                                              // grid-stride over a device-side count (usually zero)
        // Range sharding: the reads left of the first segment's look-back.  Only those that reach into the owned range scatter
        // anything (none of them for reads shorter than HALO) -- and only then is the flag raised: raising it for every batch
        // with such reads cost ranks > 0 a whole-range scan, a 4 B/base clear and a reduce without partials per step (4.6 of
        // 11.7 ms at cfg5 on 2 GPUs, seen as all-reduce wait on rank 0).
        const uint32_t n = min(*a.n_dev, a.n);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) scatter_one(a, i);
        return;
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    scatter_one(a, i);
}

// ---------------------------------------------------------------------------------------------------------------
// k_scan_add: depth[i] (+)= inclusive_prefix_sum(delta)[i].  Single pass, decoupled look-back
// (tile aggregates published through 64-bit {status,value} descriptors), dynamic tile ids for forward progress.
// Every contig's deltas sum to zero before its padding ends, so one scan over the whole array needs no resets.
// The consumed deltas are NOT zeroed here (that would be a third 4 B/base stream in a bandwidth-bound kernel): the engine
// marks the delta array stale and clears it only if another batch arrives after the scan (pd_engine.cu, delta_stale).
// ---------------------------------------------------------------------------------------------------------------
#ifndef PD_SCAN_THREADS
#define PD_SCAN_THREADS 256
#endif
#ifndef PD_SCAN_ITEMS
#define PD_SCAN_ITEMS 16
#endif
#ifndef PD_SCAN_POLL_LD
#define PD_SCAN_POLL_LD 1      // 1: look-back polls with ld.relaxed.gpu instead of an atomic read-modify-write
#endif
#ifndef PD_SCAN_MIN_CTAS
#define PD_SCAN_MIN_CTAS 8     // 8 resident CTAs/SM (32 registers, 64 warps): measured best of {1 (40 regs, 6 CTAs), 8} x
#endif                         // {items 16, 32} x {threads 128, 256, 512} -- tools/gpu_r01d.sh, profiles/r01c_scan_ab.txt
constexpr int SCAN_THREADS = PD_SCAN_THREADS;
constexpr int SCAN_ITEMS = PD_SCAN_ITEMS;                // uint32 per thread (SCAN_ROWS x uint4)
constexpr int SCAN_ROWS = SCAN_ITEMS / 4;                // rows of 128 elements a warp owns
constexpr int SCAN_WARP_ELEMS = 32 * SCAN_ITEMS;         // consecutive elements of one warp (512 at 16 items)
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;     // 4096 elements = one 16 KB tile
static_assert(SCAN_ITEMS % 4 == 0 && SEG % SCAN_WARP_ELEMS == 0, "a scan warp's elements must lie inside one segment");
static_assert(SCAN_TILE % SEG == 0 || SEG % SCAN_TILE == 0, "scan tiles and segments must nest");

__global__ void __launch_bounds__(SCAN_THREADS, PD_SCAN_MIN_CTAS) k_scan_add(uint32_t *__restrict__ depth, const uint32_t *__restrict__ delta,
                                                           uint64_t n_elems, unsigned long long *desc, unsigned int *counter,
                                                           const uint8_t *__restrict__ seg_state)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_tot[SCAN_THREADS / 32];
    __shared__ uint32_t s_prefix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tbase = (uint64_t)tile * SCAN_TILE;
    if (tbase >= n_elems) return;

    // warp w owns SCAN_WARP_ELEMS consecutive elements: SCAN_ROWS rows of 128 (lane holds 4)
    uint4 v[SCAN_ROWS];
    const uint64_t wbase = tbase + (uint64_t)warp * SCAN_WARP_ELEMS;
    uint32_t wcarry = 0;
#pragma unroll
    for (int r = 0; r < SCAN_ROWS; ++r) {
        const uint64_t e = wbase + r * 128 + lane * 4;      // n_elems is a multiple of SCAN_TILE (engine pads)
        uint4 x = __ldcs(reinterpret_cast<const uint4 *>(delta + e));
        x.y += x.x; x.z += x.y; x.w += x.z;
        uint32_t t = x.w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, t, o);
            if (lane >= o) t += y;
        }
        const uint32_t ex = t - x.w + wcarry;
        x.x += ex; x.y += ex; x.z += ex; x.w += ex;
        v[r] = x;
        wcarry += __shfl_sync(FULL, t, 31);
    }
    if (lane == 0) s_warp_tot[warp] = wcarry;
    __syncthreads();
    uint32_t woff = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const uint32_t t = s_warp_tot[w];
        if (w < warp) woff += t;
        agg += t;
    }
    if (warp == 0) {
        // publish aggregate, then look back
        if (lane == 0) {
            const unsigned long long d = ((unsigned long long)(tile == 0 ? 2u : 1u) << 32) | agg;
            atomicExch(desc + tile, d);
        }
        uint32_t excl = 0;
        if (tile > 0) {
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = look - lane;
                unsigned long long d = 0;
                uint32_t status = 2, val = 0;
                if (idx >= 0) {
#if defined(PD_WARP_EMU)
                    do { d = __atomic_load_n(desc + idx, __ATOMIC_ACQUIRE); status = (uint32_t)(d >> 32); } while (status == 0);
#elif PD_SCAN_POLL_LD
                    do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(d) : "l"(desc + idx) : "memory"); status = (uint32_t)(d >> 32); } while (status == 0);
#else
                    do { d = atomicAdd(desc + idx, 0ull); status = (uint32_t)(d >> 32); } while (status == 0);
#endif
                    val = (uint32_t)d;
                }
                // lanes up to and including the first inclusive descriptor contribute
                const unsigned incl = __ballot_sync(FULL, status == 2);
                const int stop = incl ? (__ffs(incl) - 1) : 32;
                uint32_t contrib = (lane <= stop) ? val : 0u;
                for (int o = 16; o; o >>= 1) contrib += __shfl_xor_sync(FULL, contrib, o);
                excl += contrib;
                if (incl) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(desc + tile, (2ull << 32) | (uint32_t)(excl + agg));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    const uint32_t add = s_prefix + woff;
    // a warp's elements lie inside one segment (SEG is a multiple of SCAN_WARP_ELEMS): segments no tile batch has written yet
    // take a plain store (the engine marks every segment written after this kernel)
    const bool fresh = seg_state[wbase / SEG] == 0;
#pragma unroll
    for (int r = 0; r < SCAN_ROWS; ++r) {
        const uint64_t e = wbase + r * 128 + lane * 4;
        uint4 *dp = reinterpret_cast<uint4 *>(depth + e);
        uint4 o = make_uint4(0, 0, 0, 0);
        if (!fresh) o = *dp;
        o.x += v[r].x + add; o.y += v[r].y + add; o.z += v[r].z + add; o.w += v[r].w + add;
        __stcs(dp, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_fill_untouched: segments never written by a batch become zeros.  A warp looks at the states of 32 segments at once (one
// per lane) and zero-fills, all lanes together, the few that need it -- at 50x every segment has been written and the kernel
// only reads n_seg bytes (one warp per segment cost 0.11 ms per 3 Gb just to find that out).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_untouched(uint32_t *depth, uint8_t *seg_state, uint64_t n_seg,
                                                        uint32_t hint_on, uint32_t *part_cov, unsigned long long *part_sum)
{
    const int lane = threadIdx.x & 31;
    const uint64_t s0 = ((uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
    if (s0 >= n_seg) return;
    const uint64_t s = s0 + lane;
    const bool untouched = s < n_seg && seg_state[s] == 0;
    unsigned todo = __ballot_sync(FULL, untouched);
    if (untouched) {
        if (hint_on) { part_cov[s] = 0; part_sum[s] = 0; seg_state[s] = 1; }     // min_dep >= 1: zero depth covers nothing
        else seg_state[s] = 2;
    }
    const uint4 z = make_uint4(0, 0, 0, 0);
    while (todo) {
        const int k = __ffs(todo) - 1; todo &= todo - 1;
        uint4 *p = reinterpret_cast<uint4 *>(depth + (s0 + k) * SEG);
#pragma unroll
        for (int q = lane; q < SEG / 4; q += 32) p[q] = z;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_reduce_units: one warp per unit = (interval, chunk of REDUCE_CHUNK bases).  cov/sum accumulated with one
// 64-bit atomic pair per unit.   if (depth >= minDep) { Cover++; Depth += depth; }   (:319-323, :340-344)
// ---------------------------------------------------------------------------------------------------------------
constexpr int REDUCE_CHUNK = 8192;
static_assert(REDUCE_CHUNK % SEG == 0 && REDUCE_CHUNK / SEG == 4, "a reduce chunk is four segments (k_reduce_units_part reads their states as one word)");

struct ReduceArgs {
    const uint32_t *depth;           // array base
    const uint64_t *contig_off;      // per tid, element offset (multiple of REDUCE_CHUNK and SCAN_TILE: pd_engine.cu PAD_ELEMS)
    const uint32_t *contig_len;      // per tid
    const int32_t *tid; const int32_t *start0; const int32_t *end0;
    const uint64_t *unit_prefix;     // n_iv+1: units before interval i; unit c of an interval is its part inside the
    uint64_t n_iv; uint64_t n_units; //   REDUCE_CHUNK-aligned chunk floor(start/REDUCE_CHUNK)+c of the contig
    uint32_t min_dep, mask;
    unsigned long long *cov; unsigned long long *sum;
    // partials written by k_tile_accumulate for (hint_min_dep, hint_mask); used when use_part (the caller asked for the same pair)
    uint32_t use_part; const uint8_t *seg_state; const uint32_t *part_cov; const unsigned long long *part_sum;
    const uint32_t *todo;            // non-null: only units flagged here (the others were served from the partials)
};

__device__ __forceinline__ void warp_reduce_range(const uint32_t *__restrict__ d, int64_t a, int64_t b, uint32_t min_dep,
                                                  uint32_t mask, int lane, uint32_t &cov, unsigned long long &sum)
{
    // per-lane partial sums over d[a..b), any alignment: head to 16-byte alignment, vector body, tail.  NOT yet warp-reduced.
    if (a >= b) return;
    const uint32_t *p = d + a;
    int64_t n = b - a;
    const int mis = (int)(((uintptr_t)p >> 2) & 3);
    int64_t head = mis ? (4 - mis) : 0;
    if (head > n) head = n;
    if (lane < head) { const uint32_t x = __ldg(p + lane) & mask; if (x >= min_dep) { cov++; sum += x; } }
    p += head; n -= head;
    const int64_t nv = n >> 2;
    const uint4 *p4 = reinterpret_cast<const uint4 *>(p);
    for (int64_t k = lane; k < nv; k += 32) {
        const uint4 v = __ldcs(p4 + k);
        uint32_t x;
        x = v.x & mask; if (x >= min_dep) { cov++; sum += x; }
        x = v.y & mask; if (x >= min_dep) { cov++; sum += x; }
        x = v.z & mask; if (x >= min_dep) { cov++; sum += x; }
        x = v.w & mask; if (x >= min_dep) { cov++; sum += x; }
    }
    const int64_t done = nv << 2;
    if (lane < n - done) { const uint32_t x = __ldg(p + done + lane) & mask; if (x >= min_dep) { cov++; sum += x; } }
}

// With the reduce hint most units need no depth at all: four written-once segments, four partials.  One LANE per unit then
// (a warp covers 32 consecutive units, which nearly always belong to one interval: one shuffle reduction, one atomic pair);
// units that do need depth (interval edges, re-written segments) are left to k_reduce_units, flagged in `todo`.
__global__ void __launch_bounds__(256) k_reduce_units_part(ReduceArgs a, uint32_t *todo)
{
    const int lane = threadIdx.x & 31;
    const uint64_t u = ((uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32 + lane;
    const bool live = u < a.n_units;
    uint64_t iv = 0; uint32_t cov = 0; unsigned long long sum = 0; bool slow = false;
    if (live) {
        uint64_t lo_i = 0, hi_i = a.n_iv;
        while (hi_i - lo_i > 1) {
            const uint64_t m = (lo_i + hi_i) >> 1;
            if (__ldg(a.unit_prefix + m) <= u) lo_i = m; else hi_i = m;
        }
        iv = lo_i;
        const int64_t c = (int64_t)(u - __ldg(a.unit_prefix + iv));
        const int32_t t = a.tid[iv];
        int64_t s = a.start0[iv], e = a.end0[iv];
        const int64_t len = a.contig_len[t];
        if (s < 0) s = 0;
        if (e > len) e = len;
        const int64_t chunk0 = (s / REDUCE_CHUNK + c) * REDUCE_CHUNK;
        const uint64_t sg0 = (a.contig_off[t] + (uint64_t)chunk0) / SEG;
        if (s <= chunk0 && e >= chunk0 + REDUCE_CHUNK) {
            // the whole chunk lies inside the interval: four segment partials, if all four were written exactly once
            const uint32_t st4 = *reinterpret_cast<const uint32_t *>(a.seg_state + sg0);       // REDUCE_CHUNK / SEG == 4 states, 4-byte aligned
            if (st4 == 0x01010101u) {
#pragma unroll
                for (int q = 0; q < REDUCE_CHUNK / SEG; q++) { cov += a.part_cov[sg0 + q]; sum += a.part_sum[sg0 + q]; }
            } else slow = true;
        } else slow = true;
    }
    if (slow) todo[u] = 1;
    // 32 consecutive units of one interval (the common case): reduce in the warp, one atomic pair
    const uint64_t iv0 = __shfl_sync(FULL, iv, 0);
    const bool uniform = __all_sync(FULL, !live || iv == iv0);
    if (uniform) {
        cov = __reduce_add_sync(FULL, cov);
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        if (lane == 0 && (cov | (sum != 0))) { atomicAdd(a.cov + iv0, (unsigned long long)cov); atomicAdd(a.sum + iv0, sum); }
    } else if (live && (cov | (sum != 0))) { atomicAdd(a.cov + iv, (unsigned long long)cov); atomicAdd(a.sum + iv, sum); }
}

__global__ void __launch_bounds__(256) k_reduce_units(ReduceArgs a)
{
    const uint64_t u = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (u >= a.n_units) return;
    if (a.todo && a.todo[u] == 0) return;                      // k_reduce_units_part took this unit from the partials
    const int lane = threadIdx.x & 31;
    // interval of this unit: last i with unit_prefix[i] <= u
    uint64_t lo_i = 0, hi_i = a.n_iv;
    while (hi_i - lo_i > 1) {
        const uint64_t m = (lo_i + hi_i) >> 1;
        if (__ldg(a.unit_prefix + m) <= u) lo_i = m; else hi_i = m;
    }
    const uint64_t iv = lo_i;
    const int64_t c = (int64_t)(u - __ldg(a.unit_prefix + iv));
    const int32_t t = a.tid[iv];
    int64_t s = a.start0[iv], e = a.end0[iv];
    const int64_t len = a.contig_len[t];
    if (s < 0) s = 0;
    if (e > len) e = len;
    const int64_t chunk0 = (s / REDUCE_CHUNK + c) * REDUCE_CHUNK;
    const int64_t lo = s > chunk0 ? s : chunk0;
    const int64_t hi = e < chunk0 + REDUCE_CHUNK ? e : chunk0 + REDUCE_CHUNK;
    const uint64_t off = a.contig_off[t];
    uint32_t cov = 0; unsigned long long sum = 0;
    if (a.use_part) {
        // segment by segment: a whole, written-once segment contributes its partial; anything else reads depth
        for (int64_t q = chunk0; q < chunk0 + REDUCE_CHUNK && q < hi; q += SEG) {
            const int64_t qa = lo > q ? lo : q, qb = hi < q + SEG ? hi : q + SEG;
            if (qa >= qb) continue;
            const uint64_t sg = (off + (uint64_t)q) / SEG;
            if (qb - qa == SEG && a.seg_state[sg] == 1) {
                if (lane == 0) { cov += a.part_cov[sg]; sum += a.part_sum[sg]; }
            } else warp_reduce_range(a.depth + off, qa, qb, a.min_dep, a.mask, lane, cov, sum);
        }
    } else warp_reduce_range(a.depth + off, lo, hi, a.min_dep, a.mask, lane, cov, sum);
#pragma unroll
    for (int o = 16; o; o >>= 1) { cov += __shfl_xor_sync(FULL, cov, o); sum += __shfl_xor_sync(FULL, sum, o); }
    if (lane == 0 && (cov | (sum != 0))) {
        atomicAdd(a.cov + iv, (unsigned long long)cov);
        atomicAdd(a.sum + iv, sum);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_reduce_windows: window k = [start0 + k*win, min(.., len)) of one contig (mode-6 loop :4366-4381).
// Wide windows: one warp per window.  Narrow windows (win <= 64): a warp loads 32*... handled by sub-warp groups.
// ---------------------------------------------------------------------------------------------------------------
struct WindowArgs {
    const uint32_t *depth;   // contig base
    int64_t len; int64_t start0; int64_t win; int64_t n_win;
    uint32_t min_dep, mask;
    uint32_t *cov; unsigned long long *sum;
};

__global__ void __launch_bounds__(256) k_reduce_windows(WindowArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    if (a.win >= 32 && a.win <= 160 && a.min_dep >= 1 && (unsigned long long)a.win * a.mask < (1ull << 32)) {
        // mode-6 sized windows (-w < 150): a warp takes 4 consecutive windows per iteration and issues all of their loads
        // (<= 5 per lane per window) before reducing any, then one redux.sync each for CoveredSite and TotalDepth
        // (TotalDepth of such a window fits 32 bits: win * 18-bit depth)
        // All per-window math is 32-bit and relative to the group's first base; only the group base pointer is 64-bit.
        constexpr int WB = 4, NL = 5;
        const int64_t ngroups = (a.n_win + WB - 1) / WB;
        const uint32_t win = (uint32_t)a.win, md = a.min_dep, mk = a.mask;
        // window starts are multiples of 4 elements from a 16-byte aligned base (contig offsets are multiples of 4096)
        const bool vec_ok = md == 1u && (win & 3u) == 0u && win <= 128u && (a.start0 & 3) == 0 && ((uintptr_t)a.depth & 15u) == 0;
        for (int64_t g = warp; g < ngroups; g += nwarps) {
            const int64_t k0 = g * WB;
            const int64_t lo0 = a.start0 + k0 * a.win;
            const uint32_t *__restrict__ d0 = a.depth + lo0 + lane;
            const int64_t rem64 = a.len - lo0;                                   // bases left in the contig from this group on
            const uint32_t rem = rem64 <= 0 ? 0u : (rem64 > (int64_t)(WB * 160) ? (uint32_t)(WB * 160) : (uint32_t)rem64);
            const uint32_t nw = (uint32_t)(a.n_win - k0 < WB ? a.n_win - k0 : WB);   // windows in this group
            if (vec_ok && nw == WB && rem >= WB * win) {
                // four full windows whose first bases are 16-byte aligned: one 128-bit load per lane per window
                // (lanes 0 .. win/4-1), -d 1: covered == non-zero and zeros add nothing
                const uint4 *__restrict__ d4 = reinterpret_cast<const uint4 *>(a.depth + lo0);
                const uint32_t q = win >> 2;                                      // uint4 per window (<= 32)
                uint4 v[WB];
#pragma unroll
                for (int j = 0; j < WB; ++j) v[j] = ((uint32_t)lane < q) ? __ldcs(d4 + (uint32_t)j * q + lane) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int j = 0; j < WB; ++j) {
                    const uint4 m = make_uint4(v[j].x & mk, v[j].y & mk, v[j].z & mk, v[j].w & mk);
                    uint32_t cov = nonzero4(m), sum = (m.x + m.y) + (m.z + m.w);
                    cov = __reduce_add_sync(FULL, cov);
                    sum = __reduce_add_sync(FULL, sum);
                    if (lane == 0) { a.cov[k0 + j] = cov; a.sum[k0 + j] = sum; }
                }
                continue;
            }
            uint32_t x[WB][NL];
#pragma unroll
            for (int j = 0; j < WB; ++j) {
                const uint32_t off = (uint32_t)j * win;                          // window j starts `off` bases into the group
                uint32_t wl = rem > off ? rem - off : 0u;                        // its length: min(win, bases left), 0 past n_win
                if (wl > win) wl = win;
                if ((uint32_t)j >= nw) wl = 0;
#pragma unroll
                for (int i = 0; i < NL; ++i) {
                    const uint32_t t = (uint32_t)lane + 32u * i;
                    x[j][i] = (t < wl) ? __ldcs(d0 + off + 32u * i) : 0u;
                }
            }
#pragma unroll
            for (int j = 0; j < WB; ++j) {
                uint32_t cov = 0, sum = 0;
#pragma unroll
                for (int i = 0; i < NL; ++i) {
                    const uint32_t v = x[j][i] & mk;              // lanes beyond the window loaded 0: never >= min_dep (>= 1)
                    const bool c = v >= md;
                    cov += c ? 1u : 0u; sum += c ? v : 0u;
                }
                cov = __reduce_add_sync(FULL, cov);
                sum = __reduce_add_sync(FULL, sum);
                if (lane == 0 && (uint32_t)j < nw) { a.cov[k0 + j] = cov; a.sum[k0 + j] = sum; }
            }
        }
    } else if (a.win >= 32 && a.win <= 512 && (unsigned long long)a.win * a.mask < (1ull << 32)) {
        // windows whose TotalDepth fits 32 bits: lane-strided loads, one redux.sync each for CoveredSite and TotalDepth
        for (int64_t k = warp; k < a.n_win; k += nwarps) {
            const int64_t lo = a.start0 + k * a.win;
            int64_t hi = lo + a.win; if (hi > a.len) hi = a.len;
            uint32_t cov = 0, sum = 0;
            for (int64_t q = lo + lane; q < hi; q += 32) {
                const uint32_t x = __ldcs(a.depth + q) & a.mask;
                if (x >= a.min_dep) { cov++; sum += x; }
            }
            cov = __reduce_add_sync(FULL, cov);
            sum = __reduce_add_sync(FULL, sum);
            if (lane == 0) { a.cov[k] = cov; a.sum[k] = sum; }
        }
    } else if (a.win >= 32) {
        for (int64_t k = warp; k < a.n_win; k += nwarps) {
            const int64_t lo = a.start0 + k * a.win;
            int64_t hi = lo + a.win; if (hi > a.len) hi = a.len;
            uint32_t cov = 0; unsigned long long sum = 0;
            warp_reduce_range(a.depth, lo, hi, a.min_dep, a.mask, lane, cov, sum);
#pragma unroll
            for (int o = 16; o; o >>= 1) { cov += __shfl_xor_sync(FULL, cov, o); sum += __shfl_xor_sync(FULL, sum, o); }
            if (lane == 0) { a.cov[k] = cov; a.sum[k] = sum; }
        }
    } else {
        // narrow windows: each lane owns one window
        for (int64_t k = warp * 32 + lane; k < a.n_win; k += nwarps * 32) {
            const int64_t lo = a.start0 + k * a.win;
            int64_t hi = lo + a.win; if (hi > a.len) hi = a.len;
            uint32_t cov = 0; unsigned long long sum = 0;
            for (int64_t q = lo; q < hi; ++q) { const uint32_t x = __ldg(a.depth + q) & a.mask; if (x >= a.min_dep) { cov++; sum += x; } }
            a.cov[k] = cov; a.sum[k] = sum;
        }
    }
}

}  // namespace pdk
