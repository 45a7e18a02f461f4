// pd_kernels.cuh -- sm_100a kernels of the per-base depth engine.
//
// Roles (reference = /root/reference/src/PanDepth.cpp):
//   k_cut              index a sorted batch by genomic segment (no CPU counterpart; replaces the BAI window queries :726)
//   k_tile_accumulate  filter + CIGAR walk -> per-base depth, one warp per 1024-base segment   (:732-760)
//   k_delta_scatter    same walk for unsorted / long-span batches: +-1 deltas with global reductions (:732-760)
//   k_scan_add         prefix-sum of the delta array folded into depth (the CPU increments every base instead)
//   k_fill_untouched   zero-fill segments no batch touched (the `depth[ccv]=0` loop :718-721, only where needed)
//   k_reduce_units     interval CoveredSite/TotalDepth                                         (:295-348)
//   k_reduce_windows   fixed-stride windows (mode 6 loop)                                      (:4366-4381)
//
// All arithmetic is integer (uint32 modular depth, uint64 sums); no tensor cores are involved.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdk {

constexpr int SEG = 1024;   // bases owned by one warp of k_tile_accumulate (4 KB of depth)
constexpr int HALO = 256;   // look-back window in bases; also the largest "common span" L the fast path models
constexpr int WPC = 8;      // warps per CTA in k_tile_accumulate
constexpr int TILE_SMEM_WORDS = 2 * SEG + HALO;   // per warp: H[SEG+HALO] start counts, D[SEG] corrections
constexpr unsigned FULL = 0xffffffffu;

enum { ERR_UNSORTED = 1, ERR_RANGE = 2 };

__device__ __forceinline__ bool is_match_op(uint32_t op) { return op == 0u || op == 7u || op == 8u; }   // M = X  (:745-747)
__device__ __forceinline__ bool is_skip_op(uint32_t op) { return op == 2u || op == 3u; }                  // D N    (:754-755)

__device__ __forceinline__ int32_t ldg_i32(const int32_t *p) { return __ldg(p); }
__device__ __forceinline__ uint32_t ldg_u32(const uint32_t *p) { return __ldg(p); }

__device__ __forceinline__ void red_add_u32(uint32_t *addr, uint32_t v)
{
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// k_cut: for segment j (bases [s0, s0+SEG)) of a pos-sorted batch: lo[j] = first read with pos >= s0-HALO,
// hi[j] = first read with pos >= s0+SEG.  Warp 0 of block 0 also samples 32 reads and publishes the most common
// single-run span L (0 if none dominates or L > HALO).
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
    const int32_t *pos; const uint32_t *cigar; const uint32_t *cigar_off;
    const uint16_t *flag; const uint8_t *mapq;
    uint32_t n; int32_t seg0; int32_t nseg;
    uint32_t *lo; uint32_t *hi; uint32_t *L;
    uint32_t xflags; int32_t min_mapq;
};

__global__ void __launch_bounds__(256) k_cut(CutArgs a)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < a.nseg) {
        const int64_t s0 = (int64_t)(a.seg0 + j) * SEG;
        const uint32_t l = lower_bound_pos(a.pos, a.n, s0 - HALO);
        // hi >= lo always; search the tail only
        const uint32_t h = l + lower_bound_pos(a.pos + l, a.n - l, s0 + SEG);
        a.lo[j] = l;
        a.hi[j] = h;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        uint32_t span = 0;
        if (a.n > 0) {
            const uint32_t i = (uint32_t)(((uint64_t)a.n * (uint32_t)lane) / 32u);
            uint32_t c, nops;
            if (a.cigar_off) { uint32_t c0 = a.cigar_off[i]; nops = a.cigar_off[i + 1] - c0; c = nops ? a.cigar[c0] : 0u; }
            else { nops = 1; c = a.cigar[i]; }
            bool pass = true;
            if (a.flag && (a.flag[i] & a.xflags)) pass = false;
            if (a.mapq && (int)a.mapq[i] < a.min_mapq) pass = false;
            if (pass && nops == 1 && is_match_op(c & 15u)) span = c >> 4;
        }
        const unsigned same = __match_any_sync(FULL, span);
        int votes = __popc(same);
        if (span == 0 || span > (uint32_t)HALO) votes = 0;
        int best = votes;
        for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
        const unsigned who = __ballot_sync(FULL, votes == best && best >= 8);
        uint32_t L = 0;
        if (who) L = __shfl_sync(FULL, span, __ffs(who) - 1);
        if (lane == 0) *a.L = L;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_tile_accumulate: one warp per SEG-base segment of one contig.
//
// The warp sees every read whose pos lies in [s0-HALO, s0+SEG) (lo/hi from k_cut).  Per-base depth of the segment is
//     depth(b) = #{reads: pos in (b-L, b]}            "every read is one run of the common span L"   (table H)
//              + prefix-sum of corrections D(b)        for reads that are not (filtered / other CIGARs) (table D)
// H is filled with plain (non-atomic) read-modify-writes: reads are sorted, so within a 32-read chunk equal positions
// are adjacent and one lane per distinct position adds the group size; chunks are processed sequentially by the same
// warp and H is private to it.  D takes shared-memory atomics, but only for the uncommon reads.  With L = 0 (no
// common span) every read goes through D: the plain "scatter +-1 deltas into a shared tile" scheme.
// Runs that reach past the segments that see the read (pos + HALO) are "far": the owner segment scatters them as
// +-1 deltas into the global delta array, folded in by k_scan_add at finalize.
// The segment is then scanned once (warp shuffles) and written with coalesced 16-byte stores: a plain streaming
// store if no earlier batch touched the segment (zero-initialisation is fused away), else load+add+store.
// ---------------------------------------------------------------------------------------------------------------
struct TileArgs {
    const int32_t *pos; const uint32_t *cigar; const uint32_t *cigar_off;
    const uint16_t *flag; const uint8_t *mapq;
    const uint32_t *lo; const uint32_t *hi; const uint32_t *L;
    int32_t seg0; int32_t nseg; int32_t contig_len;
    uint32_t *depth;      // contig base
    uint32_t *delta;      // contig base (global delta array)
    uint8_t *seg_state;   // contig base: 0 = never written
    uint32_t xflags; int32_t min_mapq;
    int *err; int *delta_used;
};

__global__ void __launch_bounds__(WPC * 32) k_tile_accumulate(TileArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * WPC + warp;
    if (j >= a.nseg) return;
    const uint32_t i_lo = a.lo[j], i_hi = a.hi[j];
    if (i_lo >= i_hi) return;

    uint32_t *H = smem + warp * TILE_SMEM_WORDS;
    uint32_t *D = H + SEG + HALO;
    const int64_t s0 = (int64_t)(a.seg0 + j) * SEG;
    const int64_t hbase = s0 - HALO;
    const int64_t send = s0 + SEG;
    const int L = (int)*a.L;
    const int64_t clen = a.contig_len;

    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *H4 = reinterpret_cast<uint4 *>(H);
#pragma unroll
        for (int k = lane; k < TILE_SMEM_WORDS / 4; k += 32) H4[k] = z;
    }
    __syncwarp();

    int32_t carry_p = INT32_MIN;   // last position of the previous chunk (sortedness check)
    for (uint32_t base = i_lo; base < i_hi; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < i_hi;
        const int32_t p = valid ? ldg_i32(a.pos + i) : INT32_MAX;
        uint32_t c = 0, c0 = i, nops = 1;
        bool pass = valid;
        if (valid) {
            if (a.cigar_off) { c0 = ldg_u32(a.cigar_off + i); nops = ldg_u32(a.cigar_off + i + 1) - c0; c = nops ? ldg_u32(a.cigar + c0) : 0u; }
            else c = ldg_u32(a.cigar + i);
            if (a.flag && (a.flag[i] & a.xflags)) pass = false;          // :732
            if (a.mapq && (int)a.mapq[i] < a.min_mapq) pass = false;     // :733
            if (p < 0) pass = false;
        }
        int32_t pp = __shfl_up_sync(FULL, p, 1);
        const int32_t pn = __shfl_down_sync(FULL, p, 1);
        if (lane == 0) pp = carry_p;
        if (valid && p < pp) atomicOr(a.err, ERR_UNSORTED);
        carry_p = __shfl_sync(FULL, p, 31);
        const bool first = (lane == 0) || (p != pp);
        const bool last = (lane == 31) || (p != pn);
        const unsigned mf = __ballot_sync(FULL, first);
        if (valid && last) {
            const int fl = 31 - __clz(mf & (FULL >> (31 - lane)));
            const uint32_t idx = (uint32_t)((int64_t)p - hbase);
            if (idx < (uint32_t)(SEG + HALO)) H[idx] += (uint32_t)(lane - fl + 1);
            else atomicOr(a.err, ERR_RANGE);
        }
        const bool common = pass && nops == 1 && L > 0 && (c >> 4) == (uint32_t)L && is_match_op(c & 15u);
        if (valid && !common) {
            // corrections: cancel the modelled run [p, p+L), then add the real runs
            auto emit = [&](int64_t x, uint32_t v) {
                if (x < send) atomicAdd(&D[x <= s0 ? 0 : (int)(x - s0)], v);
            };
            bool cancel_start = (L > 0);
            if (L > 0) emit((int64_t)p + L, 1u);
            if (pass) {
                const bool owner = (int64_t)p >= s0;
                const int64_t far_lo = (((int64_t)p + HALO) / SEG + 1) * SEG;
                int64_t q = p, rs = -1, re = -1;
                auto flush = [&]() {
                    if (rs < 0) return;
                    int64_t s = rs < 0 ? 0 : rs, e = re > clen ? clen : re;
                    if (s >= e) return;
                    if (s == (int64_t)p && cancel_start) cancel_start = false;   // +1 and the cancelling -1 meet at p
                    else emit(s, 1u);
                    emit(e, 0xffffffffu);
                    if (owner && e > far_lo) {
                        const int64_t gs = s > far_lo ? s : far_lo;
                        red_add_u32(a.delta + gs, 1u);
                        red_add_u32(a.delta + e, 0xffffffffu);
                        *a.delta_used = 1;
                    }
                };
                for (uint32_t k = 0; k < nops; ++k) {
                    const uint32_t w = (k == 0) ? c : ldg_u32(a.cigar + c0 + k);
                    const uint32_t op = w & 15u;
                    const int64_t n = w >> 4;
                    if (is_match_op(op)) {
                        if (rs >= 0 && re == q) re = q + n; else { flush(); rs = q; re = q + n; }
                        q += n;
                    } else if (is_skip_op(op)) q += n;
                }
                flush();
            }
            if (cancel_start) emit((int64_t)p, 0xffffffffu);
        }
        __syncwarp();
    }

    // depth just left of the segment under the common-span model: reads with pos in [s0-L, s0)
    uint32_t carry = 0;
    if (L > 0) {
        uint32_t part = 0;
        for (int k = HALO - L + lane; k < HALO; k += 32) part += H[k];
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
        carry = part;
    }
    uint8_t *st_ptr = a.seg_state + (a.seg0 + j);
    const bool fresh = (*st_ptr == 0);
    uint32_t *out = a.depth + s0;
#pragma unroll 2
    for (int r = 0; r < SEG / 128; ++r) {
        const int b = r * 128 + lane * 4;
        const uint4 h = *reinterpret_cast<const uint4 *>(H + HALO + b);
        const uint4 d = *reinterpret_cast<const uint4 *>(D + b);
        uint32_t x0 = h.x + d.x, x1 = h.y + d.y, x2 = h.z + d.z, x3 = h.w + d.w;
        if (L > 0) {
            const uint32_t *hl = H + HALO + b - L;
            x0 -= hl[0]; x1 -= hl[1]; x2 -= hl[2]; x3 -= hl[3];
        }
        x1 += x0; x2 += x1; x3 += x2;
        uint32_t t = x3;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, t, o);
            if (lane >= o) t += y;
        }
        const uint32_t ex = t - x3 + carry;
        carry += __shfl_sync(FULL, t, 31);
        uint4 v = make_uint4(x0 + ex, x1 + ex, x2 + ex, x3 + ex);
        uint4 *dst = reinterpret_cast<uint4 *>(out + b);
        if (fresh) __stcs(dst, v);
        else {
            const uint4 old = *dst;
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
            *dst = v;
        }
    }
    if (fresh && lane == 0) *st_ptr = 1;
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
};

__global__ void __launch_bounds__(256) k_delta_scatter(ScatterArgs a)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (a.flag && (a.flag[i] & a.xflags)) return;
    if (a.mapq && (int)a.mapq[i] < a.min_mapq) return;
    const int32_t p = a.pos[i];
    if (p < 0) return;
    uint32_t c0 = i, nops = 1;
    if (a.cigar_off) { c0 = a.cigar_off[i]; nops = a.cigar_off[i + 1] - c0; }
    const int64_t clen = a.contig_len;
    int64_t q = p, rs = -1, re = -1;
    auto flush = [&]() {
        if (rs < 0) return;
        const int64_t s = rs, e = re > clen ? clen : re;
        if (s >= e) return;
        red_add_u32(a.delta + s, 1u);
        red_add_u32(a.delta + e, 0xffffffffu);
    };
    for (uint32_t k = 0; k < nops; ++k) {
        const uint32_t w = a.cigar[c0 + k];
        const uint32_t op = w & 15u;
        const int64_t n = w >> 4;
        if (is_match_op(op)) {
            if (rs >= 0 && re == q) re = q + n; else { flush(); rs = q; re = q + n; }
            q += n;
        } else if (is_skip_op(op)) q += n;
    }
    flush();
}

// ---------------------------------------------------------------------------------------------------------------
// k_scan_add: depth[i] += inclusive_prefix_sum(delta)[i]; delta[i] = 0.  Single pass, decoupled look-back
// (tile aggregates published through 64-bit {status,value} descriptors), dynamic tile ids for forward progress.
// Every contig's deltas sum to zero before its padding ends, so one scan over the whole array needs no resets.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;                           // uint32 per thread (4 x uint4)
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;     // 4096 elements = one 16 KB tile

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t *__restrict__ depth, uint32_t *__restrict__ delta,
                                                           uint64_t n_elems, unsigned long long *desc, unsigned int *counter)
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

    // warp w owns 512 consecutive elements: 4 rows of 128 (lane holds 4)
    uint4 v[4];
    uint32_t row_tot[4];
    const uint64_t wbase = tbase + (uint64_t)warp * 512;
    uint32_t wcarry = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint64_t e = wbase + r * 128 + lane * 4;      // n_elems is a multiple of SCAN_TILE (engine pads)
        uint4 x = *reinterpret_cast<const uint4 *>(delta + e);
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
        row_tot[r] = __shfl_sync(FULL, t, 31);
        wcarry += row_tot[r];
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
                    do { d = atomicAdd(desc + idx, 0ull); status = (uint32_t)(d >> 32); } while (status == 0);
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
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint64_t e = wbase + r * 128 + lane * 4;
        uint4 *dp = reinterpret_cast<uint4 *>(depth + e);
        uint4 o = *dp;
        o.x += v[r].x + add; o.y += v[r].y + add; o.z += v[r].z + add; o.w += v[r].w + add;
        *dp = o;
        *reinterpret_cast<uint4 *>(delta + e) = z;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_fill_untouched: one warp per segment; segments never written by a batch become zeros.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_untouched(uint32_t *depth, uint8_t *seg_state, uint64_t n_seg)
{
    const uint64_t s = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= n_seg) return;
    if (seg_state[s] != 0) return;
    const int lane = threadIdx.x & 31;
    uint4 *p = reinterpret_cast<uint4 *>(depth + s * SEG);
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = lane; k < SEG / 4; k += 32) p[k] = z;
    __syncwarp();
    if (lane == 0) seg_state[s] = 1;
}

// ---------------------------------------------------------------------------------------------------------------
// k_reduce_units: one warp per unit = (interval, chunk of REDUCE_CHUNK bases).  cov/sum accumulated with one
// 64-bit atomic pair per unit.   if (depth >= minDep) { Cover++; Depth += depth; }   (:319-323, :340-344)
// ---------------------------------------------------------------------------------------------------------------
constexpr int REDUCE_CHUNK = 8192;

struct ReduceArgs {
    const uint32_t *depth;           // array base
    const uint64_t *contig_off;      // per tid, element offset
    const uint32_t *contig_len;      // per tid
    const int32_t *tid; const int32_t *start0; const int32_t *end0;
    const uint2 *units;              // (interval index, chunk index)
    uint64_t n_units;
    uint32_t min_dep, mask;
    unsigned long long *cov; unsigned long long *sum;
};

__device__ __forceinline__ void warp_reduce_range(const uint32_t *__restrict__ d, int64_t a, int64_t b, uint32_t min_dep,
                                                  uint32_t mask, int lane, uint32_t &cov, unsigned long long &sum)
{
    // d + a .. d + b, any alignment.  head to 16-byte alignment, vector body, tail.
    cov = 0; sum = 0;
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
    for (int o = 16; o; o >>= 1) {
        cov += __shfl_xor_sync(FULL, cov, o);
        sum += __shfl_xor_sync(FULL, sum, o);
    }
}

__global__ void __launch_bounds__(256) k_reduce_units(ReduceArgs a)
{
    const uint64_t u = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (u >= a.n_units) return;
    const int lane = threadIdx.x & 31;
    const uint2 un = a.units[u];
    const int32_t t = a.tid[un.x];
    int64_t s = a.start0[un.x], e = a.end0[un.x];
    const int64_t len = a.contig_len[t];
    if (s < 0) s = 0;
    if (e > len) e = len;
    int64_t lo = s + (int64_t)un.y * REDUCE_CHUNK, hi = lo + REDUCE_CHUNK;
    if (hi > e) hi = e;
    uint32_t cov; unsigned long long sum;
    warp_reduce_range(a.depth + a.contig_off[t], lo, hi, a.min_dep, a.mask, lane, cov, sum);
    if (lane == 0 && (cov | (sum != 0))) {
        atomicAdd(a.cov + un.x, (unsigned long long)cov);
        atomicAdd(a.sum + un.x, sum);
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
    if (a.win >= 32) {
        for (int64_t k = warp; k < a.n_win; k += nwarps) {
            const int64_t lo = a.start0 + k * a.win;
            int64_t hi = lo + a.win; if (hi > a.len) hi = a.len;
            uint32_t cov; unsigned long long sum;
            warp_reduce_range(a.depth, lo, hi, a.min_dep, a.mask, lane, cov, sum);
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
