// warp_emu.h -- TEST INFRASTRUCTURE: run warp-cooperative CUDA kernels on the CPU, one host thread per lane.
//
// Divergent code needs no special treatment: lanes only meet at the *_sync collectives, which are implemented with a
// per-warp barrier (every collective in the kernels under test uses the full mask).  Shared memory is a static object
// shared by all lane threads; __syncthreads is a per-CTA barrier.  Slow (microseconds per collective) but exact, which is
// what is needed to check the index arithmetic of pd_bgzf.cuh against zlib before a GPU is involved.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__

struct EmuDim { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuDim threadIdx, blockIdx, blockDim, gridDim;

namespace emu {

// sense-reversing barrier; waiters yield instead of sleeping (a futex round trip per collective made the emulation 10x slower)
struct Barrier {
    std::atomic<int> waiting{0}; std::atomic<uint64_t> gen{0}; int n;
    explicit Barrier(int k) : n(k) {}
    void wait()
    {
        const uint64_t my = gen.load(std::memory_order_acquire);
        if (waiting.fetch_add(1, std::memory_order_acq_rel) + 1 == n) { waiting.store(0, std::memory_order_relaxed); gen.store(my + 1, std::memory_order_release); }
        else { int spins = 0; while (gen.load(std::memory_order_acquire) == my) { if (++spins > 64) std::this_thread::yield(); } }
    }
};

struct Warp { Barrier bar{32}; uint64_t slot[32]; };
struct Cta { Barrier bar; std::vector<Warp> warps; std::vector<uint32_t> dyn; explicit Cta(int threads) : bar(threads), warps((size_t)(threads + 31) / 32), dyn(64 * 1024) {} };   // dyn: 256 KB of "dynamic shared memory"

static thread_local Warp *tl_warp = nullptr;
static thread_local Cta *tl_cta = nullptr;
static thread_local int tl_lane = 0;

// run kernel(args) over grid x block threads; CTAs one after another, the threads of a CTA concurrently
template <typename K, typename A>
void launch(unsigned grid, unsigned block, K kernel, A args)
{
    for (unsigned b = 0; b < grid; b++) {
        Cta cta((int)block);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t] {
                threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
                tl_cta = &cta; tl_warp = &cta.warps[t / 32]; tl_lane = (int)(t % 32);
                kernel(args);
                // a lane that returns early must not strand the others at a collective: kernels under test only return
                // warp-uniformly, so nothing to do here
            });
        for (auto &x : th) x.join();
    }
}

inline uint32_t *dyn_smem() { return tl_cta->dyn.data(); }

template <typename T> inline T exchange(T v, int src)
{
    Warp &w = *tl_warp;
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    w.slot[tl_lane] = raw;
    w.bar.wait();
    const uint64_t got = w.slot[src & 31];
    w.bar.wait();
    T r; memcpy(&r, &got, sizeof(T));
    return r;
}

}  // namespace emu

inline void __syncwarp(unsigned = 0xffffffffu) { emu::tl_warp->bar.wait(); }
inline void __syncthreads() { emu::tl_cta->bar.wait(); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::exchange(v, src); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    const int src = emu::tl_lane - (int)d;
    const T r = emu::exchange(v, src < 0 ? emu::tl_lane : src);
    return src < 0 ? v : r;
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::exchange(v, emu::tl_lane ^ m); }
inline unsigned __ballot_sync(unsigned, int pred)
{
    emu::Warp &w = *emu::tl_warp;
    w.slot[emu::tl_lane] = pred ? 1 : 0;
    w.bar.wait();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)w.slot[i] << i;
    w.bar.wait();
    return r;
}
inline unsigned __match_any_sync(unsigned, uint32_t v)
{
    emu::Warp &w = *emu::tl_warp;
    w.slot[emu::tl_lane] = v;
    w.bar.wait();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((uint32_t)w.slot[i] == v) r |= 1u << i;
    w.bar.wait();
    return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
    emu::Warp &w = *emu::tl_warp;
    w.slot[emu::tl_lane] = v;
    w.bar.wait();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)w.slot[i];
    w.bar.wait();
    return r;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
    emu::Warp &w = *emu::tl_warp;
    w.slot[emu::tl_lane] = v;
    w.bar.wait();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r += (unsigned)w.slot[i];
    w.bar.wait();
    return r;
}
// "shared window" addresses: the kernels under test do arithmetic on them (2 KB alignment, OR-ed offsets), so the emulation
// uses the distance from a 2 KB-aligned anchor near the static objects that stand in for shared memory: linear, and the low
// 11 bits are those of the host address
namespace emu {
static char anchor;
inline uintptr_t shared_base() { return (((uintptr_t)&anchor) - (uintptr_t)(512u << 20)) & ~(uintptr_t)2047; }
}
inline uint32_t emu_shared_addr(const void *p) { return (uint32_t)((uintptr_t)p - emu::shared_base()); }
inline const void *emu_shared_ptr(uint32_t sa) { return (const void *)(emu::shared_base() + sa); }
#define __align__(n) __attribute__((aligned(n)))
struct uint4 { uint32_t x, y, z, w; };
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline void __stcs(T *p, const T &v) { *p = v; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __brev(unsigned v)
{
    v = (v >> 16) | (v << 16); v = ((v & 0xff00ff00u) >> 8) | ((v & 0x00ff00ffu) << 8);
    v = ((v & 0xf0f0f0f0u) >> 4) | ((v & 0x0f0f0f0fu) << 4); v = ((v & 0xccccccccu) >> 2) | ((v & 0x33333333u) << 2);
    return ((v & 0xaaaaaaaau) >> 1) | ((v & 0x55555555u) << 1);
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
