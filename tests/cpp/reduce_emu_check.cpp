// CPU check of the finalize / reduce kernels (pandepth_b200/csrc/pd_kernels.cuh) through the warp emulator (warp_emu.h):
//   reduce_emu_check scatter  <contig_len> <reads> <seed> <shift> <style 0|1|2>   k_delta_scatter (whole / striped / device-side count) + k_scan_add
//   reduce_emu_check reduce   <contig_len> <intervals> <seed> <min_dep> <mask>    k_fill_untouched + k_reduce_units_part + k_reduce_units (with and without partials)
//   reduce_emu_check windows  <contig_len> <win> <start0> <min_dep> <mask> <seed> k_reduce_windows (every width class of the kernel)
//   reduce_emu_check sweep    <contig_len> <first win> <last win> <mask> <seed>   the same for every width of a range (start, -d vary with the width)
// Each is compared with the plain loops of the reference: PanDepth.cpp:732-760 (walk), :295-348 (StatChrDepthWin / LowMEM), :4366-4381
// (mode-6 windows).  Two contigs share the arrays at 8192-aligned offsets, like the engine's layout.
#define PD_WARP_EMU 1
#include "warp_emu.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include "pd_kernels.cuh"
using namespace pdk;

template <typename F> static void run(unsigned grid, unsigned block, F f) { emu::launch(grid, block, [&](int) { f(); }, 0); }
// a kernel without collectives: every thread in turn on this thread
template <typename F> static void run_serial(unsigned grid, unsigned block, F f)
{
    for (unsigned b = 0; b < grid; b++) for (unsigned t = 0; t < block; t++) { threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid; f(); }
}
static uint32_t *aligned_u32(size_t n, uint32_t fill)
{
    uint32_t *p = (uint32_t *)std::aligned_alloc(64, (n * 4 + 63) / 64 * 64);
    for (size_t i = 0; i < n; i++) p[i] = fill;
    return p;
}
// contig offsets are multiples of max(SCAN_TILE, REDUCE_CHUNK) elements (pd_engine.cu: PAD_ELEMS)
static uint64_t pad_tile(uint64_t n) { const uint64_t P = SCAN_TILE > REDUCE_CHUNK ? SCAN_TILE : REDUCE_CHUNK; return (n + P - 1) / P * P; }

static int check_scatter(int32_t clen, uint32_t n, uint64_t seed, int32_t shift, int style)
{
    std::mt19937_64 rng(seed);
    const int32_t lens[2] = {clen, clen / 3 + 17};
    const uint64_t offs[2] = {0, pad_tile((uint64_t)clen + 1)};
    const uint64_t total = offs[1] + pad_tile((uint64_t)lens[1] + 1), n_seg = total / SEG;
    uint32_t *depth = aligned_u32(total, 0xcccccccc), *delta = aligned_u32(total, 0);
    std::vector<uint32_t> before(total, 0);
    std::vector<uint8_t> seg_state(n_seg, 0);
    for (uint64_t s = 0; s < n_seg; s++) if (rng() % 3 == 0) {                       // a third of the segments hold depth from sorted batches
        seg_state[s] = 2;
        for (uint64_t b = s * SEG; b < (s + 1) * SEG; b++) depth[b] = before[b] = (uint32_t)(rng() % 50);
    }
    std::vector<uint32_t> want(total, 0);
    int bad = 0, delta_used = 0;
    auto W = [](uint32_t len, uint32_t op) { return len << 4 | op; };
    for (int t = 0; t < 2; t++) {
        const int32_t L = lens[t], sh = t == 0 ? shift : 0;
        const uint32_t nr = t == 0 ? n : n / 4 + 1;
        std::vector<int32_t> pos(nr); std::vector<uint16_t> flag(nr); std::vector<uint8_t> mapq(nr); std::vector<uint32_t> off(nr + 1, 0), cig;
        for (uint32_t i = 0; i < nr; i++) {
            pos[i] = (int32_t)(rng() % (uint64_t)(L + 600)) - 300 + sh;
            if (rng() % 40 == 0) pos[i] = -1;
            flag[i] = (uint16_t)((rng() % 8 == 0) ? 1024 : (rng() % 2 ? 16 : 0)); mapq[i] = (uint8_t)(rng() % 61);
            switch (rng() % 8) {
            case 0: case 1: case 2: cig.push_back(W(150, 0)); break;
            case 3: cig.push_back(W(1 + rng() % 40, 4)); cig.push_back(W(1 + rng() % 120, 0)); break;
            case 4: cig.push_back(W(1 + rng() % 80, 0)); cig.push_back(W(1 + rng() % 9000, 3)); cig.push_back(W(1 + rng() % 80, 0)); break;
            case 5: cig.push_back(W(1 + rng() % 60, 7)); cig.push_back(W(1 + rng() % 9, 1)); cig.push_back(W(1 + rng() % 60, 8)); cig.push_back(W(1 + rng() % 30, 2)); cig.push_back(W(1 + rng() % 60, 0)); break;
            case 6: break;
            default: cig.push_back(W(1 + rng() % 20000, 0)); break;
            }
            off[i + 1] = (uint32_t)cig.size();
        }
        if (cig.empty()) cig.push_back(0);
        const uint32_t xflags = 1796; const int32_t min_mapq = 10;
        uint32_t n_apply = nr;                                                       // style 2: a device-side count, only reads [0, n_dev)
        const uint32_t n_dev = nr / 3;
        if (style == 2) n_apply = n_dev;
        for (uint32_t i = 0; i < n_apply; i++) {
            if ((flag[i] & xflags) || mapq[i] < min_mapq || pos[i] < 0) continue;
            int64_t q = (int64_t)pos[i] - sh;
            for (uint32_t k = off[i]; k < off[i + 1]; k++) {
                const uint32_t op = cig[k] & 15u; const int64_t len = cig[k] >> 4;
                if (is_match_op(op)) { for (int64_t b = std::max<int64_t>(q, 0); b < std::min<int64_t>(q + len, L); b++) want[offs[t] + (uint64_t)b]++; q += len; }
                else if (is_skip_op(op)) q += len;
            }
        }
        ScatterArgs a{pos.data(), cig.data(), off.data(), flag.data(), mapq.data(), nr, L, delta + offs[t], xflags, min_mapq, sh, nullptr, nullptr, 0, INT64_MAX};
        if (style == 0) run_serial((nr + 255) / 256, 256, [&] { k_delta_scatter(a); });
        else if (style == 1) {                                                       // one launch per stripe of the delta array, as pd_engine.cu does
            const int64_t stripe = 10000;
            for (int64_t lo = 0; lo <= L; lo += stripe) { a.stripe_lo = lo; a.stripe_hi = lo + stripe; run_serial((nr + 255) / 256, 256, [&] { k_delta_scatter(a); }); }
        } else { a.n_dev = &n_dev; a.delta_used = &delta_used; run_serial(3, 256, [&] { k_delta_scatter(a); }); }
    }
    bool any = false; for (uint64_t b = 0; b < total && !any; b++) any = want[b] != 0;
    if (style == 2 && any && !delta_used) { printf("the device-count scatter did not raise delta_used\n"); bad++; }
    if (style == 2 && !any && delta_used) { printf("delta_used raised although nothing was scattered\n"); bad++; }
    const uint64_t n_tiles = total / SCAN_TILE;
    std::vector<unsigned long long> desc(n_tiles, 0); unsigned int counter = 0;
    run((unsigned)n_tiles, SCAN_THREADS, [&] { k_scan_add(depth, delta, total, desc.data(), &counter, seg_state.data()); });
    size_t nbad = 0, first = 0;
    for (int t = 0; t < 2; t++) for (uint64_t b = offs[t]; b < offs[t] + (uint64_t)lens[t]; b++) if (depth[b] != before[b] + want[b]) { if (!nbad) first = b; nbad++; }
    // the padding behind a contig sees only zeros: every contig's deltas cancel before its padding ends
    for (uint64_t b = offs[0] + (uint64_t)lens[0] + 1; b < offs[1]; b++) if (depth[b] != before[b]) { if (!nbad) first = b; nbad++; }
    if (nbad) { printf("%zu elements differ, first at %zu: got %u want %u\n", nbad, first, depth[first], before[first] + want[first]); bad++; }
    printf("reduce_emu_check scatter: contigs %d + %d, %u reads, shift %d, style %d, %llu scan tiles: %s\n", lens[0], lens[1], n, shift, style,
           (unsigned long long)n_tiles, bad ? "MISMATCH" : "ok");
    std::free(depth); std::free(delta);
    return bad;
}

static int check_reduce(int32_t clen, uint32_t n_iv, uint64_t seed, uint32_t min_dep, uint32_t mask)
{
    std::mt19937_64 rng(seed);
    const uint32_t lens[2] = {(uint32_t)clen, (uint32_t)clen / 2 + 4097};
    const uint64_t offs[2] = {0, pad_tile((uint64_t)clen + 1)};
    const uint64_t total = offs[1] + pad_tile((uint64_t)lens[1] + 1), n_seg = total / SEG;
    int bad = 0;
    for (int use_part = 0; use_part < 2; use_part++) {
        uint32_t *depth = aligned_u32(total, 0);
        std::vector<uint8_t> seg_state(n_seg, 2);
        std::vector<uint32_t> part_cov(n_seg, 0xdeadbeefu); std::vector<unsigned long long> part_sum(n_seg, 0xdeadbeefdeadbeefull);
        for (uint64_t b = 0; b < total; b++) { const uint64_t r = rng(); depth[b] = r % 5 == 0 ? 0u : r % 7 == 0 ? (uint32_t)(r >> 20) : (uint32_t)((r >> 8) % 90); }
        std::vector<uint32_t> truth(depth, depth + total);
        for (uint64_t s = 0; s < n_seg; s++) {
            const uint64_t r = rng() % 10;
            const int t = s * SEG >= offs[1] ? 1 : 0;
            const bool inside = (s + 1) * SEG - offs[t] <= lens[t];
            if (r < 2) {                                                             // no batch touched it: garbage until k_fill_untouched
                seg_state[s] = 0;
                for (uint64_t b = s * SEG; b < (s + 1) * SEG; b++) { depth[b] = 0xcccccccc; truth[b] = 0; }
            } else if (use_part && r < 8 && inside) {                                // written once by the tile kernel: valid partials
                uint32_t c = 0; unsigned long long sm = 0;
                for (uint64_t b = s * SEG; b < (s + 1) * SEG; b++) { const uint32_t x = depth[b] & mask; if (x >= min_dep) { c++; sm += x; } }
                seg_state[s] = 1; part_cov[s] = c; part_sum[s] = sm;
            }
        }
        run((unsigned)((n_seg + 255) / 256), 256, [&] { k_fill_untouched(depth, seg_state.data(), n_seg, (uint32_t)use_part, part_cov.data(), part_sum.data()); });
        for (uint64_t b = 0; b < total && !bad; b++) if (depth[b] != truth[b]) { printf("k_fill_untouched: element %llu is %u, expected %u\n", (unsigned long long)b, depth[b], truth[b]); bad++; }
        for (uint64_t s = 0; s < n_seg && !bad; s++) if (seg_state[s] == 0) { printf("segment %llu left untouched\n", (unsigned long long)s); bad++; }
        // intervals: whole contigs, chunk-aligned, single bases, over both ends, empty
        std::vector<int32_t> tid(n_iv), s0(n_iv), e0(n_iv);
        for (uint32_t i = 0; i < n_iv; i++) {
            const int t = (int)(rng() % 2); const int64_t L = lens[t];
            int64_t s, e;
            switch (rng() % 8) {
            case 0: s = 0; e = L; break;
            case 1: s = (int64_t)(rng() % (L / REDUCE_CHUNK + 1)) * REDUCE_CHUNK; e = s + (int64_t)(1 + rng() % 5) * REDUCE_CHUNK; break;
            case 2: s = (int64_t)(rng() % L); e = s + 1; break;
            case 3: s = -(int64_t)(rng() % 500); e = (int64_t)(rng() % (L + 500)); break;
            case 4: s = (int64_t)(rng() % L); e = L + (int64_t)(rng() % 70000); break;
            case 5: s = (int64_t)(rng() % L); e = s - (int64_t)(rng() % 3); break;                       // empty
            default: s = (int64_t)(rng() % L); e = s + (int64_t)(rng() % 40000); break;
            }
            tid[i] = t; s0[i] = (int32_t)s; e0[i] = (int32_t)e;
        }
        std::vector<uint64_t> prefix(n_iv + 1); uint64_t n_units = 0;                // pd_reduce's unit table
        for (uint32_t i = 0; i < n_iv; i++) {
            prefix[i] = n_units;
            const int64_t s = std::max<int64_t>(s0[i], 0), e = std::min<int64_t>(e0[i], lens[tid[i]]);
            if (s < e) n_units += (uint64_t)((e - 1) / REDUCE_CHUNK - s / REDUCE_CHUNK + 1);
        }
        prefix[n_iv] = n_units;
        std::vector<unsigned long long> cov(n_iv, 0), sum(n_iv, 0);
        std::vector<uint32_t> todo(n_units + 1, 0);
        ReduceArgs r{depth, offs, lens, tid.data(), s0.data(), e0.data(), prefix.data(), n_iv, n_units, min_dep, mask, cov.data(), sum.data(),
                     (uint32_t)use_part, seg_state.data(), part_cov.data(), part_sum.data(), use_part ? todo.data() : nullptr};
        if (n_units) {
            if (use_part) run((unsigned)((n_units + 255) / 256), 256, [&] { k_reduce_units_part(r, todo.data()); });
            run((unsigned)((n_units + 7) / 8), 256, [&] { k_reduce_units(r); });
        }
        uint64_t served = 0; for (uint64_t u = 0; u < n_units; u++) served += todo[u] == 0;
        for (uint32_t i = 0; i < n_iv; i++) {
            unsigned long long c = 0, sm = 0;                                        // :319-323 / :340-344
            const int64_t s = std::max<int64_t>(s0[i], 0), e = std::min<int64_t>(e0[i], lens[tid[i]]);
            for (int64_t b = s; b < e; b++) { const uint32_t x = truth[offs[tid[i]] + (uint64_t)b] & mask; if (x >= min_dep) { c++; sm += x; } }
            if (cov[i] != c || sum[i] != sm) { printf("interval %u (%d:%d-%d, partials %d): got %llu/%llu, expected %llu/%llu\n", i, tid[i], s0[i], e0[i], use_part, cov[i], sum[i], c, sm); bad++; break; }
        }
        printf("reduce_emu_check reduce: contigs %u + %u, %u intervals, %llu units%s, min_dep %u, mask %x: %s\n", lens[0], lens[1], n_iv, (unsigned long long)n_units,
               use_part ? (", " + std::to_string(served) + " served from partials alone").c_str() : "", min_dep, mask, bad ? "MISMATCH" : "ok");
        std::free(depth);
    }
    return bad;
}

static int check_windows(int64_t len, int64_t win, int64_t start0, uint32_t min_dep, uint32_t mask, uint64_t seed)
{
    std::mt19937_64 rng(seed);
    uint32_t *depth = aligned_u32((size_t)pad_tile((uint64_t)len + 1), 0);
    for (int64_t b = 0; b < len; b++) { const uint64_t r = rng(); depth[b] = r % 4 == 0 ? 0u : r % 9 == 0 ? (uint32_t)(r >> 24) : (uint32_t)((r >> 8) % 200); }
    for (int64_t b = len; b < (int64_t)pad_tile((uint64_t)len + 1); b++) depth[b] = 0x7fffffff;       // nothing behind the contig end may be counted
    const int64_t n_win = start0 < len ? (len - start0 + win - 1) / win : 0;
    std::vector<uint32_t> cov((size_t)n_win + 1, 0xdeadbeefu); std::vector<unsigned long long> sum((size_t)n_win + 1, 0xdeadbeefull);
    WindowArgs w{depth, len, start0, win, n_win, min_dep, mask, cov.data(), sum.data()};
    const int64_t warps_needed = (win >= 32 && win <= 160 && min_dep >= 1) ? (n_win + 3) / 4 : (win >= 32 ? n_win : (n_win + 31) / 32);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((warps_needed + 7) / 8, 3));   // a small grid: every warp loops
    if (n_win) run(grid, 256, [&] { k_reduce_windows(w); });
    int bad = 0;
    for (int64_t k = 0; k < n_win; k++) {                                            // :4366-4381
        unsigned long long c = 0, sm = 0;
        for (int64_t b = start0 + k * win; b < std::min(start0 + (k + 1) * win, len); b++) { const uint32_t x = depth[b] & mask; if (x >= min_dep) { c++; sm += x; } }
        if (cov[(size_t)k] != c || sum[(size_t)k] != sm) { printf("window %lld: got %u/%llu, expected %llu/%llu\n", (long long)k, cov[(size_t)k], sum[(size_t)k], c, sm); bad++; break; }
    }
    if (cov[(size_t)n_win] != 0xdeadbeefu) { printf("a window behind the last one was written\n"); bad++; }
    printf("reduce_emu_check windows: len %lld, win %lld, start %lld, min_dep %u, mask %x, %lld windows: %s\n", (long long)len, (long long)win, (long long)start0, min_dep, mask,
           (long long)n_win, bad ? "MISMATCH" : "ok");
    std::free(depth);
    return bad;
}

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: see the header of reduce_emu_check.cpp\n"); return 2; }
    const std::string mode = argv[1];
    if (mode == "scatter") return check_scatter(atoi(argv[2]), (uint32_t)atoi(argv[3]), (uint64_t)atoll(argv[4]), atoi(argv[5]), atoi(argv[6])) ? 1 : 0;
    if (mode == "reduce") return check_reduce(atoi(argv[2]), (uint32_t)atoi(argv[3]), (uint64_t)atoll(argv[4]), (uint32_t)strtoul(argv[5], nullptr, 0), (uint32_t)strtoul(argv[6], nullptr, 0)) ? 1 : 0;
    if (mode == "sweep") {                                   // every width in [lo, hi]: sweep <len> <lo> <hi> <mask> <seed>
        int bad = 0;
        for (int64_t w = atoll(argv[3]); w <= atoll(argv[4]); w++)
            bad += check_windows(atoll(argv[2]) + (w % 3), w, (w % 8 == 0) ? 0 : (w % 8 == 4) ? 64 : w % 41, (w % 4 == 2) ? 3u : 1u, (uint32_t)strtoul(argv[5], nullptr, 0), (uint64_t)atoll(argv[6]) + (uint64_t)w);
        return bad ? 1 : 0;
    }
    if (mode == "windows" && argc >= 8)
        return check_windows(atoll(argv[2]), atoll(argv[3]), atoll(argv[4]), (uint32_t)strtoul(argv[5], nullptr, 0), (uint32_t)strtoul(argv[6], nullptr, 0), (uint64_t)atoll(argv[7])) ? 1 : 0;
    return 2;
}
