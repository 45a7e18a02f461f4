// CPU check of the accumulate kernels (pandepth_b200/csrc/pd_kernels.cuh: k_cut_coarse, k_cut, k_filter_bits and all four
// instantiations of k_tile_accumulate) through the warp emulator (warp_emu.h):
//   tile_emu_check [contig_len] [reads] [seed] [has_off 0|1] [has_filter 0|1] [shift]
// A random coordinate-sorted batch (single-op or mixed CIGARs with far N skips, clips, I/D/=/X; flags, mapq; pile-ups; reads
// over the contig end) is cut and accumulated twice -- the second pass runs over segments that already hold depth (the bulk
// reduce-add branch) -- with the reduce hint on, the far-run deltas are folded in as k_scan_add would, and per-base depth and the
// per-segment (CoveredSite, TotalDepth) partials of the first pass are compared with a plain host walk (PanDepth.cpp:732-760).
#define PD_WARP_EMU 1
#include "warp_emu.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "pd_kernels.cuh"
using namespace pdk;

template <typename F> static void run(unsigned grid, unsigned block, F f) { emu::launch(grid, block, [&](int) { f(); }, 0); }

int main(int argc, char **argv)
{
    const int32_t clen = argc > 1 ? atoi(argv[1]) : 60000;
    const uint32_t n = argc > 2 ? (uint32_t)atoi(argv[2]) : 20000;
    const uint64_t seed = argc > 3 ? (uint64_t)atoll(argv[3]) : 1;
    const bool has_off = argc > 4 ? atoi(argv[4]) != 0 : true, has_filter = argc > 5 ? atoi(argv[5]) != 0 : true;
    const int32_t shift = argc > 6 ? atoi(argv[6]) : 0;                  // owned range = [shift, shift + clen) of the contig
    std::mt19937_64 rng(seed);
    std::vector<int32_t> pos(n);
    // a ranged engine (shift != 0) sees reads up to HALO left of its range in segment 0 (the ones further left go through
    // k_delta_scatter in the engine: tests/test_engine_gpu.py); reads over the end of the range / contig are clipped
    for (auto &p : pos) p = shift + (int32_t)(rng() % (uint64_t)(clen + 200 + (shift ? HALO : 0))) - (shift ? HALO : 0);
    for (uint32_t i = 0; i < n / 10; i++) pos[rng() % n] = shift + clen / 2;       // a pile-up
    for (auto &p : pos) if (p < 0) p = 0;                                          // (a range that starts less than HALO into the contig; pos -1 means unmapped)
    std::sort(pos.begin(), pos.end());
    std::vector<uint16_t> flag(n); std::vector<uint8_t> mapq(n); std::vector<uint32_t> off(n + 1, 0), cig;
    auto W = [](uint32_t len, uint32_t op) { return len << 4 | op; };
    for (uint32_t i = 0; i < n; i++) {
        flag[i] = (uint16_t)((rng() % 8 == 0) ? 1024 : (rng() % 2 ? 16 : 0)); mapq[i] = (uint8_t)(rng() % 61);
        const int k = has_off ? (int)(rng() % 10) : 0;
        switch (k) {
        case 0: case 1: case 2: case 3: cig.push_back(W(has_off ? 150 : 1 + (uint32_t)(rng() % 400), 0)); break;
        case 4: cig.push_back(W(1 + rng() % 40, 4)); cig.push_back(W(1 + rng() % 120, 0)); break;
        case 5: cig.push_back(W(1 + rng() % 80, 0)); cig.push_back(W(1 + rng() % 6000, 3)); cig.push_back(W(1 + rng() % 80, 0)); break;
        case 6: cig.push_back(W(1 + rng() % 60, 7)); cig.push_back(W(1 + rng() % 9, 1)); cig.push_back(W(1 + rng() % 60, 8)); cig.push_back(W(1 + rng() % 30, 2)); cig.push_back(W(1 + rng() % 60, 0)); break;
        case 7: cig.push_back(W(3, 2)); cig.push_back(W(1 + rng() % 100, 0)); cig.push_back(W(1, 6)); cig.push_back(W(1 + rng() % 100, 0)); cig.push_back(W(2, 5)); break;
        case 8: break;                                                              // "*"
        default: cig.push_back(W(1 + rng() % 3000, 0)); break;                      // a long read: far runs
        }
        off[i + 1] = (uint32_t)cig.size();
    }
    if (cig.empty()) cig.push_back(0);
    const uint32_t xflags = 1796; const int32_t min_mapq = 10;

    // ---- expected depth of ONE pass (uint32, modular like the reference's array), clipped to the owned range ----
    std::vector<uint32_t> want((size_t)clen, 0);
    for (uint32_t i = 0; i < n; i++) {
        if (has_filter && ((flag[i] & xflags) || mapq[i] < min_mapq)) continue;
        int64_t q = (int64_t)pos[i] - shift;
        const uint32_t c0 = has_off ? off[i] : i, c1 = has_off ? off[i + 1] : i + 1;
        for (uint32_t k = c0; k < c1; k++) {
            const uint32_t op = cig[k] & 15u; const int64_t len = cig[k] >> 4;
            if (is_match_op(op)) { for (int64_t b = std::max<int64_t>(q, 0); b < std::min<int64_t>(q + len, clen); b++) want[(size_t)b]++; q += len; }
            else if (is_skip_op(op)) q += len;
        }
    }

    // ---- the kernels, as pd_engine.cu queues them for a PD_SORTED batch ----
    const int32_t nseg = clen / SEG + 1;
    const size_t total = ((size_t)clen + 1 + 4095) / 4096 * 4096;
    std::vector<uint32_t> depth(total, 0xcccccccc), delta(total, 0), lo(nseg + 8), hi(nseg + 8), coarse(nseg / CUT_COARSE + 8), part_cov(nseg + 1, 0);
    std::vector<unsigned long long> part_sum(nseg + 1, 0);
    std::vector<uint8_t> seg_state(nseg + 1, 0);
    int flags_dev[2] = {0, 0};
    std::vector<uint32_t> bits(n / 32 + 40, 0);
    CutArgs c{pos.data(), n, 0, nseg, lo.data(), hi.data(), (int64_t)shift};
    const int32_t ncoarse = (nseg + CUT_COARSE - 1) / CUT_COARSE;
    auto cuts = [&] {   // one thread per key, no collectives
        for (int k = 0; k < ncoarse; k++) { threadIdx.x = (unsigned)(k % CUT_THREADS); blockIdx.x = (unsigned)(k / CUT_THREADS); blockDim.x = CUT_THREADS; k_cut_coarse(c, coarse.data(), ncoarse); }
        for (int j = 0; j <= nseg; j++) { threadIdx.x = (unsigned)(j % CUT_THREADS); blockIdx.x = (unsigned)(j / CUT_THREADS); blockDim.x = CUT_THREADS; k_cut(c, coarse.data(), ncoarse); }
    };
    cuts();
    if (has_filter) run((unsigned)((bits.size() * 32 + 255) / 256), 256, [&] { k_filter_bits(flag.data(), mapq.data(), n, xflags, min_mapq, bits.data(), (uint32_t)bits.size()); });
    TileArgs t{pos.data(), cig.data(), has_off ? off.data() : nullptr, has_filter ? flag.data() : nullptr, has_filter ? mapq.data() : nullptr, bits.data(),
               lo.data(), hi.data(), n, 0, nseg, clen, depth.data(), delta.data(), seg_state.data(), xflags, min_mapq, &flags_dev[0], &flags_dev[1],
               1u, 1u, 0xffffffffu, part_cov.data(), part_sum.data(), 24, shift, 0u, nullptr};
    auto tile = [&] {
        const unsigned grid = (unsigned)((nseg + WPC - 1) / WPC);
        if (has_off) { if (has_filter) run(grid, WPC * 32, [&] { k_tile_accumulate<true, true>(t); }); else run(grid, WPC * 32, [&] { k_tile_accumulate<true, false>(t); }); }
        else { if (has_filter) run(grid, WPC * 32, [&] { k_tile_accumulate<false, true>(t); }); else run(grid, WPC * 32, [&] { k_tile_accumulate<false, false>(t); }); }
    };
    int bad = 0;
    // depth as finalize leaves it: untouched segments are zero (k_fill_untouched), the delta array is prefix-summed in (k_scan_add)
    auto check = [&](uint32_t passes) {
        std::vector<uint32_t> d(depth);
        for (int32_t j = 0; j < nseg; j++) if (!seg_state[j]) for (int64_t b = (int64_t)j * SEG; b < (int64_t)(j + 1) * SEG; b++) d[(size_t)b] = 0;
        uint32_t acc = 0; size_t nbad = 0, first = 0;
        for (size_t b = 0; b < (size_t)clen; b++) { acc += delta[b]; d[b] += acc; if (d[b] != passes * want[b]) { if (!nbad) first = b; nbad++; } }
        if (nbad) { printf("pass %u: %zu bases differ, first at %zu: got %u want %u\n", passes, nbad, first, d[first], passes * want[first]); bad++; }
    };
    tile();                                                              // first pass: every touched segment is a first write
    check(1);
    // the reduce-hint partials describe what the warp wrote (table + carry; far runs of earlier segments are still in the delta
    // array -- the engine drops the partials of a contig whose deltas get scanned in): state 1 only inside the contig
    int with_part = 0;
    for (int32_t j = 0; j < nseg; j++) {
        if (seg_state[j] == 1 && (int64_t)(j + 1) * SEG > clen) { printf("segment %d reaches over the contig end but has partials\n", j); bad++; }
        if (seg_state[j] != 1) continue;
        uint32_t cov = 0; unsigned long long sum = 0;
        for (int64_t b = (int64_t)j * SEG; b < (int64_t)(j + 1) * SEG; b++) { if (depth[(size_t)b]) { cov++; sum += depth[(size_t)b]; } }
        if (part_cov[j] != cov || part_sum[j] != sum) { printf("segment %d: partials %u/%llu, written depth has %u/%llu\n", j, part_cov[j], part_sum[j], cov, sum); bad++; break; }
        with_part++;
    }
    tile();                                                              // second pass: every touched segment already holds depth
    check(2);
    for (int32_t j = 0; j < nseg; j++) if (seg_state[j] == 1) { printf("segment %d still says 'written once' after two passes\n", j); bad++; break; }
    if (flags_dev[0]) { printf("the kernel reported err bits %x on sorted input\n", flags_dev[0]); bad++; }
    // negative control: an out-of-order read must raise ERR_UNSORTED
    if (n > 64) {
        std::swap(pos[n / 2], pos[n / 2 + 40]);
        if (pos[n / 2] != pos[n / 2 + 40]) { tile(); if (!(flags_dev[0] & ERR_UNSORTED)) { printf("an out-of-order batch was not reported\n"); bad++; } }
    }
    printf("tile_emu_check: contig %d, %u sorted reads, has_off %d, has_filter %d, shift %d, %d segments (%d with partials), far runs %s: %s\n", clen, n,
           (int)has_off, (int)has_filter, shift, nseg, with_part, flags_dev[1] ? "yes" : "no", bad ? "MISMATCH" : "ok");
    return bad ? 1 : 0;
}
