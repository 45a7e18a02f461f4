// CPU check of the bucketed unsorted path's kernels (pandepth_b200/csrc/pd_bucket.cuh) through the warp emulator (warp_emu.h):
//   bucket_emu_check [contig_len] [reads] [seed] [shift]
// A random unsorted batch (multi-op CIGARs, flags, mapq; some reads left of / over the ends of the owned range) goes through
// k_expand_runs -> k_bucket_hist -> k_bucket_scan -> k_bucket_scatter -> k_bucket_fine -> k_bucket_cuts exactly as the engine
// queues them, and the result is compared with a plain host walk:
//   * the bucketed runs are the expected multiset of (start, len) runs, nothing lost, nothing invented;
//   * they are in fine-bucket (256-base) order, fine_off[] points at every bucket's first run;
//   * lo[j] / hi[j] equal the lower bounds of s0_j - 256 / s0_j over the sorted starts -- what k_cut computes for sorted input;
//   * k_tile_accumulate over the bucketed runs (unordered mode) gives the depth of the expected runs, base by base.
#define PD_WARP_EMU 1
#include "warp_emu.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <utility>

#include "pd_kernels.cuh"
#include "pd_bucket.cuh"
using namespace pdk;

template <typename F> static void run(unsigned grid, unsigned block, F f) { emu::launch(grid, block, [&](int) { f(); }, 0); }

int main(int argc, char **argv)
{
    const int32_t clen = argc > 1 ? atoi(argv[1]) : 1300000;            // three coarse buckets
    const uint32_t n = argc > 2 ? (uint32_t)atoi(argv[2]) : 40000;
    const uint64_t seed = argc > 3 ? (uint64_t)atoll(argv[3]) : 1;
    const int32_t shift = argc > 4 ? atoi(argv[4]) : 0;                 // the owned range starts at this contig coordinate
    std::mt19937_64 rng(seed);
    std::vector<int32_t> pos(n); std::vector<uint16_t> flag(n); std::vector<uint8_t> mapq(n);
    std::vector<uint32_t> off(n + 1, 0), cig;
    auto W = [](uint32_t len, uint32_t op) { return len << 4 | op; };
    for (uint32_t i = 0; i < n; i++) {
        pos[i] = (int32_t)(rng() % (uint64_t)(clen + 600)) - 300 + shift;           // some left of the range, some over its end
        if (rng() % 50 == 0) pos[i] = -1;
        flag[i] = (uint16_t)((rng() % 8 == 0) ? 1024 : (rng() % 2 ? 16 : 0)); mapq[i] = (uint8_t)(rng() % 61);
        switch (rng() % 8) {
        case 0: case 1: case 2: cig.push_back(W(150, 0)); break;
        case 3: cig.push_back(W(1 + rng() % 40, 4)); cig.push_back(W(1 + rng() % 120, 0)); break;
        case 4: cig.push_back(W(1 + rng() % 80, 0)); cig.push_back(W(1 + rng() % 3000, 3)); cig.push_back(W(1 + rng() % 80, 0)); break;
        case 5: cig.push_back(W(1 + rng() % 60, 7)); cig.push_back(W(1 + rng() % 9, 1)); cig.push_back(W(1 + rng() % 60, 8)); cig.push_back(W(1 + rng() % 30, 2)); cig.push_back(W(1 + rng() % 60, 0)); break;
        case 6: break;                                                              // "*": no CIGAR at all
        default: cig.push_back(W(1 + rng() % 900, 0)); break;
        }
        off[i + 1] = (uint32_t)cig.size();
    }
    if (cig.empty()) cig.push_back(0);
    const uint32_t cap = off[n], xflags = 1796; const int32_t min_mapq = 10;

    // ---- expected runs: the walk of scatter_one / the oracle, clipped to the owned range ----
    std::vector<std::pair<int32_t, uint32_t>> want;
    for (uint32_t i = 0; i < n; i++) {
        if ((flag[i] & xflags) || mapq[i] < min_mapq || pos[i] < 0) continue;
        int64_t q = (int64_t)pos[i] - shift, rs = 0, re = 0; bool open = false;
        auto flush = [&] { if (!open) return; int64_t s = std::max<int64_t>(rs, 0), e = std::min<int64_t>(re, clen); if (s < e) want.emplace_back((int32_t)s, (uint32_t)(e - s) << 4); };
        for (uint32_t k = off[i]; k < off[i + 1]; k++) {
            const uint32_t op = cig[k] & 15u; const int64_t len = cig[k] >> 4;
            if (is_match_op(op)) { if (open && re == q) re = q + len; else { flush(); rs = q; re = q + len; open = true; } q += len; }
            else if (is_skip_op(op)) q += len;
        }
        flush();
    }
    std::sort(want.begin(), want.end());

    // ---- the kernels, as pd_engine.cu queues them ----
    const int n_coarse = (int)(((uint64_t)clen + (1u << BK_COARSE_SHIFT) - 1) >> BK_COARSE_SHIFT);
    const size_t n_fine = (size_t)n_coarse * BK_FINE;
    const int32_t nseg = clen / SEG + 1;
    std::vector<int32_t> pa(cap + 64, 0x7badbad), pb(cap + 64, 0x7badbad); std::vector<uint32_t> ca(cap + 64), cb(cap + 64);
    std::vector<uint32_t> hist(BK_MAX_COARSE + 1, 0), base(BK_MAX_COARSE + 1, 0), cursor(BK_MAX_COARSE + 1, 0), fine(n_fine, 0xdeadbeef), lo(nseg + 1), hi(nseg + 2);
    std::vector<uint32_t> delta((size_t)clen + 8, 0); int delta_used = 0;
    ExpandArgs x{pos.data(), cig.data(), off.data(), flag.data(), mapq.data(), n, clen, shift, xflags, min_mapq, pa.data(), ca.data(), cap, delta.data(), &delta_used};
    for (unsigned b = 0; b < (n + 255) / 256; b++) for (unsigned t = 0; t < 256; t++) { threadIdx.x = t; blockIdx.x = b; blockDim.x = 256; k_expand_runs(x); }   // no collectives
    run((cap + BK_TILE - 1) / BK_TILE, 256, [&] { k_bucket_hist(pa.data(), cap, hist.data(), n_coarse); });
    run(1, 1024, [&] { k_bucket_scan(hist.data(), base.data(), cursor.data(), n_coarse); });
    run((cap + BK_STAGE - 1) / BK_STAGE, 256, [&] { k_bucket_scatter(pa.data(), ca.data(), cap, base.data(), cursor.data(), n_coarse, pb.data(), cb.data()); });
    run((unsigned)n_coarse, 512, [&] { k_bucket_fine(pb.data(), cb.data(), base.data(), fine.data(), pa.data(), ca.data()); });
    run((nseg + 1 + 255) / 256, 256, [&] { k_bucket_cuts(fine.data(), base.data() + n_coarse, (uint32_t)n_fine, 0, nseg, lo.data(), hi.data()); });

    int bad = 0;
    const uint32_t total = base[n_coarse];
    if (total != want.size()) { printf("run count %u, expected %zu\n", total, want.size()); bad++; }
    std::vector<std::pair<int32_t, uint32_t>> got;
    for (uint32_t i = 0; i < total && i < cap; i++) got.emplace_back(pa[i], ca[i]);
    // fine-bucket order and offsets
    for (uint32_t i = 1; i < got.size(); i++) if ((got[i].first >> BK_FINE_SHIFT) < (got[i - 1].first >> BK_FINE_SHIFT)) { printf("run %u out of bucket order\n", i); bad++; break; }
    for (size_t f = 0; f < n_fine && !bad; f++) {
        const uint32_t o = fine[f];
        if (o > total || (o < total && (size_t)(got[o].first >> BK_FINE_SHIFT) < f) || (o > 0 && (size_t)(got[o - 1].first >> BK_FINE_SHIFT) >= f)) { printf("fine_off[%zu] = %u is not the first run of its bucket\n", f, o); bad++; }
    }
    std::vector<std::pair<int32_t, uint32_t>> sorted = got;
    std::sort(sorted.begin(), sorted.end());
    if (sorted != want) { printf("the bucketed runs are not the expected runs\n"); bad++; }
    // cuts == lower bounds over the sorted starts
    std::vector<int32_t> starts; for (auto &r : sorted) starts.push_back(r.first);
    for (int32_t j = 0; j <= nseg && !bad; j++) {
        const int64_t s0 = (int64_t)j * SEG;
        const uint32_t wl = (uint32_t)(std::lower_bound(starts.begin(), starts.end(), (int32_t)std::max<int64_t>(s0 - HALO, INT32_MIN)) - starts.begin());
        const uint32_t wh = (uint32_t)(std::lower_bound(starts.begin(), starts.end(), (int32_t)std::min<int64_t>(s0, INT32_MAX)) - starts.begin());
        if ((j < nseg && lo[j] != wl) || hi[j] != wh) { printf("segment %d: cuts %u/%u, expected %u/%u\n", j, lo[j], hi[j], wl, wh); bad++; }
    }
    // ---- and the tile kernel over the bucketed runs (unordered = 1, device-side run count), as the engine launches it: per-base depth
    // (+ the far-run deltas of expand and tile, prefix-summed as k_scan_add would) equals the depth the expected runs give ----
    if (!bad) {
        const size_t tot = ((size_t)clen + 1 + 8191) / 8192 * 8192;
        std::vector<uint32_t> depth(tot, 0xcccccccc), dl(tot, 0), part_cov(nseg + 1, 0), wantd((size_t)clen, 0);
        std::vector<unsigned long long> part_sum(nseg + 1, 0); std::vector<uint8_t> seg_state(nseg + 1, 0);
        for (size_t b = 0; b < (size_t)clen + 1 && b < delta.size(); b++) dl[b] = delta[b];          // what k_expand_runs scattered itself (runs >= 2^28: none here)
        int flags_dev[2] = {0, 0};
        TileArgs t{pa.data(), ca.data(), nullptr, nullptr, nullptr, nullptr, lo.data(), hi.data(), cap, 0, nseg, clen, depth.data(), dl.data(), seg_state.data(), xflags, min_mapq,
                   &flags_dev[0], &flags_dev[1], 1u, 1u, 0xffffffffu, part_cov.data(), part_sum.data(), 24, 0, 1u, base.data() + n_coarse};
        run((unsigned)((nseg + WPC - 1) / WPC), WPC * 32, [&] { k_tile_accumulate<false, false>(t); });
        for (auto &r : want) for (uint32_t b = 0; b < (r.second >> 4); b++) wantd[(size_t)r.first + b]++;
        if (flags_dev[0]) { printf("the tile kernel reported err bits %x on bucketed runs\n", flags_dev[0]); bad++; }
        uint32_t acc = 0; size_t nd = 0;
        for (size_t b = 0; b < (size_t)clen; b++) { acc += dl[b]; const uint32_t d = (seg_state[b / SEG] ? depth[b] : 0u) + acc; if (d != wantd[b]) nd++; }
        if (nd) { printf("depth over the bucketed runs differs on %zu bases\n", nd); bad++; }
    }
    printf("bucket_emu_check: contig %d, %u reads (%u CIGAR words), shift %d: %zu runs in %d coarse / %zu fine buckets, %d segments: %s\n",
           clen, n, cap, shift, want.size(), n_coarse, n_fine, nseg, bad ? "MISMATCH" : "ok");
    return bad ? 1 : 0;
}
