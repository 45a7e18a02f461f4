/*
 * pd_standin.cpp -- TEST STAND-IN for libpandepth_b200.so (CPU, host-logic tests only).
 *
 * The product is the CUDA engine (pandepth_b200/csrc); it has no CPU path and fails loudly without a GPU.  The host
 * program around it (pandepth_b200/host: argument handling, region model, the three feeders, batching, report writers)
 * is plain C++ whose correctness does not depend on where depth is computed.  To let the CPU suite (-m "not gpu") run that
 * host program end to end against the golden outputs of the reference binary, tests/test_cli_cpu.py builds THIS file into a
 * temporary directory and injects it with LD_LIBRARY_PATH for the duration of a test.  It implements the C ABI of
 * include/pandepth_b200.h as a thin adapter over the oracle (oracle/pd_oracle.c).
 *
 * It is never built in-tree, never shipped, never loaded by pandepth_b200/engine.py, bench.py or smoke(), and no GPU test
 * uses it: the GPU tests run the same CLI cases against the real engine.
 */
#include "../../include/pandepth_b200.h"
#include "../../oracle/pd_oracle.h"
#include "../../pandepth_b200/csrc/pd_decode.hpp"      // the block-parallel decoder, here on a host backend

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct pd_engine {
    std::vector<uint32_t> tlen;
    std::vector<uint8_t> owned;
    std::vector<uint32_t> own_beg, own_end;              // range sharding: depth is kept whole, results are clipped to the owned range
    int32_t prog_tid = -1, prog_pos = -1;
    std::vector<std::vector<uint32_t>> depth;
    uint32_t xflags; int min_mapq;
    bool pushed = false;
    pd_stats stats{};
    std::string err;
};

static std::string g_err;
// PD_STANDIN_TRACE=<file>: one line per pd_push_bgzf call, so a test can tell the decoder ran from the fallback having run
static void trace(const char *what, long long a, long long b)
{
    if (const char *p = getenv("PD_STANDIN_TRACE")) if (FILE *f = fopen(p, "a")) { fprintf(f, "%s %lld %lld\n", what, a, b); fclose(f); }
}

static int fail(pd_engine *e, int code, const char *m) { if (e) e->err = m; else g_err = m; return code; }

extern "C" {

const char *pd_last_error(pd_engine *e) { return e ? e->err.c_str() : g_err.c_str(); }
void *pd_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void pd_host_free(void *p) { free(p); }
void pd_destroy(pd_engine *e) { delete e; }

pd_engine *pd_create_ranges(int, int n_targets, const uint32_t *target_len, const uint32_t *own_beg, const uint32_t *own_end, uint32_t xflags, int min_mapq)
{
    if (n_targets <= 0 || !target_len || !own_beg || !own_end) { g_err = "pd_create: bad arguments"; return nullptr; }
    std::vector<uint8_t> owned((size_t)n_targets);
    for (int t = 0; t < n_targets; t++) owned[(size_t)t] = own_end[t] > own_beg[t];
    pd_engine *e = pd_create(0, n_targets, target_len, owned.data(), xflags, min_mapq);
    if (e) for (int t = 0; t < n_targets; t++) { e->own_beg[(size_t)t] = std::min(own_beg[t], target_len[t]); e->own_end[(size_t)t] = std::min(own_end[t], target_len[t]); }
    return e;
}

int pd_bgzf_progress(pd_engine *e, int32_t *tid, int32_t *pos) { if (!e || !tid || !pos) return PD_ERR_ARG; *tid = e->prog_tid; *pos = e->prog_pos; return PD_OK; }

pd_engine *pd_create(int, int n_targets, const uint32_t *target_len, const uint8_t *owned, uint32_t xflags, int min_mapq)
{
    if (n_targets <= 0 || !target_len) { g_err = "pd_create: bad arguments"; return nullptr; }
    pd_engine *e = new pd_engine;
    e->tlen.assign(target_len, target_len + n_targets);
    e->owned.assign((size_t)n_targets, 1);
    if (owned) e->owned.assign(owned, owned + n_targets);
    e->own_beg.assign((size_t)n_targets, 0); e->own_end.assign(target_len, target_len + n_targets);
    e->depth.resize((size_t)n_targets);
    for (int t = 0; t < n_targets; t++) if (e->owned[(size_t)t]) e->depth[(size_t)t].assign(e->tlen[(size_t)t], 0u);
    e->xflags = xflags; e->min_mapq = min_mapq;
    return e;
}

int pd_set_stream(pd_engine *e, void *) { return e ? PD_OK : PD_ERR_ARG; }

int pd_set_reduce_hint(pd_engine *e, uint32_t, uint32_t)
{
    if (!e) return PD_ERR_ARG;
    if (e->pushed) return fail(e, PD_ERR_STATE, "pd_set_reduce_hint must be called before the first push (after pd_create or pd_reset)");
    return PD_OK;
}

int pd_reset(pd_engine *e)
{
    if (!e) return PD_ERR_ARG;
    for (auto &d : e->depth) std::fill(d.begin(), d.end(), 0u);
    e->pushed = false; e->stats = pd_stats{}; e->err.clear(); e->prog_tid = e->prog_pos = -1;
    return PD_OK;
}

int pd_push_contig(pd_engine *e, int32_t tid, int64_t n, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
                   const uint32_t *cigar_off, const uint32_t *cigar, uint32_t flags)
{
    if (!e) return PD_ERR_ARG;
    if (flags & PD_DEVICE) return fail(e, PD_ERR_ARG, "stand-in: device pointers are not supported");
    if (tid < 0 || tid >= (int32_t)e->tlen.size() || n < 0 || (n > 0 && (!pos || !cigar))) return fail(e, PD_ERR_ARG, "pd_push_contig: bad arguments");
    if (!e->owned[(size_t)tid] || n == 0 || e->tlen[(size_t)tid] == 0) return PD_OK;
    if (flags & PD_SORTED) for (int64_t i = 1; i < n; i++) if (pos[i] < pos[i - 1]) return fail(e, PD_ERR_UNSORTED, "a batch pushed with PD_SORTED was not sorted by pos");
    std::vector<int32_t> tids((size_t)n, tid);
    std::vector<uint32_t *> dp(e->tlen.size(), nullptr);
    for (size_t t = 0; t < dp.size(); t++) dp[t] = e->depth[t].empty() ? nullptr : e->depth[t].data();
    // cigar_off is absolute into `cigar` (same convention as the engine)
    pdo_accumulate((int)e->tlen.size(), e->tlen.data(), dp.data(), n, tids.data(), pos, flag, mapq, cigar_off, cigar, e->xflags, e->min_mapq);
    e->pushed = true; e->stats.reads_pushed += (uint64_t)n;
    return PD_OK;
}

int pd_push(pd_engine *e, int64_t n, const int32_t *tid, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
            const uint32_t *cigar_off, const uint32_t *cigar, uint32_t flags)
{
    if (!e) return PD_ERR_ARG;
    int64_t b = 0;
    while (b < n) {
        int64_t end = b + 1;
        while (end < n && tid[end] == tid[b]) ++end;
        if (tid[b] >= 0 && tid[b] < (int32_t)e->tlen.size()) {
            int rc = pd_push_contig(e, tid[b], end - b, pos + b, flag ? flag + b : nullptr, mapq ? mapq + b : nullptr,
                                    cigar_off ? cigar_off + b : nullptr, cigar_off ? cigar : cigar + b, flags);
            if (rc) return rc;
        }
        b = end;
    }
    return PD_OK;
}

namespace {
// pd_decode.hpp's backend interface with plain loops (what the CUDA backend does with one block per warp / thread)
struct HostBackend {
    std::vector<uint8_t> inflated_v; std::vector<pdk::BlockCount> counts_v;
    std::vector<int32_t> tid, pos; std::vector<uint32_t> cigar_off, cigar;
    uint8_t *inflated = nullptr; pdk::BlockCount *counts_host = nullptr;
    int inflate(const uint8_t *in, size_t, const pdk::ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint64_t out_total)
    {
        inflated_v.assign(out_total + 8, 0); inflated = inflated_v.data();
        static thread_local pdk::InflateTables T;       // (several stand-in engines may decode at once: PANDEPTH_DEVICES)
        for (uint32_t b = 0; b < n; b++)
            if (blk[b].out_len && pdk::inflate_raw(in + blk[b].in_off, blk[b].in_len, inflated + out_off[b], blk[b].out_len, T) != (int)blk[b].out_len) return -1;
        return 0;
    }
    void count(const pdk::ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint32_t first_skip, int32_t n_targets, uint32_t xflags, int32_t min_mapq)
    {
        counts_v.resize(n); counts_host = counts_v.data();
        for (uint32_t b = 0; b < n; b++) { const uint32_t sk = b == 0 ? first_skip : 0; counts_v[b] = pdk::count_block(inflated + out_off[b] + sk, blk[b].out_len - sk, n_targets, xflags, min_mapq); }
    }
    void emit(const pdk::ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint32_t first_skip, const uint32_t *r0, const uint32_t *c0, uint32_t n_reads, uint32_t n_cigar,
              int32_t n_targets, uint32_t xflags, int32_t min_mapq)
    {
        tid.assign(n_reads, -7); pos.assign(n_reads, -7); cigar_off.assign((size_t)n_reads + 1, 0); cigar.assign(n_cigar, 0);
        for (uint32_t b = 0; b < n; b++) { const uint32_t sk = b == 0 ? first_skip : 0; pdk::emit_block(inflated + out_off[b] + sk, blk[b].out_len - sk, n_targets, xflags, min_mapq, r0[b], c0[b], tid.data(), pos.data(), cigar_off.data(), cigar.data()); }
        cigar_off[n_reads] = n_cigar;
    }
    int32_t tid_at(uint32_t r) { return tid[r]; }
};
}  // namespace

int pd_push_bgzf(pd_engine *e, const uint8_t *bytes, size_t n_bytes, const pd_bgzf_block *blocks, int64_t n_blocks, uint32_t first_skip, uint32_t)
{
    if (!e) return PD_ERR_ARG;
    if (n_blocks <= 0) return PD_OK;
    std::vector<pdk::ChunkBlock> blk((size_t)n_blocks);
    for (int64_t b = 0; b < n_blocks; b++) blk[(size_t)b] = pdk::ChunkBlock{blocks[b].in_off, blocks[b].in_len, blocks[b].out_len};
    HostBackend be; std::vector<pdk::ContigRun> runs; uint32_t n_reads = 0, n_cigar = 0;
    const int rc = pdk::decode_chunk(be, bytes, n_bytes, blk, first_skip, (int32_t)e->tlen.size(), e->xflags, e->min_mapq, runs, n_reads, n_cigar);
    if (rc == pdk::DEC_STRADDLE || rc == pdk::DEC_CG_TAG) { trace("bgzf_fallback", n_blocks, rc); return fail(e, PD_ERR_FALLBACK, "pd_push_bgzf: chunk not handled by the block-parallel decoder"); }
    if (rc != pdk::DEC_OK) return fail(e, PD_ERR_ARG, "pd_push_bgzf: malformed BGZF / BAM data");
    for (const pdk::ContigRun &r : runs) {
        // the filter was applied by the decoder: no flag / mapq arrays; cigar_off is absolute into cigar[]
        const int prc = pd_push_contig(e, r.tid, (int64_t)(r.r1 - r.r0), be.pos.data() + r.r0, nullptr, nullptr, be.cigar_off.data() + r.r0, be.cigar.data(), r.sorted ? PD_SORTED : 0u);
        if (prc) return prc;
    }
    for (int64_t b = n_blocks; b-- > 0;) if (be.counts_host[b].n_reads) { e->prog_tid = be.counts_host[b].first_tid; e->prog_pos = be.counts_host[b].first_pos; break; }
    trace("bgzf_ok", n_blocks, n_reads);
    return PD_OK;
}

int pd_sync(pd_engine *e) { return e ? PD_OK : PD_ERR_ARG; }
int pd_finalize(pd_engine *e) { return e ? PD_OK : PD_ERR_ARG; }

int pd_reduce(pd_engine *e, int64_t n_iv, const int32_t *tid, const int32_t *s0, const int32_t *e0, uint32_t min_dep, uint32_t mask,
              uint64_t *cov, uint64_t *sum, uint32_t)
{
    if (!e) return PD_ERR_ARG;
    for (int64_t i = 0; i < n_iv; i++) {
        cov[i] = sum[i] = 0;
        if (tid[i] < 0 || tid[i] >= (int32_t)e->tlen.size()) return fail(e, PD_ERR_ARG, "pd_reduce: tid out of range");
        const auto &d = e->depth[(size_t)tid[i]];
        if (d.empty()) continue;
        const int32_t a = std::max<int32_t>(s0[i], (int32_t)e->own_beg[(size_t)tid[i]]), z = std::min<int32_t>(e0[i], (int32_t)e->own_end[(size_t)tid[i]]);
        if (a >= z) continue;                                // (clipped to the owned range: results of disjoint engines add up)
        pdo_reduce(d.data(), (uint32_t)d.size(), 1, &a, &z, min_dep, mask, cov + i, sum + i);
    }
    return PD_OK;
}

int pd_reduce_windows(pd_engine *e, int32_t tid, int32_t start0, int32_t win, int64_t n_win, uint32_t min_dep, uint32_t mask,
                      uint32_t *cov, uint64_t *sum, uint32_t)
{
    if (!e || tid < 0 || tid >= (int32_t)e->tlen.size() || win < 1) return PD_ERR_ARG;
    const auto &d = e->depth[(size_t)tid];
    for (int64_t k = 0; k < n_win; k++) {
        cov[k] = 0; sum[k] = 0;
        if (d.empty()) continue;
        const int64_t a = (int64_t)start0 + k * win, b = a + win;
        if (a >= (int64_t)d.size()) continue;
        const int32_t a32 = (int32_t)a, b32 = (int32_t)(b > (int64_t)d.size() ? (int64_t)d.size() : b);
        uint64_t c = 0;
        pdo_reduce(d.data(), (uint32_t)d.size(), 1, &a32, &b32, min_dep, mask, &c, sum + k);
        cov[k] = (uint32_t)c;
    }
    return PD_OK;
}

int pd_depth_read(pd_engine *e, int32_t tid, int32_t start0, int32_t n, uint32_t *out, uint32_t)
{
    if (!e || tid < 0 || tid >= (int32_t)e->tlen.size() || start0 < 0 || n < 0) return PD_ERR_ARG;
    if ((int64_t)start0 + n > (int64_t)e->tlen[(size_t)tid]) return fail(e, PD_ERR_ARG, "pd_depth_read: range beyond the contig");
    const auto &d = e->depth[(size_t)tid];
    if (d.empty()) memset(out, 0, (size_t)n * 4); else memcpy(out, d.data() + start0, (size_t)n * 4);
    return PD_OK;
}

int pd_get_stats(pd_engine *e, pd_stats *out) { if (!e || !out) return PD_ERR_ARG; *out = e->stats; return PD_OK; }
int pd_profile_enable(pd_engine *e, int) { return e ? PD_OK : PD_ERR_ARG; }
int pd_profile_get(pd_engine *e, pd_kernel_time *, int) { return e ? 0 : -PD_ERR_ARG; }

}  // extern "C"
