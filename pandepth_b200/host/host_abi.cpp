// host_abi.cpp -- libpandepth_host.so: the feeder's C entry points (include/pandepth_host.h).
#include "../../include/pandepth_host.h"
#include "bgzf_feed.hpp"

#include <cstdio>
#include <memory>

namespace {
std::mutex g_mu;
std::unique_ptr<pdh::ChunkPool> g_pool;
}

extern "C" int pdh_feed_bam(pd_engine *e, const char *path, int threads, pdh_feed_stats *stats, char *err, size_t err_cap)
{
    return pdh_feed_bam_range(e, path, threads, 0, UINT64_MAX, stats, err, err_cap);
}

extern "C" uint64_t pdh_bai_block_offset(const char *path, int32_t tid, int64_t pos)
{
    if (!path) return UINT64_MAX;
    pdh::BaiLinear bai;
    if (!bai.load(std::string(path) + ".bai")) return UINT64_MAX;
    // the contig may have no read at or after pos: the next contig's first read is the boundary then
    for (int t = tid; t >= 0 && (size_t)t < bai.refs.size(); t++) {
        const uint64_t v = bai.start_voffset(t, t == tid ? pos : 0);
        if (v) return v >> 16;
    }
    return UINT64_MAX;
}

extern "C" int pdh_feed_bam_range(pd_engine *e, const char *path, int threads, uint64_t file_beg, uint64_t file_end,
                                  pdh_feed_stats *stats, char *err, size_t err_cap)
{
    if (!e || !path) return PD_ERR_ARG;
    std::lock_guard<std::mutex> g(g_mu);
    if (!g_pool) g_pool.reset(new pdh::ChunkPool);
    pd_stats before{}, after{};
    pd_get_stats(e, &before);
    const pdh::BgzfFeedResult r = pdh::bgzf_feed_file(e, nullptr, path, threads < 1 ? 1 : threads, *g_pool, true, file_beg, file_end);
    pd_get_stats(e, &after);
    if (stats) { stats->file_bytes = r.bytes; stats->pieces = r.pieces; stats->bgzf_blocks = r.blocks; stats->reads = after.reads_pushed - before.reads_pushed; }
    if (err && err_cap) snprintf(err, err_cap, "%s", r.err.c_str());
    if (r.status == 1) return PD_OK;
    if (r.status == 0) { pd_reset(e); return PD_ERR_FALLBACK; }
    return PD_ERR_ARG;
}

extern "C" void pdh_release(void)
{
    std::lock_guard<std::mutex> g(g_mu);
    g_pool.reset();
}
