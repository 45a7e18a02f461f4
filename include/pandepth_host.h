/*
 * pandepth_host.h -- C entry points of the host-side feeder (libpandepth_host.so), for callers that are not the `pandepth`
 * command line (bench.py's end-to-end arm, other language bindings).  Same code as the command line uses
 * (pandepth_b200/host/bgzf_feed.hpp).
 *
 * Role in the reference: the body of a worker thread, ProDealChrBambai (PanDepth.cpp:618-790) -- "open this BAM and put all
 * of its reads into the depth arrays" -- minus the reduce, which stays pd_reduce / pd_reduce_windows.
 */
#ifndef PANDEPTH_HOST_H_
#define PANDEPTH_HOST_H_
#include <stddef.h>
#include <stdint.h>
#include "pandepth_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pdh_feed_stats {
    uint64_t file_bytes;      /* compressed bytes handed to the engine (host -> device) */
    uint64_t pieces;          /* pd_push_bgzf calls */
    uint64_t bgzf_blocks;
    uint64_t reads;           /* records that passed the engine's filter */
} pdh_feed_stats;

/* Feed every record of the BAM file `path` into `e` (whole-genome modes: every read of every owned contig counts), using
 * `threads` reader threads and the engine's device-side decoder.  The file is read with pread (page cache or disk) into
 * pinned buffers that live until pdh_release().
 * Returns PD_OK; PD_ERR_FALLBACK when the file needs the host decode path (the engine has been reset; nothing is
 * accumulated); PD_ERR_ARG for an unreadable / corrupt file.  `err` (nullable) receives the message. */
int pdh_feed_bam(pd_engine *e, const char *path, int threads, pdh_feed_stats *stats, char *err, size_t err_cap);

/* Free the pinned buffers kept between calls. */
void pdh_release(void);

#ifdef __cplusplus
}
#endif
#endif
