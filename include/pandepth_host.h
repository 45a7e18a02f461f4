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

/* Range sharding over several engines / GPUs: feed only the BGZF blocks that start in the byte range [file_beg, file_end) of
 * the file (file_beg = 0: from the first record; file_end = UINT64_MAX: to the end).  The engine (pd_create_ranges) clips
 * whatever falls outside the range it owns, so neighbouring ranks may overlap by a few blocks. */
int pdh_feed_bam_range(pd_engine *e, const char *path, int threads, uint64_t file_beg, uint64_t file_end,
                       pdh_feed_stats *stats, char *err, size_t err_cap);

/* File offset of the BGZF block holding the first read that overlaps the 16 kb window of position `pos` (0-based) of contig
 * `tid`, from the linear index of `path`.bai -- every read of that contig starting at or after `pos` (and every read that
 * overlaps it) lies at or after this offset in a coordinate-sorted file.  UINT64_MAX: no such read (or no usable index). */
uint64_t pdh_bai_block_offset(const char *path, int32_t tid, int64_t pos);

/* Free the pinned buffers kept between calls. */
void pdh_release(void);

#ifdef __cplusplus
}
#endif
#endif
