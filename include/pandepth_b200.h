/*
 * pandepth_b200.h -- C ABI of the B200-native per-base depth engine (libpandepth_b200.so).
 *
 * Drop-in boundary for PanDepth's coverage hot path.  The reference has no plugin/FFI interface; the seam
 * is the in-process worker call
 *     ThreadsVector.emplace_back(ProDealChrBambai, BamPath, paraFA04, RegionMerger, ChrNumVer, GeneData, nBgzf)
 * (/root/reference/src/PanDepth.cpp:4512; variants :4248, :2813) plus the two reducers it calls.  A host that
 * keeps PanDepth's own plumbing (htslib decode, region model, report writer) replaces exactly
 *     - the filter + CIGAR walk `depth[p]++`              PanDepth.cpp:732-760 (clones :436-460, :572-596, :4611-4710)
 *     - StatChrDepthWin / StatChrDepthLowMEM               PanDepth.cpp:295-327 / :329-348
 *     - the mode-6 fixed-stride window loop                PanDepth.cpp:4366-4381
 *     - the -a per-site read-out                           PanDepth.cpp:4278-4281
 * with the calls below.  Plain pointers and sizes only; no exceptions cross the boundary; every function
 * returns a PD_* status (0 = ok) and pd_last_error() gives the text.  One engine per GPU; an engine is
 * thread-compatible (serialise calls per handle), different engines are independent.
 *
 * Record fields are BAM's own (include/sam.h:204-217): tid, 0-based pos, flag, mapq (core.qual), and CIGAR
 * words packed len<<4|op (include/sam.h:104-115).
 */
#ifndef PANDEPTH_B200_H_
#define PANDEPTH_B200_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pd_engine pd_engine;

enum {
    PD_OK = 0,
    PD_ERR_ARG = 1,       /* bad argument */
    PD_ERR_CUDA = 2,      /* CUDA runtime failure (text in pd_last_error) */
    PD_ERR_STATE = 3,     /* call not valid in the engine's current state */
    PD_ERR_UNSORTED = 4,  /* a batch pushed with PD_SORTED was not sorted by pos */
    PD_ERR_NOMEM = 5,
    PD_ERR_FALLBACK = 6   /* pd_push_bgzf only: this input needs the host decode path (pd_reset, then pd_push_contig) */
};

/* flags for pd_push* / pd_reduce* */
#define PD_SORTED 1u   /* records are sorted by pos within each tid (coordinate-sorted BAM order): enables the tile path */
#define PD_DEVICE 2u   /* every pointer argument is a device pointer on the engine's GPU */
#define PD_ASYNC  4u   /* host buffers are pinned and stay valid until pd_sync(): copies are not waited for */
#define PD_CRC    16u  /* pd_push_bgzf: every payload is followed in `bytes` by its 8-byte BGZF trailer (CRC32, ISIZE): the engine verifies the
                          CRC32 of every inflated block on the device, like bgzf_read_block does; a mismatch is PD_ERR_ARG */
#define PD_READY  8u   /* with PD_DEVICE: the arrays are complete NOW (not merely ordered by the engine's stream), so the engine may
                          index the batch on a side stream while earlier batches are still being accumulated */

/* Create an engine on CUDA device `device` for the BAM header's targets (sam_hdr_t::target_len,
 * read at PanDepth.cpp:684).  `owned` (nullable) selects the contigs this engine accumulates (multi-GPU
 * contig sharding, the reference's ChrNumVer bins :4479-4503); pushes for other tids are ignored.
 * xflags / min_mapq are -x / -q (In3str1v::flags, ::InInt; defaults 1796 / -1).  Contigs may be up to 2^31 - 8192 bases long
 * (BAM positions are int32).  NULL on failure (pd_last_error(NULL) has the reason). */
pd_engine *pd_create(int device, int n_targets, const uint32_t *target_len, const uint8_t *owned,
                     uint32_t xflags, int min_mapq);
/* Range sharding ("genomic ranges are independent if reads overlapping a boundary are fed to both
 * owners and each owner clips to its range"): the engine owns [own_beg[t], own_end[t]) of contig t (own_end <= own_beg: nothing)
 * and holds depth only for that.  Every call keeps speaking contig coordinates: pushes may carry any reads -- what falls
 * outside the owned range is clipped away, on both sides --, pd_reduce clips its intervals to the owned range and
 * pd_depth_read returns zeros outside it, so the results of engines with disjoint ranges ADD UP to the single-engine
 * result (one all-reduce of the cov / sum arrays, or a host-side sum).  pd_reduce_windows needs wholly owned contigs. */
pd_engine *pd_create_ranges(int device, int n_targets, const uint32_t *target_len, const uint32_t *own_beg, const uint32_t *own_end,
                            uint32_t xflags, int min_mapq);
void pd_destroy(pd_engine *e);
const char *pd_last_error(pd_engine *e);

/* Use the caller's CUDA stream (a cudaStream_t) for all kernels and copies; NULL restores the engine's own. */
int pd_set_stream(pd_engine *e, void *cuda_stream);

/* Optional: announce the (min_dep, depth_mask) pair pd_reduce will be called with (-d and the depth width are known before
 * the BAM is read).  While a segment of depth is written for the first time the engine then also records its
 * CoveredSite/TotalDepth, and pd_reduce with the same pair takes whole, written-once segments from those partials instead
 * of re-reading depth (per-chr / large-interval modes: the 4 B/base reduce pass disappears).  Purely an optimisation:
 * results are identical with or without it.  Call before the first push after pd_create / pd_reset. */
int pd_set_reduce_hint(pd_engine *e, uint32_t min_dep, uint32_t depth_mask);

/* Forget all accumulated depth (start a new BAM) without reallocating. */
int pd_reset(pd_engine *e);

/* Accumulate a batch of records of ONE contig: role of the `while (sam_itr_next(...))` body,
 * PanDepth.cpp:730-761.  flag/mapq may be NULL (no filtering on that field -- the feeder already applied
 * :732-733).  cigar_off (n_reads+1 offsets) may be NULL: then read i has exactly the one op cigar[i].
 * Any order of batches and any partition of the reads gives the same depth (it is additive). */
int pd_push_contig(pd_engine *e, int32_t tid, int64_t n_reads, const int32_t *pos,
                   const uint16_t *flag, const uint8_t *mapq,
                   const uint32_t *cigar_off, const uint32_t *cigar, uint32_t flags);

/* Same with a per-record tid array (host pointers only); split into per-contig runs internally. */
int pd_push(pd_engine *e, int64_t n_reads, const int32_t *tid, const int32_t *pos,
            const uint16_t *flag, const uint8_t *mapq,
            const uint32_t *cigar_off, const uint32_t *cigar, uint32_t flags);

/* Decode AND accumulate a chunk of a BAM file on the device: `blocks` are consecutive BGZF blocks (payload offset / length
 * inside `bytes`, inflated size <= 65536) whose first record starts `first_skip` bytes into the first block's inflated data
 * (the BAM header in front of it; 0 for every later chunk).  The engine inflates the blocks (one warp per block), walks the
 * BAM records (role of the sam_read1 / sam_itr_next loops, PanDepth.cpp:4608 / :730, and of the filter :732-733 with the
 * xflags / min_mapq given to pd_create) and accumulates them: runs of blocks that hold one contig in coordinate order take
 * the tile path, blocks with a contig change or records out of order take the delta-scatter path -- so any BAM gives the
 * same depth as pd_push_contig on the same records, chunks may arrive in any order, and a file of many small contigs costs
 * one launch per chunk, not per contig.
 *   PD_ASYNC   return once the chunk is queued (up to PD_BGZF_DEPTH chunks decode concurrently); its accumulate kernels are
 *              queued, and an error of this chunk is reported, by a later pd_push_bgzf / pd_sync / pd_finalize call.
 *              `bytes` must stay valid until PD_BGZF_DEPTH later pd_push_bgzf calls, or pd_sync, have returned.
 *   PD_DEVICE  `bytes` is a device pointer (>= 8 readable bytes behind the last payload); `blocks` stays a host pointer.
 * Returns PD_ERR_FALLBACK when the chunk holds something the device path does not handle -- a record that does not start
 * on a BGZF block boundary (writers other than htslib, reads larger than a block) or a CIGAR kept in the CG tag.  Depth
 * accumulated by earlier chunks stays in the engine: the caller resets it and feeds the file through pd_push_contig.
 * A block that does not inflate to its ISIZE -- or, with PD_CRC, to bytes with another CRC32 than its trailer says -- is PD_ERR_ARG. */
typedef struct pd_bgzf_block { uint64_t in_off; uint32_t in_len; uint32_t out_len; } pd_bgzf_block;
#ifndef PD_BGZF_DEPTH
#define PD_BGZF_DEPTH 6   /* PD_ASYNC: `bytes` of a call is no longer read once this many later pd_push_bgzf calls have returned */
#endif
int pd_push_bgzf(pd_engine *e, const uint8_t *bytes, size_t n_bytes, const pd_bgzf_block *blocks, int64_t n_blocks,
                 uint32_t first_skip, uint32_t flags);

/* How far pd_push_bgzf has come in a coordinate-sorted file: (refID, pos) of the first record of the last block, among the
 * chunks whose accumulate kernels have been queued (-1, -1 before that).  A feeder that only needs the reads below some
 * position (range sharding: "everything that starts before the end of my range") stops pushing once this passes it. */
int pd_bgzf_progress(pd_engine *e, int32_t *tid, int32_t *pos);

/* Wait until every queued push has been consumed (host buffers may be reused). */
int pd_sync(pd_engine *e);

/* End of input: per-base depth is materialised in HBM (uint32, modular like the reference's
 * `unsigned int *depth`, PanDepth.cpp:717).  Idempotent; further pushes re-open the engine. */
int pd_finalize(pd_engine *e);

/* Interval reduce -- StatChrDepthWin / StatChrDepthLowMEM.  For each i over [start0[i], end0_excl[i]) of
 * contig tid[i] (0-based, clipped to the contig):
 *     d = depth & depth_mask;  if (d >= min_dep) { cov[i]++; sum[i] += d; }
 * depth_mask is 0xFFFFFFFF for the uint32 path or 0x3FFFF for the 18-bit SiteInfo paths
 * (include/DataClass.h:84-87).  A 1-based closed CDS (first, second) is start0 = first-1, end0_excl = second. */
int pd_reduce(pd_engine *e, int64_t n_iv, const int32_t *tid, const int32_t *start0, const int32_t *end0_excl,
              uint32_t min_dep, uint32_t depth_mask, uint64_t *cov, uint64_t *sum, uint32_t flags);

/* Fixed-stride windows of one contig -- the mode-6 loop PanDepth.cpp:4366-4381 (and mode-5 pseudo-genes):
 * window k is [start0 + k*win, min(start0 + (k+1)*win, contig_len)). */
int pd_reduce_windows(pd_engine *e, int32_t tid, int32_t start0, int32_t win, int64_t n_win,
                      uint32_t min_dep, uint32_t depth_mask, uint32_t *cov, uint64_t *sum, uint32_t flags);

/* Copy per-base depth [start0, start0+n) of a contig to `out` (the -a dump, PanDepth.cpp:4278-4281). */
int pd_depth_read(pd_engine *e, int32_t tid, int32_t start0, int32_t n, uint32_t *out, uint32_t flags);

/* Pinned host memory for the feeder's batches (cudaHostAlloc / cudaFreeHost). */
void *pd_host_alloc(size_t bytes);
void pd_host_free(void *p);

/* Counters since create/reset (also the basis of bench.py's gpu_launches / h2d / d2h figures). */
typedef struct pd_stats {
    uint64_t reads_pushed;
    uint64_t kernel_launches;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t hbm_bytes_allocated;
    uint64_t tile_path_batches;    /* batches that took the sorted tile kernel */
    uint64_t delta_path_batches;   /* batches that took the global delta (atomic) kernel */
    uint64_t scan_passes;          /* prefix-scan passes run by pd_finalize */
} pd_stats;
int pd_get_stats(pd_engine *e, pd_stats *out);

/* Per-kernel device timing with CUDA events on the engine's stream (off by default; adds two event records
 * per launch).  pd_profile_get fills up to `cap` entries; returns the number of kernel classes. */
typedef struct pd_kernel_time {
    char name[32];
    uint64_t launches;
    double ms;       /* summed device time */
} pd_kernel_time;
int pd_profile_enable(pd_engine *e, int on);
int pd_profile_get(pd_engine *e, pd_kernel_time *out, int cap);

#ifdef __cplusplus
}
#endif
#endif
