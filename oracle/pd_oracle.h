/*
 * pd_oracle.h -- CPU restatement of PanDepth's per-base depth path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the *checker* for the CUDA engine.  It is imported only by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 * Nothing under pandepth_b200/ may link, import or execute it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
 * outputs of the reference binary itself (oracle/_ref/pandepth_ref, built from the sources
 * under /root/reference by oracle/Makefile) on the fixtures under tests/golden/, which
 * include the hand-verified known-answer vectors of SURVEY.md App. C.
 *
 * Each function cites the reference lines (src/PanDepth.cpp unless noted) it restates.
 */
#ifndef PD_ORACLE_H_
#define PD_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Record filter + CIGAR walk -> per-base increments.
 * Restates PanDepth.cpp:732-760 (clones :436-460, :572-596, :4611-4670, :4681-4710).
 *   depth      : n_targets pointers, depth[t] has target_len[t] uint32 entries (caller zeroed).
 *   cigar_off  : n_reads+1 offsets into cigar[]; NULL means exactly one op per read (cigar[i]).
 *   flag/mapq  : NULL means "not supplied" (no filtering on that field).
 * Divergence from the reference, by design: the reference does not clip to the contig length
 * (its arrays carry +500 / +2e7 slack, :4132 / :716); we clip aligned runs to [0, target_len).
 * Returns number of reads that passed the filter. */
int64_t pdo_accumulate(int n_targets, const uint32_t *target_len, uint32_t **depth,
                       int64_t n_reads, const int32_t *tid, const int32_t *pos,
                       const uint16_t *flag, const uint8_t *mapq,
                       const uint32_t *cigar_off, const uint32_t *cigar,
                       uint32_t xflags, int min_mapq);

/* Interval reduce.  Restates StatChrDepthWin :295-327 / StatChrDepthLowMEM :329-348 and the
 * mode-6 inner loop :4373-4380.  For each interval i over depth (one contig's array):
 *   for b in [start0[i], end0[i]): d = depth[b] & mask; if (d >= min_dep) { cov++; sum += d; }
 * mask = 0xFFFFFFFF for the uint32 path (:717), 0x3FFFF for the SiteInfo 18-bit paths
 * (include/DataClass.h:84-87). Intervals are clipped to [0, len). */
void pdo_reduce(const uint32_t *depth, uint32_t len, int64_t n_iv,
                const int32_t *start0, const int32_t *end0,
                uint32_t min_dep, uint32_t mask, uint64_t *cov, uint64_t *sum);

/* Pseudo-gene tiling when no -g/-b is given.  Restates :3995-4046 (mode 0: W = 1e7, mode 5: W = -w).
 * Emits 1-based closed [start,end] tiles; returns the count (call with out==NULL to size).
 * Reproduces the reference's quirks: contigs with len < 2 yield nothing; when len % W == 1 the
 * final 1-bp tile is dropped. */
int64_t pdo_tile_contig(int64_t len, int64_t W, int32_t *out_start1, int32_t *out_end1);

/* Mode-6 (-w < 150) window rows.  Restates :4366-4381: for (j=1; j<CC; j+=W) { End=min(j-1+W,CC); }
 * Emits 0-based half-open [start0,end0) = [j-1, End); row prints Start=j, End=End, Length=End-j+1. */
int64_t pdo_mode6_windows(int64_t len, int64_t W, int32_t *out_start0, int32_t *out_end0);

/* The two printed doubles of every row.  Restates :5023-5024 / :5074-5075 / :4382-4383:
 * Coverage = cov*100.0/len ; MeanDepth = sum*1.0/len, printed fixed with precision 2. */
void pdo_format_ratio(uint64_t cov, uint64_t sum, uint64_t len, char *cov_pct /*>=32*/, char *mean_depth /*>=32*/);

#ifdef __cplusplus
}
#endif
#endif
