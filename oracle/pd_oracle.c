/*
 * pd_oracle.c -- CPU restatement of PanDepth's per-base depth path.
 * TEST INFRASTRUCTURE ONLY (see pd_oracle.h).  Plain C, single thread, no dependencies.
 * Parity status: PINNED against the reference binary (tests/test_oracle_golden.py).
 */
#include "pd_oracle.h"
#include <stdio.h>
#include <string.h>

/* BAM CIGAR packing: include/sam.h:104-115 -- op = c & 0xf, len = c >> 4. */
#define OP(c) ((c) & 0xfu)
#define OPLEN(c) ((c) >> 4)

int64_t pdo_accumulate(int n_targets, const uint32_t *target_len, uint32_t **depth,
                       int64_t n_reads, const int32_t *tid, const int32_t *pos,
                       const uint16_t *flag, const uint8_t *mapq,
                       const uint32_t *cigar_off, const uint32_t *cigar,
                       uint32_t xflags, int min_mapq)
{
    int64_t kept = 0;
    for (int64_t i = 0; i < n_reads; i++) {
        /* :732  if (aln->core.flag & flags) continue;   :733  if (qual < InInt) continue; */
        if (flag && (flag[i] & xflags)) continue;
        if (mapq && (int)mapq[i] < min_mapq) continue;
        int t = tid[i];
        if (t < 0 || t >= n_targets) continue;      /* unmapped: reference would index out of range */
        kept++;
        uint32_t *d = depth[t];
        int64_t len = target_len[t];
        int64_t p = pos[i];                          /* :736 0-based start */
        uint32_t c0 = cigar_off ? cigar_off[i] : (uint32_t)i;
        uint32_t c1 = cigar_off ? cigar_off[i + 1] : (uint32_t)i + 1;
        for (uint32_t k = c0; k < c1; k++) {
            uint32_t op = OP(cigar[k]);
            int64_t n = OPLEN(cigar[k]);
            switch (op) {                            /* :743-758 */
            case 0: case 7: case 8: {                /* M, =, X : depth[p..p+n)++ */
                int64_t e = p + n;
                int64_t a = p < 0 ? 0 : p, b = e > len ? len : e;
                for (int64_t q = a; q < b; q++) d[q]++;
                p = e;
                break;
            }
            case 2: case 3:                          /* D, N : advance only */
                p += n;
                break;
            default:                                 /* I,S,H,P,B : ignored */
                break;
            }
        }
    }
    return kept;
}

void pdo_reduce(const uint32_t *depth, uint32_t len, int64_t n_iv,
                const int32_t *start0, const int32_t *end0,
                uint32_t min_dep, uint32_t mask, uint64_t *cov, uint64_t *sum)
{
    for (int64_t i = 0; i < n_iv; i++) {
        int64_t a = start0[i], b = end0[i];
        if (a < 0) a = 0;
        if (b > (int64_t)len) b = len;
        uint64_t c = 0, s = 0;
        for (int64_t q = a; q < b; q++) {            /* :317-324 */
            uint32_t dd = depth[q] & mask;
            if (dd >= min_dep) { c++; s += dd; }
        }
        cov[i] = c;
        sum[i] = s;
    }
}

int64_t pdo_tile_contig(int64_t len, int64_t W, int32_t *out_start1, int32_t *out_end1)
{
    /* :3997-4046 -- note End is pre-set to 2 and bumped by 2 after each tile. */
    int64_t n = 0;
    int64_t Start = 1, End = 2;
    for (Start = 1; End <= len; Start += W) {
        End = Start + W - 1;
        if (End > len) End = len;
        if (out_start1) { out_start1[n] = (int32_t)Start; out_end1[n] = (int32_t)End; }
        n++;
        End += 2;
    }
    return n;
}

int64_t pdo_mode6_windows(int64_t len, int64_t W, int32_t *out_start0, int32_t *out_end0)
{
    /* :4366-4370 */
    int64_t n = 0;
    for (int64_t j = 1; j < len; j += W) {
        int64_t Start = j - 1, End = Start + W;
        if (End > len) End = len;
        if (out_start0) { out_start0[n] = (int32_t)Start; out_end0[n] = (int32_t)End; }
        n++;
    }
    return n;
}

void pdo_format_ratio(uint64_t cov, uint64_t sum, uint64_t len, char *cov_pct, char *mean_depth)
{
    double Coverage = cov * 100.0 / len;             /* :5023 */
    double MeanDepth = sum * 1.0 / len;              /* :5024 */
    snprintf(cov_pct, 32, "%.2f", Coverage);         /* fixed, setprecision(2) */
    snprintf(mean_depth, 32, "%.2f", MeanDepth);
}
