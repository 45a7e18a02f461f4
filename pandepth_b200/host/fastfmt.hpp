// fastfmt.hpp -- row formatting for the report writers (mode-6 window rows, -a per-site rows): the reference prints
// through ostream / setprecision(2) (PanDepth.cpp:4384-4389, :4278-4281); 3x10^7 window rows or 3x10^9 site rows make the
// formatter itself a hot loop, so integers are written by hand and the two ratios by an exact integer shortcut.
#pragma once
#include <cstdint>
#include <climits>
#include <cstdio>
#include <cstring>
#include <string>

namespace pdh {

// decimal digits of v at p, two at a time; returns the end
inline char *put_u64(char *p, uint64_t v)
{
    static const char LUT[201] =
        "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
    char tmp[20];
    int n = 20;
    while (v >= 100) { const unsigned r = (unsigned)(v % 100); v /= 100; n -= 2; memcpy(tmp + n, LUT + 2 * r, 2); }
    if (v >= 10) { n -= 2; memcpy(tmp + n, LUT + 2 * v, 2); }
    else tmp[--n] = (char)('0' + v);
    memcpy(p, tmp + n, (size_t)(20 - n));
    return p + (20 - n);
}

inline char *put_i64(char *p, int64_t v)
{
    if (v < 0) { *p++ = '-'; return put_u64(p, (uint64_t)0 - (uint64_t)v); }
    return put_u64(p, (uint64_t)v);
}

// printf("%.2f", (double)N / (double)D) for integers 0 <= N, 0 < D, written without going through printf.
//
// Why it is exact: v = fl(N/D) differs from the rational N/D by at most 2^-53 * N/D.  printf rounds the exact binary value
// of v to the nearest hundredth; v and N/D round alike unless a tie point (2k+1)/200 lies between them.  With X = 100*N =
// q*D + rem, the rational is q + rem/D hundredths; if 2*rem != D its distance from the nearest tie is >= 1/(200*D), which
// exceeds the error of v whenever N < 2^53/200 (we require N <= 10^13).  Exact ties (2*rem == D) are never representable
// in binary (odd numerator over 200), so v lies strictly on one side and only printf knows which: those, and anything out
// of range, take the snprintf path.
inline char *put_ratio2(char *p, int64_t N, int64_t D)
{
    if (N >= 0 && D > 0 && N <= 10000000000000ll) {
        const uint64_t X = (uint64_t)N * 100u, q = X / (uint64_t)D, rem = X % (uint64_t)D;
        if (2 * rem != (uint64_t)D) {
            const uint64_t r = q + (2 * rem > (uint64_t)D ? 1u : 0u);
            p = put_u64(p, r / 100);
            const unsigned f = (unsigned)(r % 100);
            *p++ = '.'; *p++ = (char)('0' + f / 10); *p++ = (char)('0' + f % 10);
            return p;
        }
    }
    return p + snprintf(p, 48, "%.2f", (double)N / (double)D);
}

// Rows k in [a, b) of the mode-6 window report of one contig (PanDepth.cpp:4366-4389): window k covers 1-based
// [1 + k*W, min(k*W + W, CC)];  chr, Start, End, Length, CoveredSite, TotalDepth as ints (:4364),
// Coverage = GeneCover*100.0/GeneLength and MeanDepth = GeneDepth*1.0/GeneLength with fixed/setprecision(2).
inline std::string format_window_rows(const char *chr, int64_t W, int64_t CC, const uint32_t *cov, const uint64_t *sum, int64_t a, int64_t b)
{
    const size_t nml = strlen(chr);
    std::string text;
    text.resize((size_t)(b - a) * (nml + 160));
    char *p = &text[0];
    for (int64_t k = a; k < b; k++) {
        const int64_t j = 1 + k * W;
        const int64_t End = j - 1 + W < CC ? j - 1 + W : CC;
        const int GeneLength = (int)(End - j + 1), GeneCover = (int)cov[(size_t)k], GeneDepth = (int)sum[(size_t)k];
        memcpy(p, chr, nml); p += nml;
        *p++ = '\t'; p = put_i64(p, j);
        *p++ = '\t'; p = put_i64(p, End);
        *p++ = '\t'; p = put_i64(p, GeneLength);
        *p++ = '\t'; p = put_i64(p, GeneCover);
        *p++ = '\t'; p = put_i64(p, GeneDepth);
        *p++ = '\t'; p = put_ratio2(p, (int64_t)GeneCover * 100, GeneLength);
        *p++ = '\t'; p = put_ratio2(p, (int64_t)GeneDepth, GeneLength);
        *p++ = '\n';
    }
    text.resize((size_t)(p - &text[0]));
    return text;
}

// Rows k in [a, b) of the mode-5 report of one contig (-w >= 150 pseudo-genes, :5080-5118): tile k is 1-based closed
// [1 + k*W, min(k*W + W, len)]; GeneCover is an int, Length and TotalDepth are 64-bit, Coverage = GeneCover*100.0/Length,
// MeanDepth = GeneDepth*1.0/Length.  cov[k] / sum[k] belong to tile k of this contig.
inline std::string format_tile_rows(const std::string &chr, int64_t W, int64_t len, const uint64_t *cov, const uint64_t *sum, int64_t a, int64_t b)
{
    std::string text;
    text.resize((size_t)(b - a) * (chr.size() + 176));
    char *p = &text[0];
    for (int64_t k = a; k < b; k++) {
        const int64_t Start = 1 + k * W;
        const int64_t End = Start + W - 1 < len ? Start + W - 1 : len;
        const int64_t Length = End - Start + 1;
        const int GeneCover = (int)(int64_t)cov[(size_t)k];
        memcpy(p, chr.data(), chr.size()); p += chr.size();
        *p++ = '\t'; p = put_i64(p, Start);
        *p++ = '\t'; p = put_i64(p, End);
        *p++ = '\t'; p = put_u64(p, (uint64_t)Length);
        *p++ = '\t'; p = put_i64(p, GeneCover);
        *p++ = '\t'; p = put_u64(p, sum[(size_t)k]);
        *p++ = '\t'; p = put_ratio2(p, (int64_t)GeneCover * 100, Length);
        *p++ = '\t';
        if (sum[(size_t)k] <= (uint64_t)INT64_MAX) p = put_ratio2(p, (int64_t)sum[(size_t)k], Length);
        else p += snprintf(p, 48, "%.2f", sum[(size_t)k] * 1.0 / (uint64_t)Length);
        *p++ = '\n';
    }
    text.resize((size_t)(p - &text[0]));
    return text;
}

// Rows of the -a per-site dump (:4278-4281): chr \t 0-based site \t depth (the 18-bit SiteInfo field) for sites
// first + [a, b) whose depths are d[a..b).
inline std::string format_site_rows(const std::string &chr, int64_t first, const uint32_t *d, int32_t a, int32_t b)
{
    std::string text;
    text.resize((size_t)(b - a) * (chr.size() + 32));
    char *p = &text[0];
    for (int32_t j = a; j < b; j++) {
        memcpy(p, chr.data(), chr.size()); p += chr.size();
        *p++ = '\t'; p = put_u64(p, (uint64_t)(first + j));
        *p++ = '\t'; p = put_u64(p, d[(size_t)j] & 0x3FFFFu);
        *p++ = '\n';
    }
    text.resize((size_t)(p - &text[0]));
    return text;
}

}  // namespace pdh
