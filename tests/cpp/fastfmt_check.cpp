// CPU checker for pandepth_b200/host/fastfmt.hpp: put_ratio2 / put_u64 / put_i64 against printf on random, adversarial
// (exact ties, near ties, range limits) and exhaustive small inputs.  Exit status 0 = identical everywhere.
#include "fastfmt.hpp"
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
using namespace pdh;
static long long bad = 0, checked = 0;
static void check(int64_t N, int64_t D)
{
    char a[96], b[96];
    *put_ratio2(a, N, D) = 0;
    snprintf(b, sizeof b, "%.2f", (double)N / (double)D);
    checked++;
    if (strcmp(a, b)) { if (bad++ < 20) fprintf(stderr, "MISMATCH N=%lld D=%lld ours=%s printf=%s\n", (long long)N, (long long)D, a, b); }
}
// the snprintf formulation the row formatters replace (the reference's ostream output, PanDepth.cpp:4382-4389 / :4278-4281)
static std::string window_rows_printf(const char *nm, int64_t W, int64_t CC, const uint32_t *cov, const uint64_t *sum, int64_t a, int64_t b)
{
    std::string text; char r[256];
    for (int64_t k = a; k < b; k++) {
        const int64_t j = 1 + k * W;
        const int64_t End = j - 1 + W < CC ? j - 1 + W : CC;
        const int GeneLength = (int)(End - j + 1), GeneCover = (int)cov[(size_t)k], GeneDepth = (int)sum[(size_t)k];
        const double Coverage = GeneCover * 100.0 / GeneLength, MeanDepth = GeneDepth * 1.0 / GeneLength;
        int m = snprintf(r, sizeof r, "%s\t%lld\t%lld\t%d\t%d\t%d\t%.2f\t%.2f\n", nm, (long long)j, (long long)End, GeneLength, GeneCover, GeneDepth, Coverage, MeanDepth);
        text.append(r, (size_t)m);
    }
    return text;
}
static std::string site_rows_printf(const std::string &nm, int64_t first, const uint32_t *d, int32_t a, int32_t b)
{
    std::string text; char line[64];
    for (int32_t j = a; j < b; j++) { int m = snprintf(line, sizeof line, "\t%lld\t%u\n", (long long)(first + j), d[(size_t)j] & 0x3FFFFu); text += nm; text.append(line, (size_t)m); }
    return text;
}

int main(int argc, char **argv)
{
    const int scale = argc > 1 ? atoi(argv[1]) : 10;      // 10 = full (about 15 s), 1 = quick
    {   // whole rows: every window size of mode 6 (W < 150) and a few larger ones, 18-bit depths incl. wrap-sized sums, last short window
        std::mt19937_64 rg(7);
        for (int64_t W : {1, 2, 3, 7, 33, 99, 100, 101, 128, 149, 150, 1000}) {
            const int64_t CC = 1 + W * 3000 + (int64_t)(rg() % (uint64_t)W), n_win = (CC - 1 + W - 1) / W;
            std::vector<uint32_t> cov((size_t)n_win); std::vector<uint64_t> sum((size_t)n_win);
            for (int64_t k = 0; k < n_win; k++) {
                const int64_t len = (k * W + W < CC ? W : CC - k * W);
                cov[(size_t)k] = (uint32_t)(rg() % (uint64_t)(len + 1));
                const uint64_t per = (k % 5 == 0) ? 262143u : (uint64_t)(rg() % 200);
                sum[(size_t)k] = (k % 11 == 0) ? 0u : per * cov[(size_t)k] + (rg() % 3);
                if (k % 97 == 0) sum[(size_t)k] = 0x80000000ull + (rg() % 1000);       // (int) cast goes negative, like the reference's int
            }
            const std::string a = format_window_rows("chr_x.1", W, CC, cov.data(), sum.data(), 0, n_win), b = window_rows_printf("chr_x.1", W, CC, cov.data(), sum.data(), 0, n_win);
            checked += n_win;
            if (a != b) { bad++; fprintf(stderr, "window rows differ for W=%lld\n", (long long)W); }
        }
        std::vector<uint32_t> d(300000);
        for (auto &x : d) x = (uint32_t)rg();
        for (int64_t first : {0ll, 9ll, 99999990ll, 2147000000ll}) {
            const std::string a = format_site_rows("c", first, d.data(), 5, 299999), b = site_rows_printf("c", first, d.data(), 5, 299999);
            checked += 299994;
            if (a != b) { bad++; fprintf(stderr, "site rows differ for first=%lld\n", (long long)first); }
        }
    }
    // the expressions of the report: Coverage = cov*100.0/len (N = cov*100), MeanDepth = depth*1.0/len (N = depth)
    for (int64_t D = 1; D <= 40 * scale; D++) for (int64_t N = 0; N <= 40000; N++) check(N, D);                 // exhaustive small (mode-6 sized windows)
    for (int64_t D = 1; D <= 150; D++) for (int64_t c = 0; c <= D; c++) check(c * 100, D);               // every Coverage of a window
    std::mt19937_64 rng(12345);
    for (int i = 0; i < 300000 * scale; i++) {
        const int64_t D = 1 + (int64_t)(rng() % (i % 3 == 0 ? 150 : (i % 3 == 1 ? 10000000 : 2147483647)));
        const int64_t N = (int64_t)(rng() % (i % 2 ? 262144ull * 150 : 10000000000000ull));
        check(N, D);
        // exact ties and their neighbours: N/D = (2k+1)/200
        const int64_t k = (int64_t)(rng() % 100000);
        const int64_t D2 = 200 * (1 + (int64_t)(rng() % 5000));
        const int64_t N2 = (2 * k + 1) * (D2 / 200);
        check(N2, D2); check(N2 + 1, D2); if (N2 > 0) check(N2 - 1, D2);
        check(N2 * 7, D2 * 7);
    }
    for (int64_t N : {0ll, 1ll, 9999999999999ll, 10000000000000ll, 10000000000001ll, 4611686018427387904ll, -5ll}) for (int64_t D : {1ll, 3ll, 7ll, 200ll, 2147483647ll}) check(N, D);
    for (uint64_t v : {0ull, 9ull, 10ull, 99ull, 100ull, 4294967295ull, 18446744073709551615ull, 1000000000000ull}) {
        char a[32], b[32]; *put_u64(a, v) = 0; snprintf(b, sizeof b, "%llu", (unsigned long long)v); checked++; if (strcmp(a, b)) bad++;
    }
    for (int64_t v : {0ll, -1ll, 1ll, -2147483648ll, 2147483647ll, (long long)INT64_MIN, (long long)INT64_MAX}) {
        char a[32], b[32]; *put_i64(a, v) = 0; snprintf(b, sizeof b, "%lld", (long long)v); checked++; if (strcmp(a, b)) bad++;
    }
    printf("checked %lld, mismatches %lld\n", checked, bad);
    return bad ? 1 : 0;
}
