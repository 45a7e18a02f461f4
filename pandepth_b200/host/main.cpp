// main.cpp -- `pandepth` on a B200: the reference's command line and output files, with the depth path on the GPU.
//
//   host threads (htslib decode)  --(tid, pos, cigar batches, pinned memory)-->  pd_push_contig   [role of ProDealChrBambai*]
//   pd_finalize / pd_reduce / pd_reduce_windows / pd_depth_read                                   [role of StatChrDepth*, mode-6 loop, -a]
//   report writer (same doubles, same text)                                                       [PanDepth.cpp:4879-5127]
//
// Decode parallelism is by genomic range through the BAI/CSI index (any number of threads on one contig), not by contig
// as in the reference (:4415-4523).
#include "pd_host.hpp"
#include "fastbam.hpp"
#include "fastfmt.hpp"
#include "feeder.hpp"
#include "gzwriter.hpp"
#include "../../include/pandepth_b200.h"

#include <zlib.h>
#include "sam.h"
#include "hts.h"

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <future>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>

using namespace pdh;

namespace {

bool file_exists(const std::string &p) { return access(p.c_str(), 0) == 0; }

// CPUs this process may actually use: the cgroup v2 / v1 CPU quota if there is one, else the online CPU count.  Decode
// workers beyond that only get throttled (possibly while holding the engine lock), so -t is clamped to it.
int usable_cpus()
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    double quota = 0;
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char a[64]; long long period = 0;
        if (fscanf(f, "%63s %lld", a, &period) == 2 && strcmp(a, "max") != 0 && period > 0) quota = atof(a) / (double)period;
        fclose(f);
    } else {
        long long q = -1, p = 0;
        if (FILE *fq = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (fscanf(fq, "%lld", &q) != 1) q = -1; fclose(fq); }
        if (FILE *fp = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(fp, "%lld", &p) != 1) p = 0; fclose(fp); }
        if (q > 0 && p > 0) quota = (double)q / (double)p;
    }
    if (quota >= 1.0 && quota < (double)n) n = (long)(quota + 0.5);
    return (int)n;
}

// PANDEPTH_TIMING=1: phase wall times on stderr
struct Timer {
    bool on = getenv("PANDEPTH_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    void mark(const char *what)
    {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[timing] %-22s %8.3f s  (total %.3f s)\n", what, std::chrono::duration<double>(now - last).count(),
                std::chrono::duration<double>(now - t0).count());
        last = now;
    }
};

using RegionsByContig = std::map<int, std::vector<std::pair<int, int>>>;      // per tid, 1-based closed, sorted (RegionMerger, :3912-3972)

// What the input says about itself before any read is decoded.
struct InputInfo {
    std::vector<Target> targets;
    std::map<std::string, int> paf_chr2int;        // Chr2IntMap of paf_main (:867)
    bool is_paf = false, is_bam = false, is_cram = false, sorted_hdr = false;
    sam_hdr_t *hdr = nullptr;
};

// `@HD ... SO:coordinate` (:4537-4549)
bool header_says_sorted(sam_hdr_t *h)
{
    const char *txt = h ? sam_hdr_str(h) : nullptr;
    const std::string text = txt ? txt : "";
    size_t a = text.find("\tSO:");
    if (a == std::string::npos) return false;
    a += 4;
    const size_t b = text.find_first_of("\n\t", a);
    return text.substr(a, b - a) == "coordinate";
}

// targets of a PAF: column 6/7 of the first file in first-seen order (:918-941).  With -r the reference takes them from the
// FASTA and switches to its GC report (RefIn, :878-909 / :1488) -- GC content is outside the depth path.
bool read_paf_targets(const Options &o, InputInfo &in)
{
    if (!o.reference.empty()) { std::cerr << "Error: -r with PAF input selects the GC report, which pandepth-b200 does not provide" << std::endl; return false; }
    LineReader in0(o.input);
    if (!in0.ok()) { std::cerr << "open PAF file error: " << o.input << std::endl; return false; }
    std::string line;
    while (in0.next(line)) {
        if (line.empty()) continue;
        size_t k = 0, p = 0; std::string chr, len;
        while (k < 7) {                                                  // isone>>tmp1..tmp5>>chr>>chrlength
            p = line.find_first_not_of(" \t\r\n\v\f", p);
            if (p == std::string::npos) break;
            const size_t q = line.find_first_of(" \t\r\n\v\f", p);
            if (k == 5) chr = line.substr(p, q == std::string::npos ? std::string::npos : q - p);
            if (k == 6) len = line.substr(p, q == std::string::npos ? std::string::npos : q - p);
            k++;
            if (q == std::string::npos) break;
            p = q;
        }
        if (k < 7) continue;
        if (!in.paf_chr2int.count(chr)) { in.paf_chr2int[chr] = (int)in.targets.size(); in.targets.push_back(Target{chr, (uint32_t)atoi(len.c_str())}); }
    }
    return true;
}

// targets of a SAM / BAM / CRAM: the header (:3482-3507)
bool read_hts_targets(const Options &o, InputInfo &in)
{
    htsFile *fp = hts_open(o.input.c_str(), "r");
    hts_set_log_level(HTS_LOG_OFF);
    if (!fp) { std::cerr << "Error: Failed to open the BAM/CRAM file: " << o.input << std::endl; return false; }
    if (fp->format.format == cram) hts_set_fai_filename(fp, o.reference.c_str());
    in.is_bam = fp->format.format == bam; in.is_cram = fp->format.format == cram;
    in.hdr = sam_hdr_read(fp);
    if (!in.hdr) { std::cerr << "Error: Failed to read the header for the BAM/CRAM file: " << o.input << std::endl; return false; }
    in.targets.resize((size_t)in.hdr->n_targets);
    for (int i = 0; i < in.hdr->n_targets; i++) { in.targets[(size_t)i].name = in.hdr->target_name[i]; in.targets[(size_t)i].len = in.hdr->target_len[i]; }
    in.sorted_hdr = header_says_sorted(in.hdr);
    hts_close(fp);
    return true;
}

bool has_index(const std::string &path) { return file_exists(path + ".bai") || file_exists(path + ".crai") || file_exists(path + ".csi"); }   // :4121-4124

// ---------------------------------------------------------------------------------------------------------------
// diagnostics for the CPU test-suite (PANDEPTH_DRY=1: no GPU is touched)
// ---------------------------------------------------------------------------------------------------------------
// PANDEPTH_DUMP_REGIONS=<file>: the region model as text
void dump_regions(const char *path, const Options &o, const std::vector<Target> &targets, const GeneMap &genes, const RegionsByContig &merged,
                  int mode, uint32_t mask)
{
    FILE *f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "mode %d mask %x\n", mode, (unsigned)mask);
    for (const auto &kv : genes) {
        const long long len = targets[(size_t)kv.first].len;
        if (mode == 5) {
            // implicit -w tiles (pd_host.hpp n_tiles), in the std::map<string> order of their ids; every tile is its own
            // merged region (adjacent tiles do not merge, :3919-3972) and is handed over exactly once
            const long long W = o.tile_win, nt = n_tiles(len, W);
            std::vector<std::pair<std::string, long long>> ids;
            for (long long k = 0; k < nt; k++) ids.emplace_back(targets[(size_t)kv.first].name + std::to_string(1 + k * W), k);
            std::sort(ids.begin(), ids.end());
            for (const auto &id : ids) {
                const long long S = 1 + id.second * W, E = std::min(S + W - 1, len);
                fprintf(f, "G\t%d\t%s\t%lld\t%lld\t%lld\t%lld,%lld\n", kv.first, id.first.c_str(), S, E, E - S + 1, S, E);
            }
            fprintf(f, "M\t%d", kv.first);
            for (long long k = 0; k < nt; k++) fprintf(f, "\t%lld,%lld", 1 + k * W, std::min(k * W + W, len));
            fputc('\n', f);
            for (const auto &id : ids) fprintf(f, "W\t%d\t%s\t1\n", kv.first, id.first.c_str());
            continue;
        }
        for (const auto &gk : kv.second) {
            const Gene &g = gk.second;
            fprintf(f, "G\t%d\t%s\t%d\t%d\t%llu", kv.first, gk.first.c_str(), g.start, g.end, (unsigned long long)g.length);
            for (const auto &c : g.cds) fprintf(f, "\t%d,%d", c.first, c.second);
            fputc('\n', f);
        }
        const auto &regs = merged.at(kv.first);
        fprintf(f, "M\t%d", kv.first);
        for (const auto &r : regs) fprintf(f, "\t%d,%d", r.first, r.second);
        fputc('\n', f);
        for (const auto &wk : indexed_window_counts(regs, kv.second, (int)len)) fprintf(f, "W\t%d\t%s\t%d\n", kv.first, wk.first.c_str(), wk.second);
    }
    fclose(f);
}

// PANDEPTH_SELFTEST_ROWS=<prefix>: the report writers (row formatters + parallel gzip members) on a fixed synthetic table
// -> <prefix>.win.gz, <prefix>.site.gz, <prefix>.empty.gz
void selftest_rows(const char *pre, int threads)
{
    const int64_t W = 100, CC = 10000003, n_win = (CC - 1 + W - 1) / W;
    std::vector<uint32_t> cov((size_t)n_win); std::vector<uint64_t> sum((size_t)n_win);
    for (int64_t k = 0; k < n_win; k++) {
        const int64_t len = std::min(W, CC - k * W);
        cov[(size_t)k] = (uint32_t)((k * 7919) % (len + 1));
        sum[(size_t)k] = (uint64_t)cov[(size_t)k] * (uint64_t)((k * 31) % 500);
    }
    GzWriter w;
    if (w.open(std::string(pre) + ".win.gz", threads)) {
        w.put(stat_header(6));
        const int64_t SUB = 1 << 14;
        w.put_chunks((size_t)((n_win + SUB - 1) / SUB), [&](size_t c) { const int64_t a = (int64_t)c * SUB, b = std::min(n_win, a + SUB); return format_window_rows("selftest", W, CC, cov.data(), sum.data(), a, b); });
        w.close();
    }
    const int32_t n = 600000, SUBS = 1 << 16;
    std::vector<uint32_t> d((size_t)n);
    for (int32_t j = 0; j < n; j++) d[(size_t)j] = (uint32_t)((uint64_t)j * 2654435761ull);
    GzWriter sw;
    if (sw.open(std::string(pre) + ".site.gz", threads)) {
        sw.put_chunks((size_t)((n + SUBS - 1) / SUBS), [&](size_t k) { const int32_t a = (int32_t)k * SUBS, b = std::min<int32_t>(n, a + SUBS); return format_site_rows("selftest", 5, d.data(), a, b); });
        sw.close();
    }
    GzWriter ew;                                          // a file nothing was written to is still a valid (empty) .gz
    if (ew.open(std::string(pre) + ".empty.gz", 1)) ew.close();
}

// ---------------------------------------------------------------------------------------------------------------
// feed
// ---------------------------------------------------------------------------------------------------------------
// Work items of the indexed path: genomic ranges (modes 0/5/6) or clusters of merged regions (modes 1-4, fetching exactly
// the reads the reference's `chr:first-1..second+1` queries fetch, :700-706).
void plan_items(Feeder &feeder, const RegionsByContig &merged, const std::vector<uint32_t> &tlen, bool region_modes, int threads)
{
    int64_t total_len = 0;
    for (const auto &kv : merged) total_len += tlen[(size_t)kv.first];
    // bases per work item: ~6 items per thread, but never so small that BGZF blocks are decoded twice at every boundary
    const int64_t RANGE = std::max<int64_t>(1 << 20, std::min<int64_t>(8 << 20, total_len / (6 * (int64_t)threads) + 1));
    const int64_t GAP = 64 << 10;                // merged regions closer than this share one index query
    for (const auto &kv : merged) {
        const int tid = kv.first;
        const int64_t len = tlen[(size_t)tid];
        int64_t prev_end = 0;
        if (!region_modes) {
            for (int64_t b = 0; b < len; b += RANGE) { feeder.items.push_back({tid, b, std::min(len, b + RANGE), prev_end, 0, 0}); prev_end = std::min(len, b + RANGE); }
            continue;
        }
        const auto &regs = kv.second;
        size_t a = 0;
        while (a < regs.size()) {
            size_t b = a + 1;
            int64_t beg = std::max<int64_t>((int64_t)regs[a].first - 1, 1) - 1;     // chr:first-1 .. second+1 (:700-706)
            int64_t end = std::min<int64_t>((int64_t)regs[a].second + 1, len);
            while (b < regs.size() && (int64_t)regs[b].first - 2 - end < GAP && end - beg < 2 * RANGE) { end = std::min<int64_t>((int64_t)regs[b].second + 1, len); b++; }
            if (end > beg) feeder.items.push_back({tid, beg, end, prev_end, (int)a, (int)b});
            prev_end = std::max(prev_end, end);
            a = b;
        }
    }
}

// One SAM / BAM / CRAM file into the engine: range-parallel through the index when there is one, else one pass.
void feed_hts_file(Feeder &feeder, const Options &o, const std::string &path, bool f_bam, bool f_cram, bool f_sorted, bool region_modes,
                   const RegionsByContig &merged, const std::vector<uint32_t> &tlen, size_t n_targets, bool first_file)
{
    const bool f_indexed = has_index(path) && o.use_index;
    // device-side decode of the whole file (PANDEPTH_GPU_DECODE=0 turns it off): whole-genome modes -- with -g / -b the index
    // lets the host skip most of the file --, first input file only (a fallback restarts from a clean engine), and not where
    // the reference's no-index region cursor decides which reads count (sorted header without an index)
    const char *gd = getenv("PANDEPTH_GPU_DECODE");
    if (!(gd && gd[0] == '0') && f_bam && !region_modes && first_file && (f_indexed || !f_sorted)) {
        const int r = feeder.stream_bgzf(f_indexed);         // no index: a cut file ends quietly, like the reference's sam_read1 loop
        if (r != 0) return;
    }
    feeder.filter_by_region = f_indexed && region_modes;
    if (!f_indexed) {
        if (f_sorted) std::cout << "Warning: PanDepth will run in No Index mode: " << path << std::endl;                       // :4606
        else std::cout << "Warning: Can't find index file of input BAM/CRAM. PanDepth will run in No Index mode: " << path << std::endl;   // :4677
        // BAM: own parallel stream reader; SAM / CRAM / PANDEPTH_HTS_ITER=1: htslib's sam_read1
        if (!(f_bam && !getenv("PANDEPTH_HTS_ITER") && feeder.stream_all_fast(f_sorted, n_targets))) feeder.stream_all(f_sorted, n_targets);
        return;
    }
    if (feeder.items.empty()) plan_items(feeder, merged, tlen, region_modes, o.threads);
    feeder.next = 0;
    // sorted BAM + .bai: raw-record scanner; anything else (CRAM, .csi only, PANDEPTH_HTS_ITER=1): htslib iterators
    BaiLinear bai;
    feeder.bai = nullptr; feeder.shared_idx = nullptr;
    if (f_bam && !getenv("PANDEPTH_HTS_ITER") && file_exists(path + ".bai") && bai.load(path + ".bai") && bai.refs.size() == n_targets)
        feeder.bai = &bai;
    if (!feeder.bai && !f_cram) {
        htsFile *ifp = hts_open(path.c_str(), "r");
        if (ifp) { feeder.shared_idx = sam_index_load(ifp, path.c_str()); hts_close(ifp); }
    }
    const int nthreads = (int)std::max<size_t>(1, std::min<size_t>((size_t)o.threads, feeder.items.size()));
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; i++) th.emplace_back([&] { if (feeder.bai) feeder.fast_worker(); else feeder.indexed_worker(); });
    for (auto &t : th) t.join();
    if (feeder.shared_idx) hts_idx_destroy(feeder.shared_idx);
    feeder.bai = nullptr; feeder.shared_idx = nullptr;
}

// ---------------------------------------------------------------------------------------------------------------
// report
// ---------------------------------------------------------------------------------------------------------------
struct Totals { uint64_t len = 0, cov = 0, depth = 0; };     // SS_Len, SS_Cov, SS_TotalD (:5119-5122)

bool engine_error(pd_engine *eng) { std::cerr << "Error: " << pd_last_error(eng) << std::endl; return false; }

// One engine, or several that own disjoint genomic ranges (PANDEPTH_DEVICES, range sharding): interval results add up.
struct Engines {
    std::vector<pd_engine *> v;
    pd_engine *first() const { return v[0]; }
    pd_engine *failed = nullptr;
    bool reduce(int64_t n_iv, const int32_t *tid, const int32_t *s0, const int32_t *e0, uint32_t min_dep, uint32_t mask, uint64_t *cov, uint64_t *sum)
    {
        if (v.size() == 1) { if (pd_reduce(v[0], n_iv, tid, s0, e0, min_dep, mask, cov, sum, 0) != PD_OK) { failed = v[0]; return false; } return true; }
        std::vector<uint64_t> c((size_t)n_iv), d((size_t)n_iv);
        std::fill(cov, cov + n_iv, 0); std::fill(sum, sum + n_iv, 0);
        for (pd_engine *e : v) {
            if (pd_reduce(e, n_iv, tid, s0, e0, min_dep, mask, c.data(), d.data(), 0) != PD_OK) { failed = e; return false; }
            for (int64_t i = 0; i < n_iv; i++) { cov[i] += c[(size_t)i]; sum[i] += d[(size_t)i]; }
        }
        return true;
    }
};

// -a per-site dump (:4264-4284): D2H in 4 M-base pieces, rows formatted + compressed by the thread pool
bool write_site_depth(pd_engine *eng, const std::string &path, const std::vector<Target> &targets, const GeneMap &genes, int threads)
{
    GzWriter site;
    if (!site.open(path, threads)) return true;           // like the reference: no dump, no error
    const int32_t CH = 4 << 20, SUB = 1 << 18;
    std::vector<uint32_t> d((size_t)CH);
    for (size_t t = 0; t < targets.size(); t++) {
        if (!genes.count((int)t)) continue;
        const std::string &nm = targets[t].name;
        for (int64_t s = 0; s < (int64_t)targets[t].len; s += CH) {
            const int32_t n = (int32_t)std::min<int64_t>(CH, (int64_t)targets[t].len - s);
            if (pd_depth_read(eng, (int32_t)t, (int32_t)s, n, d.data(), 0) != PD_OK) return engine_error(eng);
            site.put_chunks((size_t)((n + SUB - 1) / SUB), [&](size_t k) {
                const int32_t a = (int32_t)k * SUB, b = std::min<int32_t>(n, a + SUB);
                return format_site_rows(nm, s, d.data(), a, b);
            });
        }
    }
    return site.close();
}

// mode 6 (-w < 150): the fixed-stride window loop (:4353-4394), always on the 18-bit depth
bool report_windows(pd_engine *eng, GzWriter &out, const Options &o, const std::vector<Target> &targets, const GeneMap &genes, Totals &tot)
{
    std::vector<uint32_t> cov; std::vector<uint64_t> sum;
    for (size_t t = 0; t < targets.size(); t++) {
        if (!genes.count((int)t)) continue;
        const int64_t CC = targets[t].len, W = o.win;
        const int64_t n_win = CC > 1 ? (CC - 1 + W - 1) / W : 0;                         // for (j=1; j<CC; j+=W)
        cov.assign((size_t)n_win, 0); sum.assign((size_t)n_win, 0);
        if (n_win && pd_reduce_windows(eng, (int32_t)t, 0, (int32_t)W, n_win, (uint32_t)o.min_dep, 0x3FFFFu, cov.data(), sum.data(), 0) != PD_OK) return engine_error(eng);
        const char *nm = targets[t].name.c_str();
        const int64_t SUB = 1 << 16;
        out.put_chunks((size_t)((n_win + SUB - 1) / SUB), [&](size_t c) {
            const int64_t a = (int64_t)c * SUB, b = std::min(n_win, a + SUB);
            return format_window_rows(nm, W, CC, cov.data(), sum.data(), a, b);
        });
        for (int64_t k = 0; k < n_win; k++) {
            const int64_t j = 1 + k * W;
            const int64_t End = std::min(j - 1 + W, CC);
            tot.cov += (uint64_t)(int64_t)(int)cov[(size_t)k]; tot.len += (uint64_t)(int64_t)(int)(End - j + 1); tot.depth += (uint64_t)(int64_t)(int)sum[(size_t)k];
        }
    }
    return true;
}

// mode 5 (-w >= 150): the pseudo-gene tiles are enumerated arithmetically (pd_host.hpp n_tiles) -- one pd_reduce over all of
// them, rows formatted and compressed in parallel.  Same rows, same order (GeneStart) and same totals as report_genes would
// produce from materialised genes (:5080-5118); every tile counts exactly once (pd_host.hpp).
bool report_tiles(Engines &engs, GzWriter &out, const Options &o, const std::vector<Target> &targets, const GeneMap &genes, uint32_t mask, Totals &tot)
{
    const long long W = o.tile_win;
    std::vector<int> tids;
    std::vector<size_t> base;                                                            // first interval of each contig
    size_t n_iv = 0;
    for (const auto &kv : genes) { tids.push_back(kv.first); base.push_back(n_iv); n_iv += (size_t)n_tiles(targets[(size_t)kv.first].len, W); }
    std::vector<int32_t> iv_t(n_iv), iv_s(n_iv), iv_e(n_iv);
    for (size_t ci = 0; ci < tids.size(); ci++) {
        const long long len = targets[(size_t)tids[ci]].len, nt = n_tiles(len, W);
        for (long long k = 0; k < nt; k++) {
            iv_t[base[ci] + (size_t)k] = tids[ci];
            iv_s[base[ci] + (size_t)k] = (int32_t)(k * W);
            iv_e[base[ci] + (size_t)k] = (int32_t)std::min(k * W + W, len);
        }
    }
    std::vector<uint64_t> cov(n_iv), sum(n_iv);
    if (!engs.reduce((int64_t)n_iv, iv_t.data(), iv_s.data(), iv_e.data(), (uint32_t)o.min_dep, mask, cov.data(), sum.data())) return engine_error(engs.failed);
    for (size_t ci = 0; ci < tids.size(); ci++) {
        const long long len = targets[(size_t)tids[ci]].len, nt = n_tiles(len, W);
        const std::string &chr = targets[(size_t)tids[ci]].name;
        const uint64_t *c = cov.data() + base[ci], *d = sum.data() + base[ci];
        const long long SUB = 1 << 16;
        out.put_chunks((size_t)((nt + SUB - 1) / SUB), [&](size_t ch) {
            const long long a = (long long)ch * SUB, b = std::min(nt, a + SUB);
            return format_tile_rows(chr, W, len, c, d, a, b);
        });
        for (long long k = 0; k < nt; k++) {
            tot.cov += (uint64_t)(int64_t)(int)(int64_t)c[k];
            tot.len += (uint64_t)(std::min(k * W + W, len) - k * W);
            tot.depth += d[k];
        }
    }
    return true;
}

// modes 0-4: every CDS interval of every gene in one pd_reduce call (StatChrDepthWin / StatChrDepthLowMEM), then one row per
// contig (mode 0, :5051-5079) or per gene (:5011-5050), rows ordered by (tid, GeneStart, GeneID)
bool report_genes(Engines &engs, GzWriter &out, const Options &o, const std::vector<Target> &targets, GeneMap &genes, const RegionsByContig &merged,
                  int mode, bool whole, uint32_t mask, Totals &tot)
{
    std::vector<int32_t> iv_t, iv_s, iv_e;
    std::vector<Gene *> iv_g;
    for (auto &kv : genes) {
        std::map<std::string, int> times;
        if (!whole) times = indexed_window_counts(merged.at(kv.first), kv.second, (int)targets[(size_t)kv.first].len);
        for (auto &gk : kv.second) {
            if (!whole && times[gk.first] == 0) continue;                                 // never handed to StatChrDepthWin (:299-308)
            for (const auto &c : gk.second.cds) { iv_t.push_back(kv.first); iv_s.push_back(c.first - 1); iv_e.push_back(c.second); iv_g.push_back(&gk.second); }
        }
    }
    std::vector<uint64_t> cov(iv_t.size()), sum(iv_t.size());
    if (!engs.reduce((int64_t)iv_t.size(), iv_t.data(), iv_s.data(), iv_e.data(), (uint32_t)o.min_dep, mask, cov.data(), sum.data())) return engine_error(engs.failed);
    for (size_t i = 0; i < iv_g.size(); i++) { iv_g[i]->cover += (int64_t)cov[i]; iv_g[i]->depth += sum[i]; }

    std::vector<int> tids;
    for (auto &kv : genes) tids.push_back(kv.first);
    std::vector<uint64_t> pL(tids.size(), 0), pC(tids.size(), 0), pD(tids.size(), 0);
    out.put_chunks(tids.size(), [&](size_t ci) {                                         // one chunk per contig, rows in order
        auto &gm = genes[tids[ci]];
        const std::string &chr = targets[(size_t)tids[ci]].name;
        std::string text;
        char num[256];
        if (mode == 0) {                                                                   // :5051-5079
            uint64_t L = 0, C = 0, D = 0;
            for (auto &gk : gm) { L += gk.second.length; C += (uint64_t)(int64_t)(int)gk.second.cover; D += gk.second.depth; }
            pL[ci] = L; pC[ci] = C; pD[ci] = D;
            int m = snprintf(num, sizeof num, "%s\t%llu\t%llu\t%llu\t%s\t%s\n", chr.c_str(), (unsigned long long)L, (unsigned long long)C,
                             (unsigned long long)D, fmt_ratio_pct((double)C, (double)L).c_str(), fmt_ratio((double)D, (double)L).c_str());
            text.append(num, (size_t)m);
            return text;
        }
        std::map<int, std::string> by_start;                                               // rows by GeneStart; ties in GeneID order
        for (auto &gk : gm) {
            const Gene &g = gk.second;
            const int GeneCover = (int)g.cover;
            pC[ci] += (uint64_t)(int64_t)GeneCover; pL[ci] += g.length; pD[ci] += g.depth;
            const double Coverage = GeneCover * 100.0 / g.length, MeanDepth = g.depth * 1.0 / g.length;
            std::string line = chr + "\t" + std::to_string(g.start) + "\t" + std::to_string(g.end) + "\t" + gk.first + "\t";
            snprintf(num, sizeof num, "%llu\t%d\t%llu\t%.2f\t%.2f", (unsigned long long)g.length, GeneCover, (unsigned long long)g.depth, Coverage, MeanDepth);
            line += num;
            auto it = by_start.find(g.start);
            if (it == by_start.end()) by_start.emplace(g.start, line);
            else { it->second += "\n"; it->second += line; }
        }
        for (auto &r : by_start) { text += r.second; text += "\n"; }
        return text;
    });
    for (size_t ci = 0; ci < tids.size(); ci++) { tot.len += pL[ci]; tot.cov += pC[ci]; tot.depth += pD[ci]; }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// several GPUs in one process (PANDEPTH_DEVICES=0,1,...): range sharding
// ---------------------------------------------------------------------------------------------------------------
// The concatenated genome (the contigs that carry regions) is cut into equal ranges, one engine per device owns one range
// (pd_create_ranges) and decodes, through the .bai, the part of the file that can hold reads overlapping it: from the
// first read overlapping its first base up to the first block that starts behind its last base (pd_bgzf_progress).
// Boundary blocks go to both neighbours; the engines clip.  Returns the engines (empty: not handled, e.g. a file the device
// decoder cannot take -- the caller continues on one GPU).
std::vector<pd_engine *> feed_multi_gpu(const Options &o, const std::string &path, const std::vector<int> &devices, const std::vector<uint32_t> &tlen,
                                        const std::vector<uint8_t> &owned, uint32_t hint_min_dep, uint32_t hint_mask, bool hint, Timer &tm, std::string &err)
{
    const size_t K = devices.size(), T = tlen.size();
    BaiLinear bai;
    if (!bai.load(path + ".bai") || bai.refs.size() != T) return {};
    uint64_t G = 0;
    std::vector<uint64_t> start(T + 1, 0);
    for (size_t t = 0; t < T; t++) { start[t] = G; if (owned[t]) G += tlen[t]; }
    start[T] = G;
    std::vector<std::vector<uint32_t>> beg(K, std::vector<uint32_t>(T, 0)), end(K, std::vector<uint32_t>(T, 0));
    for (size_t k = 0; k < K; k++) {
        const uint64_t a = G * k / K, z = G * (k + 1) / K;
        for (size_t t = 0; t < T; t++) {
            if (!owned[t]) continue;
            const uint64_t lo = std::max(a, start[t]), hi = std::min(z, start[t] + tlen[t]);
            if (hi > lo) { beg[k][t] = (uint32_t)(lo - start[t]); end[k][t] = (uint32_t)(hi - start[t]); }
        }
    }
    std::vector<pd_engine *> eng(K, nullptr);
    std::vector<BgzfFeedResult> res(K);
    std::vector<ChunkPool> pools(K);
    std::vector<std::thread> th;
    const int thr = std::max(1, o.threads / (int)K);
    for (size_t k = 0; k < K; k++)
        th.emplace_back([&, k] {
            eng[k] = pd_create_ranges(devices[k], (int)T, tlen.data(), beg[k].data(), end[k].data(), o.xflags, o.min_mapq);
            if (!eng[k]) { res[k].status = -1; res[k].err = std::string("Error: ") + pd_last_error(nullptr); return; }
            if (hint) pd_set_reduce_hint(eng[k], hint_min_dep, hint_mask);
            int t0 = -1, t1 = -1;
            for (size_t t = 0; t < T; t++) if (end[k][t] > beg[k][t]) { if (t0 < 0) t0 = (int)t; t1 = (int)t; }
            if (t0 < 0) { res[k].status = 1; return; }           // more devices than bases
            uint64_t f_beg = 0;
            if (k > 0) {
                f_beg = UINT64_MAX;
                for (size_t t = (size_t)t0; t < T; t++) { const uint64_t v = bai.start_voffset((int)t, (int)t == t0 ? (int64_t)beg[k][t] : 0); if (v) { f_beg = v >> 16; break; } }
                if (f_beg == UINT64_MAX) { res[k].status = 1; return; }     // no read at or behind my first base
            }
            const bool last = k + 1 == K;
            { const char *pm = getenv("PANDEPTH_PINNED_MB"); pools[k].max_bytes = (size_t)((pm ? atof(pm) : 512.0) * 1048576.0); }
            res[k] = bgzf_feed_file(eng[k], nullptr, path, thr, pools[k], true, f_beg, UINT64_MAX, last ? INT32_MAX : t1, last ? INT32_MAX : (int32_t)end[k][(size_t)t1]);
        });
    for (auto &t : th) t.join();
    bool ok = true, fallback = false;
    for (size_t k = 0; k < K; k++) { if (res[k].status < 0) { ok = false; if (err.empty()) err = res[k].err; } if (res[k].status == 0) fallback = true; }
    if (tm.on) for (size_t k = 0; k < K; k++)
        fprintf(stderr, "[timing] device %d: %llu pieces, %.1f MB fed, pread %.3f s, pd_push_bgzf %.3f s (thread-seconds)\n", devices[k],
                (unsigned long long)res[k].pieces, res[k].bytes / 1048576.0, res[k].s_read, res[k].s_push);
    if (!ok || fallback) {
        for (pd_engine *e : eng) if (e) pd_destroy(e);
        if (ok) err.clear();                                    // fallback only: not an error, the single-GPU path takes over
        return {};
    }
    return eng;
}

}  // namespace

int main(int argc, char **argv)
{
    Options o;
    {
        const int pa = parse_args(argc, argv, o);
        if (pa == 2) return 1;                               // an option this build does not provide (-c): not a silent success
        if (pa != 0) return 0;                               // usage / argument errors leave with 0, like the reference (:3449-3450)
    }
    Timer tm;
    if (!getenv("PANDEPTH_NO_CLAMP")) {
        const int cpus = usable_cpus();
        if (o.threads > cpus) { if (tm.on) fprintf(stderr, "[timing] -t %d clamped to the %d usable CPUs (cgroup quota / online CPUs)\n", o.threads, cpus); o.threads = cpus; }
    }

    std::string ext = o.input.substr(o.input.rfind('.') == std::string::npos ? o.input.size() : o.input.rfind('.') + 1);
    if (ext == "gz") {
        std::string a = o.input.substr(0, o.input.rfind('.'));
        ext = a.substr(a.rfind('.') == std::string::npos ? a.size() : a.rfind('.') + 1);
    }
    const bool multi = o.inputs.size() > 1;                  // FileNum > 1 -> List_main / BamList_main (:3452-3458)
    if (multi) std::cout << "INFO: Run multi-file data " << std::endl;
    InputInfo in;
    in.is_paf = ext == "paf" || ext == "PAF";                // :3428 / :3474 (a trailing .gz is looked through)
    if (in.is_paf) std::cout << (multi ? "INFO: Run PAF format data " : "INFO: Run paf Format data ") << std::endl;
    if (!(in.is_paf ? read_paf_targets(o, in) : read_hts_targets(o, in))) return 1;
    const std::vector<Target> &targets = in.targets;

    GeneMap genes;
    std::string err;
    if (!build_regions(o, targets, genes, err)) { std::cerr << err << std::endl; return 1; }
    const int mode = o.region_mode;
    const bool region_modes = mode >= 1 && mode <= 4;
    RegionsByContig merged;
    for (const auto &kv : genes) merged[kv.first] = merge_regions(kv.second);

    // the reference's 18-bit SiteInfo paths: no index, -a, mode 6 (:4127, :4553), PAF (:1520) or several input files
    // (BamList_main :2687); everything else is the uint32 window path
    const bool indexed = has_index(o.input) && o.use_index;
    const bool whole = in.is_paf || multi || !indexed || o.site_output || mode == 6;
    const uint32_t mask = whole ? 0x3FFFFu : 0xFFFFFFFFu;

    if (targets.empty()) {
        // a header without @SQ lines: nothing to accumulate.  The reference still writes its files (header line, the summary
        // row with 0/0 ratios, an empty -a dump); no engine is needed for that.
        const std::string prefix0 = output_prefix(o.out_arg);
        GzWriter out0;
        const std::string stat0 = prefix0 + stat_suffix(mode);
        if (!out0.open(stat0, 1)) { std::cerr << "open OUT File error: " << stat0 << std::endl; return 0; }
        if (!in.is_paf && !(has_index(o.input) && o.use_index))
            std::cout << (in.sorted_hdr ? "Warning: PanDepth will run in No Index mode: " : "Warning: Can't find index file of input BAM/CRAM. PanDepth will run in No Index mode: ") << o.input << std::endl;
        std::cout << "INFO: Input data read done" << std::endl;
        if (o.site_output) { GzWriter s0; if (s0.open(prefix0 + ".SiteDepth.gz", 1)) s0.close(); }
        out0.put(stat_header(mode));
        char row0[256];
        volatile uint64_t z = 0;                              // 0 * 100.0 / 0 at run time, like the reference (prints -nan)
        int m0 = snprintf(row0, sizeof row0, "##RegionLength: %llu\tCoveredSite: %llu\tCoverage(%%): %.2f\tMeanDepth: %.2f\n", 0ull, 0ull, z * 100.0 / z, z * 1.0 / z);
        out0.put(std::string(row0, (size_t)m0));
        return out0.close() ? 0 : 1;
    }

    // the CUDA context (~0.5 s) is created in the background while the host threads already decode
    int device = 0;
    if (const char *d = getenv("PANDEPTH_DEVICE")) device = atoi(d);
    std::vector<int> devices;                                // PANDEPTH_DEVICES=0,1,2,...: several GPUs, range sharding
    if (const char *d = getenv("PANDEPTH_DEVICES")) { for (const char *q = d; *q;) { if (*q >= '0' && *q <= '9') { devices.push_back(atoi(q)); while (*q >= '0' && *q <= '9') q++; } else q++; } }
    if (!devices.empty()) device = devices[0];
    // cuInit brings up every VISIBLE GPU (0.5 s for one, 2+ s on a multi-GPU node): show the driver only the devices this run uses
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        std::string vis;
        if (devices.empty()) vis = std::to_string(device);
        else for (size_t k = 0; k < devices.size(); k++) vis += (k ? "," : "") + std::to_string(devices[k]);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
        if (devices.empty()) device = 0; else for (size_t k = 0; k < devices.size(); k++) devices[k] = (int)k;
    }
    std::vector<uint32_t> tlen(targets.size());
    std::vector<uint8_t> owned(targets.size(), 0);
    for (size_t t = 0; t < targets.size(); t++) { tlen[t] = targets[t].len; owned[t] = genes.count((int)t) ? 1 : 0; }
    PinnedPool pool;
    const bool dry = getenv("PANDEPTH_DRY") != nullptr;      // decode only: no GPU is touched (diagnostics)
    // -d and the depth width are known before the first read: let the engine keep per-segment partials for pd_reduce
    const bool hint_reduce = mode != 6;
    // Several GPUs: whole-genome interval modes (per-chr rows, -w >= 150 tiles) of one indexed BAM, decoded on the devices.
    // Everything else (regions, -a, -w < 150, SAM / CRAM / PAF / lists) runs on the first device.
    Engines engs;
    {
        const char *gd = getenv("PANDEPTH_GPU_DECODE");
        if (devices.size() > 1 && !dry && !multi && in.is_bam && indexed && !region_modes && (mode == 0 || mode == 5) && !o.site_output && !(gd && gd[0] == '0')) {
            std::string merr;
            engs.v = feed_multi_gpu(o, o.input, devices, tlen, owned, (uint32_t)o.min_dep, mask, hint_reduce, tm, merr);
            if (!merr.empty()) { std::cerr << merr << std::endl; return 1; }
        }
    }
    const bool multi_gpu = !engs.v.empty();
    std::shared_future<pd_engine *> eng_future = std::async(std::launch::async, [&]() -> pd_engine * {
        if (dry || multi_gpu) return nullptr;
        pd_engine *e = pd_create(device, (int)targets.size(), tlen.data(), owned.data(), o.xflags, o.min_mapq);
        if (e && hint_reduce) pd_set_reduce_hint(e, (uint32_t)o.min_dep, mask);
        if (e) pool.init(8, (1u << 20) + 1024, (1u << 20) + (1u << 18));      // 8 x 13 MB pinned; bigger batches go pageable
        return e;
    }).share();

    const std::string prefix = output_prefix(o.out_arg);
    GzWriter out;
    const std::string stat_path = prefix + stat_suffix(mode);
    if (!out.open(stat_path, o.threads)) { std::cerr << "open OUT File error: " << stat_path << std::endl; eng_future.wait(); return 0; }
    tm.mark("header+regions");
    if (const char *path = getenv("PANDEPTH_DUMP_REGIONS")) dump_regions(path, o, targets, genes, merged, mode, mask);
    if (const char *pre = getenv("PANDEPTH_SELFTEST_ROWS")) selftest_rows(pre, o.threads);

    // a failed run must not leave a truncated .gz behind at the output path
    auto drop_outputs = [&] { out.abandon(); if (o.site_output) unlink((prefix + ".SiteDepth.gz").c_str()); };

    // ---- feed ---------------------------------------------------------------------------------------------------
    Feeder feeder(o, eng_future, merged, false);
    feeder.pool = &pool;
    feeder.targets = &targets;
    feeder.hint_set = hint_reduce; feeder.hint_min_dep = (uint32_t)o.min_dep; feeder.hint_mask = mask;
    for (size_t fi = 0; fi < o.inputs.size() && !feeder.failed && !multi_gpu; fi++) {
        const std::string &path = o.inputs[fi];
        feeder.input = path;
        if (in.is_paf) { feeder.paf_stream(in.paf_chr2int, targets.size()); continue; }
        bool f_bam = in.is_bam, f_cram = in.is_cram, f_sorted = in.sorted_hdr;
        if (fi > 0) {                                        // later files: own format / sort order, the first file's targets (:2701-2862)
            htsFile *fp = hts_open(path.c_str(), "r");
            if (!fp) { feeder.fail("Error: Failed to open the BAM/CRAM file: " + path); break; }        // partial totals are not a result
            if (fp->format.format == cram) hts_set_fai_filename(fp, o.reference.c_str());
            f_bam = fp->format.format == bam; f_cram = fp->format.format == cram;
            sam_hdr_t *h2 = sam_hdr_read(fp);
            f_sorted = header_says_sorted(h2);
            if (h2) sam_hdr_destroy(h2);
            hts_close(fp);
        }
        feed_hts_file(feeder, o, path, f_bam, f_cram, f_sorted, region_modes, merged, tlen, targets.size(), fi == 0);
    }
    if (dry) {
        tm.mark("decode (dry)");
        std::cout << "dry: reads=" << feeder.reads_pushed.load() << " digest=" << feeder.dry_digest.load() << " path=" << (feeder.used_fast ? "fastbam" : "htslib") << std::endl;
        drop_outputs();                                      // a dry run writes nothing
        if (feeder.failed) { std::cerr << feeder.err << std::endl; return 1; }
        return 0;
    }
    if (!multi_gpu) feeder.finish();
    if (tm.on && !multi_gpu) fprintf(stderr, "[timing] workers: total %.3f thread-s, of which mutex wait %.3f s, inside push %.3f s (%llu pushes, %llu reads)%s\n",
                       feeder.ns_decode / 1e9, feeder.ns_wait / 1e9, feeder.ns_push / 1e9, (unsigned long long)feeder.n_push.load(), (unsigned long long)feeder.reads_pushed.load(),
                       feeder.used_device_decode ? " [device-side BGZF decode]" : "");
    pd_engine *eng = multi_gpu ? engs.v[0] : eng_future.get();
    if (!eng) { std::cerr << "Error: " << pd_last_error(nullptr) << std::endl; drop_outputs(); return 1; }
    if (!multi_gpu) engs.v.assign(1, eng);
    tm.mark(multi_gpu ? "decode+push (multi-GPU)" : "decode+push");
    if (feeder.failed) { std::cerr << feeder.err << std::endl; drop_outputs(); pd_destroy(eng); return 1; }
    for (pd_engine *e : engs.v) if (pd_finalize(e) != PD_OK) { engine_error(e); drop_outputs(); return 1; }
    std::cout << "INFO: Input data read done" << std::endl;                                                                        // :4879

    if (o.site_output && !write_site_depth(eng, prefix + ".SiteDepth.gz", targets, genes, o.threads)) { drop_outputs(); return 1; }
    tm.mark("finalize(+site dump)");

    // ---- reduce + report -----------------------------------------------------------------------------------------
    Totals tot;
    out.put(stat_header(mode));
    const bool ok = mode == 6 ? report_windows(eng, out, o, targets, genes, tot)
                  : mode == 5 ? report_tiles(engs, out, o, targets, genes, mask, tot)
                              : report_genes(engs, out, o, targets, genes, merged, mode, whole, mask, tot);
    if (!ok) { drop_outputs(); return 1; }
    {
        char row[512];
        const double Coverage = tot.cov * 100.0 / tot.len, MeanDepth = tot.depth * 1.0 / tot.len;   // :5119-5122
        int m = snprintf(row, sizeof row, "##RegionLength: %llu\tCoveredSite: %llu\tCoverage(%%): %.2f\tMeanDepth: %.2f\n",
                         (unsigned long long)tot.len, (unsigned long long)tot.cov, Coverage, MeanDepth);
        out.put(std::string(row, (size_t)m));
    }
    if (!out.close()) { drop_outputs(); return 1; }
    tm.mark("reduce+report");
    // every output is flushed and closed: leave without the teardown of the CUDA context, 24 GB of device memory and the
    // parked/pinned host buffers -- the OS reclaims all of it, and the orderly teardown only costs wall time
    fflush(stdout); fflush(stderr);
    if (!getenv("PANDEPTH_ORDERLY_EXIT")) _exit(0);
    for (pd_engine *e : engs.v) pd_destroy(e);
    if (in.hdr) sam_hdr_destroy(in.hdr);
    tm.mark("pd_destroy");
    return 0;
}
