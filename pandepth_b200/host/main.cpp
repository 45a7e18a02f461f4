// main.cpp -- `pandepth` on a B200: the reference's command line and output files, with the depth path on the GPU.
//
//   host threads (htslib decode)  --(tid, pos, cigar batches, pinned memory)-->  pd_push_contig   [role of ProDealChrBambai*]
//   pd_finalize / pd_reduce / pd_reduce_windows / pd_depth_read                                   [role of StatChrDepth*, mode-6 loop, -a]
//   report writer (same doubles, same text)                                                       [PanDepth.cpp:4879-5127]
//
// Decode parallelism is by genomic range through the BAI/CSI index (any number of threads on one contig), not by contig
// as in the reference (:4415-4523).
#include "pd_host.hpp"
#include "../../include/pandepth_b200.h"

#include <zlib.h>
#include "sam.h"
#include "hts.h"

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>

using namespace pdh;

namespace {

struct GzOut {
    gzFile fp = nullptr;
    std::string buf;
    bool open(const std::string &path) { fp = gzopen(path.c_str(), "wb"); if (fp) gzbuffer(fp, 1 << 20); buf.reserve(1 << 22); return fp != nullptr; }
    void put(const std::string &s) { buf += s; if (buf.size() > (1u << 22)) flush(); }
    void put(const char *s, size_t n) { buf.append(s, n); if (buf.size() > (1u << 22)) flush(); }
    void flush() { if (!buf.empty()) { gzwrite(fp, buf.data(), (unsigned)buf.size()); buf.clear(); } }
    void close() { if (fp) { flush(); gzclose(fp); fp = nullptr; } }
};

// growable pinned SoA batch of one contig's records
struct Batch {
    int32_t *pos = nullptr; uint32_t *cigar = nullptr; uint32_t *off = nullptr;
    size_t cap_r = 0, cap_c = 0, n = 0, nc = 0;
    bool single_op = true;
    ~Batch() { pd_host_free(pos); pd_host_free(cigar); pd_host_free(off); }
    bool reserve_reads(size_t want)
    {
        if (want <= cap_r) return true;
        size_t cap = std::max<size_t>(want, cap_r ? cap_r * 2 : (1u << 20));
        int32_t *p = (int32_t *)pd_host_alloc(cap * 4); uint32_t *o = (uint32_t *)pd_host_alloc((cap + 1) * 4);
        if (!p || !o) return false;
        if (n) { memcpy(p, pos, n * 4); memcpy(o, off, (n + 1) * 4); }
        pd_host_free(pos); pd_host_free(off); pos = p; off = o; cap_r = cap;
        return true;
    }
    bool reserve_cigar(size_t want)
    {
        if (want <= cap_c) return true;
        size_t cap = std::max<size_t>(want, cap_c ? cap_c * 2 : (1u << 20));
        uint32_t *c = (uint32_t *)pd_host_alloc(cap * 4);
        if (!c) return false;
        if (nc) memcpy(c, cigar, nc * 4);
        pd_host_free(cigar); cigar = c; cap_c = cap;
        return true;
    }
    void clear() { n = 0; nc = 0; single_op = true; }
    bool add(const bam1_t *b)
    {
        const uint32_t k = b->core.n_cigar;
        if (!reserve_reads(n + 1) || !reserve_cigar(nc + k)) return false;
        pos[n] = (int32_t)b->core.pos;                      // hts_pos_t truncated to int32 like :736
        off[n] = (uint32_t)nc;
        memcpy(cigar + nc, bam_get_cigar(b), (size_t)k * 4);
        nc += k; n++;
        off[n] = (uint32_t)nc;
        if (k != 1) single_op = false;
        return true;
    }
};

struct WorkItem { int tid; int64_t beg, end, prev_end; int r0, r1; };   // 0-based [beg,end); merged regions [r0,r1) of the contig (modes 1-4)

struct Feeder {
    const Options &o;
    pd_engine *eng;
    std::mutex eng_mu;
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    std::string err;
    std::mutex err_mu;
    std::vector<WorkItem> items;
    const std::map<int, std::vector<std::pair<int, int>>> &merged;   // per tid, 1-based closed, modes 1-4 only
    bool filter_by_region;
    std::atomic<uint64_t> reads_pushed{0};

    Feeder(const Options &oo, pd_engine *e, const std::map<int, std::vector<std::pair<int, int>>> &m, bool f)
        : o(oo), eng(e), merged(m), filter_by_region(f) {}

    void fail(const std::string &m) { std::lock_guard<std::mutex> g(err_mu); if (!failed.exchange(true)) err = m; }

    bool push(int tid, Batch &b, bool sorted)
    {
        if (b.n == 0) return true;
        std::lock_guard<std::mutex> g(eng_mu);
        int rc = pd_push_contig(eng, tid, (int64_t)b.n, b.pos, nullptr, nullptr, b.single_op ? nullptr : b.off, b.cigar, sorted ? PD_SORTED : 0u);
        if (rc != PD_OK) { fail(std::string("pd_push_contig: ") + pd_last_error(eng)); return false; }
        reads_pushed += b.n;
        b.clear();
        return true;
    }

    htsFile *open_input(sam_hdr_t **hdr, int nthreads)
    {
        htsFile *fp = hts_open(o.input.c_str(), "r");
        if (!fp) return nullptr;
        hts_set_log_level(HTS_LOG_OFF);
        if (fp->format.format == cram) {                                   // :637-643
            hts_set_fai_filename(fp, o.reference.c_str());
            hts_set_opt(fp, CRAM_OPT_DECODE_MD, 0);
            hts_set_opt(fp, CRAM_OPT_REQUIRED_FIELDS, SAM_FLAG | SAM_RNAME | SAM_POS | SAM_MAPQ | SAM_CIGAR);
        }
        if (nthreads > 0) hts_set_opt(fp, HTS_OPT_NTHREADS, nthreads);
        *hdr = sam_hdr_read(fp);
        if (!*hdr) { hts_close(fp); return nullptr; }
        return fp;
    }

    // one worker of the indexed path: own file handle + index, ranges taken from the shared queue
    void indexed_worker()
    {
        sam_hdr_t *hdr = nullptr;
        htsFile *fp = open_input(&hdr, 0);
        hts_idx_t *idx = fp ? sam_index_load(fp, o.input.c_str()) : nullptr;
        if (!fp || !idx) { fail("Error: Failed to open the index file or BAM/CRAM file: " + o.input); if (fp) hts_close(fp); return; }
        bam1_t *b = bam_init1();
        Batch batch;
        const uint32_t xf = o.xflags;
        while (!failed) {
            const size_t k = next.fetch_add(1);
            if (k >= items.size()) break;
            const WorkItem &w = items[k];
            hts_itr_t *it = sam_itr_queryi(idx, w.tid, w.beg, w.end);
            if (!it) continue;
            const std::vector<std::pair<int, int>> *regs = nullptr;
            if (filter_by_region) { auto f = merged.find(w.tid); if (f != merged.end()) regs = &f->second; }
            int rcur = w.r0;
            while (sam_itr_next(fp, it, b) >= 0) {
                if (b->core.flag & xf) continue;                              // :732
                if ((int)b->core.qual < o.min_mapq) continue;                 // :733
                const int64_t p = b->core.pos;
                if (p < w.prev_end) continue;                                 // an earlier range already took this read
                if (regs) {
                    // keep only reads the reference would fetch: overlapping a merged region's window chr:first-1..second+1 (:700-706)
                    const int64_t e = bam_endpos(b);
                    while (rcur < w.r1 && (int64_t)(*regs)[rcur].second + 1 <= p) rcur++;       // window end (0-based excl) = second+1
                    bool hit = false;
                    for (int r = rcur; r < w.r1; r++) {
                        const int64_t wb = std::max<int64_t>((int64_t)(*regs)[r].first - 1, 1) - 1;    // 0-based start of the window
                        if (wb >= e) break;
                        if ((int64_t)(*regs)[r].second + 1 > p) { hit = true; break; }
                    }
                    if (!hit) continue;
                }
                if (!batch.add(b)) { fail("out of pinned host memory"); break; }
                if (batch.n >= (4u << 20)) { if (!push(w.tid, batch, true)) break; }
            }
            hts_itr_destroy(it);
            if (!push(w.tid, batch, true)) break;
        }
        bam_destroy1(b);
        hts_idx_destroy(idx);
        sam_hdr_destroy(hdr);
        hts_close(fp);
    }

    // no-index path (:4604-4713): one pass over the file, batches split at contig changes
    bool stream_all(bool sorted)
    {
        sam_hdr_t *hdr = nullptr;
        htsFile *fp = open_input(&hdr, o.threads);
        if (!fp) { fail("Error: Failed to open the BAM/CRAM file: " + o.input); return false; }
        bam1_t *b = bam_init1();
        Batch batch;
        int cur = -1;
        while (sam_read1(fp, hdr, b) >= 0) {
            if ((int)b->core.qual < o.min_mapq) continue;
            if (b->core.flag & o.xflags) continue;
            const int t = b->core.tid;
            if (t < 0) continue;
            if (t != cur) { if (!push(cur, batch, sorted)) break; cur = t; }
            if (!batch.add(b)) { fail("out of pinned host memory"); break; }
            if (batch.n >= (4u << 20)) { if (!push(cur, batch, sorted)) break; }
        }
        if (!failed) push(cur, batch, sorted);
        bam_destroy1(b);
        sam_hdr_destroy(hdr);
        hts_close(fp);
        return !failed;
    }
};

bool file_exists(const std::string &p) { return access(p.c_str(), 0) == 0; }

}  // namespace

int main(int argc, char **argv)
{
    Options o;
    if (parse_args(argc, argv, o) != 0) return 0;

    std::string ext = o.input.substr(o.input.rfind('.') == std::string::npos ? o.input.size() : o.input.rfind('.') + 1);
    if (ext == "gz") {
        std::string a = o.input.substr(0, o.input.rfind('.'));
        ext = a.substr(a.rfind('.') == std::string::npos ? a.size() : a.rfind('.') + 1);
    }
    if (ext == "paf" || ext == "PAF") { std::cerr << "Error: PAF input is not supported by pandepth-b200 yet" << std::endl; return 1; }

    // header (:3482-3507)
    htsFile *in = hts_open(o.input.c_str(), "r");
    hts_set_log_level(HTS_LOG_OFF);
    if (!in) { std::cerr << "Error: Failed to open the BAM/CRAM file: " << o.input << std::endl; return 1; }
    if (in->format.format == cram) hts_set_fai_filename(in, o.reference.c_str());
    sam_hdr_t *hdr = sam_hdr_read(in);
    if (!hdr) { std::cerr << "Error: Failed to read the header for the BAM/CRAM file: " << o.input << std::endl; return 1; }
    std::vector<Target> targets((size_t)hdr->n_targets);
    for (int i = 0; i < hdr->n_targets; i++) { targets[i].name = hdr->target_name[i]; targets[i].len = hdr->target_len[i]; }
    bool sorted_hdr = false;
    {
        const char *txt = sam_hdr_str(hdr);
        std::string text = txt ? txt : "";
        size_t a = text.find("\tSO:");                                       // :4537-4549
        if (a != std::string::npos) { a += 4; size_t b = text.find_first_of("\n\t", a); sorted_hdr = text.substr(a, b - a) == "coordinate"; }
    }
    hts_close(in);

    GeneMap genes;
    std::string err;
    if (!build_regions(o, targets, genes, err)) { std::cerr << err << std::endl; return 1; }
    const int mode = o.region_mode;

    const std::string prefix = output_prefix(o.out_arg);
    GzOut out;
    const std::string stat_path = prefix + stat_suffix(mode);
    if (!out.open(stat_path)) { std::cerr << "open OUT File error: " << stat_path << std::endl; return 0; }

    const bool have_index = file_exists(o.input + ".bai") || file_exists(o.input + ".crai") || file_exists(o.input + ".csi");   // :4121-4124
    const bool indexed = have_index && o.use_index;
    // the reference's 18-bit SiteInfo paths: no index, -a, or mode 6 (:4127, :4553); everything else is the uint32 window path
    const bool whole = !indexed || o.site_output || mode == 6;
    const uint32_t mask = whole ? 0x3FFFFu : 0xFFFFFFFFu;

    // merged regions per contig (RegionMerger, :3912-3972)
    std::map<int, std::vector<std::pair<int, int>>> merged;
    for (const auto &kv : genes) merged[kv.first] = merge_regions(kv.second);

    int device = 0;
    if (const char *d = getenv("PANDEPTH_DEVICE")) device = atoi(d);
    std::vector<uint32_t> tlen(targets.size());
    std::vector<uint8_t> owned(targets.size(), 0);
    for (size_t t = 0; t < targets.size(); t++) { tlen[t] = targets[t].len; owned[t] = genes.count((int)t) ? 1 : 0; }
    pd_engine *eng = pd_create(device, (int)targets.size(), tlen.data(), owned.data(), o.xflags, o.min_mapq);
    if (!eng) { std::cerr << "Error: " << pd_last_error(nullptr) << std::endl; return 1; }

    // ---- feed ---------------------------------------------------------------------------------------------------
    const bool region_modes = mode >= 1 && mode <= 4;
    Feeder feeder(o, eng, merged, indexed && region_modes);
    if (indexed) {
        const int64_t RANGE = 4 << 20;                       // bases per work item when tiling a whole contig
        const int64_t GAP = 64 << 10;                        // merged regions closer than this share one index query
        for (const auto &kv : merged) {
            const int tid = kv.first;
            const int64_t len = tlen[tid];
            int64_t prev_end = 0;
            if (!region_modes) {
                for (int64_t b = 0; b < len; b += RANGE) { feeder.items.push_back({tid, b, std::min(len, b + RANGE), prev_end, 0, 0}); prev_end = std::min(len, b + RANGE); }
            } else {
                const auto &regs = kv.second;
                size_t a = 0;
                while (a < regs.size()) {
                    size_t b = a + 1;
                    int64_t beg = std::max<int64_t>((int64_t)regs[a].first - 1, 1) - 1;             // chr:first-1 .. second+1 (:700-706)
                    int64_t end = std::min<int64_t>((int64_t)regs[a].second + 1, len);
                    while (b < regs.size() && (int64_t)regs[b].first - 2 - end < GAP && end - beg < 2 * RANGE) { end = std::min<int64_t>((int64_t)regs[b].second + 1, len); b++; }
                    if (end > beg) feeder.items.push_back({tid, beg, end, prev_end, (int)a, (int)b});
                    prev_end = std::max(prev_end, end);
                    a = b;
                }
            }
        }
        const int nthreads = (int)std::max<size_t>(1, std::min<size_t>((size_t)o.threads, feeder.items.size()));
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; i++) th.emplace_back([&] { feeder.indexed_worker(); });
        for (auto &t : th) t.join();
    } else {
        if (sorted_hdr) std::cout << "Warning: PanDepth will run in No Index mode: " << o.input << std::endl;                       // :4606
        else std::cout << "Warning: Can't find index file of input BAM/CRAM. PanDepth will run in No Index mode: " << o.input << std::endl;   // :4677
        feeder.stream_all(sorted_hdr);
    }
    if (feeder.failed) { std::cerr << feeder.err << std::endl; pd_destroy(eng); return 1; }
    if (pd_finalize(eng) != PD_OK) { std::cerr << "Error: " << pd_last_error(eng) << std::endl; pd_destroy(eng); return 1; }
    std::cout << "INFO: Input data read done" << std::endl;                                                                        // :4879

    // ---- -a per-site dump (:4264-4284) ---------------------------------------------------------------------------
    if (o.site_output) {
        GzOut site;
        if (site.open(prefix + ".SiteDepth.gz")) {
            std::vector<uint32_t> d;
            const int32_t CH = 4 << 20;
            for (size_t t = 0; t < targets.size(); t++) {
                if (!genes.count((int)t)) continue;
                const std::string &nm = targets[t].name;
                for (int64_t s = 0; s < (int64_t)tlen[t]; s += CH) {
                    const int32_t n = (int32_t)std::min<int64_t>(CH, (int64_t)tlen[t] - s);
                    d.resize((size_t)n);
                    if (pd_depth_read(eng, (int32_t)t, (int32_t)s, n, d.data(), 0) != PD_OK) { std::cerr << "Error: " << pd_last_error(eng) << std::endl; return 1; }
                    char line[96];
                    for (int32_t j = 0; j < n; j++) {
                        int m = snprintf(line, sizeof line, "\t%lld\t%u\n", (long long)(s + j), d[j] & 0x3FFFFu);
                        site.put(nm); site.put(line, (size_t)m);
                    }
                }
            }
            site.close();
        }
    }

    // ---- reduce + report -----------------------------------------------------------------------------------------
    uint64_t SS_Len = 0, SS_Cov = 0, SS_TotalD = 0;
    out.put(stat_header(mode));
    char row[512];
    if (mode == 6) {                                                                         // :4353-4394
        std::vector<uint32_t> cov; std::vector<uint64_t> sum;
        for (size_t t = 0; t < targets.size(); t++) {
            if (!genes.count((int)t)) continue;
            const int64_t CC = tlen[t], W = o.win;
            const int64_t n_win = CC > 1 ? (CC - 1 + W - 1) / W : 0;                         // for (j=1; j<CC; j+=W)
            cov.assign((size_t)n_win, 0); sum.assign((size_t)n_win, 0);
            if (n_win && pd_reduce_windows(eng, (int32_t)t, 0, (int32_t)W, n_win, (uint32_t)o.min_dep, 0x3FFFFu, cov.data(), sum.data(), 0) != PD_OK) {
                std::cerr << "Error: " << pd_last_error(eng) << std::endl; return 1;
            }
            for (int64_t k = 0; k < n_win; k++) {
                const int64_t j = 1 + k * W;
                const int64_t End = std::min(j - 1 + W, CC);
                const int GeneLength = (int)(End - j + 1), GeneCover = (int)cov[k], GeneDepth = (int)sum[k];   // ints (:4364)
                const double Coverage = GeneCover * 100.0 / GeneLength, MeanDepth = GeneDepth * 1.0 / GeneLength;
                int m = snprintf(row, sizeof row, "%s\t%lld\t%lld\t%d\t%d\t%d\t%.2f\t%.2f\n", targets[t].name.c_str(), (long long)j, (long long)End,
                                 GeneLength, GeneCover, GeneDepth, Coverage, MeanDepth);
                out.put(row, (size_t)m);
                SS_Cov += (uint64_t)(int64_t)GeneCover; SS_Len += (uint64_t)(int64_t)GeneLength; SS_TotalD += (uint64_t)(int64_t)GeneDepth;
            }
        }
    } else {
        // every CDS interval of every gene in one pd_reduce call (StatChrDepthWin / StatChrDepthLowMEM)
        std::vector<int32_t> iv_t, iv_s, iv_e;
        std::vector<Gene *> iv_g;
        for (auto &kv : genes) {
            std::map<std::string, int> times;
            if (!whole) times = indexed_window_counts(merged[kv.first], kv.second, (int)tlen[kv.first]);
            for (auto &gk : kv.second) {
                if (!whole && times[gk.first] == 0) continue;                                 // never handed to StatChrDepthWin (:299-308)
                for (const auto &c : gk.second.cds) { iv_t.push_back(kv.first); iv_s.push_back(c.first - 1); iv_e.push_back(c.second); iv_g.push_back(&gk.second); }
            }
        }
        std::vector<uint64_t> cov(iv_t.size()), sum(iv_t.size());
        if (pd_reduce(eng, (int64_t)iv_t.size(), iv_t.data(), iv_s.data(), iv_e.data(), (uint32_t)o.min_dep, mask, cov.data(), sum.data(), 0) != PD_OK) {
            std::cerr << "Error: " << pd_last_error(eng) << std::endl; return 1;
        }
        for (size_t i = 0; i < iv_g.size(); i++) { iv_g[i]->cover += (int64_t)cov[i]; iv_g[i]->depth += sum[i]; }

        for (auto &kv : genes) {
            const std::string &chr = targets[kv.first].name;
            if (mode == 0) {                                                                   // :5051-5079
                uint64_t L = 0, C = 0, D = 0;
                for (auto &gk : kv.second) { L += gk.second.length; C += (uint64_t)(int64_t)(int)gk.second.cover; D += gk.second.depth; }
                SS_Len += L; SS_Cov += C; SS_TotalD += D;
                int m = snprintf(row, sizeof row, "%s\t%llu\t%llu\t%llu\t%s\t%s\n", chr.c_str(), (unsigned long long)L, (unsigned long long)C,
                                 (unsigned long long)D, fmt_ratio_pct((double)C, (double)L).c_str(), fmt_ratio((double)D, (double)L).c_str());
                out.put(row, (size_t)m);
            } else {                                                                           // :5011-5050 / :5080-5118
                std::map<int, std::string> by_start;                                           // rows by GeneStart; ties in GeneID order
                for (auto &gk : kv.second) {
                    const Gene &g = gk.second;
                    const int GeneCover = (int)g.cover;
                    SS_Cov += (uint64_t)(int64_t)GeneCover; SS_Len += g.length; SS_TotalD += g.depth;
                    const double Coverage = GeneCover * 100.0 / g.length, MeanDepth = g.depth * 1.0 / g.length;
                    std::string line = chr + "\t" + std::to_string(g.start) + "\t" + std::to_string(g.end) + "\t";
                    if (mode != 5) line += gk.first + "\t";
                    char num[160];
                    snprintf(num, sizeof num, "%llu\t%d\t%llu\t%.2f\t%.2f", (unsigned long long)g.length, GeneCover, (unsigned long long)g.depth, Coverage, MeanDepth);
                    line += num;
                    auto it = by_start.find(g.start);
                    if (it == by_start.end()) by_start.emplace(g.start, line);
                    else { it->second += "\n"; it->second += line; }
                }
                for (auto &r : by_start) { out.put(r.second); out.put("\n", 1); }
            }
        }
    }
    {
        const double Coverage = SS_Cov * 100.0 / SS_Len, MeanDepth = SS_TotalD * 1.0 / SS_Len;   // :5119-5122
        int m = snprintf(row, sizeof row, "##RegionLength: %llu\tCoveredSite: %llu\tCoverage(%%): %.2f\tMeanDepth: %.2f\n",
                         (unsigned long long)SS_Len, (unsigned long long)SS_Cov, Coverage, MeanDepth);
        out.put(row, (size_t)m);
    }
    out.close();
    pd_destroy(eng);
    sam_hdr_destroy(hdr);
    return 0;
}
