// regions.cpp -- command line, region model and naming, restated from /root/reference/src/PanDepth.cpp so that the
// interval table (and therefore every printed row) is identical to the reference's.
#include "pd_host.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

namespace pdh {

// ---------------------------------------------------------------------------------------------------------------
// small text helpers
// ---------------------------------------------------------------------------------------------------------------
namespace {

// include/comm.h:98-109 split(): tokens between any delimiter character, empties dropped
std::vector<std::string> split_any(const std::string &s, const char *delims)
{
    std::vector<std::string> out;
    size_t i = s.find_first_not_of(delims, 0);
    while (i != std::string::npos) {
        size_t j = s.find_first_of(delims, i);
        out.push_back(s.substr(i, j == std::string::npos ? std::string::npos : j - i));
        if (j == std::string::npos) break;
        i = s.find_first_not_of(delims, j);
    }
    return out;
}

std::vector<std::string> split_ws(const std::string &s) { return split_any(s, " \t\r\n\v\f"); }

std::string strip_all(std::string s, char ch)
{
    s.erase(std::remove(s.begin(), s.end(), ch), s.end());
    return s;
}

void add_interval(GeneMap &genes, int tid, const std::string &id, long long s, long long e)
{
    auto &g = genes[tid];
    auto it = g.find(id);
    if (it == g.end()) {                                   // :3601-3604
        Gene x;
        x.start = (int)s; x.end = (int)e; x.length = (uint64_t)(e - s + 1);
        x.cds.push_back({(int)s, (int)e});
        g.emplace(id, std::move(x));
    } else {                                               // :3638-3641
        Gene &x = it->second;
        if (x.start > s) x.start = (int)s;
        if (x.end < e) x.end = (int)e;
        x.length += (uint64_t)(e - s + 1);
        x.cds.push_back({(int)s, (int)e});
    }
}

void lack(const std::string &f) { std::cerr << "Error: Lack argument for [ -" << f << " ]" << std::endl; }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// command line (:84-293).  Same letters, same defaults, same oddities: any token not starting with '-' aborts, every
// '-' is stripped from a flag, -s forces the no-index path.
// ---------------------------------------------------------------------------------------------------------------
void print_help()
{
    std::cout << "Usage: pandepth -i in.bam [-g gene.gff | -b region.bed] -o outPrefix\n"
                 " Input/Output options:\n"
                 "   -i    <str>     input of sam/bam/cram or #.list file\n"
                 "   -o    <str>     prefix of output file\n"
                 " Target options:\n"
                 "   -g    <str>     input gff/gtf file for gene region\n"
                 "   -f    <str>     gff/gtf feature type to parse, CDS or exon [CDS]\n"
                 "   -b    <str>     input bed file for list of regions\n"
                 "   -w    <int>     windows size (bp)\n"
                 "   -a              output all the site depth\n"
                 " Filter options:\n"
                 "   -q    <int>     min mapping quality [0]\n"
                 "   -d    <int>     min site depth for statistics [1]\n"
                 "   -x    <int>     exclude reads with any of the bits in FLAG set [1796]\n"
                 " Other options:\n"
                 "   -t    <int>     number of host decode threads [3]\n"
                 "   -r    <str>     reference genome file for cram decode\n"
                 "   -h              show this help [pandepth-b200, CLI of PanDepth v2.26]\n"
                 " Environment:\n"
                 "   PANDEPTH_DEVICE   CUDA device ordinal [0]\n"
                 "\n";
}

int parse_args(int argc, char **argv, Options &o)
{
    if (argc <= 1) { print_help(); return 1; }
    bool have_bed = false;
    std::string gff_mode_from_sniff;
    for (int i = 1; i < argc; i++) {
        if (argv[i][0] != '-') {
            std::cerr << "Error: Command option error! Please check the provided options." << std::endl;
            return 1;
        }
        std::string flag = strip_all(argv[i], '-');
        auto need = [&](std::string &dst) -> bool {
            if (i + 1 == argc) { lack(flag); return false; }
            dst = argv[++i];
            return true;
        };
        std::string v;
        if (flag == "i") {
            if (!need(v)) return 1;
            std::string ext = v.substr(v.rfind('.') == std::string::npos ? v.size() : v.rfind('.') + 1);
            if (ext == "list" || ext == "List") {                  // ReadList, include/comm.h:31-50: one path per non-empty line
                LineReader in(v);
                if (!in.ok()) std::cerr << "open List error: " << v << std::endl;
                else { std::string line; while (in.next(line)) if (!line.empty()) o.inputs.push_back(line); }
            } else o.inputs.push_back(v);
        } else if (flag == "o") { if (!need(o.out_arg)) return 1; }
        else if (flag == "c") o.gc = true;
        else if (flag == "a") o.site_output = true;
        else if (flag == "r") { if (!need(o.reference)) return 1; }
        else if (flag == "f") { if (!need(o.feature)) return 1; }
        else if (flag == "x") { if (!need(v)) return 1; o.xflags = (uint32_t)atoi(v.c_str()); }
        else if (flag == "g") {
            if (!need(o.gff)) return 1;
            LineReader in(o.gff);
            if (!in.ok()) { std::cerr << "Error: Failed to open the file: " << o.gff << std::endl; return 1; }
            std::string line;                                       // :162-181: first 167 lines, last match wins
            int mode = 0;
            for (int k = 1; k < 168 && in.next(line); k++) {
                if (line.size() < 2 || line[0] == '#') continue;
                if (line.find("Parent") != std::string::npos) mode = 1;
                else if (line.find("transcript_id") != std::string::npos) mode = 2;
            }
            if (mode) o.region_mode = mode;
            if (o.region_mode == 0) {
                std::cerr << "Error: The format of the input GFF/GTF file is incorrect. Please check the file format: " << o.gff << std::endl;
                return 1;
            }
        } else if (flag == "b") { if (!need(o.bed)) return 1; have_bed = true; o.region_mode = 3; }
        else if (flag == "t") { if (!need(v)) return 1; o.threads = atoi(v.c_str()); }
        else if (flag == "w") {
            if (!need(v)) return 1;
            o.win = atoi(v.c_str());
            if (o.win < 1) { std::cerr << "Warning: -w should >= 1, set to 1\n"; o.win = 1; }
        } else if (flag == "q") { if (!need(v)) return 1; o.min_mapq = atoi(v.c_str()); }
        else if (flag == "s") o.use_index = false;
        else if (flag == "d") { if (!need(v)) return 1; o.min_dep = atoi(v.c_str()); if (o.min_dep < 1) o.min_dep = 1; }
        else if (flag == "help" || flag == "h") { print_help(); return 1; }
        else { std::cerr << "Error UnKnow argument -" << flag << std::endl; return 1; }
    }
    if (!o.inputs.empty()) o.input = o.inputs[0];
    if (o.input.empty() || o.out_arg.empty()) { std::cerr << "Error: lack argument -i or -o " << std::endl; return 1; }
    if (have_bed && o.gff.empty()) {                              // :262-289
        LineReader in(o.bed);
        if (!in.ok()) { std::cerr << "Error: Failed to open BED file: " << o.bed << std::endl; return 1; }
        std::string l1, l2;
        in.next(l1); in.next(l2);
        if (split_any(l1, " \t").size() == 4 || split_any(l2, " \t").size() == 4) o.region_mode = 4;
    }
    if (o.gc) { std::cerr << "Error: GC content (-c) is not part of the depth path and is not supported by pandepth-b200" << std::endl; return 2; }   // exit status 1
    if (o.threads < 1) o.threads = 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// region parsers (:3547-3905) and pseudo-gene tiling (:3975-4051)
// ---------------------------------------------------------------------------------------------------------------
bool build_regions(Options &o, const std::vector<Target> &targets, GeneMap &genes, std::string &err)
{
    std::map<std::string, int> name2tid;
    for (size_t i = 0; i < targets.size(); i++) name2tid.emplace(targets[i].name, (int)i);     // first wins (:3502)
    auto warn = [](const std::string &line) { std::cerr << line << "Warning: This region may be incorrect.\n" << std::endl; };

    if (o.region_mode != 0) {
        const std::string &path = !o.gff.empty() ? o.gff : o.bed;                                // InStr2 (:262-266)
        LineReader in(path);
        if (!in.ok()) { err = "Error: Cannot open the GFF/GTF File: " + path; return false; }
        std::string line;
        // the reference's loop variables outlive a line, so a short line inherits the previous line's fields
        std::string chr, s_str, e_str, gid;
        long long S = 0, E = 0;
        while (in.next(line)) {
            if (line.empty() || line[0] == '#') continue;
            if (o.region_mode == 1 || o.region_mode == 2) {
                if (o.region_mode == 2) { line = strip_all(line, '"'); line = strip_all(line, ';'); }  // :3657-3658
                std::vector<std::string> tok = split_ws(line);
                if (tok.size() >= 1) chr = tok[0];
                if (tok.size() < 3 || tok[2] != o.feature) continue;                                 // :3570
                S = tok.size() > 3 ? atoll(tok[3].c_str()) : 0;
                E = tok.size() > 4 ? atoll(tok[4].c_str()) : 0;
                std::string id;
                if (o.region_mode == 1) {
                    const std::string attr = tok.size() > 8 ? tok[8] : std::string();
                    std::vector<std::string> inf = split_any(attr, ",;");                             // :3575
                    if (!inf.empty()) { std::vector<std::string> t = split_any(inf[0], "="); if (!t.empty()) id = t.back(); }
                    for (size_t k = 1; k < inf.size(); k++) {
                        std::vector<std::string> t = split_any(inf[k], "=");
                        if (!t.empty() && t[0] == "Parent") id = t.back();
                    }
                } else {
                    std::vector<std::string> inf = split_any(line, "\t ");                           // :3668-3669
                    if (inf.size() > 9) id = inf[9];
                }
                auto it = name2tid.find(chr);
                if (it == name2tid.end()) { warn(line); continue; }
                add_interval(genes, it->second, id, S, E);
            } else if (o.region_mode == 3) {
                std::vector<std::string> tok = split_ws(line);
                if (tok.size() > 0) chr = tok[0];
                if (tok.size() > 1) s_str = tok[1];
                if (tok.size() > 2) e_str = tok[2];
                gid = chr + "_" + s_str + "_" + e_str;                                               // :3752
                S = atoi(s_str.c_str()); E = atoi(e_str.c_str());
                if (S > E) { warn(line); continue; }
                auto it = name2tid.find(chr);
                if (it == name2tid.end()) { warn(line); continue; }
                add_interval(genes, it->second, gid, S, E);
            } else {                                                                                 // mode 4 (:3831)
                std::vector<std::string> tok = split_ws(line);
                if (tok.size() > 0) chr = tok[0];
                if (tok.size() > 1) S = atoi(tok[1].c_str());
                if (tok.size() > 2) E = atoi(tok[2].c_str());
                if (tok.size() > 3) gid = tok[3];
                if (S > E) { warn(line); continue; }
                auto it = name2tid.find(chr);
                if (it == name2tid.end()) { warn(line); continue; }
                add_interval(genes, it->second, gid, S, E);
            }
        }
    }
    if (genes.empty()) {
        long long W = 10000000;
        if (o.win == 0) o.region_mode = 0;
        else if (o.win < 150) o.region_mode = 6;
        else { o.region_mode = 5; o.tile_win = o.win; }       // W-wide tiles stay implicit (pd_host.hpp); 10 Mb tiles mark the contigs
        for (size_t t = 0; t < targets.size(); t++) {
            const long long len = targets[t].len;
            long long Start = 1, End = 2;                      // End pre-set to 2, bumped by 2 after each tile (:3997-4044)
            for (Start = 1; End <= len; Start += W) {
                End = Start + W - 1;
                if (End > len) End = len;
                add_interval(genes, (int)t, targets[t].name + std::to_string(Start), Start, End);
                End += 2;
            }
        }
    }
    return true;
}

std::vector<std::pair<int, int>> merge_regions(const std::map<std::string, Gene> &genes)
{
    std::map<int, int> by_start;
    for (const auto &kv : genes) {
        auto it = by_start.find(kv.second.start);
        if (it == by_start.end()) by_start.emplace(kv.second.start, kv.second.end);
        else if (kv.second.end > it->second) it->second = kv.second.end;
    }
    std::vector<std::pair<int, int>> out;
    int End = 0;
    for (const auto &se : by_start) {
        if (out.empty()) { out.push_back(se); End = se.second; }
        else if (se.first > End) { out.push_back(se); End = se.second; }
        else if (se.second > End) { out.back().second = se.second; End = se.second; }
    }
    return out;
}

std::map<std::string, int> indexed_window_counts(const std::vector<std::pair<int, int>> &regions,
                                                 const std::map<std::string, Gene> &genes, int chr_len)
{
    const int W = 10000000;
    std::map<std::string, int> count;
    for (const auto &kv : genes) count[kv.first] = 0;
    if (regions.empty()) return count;
    int MeMStart = regions[0].first; if (MeMStart < 1) MeMStart = 1;
    int MeMEnd = MeMStart + W - 1; if (MeMEnd > chr_len) MeMEnd = chr_len;
    for (size_t k = 0; k < regions.size(); k++) {
        long end = (long)regions[k].second + 1; if (end > chr_len) end = chr_len;
        const bool last = (k + 1 == regions.size());
        if (end >= MeMEnd || last) {
            MeMEnd = (int)end;
            for (const auto &kv : genes) {
                const Gene &g = kv.second;
                if (MeMEnd - MeMStart != 0) { if (g.start >= MeMEnd || g.end < MeMStart) continue; }
                else { if (g.start > MeMEnd || g.end < MeMStart) continue; }
                count[kv.first]++;
            }
            if (!last) { MeMStart = regions[k + 1].first; if (MeMStart - 150 > MeMEnd) MeMStart -= 150; }
            MeMEnd = MeMStart + W; if (MeMEnd > chr_len) MeMEnd = chr_len;
        }
    }
    return count;
}

// ---------------------------------------------------------------------------------------------------------------
// naming and formatting
// ---------------------------------------------------------------------------------------------------------------
std::string output_prefix(const std::string &out_arg)
{
    // add_Asuffix appends ".gz" unless already there (include/comm.h:52-60); :4057 strips 3 chars; :4059-4066 drops .stat/.bed
    std::string p = out_arg;
    std::string ext = p.substr(p.rfind('.') == std::string::npos ? p.size() : p.rfind('.') + 1);
    if (ext == "gz") p = p.substr(0, p.size() - 3);
    ext = p.substr(p.rfind('.') == std::string::npos ? p.size() : p.rfind('.') + 1);
    if (ext == "stat" || ext == "bed") p = p.substr(0, p.rfind('.'));
    return p;
}

const char *stat_suffix(int mode)
{
    switch (mode) {
    case 3: case 4: return ".bed.stat.gz";
    case 5: case 6: return ".win.stat.gz";
    case 0: return ".chr.stat.gz";
    default: return ".gene.stat.gz";
    }
}

const char *stat_header(int mode)
{
    switch (mode) {
    case 3: return "#Chr\tStart\tEnd\tRegionID\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n";
    case 5: case 6: return "#Chr\tStart\tEnd\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n";
    case 0: return "#Chr\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n";
    default: return "#Chr\tStart\tEnd\tGeneID\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n";
    }
}

std::string fmt_ratio_pct(double cov, double len)
{
    char b[64];
    snprintf(b, sizeof b, "%.2f", cov * 100.0 / len);
    return b;
}

std::string fmt_ratio(double sum, double len)
{
    char b[64];
    snprintf(b, sizeof b, "%.2f", sum * 1.0 / len);
    return b;
}

}  // namespace pdh
