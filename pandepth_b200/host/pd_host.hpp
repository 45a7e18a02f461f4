// pd_host.hpp -- C++ host around the CUDA engine: the `pandepth` command line (same flags, same output files as the
// reference), the region model, the htslib feeder and the report writer.  New code; behaviour follows
// /root/reference/src/PanDepth.cpp (cited per function as :line).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace pdh {

// In3str1v (include/DataClass.h:22-59) with the defaults main() applies (:3444).
struct Options {
    std::string input;                 // -i (first file)
    std::vector<std::string> inputs;   // every -i / every line of a #.list file (:99-115); more than one = multi-file mode
    std::string out_arg;               // -o as given
    std::string gff;                   // -g
    std::string bed;                   // -b
    std::string feature = "CDS";       // -f
    std::string reference;             // -r (CRAM decode)
    bool site_output = false;          // -a
    bool use_index = true;             // cleared by the undocumented -s (:223-226)
    uint32_t xflags = 1796;            // -x
    int min_mapq = -1;                 // -q  (InInt preset to -1 at :3444)
    int min_dep = 1;                   // -d  (clamped >= 1, :233)
    int win = 0;                       // -w  (clamped >= 1 when given, :211-215)
    int threads = 3;                   // -t
    int region_mode = 0;               // InInt2: 0 chr, 1 gff, 2 gtf, 3 bed3, 4 bed4, 5 window>=150, 6 window<150
    long long tile_win = 0;            // mode 5: width of the -w pseudo-gene tiles, which stay implicit (see n_tiles)
    bool gc = false;                   // -c (not supported by this build; reported)
    bool help = false;
};

// Returns 0 = run, 1 = exit quietly with status 0 (help / usage error, like the reference's `return 0`).
int parse_args(int argc, char **argv, Options &o);
void print_help();

// GeneInfo (include/DataClass.h:62-80)
struct Gene {
    int start = 0, end = 0;
    uint64_t length = 0;
    std::vector<std::pair<int, int>> cds;   // 1-based closed
    int64_t cover = 0;                      // int GeneCover in the reference
    uint64_t depth = 0;
};
using GeneMap = std::map<int, std::map<std::string, Gene>>;     // tid -> GeneID -> Gene   (:3543)

struct Target { std::string name; uint32_t len; };

// plain or gzip text, line by line (the reference reads these through igzstream)
struct LineReader {
    gzFile fp = nullptr;
    std::string buf;
    explicit LineReader(const std::string &path) { fp = gzopen(path.c_str(), "rb"); }
    ~LineReader() { if (fp) gzclose(fp); }
    bool ok() const { return fp != nullptr; }
    bool next(std::string &line)
    {
        line.clear();
        char tmp[1 << 16];
        bool got = false;
        while (gzgets(fp, tmp, sizeof tmp)) {
            got = true;
            size_t n = strlen(tmp);
            if (n && tmp[n - 1] == '\n') { line.append(tmp, n - 1); return true; }
            line.append(tmp, n);
        }
        return got;
    }
};

// Pseudo-gene tiles of a contig (:3997-4044): tile k is 1-based closed [1 + k*W, min(k*W + W, len)] and exists while
// k*W + 2 <= len (`End` is pre-set to 2 and bumped by 2 after each tile) -- contigs shorter than 2 bp have none and a
// final 1-bp remainder (len % W == 1) is dropped.
inline long long n_tiles(long long len, long long W) { return len >= 2 ? (len - 2) / W + 1 : 0; }

// :3547-3905 (GFF/GTF/BED parsers), :3975-4051 (pseudo-gene tiling).  Sets o.region_mode.
// Mode 5 (-w >= 150 without -g/-b) can mean 2x10^7 tiles on a 3 Gb genome: they are NOT materialised as map entries (the
// reference's std::map costs it tens of seconds there); `genes` then holds 10 Mb tiles, only to say which contigs exist, and
// o.tile_win carries W.  The report enumerates tile k arithmetically; every tile is handed to StatChrDepthWin exactly once
// (windows of ProDealChrBambai start and end on tile boundaries, :695-784), which tests/test_regions_host_cpu.py checks
// against the generic hand-over count of the CPU model.
bool build_regions(Options &o, const std::vector<Target> &targets, GeneMap &genes, std::string &err);
// :3919-3972
std::vector<std::pair<int, int>> merge_regions(const std::map<std::string, Gene> &genes);
// How often ProDealChrBambai's windowing (:695-784) hands each gene to StatChrDepthWin (:299-308): 1, or 0 for a gene that
// starts on/after the contig-clipped window end.
std::map<std::string, int> indexed_window_counts(const std::vector<std::pair<int, int>> &regions,
                                                 const std::map<std::string, Gene> &genes, int chr_len);

// output naming :4057-4090
std::string output_prefix(const std::string &out_arg);
const char *stat_suffix(int mode);
const char *stat_header(int mode);

// fixed/setprecision(2) of cov*100.0/len and sum*1.0/len (:5023-5024)
std::string fmt_ratio_pct(double cov, double len);
std::string fmt_ratio(double sum, double len);

}  // namespace pdh
