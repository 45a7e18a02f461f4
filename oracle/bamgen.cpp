/*
 * bamgen.cpp -- deterministic synthetic BAM generator (TEST / BENCH INFRASTRUCTURE ONLY).
 *
 * Builds against the htslib the reference vendors (headers under /root/reference/include,
 * lib/libhts.a), so it lives under oracle/ and its binary goes to oracle/_ref/bamgen.
 * Shapes follow SURVEY.md section 8(d): coordinate-sorted, @HD SO:coordinate, .bai via sam_index_build3,
 * BGZF level 1, starts uniform per contig, std::mt19937_64 seeded.
 *
 *   bamgen -o out.bam -c name:len[,name:len...] -n reads [-s seed] [-l readlen] [-m uniform|mixed]
 *   bamgen -I in.sam -o out.bam      (convert a SAM fixture to BAM + .bai)
 *          [-q] (emit random SEQ+QUAL, flavour B; default SEQ='*', flavour A) [-t threads] [-S] (write SAM text)
 *   an output name ending in .cram is written as CRAM 3.x without a reference (no_ref) + .crai
 *          [-X] build a .csi index (min_shift 14) instead of the .bai
 *          [-O] mixed mode: no read may run past its contig's end (the reference's multi-file path has no slack behind its
 *               arrays: overhanging reads corrupt its heap)
 *
 * mode uniform : every read flag 0, mapq 60, CIGAR <readlen>M.
 * mode mixed   : random CIGARs over MIDNSHP=X, random flags (incl. UNMAP/SECONDARY/QCFAIL/DUP/SUPP bits),
 *                random mapq -- for differential testing against the reference binary.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <random>
#include <algorithm>
#include "sam.h"
#include "hts.h"

struct Contig { std::string name; int64_t len; };

static void die(const char *m) { fprintf(stderr, "bamgen: %s\n", m); exit(2); }

static bool is_cram_name(const std::string &p) { return p.size() > 5 && p.compare(p.size() - 5, 5, ".cram") == 0; }

static htsFile *open_out(const std::string &path, bool samtext)
{
    const bool cram = is_cram_name(path);
    htsFile *fp = hts_open(path.c_str(), samtext ? "w" : (cram ? "wc" : "wb1"));
    if (fp && cram && hts_set_opt(fp, CRAM_OPT_NO_REF, 1) != 0) die("cannot set no_ref");
    return fp;
}

static void make_mixed_cigar(std::mt19937_64 &rng, int readlen, std::vector<uint32_t> &cig, int &qlen, int &rlen)
{
    // a few ops; query-consuming ops sum to ~readlen
    cig.clear(); qlen = 0; rlen = 0;
    auto push = [&](int op, int n) { if (n > 0) cig.push_back((uint32_t)n << 4 | (uint32_t)op); };
    int r = rng() % 100;
    if (r < 8) { push(BAM_CHARD_CLIP, 1 + rng() % 30); }
    if (r % 5 == 0) { int n = 1 + rng() % 40; push(BAM_CSOFT_CLIP, n); qlen += n; }
    int nblocks = 1 + rng() % 4;
    for (int b = 0; b < nblocks; b++) {
        int n = 1 + rng() % (readlen / nblocks + 1);
        int mop = (rng() % 10 == 0) ? BAM_CEQUAL : (rng() % 10 == 1 ? BAM_CDIFF : BAM_CMATCH);
        push(mop, n); qlen += n; rlen += n;
        if (b + 1 < nblocks) {
            int k = rng() % 6;
            if (k == 0) { int d = 1 + rng() % 20; push(BAM_CDEL, d); rlen += d; }
            else if (k == 1) { int d = 1 + rng() % 400; push(BAM_CREF_SKIP, d); rlen += d; }
            else if (k == 2) { int d = 1 + rng() % 10; push(BAM_CINS, d); qlen += d; }
            else if (k == 3) { push(BAM_CPAD, 1 + rng() % 3); }
        }
    }
    if (r % 7 == 0) { int n = 1 + rng() % 40; push(BAM_CSOFT_CLIP, n); qlen += n; }
}

int main(int argc, char **argv)
{
    std::string out, cspec, mode = "uniform", insam;
    int64_t nreads = 0; uint64_t seed = 12345; int readlen = 150; bool withseq = false; int threads = 4; bool samtext = false; int min_shift = 0; bool no_overhang = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&]() -> const char * { if (i + 1 >= argc) die("missing argument"); return argv[++i]; };
        if (a == "-o") out = need();
        else if (a == "-c") cspec = need();
        else if (a == "-n") nreads = atoll(need());
        else if (a == "-s") seed = strtoull(need(), 0, 10);
        else if (a == "-l") readlen = atoi(need());
        else if (a == "-m") mode = need();
        else if (a == "-q") withseq = true;
        else if (a == "-t") threads = atoi(need());
        else if (a == "-S") samtext = true;
        else if (a == "-I") insam = need();
        else if (a == "-X") min_shift = 14;
        else if (a == "-O") no_overhang = true;
        else die("unknown option");
    }
    if (!insam.empty()) {                                   // SAM text -> BAM (+ .bai): fixture conversion
        if (out.empty()) die("-I needs -o");
        htsFile *in = hts_open(insam.c_str(), "r");
        if (!in) die("cannot open input SAM");
        sam_hdr_t *h = sam_hdr_read(in);
        if (!h) die("cannot read SAM header");
        htsFile *o = open_out(out, false);
        if (!o || sam_hdr_write(o, h) < 0) die("cannot write BAM");
        bam1_t *r = bam_init1();
        while (sam_read1(in, h, r) >= 0) if (sam_write1(o, h, r) < 0) die("write failed");
        bam_destroy1(r); hts_close(in);
        if (hts_close(o) < 0) die("close failed");
        sam_hdr_destroy(h);
        if (sam_index_build3(out.c_str(), NULL, min_shift, 1) < 0) die("index build failed");
        return 0;
    }
    if (out.empty() || cspec.empty() || nreads <= 0) die("usage: bamgen -o out.bam -c name:len,... -n reads [-s seed] [-l len] [-m uniform|mixed] [-q] [-t thr] [-S]");
    std::vector<Contig> contigs;
    {
        size_t p = 0;
        while (p < cspec.size()) {
            size_t q = cspec.find(',', p); if (q == std::string::npos) q = cspec.size();
            std::string tok = cspec.substr(p, q - p);
            size_t c = tok.rfind(':'); if (c == std::string::npos) die("bad contig spec");
            contigs.push_back({tok.substr(0, c), atoll(tok.c_str() + c + 1)});
            p = q + 1;
        }
    }
    int64_t G = 0; for (auto &c : contigs) G += c.len;
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (auto &c : contigs) text += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.len) + "\n";
    sam_hdr_t *hdr = sam_hdr_parse(text.size(), text.c_str());
    if (!hdr) die("header parse failed");

    htsFile *fp = open_out(out, samtext);
    if (!fp) die("cannot open output");
    if (threads > 1 && !samtext) hts_set_threads(fp, threads);
    if (sam_hdr_write(fp, hdr) < 0) die("header write failed");

    std::mt19937_64 rng(seed);
    bam1_t *b = bam_init1();
    const bool mixed = (mode == "mixed");
    std::vector<uint32_t> cig;
    int64_t assigned = 0;
    for (size_t ci = 0; ci < contigs.size(); ci++) {
        int64_t len = contigs[ci].len;
        int64_t n = (ci + 1 == contigs.size()) ? nreads - assigned : (int64_t)((long double)nreads * len / G);
        assigned += n;
        if (len < 1) continue;
        int64_t maxstart = len - readlen; if (maxstart < 0) maxstart = 0;
        std::vector<int32_t> starts((size_t)n);
        for (auto &s : starts) s = (int32_t)(rng() % (uint64_t)(maxstart + 1));
        std::sort(starts.begin(), starts.end());
        for (int64_t i = 0; i < n; i++) {
            int qlen = readlen, rlen = readlen; uint16_t flag = 0; uint8_t mapq = 60;
            if (mixed) {
                make_mixed_cigar(rng, readlen, cig, qlen, rlen);
                if ((int64_t)starts[i] + rlen > len + (no_overhang ? 0 : 400)) {            // keep overhang inside the reference's +500 slack (:4132)
                    int m = (int)std::min<int64_t>(readlen, std::max<int64_t>(1, len - starts[i]));
                    cig.assign(1, (uint32_t)m << 4 | BAM_CMATCH); qlen = rlen = m;
                }
                uint64_t r = rng();
                static const uint16_t bits[] = {0x1, 0x2, 0x10, 0x20, 0x40, 0x80, 0x100, 0x200, 0x400, 0x800, 0x4};
                for (int k = 0; k < 11; k++) if (((r >> (k * 4)) & 0xf) == 0) flag |= bits[k];
                mapq = (uint8_t)((r >> 48) % 61);
            } else {
                cig.assign(1, (uint32_t)readlen << 4 | BAM_CMATCH);
            }
            int l_seq = withseq ? qlen : 0;
            const char qn[] = "r"; int l_qname = 4;                  // "r\0\0\0" padded to 4
            size_t dl = l_qname + cig.size() * 4 + (size_t)(l_seq + 1) / 2 + l_seq;
            if (b->m_data < dl) { b->m_data = (uint32_t)(dl + 64); b->data = (uint8_t *)realloc(b->data, b->m_data); }
            b->l_data = (int)dl;
            memset(b->data, 0, l_qname); memcpy(b->data, qn, 1);
            memcpy(b->data + l_qname, cig.data(), cig.size() * 4);
            uint8_t *sq = b->data + l_qname + cig.size() * 4;
            for (int k = 0; k < (l_seq + 1) / 2; k++) { static const uint8_t nt[4] = {1, 2, 4, 8}; uint64_t r = rng(); sq[k] = (uint8_t)(nt[r & 3] << 4 | nt[(r >> 2) & 3]); }
            uint8_t *ql = sq + (l_seq + 1) / 2;
            for (int k = 0; k < l_seq; k++) ql[k] = (uint8_t)(2 + (rng() % 39));
            b->core.tid = (int32_t)ci; b->core.pos = starts[i]; b->core.qual = mapq; b->core.flag = flag;
            b->core.l_qname = (uint16_t)l_qname; b->core.l_extranul = 2; b->core.n_cigar = (uint32_t)cig.size();
            b->core.l_qseq = l_seq; b->core.mtid = -1; b->core.mpos = -1; b->core.isize = 0;
            b->core.bin = (uint16_t)hts_reg2bin(starts[i], starts[i] + (rlen > 0 ? rlen : 1), 14, 5);
            if (sam_write1(fp, hdr, b) < 0) die("write failed");
        }
    }
    bam_destroy1(b);
    if (hts_close(fp) < 0) die("close failed");
    sam_hdr_destroy(hdr);
    if (!samtext) { if (sam_index_build3(out.c_str(), NULL, min_shift, threads) < 0) die("index build failed"); }
    return 0;
}
