// gzwriter.hpp -- ordered output file whose chunks are formatted and compressed by a pool of threads.
// The reference writes its .stat.gz / .SiteDepth.gz through gzstream on one thread (PanDepth.cpp:4384 / :4280); here every
// chunk becomes one gzip member (RFC 1952: concatenated members are one valid stream; parity is on the decompressed bytes).
#pragma once
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#ifdef PD_HAVE_LIBDEFLATE
// libdeflate's public C API (the reference vendors lib/libdeflate.a without its header)
extern "C" {
struct libdeflate_compressor;
struct libdeflate_compressor *libdeflate_alloc_compressor(int compression_level);
size_t libdeflate_gzip_compress(struct libdeflate_compressor *c, const void *in, size_t in_nbytes, void *out, size_t out_nbytes_avail);
size_t libdeflate_gzip_compress_bound(struct libdeflate_compressor *c, size_t in_nbytes);
void libdeflate_free_compressor(struct libdeflate_compressor *c);
}
#endif

// gzip member of one text chunk (RFC 1952; concatenated members are one valid .gz stream -- parity is on decompressed bytes).
// libdeflate level 5 by default: on window rows it compresses slightly better than the zlib level 6 the reference's
// gzstream uses (ratio 7.0 vs 6.9) at a quarter of the CPU time (520 vs 1900-2000 ns per row); PANDEPTH_GZ_LEVEL=1..12
// overrides, PANDEPTH_GZ_ZLIB=1 forces the zlib path.
namespace pdh {

inline std::string gz_member(const std::string &in)
{
#ifdef PD_HAVE_LIBDEFLATE
    static const bool use_zlib = getenv("PANDEPTH_GZ_ZLIB") != nullptr;
    if (!use_zlib) {
        static const int level = [] { const char *e = getenv("PANDEPTH_GZ_LEVEL"); int l = e ? atoi(e) : 5; return l < 1 ? 1 : (l > 12 ? 12 : l); }();
        struct Holder { libdeflate_compressor *c = nullptr; ~Holder() { if (c) libdeflate_free_compressor(c); } };
        thread_local Holder h;
        if (!h.c) h.c = libdeflate_alloc_compressor(level);
        if (h.c) {
            std::string out;
            out.resize(libdeflate_gzip_compress_bound(h.c, in.size()));
            const size_t n = libdeflate_gzip_compress(h.c, in.data(), in.size(), &out[0], out.size());
            if (n) { out.resize(n); return out; }
        }
    }
#endif
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return std::string();
    std::string out;
    out.resize(deflateBound(&zs, (uLong)in.size()) + 64);
    zs.next_in = (Bytef *)in.data(); zs.avail_in = (uInt)in.size();
    zs.next_out = (Bytef *)&out[0]; zs.avail_out = (uInt)out.size();
    deflate(&zs, Z_FINISH);
    out.resize(out.size() - zs.avail_out);
    deflateEnd(&zs);
    return out;
}

struct GzWriter {
    FILE *fp = nullptr;
    int threads = 1;
    bool wrote = false;
    bool failed = false;                                     // a short write (disk full, quota): reported by close()
    std::string path;
    bool open(const std::string &p, int nthreads) { path = p; fp = fopen(p.c_str(), "wb"); threads = std::max(1, nthreads); return fp != nullptr; }
    void emit(const std::string &z) { if (z.empty()) return; if (fwrite(z.data(), 1, z.size(), fp) != z.size()) failed = true; wrote = true; }
    void put(const std::string &text) { if (text.empty()) return; emit(gz_member(text)); }
    // chunks 0..n-1, text of chunk k = produce(k); chunks must be < 2 GiB each
    template <typename F> void put_chunks(size_t n, F &&produce)
    {
        const size_t wave = (size_t)threads * 2;
        std::vector<std::string> z(wave);
        for (size_t base = 0; base < n; base += wave) {
            const size_t m = std::min(wave, n - base);
            std::atomic<size_t> next{0};
            auto work = [&] { for (;;) { size_t k = next.fetch_add(1); if (k >= m) break; std::string t = produce(base + k); z[k] = t.empty() ? std::string() : gz_member(t); } };
            std::vector<std::thread> th;
            const size_t nt = std::min<size_t>((size_t)threads, m);
            for (size_t i = 1; i < nt; i++) th.emplace_back(work);
            work();
            for (auto &t : th) t.join();
            for (size_t k = 0; k < m; k++) emit(z[k]);
        }
    }
    // a failed run: close and remove the partial file (a truncated .gz at the output path reads as a result)
    void abandon() { if (fp) { fclose(fp); fp = nullptr; } if (!path.empty()) ::remove(path.c_str()); }
    // false if anything could not be written; the caller reports it and exits non-zero
    bool close()
    {
        if (!fp) return !failed;
        if (!wrote) emit(gz_member(std::string()));
        if (fclose(fp) != 0) failed = true;
        fp = nullptr;
        if (failed) std::cerr << "Error: write failed (disk full?): " << path << std::endl;
        return !failed;
    }
};

}  // namespace pdh
