// gpu_inflate_probe.cu -- standalone probe (NOT part of the product): inflate every BGZF block of a BAM on the GPU with
// pd_inflate.cuh's k_inflate_blocks, check the bytes against zlib on the host, count records with k_count_records, and print
// the device throughput.  First GPU item of the next round for the decode side.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ipandepth_b200/csrc tools/gpu_inflate_probe.cu -lz -o build/gpu_inflate_probe
//   build/gpu_inflate_probe file.bam [max_blocks]
#include "pd_inflate.cuh"
#include <cuda_runtime.h>
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace pdk;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: gpu_inflate_probe file.bam [max_blocks]\n"); return 2; }
    const size_t max_blocks = argc > 2 ? (size_t)atoll(argv[2]) : (size_t)-1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<uint8_t> data; std::vector<uint8_t> tmp(1 << 22); size_t k;
    while ((k = fread(tmp.data(), 1, tmp.size(), f)) > 0) data.insert(data.end(), tmp.begin(), tmp.begin() + (long)k);
    fclose(f);
    std::vector<BgzfBlock> blocks; uint64_t out_total = 0;
    for (size_t p = 0; p + 18 <= data.size() && blocks.size() < max_blocks;) {
        const uint8_t *h = data.data() + p;
        if (h[0] != 31 || h[1] != 139) break;
        const uint32_t xlen = h[10] | h[11] << 8; uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) { const uint8_t *e = h + 12 + x; const uint32_t sl = e[2] | e[3] << 8; if (e[0] == 'B' && e[1] == 'C' && sl == 2) { bsize = (e[4] | e[5] << 8) + 1; break; } x += 4 + sl; }
        if (!bsize || p + bsize > data.size()) break;
        const uint32_t isize = inf_le32(h + bsize - 4);
        if (isize) { blocks.push_back(BgzfBlock{p + 12 + xlen, bsize - xlen - 20, isize, out_total}); out_total += isize; }
        p += bsize;
    }
    printf("%zu blocks, %.1f MB compressed, %.1f MB inflated\n", blocks.size(), data.size() / 1e6, out_total / 1e6);
    uint8_t *d_in, *d_out; BgzfBlock *d_blk; int *d_status; uint32_t *d_n, *d_used;
    CK(cudaMalloc(&d_in, data.size())); CK(cudaMalloc(&d_out, out_total + 64)); CK(cudaMalloc(&d_blk, blocks.size() * sizeof(BgzfBlock)));
    CK(cudaMalloc(&d_status, 4)); CK(cudaMalloc(&d_n, blocks.size() * 4)); CK(cudaMalloc(&d_used, blocks.size() * 4));
    CK(cudaMemcpy(d_in, data.data(), data.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_blk, blocks.data(), blocks.size() * sizeof(BgzfBlock), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const unsigned grid = (unsigned)((blocks.size() + INFLATE_WPC - 1) / INFLATE_WPC);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        const int zero = 0; CK(cudaMemcpy(d_status, &zero, 4, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0));
        k_inflate_blocks<<<grid, INFLATE_WPC * 32>>>(d_in, d_out, d_blk, (uint32_t)blocks.size(), d_status);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms;
    }
    int status; CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
    printf("k_inflate_blocks: status %d, %.3f ms, %.2f GB/s inflated, %.2f GB/s compressed\n", status, best, out_total / best / 1e6, data.size() / best / 1e6);
    // bytes against zlib
    std::vector<uint8_t> got(out_total); CK(cudaMemcpy(got.data(), d_out, out_total, cudaMemcpyDeviceToHost));
    size_t bad = 0; std::vector<uint8_t> want;
    for (const BgzfBlock &b : blocks) {
        want.assign(b.out_len, 0);
        z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15);
        zs.next_in = data.data() + b.in_off; zs.avail_in = b.in_len; zs.next_out = want.data(); zs.avail_out = b.out_len;
        inflate(&zs, Z_FINISH); inflateEnd(&zs);
        if (memcmp(want.data(), got.data() + b.out_off, b.out_len)) bad++;
    }
    printf("blocks differing from zlib: %zu\n", bad);
    // records (skip the BAM header if it ends inside the first block)
    uint32_t skip = 0;
    if (out_total >= 12 && !memcmp(got.data(), "BAM\1", 4)) {
        size_t q = 8 + inf_le32(got.data() + 4); uint32_t n_ref = q + 4 <= out_total ? inf_le32(got.data() + q) : 0; q += 4;
        for (uint32_t r = 0; r < n_ref && q + 4 <= out_total; r++) q += 8 + inf_le32(got.data() + q);
        if (q <= blocks[0].out_len) skip = (uint32_t)q; else printf("header spans blocks: record count skipped\n");
    }
    CK(cudaEventRecord(e0));
    k_count_records<<<(unsigned)((blocks.size() + 127) / 128), 128>>>(d_out, d_blk, (uint32_t)blocks.size(), skip, d_n, d_used);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<uint32_t> n(blocks.size()), used(blocks.size());
    CK(cudaMemcpy(n.data(), d_n, blocks.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(used.data(), d_used, blocks.size() * 4, cudaMemcpyDeviceToHost));
    uint64_t recs = 0; size_t straddle = 0;
    for (size_t i = 0; i < blocks.size(); i++) { recs += n[i]; if (used[i] != blocks[i].out_len) straddle++; }
    printf("k_count_records: %.3f ms, %llu records, %zu blocks not ending on a record boundary\n", ms, (unsigned long long)recs, straddle);
    return bad || status ? 1 : 0;
}
