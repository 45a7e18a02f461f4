// gpu_decode_probe.cu -- standalone probe (NOT part of the product): the block-parallel BGZF -> records pipeline of
// pd_decode.hpp with a CUDA backend (inflate: one block per warp; count / emit: one block per thread), checked against the
// same pipeline on a host backend, with device timings per stage.  Next round's first step for the decode side.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ipandepth_b200/csrc tools/gpu_decode_probe.cu -lz -o build/gpu_decode_probe
//   build/gpu_decode_probe file.bam [blocks_per_chunk]
#include "pd_decode.hpp"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace pdk;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

__global__ void __launch_bounds__(INFLATE_WPC * 32) k_inflate_chunk(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const ChunkBlock *__restrict__ blk,
                                                                    const uint64_t *__restrict__ out_off, uint32_t n, int *status)
{
    __shared__ InflateTables tabs[INFLATE_WPC];
    const uint32_t b = blockIdx.x * INFLATE_WPC + (threadIdx.x >> 5);
    if (b >= n || (threadIdx.x & 31) != 0) return;
    const ChunkBlock k = blk[b];
    const int r = k.out_len ? inflate_raw(in + k.in_off, k.in_len, out + out_off[b], k.out_len, tabs[threadIdx.x >> 5]) : 0;
    if (r != (int)k.out_len) atomicMin(status, r < 0 ? r : INF_ERR_OUTPUT);
}
__global__ void k_count_chunk(const uint8_t *__restrict__ data, const ChunkBlock *__restrict__ blk, const uint64_t *__restrict__ out_off, uint32_t n, uint32_t first_skip,
                              int32_t n_targets, uint32_t xflags, int32_t min_mapq, BlockCount *cnt)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint32_t sk = b == 0 ? first_skip : 0;
    cnt[b] = count_block(data + out_off[b] + sk, blk[b].out_len - sk, n_targets, xflags, min_mapq);
}
__global__ void k_emit_chunk(const uint8_t *__restrict__ data, const ChunkBlock *__restrict__ blk, const uint64_t *__restrict__ out_off, uint32_t n, uint32_t first_skip,
                             const uint32_t *__restrict__ r0, const uint32_t *__restrict__ c0, int32_t n_targets, uint32_t xflags, int32_t min_mapq,
                             int32_t *tid, int32_t *pos, uint32_t *cigar_off, uint32_t *cigar)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint32_t sk = b == 0 ? first_skip : 0;
    emit_block(data + out_off[b] + sk, blk[b].out_len - sk, n_targets, xflags, min_mapq, r0[b], c0[b], tid, pos, cigar_off, cigar);
}

struct CudaBackend {
    uint8_t *d_in = nullptr, *inflated = nullptr; size_t in_cap = 0, out_cap = 0;
    ChunkBlock *d_blk = nullptr; uint64_t *d_off = nullptr; uint32_t *d_r0 = nullptr, *d_c0 = nullptr; BlockCount *d_cnt = nullptr; size_t blk_cap = 0;
    int32_t *d_tid = nullptr, *d_pos = nullptr; uint32_t *d_coff = nullptr, *d_cig = nullptr; size_t r_cap = 0, c_cap = 0;
    int *d_status = nullptr;
    std::vector<BlockCount> counts_v; BlockCount *counts_host = nullptr;
    cudaEvent_t ev[8]; float ms_h2d = 0, ms_inflate = 0, ms_count = 0, ms_emit = 0;
    CudaBackend() { for (auto &e : ev) CK(cudaEventCreate(&e)); CK(cudaMalloc(&d_status, 4)); }
    void blocks_room(uint32_t n)
    {
        if (n <= blk_cap) return;
        cudaFree(d_blk); cudaFree(d_off); cudaFree(d_r0); cudaFree(d_c0); cudaFree(d_cnt);
        blk_cap = n + n / 2 + 64;
        CK(cudaMalloc(&d_blk, blk_cap * sizeof(ChunkBlock))); CK(cudaMalloc(&d_off, blk_cap * 8)); CK(cudaMalloc(&d_r0, blk_cap * 4)); CK(cudaMalloc(&d_c0, blk_cap * 4)); CK(cudaMalloc(&d_cnt, blk_cap * sizeof(BlockCount)));
    }
    int inflate(const uint8_t *in, size_t, const ChunkBlock *blk, const uint64_t *out_off, uint32_t n, uint64_t out_total)
    {
        // only the chunk's bytes travel: [first block's payload start, last block's payload end)
        const uint64_t a = blk[0].in_off, z = blk[n - 1].in_off + blk[n - 1].in_len;
        if (z - a > in_cap) { cudaFree(d_in); in_cap = (z - a) * 2; CK(cudaMalloc(&d_in, in_cap)); }
        if (out_total + 8 > out_cap) { cudaFree(inflated); out_cap = out_total * 2 + 8; CK(cudaMalloc(&inflated, out_cap)); }
        blocks_room(n);
        std::vector<ChunkBlock> rel(blk, blk + n);
        for (auto &b : rel) b.in_off -= a;
        const int zero = 0;
        CK(cudaEventRecord(ev[0]));
        CK(cudaMemcpyAsync(d_in, in + a, z - a, cudaMemcpyHostToDevice));
        CK(cudaMemcpyAsync(d_blk, rel.data(), n * sizeof(ChunkBlock), cudaMemcpyHostToDevice));
        CK(cudaMemcpyAsync(d_off, out_off, n * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpyAsync(d_status, &zero, 4, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(ev[1]));
        k_inflate_chunk<<<(n + INFLATE_WPC - 1) / INFLATE_WPC, INFLATE_WPC * 32>>>(d_in, inflated, d_blk, d_off, n, d_status);
        CK(cudaEventRecord(ev[2]));
        int status = 0; CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
        float t; CK(cudaEventElapsedTime(&t, ev[0], ev[1])); ms_h2d += t; CK(cudaEventElapsedTime(&t, ev[1], ev[2])); ms_inflate += t;
        return status;
    }
    void count(const ChunkBlock *, const uint64_t *, uint32_t n, uint32_t first_skip, int32_t n_targets, uint32_t xflags, int32_t min_mapq)
    {
        CK(cudaEventRecord(ev[3]));
        k_count_chunk<<<(n + 127) / 128, 128>>>(inflated, d_blk, d_off, n, first_skip, n_targets, xflags, min_mapq, d_cnt);
        CK(cudaEventRecord(ev[4]));
        counts_v.resize(n); counts_host = counts_v.data();
        CK(cudaMemcpy(counts_v.data(), d_cnt, n * sizeof(BlockCount), cudaMemcpyDeviceToHost));
        float t; CK(cudaEventElapsedTime(&t, ev[3], ev[4])); ms_count += t;
    }
    void emit(const ChunkBlock *, const uint64_t *, uint32_t n, uint32_t first_skip, const uint32_t *r0, const uint32_t *c0, uint32_t n_reads, uint32_t n_cigar,
              int32_t n_targets, uint32_t xflags, int32_t min_mapq)
    {
        if (n_reads + 1 > r_cap) { cudaFree(d_tid); cudaFree(d_pos); cudaFree(d_coff); r_cap = (size_t)n_reads * 2 + 64; CK(cudaMalloc(&d_tid, r_cap * 4)); CK(cudaMalloc(&d_pos, r_cap * 4)); CK(cudaMalloc(&d_coff, r_cap * 4)); }
        if (n_cigar > c_cap) { cudaFree(d_cig); c_cap = (size_t)n_cigar * 2 + 64; CK(cudaMalloc(&d_cig, c_cap * 4)); }
        CK(cudaMemcpyAsync(d_r0, r0, n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpyAsync(d_c0, c0, n * 4, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(ev[5]));
        k_emit_chunk<<<(n + 127) / 128, 128>>>(inflated, d_blk, d_off, n, first_skip, d_r0, d_c0, n_targets, xflags, min_mapq, d_tid, d_pos, d_coff, d_cig);
        CK(cudaEventRecord(ev[6]));
        CK(cudaMemcpyAsync(d_coff + n_reads, &n_cigar, 4, cudaMemcpyHostToDevice));
        CK(cudaEventSynchronize(ev[6]));
        float t; CK(cudaEventElapsedTime(&t, ev[5], ev[6])); ms_emit += t;
    }
    int32_t tid_at(uint32_t r) { int32_t v; CK(cudaMemcpy(&v, d_tid + r, 4, cudaMemcpyDeviceToHost)); return v; }
};

static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: gpu_decode_probe file.bam [blocks_per_chunk]\n"); return 2; }
    const size_t per_chunk = argc > 2 ? (size_t)atoll(argv[2]) : 16384;
    FILE *f = fopen(argv[1], "rb"); if (!f) { perror(argv[1]); return 2; }
    std::vector<uint8_t> data; std::vector<uint8_t> tmp(1 << 22); size_t k;
    while ((k = fread(tmp.data(), 1, tmp.size(), f)) > 0) data.insert(data.end(), tmp.begin(), tmp.begin() + (long)k);
    fclose(f);
    std::vector<ChunkBlock> blocks; uint64_t out_total = 0;
    for (size_t p = 0; p + 18 <= data.size();) {
        const uint8_t *h = data.data() + p;
        if (h[0] != 31 || h[1] != 139) break;
        const uint32_t xlen = h[10] | h[11] << 8; uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) { const uint8_t *e = h + 12 + x; const uint32_t sl = e[2] | e[3] << 8; if (e[0] == 'B' && e[1] == 'C' && sl == 2) { bsize = (e[4] | e[5] << 8) + 1; break; } x += 4 + sl; }
        if (!bsize || p + bsize > data.size()) break;
        const uint32_t isize = inf_le32(h + bsize - 4);
        if (isize) { blocks.push_back(ChunkBlock{p + 12 + xlen, bsize - xlen - 20, isize}); out_total += isize; }
        p += bsize;
    }
    // header length: inflate leading blocks on the host until the header is complete
    std::vector<uint8_t> headbuf; size_t hb = 0; size_t q = 0; uint32_t n_ref = 0; static InflateTables T;
    for (;;) {
        if (hb >= blocks.size()) { fprintf(stderr, "no records\n"); return 2; }
        const size_t o = headbuf.size(); headbuf.resize(o + blocks[hb].out_len);
        inflate_raw(data.data() + blocks[hb].in_off, blocks[hb].in_len, headbuf.data() + o, blocks[hb].out_len, T); hb++;
        if (headbuf.size() < 12) continue;
        q = 8 + inf_le32(headbuf.data() + 4); if (headbuf.size() < q + 4) continue;
        n_ref = inf_le32(headbuf.data() + q); q += 4; bool ok = true;
        for (uint32_t r = 0; r < n_ref; r++) { if (headbuf.size() < q + 4) { ok = false; break; } q += 8 + inf_le32(headbuf.data() + q); }
        if (ok && headbuf.size() >= q) break;
    }
    size_t b0 = 0; uint64_t acc = 0; while (b0 < blocks.size() && acc + blocks[b0].out_len <= q) { acc += blocks[b0].out_len; b0++; }
    const uint32_t skip0 = (uint32_t)(q - acc);
    printf("%zu blocks, %.1f MB compressed, %.1f MB inflated, %u contigs\n", blocks.size(), data.size() / 1e6, out_total / 1e6, n_ref);
    CudaBackend be;
    uint64_t n_total = 0, digest = 0; size_t n_runs = 0, n_sorted_runs = 0;
    std::vector<int32_t> tid, pos; std::vector<uint32_t> coff, cig;
    for (size_t b = b0; b < blocks.size(); b += per_chunk) {
        std::vector<ChunkBlock> chunk(blocks.begin() + (long)b, blocks.begin() + (long)std::min(blocks.size(), b + per_chunk));
        std::vector<ContigRun> runs; uint32_t n_reads = 0, n_cigar = 0;
        const int rc = decode_chunk(be, data.data(), data.size(), chunk, b == b0 ? skip0 : 0, (int32_t)n_ref, 1796u, -1, runs, n_reads, n_cigar);
        if (rc != DEC_OK) { printf("decode_chunk returned %d at block %zu\n", rc, b); return 1; }
        n_runs += runs.size(); for (auto &r : runs) n_sorted_runs += r.sorted;
        tid.resize(n_reads); pos.resize(n_reads); coff.resize(n_reads + 1); cig.resize(n_cigar);
        CK(cudaMemcpy(tid.data(), be.d_tid, (size_t)n_reads * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(pos.data(), be.d_pos, (size_t)n_reads * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(coff.data(), be.d_coff, ((size_t)n_reads + 1) * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cig.data(), be.d_cig, (size_t)n_cigar * 4, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < n_reads; i++) { uint64_t h = mix(((uint64_t)(uint32_t)tid[i] << 32) | (uint32_t)pos[i]); for (uint32_t c = coff[i]; c < coff[i + 1]; c++) h = mix(h ^ cig[c]); digest += h; }
        n_total += n_reads;
    }
    printf("GPU pipeline: %llu records, digest %llu, %zu contig runs (%zu sorted)\n", (unsigned long long)n_total, (unsigned long long)digest, n_runs, n_sorted_runs);
    printf("device ms: h2d %.2f  inflate %.2f (%.1f GB/s inflated)  count %.2f  emit %.2f   => %.1f Mreads/s through inflate+count+emit\n",
           be.ms_h2d, be.ms_inflate, out_total / be.ms_inflate / 1e6, be.ms_count, be.ms_emit, n_total / (be.ms_inflate + be.ms_count + be.ms_emit) / 1e3);
    printf("compare the digest with:  PANDEPTH_DRY=1 pandepth_b200/bin/pandepth -i %s -o /tmp/x -s   (same filter: -x 1796, no -q)\n", argv[1]);
    return 0;
}
