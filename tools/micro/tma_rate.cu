// Micro-benchmark: TMA tile-load throughput of one SM as a function of the box shape (row bytes, rows), ring depth
// and CTAs per SM, operands L2-resident.  One thread per CTA issues cp.async.bulk.tensor into a ring and waits for
// the box issued `depth` iterations earlier; prints bytes/cycle/SM and cycles per row.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../bpbreid_b200/csrc/common.cuh"
using namespace bpb;
namespace bpb {
int make_tensor_map(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swizzle);
}

__global__ void __launch_bounds__(128) tma_rate_kernel(const __grid_constant__ CUtensorMap tm, long long* out, int iters,
                                                      int depth, int box_bytes, int bw, int bh, int w_total,
                                                      int hn_total, int producers) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[64];
    if (threadIdx.x == 0) {
        for (int s = 0; s < 64; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int pw = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && pw < producers) {
        const int stage_bytes = (box_bytes + 1023) & ~1023;
        const uint32_t smem_a = smem_u32(smem) + pw * depth * stage_bytes, full_a = smem_u32(&full_bar[pw * 16]);
        int s = 0; uint32_t ph = 0;
        int hn = ((blockIdx.x * 4 + pw) * 7) % (hn_total - bh);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (it >= depth) mbar_wait_a(full_a + s * 8, ph ^ 1);     // box issued `depth` iterations ago (previous round)
            mbar_expect_tx_a(full_a + s * 8, box_bytes);
            tma_load_3d_a(smem_a + s * stage_bytes, &tm, full_a + s * 8, 0, 0, hn);
            hn += bh; if (hn + bh > hn_total) hn = 0;
            if (++s == depth) { s = 0; ph ^= 1; }
        }
        for (int k = 0; k < depth; ++k) {                              // drain
            mbar_wait_a(full_a + s * 8, ph ^ 1);
            if (++s == depth) { s = 0; ph ^= 1; }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && pw == 0) out[0] = t1 - t0;
    }
}

static void run(int row_bytes, int bw, int bh, int depth, int ctas_per_sm, int producers = 1) {
    const int c = row_bytes / 2, w_total = bw, hn_total = (32 << 20) / (row_bytes * bw);   // 32 MB tensor, L2-resident
    void* buf; cudaMalloc(&buf, 32 << 20); cudaMemset(buf, 0, 32 << 20);
    long long* d; cudaMalloc(&d, 8);
    CUtensorMap tm;
    uint64_t dims[3] = {(uint64_t)c, (uint64_t)w_total, (uint64_t)hn_total};
    uint64_t strides[2] = {(uint64_t)row_bytes, (uint64_t)row_bytes * w_total};
    uint32_t box[3] = {(uint32_t)c, (uint32_t)bw, (uint32_t)bh};
    uint32_t es[3] = {1, 1, 1};
    if (make_tensor_map(&tm, buf, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides, box, es,
                        row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) { printf("map failed\n"); return; }
    const int box_bytes = row_bytes * bw * bh;
    const int stage_bytes = (box_bytes + 1023) & ~1023;
    size_t smem = (size_t)depth * stage_bytes * producers + 1024;
    const size_t per_cta_cap = (ctas_per_sm == 1 ? 200 : 100) * 1024;
    if (smem > per_cta_cap) { cudaFree(buf); cudaFree(d); return; }
    if (ctas_per_sm == 1 && smem < 120 * 1024) smem = 120 * 1024;        // force one CTA per SM
    cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) tma_rate_kernel<<<148 * ctas_per_sm, 128, smem>>>(tm, d, iters, depth, box_bytes, bw, bh, w_total, hn_total, producers);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double per_box = double(cyc) / iters;
    printf("row=%3dB box=%3dx%-3d rows (%6d B) depth=%2d ctas/SM=%d producers=%d : %8.1f cyc/box/producer  %5.2f cyc/row/SM  %6.1f B/clk/SM%s\n",
           row_bytes, bw, bh, box_bytes, depth, ctas_per_sm, producers, per_box, per_box / (bw * bh) / (ctas_per_sm * producers),
           box_bytes / per_box * ctas_per_sm * producers, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(buf); cudaFree(d);
}

int main() {
    // one CTA per SM, 1 / 2 / 4 producer warps, small and large boxes
    for (int rb : {64, 128})
        for (int pr : {1, 2, 4})
            for (int rows : {32, 128, 192}) run(rb, 32, rows / 32, 4, 1, pr);
    for (int pr : {1, 2, 4}) run(128, 256, 1, 2, 1, pr);
    for (int pr : {1, 2}) run(64, 32, 6, 3, 2, pr);
    return 0;
}
