// Micro-benchmark: how fast can one thread issue tcgen05.mma (SS, bf16, M=128) for N = 32..256, K=16,
// with static operands in shared memory?  Prints SM cycles per MMA for 1 and 2 CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../bpbreid_b200/csrc/common.cuh"
using namespace bpb;
namespace bpb { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; }
int make_tensor_map(CUtensorMap*, const void*, CUtensorMapDataType, int, const uint64_t*, const uint64_t*, const uint32_t*, const uint32_t*, CUtensorMapSwizzle) { return 1; } }

template <int N>
__global__ void __launch_bounds__(160) rate_kernel(long long* out, int iters, int ndst, int row_bytes, int issuers, int tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc_n(&tmem_base_s, tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp >= 1 && warp <= issuers) {
        constexpr uint32_t idesc = make_idesc(false, 128, N);
        const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 128 * 128;
        const uint64_t da = make_kmajor_desc(a_addr, row_bytes), db = make_kmajor_desc(b_addr, row_bytes);
        const uint32_t bar_a = smem_u32(&bar[warp - 1]);
        const uint32_t tmem_w = tmem + (warp - 1) * N;      // private accumulators per issuing warp
        const int kper = row_bytes / 32;
        long long t0 = clock64();
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                for (int k = 0; k < 4; ++k) {
                    const uint32_t d = tmem_w + ((it * 4 + k) % ndst) * N;
                    umma<false>(d, da + 2 * (k % kper), db + 2 * (k % kper), idesc, 1u);
                }
            }
            umma_commit_a(bar_a); mbar_wait_a(bar_a, 0);
        }
        __syncwarp();
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0 && warp == 1) out[0] = (t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc_n(tmem, tmem_cols); }
}

template <int N>
void run(int ndst, int row_bytes, int issuers = 1, int ctas = 1) {
    long long* d; cudaMalloc(&d, 8);
    auto k = rate_kernel<N>;
    size_t smem = (128 + 256) * 128 + 1024 + (ctas == 1 ? 100 * 1024 : 0);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    k<<<148 * ctas, 160, smem>>>(d, iters, ndst, row_bytes, issuers, ctas == 1 ? 512 : 256);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d row_bytes=%3d accumulators=%d issuers/CTA=%d CTAs/SM=%d : %7.1f cycles per MMA per issuer, %6.1f aggregate per SM (math floor %d)%s\n", N, row_bytes, ndst, issuers, ctas,
           double(c) / (iters * 4.0), double(c) / (iters * 4.0) / (issuers * ctas), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int is : {1, 2, 3})
        for (int ct : {1, 2}) { run<32>(1, 64, is, ct); run<64>(1, 128, is, ct); }
    run<32>(1, 64, 4, 1);
    return 0;
}
