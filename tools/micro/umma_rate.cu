// Micro-benchmark: how fast can one thread issue tcgen05.mma (SS, bf16, M=128) for N = 32..256, K=16,
// with static operands in shared memory?  Prints SM cycles per MMA for 1 and 2 CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../bpbreid_b200/csrc/common.cuh"
using namespace bpb;
namespace bpb { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; }
int make_tensor_map(CUtensorMap*, const void*, CUtensorMapDataType, int, const uint64_t*, const uint64_t*, const uint32_t*, const uint32_t*, CUtensorMapSwizzle) { return 1; } }

template <int N>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int iters, int unroll_mode) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(false, 128, N);
        const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 128 * 128;
        const uint64_t da = make_kmajor_desc(a_addr, 128), db = make_kmajor_desc(b_addr, 128);
        const uint32_t bar_a = smem_u32(&bar);
        long long t0 = clock64();
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma<false>(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                if (unroll_mode == 1) umma_commit_a(bar_a), mbar_wait_a(bar_a, it & 1);   // commit+wait every 4 MMAs
            }
            if (unroll_mode == 0) { umma_commit_a(bar_a); mbar_wait_a(bar_a, 0); }
        }
        __syncwarp();
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

template <int N>
void run(int ctas_per_sm, int mode) {
    long long* d; cudaMalloc(&d, 8);
    auto k = rate_kernel<N>;
    size_t smem = (128 + 256) * 128 + 1024 + (ctas_per_sm == 1 ? 100 * 1024 : 0);   // pad to force 1 CTA/SM
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    k<<<148 * ctas_per_sm, 128, smem>>>(d, iters, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d ctas/SM=%d mode=%s : %7.1f cycles per MMA (M128 x N x K16 bf16; math floor %d)%s\n", N, ctas_per_sm,
           mode ? "commit+wait/4" : "free-run", double(c) / (iters * 4.0), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int mode = 0; mode < 2; ++mode)
        for (int c = 1; c <= 2; ++c) { run<32>(c, mode); run<64>(c, mode); run<128>(c, mode); run<256>(c, mode); }
    return 0;
}
