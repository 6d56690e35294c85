// Micro-test: may a K-major swizzled A-operand descriptor start at an arbitrary ROW of a TMA-written tile (not only at
// multiples of the 8-row swizzle atom)?  If the swizzle is a pure function of the shared-memory address bits, a start
// address of base + s*row_bytes reads rows s, s+1, ... correctly -- which would let a 3x3 convolution take its three
// horizontal taps from ONE halo tile (pixel shift = one row) instead of three TMA boxes.
// D[m][n] = sum_k A[m+s][k] * B[n][k] with B = identity  =>  D[m][n] == A[m+s][n].
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../bpbreid_b200/csrc/common.cuh"
using namespace bpb;
namespace bpb {
int make_tensor_map(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swizzle);
}

constexpr int kRowsA = 144;       // 128 + 16 spare rows

template <int KC>                 // K elements per row: 32 (64-byte rows, SW64) or 64 (128-byte rows, SW128)
__global__ void __launch_bounds__(128) shift_kernel(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB, float* out, int shift) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar, done_bar;
    __shared__ uint32_t tmem_base_s;
    constexpr int row_bytes = KC * 2;
    constexpr int N = KC;         // B = identity [N=KC][KC]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&done_bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<64>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + kRowsA * row_bytes;
    if (threadIdx.x == 0) {
        mbar_expect_tx_a(smem_u32(&full_bar), kRowsA * row_bytes + N * row_bytes);
        tma_load_2d_a(a_addr, &tmA, smem_u32(&full_bar), 0, 0);
        tma_load_2d_a(b_addr, &tmB, smem_u32(&full_bar), 0, 0);
        mbar_wait_a(smem_u32(&full_bar), 0);
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc(false, 128, N);
        const uint64_t da = make_kmajor_desc(a_addr + shift * row_bytes, row_bytes);
        const uint64_t db = make_kmajor_desc(b_addr, row_bytes);
        for (int k = 0; k < KC / 16; ++k) umma<false>(tmem, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
        umma_commit_a(smem_u32(&done_bar));
    }
    mbar_wait_a(smem_u32(&done_bar), 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem); }
}

template <int KC>
static void run() {
    constexpr int N = KC;
    std::vector<__nv_bfloat16> ha(kRowsA * KC), hb(N * KC);
    for (int r = 0; r < kRowsA; ++r)
        for (int k = 0; k < KC; ++k) ha[r * KC + k] = __float2bfloat16(static_cast<float>((r * 7 + k * 3) % 251) - 125.f);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < KC; ++k) hb[n * KC + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *da, *db; float* dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * N * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap tmA, tmB;
    uint32_t es[2] = {1, 1};
    { uint64_t dims[2] = {(uint64_t)KC, (uint64_t)kRowsA}; uint64_t st[1] = {(uint64_t)KC * 2}; uint32_t box[2] = {(uint32_t)KC, (uint32_t)kRowsA};
      if (make_tensor_map(&tmA, da, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, st, box, es, sw)) { printf("map A failed\n"); return; } }
    { uint64_t dims[2] = {(uint64_t)KC, (uint64_t)N}; uint64_t st[1] = {(uint64_t)KC * 2}; uint32_t box[2] = {(uint32_t)KC, (uint32_t)N};
      if (make_tensor_map(&tmB, db, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, st, box, es, sw)) { printf("map B failed\n"); return; } }
    const size_t smem = (kRowsA + N) * KC * 2 + 1024;
    cudaFuncSetAttribute(shift_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> hout(128 * N);
    for (int shift : {0, 1, 2, 3, 4, 5, 7, 8, 9, 16}) {
        shift_kernel<KC><<<1, 128, smem>>>(tmA, tmB, dout, shift);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n)
                if (hout[m * N + n] != __bfloat162float(ha[(m + shift) * KC + n])) ++bad;
        printf("row_bytes=%3d shift=%2d rows: %s (%d of %d elements differ)%s\n", KC * 2, shift, bad ? "MISMATCH" : "exact", bad,
               128 * N, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaFree(da); cudaFree(db); cudaFree(dout);
}

int main() {
    run<32>();
    run<64>();
    return 0;
}
