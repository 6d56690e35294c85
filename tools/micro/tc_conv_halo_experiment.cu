// 3x3 stride-1 convolution from ONE halo tile per output tile (tcgen05, sm_100a).
//
// The row-shift schedule of tc_gemm.cu still fetches every input pixel ~4.5 times from L2 (three horizontally shifted
// TMA boxes, each with a two-row halo), and the C=32 / C=64 layers of HRNet turned out to be bound by exactly that
// L2 -> SM stream (227 MB per 67 MB layer at ~7.3 TB/s, profiles/ncu_full_r1.md).  Here the (R+2) x (W+2) zero-padded
// halo of an R-row output tile is loaded ONCE per channel chunk (TMA out-of-bounds fill provides all four borders) and
// all nine taps read it through descriptor start offsets:
//
//   virtual pixel v = r * P + x   (P = W + 2; x = W, W+1 are throw-away columns)
//   tap (dy, dx) of virtual pixel v is halo row v + dy * P + dx              -- affine in v
//
// A K-major swizzled operand may start at ANY row of a TMA-written tile (the swizzle is a function of the address
// bits: tools/micro/umma_rowshift1.cu, profiles/umma_rowshift_r1.txt), so tap (dy, dx) is the same tile entered
// (dy * P + dx) rows further down.  M = 128 accumulator rows hold R * P <= 128 virtual pixels (W=32: 96 real pixels,
// W=16: 112); rows past R * P and the two throw-away columns are computed from whatever follows in shared memory and
// never stored or counted.  The whole weight matrix stays resident in shared memory (loaded once per CTA).
// L2 -> SM bytes per output pixel drop from 4.5x to ~1.8x (W=32) / ~1.5x (W=16) of the input.
//
// Roles (256 threads): warp 0 TMA producer, warps 1..3 MMA issuers (each takes every third tap into its own TMEM
// accumulator), warps 4..7 epilogue (sum of the three partial accumulators -> bf16 staging tile -> column statistics of
// the real pixels -> coalesced stores, fused train-mode BatchNorm finalize by the last CTA).
// Reference semantics replaced: the 3x3 nn.Conv2d + nn.BatchNorm2d statistics of torchreid/models/hrnet.py:67-96.
#include "common.cuh"
#include "kernels.h"

namespace bpb {

constexpr int kHM = 128;          // accumulator rows (virtual pixels) per tile
constexpr int kHW = 3;            // MMA-issuing warps
constexpr int kHThreads = 32 * (1 + kHW + 4);
constexpr int kHMaxStages = 8;

struct HaloParams {
    int n_img, H, W, cin, cout;
    int P, R, tiles_per_img, total_tiles;
    int kchunks, kc, row_bytes;
    int a_rows_smem, box_bytes, stages, tmem_cols, b_res_bytes;
    int ldo;
    void* out;
    double* bn_acc; unsigned int* bn_counter;
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;
    float* scale; float* shift;
    float eps, momentum;
    double inv_count, unbias;
    int c_lo;
};

template <int BN>
__global__ void __maxnreg__(96) halo_conv_kernel(const __grid_constant__ CUtensorMap tmA,
                                                 const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int stage_bytes = p.a_rows_smem * p.row_bytes;
    const int b_tile_bytes = BN * p.row_bytes;
    constexpr int kPitch = BN * 2 + 16;
    uint8_t* b_res = smem + p.stages * stage_bytes;
    uint8_t* epi_tile = b_res + p.b_res_bytes;                                  // [128][kPitch]
    float* cta_stats = reinterpret_cast<float*>(epi_tile + kHM * kPitch);       // [2][cout]
    const bool fused_bn = p.bn_acc != nullptr;
    if (fused_bn)
        for (int i = threadIdx.x; i < 2 * p.cout; i += blockDim.x) cta_stats[i] = 0.f;

    __shared__ uint64_t full_bar[kHMaxStages], empty_bar[kHMaxStages], acc_full[2], acc_empty[2], b_full;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float red_smem[2][128 / (BN / 2)][BN];
    __shared__ uint32_t s_pix[kHM];           // output pixel index of each accumulator row, 0xFFFFFFFF = throw-away

    const uint32_t smem_a = smem_u32(smem), b_res_a = smem_u32(b_res);
    const uint32_t full_a = smem_u32(&full_bar[0]), empty_a = smem_u32(&empty_bar[0]);
    const uint32_t accf_a = smem_u32(&acc_full[0]), acce_a = smem_u32(&acc_empty[0]), bfull_a = smem_u32(&b_full);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kHW);                      // every MMA warp commits its taps of the stage
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], kHW);
            mbar_init(&acc_empty[b], 128);
        }
        mbar_init(&b_full, 1);
        fence_barrier_init();
        mbar_expect_tx_a(bfull_a, p.b_res_bytes);
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc_n(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_smem, 0);

    // one-time load of the weight matrix: tile j = tap * kchunks + chunk  <->  columns [j * kc, (j+1) * kc) of OHWI
    if (warp >= 1 && elect_one_sync()) {
        const int total_b = 9 * p.kchunks;
        for (int j = warp - 1; j < total_b; j += (kHThreads / 32) - 1)
            tma_load_2d_a(b_res_a + j * b_tile_bytes, &tmB, bfull_a, j * p.kc, 0);
    }

    if (warp == 0) {
        // ------------------------------ TMA producer: one halo box per (tile, channel chunk) ------------------------------
        const bool leader = elect_one_sync();
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int n = t / p.tiles_per_img;
            const int h0 = (t - n * p.tiles_per_img) * p.R;
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait_a(empty_a + s * 8, ph ^ 1);
                if (leader) {
                    mbar_expect_tx_a(full_a + s * 8, p.box_bytes);
                    tma_load_4d_a(smem_a + s * stage_bytes, &tmA, full_a + s * 8, c * p.kc, -1, h0 - 1, n);
                }
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp <= kHW) {
        // ------------------------------ MMA issuers: warp w takes taps w, w+3, w+6 ------------------------------
        constexpr uint32_t idesc = make_idesc(false, kHM, BN);
        const bool leader = elect_one_sync();
        const int w = warp - 1;
        const int instrs = p.row_bytes >> 5;
        const uint64_t desc_hi = make_kmajor_desc(0, p.row_bytes);
        int s = 0;
        uint32_t ph = 0, lt = 0;
        mbar_wait_a(bfull_a, 0);
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1, use = lt >> 1;
            mbar_wait_a(acce_a + buf * 8, (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (buf * kHW + w) * BN;
            uint32_t first = 1;
            for (int c = 0; c < p.kchunks; ++c) {
                const uint32_t stage_a = smem_a + s * stage_bytes;
                mbar_wait_a(full_a + s * 8, ph);
                tc_fence_after();
                if (leader) {
                    for (int tap = w; tap < 9; tap += kHW) {
                        const int dy = tap / 3, dx = tap - dy * 3;
                        const uint32_t a_addr = stage_a + (dy * p.P + dx) * p.row_bytes;
                        const uint32_t b_addr = b_res_a + (tap * p.kchunks + c) * b_tile_bytes;
                        const uint64_t da = desc_hi | static_cast<uint64_t>((a_addr & 0x3FFFFu) >> 4);
                        const uint64_t db = desc_hi | static_cast<uint64_t>((b_addr & 0x3FFFFu) >> 4);
                        for (int k = 0; k < instrs; ++k) {
                            umma<false>(tacc, da + 2 * k, db + 2 * k, idesc, first ? 0u : 1u);
                            first = 0;
                        }
                    }
                    umma_commit_a(empty_a + s * 8);
                    if (c == p.kchunks - 1) umma_commit_a(accf_a + buf * 8);
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue (4 warps): accumulator row == virtual pixel ------------------------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = (warp - 1 - kHW) * 32 + lane;            // 0..127
        const int r_of_row = row / p.P, x_of_row = row - r_of_row * p.P;
        uint32_t lt = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
            const int n = t / p.tiles_per_img;
            const int h0 = (t - n * p.tiles_per_img) * p.R;
            const int rows_valid = min(p.R, p.H - h0);
            const bool valid = x_of_row < p.W && r_of_row < rows_valid;
            const uint32_t buf = lt & 1, use = lt >> 1;
            mbar_wait_a(accf_a + buf * 8, use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kHW * BN;
            // the previous tile's readers of the staging buffer / pixel table must be done before they are overwritten
            asm volatile("bar.sync 1, 128;" ::: "memory");
            s_pix[row] = valid ? static_cast<uint32_t>((n * p.H + h0 + r_of_row) * p.W + x_of_row) : 0xFFFFFFFFu;
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16], u1[16], u2[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld16(taddr + BN + c0, u1);
                tmem_ld16(taddr + 2 * BN + c0, u2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u1[j]) + __uint_as_float(u2[j]));
                if (c0 + 16 >= BN) {
                    tc_fence_before();
                    mbar_arrive_a(acce_a + buf * 8);
                }
                uint4* dst = reinterpret_cast<uint4*>(epi_tile + row * kPitch + c0 * 2);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint4 o;
                    o.x = valid ? pack_bf16(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])) : 0u;
                    o.y = valid ? pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])) : 0u;
                    o.z = valid ? pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])) : 0u;
                    o.w = valid ? pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])) : 0u;
                    dst[j] = o;                                  // throw-away rows hold zeros: they add nothing below
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // ---- coalesced 16-byte stores of the real pixels ----
            {
                __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
                constexpr int kChunksPerRow = BN / 8;
#pragma unroll 4
                for (int ch = et; ch < kHM * kChunksPerRow; ch += 128) {
                    const int r = ch / kChunksPerRow, cc = ch % kChunksPerRow;
                    const uint32_t pix = s_pix[r];
                    if (pix != 0xFFFFFFFFu) {
                        const uint4 o = *reinterpret_cast<const uint4*>(epi_tile + r * kPitch + cc * 16);
                        *reinterpret_cast<uint4*>(out + static_cast<size_t>(pix) * p.ldo + cc * 8) = o;
                    }
                }
            }
            // ---- column statistics of the bf16-rounded real pixels (throw-away rows are zero) ----
            if (fused_bn) {
                constexpr int kPairs = BN / 2;
                constexpr int kGroups = 128 / kPairs;
                constexpr int kRowsPer = kHM / kGroups;
                const int pair = et % kPairs, grp = et / kPairs;
                float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                const uint8_t* src = epi_tile + grp * kPitch + pair * 4;
#pragma unroll 8
                for (int i = 0; i < kRowsPer; ++i) {
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(src + i * kGroups * kPitch);
                    const float a = bf16_lo(u), b = bf16_hi(u);
                    s0 += a; s1 += b; q0 += a * a; q1 += b * b;
                }
                red_smem[0][grp][pair * 2 + 0] = s0; red_smem[0][grp][pair * 2 + 1] = s1;
                red_smem[1][grp][pair * 2 + 0] = q0; red_smem[1][grp][pair * 2 + 1] = q1;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et < BN) {
                    float sm = 0.f, qq = 0.f;
#pragma unroll
                    for (int gg = 0; gg < kGroups; ++gg) { sm += red_smem[0][gg][et]; qq += red_smem[1][gg][et]; }
                    if (et < p.cout) {
                        cta_stats[et] += sm;
                        cta_stats[p.cout + et] += qq;
                    }
                }
            }
        }
        if (fused_bn) {
            __shared__ int s_is_last;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int c = et; c < 2 * p.cout; c += 128) {
                const float v = cta_stats[c];
                if (v != 0.f) atomicAdd(p.bn_acc + c, static_cast<double>(v));
            }
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) {
                const unsigned prev = atomicAdd(p.bn_counter, 1u);
                s_is_last = prev == gridDim.x - 1;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (s_is_last) {
                __threadfence();
                for (int ch = et; ch < p.cout; ch += 128) {
                    const double sum = __ldcg(p.bn_acc + ch), sq = __ldcg(p.bn_acc + p.cout + ch);
                    const double mean = sum * p.inv_count;
                    double var = sq * p.inv_count - mean * mean;
                    if (var < 0.0) var = 0.0;
                    const int pc = ch - p.c_lo;
                    const float g = (p.gamma && pc >= 0) ? p.gamma[pc] : 1.f;
                    const float b = (p.beta && pc >= 0) ? p.beta[pc] : 0.f;
                    const float sc = g * rsqrtf(static_cast<float>(var) + p.eps);
                    p.scale[ch] = sc;
                    p.shift[ch] = b - static_cast<float>(mean) * sc;
                    if (p.running_mean && pc >= 0)
                        p.running_mean[pc] = (1.f - p.momentum) * p.running_mean[pc] + p.momentum * static_cast<float>(mean);
                    if (p.running_var && pc >= 0)
                        p.running_var[pc] = (1.f - p.momentum) * p.running_var[pc] +
                                            p.momentum * static_cast<float>(var * p.unbias);
                    p.bn_acc[ch] = 0.0;
                    p.bn_acc[p.cout + ch] = 0.0;
                }
                if (et == 0) *p.bn_counter = 0u;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_n(tmem_base, p.tmem_cols);
    }
}

template <int BN>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloParams& p, size_t smem, int ctas_per_sm,
                       cudaStream_t stream) {
    auto kern = halo_conv_kernel<BN>;
    static bool configured = false;
    if (!configured) {
        BPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        configured = true;
    }
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        BPB_CUDA(cudaGetDevice(&dev));
        BPB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int grid = num_sms * ctas_per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    kern<<<grid, kHThreads, smem, stream>>>(tmA, tmB, p);
    BPB_LAUNCH_CHECK("halo_conv_kernel");
    return 0;
}

// Returns -1 when the layer is outside this kernel's domain (the caller then uses tc_gemm.cu), 0 on success, > 0 on error.
int conv3x3_halo_bf16(const void* x, const void* w, void* y, int n, int h, int wd, int cin, int cout, const BnFuse* bn,
                      cudaStream_t stream) {
    if (cin % 32 != 0 || (cout != 32 && cout != 64)) return -1;
    const int P = wd + 2;
    if (P > 64 || h < 1) return -1;
    if (static_cast<int64_t>(n) * h * wd >= (1ll << 31) - 1) return -1;
    const int64_t b_bytes = static_cast<int64_t>(9) * cin * cout * 2;
    if (b_bytes > 80 * 1024) return -1;
    const int bn_tile = cout;
    const int kc = cin < 64 ? 32 : 64;
    if (cin % kc) return -1;

    HaloParams p{};
    p.n_img = n; p.H = h; p.W = wd; p.cin = cin; p.cout = cout;
    p.P = P;
    p.R = kHM / P < h ? kHM / P : h;
    p.tiles_per_img = (h + p.R - 1) / p.R;
    p.total_tiles = n * p.tiles_per_img;
    p.kc = kc; p.kchunks = cin / kc; p.row_bytes = kc * 2;
    p.a_rows_smem = ((kHM + 2 * P + 2) + 7) / 8 * 8;
    p.box_bytes = (p.R + 2) * P * p.row_bytes;
    p.b_res_bytes = static_cast<int>(b_bytes);
    p.tmem_cols = bn_tile == 32 ? 256 : 512;                    // 2 buffers x 3 issuers x BN columns, power of two
    p.ldo = cout;
    p.out = y;
    if (bn != nullptr) {
        BPB_CHECK(bn->workspace != nullptr && bn->scale != nullptr && bn->shift != nullptr, "conv3x3_halo: BN outputs");
        p.bn_acc = reinterpret_cast<double*>(bn->workspace);
        p.bn_counter = reinterpret_cast<unsigned int*>(p.bn_acc + 2 * static_cast<size_t>(cout));
        p.gamma = bn->gamma; p.beta = bn->beta;
        p.running_mean = bn->running_mean; p.running_var = bn->running_var;
        p.scale = bn->scale; p.shift = bn->shift;
        p.eps = bn->eps; p.momentum = bn->momentum;
        const double m = static_cast<double>(n) * h * wd;
        p.inv_count = 1.0 / m;
        p.unbias = m > 1 ? m / (m - 1) : 1.0;
        p.c_lo = bn->c_lo;
    }
    const int stage_bytes = p.a_rows_smem * p.row_bytes;
    const int fixed = p.b_res_bytes + kHM * (bn_tile * 2 + 16) + 2 * cout * 4 + 1024;
    const int ctas = 512 / p.tmem_cols;                         // 2 for BN = 32, 1 for BN = 64
    int budget = (227 * 1024) / ctas - 12 * 1024 - fixed;
    if (budget > 200 * 1024 - fixed) budget = 200 * 1024 - fixed;
    int stages = budget / stage_bytes;
    if (stages > kHMaxStages) stages = kHMaxStages;
    if (stages < 2) return -1;
    p.stages = stages;
    const size_t smem = static_cast<size_t>(stages) * stage_bytes + fixed;

    const CUtensorMapSwizzle sw = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
        uint64_t strides[3] = {(uint64_t)cin * 2, (uint64_t)wd * cin * 2, (uint64_t)h * wd * cin * 2};
        uint32_t box[4] = {(uint32_t)kc, (uint32_t)P, (uint32_t)(p.R + 2), 1};
        uint32_t es[4] = {1, 1, 1, 1};
        if (make_tensor_map(&tmA, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dims, strides, box, es, sw)) return 4;
    }
    {
        const uint64_t ktot = (uint64_t)9 * cin;
        uint64_t dims[2] = {ktot, (uint64_t)cout};
        uint64_t strides[1] = {ktot * 2};
        uint32_t box[2] = {(uint32_t)kc, (uint32_t)bn_tile};
        uint32_t es[2] = {1, 1};
        if (make_tensor_map(&tmB, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, es, sw)) return 4;
    }
    return bn_tile == 32 ? launch_halo<32>(tmA, tmB, p, smem, ctas, stream) : launch_halo<64>(tmA, tmB, p, smem, ctas, stream);
}

}  // namespace bpb
