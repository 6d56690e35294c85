// tcgen05 weight-gradient convolution for sm_100a (training backward of every nn.Conv2d on the path:
// torchreid/models/hrnet.py:60-64,80-96,116-137,222-257,459-483; torchreid/models/resnet.py:43-63,128-154;
// invoked by loss.backward() at torchreid/engine/image/part_based_engine.py:101).
//
//   dW[co, tap, ci]  +=  sum over output pixels p of   dY[p, co] * X[p (+) tap, ci]
//
// This is a GEMM whose K axis is the PIXEL axis, so both operands are "MN-major" for the tensor core: the NHWC
// tensors already have channels contiguous, i.e. K (pixels) is the strided dimension.  A TMA box of
// [128 pixels x 32|64 channels] with the 64B/128B swizzle lands in exactly the canonical MN-major swizzled layout
// (row = one pixel = one K index, 8-pixel groups SBO apart, 32|64-channel blocks LBO apart), so no transposition
// pass exists anywhere: D[co, ci] (TMEM, fp32, M = 128) += A^T[K=128 px, M] * B[K=128 px, N] per filter tap, K = 16
// pixels per tcgen05.mma.
//
// A = dY tile (the 128-pixel chunk uses the SAME image x row x width tiling as the forward kernel's output tile, so
//     the shifted input boxes are the forward kernel's A boxes), B = one shifted input box per filter tap of the
//     tap group (zero padding = TMA out-of-bounds fill, stride 2 = tensor-map element strides).
// dY may be stored zero-inserted ([N, 2Ho, 2Wo, C], the layout the stride-2 dgrad consumes): it is then read through
// element strides of 2, so that layout is written once and read by both backward kernels.
// 3x3 stride-1 layers whose chunk lies inside one image use the forward kernel's "row shift" schedule: a tap group is one
// HORIZONTAL tap; a single (rows+2)-row input box per chunk serves the three vertical taps, entered r*Wo pixel rows
// further down the tile (whole 8-row swizzle atoms: the MN-major descriptor just gets a new start address).  The
// kernel is bound by the L2 -> SM operand stream (measured 6.9 TB/s on the C=32 layers with one box per tap); this
// cuts that stream 1.6x.
// Grid: (cout tiles x cin tiles x tap groups) x K splits; every CTA adds its partial tile to the fp32 OHWI gradient
// with 16-byte vector reductions (red.global.add.v4.f32).  Channel counts < 128 on the M side replicate the channel
// block (descriptor LBO = 0): the duplicate accumulator rows are simply not read back.
//
// Roles (32 * (1 + taps + 4) threads): warp 0 TMA producer; warps 1..taps MMA issuers, ONE PER FILTER TAP of the group
// (a single thread can start a tcgen05.mma only every ~100 cycles, tools/micro/umma_rate.cu, and every tap has its own
// TMEM accumulator, so the taps issue concurrently: 24 -> 8 dependent issues per 128-pixel chunk); 4 epilogue warps.
#include "common.cuh"
#include "kernels.h"

namespace bpb {

struct WgradParams {
    int nimg, Ho, Wo;
    int tpg, bnimg, bh;           // K-chunk tiling (conv_tiling of the OUTPUT map)
    int ksize, stride, pad;
    int cin, cout;
    int cot, cit;                 // channels per tile: M side (<= 128), N side (<= 64)
    int a_row_bytes, b_row_bytes; // 64 or 128 (TMA inner box bytes == swizzle span)
    int a_blocks;                 // TMA boxes per stage on the dY side (cot / channels per box)
    int taps_g, groups;           // taps per CTA, tap groups
    int co_tiles, ci_tiles;
    int chunks, ksplit, stages;
    int dy_up;                    // dY stored zero-inserted (2x)
    int rowshift;                 // 3x3 stride-1, chunk inside one image: group = horizontal tap, ONE (bh+2)-row input box
                                  // per chunk feeds the three vertical taps through K-axis (pixel-row) start offsets
    int b_stage_bytes;            // bytes of the input-operand part of a stage
    int b_tap_stride;             // bytes between the operands of consecutive taps of the group inside a stage
    int tmem_cols;
    float* dw;                    // [cout][ksize*ksize][cin] fp32
};

// MN-major swizzled operand: 8-pixel (K) groups are `8*row_bytes` apart (SBO); channel blocks `lbo` bytes apart.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t row_bytes, uint32_t lbo_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((8u * row_bytes) >> 4) << 32;
    d |= 1ull << 46;
    d |= layout << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int CIT>
__global__ void __launch_bounds__(256) tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmY,
                                                       const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[8], empty_bar[8], acc_full;
    __shared__ uint32_t tmem_base_smem;

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int a_box_bytes = 128 * p.a_row_bytes, b_box_bytes = 128 * p.b_row_bytes;
    const int a_bytes = p.a_blocks * a_box_bytes;
    const int stage_bytes = a_bytes + p.b_stage_bytes;

    // tile coordinates
    int t = blockIdx.x;
    const int grp = t % p.groups; t /= p.groups;
    const int ci_tile = t % p.ci_tiles;
    const int co_tile = t / p.ci_tiles;
    const int co0 = co_tile * p.cot, ci0 = ci_tile * CIT;
    const int c_begin = static_cast<int>(static_cast<int64_t>(p.chunks) * blockIdx.y / p.ksplit);
    const int c_end = static_cast<int>(static_cast<int64_t>(p.chunks) * (blockIdx.y + 1) / p.ksplit);

    const uint32_t smem_a = smem_u32(smem);
    const uint32_t full_a = smem_u32(&full_bar[0]), empty_a = smem_u32(&empty_bar[0]), accf_a = smem_u32(&acc_full);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], p.taps_g); }
        mbar_init(&acc_full, p.taps_g);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmX); }
    if (warp == 1) tmem_alloc_n(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_smem, 0);

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        const bool leader = elect_one_sync();
        const int e = p.dy_up ? 2 : 1;
        const int a_elems = p.a_row_bytes >> 1;
        int s = 0;
        uint32_t ph = 0;
        for (int c = c_begin; c < c_end; ++c) {
            const int g = c / p.tpg;
            const int n0 = g * p.bnimg, h0 = (c - g * p.tpg) * p.bh;
            mbar_wait_a(empty_a + s * 8, ph ^ 1);
            if (leader) {
                const uint32_t dst = smem_a + s * stage_bytes, fb = full_a + s * 8;
                mbar_expect_tx_a(fb, stage_bytes);
                for (int b = 0; b < p.a_blocks; ++b)
                    tma_load_4d_a(dst + b * a_box_bytes, &tmY, fb, co0 + b * a_elems, 0, h0 * e, n0);
                if (p.rowshift) {
                    tma_load_4d_a(dst + a_bytes, &tmX, fb, ci0, grp - 1, h0 - 1, n0);
                } else {
                    for (int tt = 0; tt < p.taps_g; ++tt) {
                        const int tap = grp * p.taps_g + tt;
                        const int r = tap / p.ksize, sx = tap - r * p.ksize;
                        tma_load_4d_a(dst + a_bytes + tt * b_box_bytes, &tmX, fb, ci0, sx - p.pad,
                                      h0 * p.stride + r - p.pad, n0);
                    }
                }
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
    } else if (warp <= p.taps_g) {
        // ------------------------------ MMA issuers: warp 1 + t owns filter tap t of the group ------------------------
        // instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = CIT, both operands MN-major (bits 15 / 16)
        constexpr uint32_t idesc = make_idesc(false, 128, CIT) | (1u << 15) | (1u << 16);
        const bool leader = elect_one_sync();
        const int tt = warp - 1;
        const uint32_t lbo_a = p.a_blocks > 1 ? static_cast<uint32_t>(a_box_bytes) : 0u;   // 0: replicate the channel block
        const uint64_t da_hi = make_mnmajor_desc(0, p.a_row_bytes, lbo_a);
        const uint64_t db_hi = make_mnmajor_desc(0, p.b_row_bytes, 0);
        const uint32_t kstep_a = static_cast<uint32_t>(16 * p.a_row_bytes) >> 4;          // 16 pixels per MMA
        const uint32_t kstep_b = static_cast<uint32_t>(16 * p.b_row_bytes) >> 4;
        const uint32_t tacc = tmem_base + tt * CIT;
        int s = 0;
        uint32_t ph = 0;
        for (int c = c_begin; c < c_end; ++c) {
            const uint32_t stage = smem_a + s * stage_bytes;
            mbar_wait_a(full_a + s * 8, ph);
            tc_fence_after();
            if (leader) {
                const uint64_t da = da_hi | static_cast<uint64_t>((stage & 0x3FFFFu) >> 4);
                const uint64_t db = db_hi | static_cast<uint64_t>(((stage + a_bytes + tt * p.b_tap_stride) & 0x3FFFFu) >> 4);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma<false>(tacc, da + k * kstep_a, db + k * kstep_b, idesc, (c > c_begin || k > 0) ? 1u : 0u);
                umma_commit_a(empty_a + s * 8);               // the slot is free once every tap's MMAs have retired
                if (c == c_end - 1) umma_commit_a(accf_a);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
    } else if (c_end > c_begin) {
        // ------------------------------ epilogue ------------------------------
        const int q = warp & 3;
        const int row = q * 32 + lane;                       // accumulator row == output channel within the tile
        mbar_wait_a(accf_a, 0);
        tc_fence_after();
        const int taps = p.ksize * p.ksize;
        const bool row_ok = row < p.cot;                     // rows >= cot are replicas of the channel block
        float* wrow = p.dw + (static_cast<size_t>(co0 + row) * taps) * p.cin + ci0;
        for (int tt = 0; tt < p.taps_g; ++tt) {
            const int tap = p.rowshift ? tt * 3 + grp : grp * p.taps_g + tt;
#pragma unroll
            for (int c0 = 0; c0 < CIT; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tt * CIT + c0, v);
                tmem_ld_wait();
                if (row_ok) {
                    float* dst = wrow + static_cast<size_t>(tap) * p.cin + c0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
                }
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

// ----------------------------------------------------------------------------------------------
// "Tap-stacked" variant for the row-shift layers (3x3 stride-1, chunk inside one image).
// ncu on the kernel above (C=32 layer): tensor pipe 23 %, top stall mio_throttle -- the tensor core's own shared-memory
// operand reads are the bound, because with Cout < 128 the M = 128 instruction re-reads the replicated dY block.  Here the
// roles are swapped: the M side is the INPUT box, and its 128 rows are the 32|64-channel blocks of CONSECUTIVE VERTICAL
// TAPS (descriptor LBO = Wo * row bytes: block i is the same tile entered i*Wo pixel rows further down), the N side is the
// dY tile.  One MMA therefore produces D[(tap r, ci), co] for 4 (C=32) or 2 (C>=64) vertical taps at once: 8 instead of 24
// (16 for C>=64) MMAs per chunk and no replicated operand reads.  Rows of a tap index >= 3 read whatever follows the box
// in shared memory and are never written back.  Epilogue: a warp holds 32 consecutive ci of one (tap, co) per accumulator
// column -> one coalesced 128-byte red.add per column.
// ----------------------------------------------------------------------------------------------
template <int NCO>
__global__ void __launch_bounds__(224) tc_wgrad_stacked_kernel(const __grid_constant__ CUtensorMap tmY,
                                                               const __grid_constant__ CUtensorMap tmX,
                                                               const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[8], empty_bar[8], acc_full;
    __shared__ uint32_t tmem_base_smem;

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int y_box_bytes = 128 * p.a_row_bytes;                       // one dY channel block
    const int y_bytes = p.a_blocks * y_box_bytes;
    const int x_box_bytes = (128 + 2 * p.Wo) * p.b_row_bytes;          // what the TMA box writes
    const int stage_bytes = y_bytes + (128 + 3 * p.Wo) * p.b_row_bytes;  // + Wo slack rows under the garbage tap
    const int cblk = p.b_row_bytes >> 1;                               // input channels per M block: 32 or 64
    const int taps_per_acc = 128 / cblk;                               // vertical taps stacked in one accumulator
    const int nacc = p.taps_g;                                         // accumulators == MMA warps (1 or 2)

    int t = blockIdx.x;
    const int grp = t % p.groups; t /= p.groups;                       // horizontal tap
    const int ci_tile = t % p.ci_tiles;
    const int co_tile = t / p.ci_tiles;
    const int co0 = co_tile * NCO, ci0 = ci_tile * p.cit;
    const int c_begin = static_cast<int>(static_cast<int64_t>(p.chunks) * blockIdx.y / p.ksplit);
    const int c_end = static_cast<int>(static_cast<int64_t>(p.chunks) * (blockIdx.y + 1) / p.ksplit);

    const uint32_t smem_a = smem_u32(smem);
    const uint32_t full_a = smem_u32(&full_bar[0]), empty_a = smem_u32(&empty_bar[0]), accf_a = smem_u32(&acc_full);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], nacc); }
        mbar_init(&acc_full, nacc);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmX); }
    if (warp == 1) tmem_alloc_n(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_smem, 0);

    if (warp == 0) {
        const bool leader = elect_one_sync();
        const int y_elems = p.a_row_bytes >> 1;
        int s = 0;
        uint32_t ph = 0;
        for (int c = c_begin; c < c_end; ++c) {
            const int g = c / p.tpg;
            const int n0 = g * p.bnimg, h0 = (c - g * p.tpg) * p.bh;
            mbar_wait_a(empty_a + s * 8, ph ^ 1);
            if (leader) {
                const uint32_t dst = smem_a + s * stage_bytes, fb = full_a + s * 8;
                mbar_expect_tx_a(fb, y_bytes + x_box_bytes);
                for (int b = 0; b < p.a_blocks; ++b)
                    tma_load_4d_a(dst + b * y_box_bytes, &tmY, fb, co0 + b * y_elems, 0, h0, n0);
                tma_load_4d_a(dst + y_bytes, &tmX, fb, ci0, grp - 1, h0 - 1, n0);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
    } else if (warp <= nacc) {
        // M side: input box, blocks = vertical taps (LBO = Wo rows); N side: dY tile.  Both MN-major (bits 15 / 16).
        constexpr uint32_t idesc = make_idesc(false, 128, NCO) | (1u << 15) | (1u << 16);
        const bool leader = elect_one_sync();
        const int g = warp - 1;
        const uint64_t dx_hi = make_mnmajor_desc(0, p.b_row_bytes, static_cast<uint32_t>(p.Wo * p.b_row_bytes));
        const uint64_t dy_hi = make_mnmajor_desc(0, p.a_row_bytes, p.a_blocks > 1 ? static_cast<uint32_t>(y_box_bytes) : 0u);
        const uint32_t kstep_x = static_cast<uint32_t>(16 * p.b_row_bytes) >> 4;
        const uint32_t kstep_y = static_cast<uint32_t>(16 * p.a_row_bytes) >> 4;
        const uint32_t x_off = static_cast<uint32_t>(y_bytes + g * taps_per_acc * p.Wo * p.b_row_bytes);
        const uint32_t tacc = tmem_base + g * NCO;
        int s = 0;
        uint32_t ph = 0;
        for (int c = c_begin; c < c_end; ++c) {
            const uint32_t stage = smem_a + s * stage_bytes;
            mbar_wait_a(full_a + s * 8, ph);
            tc_fence_after();
            if (leader) {
                const uint64_t da = dx_hi | static_cast<uint64_t>(((stage + x_off) & 0x3FFFFu) >> 4);
                const uint64_t db = dy_hi | static_cast<uint64_t>((stage & 0x3FFFFu) >> 4);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma<false>(tacc, da + k * kstep_x, db + k * kstep_y, idesc, (c > c_begin || k > 0) ? 1u : 0u);
                umma_commit_a(empty_a + s * 8);
                if (c == c_end - 1) umma_commit_a(accf_a);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
    } else if (c_end > c_begin) {
        const int q = warp & 3;
        const int m = q * 32 + lane;                          // accumulator row = (vertical tap block, input channel)
        const int bi = m / cblk, cil = m - bi * cblk;         // bi is warp-uniform (cblk >= 32)
        mbar_wait_a(accf_a, 0);
        tc_fence_after();
        const int taps = 9;
        for (int g = 0; g < nacc; ++g) {
            const int r = g * taps_per_acc + bi;
            const int tap = r * 3 + grp;
#pragma unroll
            for (int c0 = 0; c0 < NCO; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * NCO + c0, v);
                tmem_ld_wait();
                if (r < 3) {
                    float* dst = p.dw + (static_cast<size_t>(co0 + c0) * taps + tap) * p.cin + ci0 + cil;
#pragma unroll
                    for (int j = 0; j < 16; ++j)             // lanes = 32 consecutive ci: one coalesced reduction per column
                        atomicAdd(dst + static_cast<size_t>(j) * taps * p.cin, __uint_as_float(v[j]));
                }
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

// dw [cout, ksize*ksize, cin] fp32 += wgrad(x [n,h,wd,cin] bf16, dy bf16).  dy is [n,ho,wo,cout], or with dy_up the
// zero-inserted [n,2ho,2wo,cout].  ksize in {1,3}, stride in {1,2}, pad = ksize/2.
int conv_wgrad_tc(const void* x, const void* dy, float* dw, int n, int h, int wd, int cin, int cout, int ksize,
                  int stride, int dy_up, cudaStream_t stream) {
    const bool force_plain = (dy_up & 2) != 0;               // test hook: bit 1 selects the one-accumulator-per-tap kernel
    dy_up &= 1;
    const int pad = ksize / 2;
    const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (wd + 2 * pad - ksize) / stride + 1;
    BPB_CHECK(ksize == 1 || ksize == 3, "conv_wgrad_tc: ksize %d", ksize);
    BPB_CHECK(cin % 32 == 0 && cout % 32 == 0 && (cin < 64 || cin % 64 == 0) && (cout < 64 || cout % 64 == 0),
              "conv_wgrad_tc: channels must be 32 or multiples of 64 (cin=%d cout=%d)", cin, cout);
    int bh = 0, bnimg = 0;
    BPB_CHECK(conv_tiling(ho, wo, &bh, &bnimg), "conv_wgrad_tc: output width %d must divide 128", wo);
    WgradParams p{};
    p.nimg = n; p.Ho = ho; p.Wo = wo;
    p.tpg = ho / bh; p.bnimg = bnimg; p.bh = bh;
    p.ksize = ksize; p.stride = stride; p.pad = pad;
    p.cin = cin; p.cout = cout;
    p.cot = cout < 128 ? cout : 128;
    p.cit = cin < 64 ? cin : 64;
    p.a_row_bytes = cout >= 64 ? 128 : 64;
    p.b_row_bytes = cin >= 64 ? 128 : 64;
    p.a_blocks = p.cot / (p.a_row_bytes / 2);
    p.taps_g = ksize == 3 ? 3 : 1;
    p.groups = ksize == 3 ? 3 : 1;
    p.co_tiles = cout / p.cot;
    p.ci_tiles = cin / p.cit;
    p.chunks = ((n + bnimg - 1) / bnimg) * p.tpg;
    p.dy_up = dy_up ? 1 : 0;
    p.dw = dw;
    p.rowshift = (ksize == 3 && stride == 1 && bnimg == 1 && wo % 8 == 0 && !dy_up) ? 1 : 0;
    p.b_stage_bytes = p.rowshift ? (128 + 2 * wo) * p.b_row_bytes : p.taps_g * 128 * p.b_row_bytes;
    p.b_tap_stride = p.rowshift ? wo * p.b_row_bytes : 128 * p.b_row_bytes;
    const int tiles = p.co_tiles * p.ci_tiles * p.groups;
    const int num_sms = device_sm_count();
    if (num_sms <= 0) return 2;
    int ksplit = (2 * num_sms + tiles - 1) / tiles;
    if (ksplit > p.chunks) ksplit = p.chunks;
    if (ksplit < 1) ksplit = 1;
    p.ksplit = ksplit;
    const int stage_bytes = p.a_blocks * 128 * p.a_row_bytes + p.b_stage_bytes;
    int stages = (100 * 1024) / stage_bytes;                 // ~100 KB of ring -> two CTAs per SM
    if (stages > 6) stages = 6;
    if (stages < 2) stages = 2;
    p.stages = stages;
    int cols = 32;
    while (cols < p.taps_g * p.cit) cols <<= 1;
    p.tmem_cols = cols;

    CUtensorMap tmY, tmX;
    {
        const int e = p.dy_up ? 2 : 1;
        uint64_t dims[4] = {(uint64_t)cout, (uint64_t)(wo * e), (uint64_t)(ho * e), (uint64_t)n};
        uint64_t strides[3] = {(uint64_t)cout * 2, (uint64_t)wo * e * cout * 2, (uint64_t)ho * e * wo * e * cout * 2};
        uint32_t box[4] = {(uint32_t)(p.a_row_bytes / 2), (uint32_t)(wo * e), (uint32_t)(bh * e), (uint32_t)bnimg};
        uint32_t es[4] = {1, (uint32_t)e, (uint32_t)e, 1};
        if (make_tensor_map(&tmY, dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dims, strides, box, es,
                            p.a_row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) return 4;
    }
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
        uint64_t strides[3] = {(uint64_t)cin * 2, (uint64_t)wd * cin * 2, (uint64_t)h * wd * cin * 2};
        uint32_t box[4] = {(uint32_t)p.cit, (uint32_t)(wo * stride), (uint32_t)((p.rowshift ? bh + 2 : bh) * stride),
                           (uint32_t)bnimg};
        uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        if (make_tensor_map(&tmX, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dims, strides, box, es,
                            p.b_row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) return 4;
    }
    if (p.rowshift && !force_plain) {
        // tap-stacked variant: accumulators = 1 (C=32: four 32-channel blocks = taps 0,1,2,x) or 2 (taps 0,1 | 2,x)
        p.taps_g = p.b_row_bytes == 64 ? 1 : 2;
        const int st_bytes = p.a_blocks * 128 * p.a_row_bytes + (128 + 3 * wo) * p.b_row_bytes;
        int st = (100 * 1024) / st_bytes;
        if (st > 6) st = 6;
        if (st < 2) st = 2;
        p.stages = st;
        int tc = 32;
        while (tc < p.taps_g * p.cot) tc <<= 1;
        p.tmem_cols = tc;
        const size_t sm = static_cast<size_t>(st) * st_bytes + 1024;
        dim3 sgrid(tiles, ksplit);
        const int threads = 32 * (5 + p.taps_g);
        switch (p.cot) {
            case 32:
                BPB_SET_MAX_SMEM(tc_wgrad_stacked_kernel<32>, 200 * 1024);
                tc_wgrad_stacked_kernel<32><<<sgrid, threads, sm, stream>>>(tmY, tmX, p);
                break;
            case 64:
                BPB_SET_MAX_SMEM(tc_wgrad_stacked_kernel<64>, 200 * 1024);
                tc_wgrad_stacked_kernel<64><<<sgrid, threads, sm, stream>>>(tmY, tmX, p);
                break;
            case 128:
                BPB_SET_MAX_SMEM(tc_wgrad_stacked_kernel<128>, 200 * 1024);
                tc_wgrad_stacked_kernel<128><<<sgrid, threads, sm, stream>>>(tmY, tmX, p);
                break;
            default: BPB_CHECK(false, "conv_wgrad_tc: unsupported cout tile %d", p.cot);
        }
        BPB_LAUNCH_CHECK("tc_wgrad_stacked_kernel");
        return 0;
    }
    const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024;
    dim3 grid(tiles, ksplit);
    if (p.cit == 32) {
        BPB_SET_MAX_SMEM(tc_wgrad_kernel<32>, 200 * 1024);
        tc_wgrad_kernel<32><<<grid, 32 * (5 + p.taps_g), smem, stream>>>(tmY, tmX, p);
    } else {
        BPB_SET_MAX_SMEM(tc_wgrad_kernel<64>, 200 * 1024);
        tc_wgrad_kernel<64><<<grid, 32 * (5 + p.taps_g), smem, stream>>>(tmY, tmX, p);
    }
    BPB_LAUNCH_CHECK("tc_wgrad_kernel");
    return 0;
}

}  // namespace bpb
