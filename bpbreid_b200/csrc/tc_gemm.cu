// tcgen05 implicit-GEMM convolution / dense contraction for sm_100a.
//
//   D[128 x BN] (TMEM, fp32)  +=  A[128 x KC] (smem, TMA-staged)  *  B[BN x KC]^T (smem, TMA-staged)
//
// A is either an NHWC bf16 activation tensor read through a 4-D tensor map (c, w, h, n): one TMA box per
// filter tap and channel chunk lands directly in the canonical K-major swizzled layout the tensor core
// wants (row = output pixel, 128 or 64 bytes of channels per row); zero padding comes from TMA
// out-of-bounds fill and stride-2 convolutions from the map's element strides -- or a plain row-major
// matrix [M, K] (bf16, or fp32 consumed as tf32).  B is the weight matrix [Cout, taps*Cin], K-major.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> smem tile -> coalesced global stores + BN partial sums).
// Reference semantics replaced: nn.Conv2d / nn.Linear call sites of torchreid/models/hrnet.py:67-137,
// 211-257,449-483, torchreid/models/resnet.py:105-154, torchreid/models/bpbreid.py:324-350,398-415.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace bpb {

constexpr int kTileM = 128;
constexpr int kMaxMmaWarps = 3;
constexpr int kMaxThreads = 32 * (1 + kMaxMmaWarps + 4);

struct GemmParams {
    int M;          // valid output rows (pixels)
    int N;          // valid output columns (Cout)
    int k_iters;    // taps * kchunks
    int kchunks;    // channel chunks per tap
    int kc;         // elements per chunk (row_bytes / elem size)
    int cin;        // channels per tap (column offset of a tap inside B)
    int row_bytes;  // 128 or 64: TMA inner box bytes == swizzle span
    int stages;
    int mode;       // 0: conv, one tap per stage   1: matrix (2-D A map)   2: part-concat matrix (3-D A map)
                    // 3: 3x3 stride-1 conv, three vertical taps per stage (row shift)
    int dchunks;    // mode 2: K chunks per part
    int mma_warps;  // W: MMA-issuing warps per CTA; warp w owns ring stages g = w (mod W) and its own accumulator
    int tmem_cols;  // power of two >= 2 * W * BN
    int a_rows;     // rows of the A stage: 128, or 128 + 2*Wo in mode 3
    int rshift_bytes;   // mode 3: bytes between vertically adjacent taps inside the A stage (Wo * row_bytes)
    int kwrap;      // elements after which the A channel coordinate wraps (cin for conv, K for matrix, D for parts)
    int ksize, stride, pad, Ho, Wo;
    // conv output tiling: a 128-row tile = bnimg images x bh rows x full width.  Tile t covers rows [hb*bh, hb*bh+bh)
    // of images [g*bnimg, g*bnimg+bnimg) with g = t / tpg, hb = t % tpg; tile row r is output pixel
    // (n0 + (r >> chunk_log2)) * hw + h0*Wo + (r & (chunk-1)), chunk = bh*Wo (a power of two dividing 128).
    // Matrix modes: tpg = 0 (rows are simply m0 + r).
    int tpg, bnimg, bh, chunk_log2, hw, nimg;
    int ldo;        // output leading dimension (elements)
    void* out;
    float* stats;   // optional [m_tiles, 2, N] per-tile column sum / sum of squares (EPI_RAW_STATS)
    const float* bias;
    // fused train-mode BatchNorm finalize (EPI_RAW_STATS, bn_acc != null): every CTA adds its column sums to
    // bn_acc[2][N] (fp64 atomics); the last CTA to finish turns them into scale/shift, updates the running
    // statistics and leaves bn_acc / bn_counter zeroed for the next launch.
    double* bn_acc;
    unsigned int* bn_counter;
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;
    float* scale; float* shift;
    float* save;    // optional [2][N]: batch mean / 1/sqrt(var+eps) (consumed by the BatchNorm backward)
    float eps, momentum;
    double inv_count, unbias;
    int c_lo;
    // EPI_BN_ACT (inference): out = act(raw * scale + shift [+ res]) written straight from the epilogue; scale / shift are
    // INPUTS here (running statistics folded once per weight refresh), res is an optional bf16 tensor shaped like out
    const void* res;
    int relu;
};

enum Epi { EPI_RAW_STATS = 0, EPI_F32_BIAS = 1, EPI_BN_ACT = 2 };

// Persistent, warp-specialised kernel: each CTA walks output tiles t = blockIdx.x, +gridDim.x, ... (n-tile fastest so
// consecutive tiles share their A rows in L2).  The TMA producer runs ahead across tile boundaries through the smem
// ring, the MMA warp alternates between two TMEM accumulators, and the epilogue warps drain one accumulator while the
// next tile's MMAs are in flight.
//
// mode 3 ("row shift", 3x3 stride-1 convs whose 128-pixel tile lies inside one image): one TMA box of (rows+2) image
// rows per horizontal tap s feeds the THREE vertical taps r = 0,1,2 -- tap r is the same shared-memory tile entered
// r*Wo rows further down (a multiple of the 8-row swizzle atom, so the K-major descriptor just gets a new start
// address).  A-operand traffic from L2 drops from 9 to 3*(rows+2)/rows tiles per output tile and every ring stage
// carries three taps' worth of MMAs, which amortises the single-thread issue loop.
template <int BN, bool kTf32, int kEpi>
__global__ void __maxnreg__(96) tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int taps = p.mode == 3 ? 3 : 1;                      // taps (MMA groups) per ring stage
    const int a_bytes = p.a_rows * p.row_bytes;
    const int b_tap_bytes = BN * p.row_bytes;
    const int stage_bytes = a_bytes + taps * b_tap_bytes;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int m_tiles = p.tpg == 0 ? (p.M + kTileM - 1) / kTileM : ((p.nimg + p.bnimg - 1) / p.bnimg) * p.tpg;
    const int total_tiles = m_tiles * n_tiles;
    constexpr int kPitch = BN * 2 + 16;                        // epilogue staging tile row pitch (bytes)
    uint8_t* epi_tile = smem + p.stages * stage_bytes;         // [128][kPitch], bf16-output epilogues only
    float* cta_stats = reinterpret_cast<float*>(epi_tile + kTileM * kPitch);   // [2][N] CTA-level column sums
    const bool fused_bn = kEpi == EPI_RAW_STATS && p.bn_acc != nullptr;
    constexpr bool kBnAct = kEpi == EPI_BN_ACT;
    if (fused_bn)
        for (int i = threadIdx.x; i < 2 * p.N; i += blockDim.x) cta_stats[i] = 0.f;

    __shared__ uint64_t full_bar[9], empty_bar[9], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float red_smem[2][8][BN < 32 ? 32 : BN];        // [sum|sq][row group][col]

    // shared-window addresses, computed once
    const uint32_t smem_a = smem_u32(smem);
    const uint32_t full_a = smem_u32(&full_bar[0]), empty_a = smem_u32(&empty_bar[0]);
    const uint32_t accf_a = smem_u32(&acc_full[0]), acce_a = smem_u32(&acc_empty[0]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], p.mma_warps);               // one commit per MMA warp and tile
            mbar_init(&acc_empty[b], 128);
        }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    const int W = p.mma_warps;
    if (warp == 1) tmem_alloc_n(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_smem, 0);     // warp-uniform
    // everything above touched only shared memory / TMEM / kernel parameters: the next kernel of the stream may start its
    // own prologue now, and this one must not read activations, weights or BN vectors before its predecessors are done
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        {
            // NOTE: this loop paces the whole CTA: no integer division / modulo and no address-of-shared inside
            // (measured: `g % stages`, `it / kchunks` made an iteration cost ~1450 cycles).  All 32 lanes run the
            // warp-uniform control flow, the elected lane issues (see elect_one_sync in common.cuh).
            const bool leader = elect_one_sync();
            int s = 0;
            uint32_t ph = 0;                                    // parity of the ring round
            uint32_t stage_a = smem_a, fb = full_a, eb = empty_a;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int m_tile = t / n_tiles, n_tile = t - m_tile * n_tiles;
                int n0 = 0, hbase = 0;
                if (p.mode == 0 || p.mode == 3) {
                    const int grp = m_tile / p.tpg;
                    n0 = grp * p.bnimg;
                    hbase = (m_tile - grp * p.tpg) * p.bh * p.stride - p.pad;
                }
                const int m0 = m_tile * kTileM, nb0 = n_tile * BN;
                int tap_r = 0, tap_s = 0, kc = 0, bk = 0, part = 0;
                for (int it = 0; it < p.k_iters; ++it) {
                    mbar_wait_a(eb, ph ^ 1);
                    if (leader) {
                    mbar_expect_tx_a(fb, stage_bytes);
                    if (p.mode == 3) {
                        tma_load_4d_a(stage_a, &tmA, fb, kc, tap_s - 1, hbase, n0);
                        const int col = tap_s * p.cin + kc;          // tap (r, s) lives at column (3 r + s) * cin + kc
                        tma_load_2d_a(stage_a + a_bytes, &tmB, fb, col, nb0);
                        tma_load_2d_a(stage_a + a_bytes + b_tap_bytes, &tmB, fb, col + 3 * p.cin, nb0);
                        tma_load_2d_a(stage_a + a_bytes + 2 * b_tap_bytes, &tmB, fb, col + 6 * p.cin, nb0);
                    } else {
                        if (p.mode == 0) {
                            tma_load_4d_a(stage_a, &tmA, fb, kc, tap_s - p.pad, hbase + tap_r, n0);
                        } else if (p.mode == 1) {
                            tma_load_2d_a(stage_a, &tmA, fb, kc, m0);
                        } else {
                            tma_load_3d_a(stage_a, &tmA, fb, kc, m0, part);
                        }
                        tma_load_2d_a(stage_a + a_bytes, &tmB, fb, bk, nb0);
                    }
                    }
                    bk += p.kc;                                 // B columns advance contiguously: tap*cin + kc
                    kc += p.kc;
                    if (kc == p.kwrap) {                        // end of the channel chunks of this tap / part
                        kc = 0;
                        ++part;
                        if (++tap_s == p.ksize) { tap_s = 0; ++tap_r; }
                    }
                    stage_a += stage_bytes; fb += 8; eb += 8;
                    if (++s == p.stages) { s = 0; ph ^= 1; stage_a = smem_a; fb = full_a; eb = empty_a; }
                }
            }
        }
    } else if (warp <= W) {
        // ------------------------------ MMA issuers (W warps) ------------------------------
        // A single thread can issue one tcgen05.mma per ~100 cycles (tools/micro/umma_rate.cu), 3-6x slower than the
        // 16/32-cycle math of the narrow N=32/64 tiles these convolutions need.  So W warps issue concurrently: warp w
        // consumes ring stages g = w, w+W, ... and accumulates into ITS OWN TMEM tile (no cross-thread ordering is
        // needed); the epilogue adds the W partial accumulators.  W divides k_iters, so every warp owns k_iters/W
        // stages of every tile.
        constexpr uint32_t idesc = make_idesc(kTf32, kTileM, BN);
        const bool leader = elect_one_sync();
        const int w = warp - 1;
        const int instrs = p.row_bytes >> 5;                   // 32 bytes of K per instruction (16 bf16 / 8 tf32)
        const uint64_t desc_hi = make_kmajor_desc(0, p.row_bytes);          // everything but the start address
        const uint32_t rshift16 = static_cast<uint32_t>(p.rshift_bytes) >> 4;
        const uint32_t btap16 = static_cast<uint32_t>(b_tap_bytes) >> 4;
        const int per_tile = p.k_iters / W;                    // stages this warp consumes per tile
        int s = w;                                             // ring slot of this warp's next stage
        uint32_t ph = 0, lt = 0;
        while (s >= p.stages) { s -= p.stages; ph ^= 1; }
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1, use = lt >> 1;
            mbar_wait_a(acce_a + buf * 8, (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (buf * W + w) * BN;
            for (int j = 0; j < per_tile; ++j) {
                const uint32_t stage_a = smem_a + s * stage_bytes;
                mbar_wait_a(full_a + s * 8, ph);
                tc_fence_after();
                if (leader) {
                    const uint64_t da = desc_hi | static_cast<uint64_t>((stage_a & 0x3FFFFu) >> 4);
                    const uint64_t db = desc_hi | static_cast<uint64_t>(((stage_a + a_bytes) & 0x3FFFFu) >> 4);
                    for (int r = 0; r < taps; ++r)
                        for (int k = 0; k < instrs; ++k)
                            umma<kTf32>(tacc, da + r * rshift16 + 2 * k, db + r * btap16 + 2 * k, idesc,
                                        (j > 0 || r > 0 || k > 0) ? 1u : 0u);
                    umma_commit_a(empty_a + s * 8);            // smem slot free once these MMAs retire
                    if (j == per_tile - 1) umma_commit_a(accf_a + buf * 8);
                }
                __syncwarp();
                s += W;
                while (s >= p.stages) { s -= p.stages; ph ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue (4 warps after the MMA warps) ------------------------------
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;          // row of the tile == output pixel
        const int et = (warp - 1 - W) * 32 + lane;  // 0..127 epilogue thread id
        uint32_t lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int m_tile = t / n_tiles, n_tile = t % n_tiles;
            const uint32_t buf = lt & 1, use = lt >> 1;
            const int c_base = n_tile * BN;
            // first output row of the tile and how many of its 128 rows exist (always a prefix: rows are image-major)
            int m0, valid_rows;
            if (p.tpg == 0) {
                m0 = m_tile * kTileM;
                valid_rows = min(kTileM, p.M - m0);
            } else {
                const int grp = m_tile / p.tpg;
                const int n0 = grp * p.bnimg;
                m0 = n0 * p.hw + (m_tile - grp * p.tpg) * p.bh * p.Wo;
                valid_rows = min(kTileM, (p.nimg - n0) << p.chunk_log2);
            }
            const int cmask = (1 << p.chunk_log2) - 1;
            mbar_wait_a(accf_a + buf * 8, use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * W * BN;

            if constexpr (kEpi == EPI_F32_BIAS) {
                float* out = reinterpret_cast<float*>(p.out);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                    for (int ww = 1; ww < W; ++ww) {            // add the other MMA warps' partial accumulators
                        uint32_t u[32];
                        tmem_ld32(taddr + ww * BN + c0, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
                    }
                    if (c0 + 32 >= BN) {                        // accumulator fully read: hand it back to the MMA warps
                        tc_fence_before();
                        mbar_arrive_a(acce_a + buf * 8);
                    }
                    if (row < valid_rows) {
                        float* dst = out + static_cast<size_t>(m0 + (row >> p.chunk_log2) * p.hw + (row & cmask)) * p.ldo +
                                     c_base + c0;
                        if (c_base + c0 + 32 <= p.N) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 o;
                                o.x = __uint_as_float(v[j + 0]) + (p.bias ? p.bias[c_base + c0 + j + 0] : 0.f);
                                o.y = __uint_as_float(v[j + 1]) + (p.bias ? p.bias[c_base + c0 + j + 1] : 0.f);
                                o.z = __uint_as_float(v[j + 2]) + (p.bias ? p.bias[c_base + c0 + j + 2] : 0.f);
                                o.w = __uint_as_float(v[j + 3]) + (p.bias ? p.bias[c_base + c0 + j + 3] : 0.f);
                                *reinterpret_cast<float4*>(dst + j) = o;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c_base + c0 + j < p.N)
                                    dst[j] = __uint_as_float(v[j]) + (p.bias ? p.bias[c_base + c0 + j] : 0.f);
                        }
                    }
                }
            } else {
                // previous tile's readers of the staging buffer must be done before it is overwritten
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // 16-column chunks: the W partial accumulators of a chunk are requested together and waited for once
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[16], u1[16], u2[16];
                    tmem_ld16(taddr + c0, v);
                    if (W > 1) tmem_ld16(taddr + BN + c0, u1);
                    if (W > 2) tmem_ld16(taddr + 2 * BN + c0, u2);
                    tmem_ld_wait();
                    if (W > 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u1[j]));
                    }
                    if (W > 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u2[j]));
                    }
                    if (c0 + 16 >= BN) {
                        tc_fence_before();
                        mbar_arrive_a(acce_a + buf * 8);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(epi_tile + row * kPitch + c0 * 2);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                        o.y = pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                        o.z = pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                        o.w = pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                        dst[j] = o;
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // ---- coalesced 16-byte stores of the tile (issued first: their latency hides under the statistics) ----
                {
                    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
                    constexpr int kChunksPerRow = BN / 8;
                    // kChunksPerRow divides 128, so a thread always handles the SAME 8 channels of the tile: the folded
                    // BatchNorm scale / shift of those channels live in registers (inference epilogue)
                    float bsc[8], bsh[8];
                    if constexpr (kBnAct) {
                        const int cc0 = et % kChunksPerRow;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int chn = c_base + cc0 * 8 + j;
                            bsc[j] = chn < p.N ? p.scale[chn] : 0.f;
                            bsh[j] = chn < p.N ? p.shift[chn] : 0.f;
                        }
                    }
                    const __nv_bfloat16* resp = reinterpret_cast<const __nv_bfloat16*>(p.res);
#pragma unroll 4
                    for (int ch = et; ch < kTileM * kChunksPerRow; ch += 128) {
                        const int r = ch / kChunksPerRow, cc = ch % kChunksPerRow;
                        if (r < valid_rows) {
                            uint4 o = *reinterpret_cast<const uint4*>(epi_tile + r * kPitch + cc * 16);
                            const size_t off = static_cast<size_t>(m0 + (r >> p.chunk_log2) * p.hw + (r & cmask)) * p.ldo +
                                               c_base + cc * 8;
                            if constexpr (kBnAct) {
                                // BatchNorm applied to the bf16-ROUNDED conv output: bit-identical to conv -> bn_apply
                                float f[8];
                                f[0] = bf16_lo(o.x); f[1] = bf16_hi(o.x); f[2] = bf16_lo(o.y); f[3] = bf16_hi(o.y);
                                f[4] = bf16_lo(o.z); f[5] = bf16_hi(o.z); f[6] = bf16_lo(o.w); f[7] = bf16_hi(o.w);
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = f[j] * bsc[j] + bsh[j];
                                if (resp != nullptr) {
                                    const uint4 rv = *reinterpret_cast<const uint4*>(resp + off);
                                    f[0] += bf16_lo(rv.x); f[1] += bf16_hi(rv.x); f[2] += bf16_lo(rv.y); f[3] += bf16_hi(rv.y);
                                    f[4] += bf16_lo(rv.z); f[5] += bf16_hi(rv.z); f[6] += bf16_lo(rv.w); f[7] += bf16_hi(rv.w);
                                }
                                if (p.relu) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                                }
                                o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
                                o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
                            }
                            *reinterpret_cast<uint4*>(out + off) = o;
                        }
                    }
                }
                // ---- per-tile column statistics from the bf16-rounded values ----
                if (kEpi == EPI_RAW_STATS && (p.stats != nullptr || fused_bn)) {
                    constexpr int kPairs = BN / 2;
                    constexpr int kGroups = 128 / kPairs;                   // BN=32:8, 64:4, 128:2
                    constexpr int kRowsPer = kTileM / kGroups;              // rows each thread adds up
                    const int pair = et % kPairs, grp = et / kPairs;
                    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                    const uint8_t* src = epi_tile + grp * kPitch + pair * 4;
#pragma unroll 8
                    for (int i = 0; i < kRowsPer; ++i) {
                        if (grp + i * kGroups < valid_rows) {
                            const uint32_t u = *reinterpret_cast<const uint32_t*>(src + i * kGroups * kPitch);
                            const float a = bf16_lo(u), b = bf16_hi(u);
                            s0 += a; s1 += b; q0 += a * a; q1 += b * b;
                        }
                    }
                    red_smem[0][grp][pair * 2 + 0] = s0; red_smem[0][grp][pair * 2 + 1] = s1;
                    red_smem[1][grp][pair * 2 + 0] = q0; red_smem[1][grp][pair * 2 + 1] = q1;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et < BN) {
                        float s = 0.f, qq = 0.f;
#pragma unroll
                        for (int gg = 0; gg < kGroups; ++gg) { s += red_smem[0][gg][et]; qq += red_smem[1][gg][et]; }
                        if (p.stats != nullptr) {
                            float* st = p.stats + static_cast<size_t>(m_tile) * 2 * p.N;
                            st[c_base + et] = s;
                            st[p.N + c_base + et] = qq;
                        }
                        if (fused_bn) {                          // owner thread of this column: no race
                            cta_stats[c_base + et] += s;
                            cta_stats[p.N + c_base + et] += qq;
                        }
                    }
                }
            }
        }
        if constexpr (kEpi == EPI_RAW_STATS) {
            if (fused_bn) {
                __shared__ int s_is_last;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int c = et; c < 2 * p.N; c += 128) {
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
                    for (int ch = et; ch < p.N; ch += 128) {
                        const double sum = __ldcg(p.bn_acc + ch), sq = __ldcg(p.bn_acc + p.N + ch);
                        const double mean = sum * p.inv_count;
                        double var = sq * p.inv_count - mean * mean;
                        if (var < 0.0) var = 0.0;
                        const int pc = ch - p.c_lo;
                        const float g = (p.gamma && pc >= 0) ? p.gamma[pc] : 1.f;
                        const float b = (p.beta && pc >= 0) ? p.beta[pc] : 0.f;
                        const float istd = rsqrtf(static_cast<float>(var) + p.eps);
                        const float sc = g * istd;
                        p.scale[ch] = sc;
                        p.shift[ch] = b - static_cast<float>(mean) * sc;
                        if (p.save) { p.save[ch] = static_cast<float>(mean); p.save[p.N + ch] = istd; }
                        if (p.running_mean && pc >= 0)
                            p.running_mean[pc] = (1.f - p.momentum) * p.running_mean[pc] + p.momentum * static_cast<float>(mean);
                        if (p.running_var && pc >= 0)
                            p.running_var[pc] = (1.f - p.momentum) * p.running_var[pc] +
                                                p.momentum * static_cast<float>(var * p.unbias);
                        p.bn_acc[ch] = 0.0;
                        p.bn_acc[p.N + ch] = 0.0;
                    }
                    if (et == 0) *p.bn_counter = 0u;
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
// host side
// ----------------------------------------------------------------------------------------------
template <int BN, bool kTf32, int kEpi>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
    const int stage_bytes = (p.a_rows + (p.mode == 3 ? 3 : 1) * BN) * p.row_bytes;
    const size_t epi = kEpi != EPI_F32_BIAS
                           ? static_cast<size_t>(kTileM) * (BN * 2 + 16) + (p.bn_acc ? 2 * sizeof(float) * p.N : 0)
                           : 0;
    const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + epi + 1024;
    BPB_CHECK(smem <= 200 * 1024, "tc_gemm: %zu bytes of shared memory requested", smem);
    auto kern = tc_gemm_kernel<BN, kTf32, kEpi>;
    BPB_SET_MAX_SMEM(kern, 200 * 1024);
    const int m_tiles_l = p.tpg == 0 ? (p.M + kTileM - 1) / kTileM : ((p.nimg + p.bnimg - 1) / p.bnimg) * p.tpg;
    const int total = m_tiles_l * ((p.N + BN - 1) / BN);
    // persistent: exactly as many CTAs as are co-resident (registers / smem via the occupancy API, TMEM: 2*BN
    // columns each) -- a static tile assignment must never have a second wave of CTAs
    // (shared memory is the binding limit: 56 registers x 192 threads allow 6 CTAs; the occupancy API under-reports
    //  here, so the bound is computed directly: 227 KB usable per SM, ~4 KB static + reserved per CTA)
    int per_sm = static_cast<int>((227 * 1024) / (smem + 4 * 1024));
    if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
    const int threads = 32 * (1 + p.mma_warps + 4);
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    const int num_sms = device_sm_count();
    if (num_sms <= 0) return 2;
    int grid = num_sms * per_sm;
    if (grid > total) grid = total;
    // short launches (every CTA gets at most ~2 tiles) gain from overlapping their prologue with the previous kernel's tail
    const bool pdl = total <= 2 * num_sms * per_sm;
    if (check_cuda(launch_pdl(pdl, kern, dim3(grid), dim3(threads), smem, stream, tmA, tmB, p), "tc_gemm_kernel")) return 3;
    BPB_LAUNCH_CHECK("tc_gemm_kernel");
    return 0;
}

// 128 output pixels = bnimg images x bh rows x the full output width: bh = the largest divisor of ho whose bh*wo
// divides 128 (384x128 inputs: the 24x8 map tiles as 2 images x 8 rows, the 12x4 map as 8 images x 4 rows).
bool conv_tiling(int ho, int wo, int* bh_out, int* bnimg_out) {
    if (wo < 1 || wo > 128 || 128 % wo) return false;
    int bh = 128 / wo < ho ? 128 / wo : ho;
    while (bh > 1 && (ho % bh || 128 % (bh * wo))) --bh;
    *bh_out = bh;
    *bnimg_out = 128 / (bh * wo);
    return true;
}

static int pick_stages(int stage_bytes, int budget_bytes) {
    // ring depth: the ring spans tile boundaries in the persistent kernel, so depth only has to cover the TMA latency
    int s = budget_bytes / stage_bytes;
    if (s > 8) s = 8;
    if (s < 2) s = 2;
    return s;
}

static int pow2_cols(int c) {
    int v = 32;
    while (v < c) v <<= 1;
    return v;
}

// Pick the number of MMA-issuing warps W (must divide k_iters), the TMEM allocation and the ring depth so that
// (issuers per SM) = W x (CTAs per SM) is maximal: TMEM allows 512 / pow2(2*W*BN) CTAs, shared memory the rest.
static void choose_issue_config(GemmParams& p, int bn_tile, int taps, int epi_bytes, int max_w = kMaxMmaWarps) {
    // The ring depth must be a multiple of W: then every ring slot is always consumed by the same MMA warp, which is
    // what the mbarrier parity protocol needs (a warp that shares a slot with another warp could skip a phase and
    // mistake an old completion for the one it is waiting for).
    const int stage_bytes = (p.a_rows + taps * bn_tile) * p.row_bytes;
    const int kMaxSmem = 190 * 1024;
    int best_w = 1, best_issuers = 0, best_stages = 2;
    // first pass: every MMA warp gets two ring slots (prefetch distance); second pass: one slot each
    // (a first pass that demanded two slots per warp was measured slower: it trades CTAs per SM for ring depth)
    for (int slots_per_warp = 1; slots_per_warp >= 1 && best_issuers == 0; --slots_per_warp) {
        const int w_cap = max_w;
        for (int w = (w_cap < kMaxMmaWarps ? w_cap : kMaxMmaWarps); w >= 1; --w) {
            if (p.k_iters % w) continue;
            const int cols = pow2_cols(2 * w * bn_tile);
            if (cols > 512) continue;
            int ctas = 512 / cols;
            if (ctas > 3) ctas = 3;
            for (; ctas >= 1; --ctas) {
                int budget = (227 * 1024) / ctas - 5 * 1024 - epi_bytes;
                if (budget > kMaxSmem - epi_bytes) budget = kMaxSmem - epi_bytes;
                int st = budget / stage_bytes;
                if (st > 9) st = 9;
                st = (st / w) * w;
                int min_st = slots_per_warp * w;
                if (min_st < 2) min_st = 2;
                if (st < min_st) continue;
                const int issuers = w * ctas;
                if (issuers > best_issuers) { best_issuers = issuers; best_w = w; best_stages = st; }
                break;
            }
        }
    }
    p.mma_warps = best_w;
    p.tmem_cols = pow2_cols(2 * best_w * bn_tile);
    p.stages = best_stages;
}

static int pick_bn(int m_tiles, int cout, int bn_hint) {
    if (bn_hint > 0) return bn_hint;
    int bn = cout >= 128 ? 128 : (cout >= 64 ? 64 : 32);
    // keep at least ~2 waves of CTAs on 148 SMs when the layer is small
    while (bn > 32 && m_tiles * ((cout + bn - 1) / bn) < 296) bn >>= 1;
    return bn;
}

// NHWC bf16 convolution (ksize 1 or 3, stride 1 or 2, pad = ksize/2), raw bf16 output + per-tile BN partials.
int conv_bf16_tc(const void* x, const void* w, void* y, float* stats, int n, int h, int wd, int cin, int cout,
                 int ksize, int stride, int bn_hint, const BnFuse* bn, cudaStream_t stream, const BnAct* act) {
    const int pad = ksize / 2;
    BPB_CHECK(act == nullptr || (bn == nullptr && stats == nullptr && act->scale && act->shift),
              "conv_bf16_tc: the inference epilogue excludes the training statistics outputs");
    const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (wd + 2 * pad - ksize) / stride + 1;
    BPB_CHECK(cin % 32 == 0 && cout % 32 == 0, "conv_bf16_tc: channels must be multiples of 32 (cin=%d cout=%d)", cin,
              cout);
    int bh = 0, bnimg = 0;
    BPB_CHECK(conv_tiling(ho, wo, &bh, &bnimg), "conv_bf16_tc: output width %d must divide 128", wo);
    // row-shift mode: 3x3 stride-1, tile inside one image, vertical tap offset = whole swizzle atoms (Wo % 8 == 0)
    const bool rowshift = ksize == 3 && stride == 1 && bnimg == 1 && (wo % 8) == 0 && (bn_hint & 0x100) == 0;
    bn_hint &= 0xFF;
    // One thread can issue a tcgen05.mma only every ~98 cycles (tools/micro/umma_rate.cu), far slower than the
    // 16/32-cycle math of the N=32/64 MMAs these narrow convolutions need, so MMA *issue* is the bound and the lever is
    // the number of issuing CTAs per SM: 64-byte operand rows keep the ring stages small enough for 2-3 CTAs per SM.
    // 64-byte rows for the row-shift schedule: smaller ring stages -> two issuing CTAs per SM (measured faster
    // than 128-byte rows with one CTA per SM: 15.3 vs 16.2 ms per step)
    const int kc = (rowshift || cin < 64) ? 32 : 64;
    BPB_CHECK(cin % kc == 0, "conv_bf16_tc: cin %d not a multiple of %d", cin, kc);

    GemmParams p{};
    p.M = n * ho * wo;
    p.N = cout;
    p.kchunks = cin / kc;
    p.kc = kc;
    p.cin = cin;
    p.row_bytes = kc * 2;
    p.kwrap = cin;
    p.ksize = ksize; p.stride = stride; p.pad = pad; p.Ho = ho; p.Wo = wo;
    p.ldo = cout;
    p.out = y;
    p.stats = stats;
    p.bias = nullptr;
    p.tpg = ho / bh; p.bnimg = bnimg; p.bh = bh; p.hw = ho * wo; p.nimg = n;
    p.chunk_log2 = 0;
    while ((1 << p.chunk_log2) < bh * wo) ++p.chunk_log2;
    p.mode = rowshift ? 3 : 0;
    p.a_rows = rowshift ? kTileM + 2 * wo : kTileM;
    p.rshift_bytes = rowshift ? wo * p.row_bytes : 0;
    p.k_iters = rowshift ? 3 * p.kchunks : ksize * ksize * p.kchunks;
    if (bn != nullptr) {
        BPB_CHECK(bn->workspace != nullptr && bn->scale != nullptr && bn->shift != nullptr, "conv_bf16_tc: BN outputs");
        p.bn_acc = reinterpret_cast<double*>(bn->workspace);
        p.bn_counter = reinterpret_cast<unsigned int*>(p.bn_acc + 2 * static_cast<size_t>(cout));
        p.gamma = bn->gamma; p.beta = bn->beta;
        p.running_mean = bn->running_mean; p.running_var = bn->running_var;
        p.scale = bn->scale; p.shift = bn->shift;
        p.save = bn->save_mean_invstd;
        p.eps = bn->eps; p.momentum = bn->momentum;
        p.inv_count = 1.0 / static_cast<double>(p.M);
        p.unbias = p.M > 1 ? static_cast<double>(p.M) / static_cast<double>(p.M - 1) : 1.0;
        p.c_lo = bn->c_lo;
    }
    const int m_tiles = conv_num_mtiles(n, ho, wo);
    int bn_tile = pick_bn(m_tiles, cout, bn_hint);
    if (bn_hint == 0 && ksize == 3 && bn_tile > 64) bn_tile = 64;       // narrow tiles + several issuers beat one N=128 issuer
    // Layers with many tiles (>= 4096: the stride-4 maps at batch 256) run faster as three CTAs per SM with ONE MMA
    // warp each (three epilogues, three producers in flight) than as two CTAs with three MMA warps (measured: C=32 3x3
    // 34.8 vs 38.9 us, stem conv2 84 vs 92 us); the small low-resolution layers prefer the wide CTAs.
    const int total_tiles_h = m_tiles * ((cout + bn_tile - 1) / bn_tile);
    choose_issue_config(p, bn_tile, (rowshift ? 3 : 1), kTileM * (bn_tile * 2 + 16) + (bn ? 8 * cout : 0),
                        total_tiles_h >= 4096 ? 1 : kMaxMmaWarps);

    const CUtensorMapSwizzle sw = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
        uint64_t strides[3] = {(uint64_t)cin * 2, (uint64_t)wd * cin * 2, (uint64_t)h * wd * cin * 2};
        uint32_t box[4] = {(uint32_t)kc, (uint32_t)(wo * stride), (uint32_t)((rowshift ? bh + 2 : bh) * stride),
                           (uint32_t)bnimg};
        uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        if (make_tensor_map(&tmA, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dims, strides, box, es, sw)) return 4;
    }
    {
        const uint64_t ktot = (uint64_t)ksize * ksize * cin;
        uint64_t dims[2] = {ktot, (uint64_t)cout};
        uint64_t strides[1] = {ktot * 2};
        uint32_t box[2] = {(uint32_t)kc, (uint32_t)bn_tile};
        uint32_t es[2] = {1, 1};
        if (make_tensor_map(&tmB, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, es, sw)) return 4;
    }
    if (act != nullptr) {
        p.scale = const_cast<float*>(act->scale); p.shift = const_cast<float*>(act->shift);
        p.res = act->res; p.relu = act->relu;
        switch (bn_tile) {
            case 32: return launch<32, false, EPI_BN_ACT>(tmA, tmB, p, stream);
            case 64: return launch<64, false, EPI_BN_ACT>(tmA, tmB, p, stream);
            case 128: return launch<128, false, EPI_BN_ACT>(tmA, tmB, p, stream);
            default: BPB_CHECK(false, "conv_bf16_tc: unsupported BN %d", bn_tile);
        }
    }
    switch (bn_tile) {
        case 32: return launch<32, false, EPI_RAW_STATS>(tmA, tmB, p, stream);
        case 64: return launch<64, false, EPI_RAW_STATS>(tmA, tmB, p, stream);
        case 128: return launch<128, false, EPI_RAW_STATS>(tmA, tmB, p, stream);
        default: BPB_CHECK(false, "conv_bf16_tc: unsupported BN %d", bn_tile);
    }
    return 0;
}

// out[M, N] (fp32, leading dimension ldo) = A[M, K] * W[N, K]^T + bias ; fp32 operands consumed as tf32.
// a_mode 0: A row-major with leading dimension lda.  a_mode 1: A is the part-major tensor [a_parts, M, K/a_parts]
// read as the concatenation over parts (bpbreid.py:212 concat_parts_embeddings) without materialising it.
int linear_tf32_tc(const float* a, int64_t lda, const float* w, const float* bias, float* out, int m, int n, int k,
                   int ldo, int a_mode, int a_parts, cudaStream_t stream) {
    // K need not be a multiple of the 32-float chunk in the row-major mode: the tensor maps carry the true K, so the
    // tail of the last chunk is TMA zero fill on both operands (the Linear backward contracts over the batch rows)
    BPB_CHECK(k >= 1 && (a_mode == 0 || k % 32 == 0), "linear_tf32_tc: K=%d must be a multiple of 32 in part mode", k);
    BPB_CHECK(k % 4 == 0, "linear_tf32_tc: K=%d must be a multiple of 4 (16-byte rows of W)", k);
    BPB_CHECK(ldo % 4 == 0 && ldo >= n, "linear_tf32_tc: ldo=%d must be >= N and a multiple of 4", ldo);
    GemmParams p{};
    p.M = m; p.N = n;
    p.kc = 32; p.kchunks = (k + 31) / 32; p.k_iters = p.kchunks; p.cin = k;
    p.row_bytes = 128;
    p.kwrap = k;
    p.a_rows = kTileM; p.rshift_bytes = 0;
    p.mode = a_mode == 0 ? 1 : 2; p.ksize = 1; p.stride = 1; p.pad = 0; p.Ho = 0; p.Wo = 0;
    p.dchunks = 1;
    p.tpg = 0; p.bnimg = 1; p.bh = 0; p.chunk_log2 = 7; p.hw = 0; p.nimg = 0;
    p.ldo = ldo; p.out = out; p.stats = nullptr; p.bias = bias;
    const int m_tiles = (m + kTileM - 1) / kTileM;
    int bn = n >= 128 ? 128 : (n >= 64 ? 64 : 32);
    while (bn > 32 && m_tiles * ((n + bn - 1) / bn) < 148) bn >>= 1;
    choose_issue_config(p, bn, 1, 0);
    CUtensorMap tmA, tmB;
    if (a_mode == 0) {
        BPB_CHECK(lda % 4 == 0 && lda >= k, "linear_tf32_tc: lda=%lld invalid", (long long)lda);
        uint64_t dims[2] = {(uint64_t)k, (uint64_t)m};
        uint64_t strides[1] = {(uint64_t)lda * 4};
        uint32_t box[2] = {32, 128};
        uint32_t es[2] = {1, 1};
        if (make_tensor_map(&tmA, a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B)) return 4;
    } else {
        BPB_CHECK(a_parts > 0 && k % a_parts == 0 && (k / a_parts) % 32 == 0, "linear_tf32_tc: bad part split");
        const int dd = k / a_parts;
        p.dchunks = dd / 32;
        p.kwrap = dd;
        uint64_t dims[3] = {(uint64_t)dd, (uint64_t)m, (uint64_t)a_parts};
        uint64_t strides[2] = {(uint64_t)dd * 4, (uint64_t)m * dd * 4};
        uint32_t box[3] = {32, 128, 1};
        uint32_t es[3] = {1, 1, 1};
        if (make_tensor_map(&tmA, a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B)) return 4;
    }
    {
        uint64_t dims[2] = {(uint64_t)k, (uint64_t)n};
        uint64_t strides[1] = {(uint64_t)k * 4};
        uint32_t box[2] = {32, (uint32_t)bn};
        uint32_t es[2] = {1, 1};
        if (make_tensor_map(&tmB, w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B)) return 4;
    }
    switch (bn) {
        case 32: return launch<32, true, EPI_F32_BIAS>(tmA, tmB, p, stream);
        case 64: return launch<64, true, EPI_F32_BIAS>(tmA, tmB, p, stream);
        case 128: return launch<128, true, EPI_F32_BIAS>(tmA, tmB, p, stream);
    }
    return 0;
}

int conv_num_mtiles(int n, int ho, int wo) {
    int bh = 0, bnimg = 0;
    if (!conv_tiling(ho, wo, &bh, &bnimg)) return (n * ho * wo + kTileM - 1) / kTileM;
    return ((n + bnimg - 1) / bnimg) * (ho / bh);
}

}  // namespace bpb
