// HBM-bound companions of the tensor-core convolution: batch-norm statistics / application, residual and
// cross-resolution fusion, stem convolution, bilinear upsample + concat, max-pool, instance-norm (IBN).
// All activations are NHWC bf16; every kernel moves 16-byte vectors with fully coalesced warps.
// Reference semantics: torchreid/models/hrnet.py:67-137 (BN+ReLU+residual), :262-279 (fuse sum),
// :533-538 (stem), :568-573 (upsample+cat); torchreid/models/resnet.py:342-347; resnet_fastreid.py:108-122.
#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// ----------------------------------------------------------------------------------------------
// SIMT reference convolution (validation / shapes the tensor-core tiler does not cover)
// ----------------------------------------------------------------------------------------------
__global__ void conv_ref_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                __nv_bfloat16* __restrict__ y, int n, int h, int wd, int cin, int cout, int ks,
                                int stride, int ho, int wo) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t total = static_cast<int64_t>(n) * ho * wo * cout;
    if (idx >= total) return;
    const int co = idx % cout;
    const int64_t pix = idx / cout;
    const int ox = pix % wo, oy = (pix / wo) % ho, img = pix / (static_cast<int64_t>(wo) * ho);
    const int pad = ks / 2;
    float acc = 0.f;
    for (int r = 0; r < ks; ++r) {
        const int iy = oy * stride + r - pad;
        if (iy < 0 || iy >= h) continue;
        for (int s = 0; s < ks; ++s) {
            const int ix = ox * stride + s - pad;
            if (ix < 0 || ix >= wd) continue;
            const __nv_bfloat16* xp = x + ((static_cast<int64_t>(img) * h + iy) * wd + ix) * cin;
            const __nv_bfloat16* wp = w + (static_cast<int64_t>(co) * ks * ks + r * ks + s) * cin;
            for (int c = 0; c < cin; ++c) acc += __bfloat162float(xp[c]) * __bfloat162float(wp[c]);
        }
    }
    y[idx] = __float2bfloat16(acc);
}

int conv_ref(const void* x, const void* w, void* y, int n, int h, int wd, int cin, int cout, int ksize, int stride,
             cudaStream_t stream) {
    const int pad = ksize / 2;
    const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (wd + 2 * pad - ksize) / stride + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * wo * cout;
    conv_ref_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w),
        reinterpret_cast<__nv_bfloat16*>(y), n, h, wd, cin, cout, ksize, stride, ho, wo);
    BPB_LAUNCH_CHECK("conv_ref_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Stem convolution: NCHW fp32 image (3 channels) -> NHWC bf16 raw output + per-CTA BN partial sums.
// One thread = one output pixel x 64 output channels: every weight read is a warp-wide shared-memory broadcast
// (64 FMAs per 16 LDS.128), the 128-pixel tile is staged in shared memory so the global stores are fully
// coalesced and the column statistics come from the bf16-rounded values exactly like the tensor-core conv.
// Persistent grid: each CTA loops over tiles and emits ONE row of partials [2][64] (-> bpb_bn_finalize).
// Weights: PyTorch layout [64][3][k][k] fp32.
// ----------------------------------------------------------------------------------------------
constexpr int kStemCout = 64;

template <int KS>
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                        __nv_bfloat16* __restrict__ y, float* __restrict__ partials,
                                                        int n, int h, int wd, int stride, int ho, int wo) {
    constexpr int kTaps = 3 * KS * KS;
    constexpr int kPitch = kStemCout * 2 + 16;              // bytes per staged pixel row
    extern __shared__ __align__(16) uint8_t stem_smem[];
    float* ws = reinterpret_cast<float*>(stem_smem);        // [kTaps][64]
    uint8_t* tile = stem_smem + kTaps * kStemCout * sizeof(float);      // [128][kPitch]
    __shared__ float red[2][4][kStemCout];
    for (int i = threadIdx.x; i < kTaps * kStemCout; i += blockDim.x) {
        const int co = i % kStemCout, t = i / kStemCout;    // t = ci*KS*KS + r*KS + s
        ws[i] = w[co * kTaps + t];
    }
    float s_acc = 0.f, q_acc = 0.f;                          // thread t < 64: channel sums over this CTA's tiles
    float s_acc2 = 0.f, q_acc2 = 0.f;
    __syncthreads();
    const int64_t total = static_cast<int64_t>(n) * ho * wo;
    const int64_t tiles = (total + 127) / 128;
    constexpr int pad = KS / 2;
    for (int64_t tile_i = blockIdx.x; tile_i < tiles; tile_i += gridDim.x) {
        const int64_t pix = tile_i * 128 + threadIdx.x;
        const bool valid = pix < total;
        float acc[kStemCout];
#pragma unroll
        for (int j = 0; j < kStemCout; ++j) acc[j] = 0.f;
        if (valid) {
            const int ox = pix % wo, oy = (pix / wo) % ho, im = pix / (static_cast<int64_t>(wo) * ho);
            for (int ci = 0; ci < 3; ++ci) {
                const float* plane = img + (static_cast<int64_t>(im) * 3 + ci) * h * wd;
#pragma unroll
                for (int r = 0; r < KS; ++r) {
                    const int iy = oy * stride + r - pad;
                    if (iy < 0 || iy >= h) continue;
#pragma unroll
                    for (int sx = 0; sx < KS; ++sx) {
                        const int ix = ox * stride + sx - pad;
                        if (ix < 0 || ix >= wd) continue;
                        // the image is consumed at bf16 precision like every other activation
                        const float v = __bfloat162float(__float2bfloat16(plane[iy * wd + ix]));
                        const float4* wp = reinterpret_cast<const float4*>(ws + ((ci * KS + r) * KS + sx) * kStemCout);
#pragma unroll
                        for (int j = 0; j < kStemCout / 4; ++j) {
                            const float4 wv = wp[j];
                            acc[4 * j + 0] += v * wv.x; acc[4 * j + 1] += v * wv.y;
                            acc[4 * j + 2] += v * wv.z; acc[4 * j + 3] += v * wv.w;
                        }
                    }
                }
            }
        }
        uint4* trow = reinterpret_cast<uint4*>(tile + threadIdx.x * kPitch);
#pragma unroll
        for (int j = 0; j < kStemCout / 8; ++j) {
            uint4 o;
            o.x = pack_bf16(acc[8 * j + 0], acc[8 * j + 1]); o.y = pack_bf16(acc[8 * j + 2], acc[8 * j + 3]);
            o.z = pack_bf16(acc[8 * j + 4], acc[8 * j + 5]); o.w = pack_bf16(acc[8 * j + 6], acc[8 * j + 7]);
            trow[j] = o;
        }
        __syncthreads();
        const int64_t left = total - tile_i * 128;
        const int valid_rows = left < 128 ? static_cast<int>(left) : 128;
        {   // column statistics: 32 channel pairs x 4 row groups
            const int pair = threadIdx.x & 31, grp = threadIdx.x >> 5;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            for (int r = grp; r < valid_rows; r += 4) {
                const uint32_t u = *reinterpret_cast<const uint32_t*>(tile + r * kPitch + pair * 4);
                const float a = bf16_lo(u), b = bf16_hi(u);
                s0 += a; s1 += b; q0 += a * a; q1 += b * b;
            }
            red[0][grp][pair * 2] = s0; red[0][grp][pair * 2 + 1] = s1;
            red[1][grp][pair * 2] = q0; red[1][grp][pair * 2 + 1] = q1;
        }
        // coalesced stores: 8 x 16-byte chunks per pixel row
        for (int ch = threadIdx.x; ch < 128 * 8; ch += 128) {
            const int r = ch >> 3, cc = ch & 7;
            if (r < valid_rows)
                *reinterpret_cast<uint4*>(y + (tile_i * 128 + r) * kStemCout + cc * 8) =
                    *reinterpret_cast<const uint4*>(tile + r * kPitch + cc * 16);
        }
        __syncthreads();
        if (threadIdx.x < kStemCout) {
            const int c = threadIdx.x;
            s_acc += red[0][0][c] + red[0][1][c] + red[0][2][c] + red[0][3][c];
            q_acc += red[1][0][c] + red[1][1][c] + red[1][2][c] + red[1][3][c];
        }
        (void)s_acc2; (void)q_acc2;
    }
    if (partials != nullptr && threadIdx.x < kStemCout) {
        partials[static_cast<int64_t>(blockIdx.x) * 2 * kStemCout + threadIdx.x] = s_acc;
        partials[static_cast<int64_t>(blockIdx.x) * 2 * kStemCout + kStemCout + threadIdx.x] = q_acc;
    }
}

int stem_conv_tiles(int n, int ho, int wo) {
    const int64_t tiles = (static_cast<int64_t>(n) * ho * wo + 127) / 128;
    const int64_t grid = 148 * 4;
    return static_cast<int>(tiles < grid ? tiles : grid);
}

int stem_conv(const float* img, const float* w, void* y, float* partials, int n, int h, int wd, int cout, int ksize,
              int stride, cudaStream_t stream) {
    BPB_CHECK(ksize == 3 || ksize == 7, "stem_conv: ksize %d unsupported", ksize);
    BPB_CHECK(cout == kStemCout, "stem_conv: cout %d unsupported (the stems of HRNet / ResNet have 64 channels)", cout);
    const int pad = ksize / 2;
    const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (wd + 2 * pad - ksize) / stride + 1;
    const int grid = stem_conv_tiles(n, ho, wo);
    const size_t smem = static_cast<size_t>(3) * ksize * ksize * cout * sizeof(float) + 128 * (kStemCout * 2 + 16);
    if (ksize == 3) {
        stem_conv_kernel<3><<<grid, 128, smem, stream>>>(img, w, reinterpret_cast<__nv_bfloat16*>(y), partials, n, h, wd,
                                                         stride, ho, wo);
    } else {
        BPB_SET_MAX_SMEM(stem_conv_kernel<7>, 64 * 1024);
        stem_conv_kernel<7><<<grid, 128, smem, stream>>>(img, w, reinterpret_cast<__nv_bfloat16*>(y), partials, n, h, wd,
                                                         stride, ho, wo);
    }
    BPB_LAUNCH_CHECK("stem_conv_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Column statistics of a bf16 matrix [M, C] -> partials [tiles, 2, C] (deterministic two-stage reduction)
// ----------------------------------------------------------------------------------------------
constexpr int kStatRows = 256;   // rows per partial

__global__ void __launch_bounds__(256) colstats_bf16_kernel(const uint32_t* __restrict__ x, float* __restrict__ partials,
                                                            int m, int c) {
    // thread layout: tx over channel pairs, ty over row lanes
    const int pairs = c / 2;
    const int width = pairs < 256 ? pairs : 256;      // pairs handled per pass along x
    const int lanes = 256 / width;                    // row lanes
    const int tx = threadIdx.x % width, ty = threadIdx.x / width;
    const int r0 = blockIdx.x * kStatRows;
    const int r1 = min(m, r0 + kStatRows);
    __shared__ float red[4][256];
    for (int pbase = 0; pbase < pairs; pbase += width) {
        const int pr = pbase + tx;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        if (pr < pairs && ty < lanes) {
            for (int r = r0 + ty; r < r1; r += lanes) {
                const uint32_t u = x[static_cast<int64_t>(r) * pairs + pr];
                const float a = bf16_lo(u), b = bf16_hi(u);
                s0 += a; s1 += b; q0 += a * a; q1 += b * b;
            }
        }
        red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = q0; red[3][threadIdx.x] = q1;
        __syncthreads();
        if (ty == 0 && pr < pairs) {
            for (int l = 1; l < lanes; ++l) {
                s0 += red[0][l * width + tx]; s1 += red[1][l * width + tx];
                q0 += red[2][l * width + tx]; q1 += red[3][l * width + tx];
            }
            float* st = partials + static_cast<int64_t>(blockIdx.x) * 2 * c;
            st[2 * pr] = s0; st[2 * pr + 1] = s1;
            st[c + 2 * pr] = q0; st[c + 2 * pr + 1] = q1;
        }
        __syncthreads();
    }
}

int colstats_tiles(int m) { return (m + kStatRows - 1) / kStatRows; }

int colstats_bf16(const void* x, float* partials, int m, int c, int* tiles_out, cudaStream_t stream) {
    BPB_CHECK(c % 2 == 0, "colstats_bf16: C=%d must be even", c);
    BPB_CHECK((c / 2) <= 256 ? (256 % (c / 2) == 0) : true, "colstats_bf16: C/2=%d must divide 256 when < 256", c / 2);
    const int tiles = colstats_tiles(m);
    colstats_bf16_kernel<<<tiles, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(x), partials, m, c);
    BPB_LAUNCH_CHECK("colstats_bf16_kernel");
    if (tiles_out) *tiles_out = tiles;
    return 0;
}

// ----------------------------------------------------------------------------------------------
// partials [tiles,2,C] -> BN scale/shift (train-mode biased batch statistics, nn.BatchNorm semantics:
// running_var gets the unbiased estimate)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partials, int tiles, int c,
                                                          int c_lo, double inv_count, double unbias,
                                                          const float* gamma, const float* beta, float eps,
                                                          float momentum,
                                                          float* running_mean, float* running_var, float* scale,
                                                          float* shift, float* save) {
    const int cl = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + cl;
    __shared__ double rs[8][32], rq[8][32];
    double s = 0.0, q = 0.0;
    if (ch < c) {
        for (int t = j; t < tiles; t += 8) {
            s += static_cast<double>(partials[static_cast<int64_t>(t) * 2 * c + ch]);
            q += static_cast<double>(partials[static_cast<int64_t>(t) * 2 * c + c + ch]);
        }
    }
    rs[j][cl] = s; rq[j][cl] = q;
    __syncthreads();
    if (j == 0 && ch < c) {
        for (int l = 1; l < 8; ++l) { s += rs[l][cl]; q += rq[l][cl]; }
        const double mean = s * inv_count;
        double var = q * inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        // parameters cover channels [c_lo, c) only (IBN: the first c_lo channels are instance-normalised elsewhere)
        const int pc = ch - c_lo;
        const float g = (gamma && pc >= 0) ? gamma[pc] : 1.f, b = (beta && pc >= 0) ? beta[pc] : 0.f;
        const float istd = rsqrtf(static_cast<float>(var) + eps);
        const float sc = g * istd;
        scale[ch] = sc;
        shift[ch] = b - static_cast<float>(mean) * sc;
        if (save) { save[ch] = static_cast<float>(mean); save[c + ch] = istd; }
        if (running_mean && pc >= 0)
            running_mean[pc] = (1.f - momentum) * running_mean[pc] + momentum * static_cast<float>(mean);
        if (running_var && pc >= 0)
            running_var[pc] = (1.f - momentum) * running_var[pc] + momentum * static_cast<float>(var * unbias);
    }
}

int bn_finalize(const float* partials, int tiles, int c, int c_lo, int64_t count, const float* gamma,
                const float* beta, float eps, float momentum, float* running_mean, float* running_var, float* scale,
                float* shift, float* save_mean_invstd, cudaStream_t stream) {
    const double inv = 1.0 / static_cast<double>(count);
    const double unbias = count > 1 ? static_cast<double>(count) / static_cast<double>(count - 1) : 1.0;
    bn_finalize_kernel<<<ceil_div(c, 32), 256, 0, stream>>>(partials, tiles, c, c_lo, inv, unbias, gamma, beta, eps, momentum,
                                                            running_mean, running_var, scale, shift, save_mean_invstd);
    BPB_LAUNCH_CHECK("bn_finalize_kernel");
    return 0;
}

__global__ void bn_eval_fold_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                    int c, float* scale, float* shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float sc = (gamma ? gamma[i] : 1.f) * rsqrtf(rv[i] + eps);
    scale[i] = sc;
    shift[i] = (beta ? beta[i] : 0.f) - rm[i] * sc;
}

int bn_eval_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                 float eps, int c, float* scale, float* shift, cudaStream_t stream) {
    bn_eval_fold_kernel<<<ceil_div(c, 256), 256, 0, stream>>>(gamma, beta, running_mean, running_var, eps, c, scale,
                                                              shift);
    BPB_LAUNCH_CHECK("bn_eval_fold_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// y = act( x*scale + shift (+ res | + res*res_scale + res_shift) ), 8 channels (16 B) per thread
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    return o;
}

// eight consecutive per-channel constants as two 16-byte loads
__device__ __forceinline__ void load8f(const float* __restrict__ p, float* dst) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
    dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
}

// Thread layout: tx = channel vector (8 channels) -- fixed per thread, so scale/shift live in registers and the row
// loop has no division; ty = row lane.  Warps read/write contiguous 16-byte vectors of consecutive rows.
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const uint4* __restrict__ res,
                                                       const float* __restrict__ rscale,
                                                       const float* __restrict__ rshift, uint4* __restrict__ y,
                                                       int m, int cvec, int relu) {
    pdl_trigger();
    pdl_wait();
    const int width = cvec < 256 ? cvec : 256;
    const int rows_per_block = 256 / width;
    const int tx = threadIdx.x % width, ty = threadIdx.x / width;
    if (ty >= rows_per_block) return;
    for (int cv = tx; cv < cvec; cv += width) {
        float sc[8], sh[8], rs[8], rb[8];
        load8f(scale + cv * 8, sc);
        load8f(shift + cv * 8, sh);
        if (rscale) load8f(rscale + cv * 8, rs);
        if (rshift) load8f(rshift + cv * 8, rb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (!rscale) rs[j] = 1.f;
            if (!rshift) rb[j] = 0.f;
        }
        const int step = gridDim.x * rows_per_block;
        int r = blockIdx.x * rows_per_block + ty;
        // four rows per trip: up to eight independent 16-byte loads in flight per thread (one row per trip left the
        // pass latency-bound at ~3 TB/s)
        for (; r + 3 * step < m; r += 4 * step) {
            uint4 xv[4], rv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t i = static_cast<size_t>(r + u * step) * cvec + cv;
                xv[u] = __ldg(x + i);
                if (res != nullptr) rv[u] = __ldg(res + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float f[8];
                unpack8(xv[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
                if (res != nullptr) {
                    float rr[8];
                    unpack8(rv[u], rr);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += rr[j] * rs[j] + rb[j];
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                y[static_cast<size_t>(r + u * step) * cvec + cv] = pack8(f);
            }
        }
        for (; r < m; r += step) {
            const size_t i = static_cast<size_t>(r) * cvec + cv;
            float f[8];
            unpack8(x[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
            if (res != nullptr) {
                float rr[8];
                unpack8(res[i], rr);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += rr[j] * rs[j] + rb[j];
            }
            if (relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            y[i] = pack8(f);
        }
    }
}

int bn_apply(const void* x, const float* scale, const float* shift, const void* res, const float* res_scale,
             const float* res_shift, void* y, int64_t m, int c, int relu, cudaStream_t stream) {
    BPB_CHECK(c % 8 == 0, "bn_apply: C=%d must be a multiple of 8", c);
    BPB_CHECK(m < (1ll << 31), "bn_apply: too many rows");
    const int cvec = c / 8;
    const int width = cvec < 256 ? cvec : 256;
    const int rows_per_block = 256 / width;
    BPB_CHECK(((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift) |
                reinterpret_cast<uintptr_t>(res_scale) | reinterpret_cast<uintptr_t>(res_shift)) & 15) == 0,
              "bn_apply: scale / shift vectors must be 16-byte aligned");
    int blocks = ceil_div(m, static_cast<int64_t>(rows_per_block) * 4);   // one unrolled four-row trip per thread at least
    if (blocks > 148 * 8) blocks = 148 * 8;
    const bool pdl = m * c * 2 <= (8ll << 20);                      // short passes only (see launch_pdl)
    if (check_cuda(launch_pdl(pdl, bn_apply_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const uint4*>(x), scale,
                              shift, reinterpret_cast<const uint4*>(res), res_scale, res_shift,
                              reinterpret_cast<uint4*>(y), static_cast<int>(m), cvec, relu),
                   "bn_apply_kernel"))
        return 3;
    BPB_LAUNCH_CHECK("bn_apply_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Cross-resolution fusion (hrnet.py:269-277): y = relu( sum_j term_j ),
// term_j = src_j[n, h>>sh_j, w>>sh_j, c] * scale_j[c] + shift_j[c]   (scale_j == null -> identity)
// ----------------------------------------------------------------------------------------------
struct FuseArgs {
    const uint4* src[4];
    const float* scale[4];
    const float* shift[4];
    int sh[4];
    int nsrc;
};

__global__ void __launch_bounds__(256) fuse_sum_kernel(FuseArgs a, uint4* __restrict__ y, int n, int h, int wd, int cvec,
                                                       int relu) {
    pdl_trigger();
    pdl_wait();
    const int width = cvec < 256 ? cvec : 256;
    const int rows_per_block = 256 / width;
    const int tx = threadIdx.x % width, ty = threadIdx.x / width;
    if (ty >= rows_per_block) return;
    const unsigned m = static_cast<unsigned>(n) * h * wd;
    for (int cv = tx; cv < cvec; cv += width) {
        for (unsigned r = blockIdx.x * rows_per_block + ty; r < m; r += gridDim.x * rows_per_block) {
            const unsigned ox = r % wd, t = r / wd;
            const unsigned oy = t % h, im = t / h;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < a.nsrc) {
                    const int sh = a.sh[j];
                    const unsigned hs = h >> sh, ws = wd >> sh;
                    const size_t sidx = (static_cast<size_t>(im * hs + (oy >> sh)) * ws + (ox >> sh)) * cvec + cv;
                    float f[8];
                    unpack8(a.src[j][sidx], f);
                    if (a.scale[j] != nullptr) {
                        const float* sc = a.scale[j] + cv * 8;
                        const float* sf = a.shift[j] + cv * 8;
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[q] += f[q] * sc[q] + sf[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[q] += f[q];
                    }
                }
            }
            if (relu) {
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = fmaxf(acc[q], 0.f);
            }
            y[static_cast<size_t>(r) * cvec + cv] = pack8(acc);
        }
    }
}

int fuse_sum(const void* const* srcs, const float* const* scales, const float* const* shifts, const int* shifts_log2,
             int nsrc, void* y, int n, int h, int wd, int c, int relu, cudaStream_t stream) {
    BPB_CHECK(nsrc >= 1 && nsrc <= 4, "fuse_sum: nsrc=%d out of range", nsrc);
    BPB_CHECK(c % 8 == 0, "fuse_sum: C=%d must be a multiple of 8", c);
    FuseArgs a{};
    a.nsrc = nsrc;
    for (int j = 0; j < nsrc; ++j) {
        a.src[j] = reinterpret_cast<const uint4*>(srcs[j]);
        a.scale[j] = scales ? scales[j] : nullptr;
        a.shift[j] = shifts ? shifts[j] : nullptr;
        a.sh[j] = shifts_log2[j];
        BPB_CHECK((h % (1 << a.sh[j])) == 0 && (wd % (1 << a.sh[j])) == 0, "fuse_sum: bad upsample factor");
    }
    const int cvec = c / 8;
    const int rows_per_block = 256 / (cvec < 256 ? cvec : 256);
    int blocks = ceil_div(static_cast<int64_t>(n) * h * wd, rows_per_block);
    if (blocks > 148 * 8) blocks = 148 * 8;
    const bool pdl = static_cast<int64_t>(n) * h * wd * c * 2 <= (8ll << 20);
    if (check_cuda(launch_pdl(pdl, fuse_sum_kernel, dim3(blocks), dim3(256), 0, stream, a, reinterpret_cast<uint4*>(y), n, h, wd,
                              cvec, relu),
                   "fuse_sum_kernel"))
        return 3;
    BPB_LAUNCH_CHECK("fuse_sum_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Bilinear (align_corners=True) upsample of up to 4 branch maps to (h, wd) + channel concat (hrnet.py:568-573)
// ----------------------------------------------------------------------------------------------
struct UpcatArgs {
    const uint4* src[4];
    int cvec[4], cvec_off[5], hs[4], ws[4];
    int nsrc;
};

__global__ void __launch_bounds__(256) upsample_concat_kernel(UpcatArgs a, uint4* __restrict__ y, int n, int h, int wd,
                                                              int ctot_vec) {
    const int64_t total = static_cast<int64_t>(n) * h * wd * ctot_vec;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int cv = i % ctot_vec;
        const int64_t pix = i / ctot_vec;
        const int ox = pix % wd, oy = (pix / wd) % h, im = pix / (static_cast<int64_t>(wd) * h);
        int b = 0;
#pragma unroll
        for (int j = 1; j < 4; ++j)
            if (j < a.nsrc && cv >= a.cvec_off[j]) b = j;
        const int lc = cv - a.cvec_off[b];
        const int hs = a.hs[b], ws = a.ws[b], cvb = a.cvec[b];
        const uint4* s = a.src[b] + static_cast<int64_t>(im) * hs * ws * cvb;
        if (hs == h && ws == wd) {
            y[i] = s[(static_cast<int64_t>(oy) * ws + ox) * cvb + lc];
            continue;
        }
        // PyTorch: src = dst * (in-1)/(out-1) ; i0 = floor, i1 = min(i0+1, in-1), lambda = src - i0
        const float ry = h > 1 ? static_cast<float>(hs - 1) / static_cast<float>(h - 1) : 0.f;
        const float rx = wd > 1 ? static_cast<float>(ws - 1) / static_cast<float>(wd - 1) : 0.f;
        const float sy = ry * oy, sx = rx * ox;
        const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
        const int y1 = y0 + (y0 < hs - 1 ? 1 : 0), x1 = x0 + (x0 < ws - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
        float f00[8], f01[8], f10[8], f11[8], o[8];
        unpack8(s[(static_cast<int64_t>(y0) * ws + x0) * cvb + lc], f00);
        unpack8(s[(static_cast<int64_t>(y0) * ws + x1) * cvb + lc], f01);
        unpack8(s[(static_cast<int64_t>(y1) * ws + x0) * cvb + lc], f10);
        unpack8(s[(static_cast<int64_t>(y1) * ws + x1) * cvb + lc], f11);
#pragma unroll
        for (int t = 0; t < 8; ++t)
            o[t] = (1.f - ly) * ((1.f - lx) * f00[t] + lx * f01[t]) + ly * ((1.f - lx) * f10[t] + lx * f11[t]);
        y[i] = pack8(o);
    }
}

int upsample_concat(const void* const* srcs, const int* chans, const int* hs, const int* ws, int nsrc, void* y, int n,
                    int h, int wd, cudaStream_t stream) {
    BPB_CHECK(nsrc >= 1 && nsrc <= 4, "upsample_concat: nsrc=%d out of range", nsrc);
    UpcatArgs a{};
    a.nsrc = nsrc;
    int off = 0;
    for (int j = 0; j < nsrc; ++j) {
        BPB_CHECK(chans[j] % 8 == 0, "upsample_concat: channels must be multiples of 8");
        a.src[j] = reinterpret_cast<const uint4*>(srcs[j]);
        a.cvec[j] = chans[j] / 8;
        a.cvec_off[j] = off;
        a.hs[j] = hs[j];
        a.ws[j] = ws[j];
        off += chans[j] / 8;
    }
    a.cvec_off[nsrc] = off;
    const int64_t total = static_cast<int64_t>(n) * h * wd * off;
    int blocks = ceil_div(total, 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    upsample_concat_kernel<<<blocks, 256, 0, stream>>>(a, reinterpret_cast<uint4*>(y), n, h, wd, off);
    BPB_LAUNCH_CHECK("upsample_concat_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// 3x3 stride-2 max pool (resnet.py:216 pad=1 ; resnet_fastreid.py:251 pad=0, ceil_mode=True)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h,
                                                      int wd, int cvec, int pad, int ho, int wo) {
    const int64_t total = static_cast<int64_t>(n) * ho * wo * cvec;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int cv = i % cvec;
        const int64_t pix = i / cvec;
        const int ox = pix % wo, oy = (pix / wo) % ho, im = pix / (static_cast<int64_t>(wo) * ho);
        float best[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) best[t] = -INFINITY;
        for (int r = 0; r < 3; ++r) {
            const int iy = oy * 2 + r - pad;
            if (iy < 0 || iy >= h) continue;
            for (int s = 0; s < 3; ++s) {
                const int ix = ox * 2 + s - pad;
                if (ix < 0 || ix >= wd) continue;
                float f[8];
                unpack8(x[((static_cast<int64_t>(im) * h + iy) * wd + ix) * cvec + cv], f);
#pragma unroll
                for (int t = 0; t < 8; ++t) best[t] = fmaxf(best[t], f[t]);
            }
        }
        y[i] = pack8(best);
    }
}

int maxpool3x3s2(const void* x, void* y, int n, int h, int wd, int c, int pad, int ceil_mode, cudaStream_t stream) {
    BPB_CHECK(c % 8 == 0, "maxpool: C=%d must be a multiple of 8", c);
    auto osz = [&](int in) {
        const int num = in + 2 * pad - 3;
        int o = (ceil_mode ? (num + 1) / 2 : num / 2) + 1;
        if (ceil_mode && (o - 1) * 2 >= in + pad) --o;       // PyTorch: last window must start inside the input
        return o;
    };
    const int ho = osz(h), wo = osz(wd);
    const int64_t total = static_cast<int64_t>(n) * ho * wo * (c / 8);
    int blocks = ceil_div(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    maxpool_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), n, h, wd,
                                               c / 8, pad, ho, wo);
    BPB_LAUNCH_CHECK("maxpool_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Instance-norm statistics for the IBN half (resnet_fastreid.py:108-122): per (image, channel) mean / rstd over HW
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const __nv_bfloat16* __restrict__ x,
                                                             float* __restrict__ mean_rstd, int hw, int c, int c_lo,
                                                             int c_hi, float eps) {
    // block = (image, 32-channel group); threads: 32 channels x 8 row lanes
    const int cl = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int ch = c_lo + blockIdx.y * 32 + cl;
    const int im = blockIdx.x;
    __shared__ float rs[8][32], rq[8][32];
    float s = 0.f, q = 0.f;
    if (ch < c_hi) {
        const __nv_bfloat16* base = x + static_cast<int64_t>(im) * hw * c + ch;
        for (int p = j; p < hw; p += 8) {
            const float v = __bfloat162float(base[static_cast<int64_t>(p) * c]);
            s += v; q += v * v;
        }
    }
    rs[j][cl] = s; rq[j][cl] = q;
    __syncthreads();
    if (j == 0 && ch < c_hi) {
        for (int l = 1; l < 8; ++l) { s += rs[l][cl]; q += rq[l][cl]; }
        const float mean = s / hw;
        float var = q / hw - mean * mean;
        var = fmaxf(var, 0.f);
        const int nch = c_hi - c_lo;
        mean_rstd[(static_cast<int64_t>(im) * nch + (ch - c_lo)) * 2 + 0] = mean;
        mean_rstd[(static_cast<int64_t>(im) * nch + (ch - c_lo)) * 2 + 1] = rsqrtf(var + eps);
    }
}

int instnorm_stats(const void* x, float* mean_rstd, int n, int hw, int c, int c_lo, int c_hi, float eps,
                   cudaStream_t stream) {
    dim3 grid(n, ceil_div(c_hi - c_lo, 32));
    instnorm_stats_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), mean_rstd, hw, c, c_lo,
                                                    c_hi, eps);
    BPB_LAUNCH_CHECK("instnorm_stats_kernel");
    return 0;
}

// y = relu( IBN(x) ): channels [0, half) instance-normalised (per image), [half, c) batch-normalised
__global__ void __launch_bounds__(256) ibn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ in_mr,
                                                        const float* __restrict__ in_g, const float* __restrict__ in_b,
                                                        const float* __restrict__ bn_scale,
                                                        const float* __restrict__ bn_shift, uint4* __restrict__ y,
                                                        int n, int hw, int cvec, int half, int relu) {
    const int64_t total = static_cast<int64_t>(n) * hw * cvec;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int cv = i % cvec;
        const int im = i / (static_cast<int64_t>(hw) * cvec);
        float f[8];
        unpack8(x[i], f);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int ch = cv * 8 + t;
            if (ch < half) {
                const float mean = in_mr[(static_cast<int64_t>(im) * half + ch) * 2 + 0];
                const float rstd = in_mr[(static_cast<int64_t>(im) * half + ch) * 2 + 1];
                f[t] = (f[t] - mean) * rstd * in_g[ch] + in_b[ch];
            } else {
                f[t] = f[t] * bn_scale[ch] + bn_shift[ch];
            }
            if (relu) f[t] = fmaxf(f[t], 0.f);
        }
        y[i] = pack8(f);
    }
}

int ibn_apply(const void* x, const float* in_mean_rstd, const float* in_gamma, const float* in_beta,
              const float* bn_scale, const float* bn_shift, void* y, int n, int hw, int c, int half, int relu,
              cudaStream_t stream) {
    BPB_CHECK(c % 8 == 0, "ibn_apply: C=%d must be a multiple of 8", c);
    const int64_t total = static_cast<int64_t>(n) * hw * (c / 8);
    int blocks = ceil_div(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    ibn_apply_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), in_mean_rstd, in_gamma, in_beta,
                                                 bn_scale, bn_shift, reinterpret_cast<uint4*>(y), n, hw, c / 8, half,
                                                 relu);
    BPB_LAUNCH_CHECK("ibn_apply_kernel");
    return 0;
}

}  // namespace bpb
