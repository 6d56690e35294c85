// BPBreID head: pixel->part classifier (BN folded into the 1x1 conv), softmax attention, visibility,
// the (K+3)-row weighted pooling of the feature map in ONE pass, and BN1d for the tiny per-embedding layers.
// Reference semantics: torchreid/models/bpbreid.py:147-148 (classifier+softmax), :157-192 (masks, visibility),
// :195-202 + :458-503 (global / fg / bg / parts pooling), :324-350, :398-415 (BN1d of dim-reduce / classifiers).
// The reference materialises a [N,K,C,H,W] product (bpbreid.py:493); here the map is read once.
#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// wf[k][c] = w[k][c]*scale[c] ; bf[k] = b[k] + sum_c w[k][c]*shift[c]      (BN2d folded into the 1x1 conv)
__global__ void __launch_bounds__(256) pixel_fold_kernel(const float* w, const float* b, const float* scale,
                                                         const float* shift, int c, float* wf, float* bf) {
    const int k = blockIdx.x;
    float acc = 0.f;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const float wv = w[k * c + i];
        wf[k * c + i] = wv * scale[i];
        acc += wv * shift[i];
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        bf[k] = (b ? b[k] : 0.f) + t;
    }
}

int pixel_fold(const float* w, const float* b, const float* scale, const float* shift, int kp1, int c, float* wf,
               float* bf, cudaStream_t stream) {
    pixel_fold_kernel<<<kp1, 256, 0, stream>>>(w, b, scale, shift, c, wf, bf);
    BPB_LAUNCH_CHECK("pixel_fold_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// One warp classifies 4 pixels at a time: logits -> softmax -> masks / visibility.
// `ext` (optional) = external part masks already resized to the feature map, `mode` bits:
//   1  external masks ARE the attention (learnable_attention_enabled False, bpbreid.py:149-154): no classifier, no logits
//   2  test_use_target_segmentation 'hard' (:160-167)     4  'soft' (:169-174)
// ----------------------------------------------------------------------------------------------
constexpr int kMaxKp1 = 12;
constexpr int kPixPerWarp = 4;

template <int KP1>
__global__ void __launch_bounds__(256) pixel_classify_kernel(const uint4* __restrict__ feats,
                                                             const float* __restrict__ wf_g,
                                                             const float* __restrict__ bf_g, int n, int hw, int c,
                                                             float* __restrict__ logits, float* __restrict__ probs,
                                                             float* __restrict__ fg_mask, int* __restrict__ vis_bits,
                                                             float* __restrict__ vis_cont,
                                                             const float* __restrict__ ext, int mode) {
    extern __shared__ float wf[];                     // [KP1][c]
    const bool learn = (mode & 1) == 0;               // mode bit 0: external masks ARE the attention (no classifier)
    if (learn)
        for (int i = threadIdx.x; i < KP1 * c; i += blockDim.x) wf[i] = wf_g[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cvec = c / 8;
    const int64_t total = static_cast<int64_t>(n) * hw;
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    for (int64_t base = (static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp) * kPixPerWarp; base < total;
         base += warps_total * kPixPerWarp) {
        float acc[kPixPerWarp][KP1];
#pragma unroll
        for (int p = 0; p < kPixPerWarp; ++p)
#pragma unroll
            for (int k = 0; k < KP1; ++k) acc[p][k] = 0.f;
        for (int cv = lane; learn && cv < cvec; cv += 32) {
            float f[kPixPerWarp][8];
#pragma unroll
            for (int p = 0; p < kPixPerWarp; ++p) {
                const int64_t pix = base + p < total ? base + p : total - 1;
                const uint4 u = feats[pix * cvec + cv];
                f[p][0] = bf16_lo(u.x); f[p][1] = bf16_hi(u.x); f[p][2] = bf16_lo(u.y); f[p][3] = bf16_hi(u.y);
                f[p][4] = bf16_lo(u.z); f[p][5] = bf16_hi(u.z); f[p][6] = bf16_lo(u.w); f[p][7] = bf16_hi(u.w);
            }
#pragma unroll
            for (int k = 0; k < KP1; ++k) {
                const float4 w0 = *reinterpret_cast<const float4*>(wf + k * c + cv * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(wf + k * c + cv * 8 + 4);
#pragma unroll
                for (int p = 0; p < kPixPerWarp; ++p) {
                    acc[p][k] += f[p][0] * w0.x + f[p][1] * w0.y + f[p][2] * w0.z + f[p][3] * w0.w + f[p][4] * w1.x +
                                 f[p][5] * w1.y + f[p][6] * w1.z + f[p][7] * w1.w;
                }
            }
        }
#pragma unroll
        for (int p = 0; p < kPixPerWarp; ++p)
#pragma unroll
            for (int k = 0; k < KP1; ++k) acc[p][k] = warp_sum(acc[p][k]);
        // lane p finishes pixel p
#pragma unroll
        for (int p = 0; p < kPixPerWarp; ++p) {
            if (lane == p && base + p < total) {
                const int64_t pix = base + p;
                const int im = pix / hw, loc = pix % hw;
                float pr[KP1], pm[KP1], ex[KP1];
                if (ext != nullptr) {
#pragma unroll
                    for (int k = 0; k < KP1; ++k) ex[k] = ext[(static_cast<int64_t>(im) * KP1 + k) * hw + loc];
                }
                if (learn) {
                    float l[KP1], mx = -INFINITY;
#pragma unroll
                    for (int k = 0; k < KP1; ++k) { l[k] = acc[p][k] + bf_g[k]; mx = fmaxf(mx, l[k]); }
                    float e[KP1], den = 0.f;
#pragma unroll
                    for (int k = 0; k < KP1; ++k) { e[k] = expf(l[k] - mx); den += e[k]; }
                    const float inv = 1.f / den;
#pragma unroll
                    for (int k = 0; k < KP1; ++k) {
                        pr[k] = e[k] * inv;
                        logits[(static_cast<int64_t>(im) * KP1 + k) * hw + loc] = l[k];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KP1; ++k) pr[k] = ex[k];
                }
                apply_target_segmentation<KP1>(pr, pm, ex, mode);
                int arg = 0;
                float best = -1.f, fg = -1.f;
#pragma unroll
                for (int k = 0; k < KP1; ++k) {
                    probs[(static_cast<int64_t>(im) * KP1 + k) * hw + loc] = pm[k];
                    if (pr[k] > best) { best = pr[k]; arg = k; }   // first maximum, like torch.argmax
                    if (k >= 1) fg = fmaxf(fg, pm[k]);
                    atomicMax(reinterpret_cast<int*>(vis_cont + static_cast<int64_t>(im) * KP1 + k), __float_as_int(pr[k]));
                }
                fg_mask[pix] = fg;
                atomicOr(vis_bits + im, 1 << arg);
            }
        }
    }
}

int pixel_classify(const void* feats, const float* wf, const float* bf, int n, int hw, int c, int kp1,
                   float* logits_nchw, float* probs, float* fg_mask, int* vis_bits, float* vis_cont, const float* ext,
                   int mode, cudaStream_t stream) {
    BPB_CHECK(mode == 0 || ext != nullptr, "pixel_classify: mode %d needs the external masks", mode);
    BPB_CHECK(c % 8 == 0, "pixel_classify: C=%d must be a multiple of 8", c);
    BPB_CHECK(kp1 >= 2 && kp1 <= kMaxKp1, "pixel_classify: K+1=%d out of range [2,%d]", kp1, kMaxKp1);
    BPB_CUDA(cudaMemsetAsync(vis_bits, 0, sizeof(int) * n, stream));
    BPB_CUDA(cudaMemsetAsync(vis_cont, 0, sizeof(float) * n * kp1, stream));
    const size_t smem = static_cast<size_t>(kp1) * c * sizeof(float);
    const int64_t total = static_cast<int64_t>(n) * hw;
    int blocks = ceil_div(total, 8 * kPixPerWarp);
    if (blocks > 148 * 2) blocks = 148 * 2;
#define BPB_PC_CASE(K)                                                                                              \
    case K: {                                                                                                       \
        auto kern = pixel_classify_kernel<K>;                                                                       \
        BPB_SET_MAX_SMEM(kern, 160 * 1024);                                                                                                           \
        kern<<<blocks, 256, smem, stream>>>(reinterpret_cast<const uint4*>(feats), wf, bf, n, hw, c, logits_nchw,   \
                                            probs, fg_mask, vis_bits, vis_cont, ext, mode);                         \
        break;                                                                                                      \
    }
    switch (kp1) {
        BPB_PC_CASE(2) BPB_PC_CASE(3) BPB_PC_CASE(4) BPB_PC_CASE(5) BPB_PC_CASE(6) BPB_PC_CASE(7) BPB_PC_CASE(8)
        BPB_PC_CASE(9) BPB_PC_CASE(10) BPB_PC_CASE(11) BPB_PC_CASE(12)
    }
#undef BPB_PC_CASE
    BPB_LAUNCH_CHECK("pixel_classify_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// pooled[j, n, c] for j = 0:global, 1:foreground, 2:background, 3..: parts     (one pass over the map)
//   global      = mean_p f                          (bpbreid.py:195)
//   fg / bg     = mean_p (m * f)   -- divided by H*W, NOT by the mask sum (bpbreid.py:485-486)
//   part k      = sum_p m_k f / clamp(sum_p m_k, 1e-6)                 (bpbreid.py:489-503)
// CTA = (image, 256-channel slab); 4 warps split the pixels; lane owns 8 channels.
// ----------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(128) part_pool_kernel(const uint4* __restrict__ feats, const float* __restrict__ probs,
                                                        const float* __restrict__ fg_mask, int n, int hw, int c,
                                                        float* __restrict__ pooled) {
    constexpr int KP1 = J - 2;
    extern __shared__ float wts[];                      // [hw][J-1] : fg, bg, part_0.. (global row needs no weight)
    const int im = blockIdx.x, slab = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cvec = c / 8;
    const int cv = slab * 32 + lane;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        wts[i * (J - 1) + 0] = fg_mask[static_cast<int64_t>(im) * hw + i];
#pragma unroll
        for (int k = 0; k < KP1; ++k) wts[i * (J - 1) + 1 + k] = probs[(static_cast<int64_t>(im) * KP1 + k) * hw + i];
    }
    __syncthreads();
    float acc[J][8];
    float wsum[J - 1];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[j][t] = 0.f;
#pragma unroll
    for (int j = 0; j < J - 1; ++j) wsum[j] = 0.f;
    if (cv < cvec) {
        const uint4* base = feats + static_cast<int64_t>(im) * hw * cvec + cv;
        for (int p = warp; p < hw; p += 4) {
            const uint4 u = base[static_cast<int64_t>(p) * cvec];
            float f[8];
            f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
            f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[0][t] += f[t];
#pragma unroll
            for (int j = 0; j < J - 1; ++j) {
                const float w = wts[p * (J - 1) + j];
                wsum[j] += w;
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[j + 1][t] += w * f[t];
            }
        }
    }
    // cross-warp reduction through shared memory (reuse wts after everyone is done reading it)
    __syncthreads();
    float* red = wts;                                   // [4][32][J*8 + (J-1)]
    constexpr int kPer = J * 8 + (J - 1);
    float* mine = red + (warp * 32 + lane) * kPer;
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int t = 0; t < 8; ++t) mine[j * 8 + t] = acc[j][t];
#pragma unroll
    for (int j = 0; j < J - 1; ++j) mine[J * 8 + j] = wsum[j];
    __syncthreads();
    if (warp == 0 && cv < cvec) {
        const float inv_hw = 1.f / hw;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            float norm = inv_hw;
            if (j >= 3) {
                float s = 0.f;
                for (int w = 0; w < 4; ++w) s += red[(w * 32 + lane) * kPer + J * 8 + (j - 1)];
                norm = 1.f / fmaxf(s, 1e-6f);
            }
            float o[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float s = 0.f;
                for (int w = 0; w < 4; ++w) s += red[(w * 32 + lane) * kPer + j * 8 + t];
                o[t] = s * norm;
            }
            float* dst = pooled + (static_cast<int64_t>(j) * n + im) * c + cv * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

int part_pool(const void* feats, const float* probs, const float* fg_mask, int n, int hw, int c, int kp1,
              float* pooled, cudaStream_t stream) {
    BPB_CHECK(c % 8 == 0, "part_pool: C=%d must be a multiple of 8", c);
    BPB_CHECK(kp1 >= 2 && kp1 <= kMaxKp1, "part_pool: K+1=%d out of range", kp1);
    const int J = kp1 + 2;
    size_t smem = static_cast<size_t>(hw) * (J - 1) * sizeof(float);
    const size_t red = static_cast<size_t>(128) * (J * 8 + J - 1) * sizeof(float);
    if (red > smem) smem = red;
    BPB_CHECK(smem <= 200 * 1024, "part_pool: feature map %d px x %d rows does not fit shared memory", hw, J);
    dim3 grid(n, ceil_div(c / 8, 32));
#define BPB_PP_CASE(JJ)                                                                                         \
    case JJ: {                                                                                                  \
        auto kern = part_pool_kernel<JJ>;                                                                       \
        BPB_SET_MAX_SMEM(kern, 200 * 1024);                                                                                                       \
        kern<<<grid, 128, smem, stream>>>(reinterpret_cast<const uint4*>(feats), probs, fg_mask, n, hw, c, pooled); \
        break;                                                                                                  \
    }
    switch (J) {
        BPB_PP_CASE(4) BPB_PP_CASE(5) BPB_PP_CASE(6) BPB_PP_CASE(7) BPB_PP_CASE(8) BPB_PP_CASE(9) BPB_PP_CASE(10)
        BPB_PP_CASE(11) BPB_PP_CASE(12) BPB_PP_CASE(13) BPB_PP_CASE(14)
    }
#undef BPB_PP_CASE
    BPB_LAUNCH_CHECK("part_pool_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// BatchNorm1d on a small fp32 matrix [M, C] (+ optional ReLU).  grid = C/32, block = 32 cols x 8 row lanes.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn1d_kernel(const float* __restrict__ x, const float* gamma, const float* beta,
                                                   float eps, float momentum, float* running_mean, float* running_var,
                                                   float* __restrict__ y, int m, int c, int relu, int training) {
    const int cl = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + cl;
    __shared__ float red[8][32];
    __shared__ float s_mean[32], s_rstd[32];
    float mean = 0.f, rstd = 1.f;
    if (training) {
        float s = 0.f;
        if (ch < c)
            for (int r = j; r < m; r += 8) s += x[static_cast<int64_t>(r) * c + ch];
        red[j][cl] = s;
        __syncthreads();
        if (j == 0) {
            for (int l = 1; l < 8; ++l) s += red[l][cl];
            s_mean[cl] = s / m;
        }
        __syncthreads();
        mean = s_mean[cl];
        float q = 0.f;
        if (ch < c)
            for (int r = j; r < m; r += 8) {
                const float d = x[static_cast<int64_t>(r) * c + ch] - mean;
                q += d * d;
            }
        red[j][cl] = q;
        __syncthreads();
        if (j == 0) {
            for (int l = 1; l < 8; ++l) q += red[l][cl];
            const float var = q / m;
            s_rstd[cl] = rsqrtf(var + eps);
            if (ch < c) {
                if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mean;
                if (running_var)
                    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (m > 1 ? q / (m - 1) : var);
            }
        }
        __syncthreads();
        rstd = s_rstd[cl];
    } else if (ch < c) {
        mean = running_mean[ch];
        rstd = rsqrtf(running_var[ch] + eps);
    }
    if (ch < c) {
        const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
        for (int r = j; r < m; r += 8) {
            float v = (x[static_cast<int64_t>(r) * c + ch] - mean) * rstd * g + b;
            if (relu) v = fmaxf(v, 0.f);
            y[static_cast<int64_t>(r) * c + ch] = v;
        }
    }
}

int bn1d_train(const float* x, const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
               float* running_var, float* y, int m, int c, int relu, int training, cudaStream_t stream) {
    BPB_CHECK(training || (running_mean && running_var), "bn1d: eval mode needs running statistics");
    bn1d_kernel<<<ceil_div(c, 32), 256, 0, stream>>>(x, gamma, beta, eps, momentum, running_mean, running_var, y, m, c,
                                                     relu, training);
    BPB_LAUNCH_CHECK("bn1d_kernel");
    return 0;
}

// vis_out [n, K+2] uint8 (torch.bool storage) / vis_out_f [n, K+2] float: columns {0: background, 1..K: parts, K+1: foreground}
// binary: "class k is the argmax of at least one pixel"; foreground = any class incl. background (bpbreid.py:189 takes
// the amax over all K+1 columns before slicing).  continuous: amax of the probabilities.
__global__ void vis_unpack_kernel(const int* __restrict__ vis_bits, const float* __restrict__ vis_cont, int n, int kp1,
                                  uint8_t* __restrict__ out_b, float* __restrict__ out_f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bits = vis_bits[i];
    float fmx = 0.f;
    for (int k = 0; k < kp1; ++k) {
        out_b[i * (kp1 + 1) + k] = (bits >> k) & 1;
        const float v = vis_cont[i * kp1 + k];
        out_f[i * (kp1 + 1) + k] = v;
        fmx = fmaxf(fmx, v);
    }
    out_b[i * (kp1 + 1) + kp1] = bits != 0;
    out_f[i * (kp1 + 1) + kp1] = fmx;
}

int vis_unpack(const int* vis_bits, const float* vis_cont, int n, int kp1, uint8_t* out_b, float* out_f,
               cudaStream_t stream) {
    vis_unpack_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(vis_bits, vis_cont, n, kp1, out_b, out_f);
    BPB_LAUNCH_CHECK("vis_unpack_kernel");
    return 0;
}

}  // namespace bpb
