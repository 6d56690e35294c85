// Backward of the BPBreID attention head evaluated on the un-concatenated HRNet branch maps (csrc/head_lowres.cu is
// the forward; reference semantics torchreid/models/bpbreid.py:147-202, 376-385, 458-503 and the bilinear
// upsample + concat of torchreid/models/hrnet.py:568-573, differentiated by autograd in the reference).
//
// Forward (per branch b with low-resolution map X_b, bilinear operator U_b, folded classifier Wf = Wc * scale):
//   logit = bf + sum_b U_b (X_b Wf_b^T)       prob = softmax(logit)        masks m_j in {1, max parts, bg, part_k}
//   S_j   = sum_b (U_b^T m_j)^T X_b           pooled_j = norm_j * S_j      norm = 1/HW (j<3) | 1/clamp(sum m_j, 1e-6)
//   BN2d statistics (mean, var) of the virtual upsampled map feed scale/shift of the fold.
// Backward, given G = dL/dpooled (and optionally the pixel-wise part-prediction loss):
//   (1) Gs = norm * G ; d(sum m_j) from the clamp quotient                                -> pool_bwd_prep
//   (2) dM'_b[n,q,j] = <Gs[j,n,:], X_b[n,q,:]>                                             -> head_dmask
//   (3) upsample dM', masks -> softmax backward (+ pixel CE) -> dlogit; adjoint upsample   -> head_pixel_bwd
//   (4) dWf += dLow_b^T X_b                                                                -> head_dwf
//   (5) fold backward: dWc, dbc, dgamma, dbeta and the coefficients of the BN-statistics path -> pixel_fold_bwd
//   (6) dX_b = sum_j M'_b[j] Gs[j] + dLow_b Wf_b + a_c colsum(U_b) + b_c (U_b^T U_b X_b)   -> head_branch_bwd
// Nothing at 1920 channels x stride 4 is materialised in the backward either.
#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

struct TapB {
    int i0, i1;
    float l;
};
__device__ __forceinline__ TapB make_tap_b(int o, int in_size, int out_size) {
    const float r = out_size > 1 ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.f;
    const float s = r * o;
    TapB t;
    t.i0 = static_cast<int>(s);
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l = s - t.i0;
    return t;
}
__device__ __forceinline__ void unpack8h(const uint4& u, float* f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// ----------------------------------------------------------------------------------------------
// (1) in place: g[j,n,:] *= norm_j(n) ; dms[n,j] = -<g_in[j,n,:], pooled[j,n,:]> / clamp(ms) for part rows whose
//     mask sum is above the clamp (zero otherwise and for j < 3).  grid (J, n).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_bwd_prep_kernel(float* __restrict__ g, const float* __restrict__ pooled,
                                                            const float* __restrict__ mask_sum, int n, int c, int J,
                                                            float inv_hw, float* __restrict__ dms) {
    const int j = blockIdx.x, im = blockIdx.y;
    float* gr = g + (static_cast<int64_t>(j) * n + im) * c;
    const float* pr = pooled + (static_cast<int64_t>(j) * n + im) * c;
    const float ms = mask_sum[im * J + j];
    const bool part = j >= 3;
    const float norm = part ? 1.f / fmaxf(ms, 1e-6f) : inv_hw;
    float dot = 0.f;
    for (int i = threadIdx.x; i < c; i += 256) {
        const float gv = gr[i];
        dot += gv * pr[i];
        gr[i] = gv * norm;
    }
    __shared__ float red[8];
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        dms[im * J + j] = (part && ms > 1e-6f) ? -t * norm : 0.f;
    }
}

int head_pool_bwd_prep(float* g, const float* pooled, const float* mask_sum, int n, int c, int kp1, int hw_full,
                       float* dms, cudaStream_t stream) {
    const int J = kp1 + 2;
    dim3 grid(J, n);
    pool_bwd_prep_kernel<<<grid, 256, 0, stream>>>(g, pooled, mask_sum, n, c, J, 1.f / static_cast<float>(hw_full), dms);
    BPB_LAUNCH_CHECK("pool_bwd_prep_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// (2) out[n, q, j] = sum_c gs[j, n, c_off + c] * x[n, q, c]     (per-image weight rows staged in shared memory)
//     grid (n, pixel chunks); one warp handles 4 pixels at a time.
// ----------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(256) head_dmask_kernel(const uint4* __restrict__ x, const float* __restrict__ gs,
                                                         int n, int q, int cb, int c_off, int ctot,
                                                         float* __restrict__ out) {
    extern __shared__ float wrow[];                       // [J][cb]
    const int im = blockIdx.x;
    for (int i = threadIdx.x; i < J * cb; i += blockDim.x) {
        const int j = i / cb, cc = i - j * cb;
        wrow[i] = gs[(static_cast<int64_t>(j) * n + im) * ctot + c_off + cc];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cvec = cb / 8;
    const uint4* xb = x + static_cast<int64_t>(im) * q * cvec;
    float* ob = out + static_cast<int64_t>(im) * q * J;
    for (int base = (blockIdx.y * 8 + warp) * 4; base < q; base += gridDim.y * 8 * 4) {
        float acc[4][J];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < J; ++j) acc[p][j] = 0.f;
        for (int cv = lane; cv < cvec; cv += 32) {
            float f[4][8];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int r = base + p < q ? base + p : q - 1;
                unpack8h(xb[static_cast<int64_t>(r) * cvec + cv], f[p]);
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float4 w0 = *reinterpret_cast<const float4*>(wrow + j * cb + cv * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(wrow + j * cb + cv * 8 + 4);
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    acc[p][j] += f[p][0] * w0.x + f[p][1] * w0.y + f[p][2] * w0.z + f[p][3] * w0.w + f[p][4] * w1.x +
                                 f[p][5] * w1.y + f[p][6] * w1.z + f[p][7] * w1.w;
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < J; ++j) acc[p][j] = warp_sum(acc[p][j]);
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (lane == p && base + p < q) {
#pragma unroll
                for (int j = 0; j < J; ++j) ob[static_cast<int64_t>(base + p) * J + j] = acc[p][j];
            }
    }
}

int head_dmask(const void* x, const float* gs, int n, int q, int cb, int c_off, int ctot, int kp1, float* out,
               cudaStream_t stream) {
    BPB_CHECK(cb % 8 == 0, "head_dmask: channels %d", cb);
    const int J = kp1 + 2;
    const size_t smem = static_cast<size_t>(J) * cb * sizeof(float);
    BPB_CHECK(smem <= 100 * 1024, "head_dmask: %d rows x %d channels exceed shared memory", J, cb);
    int chunks = ceil_div(q, 32);
    if (chunks > 16) chunks = 16;
    dim3 grid(n, chunks);
#define BPB_DM_CASE(JJ)                                                                                    \
    case JJ: {                                                                                             \
        auto kern = head_dmask_kernel<JJ>;                                                                 \
        BPB_SET_MAX_SMEM(kern, 100 * 1024);                                                                \
        kern<<<grid, 256, smem, stream>>>(reinterpret_cast<const uint4*>(x), gs, n, q, cb, c_off, ctot, out); \
        break;                                                                                             \
    }
    switch (J) {
        BPB_DM_CASE(4) BPB_DM_CASE(5) BPB_DM_CASE(6) BPB_DM_CASE(7) BPB_DM_CASE(8) BPB_DM_CASE(9) BPB_DM_CASE(10)
        BPB_DM_CASE(11) BPB_DM_CASE(12) BPB_DM_CASE(13) BPB_DM_CASE(14)
        default: BPB_CHECK(false, "head_dmask: K+1=%d out of range", kp1);
    }
#undef BPB_DM_CASE
    BPB_LAUNCH_CHECK("head_dmask_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// (3) one CTA per image: d(mask) at full resolution -> d(prob) -> softmax backward (+ body-part attention loss,
//     torchreid/losses/body_part_attention_loss.py:31-52 with the target preparation of part_based_engine.py:118-124)
//     -> dlogit; low-resolution logit gradients dLow_b = U_b^T dlogit; dbf += sum dlogit.
// ----------------------------------------------------------------------------------------------
struct PixelBwdGeom {
    int hb[4], wb[4];
    const float* dmt[4];       // [N][q_b][J]  (from head_dmask)
    float* dlow[4];            // [N*q_b][KP1]
    int nb;
};

template <int KP1>
__global__ void __launch_bounds__(256) pixel_head_bwd_kernel(PixelBwdGeom g, const float* __restrict__ probs,
                                                             const float* __restrict__ dms, int h, int w,
                                                             const float* __restrict__ target_masks, int hm, int wm,
                                                             float label_smoothing, float pix_coef,
                                                             const float* __restrict__ ext_dlogits,
                                                             float* __restrict__ dbf) {
    constexpr int J = KP1 + 2;
    extern __shared__ float acc_dl[];                   // concatenated [q_b][KP1] for b = 1..nb-1
    const int n = blockIdx.x;
    const int hw = h * w;
    int off[4], total = 0;
    off[0] = 0;
    for (int b = 1; b < g.nb; ++b) { off[b] = total; total += KP1 * g.hb[b] * g.wb[b]; }
    for (int i = threadIdx.x; i < total; i += blockDim.x) acc_dl[i] = 0.f;
    __shared__ float s_db[8][KP1];
    __syncthreads();
    float dbl[KP1];
#pragma unroll
    for (int k = 0; k < KP1; ++k) dbl[k] = 0.f;
    float dmsv[J];
#pragma unroll
    for (int j = 0; j < J; ++j) dmsv[j] = dms[n * J + j];
    const float ry = h > 1 ? static_cast<float>(hm - 1) / static_cast<float>(h - 1) : 0.f;
    const float rx = w > 1 ? static_cast<float>(wm - 1) / static_cast<float>(w - 1) : 0.f;
    float* dl0 = g.dlow[0] + static_cast<int64_t>(n) * hw * KP1;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
        const int oy = p / w, ox = p % w;
        float dm[J];
        const float* d0 = g.dmt[0] + (static_cast<int64_t>(n) * hw + p) * J;
#pragma unroll
        for (int j = 0; j < J; ++j) dm[j] = d0[j] + dmsv[j];
        TapB ty[4], tx[4];
        for (int b = 1; b < g.nb; ++b) {
            ty[b] = make_tap_b(oy, g.hb[b], h);
            tx[b] = make_tap_b(ox, g.wb[b], w);
            const float* db = g.dmt[b] + static_cast<int64_t>(n) * g.hb[b] * g.wb[b] * J;
            const float w00 = (1.f - ty[b].l) * (1.f - tx[b].l), w01 = (1.f - ty[b].l) * tx[b].l;
            const float w10 = ty[b].l * (1.f - tx[b].l), w11 = ty[b].l * tx[b].l;
            const float* a = db + (ty[b].i0 * g.wb[b] + tx[b].i0) * J;
            const float* bb = db + (ty[b].i0 * g.wb[b] + tx[b].i1) * J;
            const float* c = db + (ty[b].i1 * g.wb[b] + tx[b].i0) * J;
            const float* d = db + (ty[b].i1 * g.wb[b] + tx[b].i1) * J;
#pragma unroll
            for (int j = 0; j < J; ++j) dm[j] += w00 * a[j] + w01 * bb[j] + w10 * c[j] + w11 * d[j];
        }
        float pr[KP1];
        int kfg = 1;
        float fg = -1.f;
#pragma unroll
        for (int k = 0; k < KP1; ++k) {
            pr[k] = probs[(static_cast<int64_t>(n) * KP1 + k) * hw + p];
            if (k >= 1 && pr[k] > fg) { fg = pr[k]; kfg = k; }      // first maximum, like torch.max
        }
        float dp[KP1], inner = 0.f;
        dp[0] = dm[2];
#pragma unroll
        for (int k = 1; k < KP1; ++k) dp[k] = dm[2 + k] + (k == kfg ? dm[1] : 0.f);
#pragma unroll
        for (int k = 0; k < KP1; ++k) inner += pr[k] * dp[k];
        float dl[KP1];
#pragma unroll
        for (int k = 0; k < KP1; ++k) dl[k] = pr[k] * (dp[k] - inner);
        if (target_masks != nullptr) {
            // pixel-wise CE with label smoothing against argmax_k bilinear(target masks)
            const float sy = ry * oy, sx = rx * ox;
            const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
            const int y1 = y0 + (y0 < hm - 1 ? 1 : 0), x1 = x0 + (x0 < wm - 1 ? 1 : 0);
            const float ly = sy - y0, lx = sx - x0;
            int tgt = 0;
            float best = -INFINITY;
            for (int k = 0; k < KP1; ++k) {
                const float* mp = target_masks + (static_cast<int64_t>(n) * KP1 + k) * hm * wm;
                const float v = (1.f - ly) * ((1.f - lx) * mp[y0 * wm + x0] + lx * mp[y0 * wm + x1]) +
                                ly * ((1.f - lx) * mp[y1 * wm + x0] + lx * mp[y1 * wm + x1]);
                if (v > best) { best = v; tgt = k; }
            }
#pragma unroll
            for (int k = 0; k < KP1; ++k)
                dl[k] += pix_coef * (pr[k] - (k == tgt ? 1.f - label_smoothing : 0.f) - label_smoothing / KP1);
        }
        if (ext_dlogits != nullptr) {
#pragma unroll
            for (int k = 0; k < KP1; ++k) dl[k] += ext_dlogits[(static_cast<int64_t>(n) * KP1 + k) * hw + p];
        }
#pragma unroll
        for (int k = 0; k < KP1; ++k) {
            dl0[static_cast<int64_t>(p) * KP1 + k] = dl[k];
            dbl[k] += dl[k];
        }
        for (int b = 1; b < g.nb; ++b) {
            float* a = acc_dl + off[b];
            const float w00 = (1.f - ty[b].l) * (1.f - tx[b].l), w01 = (1.f - ty[b].l) * tx[b].l;
            const float w10 = ty[b].l * (1.f - tx[b].l), w11 = ty[b].l * tx[b].l;
            const int q00 = ty[b].i0 * g.wb[b] + tx[b].i0, q01 = ty[b].i0 * g.wb[b] + tx[b].i1;
            const int q10 = ty[b].i1 * g.wb[b] + tx[b].i0, q11 = ty[b].i1 * g.wb[b] + tx[b].i1;
#pragma unroll
            for (int k = 0; k < KP1; ++k) {
                atomicAdd(a + q00 * KP1 + k, w00 * dl[k]);
                if (w01 != 0.f) atomicAdd(a + q01 * KP1 + k, w01 * dl[k]);
                if (w10 != 0.f) atomicAdd(a + q10 * KP1 + k, w10 * dl[k]);
                if (w11 != 0.f) atomicAdd(a + q11 * KP1 + k, w11 * dl[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KP1; ++k) {
        const float v = warp_sum(dbl[k]);
        if ((threadIdx.x & 31) == 0) s_db[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < KP1) {
        float t = 0.f;
        for (int wv = 0; wv < 8; ++wv) t += s_db[wv][threadIdx.x];
        atomicAdd(dbf + threadIdx.x, t);
    }
    for (int b = 1; b < g.nb; ++b) {
        const int cnt = KP1 * g.hb[b] * g.wb[b];
        float* dst = g.dlow[b] + static_cast<int64_t>(n) * cnt;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = acc_dl[off[b] + i];
    }
}

int head_pixel_bwd(const float* const* dmask_t, float* const* dlow, const int* hb, const int* wb, int nb,
                   const float* probs, const float* dms, int n, int h, int w, int kp1, const float* target_masks,
                   int hm, int wm, float label_smoothing, float pix_coef, const float* ext_dlogits, float* dbf,
                   cudaStream_t stream) {
    BPB_CHECK(nb >= 1 && nb <= 4, "head_pixel_bwd: nb=%d", nb);
    BPB_CHECK(hb[0] == h && wb[0] == w, "head_pixel_bwd: branch 0 must be at the output resolution");
    PixelBwdGeom g{};
    g.nb = nb;
    size_t smem = 0;
    for (int b = 0; b < nb; ++b) {
        g.hb[b] = hb[b]; g.wb[b] = wb[b]; g.dmt[b] = dmask_t[b]; g.dlow[b] = dlow[b];
        if (b > 0) smem += static_cast<size_t>(kp1) * hb[b] * wb[b] * sizeof(float);
    }
    BPB_CHECK(smem <= 160 * 1024, "head_pixel_bwd: low-resolution logit gradients do not fit shared memory");
#define BPB_PB_CASE(K)                                                                                          \
    case K: {                                                                                                   \
        auto kern = pixel_head_bwd_kernel<K>;                                                                   \
        BPB_SET_MAX_SMEM(kern, 160 * 1024);                                                                     \
        kern<<<n, 256, smem, stream>>>(g, probs, dms, h, w, target_masks, hm, wm, label_smoothing, pix_coef,     \
                                       ext_dlogits, dbf);                                                       \
        break;                                                                                                  \
    }
    switch (kp1) {
        BPB_PB_CASE(2) BPB_PB_CASE(3) BPB_PB_CASE(4) BPB_PB_CASE(5) BPB_PB_CASE(6) BPB_PB_CASE(7) BPB_PB_CASE(8)
        BPB_PB_CASE(9) BPB_PB_CASE(10) BPB_PB_CASE(11) BPB_PB_CASE(12)
        default: BPB_CHECK(false, "head_pixel_bwd: K+1=%d out of range", kp1);
    }
#undef BPB_PB_CASE
    BPB_LAUNCH_CHECK("pixel_head_bwd_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// (4) dwf[k][c_off + c] += sum_{n,q} dlow[n,q,k] * x[n,q,c]     CTA = (image, 256-channel slab), atomics across images
// ----------------------------------------------------------------------------------------------
template <int KP1>
__global__ void __launch_bounds__(128) head_dwf_kernel(const uint4* __restrict__ x, const float* __restrict__ dlow,
                                                       int q, int cb, int c_off, int ctot, float* __restrict__ dwf) {
    extern __shared__ float wts[];                        // [q][KP1]
    const int im = blockIdx.x, slab = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cvec = cb / 8;
    const float* src = dlow + static_cast<int64_t>(im) * q * KP1;
    for (int i = threadIdx.x; i < q * KP1; i += blockDim.x) wts[i] = src[i];
    __syncthreads();
    float acc[KP1][8];
#pragma unroll
    for (int k = 0; k < KP1; ++k)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[k][t] = 0.f;
    const int cvl = slab * 32 + lane;
    if (cvl < cvec) {
        const uint4* base = x + static_cast<int64_t>(im) * q * cvec + cvl;
        for (int p = warp; p < q; p += 4) {
            float f[8];
            unpack8h(base[static_cast<int64_t>(p) * cvec], f);
#pragma unroll
            for (int k = 0; k < KP1; ++k) {
                const float wk = wts[p * KP1 + k];
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[k][t] += wk * f[t];
            }
        }
    }
    __syncthreads();
    float* red = wts;                                     // [4][32][KP1*8]
    constexpr int kPer = KP1 * 8;
    float* mine = red + (warp * 32 + lane) * kPer;
#pragma unroll
    for (int k = 0; k < KP1; ++k)
#pragma unroll
        for (int t = 0; t < 8; ++t) mine[k * 8 + t] = acc[k][t];
    __syncthreads();
    if (warp == 0 && cvl < cvec) {
#pragma unroll
        for (int k = 0; k < KP1; ++k)
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float s = 0.f;
                for (int wv = 0; wv < 4; ++wv) s += red[(wv * 32 + lane) * kPer + k * 8 + t];
                atomicAdd(dwf + static_cast<int64_t>(k) * ctot + c_off + cvl * 8 + t, s);
            }
    }
}

int head_dwf(const void* x, const float* dlow, int n, int q, int cb, int c_off, int ctot, int kp1, float* dwf,
             cudaStream_t stream) {
    BPB_CHECK(cb % 8 == 0, "head_dwf: channels %d", cb);
    size_t smem = static_cast<size_t>(q) * kp1 * sizeof(float);
    const size_t red = static_cast<size_t>(128) * kp1 * 8 * sizeof(float);
    if (red > smem) smem = red;
    BPB_CHECK(smem <= 200 * 1024, "head_dwf: %d px x %d rows does not fit shared memory", q, kp1);
    dim3 grid(n, ceil_div(cb / 8, 32));
#define BPB_DW_CASE(K)                                                                                          \
    case K: {                                                                                                   \
        auto kern = head_dwf_kernel<K>;                                                                         \
        BPB_SET_MAX_SMEM(kern, 200 * 1024);                                                                     \
        kern<<<grid, 128, smem, stream>>>(reinterpret_cast<const uint4*>(x), dlow, q, cb, c_off, ctot, dwf);     \
        break;                                                                                                  \
    }
    switch (kp1) {
        BPB_DW_CASE(2) BPB_DW_CASE(3) BPB_DW_CASE(4) BPB_DW_CASE(5) BPB_DW_CASE(6) BPB_DW_CASE(7) BPB_DW_CASE(8)
        BPB_DW_CASE(9) BPB_DW_CASE(10) BPB_DW_CASE(11) BPB_DW_CASE(12)
        default: BPB_CHECK(false, "head_dwf: K+1=%d out of range", kp1);
    }
#undef BPB_DW_CASE
    BPB_LAUNCH_CHECK("head_dwf_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// (5) backward of the BN2d fold  wf = w*scale, bf = b + w.shift,  scale = gamma*invstd, shift = beta - mean*scale:
//     dw, db, dgamma, dbeta (parameter gradients, written) and the per-channel coefficients of the batch-statistics
//     path  dL/du = coefa + coefb * u  (u = upsampled activation), coefa = (dmean - 2 mean dvar)/M, coefb = 2 dvar/M.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixel_fold_bwd_kernel(const float* __restrict__ dwf, const float* __restrict__ dbf,
                                                             const float* __restrict__ w, const float* __restrict__ gamma,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift,
                                                             const float* __restrict__ mean_invstd, int kp1, int c,
                                                             float inv_count, float* __restrict__ dw,
                                                             float* __restrict__ db, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, float* __restrict__ coefab) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch == 0 && db != nullptr)
        for (int k = 0; k < kp1; ++k) db[k] = dbf[k];
    if (ch >= c) return;
    const float sc = scale[ch], sh = shift[ch];
    const float mean = mean_invstd[ch], istd = mean_invstd[c + ch];
    float dscale = 0.f, dshift = 0.f;
    for (int k = 0; k < kp1; ++k) {
        const float dwfk = dwf[static_cast<int64_t>(k) * c + ch], wk = w[static_cast<int64_t>(k) * c + ch];
        if (dw) dw[static_cast<int64_t>(k) * c + ch] = dwfk * sc + dbf[k] * sh;
        dscale += dwfk * wk;
        dshift += dbf[k] * wk;
    }
    const float g = gamma ? gamma[ch] : 1.f;
    if (dbeta) dbeta[ch] = dshift;
    if (dgamma) dgamma[ch] = istd * (dscale - mean * dshift);
    const float dmean = -dshift * sc;
    const float dinvstd = g * (dscale - mean * dshift);
    const float dvar = -0.5f * dinvstd * istd * istd * istd;
    coefab[ch] = (dmean - 2.f * mean * dvar) * inv_count;
    coefab[c + ch] = 2.f * dvar * inv_count;
}

int pixel_fold_bwd(const float* dwf, const float* dbf, const float* w, const float* gamma, const float* scale,
                   const float* shift, const float* mean_invstd, int kp1, int c, int64_t count, float* dw, float* db,
                   float* dgamma, float* dbeta, float* coefab, cudaStream_t stream) {
    pixel_fold_bwd_kernel<<<ceil_div(c, 256), 256, 0, stream>>>(dwf, dbf, w, gamma, scale, shift, mean_invstd, kp1, c,
                                                                1.f / static_cast<float>(count), dw, db, dgamma,
                                                                dbeta, coefab);
    BPB_LAUNCH_CHECK("pixel_fold_bwd_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// (6) dx[n,q,c] = sum_j mt[n,j,q] gs[j,n,c]  +  sum_k dlow[n,q,k] wf[k,c]  +  coefa[c] cs[q]  +  coefb[c] (T x)[n,q,c]
//     cs = column sums of U_b, T = U_b^T U_b (separable tridiagonal stencil, see branch_stats_kernel); for the
//     full-resolution branch cs = 1 and T = I.  CTA = (image, 256-channel slab); lane owns 8 channels; bf16 output.
// ----------------------------------------------------------------------------------------------
template <int KP1>
__global__ void __launch_bounds__(128) head_branch_bwd_kernel(const uint4* __restrict__ x, const float* __restrict__ mt,
                                                              const float* __restrict__ gs,
                                                              const float* __restrict__ dlow,
                                                              const float* __restrict__ wf,
                                                              const float* __restrict__ coefab, int n, int hb, int wb,
                                                              int cb, int c_off, int ctot, int h, int w,
                                                              uint4* __restrict__ dx) {
    constexpr int J = KP1 + 2;
    constexpr int R = J + KP1;
    extern __shared__ float wts[];                        // [q][R] : mask rows then logit-gradient rows
    __shared__ float coef[6][64];                         // y: colsum, diag, offdiag ; x: the same
    const int im = blockIdx.x, slab = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = hb * wb, cvec = cb / 8;
    const float* msrc = mt + static_cast<int64_t>(im) * J * q;
    const float* lsrc = dlow + static_cast<int64_t>(im) * q * KP1;
    for (int i = threadIdx.x; i < q * J; i += blockDim.x) {
        const int j = i / q, qq = i - j * q;
        wts[qq * R + j] = msrc[i];
    }
    for (int i = threadIdx.x; i < q * KP1; i += blockDim.x) {
        const int qq = i / KP1, k = i - qq * KP1;
        wts[qq * R + J + k] = lsrc[i];
    }
    const bool direct = hb == h && wb == w;
    if (!direct && threadIdx.x < 2) {
        const int in = threadIdx.x == 0 ? hb : wb, out = threadIdx.x == 0 ? h : w;
        float* cs = coef[threadIdx.x * 3 + 0];
        float* dg = coef[threadIdx.x * 3 + 1];
        float* od = coef[threadIdx.x * 3 + 2];
        for (int i = 0; i < in; ++i) { cs[i] = 0.f; dg[i] = 0.f; od[i] = 0.f; }
        for (int o = 0; o < out; ++o) {
            const TapB tp = make_tap_b(o, in, out);
            const float a = 1.f - tp.l, b = tp.l;
            if (tp.i1 == tp.i0) {
                cs[tp.i0] += a + b; dg[tp.i0] += (a + b) * (a + b);
            } else {
                cs[tp.i0] += a; cs[tp.i1] += b;
                dg[tp.i0] += a * a; dg[tp.i1] += b * b; od[tp.i0] += a * b;
            }
        }
    }
    __syncthreads();
    const int cvl = slab * 32 + lane;
    if (cvl >= cvec) return;
    // per-thread channel vectors: J pooled-gradient rows of this image, KP1 folded classifier rows, two coefficients
    float vg[J][8], vw[KP1][8], ca[8], cbv[8];
    const int ch0 = c_off + cvl * 8;
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int t = 0; t < 8; ++t) vg[j][t] = gs[(static_cast<int64_t>(j) * n + im) * ctot + ch0 + t];
#pragma unroll
    for (int k = 0; k < KP1; ++k)
#pragma unroll
        for (int t = 0; t < 8; ++t) vw[k][t] = wf[static_cast<int64_t>(k) * ctot + ch0 + t];
#pragma unroll
    for (int t = 0; t < 8; ++t) { ca[t] = coefab[ch0 + t]; cbv[t] = coefab[ctot + ch0 + t]; }
    const uint4* xb = x + static_cast<int64_t>(im) * q * cvec + cvl;
    uint4* ob = dx + static_cast<int64_t>(im) * q * cvec + cvl;
    for (int p = warp; p < q; p += 4) {
        float o[8];
        const float* wr = wts + p * R;
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float cj = wr[j];
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] += cj * vg[j][t];
        }
#pragma unroll
        for (int k = 0; k < KP1; ++k) {
            const float ck = wr[J + k];
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] += ck * vw[k][t];
        }
        if (direct) {
            float f[8];
            unpack8h(xb[static_cast<int64_t>(p) * cvec], f);
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] += ca[t] + cbv[t] * f[t];
        } else {
            const int iy = p / wb, ix = p - iy * wb;
            const float* csy = coef[0]; const float* dgy = coef[1]; const float* ody = coef[2];
            const float* csx = coef[3]; const float* dgx = coef[4]; const float* odx = coef[5];
            const float wy[3] = {iy > 0 ? ody[iy - 1] : 0.f, dgy[iy], iy + 1 < hb ? ody[iy] : 0.f};
            const float wx[3] = {ix > 0 ? odx[ix - 1] : 0.f, dgx[ix], ix + 1 < wb ? odx[ix] : 0.f};
            float z[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) z[t] = 0.f;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = iy + dy;
                if (yy < 0 || yy >= hb) continue;
#pragma unroll
                for (int dxx = -1; dxx <= 1; ++dxx) {
                    const int xx = ix + dxx;
                    if (xx < 0 || xx >= wb) continue;
                    float f[8];
                    unpack8h(xb[static_cast<int64_t>(yy * wb + xx) * cvec], f);
                    const float wgt = wy[dy + 1] * wx[dxx + 1];
#pragma unroll
                    for (int t = 0; t < 8; ++t) z[t] += wgt * f[t];
                }
            }
            const float cw = csy[iy] * csx[ix];
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] += ca[t] * cw + cbv[t] * z[t];
        }
        uint4 pk;
        pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]); pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
        ob[static_cast<int64_t>(p) * cvec] = pk;
    }
}

int head_branch_bwd(const void* x, const float* mt, const float* gs, const float* dlow, const float* wf,
                    const float* coefab, int n, int hb, int wb, int cb, int c_off, int ctot, int h, int w, int kp1,
                    void* dx, cudaStream_t stream) {
    BPB_CHECK(cb % 8 == 0, "head_branch_bwd: channels %d", cb);
    BPB_CHECK((hb == h && wb == w) || (hb <= 64 && wb <= 64), "head_branch_bwd: low-resolution branch %dx%d larger than 64",
              hb, wb);
    const int R = 2 * kp1 + 2;
    const size_t smem = static_cast<size_t>(hb) * wb * R * sizeof(float);
    BPB_CHECK(smem <= 200 * 1024, "head_branch_bwd: %d px x %d rows does not fit shared memory", hb * wb, R);
    dim3 grid(n, ceil_div(cb / 8, 32));
#define BPB_HB_CASE(K)                                                                                           \
    case K: {                                                                                                    \
        auto kern = head_branch_bwd_kernel<K>;                                                                   \
        BPB_SET_MAX_SMEM(kern, 200 * 1024);                                                                      \
        kern<<<grid, 128, smem, stream>>>(reinterpret_cast<const uint4*>(x), mt, gs, dlow, wf, coefab, n, hb, wb, \
                                          cb, c_off, ctot, h, w, reinterpret_cast<uint4*>(dx));                  \
        break;                                                                                                   \
    }
    switch (kp1) {
        BPB_HB_CASE(2) BPB_HB_CASE(3) BPB_HB_CASE(4) BPB_HB_CASE(5) BPB_HB_CASE(6) BPB_HB_CASE(7) BPB_HB_CASE(8)
        BPB_HB_CASE(9) BPB_HB_CASE(10) BPB_HB_CASE(11) BPB_HB_CASE(12)
        default: BPB_CHECK(false, "head_branch_bwd: K+1=%d out of range", kp1);
    }
#undef BPB_HB_CASE
    BPB_LAUNCH_CHECK("head_branch_bwd_kernel");
    return 0;
}

}  // namespace bpb
