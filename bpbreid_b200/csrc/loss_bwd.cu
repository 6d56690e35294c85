// Backward of the GiLt loss terms and of the small fp32 head layers (training step; autograd replaced at
// torchreid/engine/image/part_based_engine.py:98-103).
//   * label-smoothed identity cross-entropy      torchreid/losses/cross_entropy_loss.py:34-56
//   * part-averaged batch-hard triplet            torchreid/losses/part_averaged_triplet_loss.py:35-224
//   * BatchNorm1d (train) of the dim-reduce / BN-classifier layers   torchreid/models/bpbreid.py:324-350, 398-415
// Everything fp32; gradients are written (or accumulated where stated) into caller-owned buffers.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// ----------------------------------------------------------------------------------------------
// dlogits[i, c] = coef_i * (softmax(x_i)_c - t_ic),  t = (1-eps) onehot + eps/C
// coef_i = loss_weight / n (unweighted)  or  loss_weight * w_i / max(sum |w|, 1e-12)  (visibility-weighted rows)
// One warp per row; columns [c, ldg) of dlogits are left untouched (the caller keeps that padding zero).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, int64_t ld,
                                                     const int64_t* __restrict__ targets,
                                                     const float* __restrict__ weights, int n, int c, float eps,
                                                     float loss_weight, float* __restrict__ dlogits, int64_t ldg) {
    __shared__ float s_wsum;
    if (weights != nullptr) {
        __shared__ float red[256];
        float w = 0.f;
        for (int i = threadIdx.x; i < n; i += 256) w += fabsf(weights[i]);
        red[threadIdx.x] = w;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) s_wsum = fmaxf(red[0], 1e-12f);
        __syncthreads();
    }
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const float coef = weights ? loss_weight * weights[i] / s_wsum : loss_weight / static_cast<float>(n);
    const float* x = logits + static_cast<int64_t>(i) * ld;
    float mx = -INFINITY;
    for (int j = lane; j < c; j += 32) mx = fmaxf(mx, x[j]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int j = lane; j < c; j += 32) se += expf(x[j] - mx);
    se = warp_sum(se);
    const float inv = 1.f / se;
    const int64_t y = targets[i];
    const float tbase = eps / static_cast<float>(c);
    float* g = dlogits + static_cast<int64_t>(i) * ldg;
    for (int j = lane; j < c; j += 32) {
        const float t = tbase + (j == y ? 1.f - eps : 0.f);
        g[j] = coef * (expf(x[j] - mx) * inv - t);
    }
}

int ce_ls_bwd(const float* logits, int64_t ld, const int64_t* targets, const float* weights, int n, int c, float eps,
              float loss_weight, float* dlogits, int64_t ldg, cudaStream_t stream) {
    ce_bwd_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(logits, ld, targets, weights, n, c, eps, loss_weight, dlogits, ldg);
    BPB_LAUNCH_CHECK("ce_bwd_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Triplet backward.  loss = (1/n_valid) sum_valid relu(hp_i - hn_i + m)  (or softplus(hp_i - hn_i) for margin <= 0),
// hp_i = dist[i, hp_idx[i]], hn_i = dist[i, hn_idx[i]], dist = (visibility-)weighted mean over parts of
// d_k(i,j) = |e_ki - e_kj| (zero where the squared distance is exactly zero, part_averaged_triplet_loss.py:88-91).
// One CTA per anchor, rows [row_lo, row_lo + rows) of the (possibly all-gathered) batch; demb [K, N, D] += (atomics:
// a sample is the hardest positive / negative of several anchors).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) triplet_bwd_kernel(const float* __restrict__ emb, const void* __restrict__ vis,
                                                          int vis_mode, const float* __restrict__ hp,
                                                          const float* __restrict__ hn, const int* __restrict__ hp_idx,
                                                          const int* __restrict__ hn_idx,
                                                          const float* __restrict__ out4, int n, int kparts, int d,
                                                          float margin, float loss_weight, int row_lo, int n_local,
                                                          float* __restrict__ demb) {
    const int il = blockIdx.x;                             // local anchor row; hp / hn / indices are per local row
    const int i = row_lo + il;
    const float hpv = hp[il], hnv = hn[il];
    if (hpv == -1.f || hnv == FLT_MAX) return;
    const float nvalid = out4[3];
    if (!(nvalid > 0.f)) return;
    float a;                                               // dL/dhp = -dL/dhn
    if (margin <= 0.f) {
        a = 1.f / (1.f + expf(hnv - hpv));                 // d/dhp log(1 + exp(hp - hn))
    } else {
        a = (hpv - hnv + margin) > 0.f ? 1.f : 0.f;
    }
    a *= loss_weight / nvalid;
    if (a == 0.f) return;
    __shared__ float red[4];
    __shared__ float s_d;
    for (int pair = 0; pair < 2; ++pair) {
        const int j = pair == 0 ? hp_idx[il] : hn_idx[il];
        const float sign = pair == 0 ? a : -a;
        if (j < 0) continue;
        // weights of the masked mean (tensortools.py:10-19) for this pair
        float den = 0.f;
        if (vis_mode == 0) {
            den = static_cast<float>(kparts);
        } else {
            for (int k = 0; k < kparts; ++k) {
                float w;
                if (vis_mode == 1) {
                    const uint8_t* v = reinterpret_cast<const uint8_t*>(vis);
                    w = (v[static_cast<int64_t>(i) * kparts + k] && v[static_cast<int64_t>(j) * kparts + k]) ? 1.f : 0.f;
                } else {
                    const float* v = reinterpret_cast<const float*>(vis);
                    w = sqrtf(v[static_cast<int64_t>(i) * kparts + k] * v[static_cast<int64_t>(j) * kparts + k]);
                }
                den += w;
            }
        }
        if (den == 0.f) continue;
        for (int k = 0; k < kparts; ++k) {
            float w = 1.f;
            if (vis_mode == 1) {
                const uint8_t* v = reinterpret_cast<const uint8_t*>(vis);
                w = (v[static_cast<int64_t>(i) * kparts + k] && v[static_cast<int64_t>(j) * kparts + k]) ? 1.f : 0.f;
            } else if (vis_mode == 2) {
                const float* v = reinterpret_cast<const float*>(vis);
                w = sqrtf(v[static_cast<int64_t>(i) * kparts + k] * v[static_cast<int64_t>(j) * kparts + k]);
            }
            if (w == 0.f) continue;                        // block-uniform
            // [shards][K][n_local][D] layout (see pair_dist_kernel); one shard == the plain [K, N, D] tensor
            const int64_t oi = ((static_cast<int64_t>(i / n_local) * kparts + k) * n_local + i % n_local) * d;
            const int64_t oj = ((static_cast<int64_t>(j / n_local) * kparts + k) * n_local + j % n_local) * d;
            const float* ei = emb + oi;
            const float* ej = emb + oj;
            float q = 0.f;
            for (int t = threadIdx.x; t < d; t += 128) {
                const float df = ei[t] - ej[t];
                q = fmaf(df, df, q);
            }
            q = warp_sum(q);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
            __syncthreads();
            if (threadIdx.x == 0) s_d = sqrtf(red[0] + red[1] + red[2] + red[3]);
            __syncthreads();
            const float dk = s_d;
            if (dk == 0.f) continue;                       // exact-zero mask: no gradient
            const float cf = sign * w / (den * dk);
            float* gi = demb + oi;
            float* gj = demb + oj;
            for (int t = threadIdx.x; t < d; t += 128) {
                const float g = cf * (ei[t] - ej[t]);
                atomicAdd(gi + t, g);
                atomicAdd(gj + t, -g);
            }
        }
    }
}

int triplet_bwd_rows(const float* emb, const void* vis, int vis_mode, const float* hp, const float* hn,
                     const int* hp_idx, const int* hn_idx, const float* out4, int n, int k, int d, float margin,
                     float loss_weight, int row_lo, int rows, int n_local, float* demb, cudaStream_t stream) {
    BPB_CHECK(vis_mode >= 0 && vis_mode <= 2, "triplet_bwd: vis_mode %d", vis_mode);
    BPB_CHECK(n_local >= 1 && n % n_local == 0 && row_lo >= 0 && rows >= 1 && row_lo + rows <= n,
              "triplet_bwd: bad row range / shard size");
    triplet_bwd_kernel<<<rows, 128, 0, stream>>>(emb, vis, vis_mode, hp, hn, hp_idx, hn_idx, out4, n, k, d, margin,
                                                 loss_weight, row_lo, n_local, demb);
    BPB_LAUNCH_CHECK("triplet_bwd_kernel");
    return 0;
}

int triplet_bwd(const float* emb_knd, const void* vis, int vis_mode, const float* hp, const float* hn,
                const int* hp_idx, const int* hn_idx, const float* out4, int n, int k, int d, float margin,
                float loss_weight, float* demb_knd, cudaStream_t stream) {
    return triplet_bwd_rows(emb_knd, vis, vis_mode, hp, hn, hp_idx, hn_idx, out4, n, k, d, margin, loss_weight, 0, n, n,
                            demb_knd, stream);
}

// ----------------------------------------------------------------------------------------------
// BatchNorm1d (train) backward on fp32 [m, c] with optional ReLU gate (mask = y > 0, y = the layer's output).
//   dx (+)= gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)) ; dgamma = sum g*xhat ; dbeta = sum g
// grid = ceil(c/32), block = 32 columns x 8 row lanes; the batch statistics are recomputed from x (tiny matrices).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn1d_bwd_kernel(const float* __restrict__ dy, int64_t ld_dy,
                                                       const float* __restrict__ x, int64_t ld_x,
                                                       const float* __restrict__ y, int64_t ld_y,
                                                       const float* __restrict__ gamma, float eps, int m, int c,
                                                       float* __restrict__ dx, int64_t ld_dx, int accumulate,
                                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int cl = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + cl;
    __shared__ float red[2][8][32];
    __shared__ float s_a[32], s_b[32];
    const bool ok = ch < c;
    // batch mean / variance of x
    float s = 0.f;
    if (ok) for (int r = j; r < m; r += 8) s += x[static_cast<int64_t>(r) * ld_x + ch];
    red[0][j][cl] = s;
    __syncthreads();
    if (j == 0) { for (int l = 1; l < 8; ++l) s += red[0][l][cl]; s_a[cl] = s / m; }
    __syncthreads();
    const float mean = s_a[cl];
    float q = 0.f;
    if (ok) for (int r = j; r < m; r += 8) { const float dd = x[static_cast<int64_t>(r) * ld_x + ch] - mean; q += dd * dd; }
    __syncthreads();
    red[0][j][cl] = q;
    __syncthreads();
    if (j == 0) { for (int l = 1; l < 8; ++l) q += red[0][l][cl]; s_b[cl] = rsqrtf(q / m + eps); }
    __syncthreads();
    const float rstd = s_b[cl];
    // sums of g and g * xhat
    float sg = 0.f, sgx = 0.f;
    if (ok)
        for (int r = j; r < m; r += 8) {
            float g = dy[static_cast<int64_t>(r) * ld_dy + ch];
            if (y != nullptr && !(y[static_cast<int64_t>(r) * ld_y + ch] > 0.f)) g = 0.f;
            sg += g;
            sgx += g * (x[static_cast<int64_t>(r) * ld_x + ch] - mean) * rstd;
        }
    __syncthreads();
    red[0][j][cl] = sg; red[1][j][cl] = sgx;
    __syncthreads();
    if (j == 0) {
        for (int l = 1; l < 8; ++l) { sg += red[0][l][cl]; sgx += red[1][l][cl]; }
        s_a[cl] = sg; s_b[cl] = sgx;
        if (ok) {
            if (dgamma) dgamma[ch] = sgx;
            if (dbeta) dbeta[ch] = sg;
        }
    }
    __syncthreads();
    if (!ok) return;
    const float mg = s_a[cl] / m, mgx = s_b[cl] / m;
    const float k = (gamma ? gamma[ch] : 1.f) * rstd;
    for (int r = j; r < m; r += 8) {
        float g = dy[static_cast<int64_t>(r) * ld_dy + ch];
        if (y != nullptr && !(y[static_cast<int64_t>(r) * ld_y + ch] > 0.f)) g = 0.f;
        const float xh = (x[static_cast<int64_t>(r) * ld_x + ch] - mean) * rstd;
        const float v = k * (g - mg - xh * mgx);
        float* o = dx + static_cast<int64_t>(r) * ld_dx + ch;
        *o = accumulate ? *o + v : v;
    }
}

int bn1d_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* y, int64_t ld_y,
             const float* gamma, float eps, int m, int c, float* dx, int64_t ld_dx, int accumulate, float* dgamma,
             float* dbeta, cudaStream_t stream) {
    bn1d_bwd_kernel<<<ceil_div(c, 32), 256, 0, stream>>>(dy, ld_dy, x, ld_x, y, ld_y, gamma, eps, m, c, dx, ld_dx,
                                                         accumulate, dgamma, dbeta);
    BPB_LAUNCH_CHECK("bn1d_bwd_kernel");
    return 0;
}

// out[c] = sum_r x[r, c]   (nn.Linear bias gradient)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t ld, int m, int c,
                                                     float* __restrict__ out) {
    const int cl = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + cl;
    __shared__ float red[8][32];
    float s = 0.f;
    if (ch < c) for (int r = j; r < m; r += 8) s += x[static_cast<int64_t>(r) * ld + ch];
    red[j][cl] = s;
    __syncthreads();
    if (j == 0 && ch < c) {
        for (int l = 1; l < 8; ++l) s += red[l][cl];
        out[ch] = s;
    }
}

int colsum_f32(const float* x, int64_t ld, int m, int c, float* out, cudaStream_t stream) {
    colsum_kernel<<<ceil_div(c, 32), 256, 0, stream>>>(x, ld, m, c, out);
    BPB_LAUNCH_CHECK("colsum_kernel");
    return 0;
}

}  // namespace bpb
