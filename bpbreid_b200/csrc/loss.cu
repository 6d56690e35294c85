// GiLt loss kernels: part-averaged batch-hard triplet (pairwise part distances -> (masked) mean over parts ->
// hardest positive / hardest negative mining with warp-shuffle arg-reductions -> margin loss) and the
// label-smoothed identity cross-entropy, plus the pixel-wise body-part attention cross-entropy.
// Reference semantics: torchreid/losses/part_averaged_triplet_loss.py:35-224, torchreid/utils/tensortools.py:10-19,
// torchreid/losses/cross_entropy_loss.py:34-56, torchreid/losses/body_part_attention_loss.py:31-52 with target
// preparation from torchreid/engine/image/part_based_engine.py:118-124.
// Everything is fp32 with the reference's sentinels (-1 = "pair not comparable", FLT_MAX = "no negative").
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

constexpr int kPT = 16;        // pair tile edge
constexpr int kDC = 64;        // feature chunk

// dot(a,b) accumulated d = 0..D-1 with one fp32 FMA chain: the diagonal (a == b) and duplicated embeddings give
// bit-identical values for |a|^2 and <a,b>, so d^2 is exactly 0 where the reference's is (":88-91" zero mask).
// Embedding layout: [shards][K][n_local][D] (shards * n_local == n; one shard = the plain [K, N, D] tensor).  With several
// GPUs every rank contributes one shard through a single all-gather and computes only ITS anchor rows
// [row_lo, row_lo + rows) of the pairwise matrix -- the reference gathers the whole batch on GPU 0
// (part_based_engine.py:82-95) and mines there; the mined pairs are identical.
__device__ __forceinline__ const float* emb_row(const float* emb, int k, int j, int kparts, int n_local, int d) {
    const int sh = j / n_local, jl = j - sh * n_local;
    return emb + ((static_cast<int64_t>(sh) * kparts + k) * n_local + jl) * d;
}

__global__ void __launch_bounds__(256) pair_dist_kernel(const float* __restrict__ emb, const void* __restrict__ vis,
                                                        int vis_mode, int n, int kparts, int d, int row_lo, int rows,
                                                        int n_local, float* __restrict__ dist) {
    __shared__ float sa[kPT][kDC + 1], sb[kPT][kDC + 1];
    __shared__ float sqa[kPT], sqb[kPT];
    const int ti = threadIdx.x / kPT, tj = threadIdx.x % kPT;
    const int il = blockIdx.y * kPT + ti;                      // local anchor row
    const int i = row_lo + il, j = blockIdx.x * kPT + tj;
    float num = 0.f, den = 0.f;
    for (int k = 0; k < kparts; ++k) {
        float dot = 0.f, qa = 0.f, qb = 0.f;   // qa/qb: only threads with tj==0 / ti==0 use them
        for (int d0 = 0; d0 < d; d0 += kDC) {
            for (int t = threadIdx.x; t < kPT * kDC; t += 256) {
                const int r = t / kDC, cc = t % kDC;
                const int ral = blockIdx.y * kPT + r, rb = blockIdx.x * kPT + r;
                const int ra = row_lo + ral;
                sa[r][cc] = (ral < rows && d0 + cc < d) ? emb_row(emb, k, ra, kparts, n_local, d)[d0 + cc] : 0.f;
                sb[r][cc] = (rb < n && d0 + cc < d) ? emb_row(emb, k, rb, kparts, n_local, d)[d0 + cc] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int cc = 0; cc < kDC; ++cc) dot = fmaf(sa[ti][cc], sb[tj][cc], dot);
            if (tj == 0)
#pragma unroll 8
                for (int cc = 0; cc < kDC; ++cc) qa = fmaf(sa[ti][cc], sa[ti][cc], qa);
            if (ti == 0)
#pragma unroll 8
                for (int cc = 0; cc < kDC; ++cc) qb = fmaf(sb[tj][cc], sb[tj][cc], qb);
            __syncthreads();
        }
        if (tj == 0) sqa[ti] = qa;
        if (ti == 0) sqb[tj] = qb;
        __syncthreads();
        float d2 = sqa[ti] - 2.f * dot + sqb[tj];
        d2 = fmaxf(d2, 0.f);
        const float zero = d2 == 0.f ? 1.f : 0.f;
        float dd = sqrtf(d2 + zero * 1e-16f) * (1.f - zero);
        float w = 1.f;
        if (vis_mode == 1 && il < rows && j < n) {
            const uint8_t* v = reinterpret_cast<const uint8_t*>(vis);
            w = (v[static_cast<int64_t>(i) * kparts + k] && v[static_cast<int64_t>(j) * kparts + k]) ? 1.f : 0.f;
        } else if (vis_mode == 2 && il < rows && j < n) {
            const float* v = reinterpret_cast<const float*>(vis);
            w = sqrtf(v[static_cast<int64_t>(i) * kparts + k] * v[static_cast<int64_t>(j) * kparts + k]);
        }
        num += dd * w;
        den += w;
        __syncthreads();
    }
    if (il < rows && j < n) {
        float out;
        if (vis_mode == 0) out = num / static_cast<float>(kparts);          // pd.mean(0)
        else out = den == 0.f ? -1.f : num / den;                            // masked_mean with the -1 sentinel
        dist[static_cast<int64_t>(il) * n + j] = out;
    }
}

// one warp per anchor: hardest positive (max) / hardest negative (min) with first-occurrence indices
__global__ void __launch_bounds__(256) mine_kernel(const float* __restrict__ dist, const int64_t* __restrict__ labels,
                                                   int n, int row_lo, int rows, float* __restrict__ hp,
                                                   float* __restrict__ hn, int* __restrict__ hp_idx,
                                                   int* __restrict__ hn_idx) {
    const int il = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (il >= rows) return;
    const int i = row_lo + il;
    const int64_t li = labels[i];
    float bp = -1.f, bn = FLT_MAX;
    int ip = 0x7fffffff, in_ = 0x7fffffff;
    for (int j = lane; j < n; j += 32) {
        const float dv = dist[static_cast<int64_t>(il) * n + j];
        const bool valid = dv != -1.f;
        const bool same = labels[j] == li;
        // reference: d*mask - (~mask)  and  d*mask + (~mask)*FLT_MAX
        const float pv = (same && j != i && valid) ? dv : -1.f;
        const float nv = (!same && valid) ? dv : FLT_MAX;
        if (pv > bp) { bp = pv; ip = j; }
        if (nv < bn) { bn = nv; in_ = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float obp = __shfl_xor_sync(0xffffffffu, bp, o);
        const int oip = __shfl_xor_sync(0xffffffffu, ip, o);
        if (obp > bp || (obp == bp && oip < ip)) { bp = obp; ip = oip; }
        const float obn = __shfl_xor_sync(0xffffffffu, bn, o);
        const int oin = __shfl_xor_sync(0xffffffffu, in_, o);
        if (obn < bn || (obn == bn && oin < in_)) { bn = obn; in_ = oin; }
    }
    if (lane == 0) {
        hp[il] = bp; hn[il] = bn;
        hp_idx[il] = bp == -1.f ? -1 : ip;
        hn_idx[il] = bn == FLT_MAX ? -1 : in_;
    }
}

// out4 = [loss, trivial_ratio, valid_ratio, n_valid]; sums_mode: [sum loss, sum trivial, n_valid, rows] (partial sums
// of one rank's anchor rows, all-reduced across ranks and turned into the ratios by triplet_finish_kernel)
__global__ void __launch_bounds__(256) triplet_finalize_kernel(const float* hp, const float* hn, int n, float margin,
                                                               int sums_mode, float* out4) {
    __shared__ float rl[256], rt[256], rv[256];
    float l = 0.f, t = 0.f, v = 0.f;
    const bool soft = margin <= 0.f;
    const float mg = soft ? 0.3f : margin;
    for (int i = threadIdx.x; i < n; i += 256) {
        if (hp[i] != -1.f && hn[i] != FLT_MAX) {
            const float hard = fmaxf(hp[i] - hn[i] + mg, 0.f);
            v += 1.f;
            t += hard == 0.f ? 1.f : 0.f;
            l += soft ? log1pf(expf(-(hn[i] - hp[i]))) : hard;
        }
    }
    rl[threadIdx.x] = l; rt[threadIdx.x] = t; rv[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            rl[threadIdx.x] += rl[threadIdx.x + s];
            rt[threadIdx.x] += rt[threadIdx.x + s];
            rv[threadIdx.x] += rv[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float nv = rv[0];
        if (sums_mode) {
            out4[0] = rl[0]; out4[1] = rt[0]; out4[2] = nv; out4[3] = static_cast<float>(n);
        } else {
            out4[0] = nv > 0.f ? rl[0] / nv : nanf("");
            out4[1] = nv > 0.f ? rt[0] / nv : 0.f;
            out4[2] = nv / static_cast<float>(n);
            out4[3] = nv;
        }
    }
}

__global__ void triplet_finish_kernel(const float* sums4, float* out4) {
    if (threadIdx.x == 0) {
        const float sl = sums4[0], st = sums4[1], nv = sums4[2], rows = sums4[3];
        out4[0] = nv > 0.f ? sl / nv : nanf("");
        out4[1] = nv > 0.f ? st / nv : 0.f;
        out4[2] = rows > 0.f ? nv / rows : 0.f;
        out4[3] = nv;
    }
}

int triplet_fwd_rows(const float* emb, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                     float margin, int row_lo, int rows, int n_local, int sums_mode, float* dist_rows, float* hp,
                     float* hn, int* hp_idx, int* hn_idx, float* out4, cudaStream_t stream) {
    BPB_CHECK(vis_mode >= 0 && vis_mode <= 2, "triplet_fwd: vis_mode %d", vis_mode);
    BPB_CHECK(vis_mode == 0 || vis != nullptr, "triplet_fwd: visibility pointer missing");
    BPB_CHECK(n_local >= 1 && n % n_local == 0, "triplet_fwd: n=%d is not a multiple of the shard size %d", n, n_local);
    BPB_CHECK(row_lo >= 0 && rows >= 1 && row_lo + rows <= n, "triplet_fwd: anchor rows [%d,+%d) outside [0,%d)", row_lo,
              rows, n);
    dim3 grid(ceil_div(n, kPT), ceil_div(rows, kPT));
    pair_dist_kernel<<<grid, 256, 0, stream>>>(emb, vis, vis_mode, n, k, d, row_lo, rows, n_local, dist_rows);
    BPB_LAUNCH_CHECK("pair_dist_kernel");
    mine_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(dist_rows, labels, n, row_lo, rows, hp, hn, hp_idx, hn_idx);
    BPB_LAUNCH_CHECK("mine_kernel");
    triplet_finalize_kernel<<<1, 256, 0, stream>>>(hp, hn, rows, margin, sums_mode, out4);
    BPB_LAUNCH_CHECK("triplet_finalize_kernel");
    return 0;
}

int triplet_fwd(const float* emb_knd, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                float margin, float* dist_nn, float* hp, float* hn, int* hp_idx, int* hn_idx, float* out4,
                cudaStream_t stream) {
    return triplet_fwd_rows(emb_knd, labels, vis, vis_mode, n, k, d, margin, 0, n, n, 0, dist_nn, hp, hn, hp_idx, hn_idx,
                            out4, stream);
}

int triplet_finish(const float* sums4, float* out4, cudaStream_t stream) {
    triplet_finish_kernel<<<1, 32, 0, stream>>>(sums4, out4);
    BPB_LAUNCH_CHECK("triplet_finish_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// label-smoothed CE: one warp per row; row_loss[i] = -sum_c t_c logp_c, t = (1-eps) onehot + eps/C
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, int64_t ld,
                                                      const int64_t* __restrict__ targets, int n, int c, float eps,
                                                      float* __restrict__ row_loss, float* __restrict__ row_hit) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const float* x = logits + static_cast<int64_t>(i) * ld;
    float mx = -INFINITY;
    int am = 0x7fffffff;
    for (int j = lane; j < c; j += 32) {
        const float v = x[j];
        if (v > mx) { mx = v; am = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
    float se = 0.f, sx = 0.f;
    for (int j = lane; j < c; j += 32) {
        const float z = x[j] - mx;
        se += expf(z);
        sx += z;
    }
    se = warp_sum(se);
    sx = warp_sum(sx);
    if (lane == 0) {
        const float lse = logf(se);
        const int64_t y = targets[i];
        const float logp_y = (x[y] - mx) - lse;
        const float sum_logp = sx - static_cast<float>(c) * lse;
        row_loss[i] = -((1.f - eps) * logp_y + (eps / static_cast<float>(c)) * sum_logp);
        row_hit[i] = am == y ? 1.f : 0.f;
    }
}

// out2 = [loss, accuracy]; weights == null -> mean over rows, else sum_i CE_i * w_i / max(sum|w|, 1e-12)
// acc_mode 1: accuracy over the rows with non-zero weight (the reference drops invisible rows before both the loss
// and the accuracy when visibility scores are boolean, GiLt_loss.py:112-114)
__global__ void __launch_bounds__(256) ce_finalize_kernel(const float* row_loss, const float* row_hit,
                                                          const float* weights, int n, int acc_mode, float* out2) {
    __shared__ float r0[256], r1[256], r2[256], r3[256];
    float l = 0.f, h = 0.f, w = 0.f, cnt = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float wi = weights ? weights[i] : 1.f;
        l += row_loss[i] * wi;
        w += fabsf(wi);
        const float sel = (acc_mode == 1 && wi == 0.f) ? 0.f : 1.f;
        h += row_hit[i] * sel;
        cnt += sel;
    }
    r3[threadIdx.x] = cnt;
    r0[threadIdx.x] = l; r1[threadIdx.x] = h; r2[threadIdx.x] = w;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            r0[threadIdx.x] += r0[threadIdx.x + s];
            r1[threadIdx.x] += r1[threadIdx.x + s];
            r2[threadIdx.x] += r2[threadIdx.x + s];
            r3[threadIdx.x] += r3[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out2[0] = weights ? r0[0] / fmaxf(r2[0], 1e-12f) : r0[0] / static_cast<float>(n);
        out2[1] = r1[0] / fmaxf(r3[0], 1.f);
    }
}

int ce_ls_fwd(const float* logits, int64_t ld, const int64_t* targets, const float* weights, int n, int c, float eps,
              int acc_mode, float* row_loss, float* out2, cudaStream_t stream) {
    // row_loss has room for 2*n floats: [loss | hit]
    ce_rows_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(logits, ld, targets, n, c, eps, row_loss, row_loss + n);
    BPB_LAUNCH_CHECK("ce_rows_kernel");
    ce_finalize_kernel<<<1, 256, 0, stream>>>(row_loss, row_loss + n, weights, n, acc_mode, out2);
    BPB_LAUNCH_CHECK("ce_finalize_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Body-part attention loss: target = argmax_k bilinear(align_corners)(external masks) ; CE(label_smoothing) / pixel
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixel_ce_kernel(const float* __restrict__ logits, const float* __restrict__ masks,
                                                       int n, int kp1, int hf, int wf, int hm, int wm, float ls,
                                                       float* __restrict__ partial) {
    const int64_t total = static_cast<int64_t>(n) * hf * wf;
    float loss = 0.f, hit = 0.f;
    const float ry = hf > 1 ? static_cast<float>(hm - 1) / static_cast<float>(hf - 1) : 0.f;
    const float rx = wf > 1 ? static_cast<float>(wm - 1) / static_cast<float>(wf - 1) : 0.f;
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ox = idx % wf, oy = (idx / wf) % hf, im = idx / (static_cast<int64_t>(wf) * hf);
        const float sy = ry * oy, sx = rx * ox;
        const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
        const int y1 = y0 + (y0 < hm - 1 ? 1 : 0), x1 = x0 + (x0 < wm - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
        int tgt = 0, pred = 0;
        float best_t = -INFINITY, best_l = -INFINITY, mx = -INFINITY;
        for (int k = 0; k < kp1; ++k) {
            const float* mp = masks + (static_cast<int64_t>(im) * kp1 + k) * hm * wm;
            const float v = (1.f - ly) * ((1.f - lx) * mp[y0 * wm + x0] + lx * mp[y0 * wm + x1]) +
                            ly * ((1.f - lx) * mp[y1 * wm + x0] + lx * mp[y1 * wm + x1]);
            if (v > best_t) { best_t = v; tgt = k; }
            const float l = logits[(static_cast<int64_t>(im) * kp1 + k) * hf * wf + oy * wf + ox];
            if (l > best_l) { best_l = l; pred = k; }
            mx = fmaxf(mx, l);
        }
        float se = 0.f, sz = 0.f, zt = 0.f;
        for (int k = 0; k < kp1; ++k) {
            const float z = logits[(static_cast<int64_t>(im) * kp1 + k) * hf * wf + oy * wf + ox] - mx;
            se += expf(z);
            sz += z;
            if (k == tgt) zt = z;
        }
        const float lse = logf(se);
        const float nll = -(zt - lse);
        const float smooth = -(sz - kp1 * lse) / kp1;
        loss += (1.f - ls) * nll + ls * smooth;
        hit += pred == tgt ? 1.f : 0.f;
    }
    __shared__ float r0[256], r1[256];
    r0[threadIdx.x] = loss; r1[threadIdx.x] = hit;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { r0[threadIdx.x] += r0[threadIdx.x + s]; r1[threadIdx.x] += r1[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = r0[0]; partial[2 * blockIdx.x + 1] = r1[0]; }
}

__global__ void pixel_ce_finalize_kernel(const float* partial, int blocks, float inv_total, float* out2) {
    if (threadIdx.x == 0) {
        double l = 0.0, h = 0.0;
        for (int b = 0; b < blocks; ++b) { l += partial[2 * b]; h += partial[2 * b + 1]; }
        out2[0] = static_cast<float>(l * inv_total);
        out2[1] = static_cast<float>(h * inv_total);
    }
}

int pixel_ce_fwd(const float* logits_nchw, const float* target_masks, int n, int kp1, int hf, int wf, int hm, int wm,
                 float label_smoothing, float* partial, float* out2, cudaStream_t stream) {
    const int64_t total = static_cast<int64_t>(n) * hf * wf;
    int blocks = ceil_div(total, 256);
    if (blocks > 256) blocks = 256;                     // `partial` holds 2*256 floats
    pixel_ce_kernel<<<blocks, 256, 0, stream>>>(logits_nchw, target_masks, n, kp1, hf, wf, hm, wm, label_smoothing,
                                                partial);
    BPB_LAUNCH_CHECK("pixel_ce_kernel");
    pixel_ce_finalize_kernel<<<1, 32, 0, stream>>>(partial, blocks, 1.f / static_cast<float>(total), out2);
    BPB_LAUNCH_CHECK("pixel_ce_finalize_kernel");
    return 0;
}

// out[0] = sum_i w[i] * *terms[i]   (GiLt_loss.py:92 weighted sum of the individual losses)
struct WsumArgs { const float* t[16]; float w[16]; int n; };
__global__ void weighted_sum_kernel(WsumArgs a, float* out) {
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < a.n; ++i) s += a.w[i] * a.t[i][0];
        out[0] = s;
    }
}
int weighted_sum(const float* const* terms, const float* weights, int n, float* out, cudaStream_t stream) {
    BPB_CHECK(n >= 1 && n <= 16, "weighted_sum: n=%d out of range", n);
    WsumArgs a{};
    a.n = n;
    for (int i = 0; i < n; ++i) { a.t[i] = terms[i]; a.w[i] = weights[i]; }
    weighted_sum_kernel<<<1, 32, 0, stream>>>(a, out);
    BPB_LAUNCH_CHECK("weighted_sum_kernel");
    return 0;
}

}  // namespace bpb
