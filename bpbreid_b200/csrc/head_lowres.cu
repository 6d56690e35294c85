// BPBreID head WITHOUT materialising the 1920-channel stride-4 map (HRNet).
//
// The reference upsamples branches 1..3 bilinearly (align_corners) to 64x32, concatenates to [N,1920,64,32]
// (hrnet.py:568-573), then BatchNorms / classifies / pools that map (bpbreid.py:147-202, 376-385, 458-503).
// Upsampling U_b, the 1x1 classifier and the pooling sums are all LINEAR, so with X_b the low-resolution branch map
//   * logits      = sum_b U_b (Wf_b X_b) + bf            -> classify at LOW resolution, upsample K+1 channels
//   * pooled_j    = sum_p m_j[p] (U_b X_b)[p] = (U_b^T m_j) . X_b   -> pool at LOW resolution with transposed masks
//   * BN mean     = linear as well; the BN *variance* is quadratic, so sum (U_b x)^2 is accumulated on the fly
//     from a shared-memory copy of the branch slab (no write of the upsampled values).
// HBM traffic: 3 reads of the branch maps (0.98 MB/img bf16) instead of writing 7.9 MB/img and reading it 3 times.
// All arithmetic fp32; results equal the reference's up to fp32 re-association (the reference rounds nothing either).
#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

struct Tap {
    int i0, i1;
    float l;
};
// PyTorch upsample_bilinear2d, align_corners=True: src = dst * (in-1)/(out-1)
__device__ __forceinline__ Tap make_tap(int o, int in_size, int out_size) {
    const float r = out_size > 1 ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.f;
    const float s = r * o;
    Tap t;
    t.i0 = static_cast<int>(s);
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l = s - t.i0;
    return t;
}

__device__ __forceinline__ void unpack8f(const uint4& u, float* f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// ----------------------------------------------------------------------------------------------
// (1) per-image partial BN statistics of the (virtually) upsampled branch: partial[n][2][ctot] at channel c_off
//     grid (N, Cb/64), block 256 = 8 channel-vectors x 32 pixel lanes
// ----------------------------------------------------------------------------------------------
// The variance term is quadratic in the branch values, but the bilinear operator is separable, U = Uy (x) Ux, so
//   sum_hi (U x)^2 = x^T (Ty (x) Tx) x,   Ty = Uy^T Uy, Tx = Ux^T Ux  (tridiagonal: an output row touches <= 2 input rows)
// i.e. a 3x3 stencil at LOW resolution (9 MACs per low-res value instead of 5 flops per HIGH-res value: 8-32x fewer
// for the stride-8/16/32 branches); the mean term uses the column sums of Uy and Ux.
__global__ void __launch_bounds__(256) branch_stats_kernel(const uint4* __restrict__ x, int hb, int wb, int cb, int h,
                                                           int w, float* __restrict__ partial, int c_off, int ctot) {
    extern __shared__ uint4 slab[];                     // [hb*wb][8] vectors (64 channels) -- unused when hb == h
    __shared__ float coef[6][64];                       // y: colsum, diag, offdiag(i,i+1) ; x: the same
    const int n = blockIdx.x, sl = blockIdx.y;
    const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int cvec = cb / 8;
    const uint4* base = x + static_cast<int64_t>(n) * hb * wb * cvec + sl * 8;
    const bool direct = (hb == h && wb == w);
    float s[8], q[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { s[t] = 0.f; q[t] = 0.f; }
    if (direct) {
#pragma unroll 4
        for (int p = pl; p < h * w; p += 32) {
            float v[8];
            unpack8f(base[static_cast<int64_t>(p) * cvec + cv], v);
#pragma unroll
            for (int t = 0; t < 8; ++t) { s[t] += v[t]; q[t] += v[t] * v[t]; }
        }
    } else {
        for (int i = threadIdx.x; i < hb * wb * 8; i += 256) slab[i] = base[static_cast<int64_t>(i >> 3) * cvec + (i & 7)];
        if (threadIdx.x < 2) {                          // thread 0: y axis, thread 1: x axis
            const int in = threadIdx.x == 0 ? hb : wb, out = threadIdx.x == 0 ? h : w;
            float* cs = coef[threadIdx.x * 3 + 0];
            float* dg = coef[threadIdx.x * 3 + 1];
            float* od = coef[threadIdx.x * 3 + 2];
            for (int i = 0; i < in; ++i) { cs[i] = 0.f; dg[i] = 0.f; od[i] = 0.f; }
            for (int o = 0; o < out; ++o) {
                const Tap tp = make_tap(o, in, out);
                const float a = 1.f - tp.l, b = tp.l;
                if (tp.i1 == tp.i0) {                   // last input row: both taps coincide
                    cs[tp.i0] += a + b; dg[tp.i0] += (a + b) * (a + b);
                } else {
                    cs[tp.i0] += a; cs[tp.i1] += b;
                    dg[tp.i0] += a * a; dg[tp.i1] += b * b; od[tp.i0] += a * b;
                }
            }
        }
        __syncthreads();
        const float* csy = coef[0]; const float* dgy = coef[1]; const float* ody = coef[2];
        const float* csx = coef[3]; const float* dgx = coef[4]; const float* odx = coef[5];
        for (int p = pl; p < hb * wb; p += 32) {
            const int iy = p / wb, ix = p - iy * wb;
            // stencil weights of the three rows / columns around (iy, ix); zero outside the map
            const float wy[3] = {iy > 0 ? ody[iy - 1] : 0.f, dgy[iy], iy + 1 < hb ? ody[iy] : 0.f};
            const float wx[3] = {ix > 0 ? odx[ix - 1] : 0.f, dgx[ix], ix + 1 < wb ? odx[ix] : 0.f};
            float z[8], ctr[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) z[t] = 0.f;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = iy + dy;
                if (yy < 0 || yy >= hb) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = ix + dx;
                    if (xx < 0 || xx >= wb) continue;
                    float v[8];
                    unpack8f(slab[(yy * wb + xx) * 8 + cv], v);
                    const float wgt = wy[dy + 1] * wx[dx + 1];
#pragma unroll
                    for (int t = 0; t < 8; ++t) z[t] += wgt * v[t];
                    if (dy == 0 && dx == 0) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) ctr[t] = v[t];
                    }
                }
            }
            const float cw = csy[iy] * csx[ix];
#pragma unroll
            for (int t = 0; t < 8; ++t) { s[t] += cw * ctr[t]; q[t] += ctr[t] * z[t]; }
        }
    }
    __shared__ float red[2][32][64];
#pragma unroll
    for (int t = 0; t < 8; ++t) { red[0][pl][cv * 8 + t] = s[t]; red[1][pl][cv * 8 + t] = q[t]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
        float acc = 0.f;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) acc += red[which][l][ch];
        partial[(static_cast<int64_t>(n) * 2 + which) * ctot + c_off + sl * 64 + ch] = acc;
    }
}

int head_branch_stats(const void* x, int n, int hb, int wb, int cb, int h, int w, float* partial, int c_off, int ctot,
                      cudaStream_t stream) {
    BPB_CHECK(cb % 64 == 0, "head_branch_stats: channels %d must be a multiple of 64", cb);
    const bool direct = hb == h && wb == w;
    const size_t smem = direct ? 0 : static_cast<size_t>(hb) * wb * 8 * sizeof(uint4);
    BPB_CHECK(smem <= 160 * 1024, "head_branch_stats: branch map %dx%d too large for shared memory", hb, wb);
    BPB_CHECK(direct || (hb <= 64 && wb <= 64), "head_branch_stats: low-resolution branch %dx%d larger than 64", hb, wb);
    BPB_SET_MAX_SMEM(branch_stats_kernel, 160 * 1024);
    dim3 grid(n, cb / 64);
    branch_stats_kernel<<<grid, 256, smem, stream>>>(reinterpret_cast<const uint4*>(x), hb, wb, cb, h, w, partial, c_off,
                                                     ctot);
    BPB_LAUNCH_CHECK("branch_stats_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// (2) low-resolution logits of one branch: out[row][k] = sum_c wf[k][c_off + c] * x[row][c]   (no bias)
//     one warp handles 4 rows (pixels) at a time, folded weights of the branch staged in shared memory
// ----------------------------------------------------------------------------------------------
template <int KP1>
__global__ void __launch_bounds__(256) lowres_logits_kernel(const uint4* __restrict__ x, const float* __restrict__ wf_g,
                                                            int64_t rows, int cb, int c_off, int ctot,
                                                            float* __restrict__ out) {
    extern __shared__ float wf[];                        // [KP1][cb]
    for (int i = threadIdx.x; i < KP1 * cb; i += blockDim.x) wf[i] = wf_g[(i / cb) * ctot + c_off + (i % cb)];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cvec = cb / 8;
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    for (int64_t base = (static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp) * 4; base < rows;
         base += warps_total * 4) {
        float acc[4][KP1];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int k = 0; k < KP1; ++k) acc[p][k] = 0.f;
        for (int cv = lane; cv < cvec; cv += 32) {
            float f[4][8];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int64_t r = base + p < rows ? base + p : rows - 1;
                unpack8f(x[r * cvec + cv], f[p]);
            }
#pragma unroll
            for (int k = 0; k < KP1; ++k) {
                const float4 w0 = *reinterpret_cast<const float4*>(wf + k * cb + cv * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(wf + k * cb + cv * 8 + 4);
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    acc[p][k] += f[p][0] * w0.x + f[p][1] * w0.y + f[p][2] * w0.z + f[p][3] * w0.w + f[p][4] * w1.x +
                                 f[p][5] * w1.y + f[p][6] * w1.z + f[p][7] * w1.w;
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int k = 0; k < KP1; ++k) acc[p][k] = warp_sum(acc[p][k]);
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (lane == p && base + p < rows) {
#pragma unroll
                for (int k = 0; k < KP1; ++k) out[(base + p) * KP1 + k] = acc[p][k];
            }
    }
}

int head_lowres_logits(const void* x, const float* wf, int64_t rows, int cb, int c_off, int ctot, int kp1, float* out,
                       cudaStream_t stream) {
    BPB_CHECK(cb % 8 == 0 && kp1 >= 2 && kp1 <= 12, "head_lowres_logits: bad shape (cb=%d kp1=%d)", cb, kp1);
    const size_t smem = static_cast<size_t>(kp1) * cb * sizeof(float);
    int blocks = ceil_div(rows, 8 * 4);
    if (blocks > 148 * 4) blocks = 148 * 4;
#define BPB_LL_CASE(K)                                                                                         \
    case K: {                                                                                                  \
        auto kern = lowres_logits_kernel<K>;                                                                   \
        BPB_SET_MAX_SMEM(kern, 100 * 1024);                                                                                                      \
        kern<<<blocks, 256, smem, stream>>>(reinterpret_cast<const uint4*>(x), wf, rows, cb, c_off, ctot, out); \
        break;                                                                                                 \
    }
    switch (kp1) {
        BPB_LL_CASE(2) BPB_LL_CASE(3) BPB_LL_CASE(4) BPB_LL_CASE(5) BPB_LL_CASE(6) BPB_LL_CASE(7) BPB_LL_CASE(8)
        BPB_LL_CASE(9) BPB_LL_CASE(10) BPB_LL_CASE(11) BPB_LL_CASE(12)
    }
#undef BPB_LL_CASE
    BPB_LAUNCH_CHECK("lowres_logits_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// (3)+(4) one CTA per image: combine the branch logits at full resolution (bilinear), softmax, masks, visibility,
//         mask sums, and the TRANSPOSED-upsampled masks M'_b[j][q] = sum_p U_b[p][q] m_j[p] for every branch
//         (rows j: 0 global(=1), 1 foreground, 2 background, 3.. parts).
// ----------------------------------------------------------------------------------------------
struct BranchGeom {
    int hb[4], wb[4];
    const float* lg[4];        // low-res logits [N*hb*wb][KP1]
    float* mt[4];              // transposed masks [N][J][hb*wb]
    int nb;
};

template <int KP1>
__global__ void __launch_bounds__(256) pixel_head_kernel(BranchGeom g, const float* __restrict__ bf, int h, int w,
                                                         float* __restrict__ logits, float* __restrict__ probs,
                                                         float* __restrict__ fg_mask, int* __restrict__ vis_bits,
                                                         float* __restrict__ vis_cont, float* __restrict__ mask_sum,
                                                         const float* __restrict__ ext, int mode) {
    constexpr int J = KP1 + 2;
    const bool learn = (mode & 1) == 0;
    extern __shared__ float acc_mt[];                   // concatenated [J][Q_b] for b = 1..nb-1
    const int n = blockIdx.x;
    const int hw = h * w;
    int off[4], total = 0;
    off[0] = 0;
    for (int b = 1; b < g.nb; ++b) { off[b] = total; total += J * g.hb[b] * g.wb[b]; }
    for (int i = threadIdx.x; i < total; i += blockDim.x) acc_mt[i] = 0.f;
    __shared__ int s_bits;
    __shared__ float s_vmax[KP1];
    __shared__ float s_msum[8][J];
    if (threadIdx.x == 0) s_bits = 0;
    if (threadIdx.x < KP1) s_vmax[threadIdx.x] = 0.f;
    __syncthreads();
    float msum[J];
#pragma unroll
    for (int j = 0; j < J; ++j) msum[j] = 0.f;
    int bits = 0;
    float vmax[KP1];
#pragma unroll
    for (int k = 0; k < KP1; ++k) vmax[k] = 0.f;
    float* mt0 = g.mt[0] + static_cast<int64_t>(n) * J * hw;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
        const int oy = p / w, ox = p % w;
        float l[KP1];
        if (learn) {
            const float* l0 = g.lg[0] + (static_cast<int64_t>(n) * hw + p) * KP1;
#pragma unroll
            for (int k = 0; k < KP1; ++k) l[k] = bf[k] + l0[k];
        }
        Tap ty[4], tx[4];
        for (int b = 1; b < g.nb; ++b) {
            ty[b] = make_tap(oy, g.hb[b], h);
            tx[b] = make_tap(ox, g.wb[b], w);
            if (!learn) continue;
            const float* lb = g.lg[b] + static_cast<int64_t>(n) * g.hb[b] * g.wb[b] * KP1;
            const float w00 = (1.f - ty[b].l) * (1.f - tx[b].l), w01 = (1.f - ty[b].l) * tx[b].l;
            const float w10 = ty[b].l * (1.f - tx[b].l), w11 = ty[b].l * tx[b].l;
            const float* a = lb + (ty[b].i0 * g.wb[b] + tx[b].i0) * KP1;
            const float* bb = lb + (ty[b].i0 * g.wb[b] + tx[b].i1) * KP1;
            const float* c = lb + (ty[b].i1 * g.wb[b] + tx[b].i0) * KP1;
            const float* d = lb + (ty[b].i1 * g.wb[b] + tx[b].i1) * KP1;
#pragma unroll
            for (int k = 0; k < KP1; ++k) l[k] += w00 * a[k] + w01 * bb[k] + w10 * c[k] + w11 * d[k];
        }
        float pr[KP1], pm[KP1], ex[KP1];
        if (ext != nullptr) {
#pragma unroll
            for (int k = 0; k < KP1; ++k) ex[k] = ext[(static_cast<int64_t>(n) * KP1 + k) * hw + p];
        }
        if (learn) {
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < KP1; ++k) mx = fmaxf(mx, l[k]);
            float e[KP1], den = 0.f;
#pragma unroll
            for (int k = 0; k < KP1; ++k) { e[k] = expf(l[k] - mx); den += e[k]; }
            const float inv = 1.f / den;
#pragma unroll
            for (int k = 0; k < KP1; ++k) {
                pr[k] = e[k] * inv;
                logits[(static_cast<int64_t>(n) * KP1 + k) * hw + p] = l[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < KP1; ++k) pr[k] = ex[k];
        }
        apply_target_segmentation<KP1>(pr, pm, ex, mode);
        float m[J];                                       // mask rows: global, fg, bg, parts
        m[0] = 1.f;
        int arg = 0;
        float best = -1.f, fg = -1.f;
#pragma unroll
        for (int k = 0; k < KP1; ++k) {
            probs[(static_cast<int64_t>(n) * KP1 + k) * hw + p] = pm[k];
            if (pr[k] > best) { best = pr[k]; arg = k; }
            if (k >= 1) fg = fmaxf(fg, pm[k]);
            vmax[k] = fmaxf(vmax[k], pr[k]);
            if (k == 0) m[2] = pm[k]; else m[2 + k] = pm[k];
        }
        m[1] = fg;
        fg_mask[static_cast<int64_t>(n) * hw + p] = fg;
        bits |= 1 << arg;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            msum[j] += m[j];
            mt0[j * hw + p] = m[j];
        }
        for (int b = 1; b < g.nb; ++b) {
            const int qb = g.hb[b] * g.wb[b];
            float* a = acc_mt + off[b];
            const float w00 = (1.f - ty[b].l) * (1.f - tx[b].l), w01 = (1.f - ty[b].l) * tx[b].l;
            const float w10 = ty[b].l * (1.f - tx[b].l), w11 = ty[b].l * tx[b].l;
            const int q00 = ty[b].i0 * g.wb[b] + tx[b].i0, q01 = ty[b].i0 * g.wb[b] + tx[b].i1;
            const int q10 = ty[b].i1 * g.wb[b] + tx[b].i0, q11 = ty[b].i1 * g.wb[b] + tx[b].i1;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                atomicAdd(a + j * qb + q00, w00 * m[j]);
                if (w01 != 0.f) atomicAdd(a + j * qb + q01, w01 * m[j]);
                if (w10 != 0.f) atomicAdd(a + j * qb + q10, w10 * m[j]);
                if (w11 != 0.f) atomicAdd(a + j * qb + q11, w11 * m[j]);
            }
        }
    }
    // block reductions: visibility bits / maxima, mask sums
    atomicOr(&s_bits, bits);
#pragma unroll
    for (int k = 0; k < KP1; ++k) atomicMax(reinterpret_cast<int*>(&s_vmax[k]), __float_as_int(vmax[k]));
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float v = warp_sum(msum[j]);
        if ((threadIdx.x & 31) == 0) s_msum[threadIdx.x >> 5][j] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) vis_bits[n] = s_bits;
    if (threadIdx.x < KP1) vis_cont[n * KP1 + threadIdx.x] = s_vmax[threadIdx.x];
    if (threadIdx.x < J) {
        float t = 0.f;
        for (int wv = 0; wv < 8; ++wv) t += s_msum[wv][threadIdx.x];
        mask_sum[n * J + threadIdx.x] = t;
    }
    for (int b = 1; b < g.nb; ++b) {
        const int cnt = J * g.hb[b] * g.wb[b];
        float* dst = g.mt[b] + static_cast<int64_t>(n) * cnt;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = acc_mt[off[b] + i];
    }
}

int head_pixel(const float* const* lowres_logits, float* const* mask_t, const int* hb, const int* wb, int nb,
               const float* bf, int n, int h, int w, int kp1, float* logits, float* probs, float* fg_mask,
               int* vis_bits, float* vis_cont, float* mask_sum, const float* ext, int mode, cudaStream_t stream) {
    BPB_CHECK(nb >= 1 && nb <= 4, "head_pixel: nb=%d", nb);
    BPB_CHECK(mode == 0 || ext != nullptr, "head_pixel: mode %d needs the external masks", mode);
    BPB_CHECK(hb[0] == h && wb[0] == w, "head_pixel: branch 0 must be at the output resolution");
    BranchGeom g{};
    g.nb = nb;
    size_t smem = 0;
    for (int b = 0; b < nb; ++b) {
        g.hb[b] = hb[b]; g.wb[b] = wb[b]; g.lg[b] = lowres_logits ? lowres_logits[b] : nullptr; g.mt[b] = mask_t[b];
        if (b > 0) smem += static_cast<size_t>(kp1 + 2) * hb[b] * wb[b] * sizeof(float);
    }
    BPB_CHECK(smem <= 160 * 1024, "head_pixel: transposed masks do not fit shared memory");
#define BPB_PH_CASE(K)                                                                                         \
    case K: {                                                                                                  \
        auto kern = pixel_head_kernel<K>;                                                                      \
        BPB_SET_MAX_SMEM(kern, 160 * 1024);                                                                                                      \
        kern<<<n, 256, smem, stream>>>(g, bf, h, w, logits, probs, fg_mask, vis_bits, vis_cont, mask_sum, ext,  \
                                       mode);                                                                  \
        break;                                                                                                 \
    }
    switch (kp1) {
        BPB_PH_CASE(2) BPB_PH_CASE(3) BPB_PH_CASE(4) BPB_PH_CASE(5) BPB_PH_CASE(6) BPB_PH_CASE(7) BPB_PH_CASE(8)
        BPB_PH_CASE(9) BPB_PH_CASE(10) BPB_PH_CASE(11) BPB_PH_CASE(12)
        default: BPB_CHECK(false, "head_pixel: K+1=%d out of range", kp1);
    }
#undef BPB_PH_CASE
    BPB_LAUNCH_CHECK("pixel_head_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// (5) pooled[j][n][c_off + c] = norm_j * sum_q mt[n][j][q] * x[n][q][c]
//     norm: rows 0..2 -> 1/(h*w) ; parts -> 1/clamp(mask_sum[n][j], 1e-6)        (bpbreid.py:485-486, :498-501)
//     CTA = (image, 256-channel slab); 4 warps split the low-res pixels; lane owns 8 channels.
// ----------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(128) branch_pool_kernel(const uint4* __restrict__ x, const float* __restrict__ mt,
                                                          const float* __restrict__ mask_sum, int n, int q, int cb,
                                                          int c_off, int ctot, float inv_hw,
                                                          float* __restrict__ pooled) {
    extern __shared__ float wts[];                        // [q][J]
    const int im = blockIdx.x, slab = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cvec = cb / 8;
    const float* src = mt + static_cast<int64_t>(im) * J * q;
    for (int i = threadIdx.x; i < q * J; i += blockDim.x) {
        const int j = i / q, qq = i % q;
        wts[qq * J + j] = src[i];
    }
    __syncthreads();
    float acc[J][8];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[j][t] = 0.f;
    // a slab narrower than 32 channel vectors (C = 128) lets the spare lanes take other pixels
    int lpp = cvec - slab * 32 >= 32 ? 32 : cvec - slab * 32;         // lanes per pixel
    if (lpp <= 0 || (32 % lpp) != 0) lpp = 32;
    const int ppw = 32 / lpp;                                         // pixels per warp iteration
    const int sub = lane / lpp, cl = lane - sub * lpp;
    const int cvl = slab * 32 + cl;
    if (cvl < cvec) {
        const uint4* base = x + static_cast<int64_t>(im) * q * cvec + cvl;
        const int step = 4 * ppw;
        int p = warp * ppw + sub;
        for (; p + 3 * step < q; p += 4 * step) {                     // four loads in flight
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = base[static_cast<int64_t>(p + k * step) * cvec];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f[8];
                unpack8f(u[k], f);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float wj = wts[(p + k * step) * J + j];
#pragma unroll
                    for (int t = 0; t < 8; ++t) acc[j][t] += wj * f[t];
                }
            }
        }
        for (; p < q; p += step) {
            float f[8];
            unpack8f(base[static_cast<int64_t>(p) * cvec], f);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float wj = wts[p * J + j];
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[j][t] += wj * f[t];
            }
        }
    }
    __syncthreads();
    float* red = wts;                                     // [4][32][J*8]
    constexpr int kPer = J * 8;
    float* mine = red + (warp * 32 + lane) * kPer;
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int t = 0; t < 8; ++t) mine[j * 8 + t] = acc[j][t];
    __syncthreads();
    if (warp == 0 && sub == 0 && cvl < cvec) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float norm = j < 3 ? inv_hw : 1.f / fmaxf(mask_sum[im * J + j], 1e-6f);
            float o[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float sum = 0.f;
                for (int wv = 0; wv < 4; ++wv)
                    for (int sb = 0; sb < ppw; ++sb) sum += red[(wv * 32 + sb * lpp + cl) * kPer + j * 8 + t];
                o[t] = sum * norm;
            }
            float* dst = pooled + (static_cast<int64_t>(j) * n + im) * ctot + c_off + cvl * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

int head_branch_pool(const void* x, const float* mt, const float* mask_sum, int n, int q, int cb, int c_off, int ctot,
                     int kp1, int hw_full, float* pooled, cudaStream_t stream) {
    BPB_CHECK(cb % 8 == 0, "head_branch_pool: channels %d", cb);
    const int J = kp1 + 2;
    size_t smem = static_cast<size_t>(q) * J * sizeof(float);
    const size_t red = static_cast<size_t>(128) * J * 8 * sizeof(float);
    if (red > smem) smem = red;
    BPB_CHECK(smem <= 200 * 1024, "head_branch_pool: %d px x %d rows does not fit shared memory", q, J);
    dim3 grid(n, ceil_div(cb / 8, 32));
    const float inv_hw = 1.f / static_cast<float>(hw_full);
#define BPB_BP_CASE(JJ)                                                                                        \
    case JJ: {                                                                                                 \
        auto kern = branch_pool_kernel<JJ>;                                                                    \
        BPB_SET_MAX_SMEM(kern, 200 * 1024);                                                                                                      \
        kern<<<grid, 128, smem, stream>>>(reinterpret_cast<const uint4*>(x), mt, mask_sum, n, q, cb, c_off, ctot, \
                                          inv_hw, pooled);                                                     \
        break;                                                                                                 \
    }
    switch (J) {
        BPB_BP_CASE(4) BPB_BP_CASE(5) BPB_BP_CASE(6) BPB_BP_CASE(7) BPB_BP_CASE(8) BPB_BP_CASE(9) BPB_BP_CASE(10)
        BPB_BP_CASE(11) BPB_BP_CASE(12) BPB_BP_CASE(13) BPB_BP_CASE(14)
        default: BPB_CHECK(false, "head_branch_pool: K+1=%d out of range", kp1);
    }
#undef BPB_BP_CASE
    BPB_LAUNCH_CHECK("branch_pool_kernel");
    return 0;
}

}  // namespace bpb
