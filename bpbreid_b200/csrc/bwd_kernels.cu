// Backward companions of the tensor-core convolutions (training step, loss.backward() at
// torchreid/engine/image/part_based_engine.py:98-103): ReLU / residual gradient merge, train-mode BatchNorm backward
// (two HBM passes: per-channel reductions with a fused last-block finalize, then the elementwise gradient), stem
// weight gradient, weight (re)packing for the forward / dgrad tensor-core operands fused with the Adam update.
// Activations and their gradients are NHWC bf16; parameter gradients fp32.
// Reference semantics: autograd of nn.BatchNorm2d (training), nn.ReLU, residual adds and the HRNet fuse sum
// (torchreid/models/hrnet.py:80-96,116-137,262-279); torch.optim.Adam as built by torchreid/optim/optimizer.py.
#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

__device__ __forceinline__ void unpack8b(const uint4& u, float* f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8b(const float* f) {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    return o;
}

// ----------------------------------------------------------------------------------------------
// g = (sum_s src_s) * [y > 0]        (merge of the gradients of a tensor's consumers + ReLU backward)
// ----------------------------------------------------------------------------------------------
struct GradSumArgs {
    const uint4* src[4];
    int nsrc;
};

__device__ __forceinline__ uint4 relu_grad_sum_one(const uint4* v, int nsrc, bool gate, const uint4& yv) {
    float acc[8];
    unpack8b(v[0], acc);
#pragma unroll
    for (int s = 1; s < 4; ++s) {
        if (s < nsrc) {
            float f[8];
            unpack8b(v[s], f);
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[t] += f[t];
        }
    }
    if (gate) {
        float yy[8];
        unpack8b(yv, yy);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = yy[t] > 0.f ? acc[t] : 0.f;
    }
    return pack8b(acc);
}

__global__ void __launch_bounds__(256) relu_grad_sum_kernel(GradSumArgs a, const uint4* __restrict__ y,
                                                            uint4* __restrict__ out, int64_t total) {
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (; i + step < total; i += 2 * step) {                 // two vectors per trip: all loads issued before any use
        uint4 v0[4], v1[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s < a.nsrc) {
                v0[s] = __ldg(a.src[s] + i);
                v1[s] = __ldg(a.src[s] + i + step);
            }
        }
        const uint4 y0 = y ? __ldg(y + i) : zero, y1 = y ? __ldg(y + i + step) : zero;
        out[i] = relu_grad_sum_one(v0, a.nsrc, y != nullptr, y0);
        out[i + step] = relu_grad_sum_one(v1, a.nsrc, y != nullptr, y1);
    }
    for (; i < total; i += step) {
        uint4 v0[4];
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s < a.nsrc) v0[s] = a.src[s][i];
        const uint4 y0 = y ? y[i] : zero;
        out[i] = relu_grad_sum_one(v0, a.nsrc, y != nullptr, y0);
    }
}

int relu_grad_sum(const void* const* srcs, int nsrc, const void* y, void* out, int64_t m, int c, cudaStream_t stream) {
    BPB_CHECK(nsrc >= 1 && nsrc <= 4, "relu_grad_sum: nsrc=%d out of range", nsrc);
    BPB_CHECK(c % 8 == 0, "relu_grad_sum: C=%d must be a multiple of 8", c);
    GradSumArgs a{};
    a.nsrc = nsrc;
    for (int s = 0; s < nsrc; ++s) a.src[s] = reinterpret_cast<const uint4*>(srcs[s]);
    const int64_t total = m * (c / 8);
    int blocks = ceil_div(total, 512);
    if (blocks > 148 * 16) blocks = 148 * 16;
    relu_grad_sum_kernel<<<blocks, 256, 0, stream>>>(a, reinterpret_cast<const uint4*>(y), reinterpret_cast<uint4*>(out),
                                                     total);
    BPB_LAUNCH_CHECK("relu_grad_sum_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Train-mode BatchNorm backward.  y = gamma * xhat + beta, xhat = (raw - mean) * invstd, batch statistics.
//   g      = dL/dy at the BN output; optionally gated by the ReLU that follows (relu_affine: raw*scale+shift > 0) and /
//            or gathered from a 2^up-times finer map (adjoint of the nearest upsample inside the HRNet fuse sum)
//   dbeta  = sum g ; dgamma = sum g*xhat
//   draw   = gamma*invstd * (g - dbeta/M - xhat * dgamma/M)
// Pass 1 (reduce) accumulates the two sums per channel (fp32 in registers -> fp64 atomics, last block finalizes into
// coef[2][C] = {dbeta/M, dgamma/M} and the parameter gradients); pass 2 (apply) writes draw as bf16, optionally in the
// zero-inserted 2x layout the stride-2 dgrad / wgrad consume.
// ----------------------------------------------------------------------------------------------
struct BnBwdArgs {
    const uint4* dy;
    const uint4* raw;
    const float* mean_invstd;   // [2][C]
    const float* scale;         // gamma * invstd  [C]
    const float* shift;
    int n, h, w, cvec;          // geometry of `raw`
    int up_log2;                // dy lives on a (h << up) x (w << up) map
    int relu_affine;
    int c_lo;                   // parameters cover channels [c_lo, C)
};

// eight consecutive per-channel constants as two 16-byte loads (40 scalar loads per thread made the preamble of these
// kernels longer than the streaming pass on the small maps)
__device__ __forceinline__ void load8f(const float* __restrict__ p, float* dst) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
    dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
}

__device__ __forceinline__ void bn_bwd_load_g(const BnBwdArgs& a, unsigned r, int cv, const float* rawf,
                                              const float* sc, const float* sh, float* g) {
    if (a.up_log2 == 0) {
        unpack8b(a.dy[static_cast<size_t>(r) * a.cvec + cv], g);
    } else {
        const unsigned ox = r % a.w, t = r / a.w;
        const unsigned oy = t % a.h, im = t / a.h;
        const int f = 1 << a.up_log2;
        const unsigned hh = a.h << a.up_log2, ww = a.w << a.up_log2;
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = 0.f;
        for (int dy = 0; dy < f; ++dy)
            for (int dx = 0; dx < f; ++dx) {
                float v[8];
                unpack8b(a.dy[(static_cast<size_t>(im * hh + oy * f + dy) * ww + ox * f + dx) * a.cvec + cv], v);
#pragma unroll
                for (int q = 0; q < 8; ++q) g[q] += v[q];
            }
    }
    if (a.relu_affine) {
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = (rawf[q] * sc[q] + sh[q]) > 0.f ? g[q] : 0.f;
    }
}

// Sum `nrows` rows of `cols` floats (row stride `cols`) with all 256 threads: thread = (4-column group, row slice), up to
// four independent 16-byte loads in flight, fp64 accumulation, slices combined through shared memory.  The column totals
// land in dred[0 .. min(cols, 1024)) for the chunk starting at column `c0`; `emit(col, total)` is called once per column.
template <typename Emit>
__device__ __forceinline__ void sum_partial_rows(const float* __restrict__ base, int nrows, int cols, double* dred,
                                                 Emit emit) {
    const int cols4 = cols >> 2;
    const int per = cols4 < 256 ? cols4 : 256;                     // 4-column groups handled per chunk
    const int nsl = 256 / per;                                     // row slices
    const int g = static_cast<int>(threadIdx.x) % per, sl = static_cast<int>(threadIdx.x) / per;
    for (int g0 = 0; g0 < cols4; g0 += per) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (sl < nsl && g0 + g < cols4) {
            const float4* p = reinterpret_cast<const float4*>(base) + g0 + g;
            int r = sl;
            for (; r + 3 * nsl < nrows; r += 4 * nsl) {
                const float4 v0 = __ldcg(p + static_cast<size_t>(r) * cols4);
                const float4 v1 = __ldcg(p + static_cast<size_t>(r + nsl) * cols4);
                const float4 v2 = __ldcg(p + static_cast<size_t>(r + 2 * nsl) * cols4);
                const float4 v3 = __ldcg(p + static_cast<size_t>(r + 3 * nsl) * cols4);
                a0 += (static_cast<double>(v0.x) + static_cast<double>(v1.x)) + (static_cast<double>(v2.x) + static_cast<double>(v3.x));
                a1 += (static_cast<double>(v0.y) + static_cast<double>(v1.y)) + (static_cast<double>(v2.y) + static_cast<double>(v3.y));
                a2 += (static_cast<double>(v0.z) + static_cast<double>(v1.z)) + (static_cast<double>(v2.z) + static_cast<double>(v3.z));
                a3 += (static_cast<double>(v0.w) + static_cast<double>(v1.w)) + (static_cast<double>(v2.w) + static_cast<double>(v3.w));
            }
            for (; r < nrows; r += nsl) {
                const float4 v = __ldcg(p + static_cast<size_t>(r) * cols4);
                a0 += static_cast<double>(v.x);
                a1 += static_cast<double>(v.y);
                a2 += static_cast<double>(v.z);
                a3 += static_cast<double>(v.w);
            }
        }
        __syncthreads();                                           // dred may still be read by the previous chunk
        if (sl < nsl) {
            double* d = dred + (static_cast<size_t>(sl) * per + g) * 4;
            d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < per * 4; i += 256) {
            const int col = g0 * 4 + i;
            if (col < cols) {
                double t = 0.0;
                for (int l = 0; l < nsl; ++l) t += dred[static_cast<size_t>(l) * per * 4 + i];
                emit(col, t);
            }
        }
    }
}

constexpr int kBnRedBlocks = 148 * 4;      // reduce pass: one partial row per block
constexpr int kBnRedGroup = 16;            // blocks per first-level group of the two-level deterministic row sum
constexpr int kBnRedGroups = (kBnRedBlocks + kBnRedGroup - 1) / kBnRedGroup;

// workspace: [kBnRedBlocks][2C] block rows | [kBnRedGroups][2C] group rows | kBnRedGroups + 1 counters
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(BnBwdArgs a, float* __restrict__ ws, double inv_count,
                                                               float* __restrict__ coef, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta) {
    const int width = a.cvec < 256 ? a.cvec : 256;
    const int rows_per_block = 256 / width;
    const int tx = threadIdx.x % width, ty = threadIdx.x / width;
    const unsigned m = static_cast<unsigned>(a.n) * a.h * a.w;
    const int c = a.cvec * 8;
    const int cols = 2 * c;
    __shared__ __align__(16) float red[2][256][8];
    __shared__ int s_flag;
    float* ws2 = ws + static_cast<size_t>(kBnRedBlocks) * cols;
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws2 + static_cast<size_t>(kBnRedGroups) * cols);
    // q accumulates g * (raw - mean); the 1/std factor is applied once per channel by the finalize
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    const int cv = tx;
    if (ty < rows_per_block && cv < a.cvec) {
        float mean[8], sc[8], sh[8];
        load8f(a.mean_invstd + cv * 8, mean);
        if (a.relu_affine) {
            load8f(a.scale + cv * 8, sc);
            load8f(a.shift + cv * 8, sh);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { sc[j] = 0.f; sh[j] = 1.f; }     // no ReLU: gate always open
        }
        const unsigned step = gridDim.x * rows_per_block;
        unsigned r = blockIdx.x * rows_per_block + ty;
        if (a.up_log2 == 0) {
            // four rows per trip: eight independent 16-byte loads in flight per thread.  (One row per trip left the
            // pass latency-bound at ~3 TB/s: 24 % active warps, long-scoreboard on top in ncu.)
            for (; r + 3 * step < m; r += 4 * step) {
                uint4 xr[4], dr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    xr[u] = __ldg(a.raw + static_cast<size_t>(r + u * step) * a.cvec + cv);
                    dr[u] = __ldg(a.dy + static_cast<size_t>(r + u * step) * a.cvec + cv);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float x[8], g[8];
                    unpack8b(xr[u], x);
                    unpack8b(dr[u], g);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float gg = (x[j] * sc[j] + sh[j]) > 0.f ? g[j] : 0.f;
                        s[j] += gg;
                        q[j] += gg * (x[j] - mean[j]);
                    }
                }
            }
        }
        for (; r < m; r += step) {
            float rawf[8], g[8];
            unpack8b(a.raw[static_cast<size_t>(r) * a.cvec + cv], rawf);
            bn_bwd_load_g(a, r, cv, rawf, sc, sh, g);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] += g[j];
                q[j] += g[j] * (rawf[j] - mean[j]);
            }
        }
    }
    // block reduction over the row lanes: threads that share a channel vector are `width` apart.  Inside a warp that is a
    // shuffle tree (a serial loop over up to 64 row lanes on `width` threads cost more than the whole HBM pass for the
    // narrow C = 32 / 64 layers); across warps one pass over at most 8 partial rows in shared memory.
    if (width < 32) {
        for (int off = width; off < 32; off <<= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
                q[j] += __shfl_xor_sync(0xffffffffu, q[j], off);
            }
        }
    }
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int parts = width < 32 ? 8 : rows_per_block;       // partial rows to add up (one per warp / per row lane)
    if (width < 32) {
        if (lane < width) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { red[0][wrp * width + lane][j] = s[j]; red[1][wrp * width + lane][j] = q[j]; }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { red[0][threadIdx.x][j] = s[j]; red[1][threadIdx.x][j] = q[j]; }
    }
    __syncthreads();
    if (threadIdx.x < width && cv < a.cvec) {               // row 0 of the partials is this thread's own register total
        for (int l = 1; l < parts; ++l)
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += red[0][l * width + tx][j]; q[j] += red[1][l * width + tx][j]; }
        // one row of per-block partials (no atomics: ~1200 blocks adding into the same few cache lines serialised at
        // the L2 and made this kernel slower than the HBM pass it does)
        float* mine = ws + static_cast<size_t>(blockIdx.x) * cols;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mine[cv * 8 + j] = s[j];
            mine[c + cv * 8 + j] = q[j];
        }
    }
    // Two-level deterministic sum of the block rows: the last block of every group of kBnRedGroup blocks folds the
    // group's rows into one group row, the last group to finish folds the group rows and finalizes.  (One block walking
    // all rows serially was a chain of dependent L2 round trips that outlasted the streaming pass on every map below
    // ~30 MB: 22 us for an 8 MB map.)
    const unsigned ngroups = (gridDim.x + kBnRedGroup - 1) / kBnRedGroup;
    const unsigned grp = blockIdx.x / kBnRedGroup;
    const unsigned grp_size = min(static_cast<unsigned>(kBnRedGroup), gridDim.x - grp * kBnRedGroup);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_flag = atomicAdd(counters + 1 + grp, 1u) == grp_size - 1;
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    double* dred = reinterpret_cast<double*>(&red[0][0][0]);               // 16 KB = 2048 doubles
    float* grow = ws2 + static_cast<size_t>(grp) * cols;
    sum_partial_rows(ws + static_cast<size_t>(grp) * kBnRedGroup * cols, static_cast<int>(grp_size), cols, dred,
                     [&](int col, double t) { grow[col] = static_cast<float>(t); });
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        counters[1 + grp] = 0u;
        s_flag = atomicAdd(counters, 1u) == ngroups - 1;
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    sum_partial_rows(ws2, static_cast<int>(ngroups), cols, dred, [&](int col, double t) {
        if (col < c) {
            coef[col] = static_cast<float>(t * inv_count);
            if (dbeta && col >= a.c_lo) dbeta[col - a.c_lo] = static_cast<float>(t);
        } else {
            const int ch = col - c;
            const double tq = t * static_cast<double>(a.mean_invstd[c + ch]);
            coef[c + ch] = static_cast<float>(tq * inv_count);
            if (dgamma && ch >= a.c_lo) dgamma[ch - a.c_lo] = static_cast<float>(tq);
        }
    });
    if (threadIdx.x == 0) counters[0] = 0u;
}

__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(BnBwdArgs a, const float* __restrict__ coef,
                                                              uint4* __restrict__ out, int zero_insert) {
    const int width = a.cvec < 256 ? a.cvec : 256;
    const int rows_per_block = 256 / width;
    const int tx = threadIdx.x % width, ty = threadIdx.x / width;
    if (ty >= rows_per_block) return;
    const unsigned m = static_cast<unsigned>(a.n) * a.h * a.w;
    const int c = a.cvec * 8;
    const int cv = tx;
    if (cv >= a.cvec) return;
    // draw = k*(g - c1 - (raw - mean)*istd*c2) = k*g + raw*B + D  with  B = -k*c2*istd,  D = -k*c1 - mean*B
    float sc[8], sh[8], kk[8], bb[8], dd[8];
    {
        float mean[8], istd[8], c1[8], c2[8];
        load8f(a.mean_invstd + cv * 8, mean);
        load8f(a.mean_invstd + c + cv * 8, istd);
        load8f(a.scale + cv * 8, kk);
        load8f(coef + cv * 8, c1);
        load8f(coef + c + cv * 8, c2);
        if (a.relu_affine) load8f(a.shift + cv * 8, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = a.relu_affine ? kk[j] : 0.f;
            if (!a.relu_affine) sh[j] = 1.f;                       // no ReLU: gate always open
            bb[j] = -kk[j] * c2[j] * istd[j];
            dd[j] = -kk[j] * c1[j] - mean[j] * bb[j];
        }
    }
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const unsigned step = gridDim.x * rows_per_block;
    unsigned r = blockIdx.x * rows_per_block + ty;
    if (a.up_log2 == 0 && !zero_insert) {
        for (; r + 3 * step < m; r += 4 * step) {                 // four rows per trip (eight independent loads in flight)
            uint4 xr[4], dr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                xr[u] = __ldg(a.raw + static_cast<size_t>(r + u * step) * a.cvec + cv);
                dr[u] = __ldg(a.dy + static_cast<size_t>(r + u * step) * a.cvec + cv);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float x[8], g[8], o[8];
                unpack8b(xr[u], x);
                unpack8b(dr[u], g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float gg = (x[j] * sc[j] + sh[j]) > 0.f ? g[j] : 0.f;
                    o[j] = kk[j] * gg + (x[j] * bb[j] + dd[j]);
                }
                out[static_cast<size_t>(r + u * step) * a.cvec + cv] = pack8b(o);
            }
        }
    }
    for (; r < m; r += step) {
        float rawf[8], g[8], o[8];
        unpack8b(a.raw[static_cast<size_t>(r) * a.cvec + cv], rawf);
        bn_bwd_load_g(a, r, cv, rawf, sc, sh, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = kk[j] * g[j] + (rawf[j] * bb[j] + dd[j]);
        if (!zero_insert) {
            out[static_cast<size_t>(r) * a.cvec + cv] = pack8b(o);
        } else {
            const unsigned ox = r % a.w, t = r / a.w;
            const unsigned oy = t % a.h, im = t / a.h;
            const size_t base = (static_cast<size_t>(im * 2 * a.h + 2 * oy) * (2 * a.w) + 2 * ox) * a.cvec + cv;
            out[base] = pack8b(o);
            out[base + a.cvec] = zero;
            out[base + static_cast<size_t>(2 * a.w) * a.cvec] = zero;
            out[base + static_cast<size_t>(2 * a.w) * a.cvec + a.cvec] = zero;
        }
    }
}

static int bn_bwd_grid(int64_t m, int cvec) {
    const int width = cvec < 256 ? cvec : 256;
    int blocks = ceil_div(m, static_cast<int64_t>(256 / width) * 4);      // one unrolled four-row trip per thread at least
    if (blocks > 148 * 8) blocks = 148 * 8;
    return blocks;
}

int bn_bwd_reduce(const void* dy, const void* raw, const float* mean_invstd, const float* scale, const float* shift,
                  int n, int h, int w, int c, int up_log2, int relu_affine, int c_lo, void* workspace, float* coef,
                  float* dgamma, float* dbeta, cudaStream_t stream) {
    BPB_CHECK(c % 8 == 0 && c <= 2048, "bn_bwd_reduce: C=%d must be a multiple of 8 and <= 2048", c);
    BPB_CHECK(up_log2 >= 0 && up_log2 <= 3, "bn_bwd_reduce: up_log2=%d", up_log2);
    BPB_CHECK(((reinterpret_cast<uintptr_t>(mean_invstd) | reinterpret_cast<uintptr_t>(scale) |
                reinterpret_cast<uintptr_t>(shift) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
              "bn_bwd_reduce: mean_invstd / scale / shift / workspace must be 16-byte aligned");
    BnBwdArgs a{reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(raw), mean_invstd, scale, shift,
                n, h, w, c / 8, up_log2, relu_affine, c_lo};
    const int64_t m = static_cast<int64_t>(n) * h * w;
    const int width = c / 8 < 256 ? c / 8 : 256;
    // every thread should see at least ~4 rows (one unrolled trip), otherwise as many blocks as the row sum supports
    int blocks = ceil_div(m, static_cast<int64_t>(256 / width) * 4);
    if (blocks > kBnRedBlocks) blocks = kBnRedBlocks;
    if (blocks < 1) blocks = 1;
    bn_bwd_reduce_kernel<<<blocks, 256, 0, stream>>>(a, reinterpret_cast<float*>(workspace),
                                                     1.0 / static_cast<double>(m), coef, dgamma, dbeta);
    BPB_LAUNCH_CHECK("bn_bwd_reduce_kernel");
    return 0;
}

int bn_bwd_apply(const void* dy, const void* raw, const float* mean_invstd, const float* scale, const float* shift,
                 const float* coef, int n, int h, int w, int c, int up_log2, int relu_affine, int zero_insert,
                 void* draw, cudaStream_t stream) {
    BPB_CHECK(c % 8 == 0 && c <= 2048, "bn_bwd_apply: C=%d must be a multiple of 8 and <= 2048", c);
    BPB_CHECK(((reinterpret_cast<uintptr_t>(mean_invstd) | reinterpret_cast<uintptr_t>(scale) |
                reinterpret_cast<uintptr_t>(shift) | reinterpret_cast<uintptr_t>(coef)) & 15) == 0,
              "bn_bwd_apply: mean_invstd / scale / shift / coef must be 16-byte aligned");
    BnBwdArgs a{reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(raw), mean_invstd, scale, shift,
                n, h, w, c / 8, up_log2, relu_affine, 0};
    const int64_t m = static_cast<int64_t>(n) * h * w;
    bn_bwd_apply_kernel<<<bn_bwd_grid(m, c / 8), 256, 0, stream>>>(a, coef, reinterpret_cast<uint4*>(draw), zero_insert);
    BPB_LAUNCH_CHECK("bn_bwd_apply_kernel");
    return 0;
}

int64_t bn_bwd_workspace_bytes(int c) {
    return (static_cast<int64_t>(kBnRedBlocks) + kBnRedGroups) * 2 * c * 4 + (kBnRedGroups + 1) * 4 + 64;
}

// ----------------------------------------------------------------------------------------------
// Stem weight gradient (3 -> 64 channels, NCHW fp32 image, PyTorch weight layout [64][3][k][k]):
//   dW[co][ci][r][s] += sum_pix draw[pix][co] * img[n][ci][oy*stride + r - pad][ox*stride + s - pad]
// CTA: 256 threads = 64 output channels x 4 tap groups; 64-pixel chunks staged through shared memory.
// ----------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ img,
                                                         const __nv_bfloat16* __restrict__ draw,
                                                         float* __restrict__ dw, int n, int h, int wd, int stride,
                                                         int ho, int wo) {
    constexpr int kTaps = 3 * KS * KS;
    constexpr int kPer = (kTaps + 3) / 4;                    // taps per thread group
    constexpr int kPix = KS == 3 ? 64 : 32;          // pixels per staged chunk (static shared memory <= 48 KB)
    __shared__ float s_img[kPix][kTaps + 1];
    __shared__ float s_dy[kPix][64];
    const int co = threadIdx.x & 63, grp = threadIdx.x >> 6;
    float acc[kPer];
#pragma unroll
    for (int t = 0; t < kPer; ++t) acc[t] = 0.f;
    const int64_t total = static_cast<int64_t>(n) * ho * wo;
    constexpr int pad = KS / 2;
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * kPix; base < total;
         base += static_cast<int64_t>(gridDim.x) * kPix) {
        __syncthreads();
        for (int i = threadIdx.x; i < kPix * kTaps; i += 256) {
            const int px = i / kTaps, tp = i % kTaps;
            const int64_t pix = base + px;
            float v = 0.f;
            if (pix < total) {
                const int ox = pix % wo, oy = (pix / wo) % ho, im = pix / (static_cast<int64_t>(wo) * ho);
                const int ci = tp / (KS * KS), r = (tp / KS) % KS, sx = tp % KS;
                const int iy = oy * stride + r - pad, ix = ox * stride + sx - pad;
                if (iy >= 0 && iy < h && ix >= 0 && ix < wd)
                    // the forward consumed the image at bf16 precision (stem_conv_kernel)
                    v = __bfloat162float(__float2bfloat16(img[((static_cast<int64_t>(im) * 3 + ci) * h + iy) * wd + ix]));
            }
            s_img[px][tp] = v;
        }
        for (int i = threadIdx.x; i < kPix * 64; i += 256) {
            const int px = i >> 6, cc = i & 63;
            const int64_t pix = base + px;
            s_dy[px][cc] = pix < total ? __bfloat162float(draw[pix * 64 + cc]) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int px = 0; px < kPix; ++px) {
            const float d = s_dy[px][co];
#pragma unroll
            for (int t = 0; t < kPer; ++t) {
                const int tp = grp * kPer + t;
                if (tp < kTaps) acc[t] += d * s_img[px][tp];
            }
        }
    }
#pragma unroll
    for (int t = 0; t < kPer; ++t) {
        const int tp = grp * kPer + t;
        if (tp < kTaps) atomicAdd(dw + co * kTaps + tp, acc[t]);
    }
}

// 3x3 stem, register-tiled: thread = (4 output channels) x (tap group of 7 taps, padded to 8) x (pixel phase of 4).
// The first version (one channel x 7 taps per thread, scalar shared-memory loads) issued 8 LDS per 7 FMA and ran at the
// shared-memory issue rate (890 us at batch 256); here one pixel costs 3 x LDS.128 for 32 FMA.
__global__ void __launch_bounds__(256) stem_wgrad3_kernel(const float* __restrict__ img,
                                                          const __nv_bfloat16* __restrict__ draw, float* __restrict__ dw,
                                                          int n, int h, int wd, int stride, int ho, int wo) {
    constexpr int kTaps = 27, kPix = 64;
    __shared__ __align__(16) float s_dy[kPix][64];            // reused for the cross-phase reduction at the end
    __shared__ float s_img[32][kPix];                         // [slot][px]: tap t -> slot (t / 7) * 8 + t % 7
    __shared__ int s_base[kPix], s_iy[kPix], s_ix[kPix];      // per pixel: image plane offset, top-left input coords
    const int cg = threadIdx.x & 15, tg = (threadIdx.x >> 4) & 3, ph = threadIdx.x >> 6;
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
    for (int i = threadIdx.x; i < 32 * kPix; i += 256) (&s_img[0][0])[i] = 0.f;     // padding slots stay zero
    const int64_t total = static_cast<int64_t>(n) * ho * wo;
    const int hw_out = ho * wo;
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * kPix; base < total;
         base += static_cast<int64_t>(gridDim.x) * kPix) {
        __syncthreads();
        if (threadIdx.x < kPix) {                                 // pixel coordinates once per pixel (no per-tap division)
            const int64_t pix = base + threadIdx.x;
            if (pix < total) {
                const int im = static_cast<int>(pix / hw_out), rem = static_cast<int>(pix - static_cast<int64_t>(im) * hw_out);
                const int oy = rem / wo, ox = rem - oy * wo;
                s_base[threadIdx.x] = im * 3 * h * wd;
                s_iy[threadIdx.x] = oy * stride - 1;
                s_ix[threadIdx.x] = ox * stride - 1;
            } else {
                s_base[threadIdx.x] = -1;
            }
        }
        for (int i = threadIdx.x; i < kPix * 8; i += 256) {       // 8 channels (16 bytes of bf16) per thread
            const int px = i >> 3, c8 = i & 7;
            const int64_t pix = base + px;
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (pix < total) unpack8b(*reinterpret_cast<const uint4*>(draw + pix * 64 + c8 * 8), f);
#pragma unroll
            for (int t = 0; t < 8; ++t) s_dy[px][c8 * 8 + t] = f[t];
        }
        __syncthreads();
        // lanes walk consecutive pixels of one tap: stride-`stride` reads of one image row, conflict-free stores
        for (int i = threadIdx.x; i < kTaps * kPix; i += 256) {
            const int tp = i >> 6, px = i & 63;
            const int ci = tp / 9, r = (tp / 3) % 3, sx = tp % 3;
            float v = 0.f;
            const int b0 = s_base[px];
            if (b0 >= 0) {
                const int iy = s_iy[px] + r, ix = s_ix[px] + sx;
                if (iy >= 0 && iy < h && ix >= 0 && ix < wd)
                    // the forward consumed the image at bf16 precision (stem_conv_kernel)
                    v = __bfloat162float(__float2bfloat16(img[static_cast<size_t>(b0) + (static_cast<size_t>(ci) * h + iy) * wd + ix]));
            }
            s_img[(tp / 7) * 8 + tp % 7][px] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int px = ph; px < kPix; px += 4) {
            const float4 d = *reinterpret_cast<const float4*>(&s_dy[px][cg * 4]);
            const float dv[4] = {d.x, d.y, d.z, d.w};
            float av[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) av[b] = s_img[tg * 8 + b][px];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] += dv[a] * av[b];
        }
    }
    // sum the four pixel phases through shared memory, then one atomic per (channel, tap) and block
    __syncthreads();
    float* red = &s_dy[0][0];                                     // 4096 floats: [phase 0..1][64 co][32 slots] x 2 rounds
    for (int round = 0; round < 2; ++round) {
        if ((ph >> 1) == round) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) red[((ph & 1) * 64 + cg * 4 + a) * 32 + tg * 8 + b] = acc[a][b];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 32; i += 256) {
            const int co = i >> 5, slot = i & 31;
            const int tp = (slot >> 3) * 7 + (slot & 7);
            if ((slot & 7) < 7 && tp < kTaps) atomicAdd(dw + co * kTaps + tp, red[co * 32 + slot] + red[(64 + co) * 32 + slot]);
        }
        __syncthreads();
    }
}

int stem_wgrad(const float* img, const void* draw, float* dw, int n, int h, int wd, int cout, int ksize, int stride,
               cudaStream_t stream) {
    BPB_CHECK(ksize == 3 || ksize == 7, "stem_wgrad: ksize %d unsupported", ksize);
    BPB_CHECK(cout == 64, "stem_wgrad: cout %d unsupported", cout);
    const int pad = ksize / 2;
    const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (wd + 2 * pad - ksize) / stride + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * wo;
    int blocks = ceil_div(total, 64);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (ksize == 3)
        stem_wgrad3_kernel<<<blocks, 256, 0, stream>>>(img, reinterpret_cast<const __nv_bfloat16*>(draw), dw, n, h, wd,
                                                       stride, ho, wo);
    else
        stem_wgrad_kernel<7><<<blocks, 256, 0, stream>>>(img, reinterpret_cast<const __nv_bfloat16*>(draw), dw, n, h, wd,
                                                         stride, ho, wo);
    BPB_LAUNCH_CHECK("stem_wgrad_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Parameter update + operand packing over ONE flat fp32 parameter buffer.
// Segment table (device): every parameter tensor is a segment [off, off+numel).  kind 1 = tensor-core conv weight
// [cout][cin][k][k]: its gradient lives in the flat gradient buffer in OHWI order (what tc_wgrad_kernel writes) and
// two bf16 operand copies are refreshed after the update -- OHWI for the forward conv and [cin][flipped taps][cout]
// for the data gradient (which is then the forward kernel run on dY).
// Adam: torch.optim.Adam semantics (L2 weight decay added to the gradient, bias correction, eps outside the sqrt).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_segment(const ParamSegment* __restrict__ segs, int nseg, int64_t i) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].off <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) adam_pack_kernel(const ParamSegment* __restrict__ segs, int nseg, int64_t total,
                                                        float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, float lr,
                                                        float beta1, float beta2, float eps, float wd, float bc1,
                                                        float bc2, float grad_scale, int do_update) {
    // blocks walk contiguous 1024-element chunks: the binary search runs once per chunk (uniform across the block), the
    // elements of the chunk only step forward from there; index arithmetic in 32 bits (a tensor has < 2^31 elements)
    constexpr int kChunk = 1024;
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * kChunk; base < total;
         base += static_cast<int64_t>(gridDim.x) * kChunk) {
        int sidx = find_segment(segs, nseg, base);
#pragma unroll
        for (int j = 0; j < kChunk / 256; ++j) {
            const int64_t i = base + j * 256 + threadIdx.x;
            if (i >= total) break;
            while (sidx + 1 < nseg && segs[sidx + 1].off <= i) ++sidx;
            const ParamSegment sg = segs[sidx];
            const int64_t li64 = i - sg.off;
            if (li64 >= sg.numel) continue;                        // alignment padding between segments
            const unsigned li = static_cast<unsigned>(li64);
            int64_t gi = i;
            unsigned co = 0, ci = 0, r = 0, sx = 0;
            const unsigned k = sg.k, kk = k * k, cin = sg.cin;
            if (sg.kind == 1) {
                co = li / (cin * kk);
                const unsigned rem = li - co * cin * kk;
                ci = rem / kk;
                const unsigned tap = rem - ci * kk;
                r = tap / k; sx = tap - r * k;
                gi = sg.off + (static_cast<int64_t>(co) * kk + tap) * cin + ci;
            }
            float w = p[i];
            if (do_update && sg.trainable) {
                float gr = g[gi] * grad_scale + wd * w;
                const float mm = beta1 * m[i] + (1.f - beta1) * gr;
                const float vv = beta2 * v[i] + (1.f - beta2) * gr * gr;
                m[i] = mm; v[i] = vv;
                w -= lr * (mm / bc1) / (sqrtf(vv / bc2) + eps);
                p[i] = w;
            }
            if (sg.kind == 1) {
                const __nv_bfloat16 wb = __float2bfloat16(w);
                const uint16_t wbits = *reinterpret_cast<const uint16_t*>(&wb);
                if (sg.wf) sg.wf[(static_cast<int64_t>(co) * kk + r * k + sx) * cin + ci] = wbits;
                if (sg.wd) sg.wd[(static_cast<int64_t>(ci) * kk + (k - 1 - r) * k + (k - 1 - sx)) * sg.cout + co] = wbits;
            }
        }
    }
}

int adam_pack(const void* segs, int nseg, int64_t total, float* p, const float* g, float* m, float* v, float lr,
              float beta1, float beta2, float eps, float wd, int step, float grad_scale, int do_update,
              cudaStream_t stream) {
    BPB_CHECK(nseg >= 1 && total >= 1, "adam_pack: empty parameter set");
    BPB_CHECK(!do_update || (g && m && v && step >= 1), "adam_pack: update needs g/m/v and step >= 1");
    const float bc1 = do_update ? 1.f - powf(beta1, static_cast<float>(step)) : 1.f;
    const float bc2 = do_update ? 1.f - powf(beta2, static_cast<float>(step)) : 1.f;
    int blocks = ceil_div(total, 1024);
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_pack_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const ParamSegment*>(segs), nseg, total, p, g, m, v,
                                                 lr, beta1, beta2, eps, wd, bc1, bc2, grad_scale, do_update);
    BPB_LAUNCH_CHECK("adam_pack_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// out[c][r] = in[r][c] (fp32, zero padding of out's extra rows / columns): operand staging for the tf32 GEMMs of the
// Linear-layer backward (the tensor-core kernel wants both operands K-major)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ in, int64_t ld_in, int rows,
                                                            int cols, float* __restrict__ out, int64_t ld_out,
                                                            int out_rows, int out_cols) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;    // bx over input cols, by over input rows
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = by + j, c = bx + tx;
        tile[j][tx] = (r < rows && c < cols) ? in[static_cast<int64_t>(r) * ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int oc = bx + j, orow = by + tx;               // out[oc][orow]
        if (oc < out_rows && orow < out_cols) out[static_cast<int64_t>(oc) * ld_out + orow] = tile[tx][j];
    }
}

int transpose_pad(const float* in, int64_t ld_in, int rows, int cols, float* out, int64_t ld_out, int out_rows,
                  int out_cols, cudaStream_t stream) {
    BPB_CHECK(out_rows >= cols && out_cols >= rows, "transpose_pad: output smaller than the transposed input");
    dim3 grid(ceil_div(out_rows, 32), ceil_div(out_cols, 32));
    transpose_pad_kernel<<<grid, 256, 0, stream>>>(in, ld_in, rows, cols, out, ld_out, out_rows, out_cols);
    BPB_LAUNCH_CHECK("transpose_pad_kernel");
    return 0;
}

}  // namespace bpb
