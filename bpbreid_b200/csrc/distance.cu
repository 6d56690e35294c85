// Query x gallery part-based distance (evaluation).
//   part_dist[p,i,j] = sqrt(relu(|q_ip|^2 - 2 q_ip.g_jp + |g_jp|^2))     (or 1 - q.g for cosine)
//   dist[i,j]        = visibility-masked mean (or max) over parts, -1 -> max(part_dist)+1
// Reference semantics: torchreid/metrics/distance.py:87-247 and torchreid/utils/tensortools.py:1-19.
// The reference loops over gallery chunks of 500 with a D2H copy each; here one launch covers the whole
// matrix: P tf32 tcgen05 GEMMs per 128x128 output tile with the sqrt / mask / mean epilogue fused, outputs
// written once with full 128-byte lines, and a sparse fix-up pass that rewrites only the sentinel entries.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

enum { VIS_NONE = 0, VIS_BOOL = 1, VIS_FLOAT = 2 };
enum { METRIC_EUCLIDEAN = 0, METRIC_COSINE = 1 };
enum { STRAT_MEAN = 0, STRAT_MAX = 1 };

// scratch layout (floats): [0] running max (float bits)  [1..] unused
//                          then row norms: qs[P][Nq], gs[P][Ng]

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    int* a = reinterpret_cast<int*>(addr);
    int old = *a;
    while (__int_as_float(old) < v) {
        const int assumed = old;
        old = atomicCAS(a, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}

// |x|^2 per (row, part): x [rows, P, D] -> out [P, rows]
// tf32 != 0: square the values the tensor core will actually see (tcgen05 kind::tf32 ignores the low 13 mantissa
// bits), so that |q|^2 - 2 q.g + |g|^2 is the squared distance between the *truncated* vectors: exact duplicates give
// ~0 again and the error becomes relative to the distance instead of to the norms.
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ x, int rows, int p, int d, int tf32,
                                                     float* __restrict__ out) {
    const int64_t w = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= static_cast<int64_t>(rows) * p) return;
    const int r = w / p, pp = w % p;
    const float* v = x + (static_cast<int64_t>(r) * p + pp) * d;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) {
        const float t = tf32 ? __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u) : v[i];
        s += t * t;
    }
    s = warp_sum(s);
    if (lane == 0) out[static_cast<int64_t>(pp) * rows + r] = s;
}

__global__ void init_max_kernel(float* scratch) { scratch[0] = -1.f; }

// ----------------------------------------------------------------------------------------------
// SIMT path (any D; also the on-device cross-check of the tensor-core path)
// ----------------------------------------------------------------------------------------------
constexpr int kST = 64, kSK = 16;

__global__ void __launch_bounds__(256) part_distance_simt_kernel(
    const float* __restrict__ qf, const float* __restrict__ gf, const void* __restrict__ qv,
    const void* __restrict__ gv, int vis_mode, int nq, int ng, int p, int d, int metric, int strat,
    const float* __restrict__ qs, const float* __restrict__ gs, float* __restrict__ dist,
    float* __restrict__ part_dist, float* __restrict__ scratch) {
    __shared__ float sa[kSK][kST + 4], sb[kSK][kST + 4];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int i0 = blockIdx.y * kST, j0 = blockIdx.x * kST;
    float num[4][4], den[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { num[a][b] = strat == STRAT_MAX ? -1.f : 0.f; den[a][b] = 0.f; }
    float local_max = -1.f;
    for (int pp = 0; pp < p; ++pp) {
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        for (int k0 = 0; k0 < d; k0 += kSK) {
            for (int t = threadIdx.x; t < kST * kSK; t += 256) {
                const int r = t / kSK, kk = t % kSK;
                const int qi = i0 + r, gj = j0 + r;
                sa[kk][r] = (qi < nq && k0 + kk < d) ? qf[(static_cast<int64_t>(qi) * p + pp) * d + k0 + kk] : 0.f;
                sb[kk][r] = (gj < ng && k0 + kk < d) ? gf[(static_cast<int64_t>(gj) * p + pp) * d + k0 + kk] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kSK; ++kk) {
                float av[4], bv[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) av[a] = sa[kk][ty * 4 + a];
#pragma unroll
                for (int b = 0; b < 4; ++b) bv[b] = sb[kk][tx * 4 + b];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = i0 + ty * 4 + a;
            if (i >= nq) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = j0 + tx * 4 + b;
                if (j >= ng) continue;
                float dv;
                if (metric == METRIC_EUCLIDEAN) {
                    const float d2 = qs[static_cast<int64_t>(pp) * nq + i] - 2.f * acc[a][b] +
                                     gs[static_cast<int64_t>(pp) * ng + j];
                    dv = sqrtf(fmaxf(d2, 0.f));
                } else {
                    dv = 1.f - acc[a][b];
                }
                float w = 1.f;
                if (vis_mode == VIS_BOOL) {
                    w = (reinterpret_cast<const uint8_t*>(qv)[static_cast<int64_t>(i) * p + pp] &&
                         reinterpret_cast<const uint8_t*>(gv)[static_cast<int64_t>(j) * p + pp]) ? 1.f : 0.f;
                } else if (vis_mode == VIS_FLOAT) {
                    w = sqrtf(reinterpret_cast<const float*>(qv)[static_cast<int64_t>(i) * p + pp] *
                              reinterpret_cast<const float*>(gv)[static_cast<int64_t>(j) * p + pp]);
                }
                const float pd_out = (vis_mode == VIS_BOOL && w == 0.f) ? -1.f : dv;
                if (part_dist) part_dist[(static_cast<int64_t>(pp) * nq + i) * ng + j] = pd_out;
                local_max = fmaxf(local_max, pd_out);
                if (strat == STRAT_MAX) {
                    num[a][b] = fmaxf(num[a][b], vis_mode == VIS_NONE ? dv : pd_out);
                } else {
                    num[a][b] += dv * w;
                    den[a][b] += w;
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i >= nq) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + tx * 4 + b;
            if (j >= ng) continue;
            float o;
            if (strat == STRAT_MAX) o = num[a][b];
            else if (vis_mode == VIS_NONE) o = num[a][b] / static_cast<float>(p);
            else o = den[a][b] == 0.f ? -1.f : num[a][b] / den[a][b];
            dist[static_cast<int64_t>(i) * ng + j] = o;
        }
    }
    // block max -> global
    __shared__ float rmax[256];
    rmax[threadIdx.x] = local_max;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) rmax[threadIdx.x] = fmaxf(rmax[threadIdx.x], rmax[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) atomic_max_float(scratch, rmax[0]);
}

// ----------------------------------------------------------------------------------------------
// tcgen05 path: CTA = 128 queries x 128 gallery rows, loop over parts; per part a tf32 K-loop (D/32 TMA stages)
// accumulates q.g^T into one of two TMEM buffers while 8 epilogue warps drain the other one.
// warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue (two warps per TMEM lane quarter, 64 columns each)
// ----------------------------------------------------------------------------------------------
constexpr int kDT = 128;           // tile edge (queries and gallery)
constexpr int kDThreads = 320;
constexpr int kDStages = 4;
constexpr int kDStageBytes = 2 * kDT * 128;

struct DistParams {
    int nq, ng, p, d, vis_mode, metric, strat;
    const void* qv; const void* gv;
    const float* qs; const float* gs;
    float* dist; float* part_dist; float* scratch;
};

__global__ void __launch_bounds__(kDThreads) part_distance_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmG,
                                                                     const DistParams a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[kDStages], empty_bar[kDStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_gs[2][kDT];          // |g|^2 of the current part (double buffered by part parity)
    __shared__ float s_gvf[2][kDT];         // gallery visibility (as float) of the current part
    __shared__ float s_max[8];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = blockIdx.y * kDT, j0 = blockIdx.x * kDT;
    const int kchunks = a.d / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 256); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmG); }
    if (warp == 1) tmem_alloc<256>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int pp = 0; pp < a.p; ++pp) {
                for (int kc = 0; kc < kchunks; ++kc, ++it) {
                    const int s = it % kDStages;
                    const uint32_t round = it / kDStages;
                    mbar_wait(&empty_bar[s], (round & 1) ^ 1);
                    uint8_t* dst = smem + s * kDStageBytes;
                    mbar_expect_tx(&full_bar[s], kDStageBytes);
                    tma_load_3d(dst, &tmQ, &full_bar[s], kc * 32, pp, i0);
                    tma_load_3d(dst + kDT * 128, &tmG, &full_bar[s], kc * 32, pp, j0);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(true, kDT, kDT);
        int it = 0;
        for (int pp = 0; pp < a.p; ++pp) {
            const int buf = pp & 1;
            const uint32_t use = pp >> 1;              // how many times this buffer has been used before
            mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
            tc_fence_after();
            for (int kc = 0; kc < kchunks; ++kc, ++it) {
                const int s = it % kDStages;
                const uint32_t round = it / kDStages;
                mbar_wait(&full_bar[s], round & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t q_addr = smem_u32(smem + s * kDStageBytes);
                    const uint64_t dq = make_kmajor_desc(q_addr, 128);
                    const uint64_t dg = make_kmajor_desc(q_addr + kDT * 128, 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma<true>(tmem_base + buf * kDT, dq + 2 * k, dg + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                    if (kc == kchunks - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        const int ew = warp - 2;                 // 0..7
        const int q = warp & 3;                  // TMEM lane quarter
        const int half = ew >> 2;                // which 64-column half of the tile this warp drains
        const int row = q * 32 + lane;
        const int i = i0 + row;
        const bool row_ok = i < a.nq;
        const int et = ew * 32 + lane;           // 0..255
        float num[64];
#pragma unroll
        for (int t = 0; t < 64; ++t) num[t] = a.strat == STRAT_MAX ? -1.f : 0.f;
        float local_max = -1.f;
        for (int pp = 0; pp < a.p; ++pp) {
            const int buf = pp & 1;
            const uint32_t use = pp >> 1;
            // stage per-part gallery-side vectors (256 epilogue threads, 128 entries each)
            if (et < kDT) {
                const int j = j0 + et;
                s_gs[buf][et] = (j < a.ng && a.metric == METRIC_EUCLIDEAN) ? a.gs[static_cast<int64_t>(pp) * a.ng + j] : 0.f;
            } else {
                const int jj = et - kDT, j = j0 + jj;
                float w = 1.f;
                if (j < a.ng) {
                    if (a.vis_mode == VIS_BOOL) w = reinterpret_cast<const uint8_t*>(a.gv)[static_cast<int64_t>(j) * a.p + pp] ? 1.f : 0.f;
                    else if (a.vis_mode == VIS_FLOAT) w = reinterpret_cast<const float*>(a.gv)[static_cast<int64_t>(j) * a.p + pp];
                }
                s_gvf[buf][jj] = w;
            }
            float qsv = 0.f, qw = 1.f;
            if (row_ok) {
                if (a.metric == METRIC_EUCLIDEAN) qsv = a.qs[static_cast<int64_t>(pp) * a.nq + i];
                if (a.vis_mode == VIS_BOOL) qw = reinterpret_cast<const uint8_t*>(a.qv)[static_cast<int64_t>(i) * a.p + pp] ? 1.f : 0.f;
                else if (a.vis_mode == VIS_FLOAT) qw = reinterpret_cast<const float*>(a.qv)[static_cast<int64_t>(i) * a.p + pp];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kDT + half * 64;
#pragma unroll
            for (int cch = 0; cch < 2; ++cch) {
                uint32_t v[32];
                tmem_ld32(taddr + cch * 32, v);
                tmem_ld_wait();
                const int jb = half * 64 + cch * 32;
                float o[32];
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const float dot = __uint_as_float(v[t]);
                    float dv;
                    if (a.metric == METRIC_EUCLIDEAN) dv = sqrtf(fmaxf(qsv - 2.f * dot + s_gs[buf][jb + t], 0.f));
                    else dv = 1.f - dot;
                    float w = 1.f;
                    if (a.vis_mode == VIS_BOOL) w = qw * s_gvf[buf][jb + t];
                    else if (a.vis_mode == VIS_FLOAT) w = sqrtf(qw * s_gvf[buf][jb + t]);
                    const float pd_out = (a.vis_mode == VIS_BOOL && w == 0.f) ? -1.f : dv;
                    o[t] = pd_out;
                    if (j0 + jb + t < a.ng) local_max = fmaxf(local_max, pd_out);
                    if (a.strat == STRAT_MAX) num[cch * 32 + t] = fmaxf(num[cch * 32 + t], a.vis_mode == VIS_NONE ? dv : pd_out);
                    else num[cch * 32 + t] += dv * w;
                }
                if (a.part_dist && row_ok) {
                    float* dst = a.part_dist + (static_cast<int64_t>(pp) * a.nq + i) * a.ng + j0 + jb;
                    if (j0 + jb + 32 <= a.ng && (a.ng & 3) == 0) {
#pragma unroll
                        for (int t = 0; t < 32; t += 4)
                            *reinterpret_cast<float4*>(dst + t) = make_float4(o[t], o[t + 1], o[t + 2], o[t + 3]);
                    } else {
#pragma unroll
                        for (int t = 0; t < 32; ++t)
                            if (j0 + jb + t < a.ng) dst[t] = o[t];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
        if (!row_ok) local_max = -1.f;
        // combined distance
        if (row_ok) {
#pragma unroll
            for (int cch = 0; cch < 2; ++cch) {
                const int jb = half * 64 + cch * 32;
                float o[32];
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const int j = j0 + jb + t;
                    float r = num[cch * 32 + t];
                    if (a.strat != STRAT_MAX) {
                        if (a.vis_mode == VIS_NONE) {
                            r = r / static_cast<float>(a.p);
                        } else if (j < a.ng) {
                            float den = 0.f;
                            for (int pp = 0; pp < a.p; ++pp) {
                                if (a.vis_mode == VIS_BOOL) {
                                    den += (reinterpret_cast<const uint8_t*>(a.qv)[static_cast<int64_t>(i) * a.p + pp] &&
                                            reinterpret_cast<const uint8_t*>(a.gv)[static_cast<int64_t>(j) * a.p + pp]) ? 1.f : 0.f;
                                } else {
                                    den += sqrtf(reinterpret_cast<const float*>(a.qv)[static_cast<int64_t>(i) * a.p + pp] *
                                                 reinterpret_cast<const float*>(a.gv)[static_cast<int64_t>(j) * a.p + pp]);
                                }
                            }
                            r = den == 0.f ? -1.f : r / den;
                        }
                    }
                    o[t] = r;
                }
                float* dst = a.dist + static_cast<int64_t>(i) * a.ng + j0 + jb;
                if (j0 + jb + 32 <= a.ng && (a.ng & 3) == 0) {
#pragma unroll
                    for (int t = 0; t < 32; t += 4)
                        *reinterpret_cast<float4*>(dst + t) = make_float4(o[t], o[t + 1], o[t + 2], o[t + 3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 32; ++t)
                        if (j0 + jb + t < a.ng) dst[t] = o[t];
                }
            }
        }
        local_max = warp_max(local_max);
        if (lane == 0) s_max[ew] = local_max;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
            float m = s_max[0];
            for (int t = 1; t < 8; ++t) m = fmaxf(m, s_max[t]);
            atomic_max_float(a.scratch, m);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ----------------------------------------------------------------------------------------------
// Sparse fix-up: every -1 becomes max+1 (distance.py:173-176 / :214-217).  Positions are known from the
// visibility vectors alone, so the big matrices are only written, never re-read.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fixup_part_kernel(float* __restrict__ part_dist, const uint8_t* __restrict__ qv,
                                                         const uint8_t* __restrict__ gv, int nq, int ng, int p,
                                                         const float* __restrict__ scratch) {
    const int pp = blockIdx.y, i = blockIdx.x;
    const float mv = scratch[0] + 1.f;
    float* row = part_dist + (static_cast<int64_t>(pp) * nq + i) * ng;
    const bool qvis = qv[static_cast<int64_t>(i) * p + pp] != 0;
    for (int j = threadIdx.x; j < ng; j += blockDim.x)
        if (!qvis || gv[static_cast<int64_t>(j) * p + pp] == 0) row[j] = mv;
}

__global__ void __launch_bounds__(256) fixup_dist_kernel(float* __restrict__ dist, const void* __restrict__ qv,
                                                         const void* __restrict__ gv, int vis_mode, int nq, int ng,
                                                         int p, const float* __restrict__ scratch) {
    const int i = blockIdx.x;
    const float mv = scratch[0] + 1.f;
    float* row = dist + static_cast<int64_t>(i) * ng;
    if (vis_mode == VIS_BOOL) {
        const uint8_t* q = reinterpret_cast<const uint8_t*>(qv) + static_cast<int64_t>(i) * p;
        unsigned qm = 0;
        for (int pp = 0; pp < p; ++pp) qm |= (q[pp] ? 1u : 0u) << pp;
        for (int j = threadIdx.x; j < ng; j += blockDim.x) {
            const uint8_t* g = reinterpret_cast<const uint8_t*>(gv) + static_cast<int64_t>(j) * p;
            unsigned gm = 0;
            for (int pp = 0; pp < p; ++pp) gm |= (g[pp] ? 1u : 0u) << pp;
            if ((qm & gm) == 0) row[j] = mv;
        }
    } else {
        const float* q = reinterpret_cast<const float*>(qv) + static_cast<int64_t>(i) * p;
        for (int j = threadIdx.x; j < ng; j += blockDim.x) {
            const float* g = reinterpret_cast<const float*>(gv) + static_cast<int64_t>(j) * p;
            float den = 0.f;
            for (int pp = 0; pp < p; ++pp) den += sqrtf(q[pp] * g[pp]);
            if (den == 0.f) row[j] = mv;
        }
    }
}

int part_distance(const float* qf, const float* gf, const void* qv, const void* gv, int vis_mode, int nq, int ng,
                  int p, int d, int metric, int strat, int impl, float* dist, float* part_dist, float* scratch,
                  cudaStream_t stream) {
    BPB_CHECK(vis_mode >= 0 && vis_mode <= 2, "part_distance: vis_mode %d", vis_mode);
    BPB_CHECK(vis_mode == VIS_NONE || (qv && gv), "part_distance: visibility pointers missing");
    BPB_CHECK(metric == METRIC_EUCLIDEAN || metric == METRIC_COSINE, "part_distance: unknown metric %d", metric);
    BPB_CHECK(strat == STRAT_MEAN || strat == STRAT_MAX, "part_distance: unknown strategy %d", strat);
    BPB_CHECK(p >= 1 && p <= 32, "part_distance: P=%d out of range", p);
    BPB_CHECK(!(vis_mode == VIS_FLOAT && strat == STRAT_MAX), "part_distance: 'max' needs boolean visibility");
    float* qs = scratch + 4;
    float* gs = qs + static_cast<int64_t>(p) * nq;
    init_max_kernel<<<1, 1, 0, stream>>>(scratch);
    const bool use_tc = impl != 1 && d % 32 == 0;
    if (metric == METRIC_EUCLIDEAN) {
        sqnorm_kernel<<<ceil_div(static_cast<int64_t>(nq) * p, 8), 256, 0, stream>>>(qf, nq, p, d, use_tc ? 1 : 0, qs);
        sqnorm_kernel<<<ceil_div(static_cast<int64_t>(ng) * p, 8), 256, 0, stream>>>(gf, ng, p, d, use_tc ? 1 : 0, gs);
        BPB_LAUNCH_CHECK("sqnorm_kernel");
    }
    if (use_tc) {
        CUtensorMap tmQ, tmG;
        uint32_t box[3] = {32, 1, (uint32_t)kDT};
        uint32_t es[3] = {1, 1, 1};
        {
            uint64_t dims[3] = {(uint64_t)d, (uint64_t)p, (uint64_t)nq};
            uint64_t strides[2] = {(uint64_t)d * 4, (uint64_t)p * d * 4};
            if (make_tensor_map(&tmQ, qf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dims, strides, box, es,
                                CU_TENSOR_MAP_SWIZZLE_128B)) return 4;
        }
        {
            uint64_t dims[3] = {(uint64_t)d, (uint64_t)p, (uint64_t)ng};
            uint64_t strides[2] = {(uint64_t)d * 4, (uint64_t)p * d * 4};
            if (make_tensor_map(&tmG, gf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dims, strides, box, es,
                                CU_TENSOR_MAP_SWIZZLE_128B)) return 4;
        }
        DistParams a{nq, ng, p, d, vis_mode, metric, strat, qv, gv, qs, gs, dist, part_dist, scratch};
        static bool configured = false;
        const size_t smem = kDStages * kDStageBytes + 1024;
        if (!configured) {
            BPB_CUDA(cudaFuncSetAttribute(part_distance_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
            configured = true;
        }
        dim3 grid(ceil_div(ng, kDT), ceil_div(nq, kDT));
        part_distance_tc_kernel<<<grid, kDThreads, smem, stream>>>(tmQ, tmG, a);
        BPB_LAUNCH_CHECK("part_distance_tc_kernel");
    } else {
        dim3 grid(ceil_div(ng, kST), ceil_div(nq, kST));
        part_distance_simt_kernel<<<grid, 256, 0, stream>>>(qf, gf, qv, gv, vis_mode, nq, ng, p, d, metric, strat, qs,
                                                            gs, dist, part_dist, scratch);
        BPB_LAUNCH_CHECK("part_distance_simt_kernel");
    }
    if (vis_mode != VIS_NONE) {
        if (vis_mode == VIS_BOOL && part_dist) {
            dim3 g2(nq, p);
            fixup_part_kernel<<<g2, 256, 0, stream>>>(part_dist, reinterpret_cast<const uint8_t*>(qv),
                                                      reinterpret_cast<const uint8_t*>(gv), nq, ng, p, scratch);
            BPB_LAUNCH_CHECK("fixup_part_kernel");
        }
        fixup_dist_kernel<<<nq, 256, 0, stream>>>(dist, qv, gv, vis_mode, nq, ng, p, scratch);
        BPB_LAUNCH_CHECK("fixup_dist_kernel");
    }
    return 0;
}

}  // namespace bpb
