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

// scratch layout (floats): [0] running max (float bits)  [1..3] unused
//                          then row norms qs[P][Nq], gs[P][Ng], then visibility bitmasks qmask[Nq], gmask[Ng] (u32)

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

// sqrt.approx (MUFU.SQRT, <= 1 ulp-ish): the tf32 dot product in front of it is accurate to ~1e-3 relative, the IEEE
// software sqrt would only add ~8 dependent instructions per output to the epilogue (the kernel's busiest role)
__device__ __forceinline__ float sqrt_fast(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 32-byte streaming store (STG.EF.256): a lane writes one full sector (half the LSU requests of float4 stores) and the
// evict-first hint keeps the 1.9 GB output stream from pushing the L2-resident gallery tiles out
__device__ __forceinline__ void st_global_v8(float* p, const float (&o)[8]) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]),
                 "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                 : "memory");
}

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
// tcgen05 path.  Persistent CTAs (one per SM); a tile = 128 queries x 256 gallery rows, all P parts.
//   warp 0      TMA producer: per K-chunk of 32 floats one 128x128B query box + one 256x128B gallery box (SW128)
//   warp 1      MMA issuer: 4 x tcgen05.mma kind::tf32 M128 N256 K8 per chunk into one of two 256-column TMEM
//               accumulators (part p+1 accumulates while part p drains)
//   warps 2..17 epilogue: 4 warps per TMEM lane quarter, 64 columns each: sqrt / mask / running numerator in
//               registers, part_dist row segments written straight from registers, combined distance at tile end
// The kernel is bound by the L2 -> SM operand stream ((128+256) rows x D x 4 B per part per tile, ~9.7 GB at
// 4096 x 16384 x 6 x 512 against a ~12 TB/s L2 fabric), not by HBM (2.1 GB) or the tensor pipe; the tile raster keeps
// a group of 8 gallery tiles L2-resident while the query tiles stream past, so HBM sees every operand about once.
// ----------------------------------------------------------------------------------------------
constexpr int kDM = 128;           // queries per tile
constexpr int kDN = 256;           // gallery rows per tile
constexpr int kDStages = 4;
constexpr int kDStageBytes = (kDM + kDN) * 128;
constexpr int kDEpiWarps = 16;
constexpr int kDEpiThreads = kDEpiWarps * 32;
constexpr int kDThreads = 64 + kDEpiThreads;
constexpr int kDMaxParts = 16;     // per-tile gallery vectors of all parts live in shared memory
constexpr int kDGroup = 8;         // gallery tiles per raster group
constexpr size_t kDSmemMax = kDStages * kDStageBytes + 1024 + (2 * kDMaxParts + 1) * kDN * 4;

struct DistParams {
    int nq, ng, p, d, vis_mode, metric, strat;
    const void* qv; const void* gv;
    const float* qs; const float* gs;
    const uint32_t* qmask; const uint32_t* gmask;
    float* dist; float* part_dist; float* scratch;
    int tiles_q, tiles_g, num_tiles;
};

__device__ __forceinline__ void dist_tile_coords(int t, int tiles_q, int tiles_g, int& tq, int& tg) {
    const int group_sz = kDGroup * tiles_q;
    const int grp = t / group_sz;
    const int within = t - grp * group_sz;
    const int g_in_grp = min(kDGroup, tiles_g - grp * kDGroup);
    tq = within / g_in_grp;
    tg = grp * kDGroup + (within - tq * g_in_grp);
}

// visibility bitmask per row (bit p = part p visible); boolean visibility only
__global__ void __launch_bounds__(256) vis_mask_kernel(const uint8_t* __restrict__ v, int rows, int p,
                                                       uint32_t* __restrict__ mask) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    uint32_t m = 0;
    for (int pp = 0; pp < p; ++pp) m |= (v[static_cast<int64_t>(r) * p + pp] ? 1u : 0u) << pp;
    mask[r] = m;
}

template <int kVis, bool kEuclid, bool kMax>
__global__ void __launch_bounds__(kDThreads, 1) part_distance_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                        const __grid_constant__ CUtensorMap tmG,
                                                                        const DistParams a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* s_gs = reinterpret_cast<float*>(smem + kDStages * kDStageBytes);      // [P][256] gallery |g|^2
    float* s_gw = s_gs + a.p * kDN;                                               // [P][256] gallery visibility
    uint32_t* s_gm = reinterpret_cast<uint32_t*>(s_gw + a.p * kDN);               // [256] gallery part bitmask
    __shared__ uint64_t full_bar[kDStages], empty_bar[kDStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_max[kDEpiWarps];

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int kchunks = a.d >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kDEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmG); }
    if (warp == 1) tmem_alloc<512>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_smem, 0);     // warp-uniform
    const uint32_t smem_a = smem_u32(smem);
    const uint32_t full_a = smem_u32(&full_bar[0]), empty_a = smem_u32(&empty_bar[0]);
    const uint32_t accf_a = smem_u32(&acc_full[0]), acce_a = smem_u32(&acc_empty[0]);

    if (warp == 0) {
        // all 32 lanes run the (warp-uniform) control flow, the elected lane issues: keeps TMA / MMA operands in
        // uniform registers (see elect_one_sync in common.cuh)
        const bool leader = elect_one_sync();
        {
            uint32_t s = 0, ph = 0;
            for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
                int tq, tg;
                dist_tile_coords(t, a.tiles_q, a.tiles_g, tq, tg);
                const int i0 = tq * kDM, j0 = tg * kDN;
                for (int pp = 0; pp < a.p; ++pp) {
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait_a(empty_a + s * 8, ph ^ 1);
                        const uint32_t dst = smem_a + s * kDStageBytes;
                        if (leader) {
                            mbar_expect_tx_a(full_a + s * 8, kDStageBytes);
                            tma_load_3d_a(dst, &tmQ, full_a + s * 8, kc * 32, pp, i0);
                            tma_load_3d_a(dst + kDM * 128, &tmG, full_a + s * 8, kc * 32, pp, j0);
                        }
                        if (++s == kDStages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one_sync();
        {
            constexpr uint32_t idesc = make_idesc(true, kDM, kDN);
            const uint64_t desc_hi = make_kmajor_desc(0, 128);
            uint32_t s = 0, ph = 0, acc_it = 0;
            for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
                for (int pp = 0; pp < a.p; ++pp, ++acc_it) {
                    const uint32_t buf = acc_it & 1, use = acc_it >> 1;
                    mbar_wait_a(acce_a + buf * 8, (use & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + buf * kDN;
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait_a(full_a + s * 8, ph);
                        tc_fence_after();
                        const uint32_t q_addr = smem_a + s * kDStageBytes;
                        const uint64_t dq = desc_hi | static_cast<uint64_t>((q_addr & 0x3FFFFu) >> 4);
                        const uint64_t dg = desc_hi | static_cast<uint64_t>(((q_addr + kDM * 128) & 0x3FFFFu) >> 4);
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma<true>(tacc, dq + 2 * k, dg + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                            umma_commit_a(empty_a + s * 8);
                            if (kc == kchunks - 1) umma_commit_a(accf_a + buf * 8);
                        }
                        __syncwarp();
                        if (++s == kDStages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else {
        const int ew = warp - 2;                 // 0..15
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int cg = ew >> 2;                  // which 64-column group of the tile this warp drains
        const int row = q * 32 + lane;
        const int et = ew * 32 + lane;           // 0..511
        float local_max = -1.f;
        uint32_t acc_it = 0;
        // 32-byte vector stores need 32-byte aligned row segments
        const bool vec_ok = (a.ng & 7) == 0 && (reinterpret_cast<uintptr_t>(a.dist) & 31) == 0 &&
                            (a.part_dist == nullptr || (reinterpret_cast<uintptr_t>(a.part_dist) & 31) == 0);
        for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
            int tq, tg;
            dist_tile_coords(t, a.tiles_q, a.tiles_g, tq, tg);
            const int i0 = tq * kDM, j0 = tg * kDN;
            const int i = i0 + row;
            const bool row_ok = i < a.nq;
            const int jc = j0 + cg * 64;           // first gallery column of this thread
            // ---- per-tile gallery-side vectors of all parts ----
            asm volatile("bar.sync 1, %0;" ::"n"(kDEpiThreads) : "memory");       // previous tile done with them
            for (int idx = et; idx < a.p * kDN; idx += kDEpiThreads) {
                const int pp = idx >> 8, jj = idx & (kDN - 1), j = j0 + jj;
                float gsv = 0.f, w = 1.f;
                if (j < a.ng) {
                    if (kEuclid) gsv = a.gs[static_cast<int64_t>(pp) * a.ng + j];
                    if (kVis == VIS_BOOL) w = (a.gmask[j] >> pp) & 1u ? 1.f : 0.f;
                    else if (kVis == VIS_FLOAT) w = reinterpret_cast<const float*>(a.gv)[static_cast<int64_t>(j) * a.p + pp];
                }
                s_gs[idx] = gsv;
                s_gw[idx] = w;
            }
            if (et < kDN) s_gm[et] = (kVis == VIS_BOOL && j0 + et < a.ng) ? a.gmask[j0 + et] : 0xFFFFFFFFu;
            asm volatile("bar.sync 1, %0;" ::"n"(kDEpiThreads) : "memory");
            const uint32_t qm = (kVis == VIS_BOOL && row_ok) ? a.qmask[i] : 0xFFFFFFFFu;

            float num[64];
#pragma unroll
            for (int u = 0; u < 64; ++u) num[u] = kMax ? -1.f : 0.f;
            for (int pp = 0; pp < a.p; ++pp, ++acc_it) {
                const uint32_t buf = acc_it & 1, use = acc_it >> 1;
                float qsv = 0.f, qw = 1.f;
                if (row_ok) {
                    if (kEuclid) qsv = a.qs[static_cast<int64_t>(pp) * a.nq + i];
                    if (kVis == VIS_BOOL) qw = (qm >> pp) & 1u ? 1.f : 0.f;
                    else if (kVis == VIS_FLOAT) qw = reinterpret_cast<const float*>(a.qv)[static_cast<int64_t>(i) * a.p + pp];
                }
                const float* gs_p = s_gs + pp * kDN + cg * 64;
                const float* gw_p = s_gw + pp * kDN + cg * 64;
                // (a part the query does not show is rewritten as a whole row by the fix-up: skip its stores here)
                float* prow = (a.part_dist && row_ok && !(kVis == VIS_BOOL && qw == 0.f))
                                  ? a.part_dist + (static_cast<int64_t>(pp) * a.nq + i) * a.ng + jc : nullptr;
                mbar_wait_a(accf_a + buf * 8, use & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kDN + cg * 64;
#pragma unroll
                for (int cch = 0; cch < 4; ++cch) {
                    uint32_t v[16];
                    tmem_ld16(taddr + cch * 16, v);
                    tmem_ld_wait();
                    if (cch == 3) {                    // accumulator drained: hand the buffer back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_a(acce_a + buf * 8);
                    }
#pragma unroll
                    for (int u8 = 0; u8 < 16; u8 += 8) {
                        float o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int u = cch * 16 + u8 + e;
                            const float dot = __uint_as_float(v[u8 + e]);
                            float dv;
                            if (kEuclid) dv = sqrt_fast(fmaxf(qsv - 2.f * dot + gs_p[u], 0.f));
                            else dv = 1.f - dot;
                            float w = 1.f;
                            if (kVis == VIS_BOOL) w = qw * gw_p[u];
                            else if (kVis == VIS_FLOAT) w = sqrtf(qw * gw_p[u]);
                            const float pd_out = (kVis == VIS_BOOL && w == 0.f) ? -1.f : dv;
                            o[e] = pd_out;
                            if (row_ok && jc + u < a.ng) local_max = fmaxf(local_max, pd_out);
                            if (kMax) num[u] = fmaxf(num[u], kVis == VIS_NONE ? dv : pd_out);
                            else num[u] += dv * w;
                        }
                        if (prow) {
                            const int jb = jc + cch * 16 + u8;
                            if (vec_ok && jb + 8 <= a.ng) {
                                st_global_v8(prow + cch * 16 + u8, o);
                            } else {
#pragma unroll
                                for (int e = 0; e < 8; ++e)
                                    if (jb + e < a.ng) prow[cch * 16 + u8 + e] = o[e];
                            }
                        }
                    }
                }
            }
            // ---- combined distance ----
            if (row_ok) {
                float* drow = a.dist + static_cast<int64_t>(i) * a.ng + jc;
#pragma unroll
                for (int u8 = 0; u8 < 64; u8 += 8) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int u = u8 + e;
                        float r = num[u];
                        if (!kMax) {
                            if (kVis == VIS_NONE) {
                                r = r / static_cast<float>(a.p);
                            } else if (kVis == VIS_BOOL) {
                                const int den = __popc(qm & s_gm[cg * 64 + u]);
                                r = den == 0 ? -1.f : r / static_cast<float>(den);
                            } else {
                                float den = 0.f;
                                for (int pp = 0; pp < a.p; ++pp)
                                    den += sqrtf(reinterpret_cast<const float*>(a.qv)[static_cast<int64_t>(i) * a.p + pp] *
                                                 s_gw[pp * kDN + cg * 64 + u]);
                                r = den == 0.f ? -1.f : r / den;
                            }
                        }
                        o[e] = r;
                    }
                    const int jb = jc + u8;
                    if (vec_ok && jb + 8 <= a.ng) {
                        st_global_v8(drow + u8, o);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (jb + e < a.ng) drow[u8 + e] = o[e];
                    }
                }
            }
        }
        local_max = warp_max(local_max);
        if (lane == 0) s_max[ew] = local_max;
        asm volatile("bar.sync 1, %0;" ::"n"(kDEpiThreads) : "memory");
        if (et == 0) {
            float m = s_max[0];
            for (int u = 1; u < kDEpiWarps; ++u) m = fmaxf(m, s_max[u]);
            atomic_max_float(a.scratch, m);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ----------------------------------------------------------------------------------------------
// Sparse fix-up: every -1 becomes max+1 (distance.py:173-176 / :214-217).  Positions are known from the
// visibility vectors alone, so the big matrices are only written, never re-read.
// Boolean visibility: one pass over (i, j) with the precomputed part bitmasks covers part_dist and dist.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fixup_bool_kernel(float* __restrict__ dist, float* __restrict__ part_dist,
                                                         const uint32_t* __restrict__ qmask,
                                                         const uint32_t* __restrict__ gmask, int nq, int ng, int p,
                                                         const float* __restrict__ scratch) {
    const int i = blockIdx.x;
    const float mv = scratch[0] + 1.f;
    const uint32_t qm = qmask[i];
    const uint32_t all = p >= 32 ? 0xFFFFFFFFu : ((1u << p) - 1u);
    float* drow = dist + static_cast<int64_t>(i) * ng;
    for (int j = threadIdx.x; j < ng; j += blockDim.x) {
        const uint32_t vis = qm & gmask[j];
        if (vis == 0) drow[j] = mv;
        if (part_dist) {
            uint32_t inv = ~vis & all;
            while (inv) {
                const int pp = __ffs(inv) - 1;
                inv &= inv - 1;
                part_dist[(static_cast<int64_t>(pp) * nq + i) * ng + j] = mv;
            }
        }
    }
}

__global__ void __launch_bounds__(256) fixup_dist_float_kernel(float* __restrict__ dist, const float* __restrict__ qv,
                                                               const float* __restrict__ gv, int nq, int ng, int p,
                                                               const float* __restrict__ scratch) {
    const int i = blockIdx.x;
    const float mv = scratch[0] + 1.f;
    float* row = dist + static_cast<int64_t>(i) * ng;
    const float* q = qv + static_cast<int64_t>(i) * p;
    for (int j = threadIdx.x; j < ng; j += blockDim.x) {
        const float* g = gv + static_cast<int64_t>(j) * p;
        float den = 0.f;
        for (int pp = 0; pp < p; ++pp) den += sqrtf(q[pp] * g[pp]);
        if (den == 0.f) row[j] = mv;
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
    // continuous visibility scores always take the weighted mean in the reference, whatever dist_combine_strat says
    // (distance.py:181-219 never looks at it)
    if (vis_mode == VIS_FLOAT) strat = STRAT_MEAN;
    float* qs = scratch + 4;
    float* gs = qs + static_cast<int64_t>(p) * nq;
    uint32_t* qmask = reinterpret_cast<uint32_t*>(gs + static_cast<int64_t>(p) * ng);
    uint32_t* gmask = qmask + nq;
    init_max_kernel<<<1, 1, 0, stream>>>(scratch);
    const bool use_tc = impl != 1 && d % 32 == 0 && p <= kDMaxParts;
    if (metric == METRIC_EUCLIDEAN) {
        sqnorm_kernel<<<ceil_div(static_cast<int64_t>(nq) * p, 8), 256, 0, stream>>>(qf, nq, p, d, use_tc ? 1 : 0, qs);
        sqnorm_kernel<<<ceil_div(static_cast<int64_t>(ng) * p, 8), 256, 0, stream>>>(gf, ng, p, d, use_tc ? 1 : 0, gs);
        BPB_LAUNCH_CHECK("sqnorm_kernel");
    }
    if (vis_mode == VIS_BOOL) {
        vis_mask_kernel<<<ceil_div(nq, 256), 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(qv), nq, p, qmask);
        vis_mask_kernel<<<ceil_div(ng, 256), 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(gv), ng, p, gmask);
        BPB_LAUNCH_CHECK("vis_mask_kernel");
    }
    if (use_tc) {
        CUtensorMap tmQ, tmG;
        uint32_t es[3] = {1, 1, 1};
        {
            uint32_t box[3] = {32, 1, (uint32_t)kDM};
            uint64_t dims[3] = {(uint64_t)d, (uint64_t)p, (uint64_t)nq};
            uint64_t strides[2] = {(uint64_t)d * 4, (uint64_t)p * d * 4};
            if (make_tensor_map(&tmQ, qf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dims, strides, box, es,
                                CU_TENSOR_MAP_SWIZZLE_128B)) return 4;
        }
        {
            uint32_t box[3] = {32, 1, (uint32_t)kDN};
            uint64_t dims[3] = {(uint64_t)d, (uint64_t)p, (uint64_t)ng};
            uint64_t strides[2] = {(uint64_t)d * 4, (uint64_t)p * d * 4};
            if (make_tensor_map(&tmG, gf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dims, strides, box, es,
                                CU_TENSOR_MAP_SWIZZLE_128B)) return 4;
        }
        const int tiles_q = ceil_div(nq, kDM), tiles_g = ceil_div(ng, kDN);
        DistParams a{nq, ng, p, d, vis_mode, metric, strat, qv, gv, qs, gs, qmask, gmask, dist, part_dist, scratch,
                     tiles_q, tiles_g, tiles_q * tiles_g};
        const int num_sms = device_sm_count();
    if (num_sms <= 0) return 2;
        const size_t smem = kDStages * kDStageBytes + 1024 + (2 * p + 1) * kDN * 4;
        const int grid = a.num_tiles < num_sms ? a.num_tiles : num_sms;
        const bool eu = metric == METRIC_EUCLIDEAN, mx = strat == STRAT_MAX;
#define BPB_DIST_LAUNCH(VIS, EU, MX)                                                                              \
    do {                                                                                                          \
        auto kern = part_distance_tc_kernel<VIS, EU, MX>;                                                         \
        BPB_SET_MAX_SMEM(kern, kDSmemMax);                                                                        \
        kern<<<grid, kDThreads, smem, stream>>>(tmQ, tmG, a);                                                     \
    } while (0)
#define BPB_DIST_VIS(VIS)                                                                                         \
    do {                                                                                                          \
        if (eu && !mx) BPB_DIST_LAUNCH(VIS, true, false);                                                         \
        else if (eu && mx) BPB_DIST_LAUNCH(VIS, true, true);                                                      \
        else if (!eu && !mx) BPB_DIST_LAUNCH(VIS, false, false);                                                  \
        else BPB_DIST_LAUNCH(VIS, false, true);                                                                   \
    } while (0)
        if (vis_mode == VIS_NONE) BPB_DIST_VIS(VIS_NONE);
        else if (vis_mode == VIS_BOOL) BPB_DIST_VIS(VIS_BOOL);
        else BPB_DIST_VIS(VIS_FLOAT);
#undef BPB_DIST_VIS
#undef BPB_DIST_LAUNCH
        BPB_LAUNCH_CHECK("part_distance_tc_kernel");
    } else {
        dim3 grid(ceil_div(ng, kST), ceil_div(nq, kST));
        part_distance_simt_kernel<<<grid, 256, 0, stream>>>(qf, gf, qv, gv, vis_mode, nq, ng, p, d, metric, strat, qs,
                                                            gs, dist, part_dist, scratch);
        BPB_LAUNCH_CHECK("part_distance_simt_kernel");
    }
    if (vis_mode == VIS_BOOL) {
        fixup_bool_kernel<<<nq, 256, 0, stream>>>(dist, part_dist, qmask, gmask, nq, ng, p, scratch);
        BPB_LAUNCH_CHECK("fixup_bool_kernel");
    } else if (vis_mode == VIS_FLOAT) {
        fixup_dist_float_kernel<<<nq, 256, 0, stream>>>(dist, reinterpret_cast<const float*>(qv),
                                                        reinterpret_cast<const float*>(gv), nq, ng, p, scratch);
        BPB_LAUNCH_CHECK("fixup_dist_float_kernel");
    }
    return 0;
}

}  // namespace bpb
