// CMC / mAP ranking (market1501 protocol) without materialising the argsort, plus a stable row argsort.
// Reference semantics: torchreid/metrics/rank.py:97-159 (numpy, fp64) and its native twin
// torchreid/metrics/rank_cylib/rank_cy.pyx:154-249 (Cython, fp32 accumulators).
//
// Key observation: cmc/AP only need the *ranks of the matching gallery entries*.  For a query q the position
// of match m in the stable ascending order of the kept gallery is
//     rank_m = 1 + #{ kept j : d_j < d_m  or  (d_j == d_m and j < m) }
// so one streaming pass over the distance row (HBM-bound, integer-exact) replaces the O(Ng log Ng) sort:
// bit-identical to np.argsort(kind='stable') based evaluation.
#include "common.cuh"
#include "kernels.h"

namespace bpb {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

constexpr int kMaxMatches = 2048;
constexpr int kMatchBatch = 16;
constexpr int kMaxJunk = 512;     // same pid AND same camera entries per query (removed from the ranking, rank.py:124-127)

__global__ void __launch_bounds__(256) rank_query_kernel(const float* __restrict__ dist,
                                                         const int64_t* __restrict__ q_pids,
                                                         const int64_t* __restrict__ g_pids,
                                                         const int64_t* __restrict__ q_camids,
                                                         const int64_t* __restrict__ g_camids, int nq, int ng,
                                                         double* __restrict__ ap64, float* __restrict__ ap32,
                                                         int* __restrict__ first_rank, int* __restrict__ overflow) {
    __shared__ int m_idx[kMaxMatches];
    __shared__ float m_d[kMaxMatches];
    __shared__ int m_rank[kMaxMatches];
    __shared__ int m_count, j_count;
    __shared__ int cnt[kMatchBatch];
    __shared__ int junk_idx[kMaxJunk];
    __shared__ float junk_d[kMaxJunk];
    const int q = blockIdx.x;
    const int64_t qp = q_pids[q], qc = q_camids[q];
    const float* row = dist + static_cast<int64_t>(q) * ng;
    if (threadIdx.x == 0) { m_count = 0; j_count = 0; }
    __syncthreads();
    // 1. matches = same pid, different camera (kept & relevant)
    //    junk = same pid, same camera (dropped).  The gallery ids are read ONCE: the counting pass below runs over the
    //    whole row and the few junk entries are subtracted afterwards.
    for (int j = threadIdx.x; j < ng; j += blockDim.x) {
        if (g_pids[j] == qp) {
            if (g_camids[j] != qc) {
                const int slot = atomicAdd(&m_count, 1);
                if (slot < kMaxMatches) { m_idx[slot] = j; m_d[slot] = row[j]; }
            } else {
                const int slot = atomicAdd(&j_count, 1);
                if (slot < kMaxJunk) { junk_idx[slot] = j; junk_d[slot] = row[j]; }
            }
        }
    }
    __syncthreads();
    const int nm = m_count;
    const int nj = j_count;
    const bool junk_listed = nj <= kMaxJunk;              // else: fall back to re-reading the ids in the counting pass
    if (nm == 0 || nm > kMaxMatches) {
        if (threadIdx.x == 0) {
            first_rank[q] = -1; ap64[q] = 0.0; ap32[q] = 0.f;
            if (nm > kMaxMatches) atomicExch(overflow, 1);
        }
        return;
    }
    // 2. rank of every match: count kept gallery entries that precede it in the stable order
    for (int b0 = 0; b0 < nm; b0 += kMatchBatch) {
        const int nb = min(kMatchBatch, nm - b0);
        if (threadIdx.x < kMatchBatch) cnt[threadIdx.x] = 0;
        __syncthreads();
        float md[kMatchBatch];
        int mi[kMatchBatch], local[kMatchBatch];
#pragma unroll
        for (int t = 0; t < kMatchBatch; ++t) {
            md[t] = t < nb ? m_d[b0 + t] : 0.f;
            mi[t] = t < nb ? m_idx[b0 + t] : -1;
            local[t] = 0;
        }
        if (junk_listed) {
#pragma unroll 4
            for (int j = threadIdx.x; j < ng; j += blockDim.x) {
                const float dj = row[j];
#pragma unroll
                for (int t = 0; t < kMatchBatch; ++t) local[t] += (dj < md[t] || (dj == md[t] && j < mi[t])) ? 1 : 0;
            }
            // take the junk entries back out (each was counted once above)
            for (int u = threadIdx.x; u < nj; u += blockDim.x) {
                const float dj = junk_d[u];
                const int j = junk_idx[u];
#pragma unroll
                for (int t = 0; t < kMatchBatch; ++t) local[t] -= (dj < md[t] || (dj == md[t] && j < mi[t])) ? 1 : 0;
            }
        } else {
            for (int j = threadIdx.x; j < ng; j += blockDim.x) {
                const bool kept = !(g_pids[j] == qp && g_camids[j] == qc);
                if (!kept) continue;
                const float dj = row[j];
#pragma unroll
                for (int t = 0; t < kMatchBatch; ++t) local[t] += (dj < md[t] || (dj == md[t] && j < mi[t])) ? 1 : 0;
            }
        }
#pragma unroll
        for (int t = 0; t < kMatchBatch; ++t) {
            int v = local[t];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && t < nb) atomicAdd(&cnt[t], v);
        }
        __syncthreads();
        if (threadIdx.x < nb) m_rank[b0 + threadIdx.x] = cnt[threadIdx.x] + 1;     // 1-based position
        __syncthreads();
    }
    // 3. AP = (1/nm) * sum over matches in rank order of (k / rank_k), k = 1-based index among matches
    //    (ranks are distinct, so k_m = 1 + #{m' : rank_m' < rank_m})
    __shared__ double red64[256];
    __shared__ int red_min[256];
    double acc = 0.0;
    int mn = 0x7fffffff;
    for (int m = threadIdx.x; m < nm; m += blockDim.x) {
        const int r = m_rank[m];
        int k = 1;
        for (int o = 0; o < nm; ++o) k += m_rank[o] < r ? 1 : 0;
        acc += static_cast<double>(k) / static_cast<double>(r);
        mn = min(mn, r);
        reinterpret_cast<int*>(m_d)[k - 1] = m;    // reuse m_d: match holding the k-th smallest rank
    }
    red64[threadIdx.x] = acc; red_min[threadIdx.x] = mn;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            red64[threadIdx.x] += red64[threadIdx.x + s];
            red_min[threadIdx.x] = min(red_min[threadIdx.x], red_min[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ap64[q] = red64[0] / static_cast<double>(nm);
        first_rank[q] = red_min[0] - 1;      // 0-based rank of the first correct match
        // Cython emulation (rank_cy.pyx:219-225): float accumulator, terms visited in rank order
        // (serial; match lists are short)
        float s32 = 0.f, nrel = 0.f;
        for (int k = 1; k <= nm; ++k) {
            const int mm = reinterpret_cast<const int*>(m_d)[k - 1];
            const float cum = static_cast<float>(k);           // tmp_cmc[g] (float cumsum of 0/1)
            s32 = static_cast<float>(static_cast<double>(s32) +
                                     (static_cast<double>(cum) / (static_cast<double>(m_rank[mm] - 1) + 1.0)) * 1.0);
            nrel += 1.f;
        }
        ap32[q] = s32 / nrel;
    }
}

__global__ void __launch_bounds__(256) rank_finalize_kernel(const double* ap64, const float* ap32,
                                                            const int* first_rank, int nq, int max_rank, float* cmc,
                                                            double* map_out, int64_t* num_valid, float* map32_out) {
    __shared__ int hist[1024];
    __shared__ int nvalid_s;
    for (int r = threadIdx.x; r < max_rank; r += blockDim.x) hist[r] = 0;
    if (threadIdx.x == 0) nvalid_s = 0;
    __syncthreads();
    int local_valid = 0;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        const int fr = first_rank[q];
        if (fr >= 0) {
            ++local_valid;
            if (fr < max_rank) atomicAdd(&hist[fr], 1);
        }
    }
    atomicAdd(&nvalid_s, local_valid);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nv = nvalid_s;
        *num_valid = nv;
        int run = 0;
        for (int r = 0; r < max_rank; ++r) {
            run += hist[r];
            cmc[r] = nv > 0 ? static_cast<float>(run) / static_cast<float>(nv) : 0.f;
        }
    }
    // mAP: the float32 sum must run in query order (rank_cy.pyx accumulates sequentially in float32), so one thread
    // adds; the operands are staged through shared memory by the whole CTA so the serial chain never waits on DRAM.
    __shared__ double s_ap64[1024];
    __shared__ float s_ap32[1024];
    double s = 0.0;
    float s32 = 0.f;
    for (int base = 0; base < nq; base += 1024) {
        __syncthreads();
        for (int k = threadIdx.x; k < 1024 && base + k < nq; k += blockDim.x) {
            const bool ok = first_rank[base + k] >= 0;
            s_ap64[k] = ok ? ap64[base + k] : -1.0;
            s_ap32[k] = ap32[base + k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int n = min(1024, nq - base);
            for (int k = 0; k < n; ++k)
                if (s_ap64[k] >= 0.0) { s += s_ap64[k]; s32 += s_ap32[k]; }
        }
    }
    if (threadIdx.x == 0) {
        const int nv = nvalid_s;
        *map_out = nv > 0 ? s / nv : 0.0;
        *map32_out = nv > 0 ? s32 / static_cast<float>(nv) : 0.f;
    }
}

int rank_market1501(const float* dist, const int64_t* q_pids, const int64_t* g_pids, const int64_t* q_camids,
                    const int64_t* g_camids, int nq, int ng, int max_rank, double* ap64, float* ap32,
                    int* first_rank, float* cmc, double* map_out, int64_t* num_valid, float* map32_out,
                    int* overflow_flag, cudaStream_t stream) {
    BPB_CHECK(max_rank >= 1 && max_rank <= 1024, "rank_market1501: max_rank %d out of range [1,1024]", max_rank);
    BPB_CUDA(cudaMemsetAsync(overflow_flag, 0, sizeof(int), stream));
    rank_query_kernel<<<nq, 256, 0, stream>>>(dist, q_pids, g_pids, q_camids, g_camids, nq, ng, ap64, ap32, first_rank,
                                              overflow_flag);
    BPB_LAUNCH_CHECK("rank_query_kernel");
    rank_finalize_kernel<<<1, 256, 0, stream>>>(ap64, ap32, first_rank, nq, max_rank, cmc, map_out, num_valid,
                                                map32_out);
    BPB_LAUNCH_CHECK("rank_finalize_kernel");
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Stable ascending argsort of every row (ties -> ascending index), bitonic network in shared memory on
// 64-bit keys (ordered float bits << 32 | index).  One CTA per row; Ng <= 16384 (split arrays up to 32768, below).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ordered_bits(float f) {
    if (f == 0.f) f = 0.f;                       // -0.0 == +0.0 for comparison sorts
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(1024) argsort_rows_kernel(const float* __restrict__ dist, int ng, int npow2,
                                                            int32_t* __restrict__ indices) {
    extern __shared__ uint64_t keys[];
    const float* row = dist + static_cast<int64_t>(blockIdx.x) * ng;
    for (int j = threadIdx.x; j < npow2; j += blockDim.x)
        keys[j] = j < ng ? (static_cast<uint64_t>(ordered_bits(row[j])) << 32) | static_cast<uint32_t>(j)
                         : 0xFFFFFFFFFFFFFFFFull;
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < npow2 / 2; t += blockDim.x) {
                const int lo = ((t / j) * 2 * j) + (t % j);
                const int hi = lo + j;
                const bool up = (lo & k) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    int32_t* out = indices + static_cast<int64_t>(blockIdx.x) * ng;
    for (int j = threadIdx.x; j < ng; j += blockDim.x) out[j] = static_cast<int32_t>(keys[j] & 0xFFFFFFFFu);
}

// Same network for 16384 < Ng <= 32768 (Market-1501's gallery has 19 732 entries, DukeMTMC 17 661): the 64-bit keys no
// longer fit, so the ordered float bits and the 16-bit column index live in two shared arrays (192 KB) and are compared
// lexicographically.
__global__ void __launch_bounds__(1024) argsort_rows_split_kernel(const float* __restrict__ dist, int ng, int npow2,
                                                                  int32_t* __restrict__ indices) {
    extern __shared__ uint32_t skeys[];
    uint16_t* sidx = reinterpret_cast<uint16_t*>(skeys + npow2);
    const float* row = dist + static_cast<int64_t>(blockIdx.x) * ng;
    for (int j = threadIdx.x; j < npow2; j += blockDim.x) {
        skeys[j] = j < ng ? ordered_bits(row[j]) : 0xFFFFFFFFu;
        sidx[j] = static_cast<uint16_t>(j);                   // padding sorts behind every real entry with the same key
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < npow2 / 2; t += blockDim.x) {
                const int lo = ((t / j) * 2 * j) + (t % j);
                const int hi = lo + j;
                const bool up = (lo & k) == 0;
                const uint32_t ka = skeys[lo], kb = skeys[hi];
                const uint16_t ia = sidx[lo], ib = sidx[hi];
                const bool gt = ka > kb || (ka == kb && ia > ib);
                if (gt == up) { skeys[lo] = kb; skeys[hi] = ka; sidx[lo] = ib; sidx[hi] = ia; }
            }
            __syncthreads();
        }
    }
    int32_t* out = indices + static_cast<int64_t>(blockIdx.x) * ng;
    for (int j = threadIdx.x; j < ng; j += blockDim.x) out[j] = static_cast<int32_t>(sidx[j]);
}

int argsort_rows(const float* dist, int nq, int ng, int32_t* indices, cudaStream_t stream) {
    int npow2 = 1;
    while (npow2 < ng) npow2 <<= 1;
    BPB_CHECK(npow2 <= 32768, "argsort_rows: Ng=%d exceeds the 32768-entry shared-memory sorter", ng);
    if (npow2 <= 16384) {
        const size_t smem = static_cast<size_t>(npow2) * sizeof(uint64_t);
        BPB_SET_MAX_SMEM(argsort_rows_kernel, 128 * 1024);
        argsort_rows_kernel<<<nq, 1024, smem, stream>>>(dist, ng, npow2, indices);
    } else {
        // a real entry with key 0xFFFFFFFF (a NaN with all payload bits set) would tie with the padding: indices of the
        // padding are >= ng and still sort behind it, so the first ng outputs are the real entries
        const size_t smem = static_cast<size_t>(npow2) * (sizeof(uint32_t) + sizeof(uint16_t));
        BPB_SET_MAX_SMEM(argsort_rows_split_kernel, 192 * 1024);
        argsort_rows_split_kernel<<<nq, 1024, smem, stream>>>(dist, ng, npow2, indices);
    }
    BPB_LAUNCH_CHECK("argsort_rows_kernel");
    return 0;
}

}  // namespace bpb
