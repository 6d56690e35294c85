// k-reciprocal re-ranking on the device (reference: torchreid/utils/rerank.py:31-113, called from
// part_based_engine.py:218-226 with the query-gallery, query-query and gallery-gallery part-based distance matrices).
//
// The reference builds dense (Nq+Ng)^2 numpy matrices for the weights V and loops in Python.  Here only the normalised
// distance matrix is dense; the neighbour table is a top-(k1+1) selection per row (no full sort), V is kept as
// fixed-capacity sparse rows (a row has at most (k1+1)(1+k1/2+1) non-zeros), the local query expansion is a sorted
// merge of k2 sparse rows, and the Jaccard pass walks a column-compressed copy of V.  Integer results (neighbour
// table, expanded sets) are bit-exact against the reference; the floating-point stages follow its operation order
// (separate IEEE multiply / divide / add, sums in ascending column order) so they agree to the last few ulps.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>

#include "common.cuh"
#include "kernels.h"

namespace bpb {
namespace {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

__device__ __forceinline__ float sq_rn(float x) { return __fmul_rn(x, x); }

// ---- column / row maxima of the squared block matrix (rerank.py:44-45: np.max(original_dist, axis=0)) ----------
__global__ void __launch_bounds__(256) rr_colmax_kernel(const float* __restrict__ m, int rows, int cols, int rows_per_block,
                                                        float* __restrict__ cmax) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= cols) return;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float best = 0.f;
    for (int r = r0; r < r1; ++r) best = fmaxf(best, sq_rn(m[static_cast<size_t>(r) * cols + c]));
    atomicMax(reinterpret_cast<unsigned int*>(cmax + c), __float_as_uint(best));      // squares are >= 0
}

__global__ void __launch_bounds__(256) rr_rowmax_kernel(const float* __restrict__ m, int cols, float* __restrict__ cmax) {
    const float* row = m + static_cast<size_t>(blockIdx.x) * cols;
    float best = 0.f;
    for (int c = threadIdx.x; c < cols; c += 256) best = fmaxf(best, sq_rn(row[c]));
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = fmaxf(best, red[w]);
        atomicMax(reinterpret_cast<unsigned int*>(cmax + blockIdx.x), __float_as_uint(best));
    }
}

// dst[r0 + r][c0 + c] = src(r, c)^2 / cmax[r0 + r]   with src(r, c) = transpose ? src[c][r] : src[r][c]
// (rerank.py:44-46: square, divide every column by its maximum, transpose)
__global__ void __launch_bounds__(256) rr_build_kernel(const float* __restrict__ src, int src_ld, int rows, int cols,
                                                       int transpose, float* __restrict__ dst, int dst_ld, int r0, int c0,
                                                       const float* __restrict__ cmax) {
    __shared__ float tile[32][33];
    const int tr = blockIdx.y * 32, tc = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 32 x 8
    if (transpose) {
        for (int y = ty; y < 32; y += 8) {                                // read src[c][r] rows coalesced along r
            const int c = tc + y, r = tr + tx;
            tile[y][tx] = (c < cols && r < rows) ? src[static_cast<size_t>(c) * src_ld + r] : 0.f;
        }
        __syncthreads();
        for (int y = ty; y < 32; y += 8) {
            const int r = tr + y, c = tc + tx;
            if (r < rows && c < cols)
                dst[static_cast<size_t>(r0 + r) * dst_ld + c0 + c] = __fdiv_rn(sq_rn(tile[tx][y]), cmax[r0 + r]);
        }
    } else {
        for (int y = ty; y < 32; y += 8) {
            const int r = tr + y, c = tc + tx;
            if (r < rows && c < cols)
                dst[static_cast<size_t>(r0 + r) * dst_ld + c0 + c] =
                    __fdiv_rn(sq_rn(src[static_cast<size_t>(r) * src_ld + c]), cmax[r0 + r]);
        }
    }
}

// ---- neighbour table: the first k entries of the stable argsort of every row (rerank.py:47) ----------------------
// One CTA per row; the row is cached in shared memory when it fits.  Round t picks the smallest (value, index) pair
// that is lexicographically greater than the pair picked in round t-1: no marking, ties resolved by index.
__global__ void __launch_bounds__(256) rr_topk_kernel(const float* __restrict__ orig, int a, int k, int cached,
                                                      int* __restrict__ rank) {
    extern __shared__ float srow[];
    __shared__ float red_v[8];
    __shared__ int red_i[8];
    __shared__ float s_pv;
    __shared__ int s_pi;
    for (int i = blockIdx.x; i < a; i += gridDim.x) {
        const float* grow = orig + static_cast<size_t>(i) * a;
        if (cached) {
            __syncthreads();
            for (int j = threadIdx.x; j < a; j += 256) srow[j] = grow[j];
            __syncthreads();
        }
        const float* row = cached ? srow : grow;
        float pv = -FLT_MAX;
        int pi = -1;
        for (int t = 0; t < k; ++t) {
            float bv = FLT_MAX;
            int bi = INT_MAX;
            for (int j = threadIdx.x; j < a; j += 256) {
                const float v = row[j];
                const bool after = (v > pv) || (v == pv && j > pi);
                if (after && ((v < bv) || (v == bv && j < bi))) { bv = v; bi = j; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if ((ov < bv) || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = bv; red_i[threadIdx.x >> 5] = bi; }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < 8; ++w)
                    if ((red_v[w] < bv) || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
                s_pv = bv;
                s_pi = bi;
                rank[static_cast<size_t>(i) * k + t] = bi;
            }
            __syncthreads();
            pv = s_pv;
            pi = s_pi;
        }
    }
}

// ---- expanded k-reciprocal sets and their Gaussian weights (rerank.py:54-79) --------------------------------------
// One warp per row.  R(i) = {f in N(i, kk): i in N(f, kk)}; for every c in R(i): R_half(c) likewise with `half`
// neighbours, appended when |R_half(c) & R(i)| > 2/3 |R_half(c)|; duplicates removed; V[i, e] = exp(-orig[i, e]) / sum.
constexpr int kRrWarps = 4;

__global__ void __launch_bounds__(kRrWarps * 32) rr_kreciprocal_kernel(const float* __restrict__ orig,
                                                                       const int* __restrict__ rank, int a, int rk, int kk,
                                                                       int half, int cap, int* __restrict__ vidx,
                                                                       float* __restrict__ vval, int* __restrict__ vcnt) {
    extern __shared__ int s_sets[];                                       // per warp: R[32] | E[cap] | U[cap]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int* R = s_sets + warp * (32 + 2 * cap);
    int* E = R + 32;
    int* U = E + cap;
    const int i = blockIdx.x * kRrWarps + warp;
    if (i >= a) return;
    const int c = lane < kk ? rank[static_cast<size_t>(i) * rk + lane] : -1;
    bool rec = false;
    if (lane < kk)
        for (int t = 0; t < kk; ++t) rec |= rank[static_cast<size_t>(c) * rk + t] == i;
    const unsigned m1 = __ballot_sync(0xffffffffu, rec);
    const int nr = __popc(m1);
    if (rec) {
        const int pos = __popc(m1 & lt);
        R[pos] = c;
        E[pos] = c;
    }
    int ne = nr;
    __syncwarp();
    for (int x = 0; x < nr; ++x) {
        const int cand = R[x];
        const int f = lane < half ? rank[static_cast<size_t>(cand) * rk + lane] : -1;
        bool rc = false;
        if (lane < half)
            for (int t = 0; t < half; ++t) rc |= rank[static_cast<size_t>(f) * rk + t] == cand;
        const unsigned m2 = __ballot_sync(0xffffffffu, rc);
        const int n2 = __popc(m2);
        bool common = false;
        if (rc)
            for (int y = 0; y < nr; ++y) common |= R[y] == f;
        const int ncommon = __popc(__ballot_sync(0xffffffffu, common));
        if (static_cast<double>(ncommon) > 2.0 / 3.0 * static_cast<double>(n2)) {      // rerank.py:71-74
            if (rc) E[ne + __popc(m2 & lt)] = f;
            ne += n2;
        }
        __syncwarp();
    }
    // np.unique: keep first occurrences (the order inside a row does not matter to any later stage: the expansion
    // kernel sorts by column)
    int nu = 0;
    for (int base = 0; base < ne; base += 32) {
        const int e = base + lane;
        bool first = e < ne;
        if (first) {
            const int v = E[e];
            for (int y = 0; y < e; ++y) first &= E[y] != v;
        }
        const unsigned mf = __ballot_sync(0xffffffffu, first);
        if (first) U[nu + __popc(mf & lt)] = E[e];
        nu += __popc(mf);
    }
    __syncwarp();
    float sum = 0.f;
    const float* row = orig + static_cast<size_t>(i) * a;
    for (int u = lane; u < nu; u += 32) {
        const float w = expf(-row[U[u]]);
        vval[static_cast<size_t>(i) * cap + u] = w;
        vidx[static_cast<size_t>(i) * cap + u] = U[u];
        sum += w;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    for (int u = lane; u < nu; u += 32)
        vval[static_cast<size_t>(i) * cap + u] = __fdiv_rn(vval[static_cast<size_t>(i) * cap + u], sum);
    if (lane == 0) vcnt[i] = nu;
}

// ---- local query expansion (rerank.py:81-86): V_qe[i] = mean of the rows V[N(i, k2)] ----------------------------
// Sparse merge: the <= k2*cap entries are sorted by (column, neighbour position) in shared memory, runs of equal column
// are summed in neighbour order (the order numpy's axis-0 reduction adds them in) and divided by k2.
constexpr int kRrSort = 4096;

__global__ void __launch_bounds__(256) rr_qexpand_kernel(const int* __restrict__ rank, int rk, int k2, int identity,
                                                         const int* __restrict__ vidx, const float* __restrict__ vval,
                                                         const int* __restrict__ vcnt, int cap, int a, int qcap,
                                                         int* __restrict__ qidx, float* __restrict__ qval,
                                                         int* __restrict__ qcnt) {
    __shared__ unsigned int keys[kRrSort];
    __shared__ float vals[kRrSort];
    __shared__ int s_off[33];
    __shared__ int s_heads[256];
    for (int i = blockIdx.x; i < a; i += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int o = 0;
            for (int j = 0; j < k2; ++j) {
                s_off[j] = o;
                o += vcnt[identity ? i : rank[static_cast<size_t>(i) * rk + j]];
            }
            s_off[k2] = o;
        }
        __syncthreads();
        const int total = s_off[k2];
        int npow2 = 32;
        while (npow2 < total) npow2 <<= 1;
        for (int p = threadIdx.x; p < npow2; p += 256) { keys[p] = 0xFFFFFFFFu; vals[p] = 0.f; }
        __syncthreads();
        for (int j = 0; j < k2; ++j) {
            const int src = identity ? i : rank[static_cast<size_t>(i) * rk + j];
            const int n = s_off[j + 1] - s_off[j];
            for (int e = threadIdx.x; e < n; e += 256) {
                keys[s_off[j] + e] = static_cast<unsigned int>(vidx[static_cast<size_t>(src) * cap + e]) * 32u + j;
                vals[s_off[j] + e] = vval[static_cast<size_t>(src) * cap + e];
            }
        }
        __syncthreads();
        for (int k = 2; k <= npow2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < npow2 / 2; t += 256) {
                    const int lo = ((t / j) * 2 * j) + (t % j);
                    const int hi = lo + j;
                    const bool up = (lo & k) == 0;
                    const unsigned int ka = keys[lo], kb = keys[hi];
                    if ((ka > kb) == up) {
                        keys[lo] = kb; keys[hi] = ka;
                        const float va = vals[lo]; vals[lo] = vals[hi]; vals[hi] = va;
                    }
                }
                __syncthreads();
            }
        }
        // heads of the runs of equal column: thread t owns the contiguous chunk [t*per, (t+1)*per)
        const int per = npow2 / 256 > 0 ? npow2 / 256 : 1;
        const int p0 = threadIdx.x * per, p1 = min(total, p0 + per);
        int nh = 0;
        for (int p = p0; p < p1; ++p) nh += (p == 0 || (keys[p] >> 5) != (keys[p - 1] >> 5)) ? 1 : 0;
        s_heads[threadIdx.x] = nh;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int t = 0; t < 256; ++t) { const int h = s_heads[t]; s_heads[t] = run; run += h; }
            qcnt[i] = run;
        }
        __syncthreads();
        int out = s_heads[threadIdx.x];
        const float inv_n = static_cast<float>(k2);
        for (int p = p0; p < p1; ++p) {
            if (p == 0 || (keys[p] >> 5) != (keys[p - 1] >> 5)) {
                const unsigned int col = keys[p] >> 5;
                float s = vals[p];
                for (int e = p + 1; e < total && (keys[e] >> 5) == col; ++e) s = __fadd_rn(s, vals[e]);
                qidx[static_cast<size_t>(i) * qcap + out] = static_cast<int>(col);
                qval[static_cast<size_t>(i) * qcap + out] = __fdiv_rn(s, inv_n);
                ++out;
            }
        }
    }
}

// ---- column-compressed copy of V (rerank.py:88-90: invIndex) -------------------------------------------------------
__global__ void __launch_bounds__(256) rr_csc_count_kernel(const int* __restrict__ qidx, const int* __restrict__ qcnt,
                                                           int qcap, int a, int* __restrict__ colcnt) {
    for (int i = blockIdx.x; i < a; i += gridDim.x)
        for (int e = threadIdx.x; e < qcnt[i]; e += 256) atomicAdd(colcnt + qidx[static_cast<size_t>(i) * qcap + e], 1);
}

__global__ void __launch_bounds__(256) rr_csc_fill_kernel(const int* __restrict__ qidx, const float* __restrict__ qval,
                                                          const int* __restrict__ qcnt, int qcap, int a,
                                                          const int* __restrict__ col_ptr, int* __restrict__ cursor,
                                                          int* __restrict__ crow, float* __restrict__ cval) {
    for (int i = blockIdx.x; i < a; i += gridDim.x)
        for (int e = threadIdx.x; e < qcnt[i]; e += 256) {
            const int c = qidx[static_cast<size_t>(i) * qcap + e];
            const int pos = col_ptr[c] + atomicAdd(cursor + c, 1);
            crow[pos] = i;
            cval[pos] = qval[static_cast<size_t>(i) * qcap + e];
        }
}

// ---- Jaccard distance + final combination (rerank.py:92-111) ------------------------------------------------------
// One CTA per query row; t[g] accumulates min(V[i,c], V[g,c]) over the non-zero columns c of row i in ascending order
// (one barrier per column: the entries of one column have distinct rows g, so there are no conflicts and the
// summation order per g is the reference's).
__global__ void __launch_bounds__(256) rr_jaccard_kernel(const float* __restrict__ orig, const int* __restrict__ qidx,
                                                         const float* __restrict__ qval, const int* __restrict__ qcnt,
                                                         int qcap, const int* __restrict__ col_ptr,
                                                         const int* __restrict__ crow, const float* __restrict__ cval,
                                                         int nq, int a, float one_minus_lambda, float lambda, int cached,
                                                         float* __restrict__ scratch, float* __restrict__ out) {
    extern __shared__ float st[];
    float* t = cached ? st : scratch + static_cast<size_t>(blockIdx.x) * a;
    const int ng = a - nq;
    for (int i = blockIdx.x; i < nq; i += gridDim.x) {
        __syncthreads();
        for (int g = threadIdx.x; g < a; g += 256) t[g] = 0.f;
        __syncthreads();
        const int n = qcnt[i];
        for (int e = 0; e < n; ++e) {
            const int c = qidx[static_cast<size_t>(i) * qcap + e];
            const float v = qval[static_cast<size_t>(i) * qcap + e];
            const int beg = col_ptr[c], end = col_ptr[c + 1];
            for (int p = beg + threadIdx.x; p < end; p += 256) {
                const int g = crow[p];
                t[g] = __fadd_rn(t[g], fminf(v, cval[p]));
            }
            __syncthreads();
        }
        const float* row = orig + static_cast<size_t>(i) * a + nq;
        for (int j = threadIdx.x; j < ng; j += 256) {
            const float tj = t[nq + j];
            const float jac = __fsub_rn(1.f, __fdiv_rn(tj, __fsub_rn(2.f, tj)));
            out[static_cast<size_t>(i) * ng + j] = __fadd_rn(__fmul_rn(jac, one_minus_lambda), __fmul_rn(row[j], lambda));
        }
    }
}

constexpr int kSmemRow = 200 * 1024;

}  // namespace

int rerank_build(const float* qg, const float* qq, const float* gg, int nq, int ng, float* cmax, float* orig,
                 cudaStream_t stream) {
    BPB_CHECK(nq > 0 && ng > 0, "rerank_build: empty query or gallery set (nq=%d ng=%d)", nq, ng);
    const int a = nq + ng;
    BPB_CUDA(cudaMemsetAsync(cmax, 0, sizeof(float) * a, stream));
    const int rpb = 256;
    rr_colmax_kernel<<<dim3(ceil_div(nq, 256), ceil_div(nq, rpb)), 256, 0, stream>>>(qq, nq, nq, rpb, cmax);
    rr_rowmax_kernel<<<nq, 256, 0, stream>>>(qg, ng, cmax);
    rr_colmax_kernel<<<dim3(ceil_div(ng, 256), ceil_div(nq, rpb)), 256, 0, stream>>>(qg, nq, ng, rpb, cmax + nq);
    rr_colmax_kernel<<<dim3(ceil_div(ng, 256), ceil_div(ng, rpb)), 256, 0, stream>>>(gg, ng, ng, rpb, cmax + nq);
    BPB_LAUNCH_CHECK("rr_colmax_kernel");
    // rows [0, nq): [ qq^T | qg ] ; rows [nq, a): [ qg^T | gg^T ]
    rr_build_kernel<<<dim3(ceil_div(nq, 32), ceil_div(nq, 32)), 256, 0, stream>>>(qq, nq, nq, nq, 1, orig, a, 0, 0, cmax);
    rr_build_kernel<<<dim3(ceil_div(ng, 32), ceil_div(nq, 32)), 256, 0, stream>>>(qg, ng, nq, ng, 0, orig, a, 0, nq, cmax);
    rr_build_kernel<<<dim3(ceil_div(nq, 32), ceil_div(ng, 32)), 256, 0, stream>>>(qg, ng, ng, nq, 1, orig, a, nq, 0, cmax);
    rr_build_kernel<<<dim3(ceil_div(ng, 32), ceil_div(ng, 32)), 256, 0, stream>>>(gg, ng, ng, ng, 1, orig, a, nq, nq, cmax);
    BPB_LAUNCH_CHECK("rr_build_kernel");
    return 0;
}

int rerank_topk(const float* orig, int a, int k, int* rank, cudaStream_t stream) {
    BPB_CHECK(k >= 1 && k <= a, "rerank_topk: k=%d out of range for %d rows", k, a);
    const size_t row_bytes = static_cast<size_t>(a) * sizeof(float);
    const int cached = row_bytes <= kSmemRow ? 1 : 0;
    BPB_SET_MAX_SMEM(rr_topk_kernel, kSmemRow);
    const int per_sm = cached ? static_cast<int>(kSmemRow / (row_bytes > 1024 ? row_bytes : 1024)) : 8;
    int blocks = device_sm_count() * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    if (blocks > a) blocks = a;
    rr_topk_kernel<<<blocks, 256, cached ? row_bytes : 0, stream>>>(orig, a, k, cached, rank);
    BPB_LAUNCH_CHECK("rr_topk_kernel");
    return 0;
}

int rerank_row_capacity(int k1) {
    // (k1+1) members of R(i) + up to (k1+1) appended half-sets of round(k1/2)+1 members, rounded up to 32
    const int half = static_cast<int>(nearbyint(k1 / 2.0)) + 1;
    const int cap = (k1 + 1) * (1 + half);
    return (cap + 31) / 32 * 32;
}

int rerank_kreciprocal(const float* orig, const int* rank, int a, int rk, int k1, int* vidx, float* vval, int* vcnt,
                       cudaStream_t stream) {
    BPB_CHECK(k1 >= 1 && k1 + 1 <= 32, "rerank_kreciprocal: k1=%d (one warp lane per neighbour: k1 <= 31)", k1);
    const int kk = k1 + 1 < a ? k1 + 1 : a;
    int half = static_cast<int>(nearbyint(k1 / 2.0)) + 1;                  // np.around: half to even (rerank.py:65)
    if (half > a) half = a;
    BPB_CHECK(rk >= kk, "rerank_kreciprocal: neighbour table has %d columns, needs %d", rk, kk);
    const int cap = rerank_row_capacity(k1);
    const size_t smem = static_cast<size_t>(kRrWarps) * (32 + 2 * cap) * sizeof(int);
    BPB_CHECK(smem <= 200 * 1024, "rerank_kreciprocal: k1=%d needs %zu bytes of shared memory", k1, smem);
    BPB_SET_MAX_SMEM(rr_kreciprocal_kernel, 200 * 1024);
    rr_kreciprocal_kernel<<<ceil_div(a, kRrWarps), kRrWarps * 32, smem, stream>>>(orig, rank, a, rk, kk, half, cap, vidx,
                                                                                  vval, vcnt);
    BPB_LAUNCH_CHECK("rr_kreciprocal_kernel");
    return 0;
}

int rerank_qexpand(const int* rank, int rk, int k2, const int* vidx, const float* vval, const int* vcnt, int k1, int a,
                   int* qidx, float* qval, int* qcnt, cudaStream_t stream) {
    const int cap = rerank_row_capacity(k1);
    // k2 == 1: the reference skips the expansion (rerank.py:81); the pass still runs to sort every row by column
    const int identity = k2 == 1 ? 1 : 0;
    const int k2e = k2 < a ? k2 : a;
    BPB_CHECK(k2e >= 1 && k2e <= 32 && k2e <= rk, "rerank_qexpand: k2=%d out of range (<= 32, <= %d table columns)", k2, rk);
    BPB_CHECK(k2e * cap <= kRrSort, "rerank_qexpand: k2*capacity = %d exceeds the %d-entry merge buffer", k2e * cap, kRrSort);
    BPB_CHECK(a < (1 << 27), "rerank_qexpand: %d rows exceed the 27-bit column key", a);
    int blocks = device_sm_count() * 4;
    if (blocks > a) blocks = a;
    rr_qexpand_kernel<<<blocks, 256, 0, stream>>>(rank, rk, k2e, identity, vidx, vval, vcnt, cap, a, k2e * cap, qidx, qval, qcnt);
    BPB_LAUNCH_CHECK("rr_qexpand_kernel");
    return 0;
}

int rerank_csc_count(const int* qidx, const int* qcnt, int qcap, int a, int* colcnt, cudaStream_t stream) {
    BPB_CUDA(cudaMemsetAsync(colcnt, 0, sizeof(int) * a, stream));
    int blocks = device_sm_count() * 8;
    if (blocks > a) blocks = a;
    rr_csc_count_kernel<<<blocks, 256, 0, stream>>>(qidx, qcnt, qcap, a, colcnt);
    BPB_LAUNCH_CHECK("rr_csc_count_kernel");
    return 0;
}

int rerank_csc_fill(const int* qidx, const float* qval, const int* qcnt, int qcap, int a, const int* col_ptr, int* cursor,
                    int* crow, float* cval, cudaStream_t stream) {
    BPB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * a, stream));
    int blocks = device_sm_count() * 8;
    if (blocks > a) blocks = a;
    rr_csc_fill_kernel<<<blocks, 256, 0, stream>>>(qidx, qval, qcnt, qcap, a, col_ptr, cursor, crow, cval);
    BPB_LAUNCH_CHECK("rr_csc_fill_kernel");
    return 0;
}

int rerank_jaccard_blocks(int nq, int a) {
    const size_t row_bytes = static_cast<size_t>(a) * sizeof(float);
    int per_sm = row_bytes <= kSmemRow ? static_cast<int>(kSmemRow / (row_bytes > 1024 ? row_bytes : 1024)) : 2;
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int blocks = device_sm_count() * per_sm;
    return blocks < nq ? blocks : nq;
}

int rerank_jaccard(const float* orig, const int* qidx, const float* qval, const int* qcnt, int qcap, const int* col_ptr,
                   const int* crow, const float* cval, int nq, int a, float one_minus_lambda, float lambda, float* scratch,
                   float* out, cudaStream_t stream) {
    const size_t row_bytes = static_cast<size_t>(a) * sizeof(float);
    const int cached = row_bytes <= kSmemRow ? 1 : 0;
    BPB_CHECK(cached || scratch != nullptr, "rerank_jaccard: %d rows need a global scratch of blocks*a floats", a);
    BPB_SET_MAX_SMEM(rr_jaccard_kernel, kSmemRow);
    rr_jaccard_kernel<<<rerank_jaccard_blocks(nq, a), 256, cached ? row_bytes : 0, stream>>>(
        orig, qidx, qval, qcnt, qcap, col_ptr, crow, cval, nq, a, one_minus_lambda, lambda, cached, scratch, out);
    BPB_LAUNCH_CHECK("rr_jaccard_kernel");
    return 0;
}

}  // namespace bpb
