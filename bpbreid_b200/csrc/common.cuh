// bpbreid_b200 -- shared device/host helpers for the sm_100a kernels.
// Thin inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc/mma/commit/ld) and the
// UMMA shared-memory / instruction descriptors.  Field layouts follow the PTX ISA "tcgen05" chapter
// (matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64);
//  instruction descriptor: c_format [4,6), a_format [7,10), b_format [10,13), a/b major [15],[16],
//  N>>3 [17,23), M>>4 [24,29)).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace bpb {

// ----------------------------------------------------------------------------------------------
// error plumbing (no exceptions cross the C ABI)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define BPB_CHECK(cond, ...)                  \
    do {                                      \
        if (!(cond)) {                        \
            ::bpb::set_error(__VA_ARGS__);    \
            return 1;                         \
        }                                     \
    } while (0)

#define BPB_CUDA(call)                                             \
    do {                                                           \
        if (::bpb::check_cuda((call), #call)) return 2;            \
    } while (0)

// Programmatic dependent launch: a kernel launched through launch_pdl may be scheduled while its predecessor in the stream
// is still running; it must execute pdl_wait() before it touches anything the predecessor (or any earlier kernel) wrote.
// pdl_trigger() in the predecessor allows that early scheduling (it is implied at exit).  Both are no-ops for kernels
// launched the ordinary way.  The point is to hide the ~2 us launch + prologue (barrier init, TMEM allocation, descriptor
// prefetch) of every kernel of a dependent chain behind the tail of the previous one.
#ifndef BPB_PDL
#define BPB_PDL 1
#endif
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() {
#if BPB_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#if BPB_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// `allow`: measured on the HRNet step (profiles/pdl_ab_r2.txt), early scheduling pays when the kernels of a chain are short
// (batch 32 per GPU: 3.79 -> 3.49 ms per step) and costs ~0.6 % when they are long (batch 256: the waiting CTAs of the
// next kernel hold shared memory / TMEM of SMs the running kernel could still use), so callers pass a size test.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool allow, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (BPB_PDL && allow) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#define BPB_LAUNCH_CHECK(name)                                     \
    do {                                                           \
        if (::bpb::check_cuda(cudaGetLastError(), name)) return 3; \
    } while (0)

// Per-DEVICE (not per-process) one-time kernel attribute setup: cudaFuncSetAttribute applies to the current device
// only, and a DataParallel-style host drives several devices from one process (SURVEY.md section 8b "Threading").
// Idempotent, so the benign race between two host threads costs at most a redundant call.
#define BPB_SET_MAX_SMEM(kern, bytes)                                                                       \
    do {                                                                                                    \
        static unsigned long long bpb_done_[2] = {0ull, 0ull};                                              \
        int bpb_dev_ = 0;                                                                                   \
        BPB_CUDA(cudaGetDevice(&bpb_dev_));                                                                 \
        const int bpb_w_ = (bpb_dev_ >> 6) & 1;                                                             \
        if (!((bpb_done_[bpb_w_] >> (bpb_dev_ & 63)) & 1ull)) {                                             \
            BPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            bpb_done_[bpb_w_] |= 1ull << (bpb_dev_ & 63);                                                   \
        }                                                                                                   \
    } while (0)

// SM count of the CURRENT device (cached per device), 0 on error (bpb::set_error holds the message)
int device_sm_count();

// Encodes a tiled tensor map (driver entry point fetched at run time: no link dependency on libcuda).
// dims/strides are innermost-first; strides_bytes has rank-1 entries (stride of dim 1..rank-1).
int make_tensor_map(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    const uint32_t* elem_strides, CUtensorMapSwizzle swizzle);

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a descriptor/pipeline bug must surface as a trap (-> CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {      // ~2 s of SM clocks: pipeline bug, not load
            printf("bpb: mbarrier timeout block(%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x);
            __trap();
        }
    }
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}

// ---- address-based variants: the hot single-thread loops pre-compute shared-window addresses once (taking the
//      address of a __shared__ object costs an S2UR + ULEA sequence every time otherwise) ----
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_a(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_a(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            printf("bpb: mbarrier timeout block(%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x);
            __trap();
        }
    }
}
// whole-warp wait with a single polling lane: the other 31 lanes park at the warp barrier instead of hammering the
// mbarrier (a dozen spinning warps per CTA otherwise compete with TMA / UMMA / epilogue shared-memory traffic)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, bool leader) {
    if (leader) mbar_wait_a(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                              int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// One elected lane of a fully converged warp (elect.sync).  Issue loops are written so that the WHOLE warp runs the
// control flow with warp-uniform values and only the tcgen05 / TMA instruction itself is predicated on the elected
// lane: then ptxas keeps descriptors and coordinates in uniform registers.  A loop nested under `if (lane == 0)`
// instead makes every UTCHMMA / UTMALDG a vote + ELECT + 5x R2UR "waterfall" (~100 cycles per tcgen05.mma, ~300 per
// TMA box: tools/micro/umma_rate.cu, tma_rate.cu).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// warp index as a provably warp-uniform value (the shuffle is what tells the compiler)
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }

// ---- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_n(uint32_t* smem_dst, uint32_t cols) {   // whole warp, cols = 2^k >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_n(uint32_t taddr, uint32_t cols) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {      // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; kind::f16 covers bf16/f16 inputs, kind::tf32 fp32 inputs read as tf32.
template <bool kTf32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                     uint32_t accumulate) {
    if constexpr (kTf32) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane_base+i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile whose rows are `row_bytes` (= swizzle span: 128/64/32) wide: 8-row groups are
// `8*row_bytes` apart (SBO); LBO unused (single swizzle atom along K).
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((8u * row_bytes) >> 4) << 32;
    d |= 1ull << 46;            // descriptor version (sm_100)
    d |= layout << 61;
    return d;
}

// kind::f16 (bf16 x bf16 -> f32) or kind::tf32 (tf32 x tf32 -> f32), both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc(bool tf32, uint32_t M, uint32_t N) {
    uint32_t d = 0;
    d |= 1u << 4;                         // accumulator fp32
    d |= (tf32 ? 2u : 1u) << 7;           // A format: bf16 = 1, tf32 = 2
    d |= (tf32 ? 2u : 1u) << 10;          // B format
    d |= (N >> 3) << 17;
    d |= (M >> 4) << 24;
    return d;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// Mask algebra of BPBreID.forward for externally supplied part masks (torchreid/models/bpbreid.py:157-175).
//   pr[]  in: per-pixel attention probabilities (softmax of the pixel classifier, or the external masks themselves)
//         out: the probabilities the VISIBILITY scores are taken from ('hard' rewrites the part entries in place, :167)
//   pm[]  out: masks used for pooling / returned in the masks dict: pm[0] background, pm[k] part k
//   ex[]  external masks resized to the feature map (only read when mode & 6)
template <int KP1>
__device__ __forceinline__ void apply_target_segmentation(float* pr, float* pm, const float* ex, int mode) {
#pragma unroll
    for (int k = 0; k < KP1; ++k) pm[k] = pr[k];
    if (mode & 2) {                                   // 'hard': keep the parts only where the target segmentation fires
        float mx = -INFINITY;
#pragma unroll
        for (int k = 1; k < KP1; ++k) mx = fmaxf(mx, ex[k]);
        const bool target = mx > ex[0];
        if (!target) {
#pragma unroll
            for (int k = 1; k < KP1; ++k) { pr[k] = 1e-12f; pm[k] = 1e-12f; }
        }
        pm[0] = target ? 0.f : 1.f;                   // background_masks = ~target_segmentation_mask
    }
    if (mode & 4) {                                   // 'soft'
#pragma unroll
        for (int k = 1; k < KP1; ++k) pm[k] = pr[k] * ex[k];
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

#endif  // __CUDACC__

}  // namespace bpb
