// Evaluation-side glue kernels: L2 normalisation of the test embeddings (Engine.normalize,
// torchreid/engine/engine.py:558-559 == F.normalize(p=2, dim=-1, eps=1e-12)) on the device, so that the feature
// extraction loop (torchreid/engine/image/part_based_engine.py:132-166) can hand q/g features to bpb_part_distance
// without a host round trip.
#include "common.cuh"
#include "kernels.h"

namespace bpb {

// one warp per row of d floats: y = x / max(|x|_2, eps)
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           int64_t rows, int d, float eps) {
    const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* xr = x + r * d;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) s = fmaf(xr[i], xr[i], s);
    s = warp_sum(s);
    const float inv = 1.f / fmaxf(sqrtf(s), eps);
    float* yr = y + r * d;
    for (int i = lane; i < d; i += 32) yr[i] = xr[i] * inv;
}

int l2_normalize(const float* x, float* y, int64_t rows, int d, float eps, cudaStream_t stream) {
    BPB_CHECK(rows >= 0 && d >= 1, "l2_normalize: bad shape rows=%lld d=%d", (long long)rows, d);
    if (rows == 0) return 0;
    const int64_t blocks = (rows + 7) / 8;
    BPB_CHECK(blocks < (1ll << 31), "l2_normalize: too many rows");
    l2_normalize_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, y, rows, d, eps);
    BPB_LAUNCH_CHECK("l2_normalize_kernel");
    return 0;
}

}  // namespace bpb
