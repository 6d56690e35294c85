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

// out[n, c, ho, wo] = bilinear(align_corners=True) resize of in[n, c, hi, wi]  (fp32; nn.functional.interpolate as used on
// the external part masks at torchreid/models/bpbreid.py:153, 162-163, 171-172)
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              int64_t planes, int hi, int wi, int ho, int wo) {
    const int64_t total = planes * ho * wo;
    const float ry = ho > 1 ? static_cast<float>(hi - 1) / static_cast<float>(ho - 1) : 0.f;
    const float rx = wo > 1 ? static_cast<float>(wi - 1) / static_cast<float>(wo - 1) : 0.f;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ox = i % wo, oy = (i / wo) % ho;
        const int64_t pl = i / (static_cast<int64_t>(wo) * ho);
        const float sy = ry * oy, sx = rx * ox;
        const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
        const int y1 = y0 + (y0 < hi - 1 ? 1 : 0), x1 = x0 + (x0 < wi - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
        const float* p = in + pl * hi * wi;
        out[i] = (1.f - ly) * ((1.f - lx) * p[y0 * wi + x0] + lx * p[y0 * wi + x1]) +
                 ly * ((1.f - lx) * p[y1 * wi + x0] + lx * p[y1 * wi + x1]);
    }
}

int resize_bilinear_f32(const float* in, float* out, int64_t planes, int hi, int wi, int ho, int wo, cudaStream_t stream) {
    BPB_CHECK(planes >= 1 && hi >= 1 && wi >= 1 && ho >= 1 && wo >= 1, "resize_bilinear_f32: bad shape");
    const int64_t total = planes * ho * wo;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    resize_bilinear_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, out, planes, hi, wi, ho, wo);
    BPB_LAUNCH_CHECK("resize_bilinear_kernel");
    return 0;
}

}  // namespace bpb
