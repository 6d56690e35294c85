// Per-sample input transforms on the device (reference: torchreid/data/transforms.py:24-163 and
// torchreid/data/masks_transforms/mask_transform.py:21-85).  One launch per batch for the images, one for the masks;
// every sample may have its own source size (descriptor tables), augmentation parameters are drawn by the host.
//
// Images: cv2.resize(INTER_LINEAR) on uint8 as albumentations.Resize does it -- OpenCV's 11-bit fixed-point scheme
// (weights rounded half-to-even to shorts, horizontal taps clamped onto the border pixel, vertical ROW INDICES clipped
// separately, ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2), bit-exact -- then flip, zero pad + crop, Normalize
// (float32: (x - 255*mean) * (1/(255*std)), separate IEEE ops), CoarseDropout fill (behind Normalize in the reference's
// list, so the fill value is written un-normalised), NCHW float32 out.
// Masks: cv2.resize(INTER_NEAREST) of the 36 confidence fields to the image size, flip / pad + crop / erase with 0,
// part grouping (max or sum, clamp to [0,1]), background channel, softmax(weight * m) over the K+1 channels, nearest
// resize to (H/scale, W/scale) -- composed into one gather per output pixel.
#include <cstdint>

#include "common.cuh"
#include "kernels.h"

namespace bpb {
namespace {

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// geometry row of one sample (int32 x 10): source h, w | flip | crop x, y (-1: no crop) | erase x0, y0, x1, y1 | unused
constexpr int kGeom = 10;

struct LinCoef {
    int i0, i1;     // source taps
    int w0, w1;     // 11-bit fixed-point weights
};

// resize.cpp INTER_LINEAR coefficient of destination coordinate d (dst -> src); horizontal taps clamp the weights too
__device__ __forceinline__ LinCoef lin_coef(int d, int dst, int src, bool horizontal) {
    const double scale = static_cast<double>(src) / static_cast<double>(dst);
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(floorf(f));
    f = __fsub_rn(f, static_cast<float>(s));
    LinCoef c;
    if (horizontal) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        c.i0 = s;
        c.i1 = min(s + 1, src - 1);
    } else {
        c.i0 = min(max(s, 0), src - 1);
        c.i1 = min(max(s + 1, 0), src - 1);
    }
    c.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));      // saturate_cast<short>: round half to even
    c.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
    return c;
}

struct NormArgs {
    float mean255[3], inv_std255[3], fill[3];
};

__global__ void __launch_bounds__(256) preprocess_images_kernel(const uint8_t* __restrict__ pix,
                                                                const int64_t* __restrict__ offsets,
                                                                const int32_t* __restrict__ geom, int height, int width,
                                                                int pad, NormArgs na, float* __restrict__ out) {
    const int n = blockIdx.y;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= height * width) return;
    const int oy = p / width, ox = p % width;
    const int32_t* g = geom + static_cast<size_t>(n) * kGeom;
    const int sh = g[0], sw = g[1], flip = g[2], cx = g[3], cy = g[4];
    float* o = out + (static_cast<size_t>(n) * 3) * height * width + p;
    const size_t plane = static_cast<size_t>(height) * width;
    if (ox >= g[5] && ox < g[7] && oy >= g[6] && oy < g[8]) {          // CoarseDropout rectangle (transforms.py:112-117)
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c * plane] = na.fill[c];
        return;
    }
    int px = ox, py = oy;
    bool inside = true;
    if (cx >= 0) {                                                     // PadIfNeeded(0) + RandomCrop (transforms.py:92-95)
        px = ox + cx - pad;
        py = oy + cy - pad;
        inside = px >= 0 && px < width && py >= 0 && py < height;
    }
    int v[3] = {0, 0, 0};
    if (inside) {
        if (flip) px = width - 1 - px;
        const uint8_t* src = pix + offsets[n];
        if (sh == height && sw == width) {                             // cv2.resize copies
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = src[(static_cast<size_t>(py) * sw + px) * 3 + c];
        } else {
            const LinCoef hx = lin_coef(px, width, sw, true), vy = lin_coef(py, height, sh, false);
            const uint8_t* r0 = src + static_cast<size_t>(vy.i0) * sw * 3;
            const uint8_t* r1 = src + static_cast<size_t>(vy.i1) * sw * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int s0 = r0[hx.i0 * 3 + c] * hx.w0 + r0[hx.i1 * 3 + c] * hx.w1;
                const int s1 = r1[hx.i0 * 3 + c] * hx.w0 + r1[hx.i1 * 3 + c] * hx.w1;
                const int r = (((vy.w0 * (s0 >> 4)) >> 16) + ((vy.w1 * (s1 >> 4)) >> 16) + 2) >> 2;
                v[c] = min(max(r, 0), 255);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        o[c * plane] = __fmul_rn(__fsub_rn(static_cast<float>(v[c]), na.mean255[c]), na.inv_std255[c]);
}

constexpr int kMaxParts = 16;
constexpr int kFields = 36;

struct MaskArgs {
    unsigned long long group_bits[kMaxParts];
    int parts;
    int combine_sum;          // 0: max, 1: sum
    int strategy;             // 0: sum, 1: threshold, 2: diff_from_max
    float threshold;
    float softmax_weight;     // <= 0: divide by the channel sum instead
};

__global__ void __launch_bounds__(128) preprocess_masks_kernel(const float* __restrict__ fields,
                                                               const int64_t* __restrict__ offsets,
                                                               const int32_t* __restrict__ geom, int height, int width,
                                                               int ho, int wo, int pad, MaskArgs ma,
                                                               float* __restrict__ out) {
    const int n = blockIdx.y;
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= ho * wo) return;
    const int oy = p / wo, ox = p % wo;
    const int32_t* g = geom + static_cast<size_t>(n) * kGeom;
    const int mh = g[0], mw = g[1], flip = g[2], cx = g[3], cy = g[4];
    // ResizeMasks: F.interpolate(mode='nearest') from (height, width) (mask_transform.py:52)
    const int y1 = min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(oy), static_cast<float>(height) / static_cast<float>(ho)))), height - 1);
    const int x1 = min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(ox), static_cast<float>(width) / static_cast<float>(wo)))), width - 1);
    bool live = !(x1 >= g[5] && x1 < g[7] && y1 >= g[6] && y1 < g[8]);             // mask_fill_value = 0
    int px = x1, py = y1;
    if (live && cx >= 0) {
        px = x1 + cx - pad;
        py = y1 + cy - pad;
        live = px >= 0 && px < width && py >= 0 && py < height;                     // mask_value = 0 in the padding
    }
    float f[kFields];
#pragma unroll
    for (int k = 0; k < kFields; ++k) f[k] = 0.f;
    if (live) {
        if (flip) px = width - 1 - px;
        // albumentations.Resize on masks: cv2 INTER_NEAREST, min(floor(x * (1 / (dst / src))), src - 1) in double
        const int sx = min(static_cast<int>(floor(px * (1.0 / (static_cast<double>(width) / mw)))), mw - 1);
        const int sy = min(static_cast<int>(floor(py * (1.0 / (static_cast<double>(height) / mh)))), mh - 1);
        const float* src = fields + offsets[n] + static_cast<size_t>(sy) * mw + sx;
        const size_t plane = static_cast<size_t>(mh) * mw;
#pragma unroll
        for (int k = 0; k < kFields; ++k) f[k] = src[k * plane];
    }
    float m[kMaxParts + 1];
    float mx = -3.4e38f, sum = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxParts; ++q) {
        if (q < ma.parts) {
            float v = ma.combine_sum ? 0.f : -3.4e38f;
#pragma unroll
            for (int k = 0; k < kFields; ++k)
                if ((ma.group_bits[q] >> k) & 1ull) v = ma.combine_sum ? __fadd_rn(v, f[k]) : fmaxf(v, f[k]);
            v = fminf(fmaxf(v, 0.f), 1.f);
            m[q + 1] = v;
            mx = fmaxf(mx, v);
            sum = __fadd_rn(sum, v);
        }
    }
    if (ma.strategy == 0) m[0] = fminf(fmaxf(__fsub_rn(1.f, sum), 0.f), 1.f);       // mask_transform.py:67-79
    else if (ma.strategy == 1) m[0] = mx < ma.threshold ? 1.f : 0.f;
    else m[0] = fminf(fmaxf(__fsub_rn(1.f, mx), 0.f), 1.f);
    const int ch = ma.parts + 1;
    float* o = out + (static_cast<size_t>(n) * ch) * ho * wo + p;
    const size_t oplane = static_cast<size_t>(ho) * wo;
    if (ma.softmax_weight > 0.f) {                                                   // F.softmax(masks * w, dim=0)
        float zmax = -3.4e38f;
#pragma unroll
        for (int q = 0; q <= kMaxParts; ++q)
            if (q < ch) { m[q] = __fmul_rn(m[q], ma.softmax_weight); zmax = fmaxf(zmax, m[q]); }
        float den = 0.f;
#pragma unroll
        for (int q = 0; q <= kMaxParts; ++q)
            if (q < ch) { m[q] = expf(__fsub_rn(m[q], zmax)); den = __fadd_rn(den, m[q]); }
#pragma unroll
        for (int q = 0; q <= kMaxParts; ++q)
            if (q < ch) o[q * oplane] = __fdiv_rn(m[q], den);
    } else {                                                                         // masks / masks.sum(dim=0)
        float den = 0.f;
#pragma unroll
        for (int q = 0; q <= kMaxParts; ++q)
            if (q < ch) den = __fadd_rn(den, m[q]);
#pragma unroll
        for (int q = 0; q <= kMaxParts; ++q)
            if (q < ch) o[q * oplane] = __fdiv_rn(m[q], den);
    }
}

}  // namespace

int preprocess_images(const uint8_t* pixels, const int64_t* offsets, const int32_t* geom, int n, int height, int width,
                      int pad, const float* mean255, const float* inv_std255, const float* erase_fill, float* out,
                      cudaStream_t stream) {
    BPB_CHECK(n > 0 && height > 0 && width > 0 && pad >= 0, "preprocess_images: bad sizes n=%d %dx%d pad=%d", n, height, width, pad);
    NormArgs na;
    for (int c = 0; c < 3; ++c) { na.mean255[c] = mean255[c]; na.inv_std255[c] = inv_std255[c]; na.fill[c] = erase_fill[c]; }
    preprocess_images_kernel<<<dim3(ceil_div(static_cast<int64_t>(height) * width, 256), n), 256, 0, stream>>>(
        pixels, offsets, geom, height, width, pad, na, out);
    BPB_LAUNCH_CHECK("preprocess_images_kernel");
    return 0;
}

int preprocess_masks(const float* fields, const int64_t* offsets, const int32_t* geom, int n, int height, int width,
                     int out_h, int out_w, int pad, const uint64_t* group_bits, int parts, int combine_sum, int strategy,
                     float threshold, float softmax_weight, float* out, cudaStream_t stream) {
    BPB_CHECK(parts >= 1 && parts <= kMaxParts, "preprocess_masks: %d parts (1..%d)", parts, kMaxParts);
    BPB_CHECK(strategy >= 0 && strategy <= 2, "preprocess_masks: background strategy %d", strategy);
    BPB_CHECK(n > 0 && out_h > 0 && out_w > 0, "preprocess_masks: bad sizes");
    MaskArgs ma{};
    for (int q = 0; q < parts; ++q) {
        BPB_CHECK((group_bits[q] >> kFields) == 0 && group_bits[q] != 0, "preprocess_masks: group %d selects no / unknown fields", q);
        ma.group_bits[q] = group_bits[q];
    }
    ma.parts = parts;
    ma.combine_sum = combine_sum;
    ma.strategy = strategy;
    ma.threshold = threshold;
    ma.softmax_weight = softmax_weight;
    preprocess_masks_kernel<<<dim3(ceil_div(static_cast<int64_t>(out_h) * out_w, 128), n), 128, 0, stream>>>(
        fields, offsets, geom, height, width, out_h, out_w, pad, ma, out);
    BPB_LAUNCH_CHECK("preprocess_masks_kernel");
    return 0;
}

}  // namespace bpb
