// C ABI shim (include/bpbreid_sm100.h) + error plumbing + tensor-map encoding.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bpbreid_sm100.h"
#include "common.cuh"
#include "kernels.h"

namespace bpb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return 1;
}

int device_sm_count() {
    static int cache[128] = {0};
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 0;
    if (dev < 0 || dev >= 128) dev = 0;
    if (cache[dev] == 0) {
        int n = 0;
        if (check_cuda(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return 0;
        cache[dev] = n;
    }
    return cache[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || p == nullptr) {
            set_error("cuTensorMapEncodeTiled driver entry point unavailable");
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return 1;
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = elem_strides[i]; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    const CUresult r = fn(out, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, s, b, e,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u] "
                  "estride=[%u,%u,%u,%u] base=%p",
                  static_cast<int>(r), rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
                  (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0), b[0],
                  rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0, e[0], rank > 1 ? e[1] : 0,
                  rank > 2 ? e[2] : 0, rank > 3 ? e[3] : 0, base);
        return 1;
    }
    return 0;
}

}  // namespace bpb

using namespace bpb;
#define S(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" {

const char* bpb_last_error(void) { return g_err; }
int bpb_abi_version(void) { return 2; }

int bpb_conv_bf16(const void* x, const void* w, void* y, float* stats, int n, int h, int wd, int cin, int cout,
                  int ksize, int stride, int bn_hint, void* stream) {
    return conv_bf16_tc(x, w, y, stats, n, h, wd, cin, cout, ksize, stride, bn_hint, nullptr, S(stream));
}
int bpb_conv_bn_bf16(const void* x, const void* w, void* y, float* stats, int n, int h, int wd, int cin, int cout,
                     int ksize, int stride, int bn_hint, void* bn_workspace, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var, int c_lo, float* scale,
                     float* shift, float* save_mean_invstd, void* stream) {
    BnFuse bn{bn_workspace, gamma, beta, running_mean, running_var, scale, shift, eps, momentum, c_lo, save_mean_invstd};
    return conv_bf16_tc(x, w, y, stats, n, h, wd, cin, cout, ksize, stride, bn_hint, &bn, S(stream));
}
int bpb_conv_bn_act_bf16(const void* x, const void* w, void* y, int n, int h, int wd, int cin, int cout, int ksize,
                         int stride, const float* scale, const float* shift, const void* res, int relu, void* stream) {
    BnAct act{scale, shift, res, relu};
    return conv_bf16_tc(x, w, y, nullptr, n, h, wd, cin, cout, ksize, stride, 0, nullptr, S(stream), &act);
}
int64_t bpb_conv_bn_workspace_bytes(int cout) { return static_cast<int64_t>(cout) * 2 * 8 + 16; }
int bpb_conv_num_mtiles(int n, int ho, int wo) { return conv_num_mtiles(n, ho, wo); }
int bpb_conv_ref_bf16(const void* x, const void* w, void* y, int n, int h, int wd, int cin, int cout, int ksize,
                      int stride, void* stream) {
    return conv_ref(x, w, y, n, h, wd, cin, cout, ksize, stride, S(stream));
}
int bpb_stem_conv(const float* img, const float* w, void* y, float* partials, int n, int h, int wd, int cout, int ksize,
                  int stride, void* stream) {
    return stem_conv(img, w, y, partials, n, h, wd, cout, ksize, stride, S(stream));
}
int bpb_stem_conv_tiles(int n, int ho, int wo) { return stem_conv_tiles(n, ho, wo); }
int bpb_colstats_bf16(const void* x, float* partials, int m, int c, void* stream) {
    return colstats_bf16(x, partials, m, c, nullptr, S(stream));
}
int bpb_colstats_tiles(int m) { return colstats_tiles(m); }
int bpb_bn_finalize(const float* partials, int tiles, int c, int c_lo, int64_t count, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                    float* scale, float* shift, float* save_mean_invstd, void* stream) {
    return bn_finalize(partials, tiles, c, c_lo, count, gamma, beta, eps, momentum, running_mean, running_var, scale,
                       shift, save_mean_invstd, S(stream));
}
int bpb_bn_eval_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     float eps, int c, float* scale, float* shift, void* stream) {
    return bn_eval_fold(gamma, beta, running_mean, running_var, eps, c, scale, shift, S(stream));
}
int bpb_bn_apply(const void* x, const float* scale, const float* shift, const void* res, const float* res_scale,
                 const float* res_shift, void* y, int64_t m, int c, int relu, void* stream) {
    return bn_apply(x, scale, shift, res, res_scale, res_shift, y, m, c, relu, S(stream));
}
int bpb_fuse_sum(const void* const* srcs, const float* const* scales, const float* const* shifts,
                 const int* shifts_log2, int nsrc, void* y, int n, int h, int wd, int c, int relu, void* stream) {
    return fuse_sum(srcs, scales, shifts, shifts_log2, nsrc, y, n, h, wd, c, relu, S(stream));
}
int bpb_upsample_concat(const void* const* srcs, const int* chans, const int* hs, const int* ws, int nsrc, void* y,
                        int n, int h, int wd, void* stream) {
    return upsample_concat(srcs, chans, hs, ws, nsrc, y, n, h, wd, S(stream));
}
int bpb_maxpool3x3s2(const void* x, void* y, int n, int h, int wd, int c, int pad, int ceil_mode, void* stream) {
    return maxpool3x3s2(x, y, n, h, wd, c, pad, ceil_mode, S(stream));
}
int bpb_instnorm_stats(const void* x, float* mean_rstd, int n, int hw, int c, int c_lo, int c_hi, float eps,
                       void* stream) {
    return instnorm_stats(x, mean_rstd, n, hw, c, c_lo, c_hi, eps, S(stream));
}
int bpb_ibn_apply(const void* x, const float* in_mean_rstd, const float* in_gamma, const float* in_beta,
                  const float* bn_scale, const float* bn_shift, void* y, int n, int hw, int c, int half, int relu,
                  void* stream) {
    return ibn_apply(x, in_mean_rstd, in_gamma, in_beta, bn_scale, bn_shift, y, n, hw, c, half, relu, S(stream));
}

int bpb_pixel_fold(const float* w, const float* b, const float* scale, const float* shift, int kp1, int c, float* wf,
                   float* bf, void* stream) {
    return pixel_fold(w, b, scale, shift, kp1, c, wf, bf, S(stream));
}
int bpb_pixel_classify(const void* feats, const float* wf, const float* bf, int n, int hw, int c, int kp1,
                       float* logits_nchw, float* probs, float* fg_mask, int* vis_bits, float* vis_cont,
                       const float* ext_masks, int mode, void* stream) {
    return pixel_classify(feats, wf, bf, n, hw, c, kp1, logits_nchw, probs, fg_mask, vis_bits, vis_cont, ext_masks, mode,
                          S(stream));
}
int bpb_part_pool(const void* feats, const float* probs, const float* fg_mask, int n, int hw, int c, int kp1,
                  float* pooled, void* stream) {
    return part_pool(feats, probs, fg_mask, n, hw, c, kp1, pooled, S(stream));
}
int bpb_linear_tf32(const float* a, int64_t lda, const float* w, const float* bias, float* out, int m, int n, int k,
                    int ldo, int a_mode, int a_parts, void* stream) {
    return linear_tf32_tc(a, lda, w, bias, out, m, n, k, ldo, a_mode, a_parts, S(stream));
}
int bpb_bn1d(const float* x, const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
             float* running_var, float* y, int m, int c, int relu, int training, void* stream) {
    return bn1d_train(x, gamma, beta, eps, momentum, running_mean, running_var, y, m, c, relu, training, S(stream));
}

int bpb_head_branch_stats(const void* x, int n, int hb, int wb, int cb, int h, int w, float* partial, int c_off,
                          int ctot, void* stream) {
    return head_branch_stats(x, n, hb, wb, cb, h, w, partial, c_off, ctot, S(stream));
}
int bpb_head_lowres_logits(const void* x, const float* wf, int64_t rows, int cb, int c_off, int ctot, int kp1,
                           float* out, void* stream) {
    return head_lowres_logits(x, wf, rows, cb, c_off, ctot, kp1, out, S(stream));
}
int bpb_head_pixel(const float* const* lowres_logits, float* const* mask_t, const int* hb, const int* wb, int nb,
                   const float* bf, int n, int h, int w, int kp1, float* logits, float* probs, float* fg_mask,
                   int* vis_bits, float* vis_cont, float* mask_sum, const float* ext_masks, int mode, void* stream) {
    return head_pixel(lowres_logits, mask_t, hb, wb, nb, bf, n, h, w, kp1, logits, probs, fg_mask, vis_bits, vis_cont,
                      mask_sum, ext_masks, mode, S(stream));
}
int bpb_head_branch_pool(const void* x, const float* mt, const float* mask_sum, int n, int q, int cb, int c_off,
                         int ctot, int kp1, int hw_full, float* pooled, void* stream) {
    return head_branch_pool(x, mt, mask_sum, n, q, cb, c_off, ctot, kp1, hw_full, pooled, S(stream));
}
int bpb_vis_unpack(const int* vis_bits, const float* vis_cont, int n, int kp1, uint8_t* out_b, float* out_f,
                   void* stream) {
    return vis_unpack(vis_bits, vis_cont, n, kp1, out_b, out_f, S(stream));
}
int bpb_weighted_sum(const float* const* terms, const float* weights, int n, float* out, void* stream) {
    return weighted_sum(terms, weights, n, out, S(stream));
}

int bpb_triplet_fwd(const float* emb_knd, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                    float margin, float* dist_nn, float* hp, float* hn, int* hp_idx, int* hn_idx, float* out4,
                    void* stream) {
    return triplet_fwd(emb_knd, labels, vis, vis_mode, n, k, d, margin, dist_nn, hp, hn, hp_idx, hn_idx, out4,
                       S(stream));
}
int bpb_triplet_fwd_rows(const float* emb, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                         float margin, int row_lo, int rows, int n_local, int sums_mode, float* dist_rows, float* hp,
                         float* hn, int* hp_idx, int* hn_idx, float* out4, void* stream) {
    return triplet_fwd_rows(emb, labels, vis, vis_mode, n, k, d, margin, row_lo, rows, n_local, sums_mode, dist_rows, hp,
                            hn, hp_idx, hn_idx, out4, S(stream));
}
int bpb_triplet_finish(const float* sums4, float* out4, void* stream) { return triplet_finish(sums4, out4, S(stream)); }
int bpb_ce_ls_fwd(const float* logits, int64_t ld, const int64_t* targets, const float* weights, int n, int c,
                  float eps, int acc_mode, float* scratch, float* out2, void* stream) {
    return ce_ls_fwd(logits, ld, targets, weights, n, c, eps, acc_mode, scratch, out2, S(stream));
}
int bpb_pixel_ce_fwd(const float* logits_nchw, const float* target_masks, int n, int kp1, int hf, int wf, int hm,
                     int wm, float label_smoothing, float* partial, float* out2, void* stream) {
    return pixel_ce_fwd(logits_nchw, target_masks, n, kp1, hf, wf, hm, wm, label_smoothing, partial, out2, S(stream));
}

int bpb_part_distance(const float* qf, const float* gf, const void* qv, const void* gv, int vis_mode, int nq, int ng,
                      int p, int d, int metric, int strat, int impl, float* dist, float* part_dist, float* scratch,
                      void* stream) {
    return part_distance(qf, gf, qv, gv, vis_mode, nq, ng, p, d, metric, strat, impl, dist, part_dist, scratch,
                         S(stream));
}
int64_t bpb_part_distance_scratch_floats(int nq, int ng, int p) {
    return 4 + static_cast<int64_t>(p + 1) * (static_cast<int64_t>(nq) + ng);
}
int bpb_rank_market1501(const float* dist, const int64_t* q_pids, const int64_t* g_pids, const int64_t* q_camids,
                        const int64_t* g_camids, int nq, int ng, int max_rank, double* ap64, float* ap32,
                        int* first_rank, float* cmc, double* map_out, int64_t* num_valid, float* map32_out,
                        int* overflow_flag, void* stream) {
    return rank_market1501(dist, q_pids, g_pids, q_camids, g_camids, nq, ng, max_rank, ap64, ap32, first_rank, cmc,
                           map_out, num_valid, map32_out, overflow_flag, S(stream));
}
int bpb_argsort_rows(const float* dist, int nq, int ng, int32_t* indices, void* stream) {
    return argsort_rows(dist, nq, ng, indices, S(stream));
}

/* ---- training backward ---- */
int bpb_conv_wgrad_bf16(const void* x, const void* dy, float* dw, int n, int h, int wd, int cin, int cout, int ksize,
                        int stride, int dy_up, void* stream) {
    return conv_wgrad_tc(x, dy, dw, n, h, wd, cin, cout, ksize, stride, dy_up, S(stream));
}
int bpb_relu_grad_sum(const void* const* srcs, int nsrc, const void* y, void* out, int64_t m, int c, void* stream) {
    return relu_grad_sum(srcs, nsrc, y, out, m, c, S(stream));
}
int64_t bpb_bn_bwd_workspace_bytes(int c) { return bn_bwd_workspace_bytes(c); }
int bpb_bn_bwd_reduce(const void* dy, const void* raw, const float* mean_invstd, const float* scale, const float* shift,
                      int n, int h, int w, int c, int up_log2, int relu_affine, int c_lo, void* workspace, float* coef,
                      float* dgamma, float* dbeta, void* stream) {
    return bn_bwd_reduce(dy, raw, mean_invstd, scale, shift, n, h, w, c, up_log2, relu_affine, c_lo, workspace, coef,
                         dgamma, dbeta, S(stream));
}
int bpb_bn_bwd_apply(const void* dy, const void* raw, const float* mean_invstd, const float* scale, const float* shift,
                     const float* coef, int n, int h, int w, int c, int up_log2, int relu_affine, int zero_insert,
                     void* draw, void* stream) {
    return bn_bwd_apply(dy, raw, mean_invstd, scale, shift, coef, n, h, w, c, up_log2, relu_affine, zero_insert, draw,
                        S(stream));
}
int bpb_stem_wgrad(const float* img, const void* draw, float* dw, int n, int h, int wd, int cout, int ksize, int stride,
                   void* stream) {
    return stem_wgrad(img, draw, dw, n, h, wd, cout, ksize, stride, S(stream));
}
int bpb_adam_pack(const void* segs, int nseg, int64_t total, float* p, const float* g, float* m, float* v, float lr,
                  float beta1, float beta2, float eps, float wd, int step, float grad_scale, int do_update,
                  void* stream) {
    return adam_pack(segs, nseg, total, p, g, m, v, lr, beta1, beta2, eps, wd, step, grad_scale, do_update, S(stream));
}
int bpb_memset_zero(void* p, int64_t bytes, void* stream) {
    BPB_CUDA(cudaMemsetAsync(p, 0, static_cast<size_t>(bytes), S(stream)));
    return 0;
}
int bpb_param_segment_bytes(void) { return static_cast<int>(sizeof(ParamSegment)); }
int bpb_transpose_pad(const float* in, int64_t ld_in, int rows, int cols, float* out, int64_t ld_out, int out_rows,
                      int out_cols, void* stream) {
    return transpose_pad(in, ld_in, rows, cols, out, ld_out, out_rows, out_cols, S(stream));
}

int bpb_ce_ls_bwd(const float* logits, int64_t ld, const int64_t* targets, const float* weights, int n, int c,
                  float eps, float loss_weight, float* dlogits, int64_t ldg, void* stream) {
    return ce_ls_bwd(logits, ld, targets, weights, n, c, eps, loss_weight, dlogits, ldg, S(stream));
}
int bpb_triplet_bwd(const float* emb_knd, const void* vis, int vis_mode, const float* hp, const float* hn,
                    const int* hp_idx, const int* hn_idx, const float* out4, int n, int k, int d, float margin,
                    float loss_weight, float* demb_knd, void* stream) {
    return triplet_bwd(emb_knd, vis, vis_mode, hp, hn, hp_idx, hn_idx, out4, n, k, d, margin, loss_weight, demb_knd,
                       S(stream));
}
int bpb_triplet_bwd_rows(const float* emb, const void* vis, int vis_mode, const float* hp, const float* hn,
                         const int* hp_idx, const int* hn_idx, const float* out4, int n, int k, int d, float margin,
                         float loss_weight, int row_lo, int rows, int n_local, float* demb, void* stream) {
    return triplet_bwd_rows(emb, vis, vis_mode, hp, hn, hp_idx, hn_idx, out4, n, k, d, margin, loss_weight, row_lo, rows,
                            n_local, demb, S(stream));
}
int bpb_bn1d_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* y, int64_t ld_y,
                 const float* gamma, float eps, int m, int c, float* dx, int64_t ld_dx, int accumulate, float* dgamma,
                 float* dbeta, void* stream) {
    return bn1d_bwd(dy, ld_dy, x, ld_x, y, ld_y, gamma, eps, m, c, dx, ld_dx, accumulate, dgamma, dbeta, S(stream));
}
int bpb_colsum_f32(const float* x, int64_t ld, int m, int c, float* out, void* stream) {
    return colsum_f32(x, ld, m, c, out, S(stream));
}
int bpb_head_pool_bwd_prep(float* g, const float* pooled, const float* mask_sum, int n, int c, int kp1, int hw_full,
                           float* dms, void* stream) {
    return head_pool_bwd_prep(g, pooled, mask_sum, n, c, kp1, hw_full, dms, S(stream));
}
int bpb_head_dmask(const void* x, const float* gs, int n, int q, int cb, int c_off, int ctot, int kp1, float* out,
                   void* stream) {
    return head_dmask(x, gs, n, q, cb, c_off, ctot, kp1, out, S(stream));
}
int bpb_head_pixel_bwd(const float* const* dmask_t, float* const* dlow, const int* hb, const int* wb, int nb,
                       const float* probs, const float* dms, int n, int h, int w, int kp1, const float* target_masks,
                       int hm, int wm, float label_smoothing, float pix_coef, const float* ext_dlogits, float* dbf,
                       void* stream) {
    return head_pixel_bwd(dmask_t, dlow, hb, wb, nb, probs, dms, n, h, w, kp1, target_masks, hm, wm, label_smoothing,
                          pix_coef, ext_dlogits, dbf, S(stream));
}
int bpb_head_dwf(const void* x, const float* dlow, int n, int q, int cb, int c_off, int ctot, int kp1, float* dwf,
                 void* stream) {
    return head_dwf(x, dlow, n, q, cb, c_off, ctot, kp1, dwf, S(stream));
}
int bpb_pixel_fold_bwd(const float* dwf, const float* dbf, const float* w, const float* gamma, const float* scale,
                       const float* shift, const float* mean_invstd, int kp1, int c, int64_t count, float* dw,
                       float* db, float* dgamma, float* dbeta, float* coefab, void* stream) {
    return pixel_fold_bwd(dwf, dbf, w, gamma, scale, shift, mean_invstd, kp1, c, count, dw, db, dgamma, dbeta, coefab,
                          S(stream));
}
int bpb_head_branch_bwd(const void* x, const float* mt, const float* gs, const float* dlow, const float* wf,
                        const float* coefab, int n, int hb, int wb, int cb, int c_off, int ctot, int h, int w, int kp1,
                        void* dx, void* stream) {
    return head_branch_bwd(x, mt, gs, dlow, wf, coefab, n, hb, wb, cb, c_off, ctot, h, w, kp1, dx, S(stream));
}

int bpb_resize_bilinear_f32(const float* in, float* out, int64_t planes, int hi, int wi, int ho, int wo, void* stream) {
    return resize_bilinear_f32(in, out, planes, hi, wi, ho, wo, S(stream));
}
int bpb_l2_normalize(const float* x, float* y, int64_t rows, int d, float eps, void* stream) {
    return l2_normalize(x, y, rows, d, eps, S(stream));
}


/* ---- input pipeline ---- */
int bpb_preprocess_images(const uint8_t* pixels, const int64_t* offsets, const int32_t* geom, int n, int height, int width,
                          int pad, const float* mean255, const float* inv_std255, const float* erase_fill, float* out,
                          void* stream) {
    return preprocess_images(pixels, offsets, geom, n, height, width, pad, mean255, inv_std255, erase_fill, out, S(stream));
}
int bpb_preprocess_masks(const float* fields, const int64_t* offsets, const int32_t* geom, int n, int height, int width,
                         int out_h, int out_w, int pad, const uint64_t* group_bits, int parts, int combine_sum,
                         int strategy, float threshold, float softmax_weight, float* out, void* stream) {
    return preprocess_masks(fields, offsets, geom, n, height, width, out_h, out_w, pad, group_bits, parts, combine_sum,
                            strategy, threshold, softmax_weight, out, S(stream));
}

/* ---- k-reciprocal re-ranking ---- */
int bpb_rerank_build(const float* qg, const float* qq, const float* gg, int nq, int ng, float* colmax, float* orig,
                     void* stream) {
    return rerank_build(qg, qq, gg, nq, ng, colmax, orig, S(stream));
}
int bpb_rerank_topk(const float* orig, int a, int k, int32_t* rank, void* stream) {
    return rerank_topk(orig, a, k, rank, S(stream));
}
int bpb_rerank_row_capacity(int k1) { return rerank_row_capacity(k1); }
int bpb_rerank_kreciprocal(const float* orig, const int32_t* rank, int a, int rank_cols, int k1, int32_t* vidx, float* vval,
                           int32_t* vcnt, void* stream) {
    return rerank_kreciprocal(orig, rank, a, rank_cols, k1, vidx, vval, vcnt, S(stream));
}
int bpb_rerank_qexpand(const int32_t* rank, int rank_cols, int k2, const int32_t* vidx, const float* vval,
                       const int32_t* vcnt, int k1, int a, int32_t* qidx, float* qval, int32_t* qcnt, void* stream) {
    return rerank_qexpand(rank, rank_cols, k2, vidx, vval, vcnt, k1, a, qidx, qval, qcnt, S(stream));
}
int bpb_rerank_csc_count(const int32_t* qidx, const int32_t* qcnt, int qcap, int a, int32_t* colcnt, void* stream) {
    return rerank_csc_count(qidx, qcnt, qcap, a, colcnt, S(stream));
}
int bpb_rerank_csc_fill(const int32_t* qidx, const float* qval, const int32_t* qcnt, int qcap, int a,
                        const int32_t* col_ptr, int32_t* cursor, int32_t* crow, float* cval, void* stream) {
    return rerank_csc_fill(qidx, qval, qcnt, qcap, a, col_ptr, cursor, crow, cval, S(stream));
}
int bpb_rerank_jaccard_blocks(int nq, int a) { return rerank_jaccard_blocks(nq, a); }
int bpb_rerank_jaccard(const float* orig, const int32_t* qidx, const float* qval, const int32_t* qcnt, int qcap,
                       const int32_t* col_ptr, const int32_t* crow, const float* cval, int nq, int a,
                       float one_minus_lambda, float lambda, float* scratch, float* out, void* stream) {
    return rerank_jaccard(orig, qidx, qval, qcnt, qcap, col_ptr, crow, cval, nq, a, one_minus_lambda, lambda, scratch, out,
                          S(stream));
}

}  // extern "C"
