// Internal C++ entry points behind the C ABI (include/bpbreid_sm100.h).  Every function enqueues work on
// `stream`, never synchronises, returns 0 on success (bpb::set_error + non-zero otherwise).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bpb {

// ---- tc_gemm.cu -------------------------------------------------------------------------------
// Train-mode BatchNorm fused into the convolution launch (last CTA finalizes): see GemmParams in tc_gemm.cu.
struct BnFuse {
    void* workspace;            // 2*cout doubles + 1 uint32 (zero-initialised once; the kernel leaves it zeroed)
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;
    float* scale; float* shift;
    float eps, momentum;
    int c_lo;                   // parameters cover channels [c_lo, cout) (IBN: first half is instance-normalised)
    float* save_mean_invstd;    // optional [2][cout]: batch mean and 1/sqrt(var+eps) for the backward pass
};
// Inference epilogue: y = act(conv * scale + shift [+ res]) written by the conv launch itself (no bn_apply pass).
struct BnAct {
    const float* scale; const float* shift;   // folded running statistics (bn_eval_fold)
    const void* res;                           // optional bf16 residual shaped like y
    int relu;
};
int conv_bf16_tc(const void* x, const void* w, void* y, float* stats, int n, int h, int wd, int cin, int cout,
                 int ksize, int stride, int bn_hint, const BnFuse* bn, cudaStream_t stream, const BnAct* act = nullptr);
int linear_tf32_tc(const float* a, int64_t lda, const float* w, const float* bias, float* out, int m, int n, int k,
                   int ldo, int a_mode, int a_parts, cudaStream_t stream);
int conv_num_mtiles(int n, int ho, int wo);
bool conv_tiling(int ho, int wo, int* bh_out, int* bnimg_out);

// ---- nn_kernels.cu ----------------------------------------------------------------------------
int conv_ref(const void* x, const void* w, void* y, int n, int h, int wd, int cin, int cout, int ksize, int stride,
             cudaStream_t stream);
int stem_conv(const float* img, const float* w, void* y, float* partials, int n, int h, int wd, int cout, int ksize,
              int stride, cudaStream_t stream);
int stem_conv_tiles(int n, int ho, int wo);
int colstats_bf16(const void* x, float* partials, int m, int c, int* tiles_out, cudaStream_t stream);
int colstats_tiles(int m);
int bn_finalize(const float* partials, int tiles, int c, int c_lo, int64_t count, const float* gamma,
                const float* beta, float eps, float momentum, float* running_mean, float* running_var, float* scale,
                float* shift, float* save_mean_invstd, cudaStream_t stream);
int bn_eval_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                 float eps, int c, float* scale, float* shift, cudaStream_t stream);
int bn_apply(const void* x, const float* scale, const float* shift, const void* res, const float* res_scale,
             const float* res_shift, void* y, int64_t m, int c, int relu, cudaStream_t stream);
int fuse_sum(const void* const* srcs, const float* const* scales, const float* const* shifts, const int* shifts_log2,
             int nsrc, void* y, int n, int h, int wd, int c, int relu, cudaStream_t stream);
int upsample_concat(const void* const* srcs, const int* chans, const int* hs, const int* ws, int nsrc, void* y, int n,
                    int h, int wd, cudaStream_t stream);
int maxpool3x3s2(const void* x, void* y, int n, int h, int wd, int c, int pad, int ceil_mode, cudaStream_t stream);
int instnorm_stats(const void* x, float* mean_rstd, int n, int hw, int c, int c_lo, int c_hi, float eps,
                   cudaStream_t stream);
int ibn_apply(const void* x, const float* in_mean_rstd, const float* in_gamma, const float* in_beta,
              const float* bn_scale, const float* bn_shift, void* y, int n, int hw, int c, int half, int relu,
              cudaStream_t stream);

// ---- head.cu ----------------------------------------------------------------------------------
int pixel_fold(const float* w, const float* b, const float* scale, const float* shift, int kp1, int c, float* wf,
               float* bf, cudaStream_t stream);
int pixel_classify(const void* feats, const float* wf, const float* bf, int n, int hw, int c, int kp1,
                   float* logits_nchw, float* probs, float* fg_mask, int* vis_bits, float* vis_cont, const float* ext,
                   int mode, cudaStream_t stream);
int part_pool(const void* feats, const float* probs, const float* fg_mask, int n, int hw, int c, int kp1,
              float* pooled, cudaStream_t stream);
int bn1d_train(const float* x, const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
               float* running_var, float* y, int m, int c, int relu, int training, cudaStream_t stream);

int vis_unpack(const int* vis_bits, const float* vis_cont, int n, int kp1, uint8_t* out_b, float* out_f,
               cudaStream_t stream);

// ---- head_lowres.cu ---------------------------------------------------------------------------
int head_branch_stats(const void* x, int n, int hb, int wb, int cb, int h, int w, float* partial, int c_off, int ctot,
                      cudaStream_t stream);
int head_lowres_logits(const void* x, const float* wf, int64_t rows, int cb, int c_off, int ctot, int kp1, float* out,
                       cudaStream_t stream);
int head_pixel(const float* const* lowres_logits, float* const* mask_t, const int* hb, const int* wb, int nb,
               const float* bf, int n, int h, int w, int kp1, float* logits, float* probs, float* fg_mask,
               int* vis_bits, float* vis_cont, float* mask_sum, const float* ext, int mode, cudaStream_t stream);
int head_branch_pool(const void* x, const float* mt, const float* mask_sum, int n, int q, int cb, int c_off, int ctot,
                     int kp1, int hw_full, float* pooled, cudaStream_t stream);

// ---- loss.cu ----------------------------------------------------------------------------------
int triplet_fwd(const float* emb_knd, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                float margin, float* dist_nn, float* hp, float* hn, int* hp_idx, int* hn_idx, float* out4,
                cudaStream_t stream);
int triplet_fwd_rows(const float* emb, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                     float margin, int row_lo, int rows, int n_local, int sums_mode, float* dist_rows, float* hp,
                     float* hn, int* hp_idx, int* hn_idx, float* out4, cudaStream_t stream);
int triplet_finish(const float* sums4, float* out4, cudaStream_t stream);
int ce_ls_fwd(const float* logits, int64_t ld, const int64_t* targets, const float* weights, int n, int c, float eps,
              int acc_mode, float* row_loss, float* out2, cudaStream_t stream);
int pixel_ce_fwd(const float* logits_nchw, const float* target_masks, int n, int kp1, int hf, int wf, int hm, int wm,
                 float label_smoothing, float* partial, float* out2, cudaStream_t stream);

int weighted_sum(const float* const* terms, const float* weights, int n, float* out, cudaStream_t stream);

// ---- distance.cu ------------------------------------------------------------------------------
int part_distance(const float* qf, const float* gf, const void* qv, const void* gv, int vis_mode, int nq, int ng,
                  int p, int d, int metric, int strat, int impl, float* dist, float* part_dist, float* scratch,
                  cudaStream_t stream);

// ---- rank.cu ----------------------------------------------------------------------------------
int rank_market1501(const float* dist, const int64_t* q_pids, const int64_t* g_pids, const int64_t* q_camids,
                    const int64_t* g_camids, int nq, int ng, int max_rank, double* ap64, float* ap32,
                    int* first_rank, float* cmc, double* map_out, int64_t* num_valid, float* map32_out,
                    int* overflow_flag, cudaStream_t stream);
int argsort_rows(const float* dist, int nq, int ng, int32_t* indices, cudaStream_t stream);

// ---- tc_wgrad.cu ------------------------------------------------------------------------------
int conv_wgrad_tc(const void* x, const void* dy, float* dw, int n, int h, int wd, int cin, int cout, int ksize,
                  int stride, int dy_up, cudaStream_t stream);

// ---- bwd_kernels.cu ---------------------------------------------------------------------------
// one parameter tensor inside the flat fp32 parameter buffer (device table consumed by adam_pack)
struct ParamSegment {
    int64_t off, numel;
    int kind;                   // 0: plain tensor, 1: tensor-core conv weight [cout][cin][k][k]
    int trainable;
    int cout, cin, k, pad_;
    uint16_t* wf;               // bf16 OHWI copy (forward operand) or null
    uint16_t* wd;               // bf16 [cin][flipped taps][cout] copy (dgrad operand) or null
};
int relu_grad_sum(const void* const* srcs, int nsrc, const void* y, void* out, int64_t m, int c, cudaStream_t stream);
int bn_bwd_reduce(const void* dy, const void* raw, const float* mean_invstd, const float* scale, const float* shift,
                  int n, int h, int w, int c, int up_log2, int relu_affine, int c_lo, void* workspace, float* coef,
                  float* dgamma, float* dbeta, cudaStream_t stream);
int64_t bn_bwd_workspace_bytes(int c);
int bn_bwd_apply(const void* dy, const void* raw, const float* mean_invstd, const float* scale, const float* shift,
                 const float* coef, int n, int h, int w, int c, int up_log2, int relu_affine, int zero_insert,
                 void* draw, cudaStream_t stream);
int stem_wgrad(const float* img, const void* draw, float* dw, int n, int h, int wd, int cout, int ksize, int stride,
               cudaStream_t stream);
int adam_pack(const void* segs, int nseg, int64_t total, float* p, const float* g, float* m, float* v, float lr,
              float beta1, float beta2, float eps, float wd, int step, float grad_scale, int do_update,
              cudaStream_t stream);
int transpose_pad(const float* in, int64_t ld_in, int rows, int cols, float* out, int64_t ld_out, int out_rows,
                  int out_cols, cudaStream_t stream);

// ---- loss_bwd.cu ------------------------------------------------------------------------------
int ce_ls_bwd(const float* logits, int64_t ld, const int64_t* targets, const float* weights, int n, int c, float eps,
              float loss_weight, float* dlogits, int64_t ldg, cudaStream_t stream);
int triplet_bwd(const float* emb_knd, const void* vis, int vis_mode, const float* hp, const float* hn,
                const int* hp_idx, const int* hn_idx, const float* out4, int n, int k, int d, float margin,
                float loss_weight, float* demb_knd, cudaStream_t stream);
int triplet_bwd_rows(const float* emb, const void* vis, int vis_mode, const float* hp, const float* hn,
                     const int* hp_idx, const int* hn_idx, const float* out4, int n, int k, int d, float margin,
                     float loss_weight, int row_lo, int rows, int n_local, float* demb, cudaStream_t stream);
int bn1d_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* y, int64_t ld_y,
             const float* gamma, float eps, int m, int c, float* dx, int64_t ld_dx, int accumulate, float* dgamma,
             float* dbeta, cudaStream_t stream);
int colsum_f32(const float* x, int64_t ld, int m, int c, float* out, cudaStream_t stream);

// ---- head_bwd.cu ------------------------------------------------------------------------------
int head_pool_bwd_prep(float* g, const float* pooled, const float* mask_sum, int n, int c, int kp1, int hw_full,
                       float* dms, cudaStream_t stream);
int head_dmask(const void* x, const float* gs, int n, int q, int cb, int c_off, int ctot, int kp1, float* out,
               cudaStream_t stream);
int head_pixel_bwd(const float* const* dmask_t, float* const* dlow, const int* hb, const int* wb, int nb,
                   const float* probs, const float* dms, int n, int h, int w, int kp1, const float* target_masks,
                   int hm, int wm, float label_smoothing, float pix_coef, const float* ext_dlogits, float* dbf,
                   cudaStream_t stream);
int head_dwf(const void* x, const float* dlow, int n, int q, int cb, int c_off, int ctot, int kp1, float* dwf,
             cudaStream_t stream);
int pixel_fold_bwd(const float* dwf, const float* dbf, const float* w, const float* gamma, const float* scale,
                   const float* shift, const float* mean_invstd, int kp1, int c, int64_t count, float* dw, float* db,
                   float* dgamma, float* dbeta, float* coefab, cudaStream_t stream);
int head_branch_bwd(const void* x, const float* mt, const float* gs, const float* dlow, const float* wf,
                    const float* coefab, int n, int hb, int wb, int cb, int c_off, int ctot, int h, int w, int kp1,
                    void* dx, cudaStream_t stream);

// ---- eval_kernels.cu --------------------------------------------------------------------------
int l2_normalize(const float* x, float* y, int64_t rows, int d, float eps, cudaStream_t stream);

// input pipeline (input_pipeline.cu); mean255 / inv_std255 / erase_fill / group_bits are HOST arrays
int preprocess_images(const uint8_t* pixels, const int64_t* offsets, const int32_t* geom, int n, int height, int width,
                      int pad, const float* mean255, const float* inv_std255, const float* erase_fill, float* out,
                      cudaStream_t stream);
int preprocess_masks(const float* fields, const int64_t* offsets, const int32_t* geom, int n, int height, int width,
                     int out_h, int out_w, int pad, const uint64_t* group_bits, int parts, int combine_sum, int strategy,
                     float threshold, float softmax_weight, float* out, cudaStream_t stream);

// k-reciprocal re-ranking (rerank.cu)
int rerank_build(const float* qg, const float* qq, const float* gg, int nq, int ng, float* cmax, float* orig,
                 cudaStream_t stream);
int rerank_topk(const float* orig, int a, int k, int* rank, cudaStream_t stream);
int rerank_row_capacity(int k1);
int rerank_kreciprocal(const float* orig, const int* rank, int a, int rk, int k1, int* vidx, float* vval, int* vcnt,
                       cudaStream_t stream);
int rerank_qexpand(const int* rank, int rk, int k2, const int* vidx, const float* vval, const int* vcnt, int k1, int a,
                   int* qidx, float* qval, int* qcnt, cudaStream_t stream);
int rerank_csc_count(const int* qidx, const int* qcnt, int qcap, int a, int* colcnt, cudaStream_t stream);
int rerank_csc_fill(const int* qidx, const float* qval, const int* qcnt, int qcap, int a, const int* col_ptr, int* cursor,
                    int* crow, float* cval, cudaStream_t stream);
int rerank_jaccard_blocks(int nq, int a);
int rerank_jaccard(const float* orig, const int* qidx, const float* qval, const int* qcnt, int qcap, const int* col_ptr,
                   const int* crow, const float* cval, int nq, int a, float one_minus_lambda, float lambda, float* scratch,
                   float* out, cudaStream_t stream);
int resize_bilinear_f32(const float* in, float* out, int64_t planes, int hi, int wi, int ho, int wo, cudaStream_t stream);

}  // namespace bpb
