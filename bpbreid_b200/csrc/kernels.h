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
};
int conv_bf16_tc(const void* x, const void* w, void* y, float* stats, int n, int h, int wd, int cin, int cout,
                 int ksize, int stride, int bn_hint, const BnFuse* bn, cudaStream_t stream);
int linear_tf32_tc(const float* a, int64_t lda, const float* w, const float* bias, float* out, int m, int n, int k,
                   int ldo, int a_mode, int a_parts, cudaStream_t stream);
int conv3x3_halo_bf16(const void* x, const void* w, void* y, int n, int h, int wd, int cin, int cout, const BnFuse* bn,
                      cudaStream_t stream);
int conv_num_mtiles(int n, int ho, int wo);

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
                float* shift, cudaStream_t stream);
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
                   float* logits_nchw, float* probs, float* fg_mask, int* vis_bits, float* vis_cont,
                   cudaStream_t stream);
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
               int* vis_bits, float* vis_cont, float* mask_sum, cudaStream_t stream);
int head_branch_pool(const void* x, const float* mt, const float* mask_sum, int n, int q, int cb, int c_off, int ctot,
                     int kp1, int hw_full, float* pooled, cudaStream_t stream);

// ---- loss.cu ----------------------------------------------------------------------------------
int triplet_fwd(const float* emb_knd, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                float margin, float* dist_nn, float* hp, float* hn, int* hp_idx, int* hn_idx, float* out4,
                cudaStream_t stream);
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

}  // namespace bpb
