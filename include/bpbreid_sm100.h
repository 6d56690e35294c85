/* bpbreid_sm100.h -- C ABI of libbpbreid_sm100.so (B200 / sm_100a).
 *
 * Drop-in boundary for the BPBReID hot path.  The reference is pure Python: its "FFI" for
 * this path is the set of tensor-level call sites listed beside each entry point below, so every function
 * takes raw DEVICE pointers + explicit shapes + a cudaStream_t (passed as void*), never a torch type.
 *
 * Conventions
 *   - the caller owns every buffer (torch tensors in practice, so the caching allocator / CUDA graphs work);
 *   - work is enqueued on `stream`; nothing synchronises; functions are re-entrant;
 *   - return value 0 = ok, non-zero = error, message via bpb_last_error() (thread-local);
 *   - activations are NHWC bf16 ("bf16*" below = uint16_t storage), statistics / heads / losses are fp32;
 *   - visibility vectors are uint8 (torch.bool storage) when vis_mode == 1, float when vis_mode == 2.
 */
#ifndef BPBREID_SM100_H_
#define BPBREID_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* bpb_last_error(void);
int bpb_abi_version(void);

/* ---- backbone: replaces nn.Conv2d / nn.BatchNorm2d / ReLU / residual call sites --------------------------- */

/* NHWC bf16 convolution on the tcgen05 tensor cores, ksize in {1,3}, stride in {1,2}, pad = ksize/2, no bias.
 * Writes the raw (pre-BN) output as bf16 [N*Ho*Wo, cout] and per-128-pixel-tile column sums / sums of squares
 * `stats` [m_tiles, 2, cout] (m_tiles = bpb_conv_num_mtiles).  w is OHWI bf16 [cout, ksize*ksize*cin].
 * Replaces: torchreid/models/hrnet.py:60-64,80-96,116-137,222-257,459-483; torchreid/models/resnet.py:43-63,128-154. */
int bpb_conv_bf16(const void* x, const void* w, void* y, float* stats, int n, int h, int wd, int cin, int cout,
                  int ksize, int stride, int bn_hint, void* stream);
/* bn_hint: 0 = automatic tile width; 32/64/128 forces the N tile; bit 8 (0x100) forces the one-tap-per-stage schedule
 * for 3x3 stride-1 layers.  The bits only select between numerically equivalent schedules (test hooks). */
int bpb_conv_num_mtiles(int n, int ho, int wo);
/* Same convolution with the train-mode BatchNorm that follows it fused into the launch (conv + nn.BatchNorm2d call
 * pairs, e.g. hrnet.py:82-83): every CTA adds its per-channel sums to `bn_workspace` (fp64 atomics) and the last CTA
 * turns them into scale/shift [cout] (biased batch variance), updates running_mean/var like nn.BatchNorm2d and
 * leaves the workspace zeroed.  bn_workspace: bpb_conv_bn_workspace_bytes(cout) bytes, zero-initialised ONCE by the
 * caller.  gamma/beta/running_* cover channels [c_lo, cout).  `stats` (per-tile partials) is optional. */
int bpb_conv_bn_bf16(const void* x, const void* w, void* y, float* stats, int n, int h, int wd, int cin, int cout,
                     int ksize, int stride, int bn_hint, void* bn_workspace, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var, int c_lo, float* scale,
                     float* shift, float* save_mean_invstd, void* stream);
/* save_mean_invstd: optional [2][cout] fp32 -- batch mean and 1/sqrt(var + eps), kept for bpb_bn_bwd_*. */
int64_t bpb_conv_bn_workspace_bytes(int cout);
/* Inference: the same convolution with the (running-statistics) BatchNorm, the optional residual add and the ReLU in
 * its epilogue: y = act(conv(x) * scale + shift [+ res]) as bf16 -- one launch instead of conv + bpb_bn_apply.
 * scale / shift [cout] from bpb_bn_eval_fold; res NULL or a bf16 tensor shaped like y.  The BatchNorm is applied to the
 * bf16-rounded convolution result, so the output is bit-identical to the two-launch path.
 * Replaces conv -> bn -> (+ residual) -> relu call chains in eval mode (hrnet.py:80-96,116-137; resnet.py:128-154). */
int bpb_conv_bn_act_bf16(const void* x, const void* w, void* y, int n, int h, int wd, int cin, int cout, int ksize,
                         int stride, const float* scale, const float* shift, const void* res, int relu, void* stream);
/* SIMT reference convolution, same operands (validation / shapes the tensor-core tiler rejects). */
int bpb_conv_ref_bf16(const void* x, const void* w, void* y, int n, int h, int wd, int cin, int cout, int ksize,
                      int stride, void* stream);
/* Stem: NCHW fp32 image (3 ch) -> NHWC bf16 raw output; w is the PyTorch [cout,3,k,k] fp32 weight.
 * Replaces hrnet.py:533 (3x3 s2) and resnet.py:199-201 / resnet_fastreid.py:246 (7x7 s2). */
int bpb_stem_conv(const float* img, const float* w, void* y, float* partials, int n, int h, int wd, int cout, int ksize,
                  int stride, void* stream);
/* rows of `partials` ([rows, 2, 64] per-CTA column sums / sums of squares, may be NULL) written by bpb_stem_conv */
int bpb_stem_conv_tiles(int n, int ho, int wo);
/* Column statistics of a bf16 matrix [m, c] -> partials [bpb_colstats_tiles(m), 2, c]. */
int bpb_colstats_bf16(const void* x, float* partials, int m, int c, void* stream);
int bpb_colstats_tiles(int m);
/* partials -> BN scale/shift (train-mode batch statistics, biased variance; running stats updated with momentum and
 * the unbiased variance exactly like nn.BatchNorm2d; pass NULL to skip).  Parameters cover channels [c_lo, c).
 * Replaces the nn.BatchNorm2d call sites of hrnet.py:73,77,106-111 etc. in training mode. */
int bpb_bn_finalize(const float* partials, int tiles, int c, int c_lo, int64_t count, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                    float* scale, float* shift, float* save_mean_invstd, void* stream);
/* eval mode: scale/shift from the running statistics. */
int bpb_bn_eval_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     float eps, int c, float* scale, float* shift, void* stream);
/* y = act(x*scale + shift [+ res | + res*res_scale + res_shift]) on bf16 [m, c]  (hrnet.py:84-96,123-137). */
int bpb_bn_apply(const void* x, const float* scale, const float* shift, const void* res, const float* res_scale,
                 const float* res_shift, void* y, int64_t m, int c, int relu, void* stream);
/* Cross-resolution fuse sum (hrnet.py:269-277): y = relu(sum_j up_nearest_{2^sh_j}(src_j*scale_j+shift_j)). */
int bpb_fuse_sum(const void* const* srcs, const float* const* scales, const float* const* shifts,
                 const int* shifts_log2, int nsrc, void* y, int n, int h, int wd, int c, int relu, void* stream);
/* Bilinear(align_corners=True) upsample + channel concat of the branch maps (hrnet.py:568-573). */
int bpb_upsample_concat(const void* const* srcs, const int* chans, const int* hs, const int* ws, int nsrc, void* y,
                        int n, int h, int wd, void* stream);
/* 3x3 stride-2 max pool (resnet.py:216 pad 1; resnet_fastreid.py:251 pad 0 + ceil_mode). */
int bpb_maxpool3x3s2(const void* x, void* y, int n, int h, int wd, int c, int pad, int ceil_mode, void* stream);
/* IBN (resnet_fastreid.py:108-122): per-image statistics of channels [c_lo,c_hi) and the fused IN|BN apply. */
int bpb_instnorm_stats(const void* x, float* mean_rstd, int n, int hw, int c, int c_lo, int c_hi, float eps,
                       void* stream);
int bpb_ibn_apply(const void* x, const float* in_mean_rstd, const float* in_gamma, const float* in_beta,
                  const float* bn_scale, const float* bn_shift, void* y, int n, int hw, int c, int half, int relu,
                  void* stream);

/* ---- head: replaces torchreid/models/bpbreid.py:147-219 -------------------------------------------------- */

/* Fold BN2d into the 1x1 pixel classifier (bpbreid.py:383-385). */
int bpb_pixel_fold(const float* w, const float* b, const float* scale, const float* shift, int kp1, int c, float* wf,
                   float* bf, void* stream);
/* Pixel classifier + softmax + masks + visibility (bpbreid.py:147-148,157-192).  feats bf16 [n, hw, c].
 * logits/probs fp32 [n, K+1, hw]; fg_mask [n, hw]; vis_bits[n] bit k set iff class k is the argmax somewhere;
 * vis_cont [n, K+1] = amax of the probabilities.
 * ext_masks (optional): external part masks already resized to the feature map, fp32 [n, K+1, hw]
 * (bpb_resize_bilinear_f32); mode bit 0: they ARE the attention (learnable_attention_enabled False, bpbreid.py:149-154;
 * wf / bf / logits unused), bit 1: test_use_target_segmentation 'hard' (:160-167), bit 2: 'soft' (:169-174).
 * With mode != 0 `probs` holds the masks used for pooling: row 0 background mask, rows 1..K part masks. */
int bpb_pixel_classify(const void* feats, const float* wf, const float* bf, int n, int hw, int c, int kp1,
                       float* logits_nchw, float* probs, float* fg_mask, int* vis_bits, float* vis_cont,
                       const float* ext_masks, int mode, void* stream);
/* Global / foreground / background / K part pooling in one pass (bpbreid.py:195-202,458-503).
 * pooled fp32 [K+3, n, c] with rows {0: global, 1: fg, 2: bg, 3..: parts}. */
int bpb_part_pool(const void* feats, const float* probs, const float* fg_mask, int n, int hw, int c, int kp1,
                  float* pooled, void* stream);
/* out[m, n] = A[m, k] W[n, k]^T + bias on tcgen05 (tf32) -- nn.Linear of bpbreid.py:329-334,409.
 * a_mode 1: A is [a_parts, m, k/a_parts] read as the concat over parts (bpbreid.py:212). */
int bpb_linear_tf32(const float* a, int64_t lda, const float* w, const float* bias, float* out, int m, int n, int k,
                    int ldo, int a_mode, int a_parts, void* stream);
/* nn.BatchNorm1d (+ReLU) on fp32 [m, c] (bpbreid.py:335-336,413). */
int bpb_bn1d(const float* x, const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
             float* running_var, float* y, int m, int c, int relu, int training, void* stream);

/* ---- head on the UN-concatenated HRNet branch maps (linear algebra identity, csrc/head_lowres.cu): replaces the same
 *      reference lines as the four entry points above plus hrnet.py:568-573, without materialising [N,1920,H/4,W/4] ---- */
/* per-image partial BN2d statistics of the bilinearly-upsampled branch -> partial[n][2][ctot] at channel c_off
 * (reduce with bpb_bn_finalize(tiles = n)). */
int bpb_head_branch_stats(const void* x, int n, int hb, int wb, int cb, int h, int w, float* partial, int c_off,
                          int ctot, void* stream);
/* out[row][k] = sum_c wf[k][c_off+c] x[row][c] : the folded pixel classifier applied at the branch's own resolution. */
int bpb_head_lowres_logits(const void* x, const float* wf, int64_t rows, int cb, int c_off, int ctot, int kp1,
                           float* out, void* stream);
/* full-resolution logits = bf + sum_b bilinear(lowres_logits[b]) -> softmax -> masks / visibility (same outputs as
 * bpb_pixel_classify) + mask_sum[n][K+3] and the transposed-upsampled masks mask_t[b][n][K+3][hb*wb]
 * (rows: global, foreground, background, parts).  ext_masks / mode as for bpb_pixel_classify (mode bit 0:
 * lowres_logits may be NULL). */
int bpb_head_pixel(const float* const* lowres_logits, float* const* mask_t, const int* hb, const int* wb, int nb,
                   const float* bf, int n, int h, int w, int kp1, float* logits, float* probs, float* fg_mask,
                   int* vis_bits, float* vis_cont, float* mask_sum, const float* ext_masks, int mode, void* stream);
/* pooled[j][n][c_off+c] = norm_j * sum_q mask_t[n][j][q] x[n][q][c]  (same pooled layout as bpb_part_pool). */
int bpb_head_branch_pool(const void* x, const float* mt, const float* mask_sum, int n, int q, int cb, int c_off,
                         int ctot, int kp1, int hw_full, float* pooled, void* stream);

/* Visibility scores (bpbreid.py:182-192): out_b / out_f [n, K+2] = {background, part_1..K, foreground}. */
int bpb_vis_unpack(const int* vis_bits, const float* vis_cont, int n, int kp1, uint8_t* out_b, float* out_f,
                   void* stream);

/* ---- GiLt loss: replaces torchreid/losses/{part_averaged_triplet_loss,cross_entropy_loss,
 *      body_part_attention_loss}.py ------------------------------------------------------------------------ */

/* emb [K, n, d] fp32 (part-major), labels int64 [n], vis [n, K] (uint8 / float) or NULL.
 * Outputs: dist [n, n], hardest positive / negative value + index per anchor, out4 = {loss, trivial_ratio,
 * valid_ratio, n_valid} (loss = NaN when n_valid == 0: the caller warns and returns None like
 * part_averaged_triplet_loss.py:161-163).  margin <= 0 selects the soft-margin variant. */
int bpb_triplet_fwd(const float* emb_knd, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                    float margin, float* dist_nn, float* hp, float* hn, int* hp_idx, int* hn_idx, float* out4,
                    void* stream);
/* Sharded variant for data-parallel training (one process per GPU): `emb` is the all-gathered tensor in
 * [shards][K][n_local][D] layout (shards * n_local = n), labels / vis cover all n samples, and only the anchor rows
 * [row_lo, row_lo + rows) are computed: dist_rows [rows, n], hp/hn/hp_idx/hn_idx [rows] (indices are global).
 * sums_mode 1: out4 = {sum loss, sum trivial, n_valid, rows} partial sums to all-reduce, then bpb_triplet_finish.
 * The mined pairs equal those of the reference's gather-on-GPU-0 mining (part_based_engine.py:82-95). */
int bpb_triplet_fwd_rows(const float* emb, const int64_t* labels, const void* vis, int vis_mode, int n, int k, int d,
                         float margin, int row_lo, int rows, int n_local, int sums_mode, float* dist_rows, float* hp,
                         float* hn, int* hp_idx, int* hn_idx, float* out4, void* stream);
int bpb_triplet_finish(const float* sums4, float* out4, void* stream);
/* Label-smoothed CE (cross_entropy_loss.py:34-56); scratch has 2n floats; out2 = {loss, top-1 accuracy}.
 * weights NULL: mean over rows; else sum_i CE_i w_i / |w|_1.  acc_mode 1: accuracy over rows with w != 0. */
int bpb_ce_ls_fwd(const float* logits, int64_t ld, const int64_t* targets, const float* weights, int n, int c,
                  float eps, int acc_mode, float* scratch, float* out2, void* stream);
/* Body part attention loss (part_based_engine.py:118-124 + body_part_attention_loss.py:31-52);
 * partial has 512 floats; out2 = {loss, pixel accuracy}. */
int bpb_pixel_ce_fwd(const float* logits_nchw, const float* target_masks, int n, int kp1, int hf, int wf, int hm,
                     int wm, float label_smoothing, float* partial, float* out2, void* stream);

/* out[0] = sum_i weights[i] * terms[i][0]  (GiLt_loss.py:92; `terms` is a host array of device pointers). */
int bpb_weighted_sum(const float* const* terms, const float* weights, int n, float* out, void* stream);

/* ---- evaluation: replaces torchreid/metrics/distance.py:87-247 and torchreid/metrics/rank.py:97-159 /
 *      torchreid/metrics/rank_cylib/rank_cy.pyx:154-249 ----------------------------------------------------- */

/* qf [nq, p, d], gf [ng, p, d] fp32.  metric 0 euclidean / 1 cosine; strat 0 mean / 1 max; impl 0 tcgen05 (tf32),
 * 1 SIMT fp32.  dist [nq, ng]; part_dist [p, nq, ng] or NULL; scratch >= bpb_part_distance_scratch_floats().
 * impl 0 numerics: tf32 products with fp32 accumulation, norms taken from the tf32-truncated inputs, sqrt.approx
 * (absolute error <= ~2e-3 on unit-norm features, tests/test_gpu_path.py); impl 1 is plain fp32 with IEEE sqrt. */
int bpb_part_distance(const float* qf, const float* gf, const void* qv, const void* gv, int vis_mode, int nq, int ng,
                      int p, int d, int metric, int strat, int impl, float* dist, float* part_dist, float* scratch,
                      void* stream);
int64_t bpb_part_distance_scratch_floats(int nq, int ng, int p);
/* market1501 CMC/mAP.  Per-query outputs ap64/ap32/first_rank [nq]; cmc [max_rank]; map_out (fp64 accumulation,
 * rank.py semantics), map32_out (fp32 accumulation, rank_cy.pyx semantics); num_valid; overflow_flag set when a
 * query has more matches than the kernel's list can hold. */
int bpb_rank_market1501(const float* dist, const int64_t* q_pids, const int64_t* g_pids, const int64_t* q_camids,
                        const int64_t* g_camids, int nq, int ng, int max_rank, double* ap64, float* ap32,
                        int* first_rank, float* cmc, double* map_out, int64_t* num_valid, float* map32_out,
                        int* overflow_flag, void* stream);
/* Stable ascending argsort of every row (== np.argsort(kind='stable')); ng <= 32768 (Market-1501: 19 732). */
int bpb_argsort_rows(const float* dist, int nq, int ng, int32_t* indices, void* stream);

/* ---- training backward: replaces autograd's `loss.backward()` + `optimizer.step()` for this path
 *      (torchreid/engine/image/part_based_engine.py:98-103) ---------------------------------------------------- */

/* Weight gradient of the NHWC bf16 convolution on the tcgen05 tensor cores (both operands MN-major: the pixel axis
 * is the contraction axis, no transposition pass).  dw [cout, ksize*ksize, cin] fp32 is ACCUMULATED (red.add): zero it
 * per step.  dy is the gradient w.r.t. the raw conv output, [n, ho, wo, cout] bf16, or with dy_up = 1 the
 * zero-inserted [n, 2ho, 2wo, cout] layout produced by bpb_bn_bwd_apply(zero_insert = 1) for stride-2 layers.
 * The DATA gradient needs no entry point of its own: it is bpb_conv_bf16 run on dy (zero-inserted for stride 2)
 * with the [cin][flipped taps][cout] weight copy that bpb_adam_pack maintains, ksize taps, stride 1.
 * Replaces: autograd of nn.Conv2d at hrnet.py:60-64,80-96,116-137,222-257,459-483; resnet.py:43-63,128-154. */
int bpb_conv_wgrad_bf16(const void* x, const void* dy, float* dw, int n, int h, int wd, int cin, int cout, int ksize,
                        int stride, int dy_up, void* stream);
/* out = (sum of up to 4 bf16 [m, c] gradient tensors) * [y > 0]   (y NULL: plain sum).  Merges the gradients of a
 * tensor's consumers and applies the backward of the ReLU that produced it (hrnet.py:92-96,133-137,277). */
int bpb_relu_grad_sum(const void* const* srcs, int nsrc, const void* y, void* out, int64_t m, int c, void* stream);
/* Train-mode BatchNorm2d backward in two passes over raw [n,h,w,c] bf16 (the conv output the forward normalised):
 * reduce -> coef[2][c] = {sum g / M, sum g*xhat / M}, dgamma/dbeta [c - c_lo] (written, fp32);
 * apply  -> draw = scale * (g - coef0 - xhat * coef1) as bf16, compact or zero-inserted (stride-2 producers).
 * g is dy, optionally gated by the following ReLU (relu_affine: raw*scale+shift > 0) and, with up_log2 > 0, summed over
 * the 2^up x 2^up block of the finer map dy lives on (adjoint of the nearest upsample of hrnet.py:239-241).
 * workspace: bpb_bn_bwd_workspace_bytes(c) bytes (per-block and per-group partial rows of a two-level deterministic sum +
 * their counters), zero-initialised once; mean_invstd / scale / shift / coef / workspace 16-byte aligned. */
int64_t bpb_bn_bwd_workspace_bytes(int c);
int bpb_bn_bwd_reduce(const void* dy, const void* raw, const float* mean_invstd, const float* scale, const float* shift,
                      int n, int h, int w, int c, int up_log2, int relu_affine, int c_lo, void* workspace, float* coef,
                      float* dgamma, float* dbeta, void* stream);
int bpb_bn_bwd_apply(const void* dy, const void* raw, const float* mean_invstd, const float* scale, const float* shift,
                     const float* coef, int n, int h, int w, int c, int up_log2, int relu_affine, int zero_insert,
                     void* draw, void* stream);
/* Weight gradient of the stem convolution (NCHW fp32 image, PyTorch [64,3,k,k] layout, accumulated). */
int bpb_stem_wgrad(const float* img, const void* draw, float* dw, int n, int h, int wd, int cout, int ksize, int stride,
                   void* stream);
/* torch.optim.Adam step (torchreid/optim/optimizer.py:96-103: L2 weight decay) over ONE flat fp32 parameter buffer +
 * refresh of the bf16 tensor-core operand copies of every conv weight.  `segs`: device array of nseg records of
 * bpb_param_segment_bytes() bytes {int64 off, int64 numel, int32 kind, trainable, cout, cin, k, pad, void* wf, void* wd};
 * kind 1 = conv weight whose gradient is stored OHWI (as bpb_conv_wgrad_bf16 writes it).  do_update = 0: pack only. */
int bpb_adam_pack(const void* segs, int nseg, int64_t total, float* p, const float* g, float* m, float* v, float lr,
                  float beta1, float beta2, float eps, float wd, int step, float grad_scale, int do_update,
                  void* stream);
int bpb_param_segment_bytes(void);
/* cudaMemsetAsync(p, 0, bytes) as a recordable launch (gradient accumulators are zeroed per step). */
int bpb_memset_zero(void* p, int64_t bytes, void* stream);
/* out[c][r] = in[r][c] with zero padding up to [out_rows, out_cols]: operand staging for the tf32 GEMMs of the
 * nn.Linear backward (bpbreid.py:324-350, 398-415). */
int bpb_transpose_pad(const float* in, int64_t ld_in, int rows, int cols, float* out, int64_t ld_out, int out_rows,
                      int out_cols, void* stream);

/* Loss gradients.  dlogits [n, ldg] = coef_i (softmax - smoothed target), coef_i = loss_weight/n or
 * loss_weight * w_i / max(sum|w|, 1e-12) (cross_entropy_loss.py:34-56); columns >= c are not touched. */
int bpb_ce_ls_bwd(const float* logits, int64_t ld, const int64_t* targets, const float* weights, int n, int c,
                  float eps, float loss_weight, float* dlogits, int64_t ldg, void* stream);
/* demb [K, N, D] += gradient of loss_weight * (part-averaged batch-hard triplet) through the mined pairs recorded by
 * bpb_triplet_fwd (hp/hn/hp_idx/hn_idx/out4); exact-zero distances carry no gradient
 * (part_averaged_triplet_loss.py:88-91, 95-195). */
int bpb_triplet_bwd(const float* emb_knd, const void* vis, int vis_mode, const float* hp, const float* hn,
                    const int* hp_idx, const int* hn_idx, const float* out4, int n, int k, int d, float margin,
                    float loss_weight, float* demb_knd, void* stream);
/* Sharded variant: anchors [row_lo, row_lo + rows), demb in the same [shards][K][n_local][D] layout as emb (+=);
 * all-reduce demb across ranks, the rank's own shard is then the gradient of its local embeddings. */
int bpb_triplet_bwd_rows(const float* emb, const void* vis, int vis_mode, const float* hp, const float* hn,
                         const int* hp_idx, const int* hn_idx, const float* out4, int n, int k, int d, float margin,
                         float loss_weight, int row_lo, int rows, int n_local, float* demb, void* stream);
/* Train-mode nn.BatchNorm1d backward on fp32 [m, c] (bpbreid.py:336-338, 404-409); y != NULL gates by the following
 * ReLU (y = the layer output); dx is written or accumulated; dgamma / dbeta optional. */
int bpb_bn1d_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* y, int64_t ld_y,
                 const float* gamma, float eps, int m, int c, float* dx, int64_t ld_dx, int accumulate, float* dgamma,
                 float* dbeta, void* stream);
int bpb_colsum_f32(const float* x, int64_t ld, int m, int c, float* out, void* stream);
/* Backward of the branch-map head (forward: bpb_head_branch_stats .. bpb_head_branch_pool), see csrc/head_bwd.cu:
 * pool_bwd_prep scales g = dL/dpooled [K+3, n, c] in place by the pooling norm and emits d(mask sums) [n, K+3];
 * dmask: out[n, q, K+3] = <g[j, n, c_off:], x[n, q, :]>;  pixel_bwd: per image mask/softmax backward (+ the
 * body-part attention loss when target_masks != NULL, coefficient pix_coef = weight / (n*h*w)) -> low-resolution
 * logit gradients dlow[b] [n*q_b, K+1], dbf [K+1] +=;  dwf [K+1, ctot] += dlow^T x;  fold_bwd: parameter gradients of
 * PixelToPartClassifier (bpbreid.py:376-385) + the batch-statistics coefficients coefab [2][ctot];
 * branch_bwd: dx[b] (bf16, same shape as the branch map). */
int bpb_head_pool_bwd_prep(float* g, const float* pooled, const float* mask_sum, int n, int c, int kp1, int hw_full,
                           float* dms, void* stream);
int bpb_head_dmask(const void* x, const float* gs, int n, int q, int cb, int c_off, int ctot, int kp1, float* out,
                   void* stream);
int bpb_head_pixel_bwd(const float* const* dmask_t, float* const* dlow, const int* hb, const int* wb, int nb,
                       const float* probs, const float* dms, int n, int h, int w, int kp1, const float* target_masks,
                       int hm, int wm, float label_smoothing, float pix_coef, const float* ext_dlogits, float* dbf,
                       void* stream);
int bpb_head_dwf(const void* x, const float* dlow, int n, int q, int cb, int c_off, int ctot, int kp1, float* dwf,
                 void* stream);
int bpb_pixel_fold_bwd(const float* dwf, const float* dbf, const float* w, const float* gamma, const float* scale,
                       const float* shift, const float* mean_invstd, int kp1, int c, int64_t count, float* dw,
                       float* db, float* dgamma, float* dbeta, float* coefab, void* stream);
int bpb_head_branch_bwd(const void* x, const float* mt, const float* gs, const float* dlow, const float* wf,
                        const float* coefab, int n, int hb, int wb, int cb, int c_off, int ctot, int h, int w, int kp1,
                        void* dx, void* stream);

/* out [planes, ho, wo] = bilinear (align_corners=True) resize of in [planes, hi, wi], fp32: nn.functional.interpolate of
 * the external part masks (bpbreid.py:153, 162-163, 171-172). */
int bpb_resize_bilinear_f32(const float* in, float* out, int64_t planes, int hi, int wi, int ho, int wo, void* stream);
/* y[r, :] = x[r, :] / max(|x[r, :]|_2, eps) for `rows` rows of d floats (in place allowed).
 * Replaces Engine.normalize = F.normalize(p=2, dim=-1) (torchreid/engine/engine.py:558-559), applied to the
 * [N, P, D] test embeddings at torchreid/engine/image/part_based_engine.py:205-208. */
int bpb_l2_normalize(const float* x, float* y, int64_t rows, int d, float eps, void* stream);

/* ---- input pipeline: torchreid/data/transforms.py:73-163 (build_transforms' per-sample list) and
 * torchreid/data/masks_transforms/mask_transform.py:21-85, one launch per batch -------------------------------------
 * Every sample has its own source size: `offsets[n]` (elements from the base pointer) and `geom[n][10]` =
 * {src_h, src_w, flip, crop_x, crop_y (-1: no crop), erase_x0, erase_y0, erase_x1, erase_y1, 0}; the host draws the
 * augmentation parameters.  mean255 / inv_std255 / erase_fill / group_bits are HOST arrays (copied into the launch).
 *  images: RGB uint8 HWC -> cv2.resize(INTER_LINEAR) fixed-point, bit-exact -> flip -> zero pad `pad` + crop ->
 *          (x - mean255) * inv_std255 (albumentations.Normalize) -> erase rectangle := erase_fill -> out [n][3][H][W] f32
 *  masks:  fields float [36][src_h][src_w] per sample -> cv2 INTER_NEAREST to (H, W) -> flip / pad+crop / erase with 0 ->
 *          parts = clamp(max|sum over group_bits[k]) (MaskGroupingTransform) -> background (strategy 0 sum,
 *          1 threshold, 2 diff_from_max; AddBackgroundMask) -> softmax(weight*m) (weight <= 0: m / sum) ->
 *          nearest resize to (out_h, out_w) (ResizeMasks) -> out [n][parts+1][out_h][out_w] f32 */
int bpb_preprocess_images(const uint8_t* pixels, const int64_t* offsets, const int32_t* geom, int n, int height, int width,
                          int pad, const float* mean255, const float* inv_std255, const float* erase_fill, float* out,
                          void* stream);
int bpb_preprocess_masks(const float* fields, const int64_t* offsets, const int32_t* geom, int n, int height, int width,
                         int out_h, int out_w, int pad, const uint64_t* group_bits, int parts, int combine_sum,
                         int strategy, float threshold, float softmax_weight, float* out, void* stream);

/* ---- k-reciprocal re-ranking: torchreid/utils/rerank.py:31-113, called at part_based_engine.py:218-226 -------------
 * Stages of one re_ranking(q_g, q_q, g_g, k1, k2, lambda) call, all device-resident (A = nq + ng):
 *  build       orig[A][A] = transpose(D^2 / colmax(D^2)), D = [[qq, qg], [qg^T, gg]]   (rerank.py:36-46); colmax: A floats
 *  topk        rank[A][k] = first k columns of the stable argsort of every row of orig  (rerank.py:47; the reference's
 *              introsort may permute exact ties, this is the (value, index) order)
 *  kreciprocal expanded k-reciprocal sets + weights exp(-orig)/sum as sparse rows of capacity
 *              bpb_rerank_row_capacity(k1): vidx/vval [A][cap], vcnt [A]               (rerank.py:54-79); k1 <= 31
 *  qexpand     local query expansion: mean of the rows V[rank[i, :k2]] as sorted sparse rows of capacity
 *              min(k2, A)*cap: qidx/qval/qcnt (k2 == 1: rows copied and sorted)         (rerank.py:81-86); k2*cap <= 4096
 *  csc_count / csc_fill   column-compressed copy (rerank.py:88-90): colcnt[A]; the caller turns it into the exclusive
 *              prefix col_ptr[A+1]; cursor[A] scratch; crow/cval sized col_ptr[A]
 *  jaccard     out[nq][ng] = (1 - t/(2-t))*(1-lambda) + orig[:nq, nq:]*lambda, t = sum_c min(V[i,c], V[g,c])
 *              (rerank.py:92-111); scratch: bpb_rerank_jaccard_blocks(nq, A)*A floats, only read when A*4 > 200 KB */
int bpb_rerank_build(const float* qg, const float* qq, const float* gg, int nq, int ng, float* colmax, float* orig,
                     void* stream);
int bpb_rerank_topk(const float* orig, int a, int k, int32_t* rank, void* stream);
int bpb_rerank_row_capacity(int k1);
int bpb_rerank_kreciprocal(const float* orig, const int32_t* rank, int a, int rank_cols, int k1, int32_t* vidx, float* vval,
                           int32_t* vcnt, void* stream);
int bpb_rerank_qexpand(const int32_t* rank, int rank_cols, int k2, const int32_t* vidx, const float* vval,
                       const int32_t* vcnt, int k1, int a, int32_t* qidx, float* qval, int32_t* qcnt, void* stream);
int bpb_rerank_csc_count(const int32_t* qidx, const int32_t* qcnt, int qcap, int a, int32_t* colcnt, void* stream);
int bpb_rerank_csc_fill(const int32_t* qidx, const float* qval, const int32_t* qcnt, int qcap, int a,
                        const int32_t* col_ptr, int32_t* cursor, int32_t* crow, float* cval, void* stream);
int bpb_rerank_jaccard_blocks(int nq, int a);
int bpb_rerank_jaccard(const float* orig, const int32_t* qidx, const float* qval, const int32_t* qcnt, int qcap,
                       const int32_t* col_ptr, const int32_t* crow, const float* cval, int nq, int a,
                       float one_minus_lambda, float lambda, float* scratch, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BPBREID_SM100_H_ */
