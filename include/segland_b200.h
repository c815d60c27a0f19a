/*
 * segland_b200.h -- C ABI of libsegland_b200.so
 *
 * The drop-in boundary for SegLand's POP-head + dense post-processing hot path
 * (SURVEY.md section 8).  The reference has no FFI layer of its own: its "operator
 * surface" is a set of Python functions.  Each entry point below names the
 * reference function (file:line under the SegLand tree) whose arithmetic it
 * replaces; segland_b200/ops.py binds them with ctypes and keeps the reference's
 * Python signatures (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - the library never allocates, frees or synchronises: the caller owns every
 *     buffer and passes the CUDA stream (a cudaStream_t cast to void*, NULL = the
 *     legacy default stream) the work is queued on;
 *   - the device is the caller's current CUDA device; entry points are re-entrant
 *     and may be called concurrently on different streams / devices;
 *   - return value: 0 on success, a negative SL_E* code for argument errors
 *     (nothing was launched), or a positive cudaError_t from the launch;
 *   - tensors are dense, row-major, in the reference's own layout (NCHW);
 *   - "bf16" buffers are raw uint16_t bit patterns of __nv_bfloat16.
 *   - accumulating outputs (confusion matrices, inter/union meters) are ADDED to:
 *     zero them once per sweep, not per call.
 */
#ifndef SEGLAND_B200_H
#define SEGLAND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SL_ABI_VERSION 1

/* Environment switches read by the library (debugging and A/B measurements only; defaults are the fast paths).
 * They are read ONCE per process (no getenv on the launch path); sl_env_reload() re-reads them.
 *   SL_POST_PRUNE=1     sl_upsample_argmax's prediction-only path on the per-cell class-pruning kernel (post_prune.cu)
 *                       instead of the row-cached kernel: identical results; see profiles/ for when it pays
 *   SL_POST_REGS=0      sl_upsample_argmax's prediction-only path (K = 8 / 12, up-sampling by >= 2x) on the
 *                       row-cached shared-memory kernel instead of the register-resident one (post_regs.cu): identical results
 *   SL_TC_PAIR=0        (builds with -DSL_AB_VARIANTS only) background MLP on the single-CTA tcgen05 kernel instead of
 *                       the cta_group::2 pair kernel
 *   SL_TC_PAIR=1        (builds with -DSL_AB_VARIANTS only) pair kernel with one (A, B) operand pair per MMA pass and
 *                       pipeline stage (the schedule before the de-duplicated stages; 4-10 % slower)
 *   SL_TC_SMALL=0       C <= 128: streaming kernel instead of the weights-resident narrow-head kernel
 *   SL_PREP_SPLIT=1     sl_pop_prepare as five separate launches instead of three
 *   SL_POST_FUSED_CM=1/0 sl_upsample_argmax counts the confusion matrix inside the interpolation kernel (1, the default
 *                       for every kernel) or in a second launch over (label, pred) (0)
 *   SL_TC_DEBUG=<bits>  knock-outs inside the single-CTA kernel (timing experiments, results INVALID)
 *   SL_FG_MMA=0         sl_pop_fg_lowres on the FFMA2 CUDA-core kernel instead of the mma.sync kernel
 *   SL_TAIL_FUSED=0     sl_tail_bn_relu_conv as two kernels (bf16 hi/lo planes in the workspace + generic GEMM)
 *   SL_SMALL_DBG=<ptr>  device pointer (decimal) of a 64x16 int64 buffer receiving per-tile cycle stamps of CTA 0
 */

#if defined(__GNUC__)
#define SL_API __attribute__((visibility("default")))
#else
#define SL_API
#endif

#define SL_OK 0
#define SL_EINVAL (-1)      /* a size/shape argument is out of the supported range   */
#define SL_ENULL (-2)       /* a required pointer is NULL                             */
#define SL_EALIGN (-3)      /* a pointer/extent violates the stated alignment         */
#define SL_EUNSUPPORTED (-4)/* the device is not sm_100 (this library is B200-only)  */

#define SL_MAX_CLASSES 32   /* 1 + Kb + Kn must be <= 32 (OEM: 12)                    */
#define SL_MAX_FUSE 16      /* fusemat model count M <= 16                            */
#define SL_MAX_WINDOWS_1D 32 /* sliding-window origins per axis                       */
#define SL_MAX_VIEWS 8      /* flipped views per window                               */

SL_API int sl_abi_version(void);
/* Static string for any code an entry point can return (SL_E* or cudaError_t). */
SL_API const char *sl_error_string(int code);
/* 0 iff the current device is compute capability 10.x; SL_EUNSUPPORTED otherwise. */
SL_API int sl_check_device(void);
/* Re-read the SL_* environment switches (they are otherwise cached at first use). */
SL_API int sl_env_reload(void);

/* ---------------------------------------------------------------------------
 * (a1/a2) POP head -- GFSS_Model.orthogonal_decompose + classifier/classifier_n
 *   networks/pspnet_pop.py:95-121 (decompose), :46-52,57-63 (3-layer 1x1 MLP),
 *   :143-159 (forward_all), :171-182 (forward_base).  Identical bodies in
 *   networks/{pspplus,deeplab,vggunet,seghr,swin,convnext,lsk}_pop.py.
 *
 * sl_pop_prepare: everything that depends only on weights/prototypes.
 *   protos  [K,C] fp32 raw prototypes, base classes first then novel
 *           (cat[base_emb, novel_emb]); K = Kb + Kn >= 1; C % 8 == 0, 8 <= C <= 1024.
 *   W1_fg/W2_fg [C,C], w3_fg [C]: the MLP applied to classes [0,Kb)  (self.classifier)
 *   W1_bg/W2_bg/w3_bg: the MLP applied to the background vector and to classes
 *           [Kb,K) (self.classifier_n in ft mode; pass the same pointers as *_fg in
 *           base mode, pspnet_pop.py:178-182).
 * outputs
 *   s_hat   [K,C]  F.normalize(protos, p=2, dim=-1), eps 1e-12 (pspnet_pop.py:106,113)
 *   alpha,beta [K] alpha_k = MLP(+s_hat_k), beta_k = MLP(-s_hat_k): the MLP is
 *           bias-free with ReLUs, hence positively homogeneous, so the logit of the
 *           rank-1 foreground vector p*s_hat_k is p>=0 ? p*alpha_k : -p*beta_k.
 *   W1p_t   [C_in][C_out] fp32, transposed W1' = W1_bg (I - S_hat^T S_hat): folds
 *           "bg = q - sum_k p_k s_hat_k" (pspnet_pop.py:112,118) into layer 1.
 *   W2_t    [C_in][C_out] fp32, transposed W2_bg.
 *   W1p_hi/lo, W2_hi/lo  [C_out][C_in] bf16 split (hi = bf16(W), lo = bf16(W - hi)) for
 *           the tensor-core path; may all be NULL when only the SIMT path is used.
 *   W1p_f16, W2_f16  [C_out][C_in] IEEE fp16 copies of W1' and W2 (both or neither; NULL when
 *           unused).  SL_TC_BALANCED reads W2_f16.
 *   ws      scratch, sl_pop_prepare_ws_bytes(K, C) bytes (the two hidden layers of the 2K
 *           +-s_hat_k vectors).
 */
SL_API size_t sl_pop_prepare_ws_bytes(int K, int C);
SL_API int sl_pop_prepare(const float *protos, int K, int Kb, int C,
                   const float *W1_fg, const float *W2_fg, const float *w3_fg,
                   const float *W1_bg, const float *W2_bg, const float *w3_bg,
                   float *s_hat, float *alpha, float *beta,
                   float *W1p_t, float *W2_t,
                   uint16_t *W1p_hi, uint16_t *W1p_lo, uint16_t *W2_hi, uint16_t *W2_lo,
                   uint16_t *W1p_f16, uint16_t *W2_f16, float *ws, void *stream);

/* K foreground logits at feature resolution (HBM-bound; the projections run on the legacy tensor path -- mma.sync
 * m16n8k16 with the fp32 prototypes split into three bf16 terms, fp32 accumulation -- so that the K FMAs per feature
 * element do not bind before HBM does).
 *   feat   [B,C,N] bf16 (features.flatten(2), pspnet_pop.py:148); C % 8 == 0, C <= 512,
 *          N % 8 == 0, 16-byte aligned.
 *   logits [B,Ktot,N] fp32; channel ch_map[k] (host array of K ints) receives class
 *          k's logit -- forward_all's order [bg, base.., novel..] (pspnet_pop.py:159)
 *          is ch_map[k] = 1 + k.
 */
SL_API int sl_pop_fg_lowres(const uint16_t *feat, int B, int C, int N,
                     const float *s_hat, const float *alpha, const float *beta, int K,
                     float *logits, int Ktot, const int *ch_map_host, void *stream);

/* Background logit (class 0), exact fp32 CUDA-core path:
 *   logit_0 = w3 . relu(W2 relu(W1' q))   written to channel `ch` of logits.
 */
SL_API int sl_pop_bg_simt(const uint16_t *feat, int B, int C, int N,
                   const float *W1p_t, const float *W2_t, const float *w3_bg,
                   float *logits, int Ktot, int ch, void *stream);

/* Background logit on tcgen05 tensor cores, fp32 accumulation in TMEM.
 * C % 32 == 0, 32 <= C <= 512, N % 8 == 0 (a partial last 128-pixel tile is zero-filled by TMA); SL_EINVAL
 * outside that range (callers use _simt).
 *   precision  SL_TC_PRECISE  split-bf16 operands, 2 + 3 MMA passes: ~5e-6 of the fp32 reference.
 *              SL_TC_BALANCED layer 1 split-bf16 (2 passes), layer 2 single-pass fp16 (3 passes):
 *                             ~3e-4 relative to the tensor maximum.  Needs W2_f16.
 *              SL_TC_MID      layer 1 split-bf16 (2 passes), layer 2 with the hidden layer split into fp16 hi + lo
 *                             planes against single fp16 weights (2 passes, 4 in total): the only rounding left is
 *                             W2 -> fp16 (2^-12 per weight).  Needs W2_f16.  See profiles/r2_pass_probe.txt.
 *              The fp16 modes require |relu(W1' q)| < 65504 (fp16 range).
 *   Unused weight pointers for the chosen mode may be NULL.
 *   h1_ws: scratch for the hidden layer, sl_pop_bg_tc_ws_bytes(B, C, N) bytes, 128-byte aligned
 *          (two 128-pixel tiles per SM; it stays L2-resident).
 */
#define SL_TC_PRECISE 0
#define SL_TC_BALANCED 1
#define SL_TC_MID 2
SL_API size_t sl_pop_bg_tc_ws_bytes(int B, int C, int N);
SL_API int sl_pop_bg_tc(const uint16_t *feat, int B, int C, int N,
                 const uint16_t *W1p_hi, const uint16_t *W1p_lo,
                 const uint16_t *W2_hi, const uint16_t *W2_lo,
                 const uint16_t *W2_f16, const float *w3_bg, int precision, uint16_t *h1_ws, float *logits, int Ktot, int ch, void *stream);

/* Backward of the POP head for training: the per-pixel part of autograd through
 * forward_novel / forward_base (networks/pspnet_pop.py:191-219, :161-189), i.e. through
 * orthogonal_decompose (:95-121) and classifier / classifier_n (:46-63), in exact fp32 on the CUDA cores.
 * Inputs are the forward's operands -- features, s_hat [K,C], alpha / beta [K], the FOLDED first layer
 * W1p = W1 (I - S_hat^T S_hat) [C_out][C_in] row-major, W2 [C,C], w3 [C] of the background MLP -- and
 * g_logits [B,Ktot,N] = dL/dlogits in the forward's channel layout (fg_ch[k] = channel of class k, bg_ch).
 * Outputs (overwritten): d_s_hat [K,C] (through the projections p_k = s_hat_k . q only; the dependence of
 * W1p on s_hat is differentiated by the caller from dW1p), d_alpha, d_beta [K], dW1p, dW2 [C,C], dw3 [C] and,
 * when d_feat != NULL, dL/dfeatures [B,C,N] fp32.  Activations are recomputed; ws holds them
 * (sl_pop_head_bwd_ws_bytes, 128-byte aligned).  C % 8 == 0, C <= 512, N % 8 == 0, B*N < 2^31.
 * mode: SL_BWD_AUTO runs the GEMMs on tcgen05 as split-bf16 products (2-3 passes, fp32 accumulation, ~1e-5 of
 *       fp32; C >= 32) and falls back to the exact CUDA-core SGEMMs for narrower heads; SL_BWD_SIMT forces
 *       the exact fp32 path.
 */
#define SL_BWD_AUTO 0
#define SL_BWD_SIMT 1
SL_API size_t sl_pop_head_bwd_ws_bytes(int B, int C, int N, int K);
SL_API int sl_pop_head_bwd(const uint16_t *feat, int B, int C, int N,
                    const float *s_hat, const float *alpha, const float *beta, int K, const int *fg_ch_host,
                    const float *W1p, const float *W2, const float *w3,
                    const float *g_logits, int Ktot, int bg_ch,
                    float *d_s_hat, float *d_alpha, float *d_beta, float *dW1p, float *dW2, float *dw3,
                    float *d_feat, int mode, void *ws, void *stream);

/* Backward of sl_pop_prepare (training): the parameter-side chain s_hat = normalize(protos), alpha/beta =
 * MLP(+-s_hat), W1' = W1_bg (I - S_hat^T S_hat) (pspnet_pop.py:106,113; :46-63; :112,118), given the
 * outputs of sl_pop_head_bwd: d_s_hat [K,C], d_alpha, d_beta [K], dW1p [C,C] and the background MLP's direct
 * gradients dW2_direct [C,C], dw3_direct [C].  Outputs (overwritten): d_protos [K,C] (rows [0,Kb) = base_emb,
 * [Kb,K) = novel_emb), dW1/dW2/dw3 of the background MLP (classifier_n in ft mode) and, unless NULL, of the
 * foreground MLP (classifier; pass NULL when it is frozen).  When both MLPs are the same tensors (base mode)
 * pass NULL for the *_fg outputs: everything accumulates into *_bg.  ws: sl_pop_prepare_bwd_ws_bytes(K, C).
 */
SL_API size_t sl_pop_prepare_bwd_ws_bytes(int K, int C);
SL_API int sl_pop_prepare_bwd(const float *protos, int K, int Kb, int C,
                       const float *W1_fg, const float *W2_fg, const float *w3_fg,
                       const float *W1_bg, const float *W2_bg, const float *w3_bg,
                       const float *d_s_hat, const float *d_alpha, const float *d_beta, const float *dW1p,
                       const float *dW2_direct, const float *dw3_direct,
                       float *d_protos, float *dW1_fg, float *dW2_fg, float *dw3_fg,
                       float *dW1_bg, float *dW2_bg, float *dw3_bg, float *ws, void *stream);

/* The whole POP head in one launch (tensor-core path): sl_pop_bg_tc with the K <= 12 foreground logits of
 * sl_pop_fg_lowres computed by extra warps from the feature tiles already staged for the MMAs, so the
 * features are read from HBM once.  Arguments as in the two calls it replaces; ch_map_host[k] is the
 * output channel of class k, bg_ch the background channel.  Same shape limits and scratch as sl_pop_bg_tc.
 */
SL_API int sl_pop_head_tc(const uint16_t *feat, int B, int C, int N,
                   const uint16_t *W1p_hi, const uint16_t *W1p_lo,
                   const uint16_t *W2_hi, const uint16_t *W2_lo,
                   const uint16_t *W2_f16, const float *w3_bg, int precision,
                   const float *s_hat, const float *alpha, const float *beta, int K,
                   const int *ch_map_host, uint16_t *h1_ws, float *logits, int Ktot, int bg_ch,
                   void *stream);

/* Test-time view aggregation at feature resolution (spec: this repo -- the reference
 * has no flip/sliding-window inference): out = scale * sum_v unflip(view_v).
 *   views [V,B,K,h,w] fp32; flip_host[v] bit0 = horizontal flip, bit1 = vertical flip.
 */
SL_API int sl_views_reduce(const float *views, int V, int B, int K, int h, int w,
                    const int *flip_host, float scale, float *out, void *stream);

/* Sliding-window + flip aggregation of per-crop logits at feature resolution (spec: this repo -- north_star
 * names "engine.py's sliding-window/flip aggregation", the reference has none: engine.py:23-143 is argparse/DDP
 * only and eval_base.py:162-170 / eval_ft.py:162-172 run whole tiles).  A tile's canvas [h,w] (feature
 * pixels) is covered by an ny x nx grid of crops of hc x wc feature pixels with ascending origins oy_host[ny],
 * ox_host[nx] (first = 0, last = h - hc / w - wc: the last window is pulled back to the border, consecutive windows
 * touch or overlap); every window has V views, view v flipped by flip_host[v] (bit0 = horizontal, bit1 = vertical).
 *   crops  entry e = (gy*nx + gx)*V + v, image b, class k at crops + e*stride_e + b*stride_b + k*hc*wc (fp32,
 *          [hc][wc] row-major): both [E,B,K,hc,wc] (stride_e = B*K*hc*wc, stride_b = K*hc*wc) and [B,E,K,hc,wc]
 *          (stride_b = E*K*hc*wc, stride_e = K*hc*wc) layouts work.
 *   canvas [B,K,h,w] fp32 = sum of the un-flipped covering crops, added in entry order, divided by their number
 *          (OVERWRITTEN); it feeds sl_upsample_argmax.  count [h,w] fp32 or NULL receives the overlap counts.
 * Gather formulation: no atomics, fixed summation order, bit-reproducible.
 */
SL_API int sl_window_accumulate(const float *crops, long long stride_e, long long stride_b, int B, int K,
                         int hc, int wc, const int *oy_host, int ny, const int *ox_host, int nx,
                         const int *flip_host, int V, int h, int w, float *canvas, float *count, void *stream);

/* ---------------------------------------------------------------------------
 * (a4/a5) F.interpolate(bilinear, align_corners=True) -> argmax -> confusion
 *   eval_base.py:168-178, eval_ft.py:168-183, ft_pop.py:327-331.
 *   logits_lr [B,K,h,w] fp32; output size H x W; 2 <= K <= SL_MAX_CLASSES.
 *   label   [B,H,W] u8 or NULL; pixels == ignore_label are excluded from cm.
 *   pred    [B,H,W] u8 or NULL: first-maximum argmax (np.argmax semantics).
 *   conf    [B,H,W] fp32 or NULL: max softmax probability (spec: this repo).
 *   probs   [B,K,H,W] fp32 or NULL: softmax over channels (spec: this repo).
 *   logits_hr [B,K,H,W] fp32 or NULL: the upsampled logits themselves
 *           (what eval_base.py:190-191 dumps to .mat for fusemat.py).
 *   cm      [K,K] int64 or NULL, row = gt, col = pred, ACCUMULATED
 *           (get_confusion_matrix, utils/pyt_utils.py:182-200); requires label.
 */
SL_API int sl_upsample_argmax(const float *logits_lr, int B, int K, int h, int w, int H, int W,
                       const uint8_t *label, int ignore_label,
                       uint8_t *pred, float *conf, float *probs, float *logits_hr,
                       long long *cm, void *stream);

/* (a10) pseudo-labelling of base-image background, pspnet_pop.py:221-231:
 *   idx = argmax(upsample(preds2[b])); idx[idx>0] += n_base; mask[b][mask[b]==0] = idx.
 *   preds2 [B,K2,h,w] fp32 (K2 = 1+Kn); mask [B,H,W] int64, updated IN PLACE.
 */
SL_API int sl_pseudo_label(const float *preds2, int B, int K2, int h, int w, int H, int W,
                    int n_base, long long *mask, void *stream);

/* (a5) get_confusion_matrix(gt, pred, K), utils/pyt_utils.py:182-200, on label maps:
 *   cm[gt*K+pred] += 1 for every i < n with gt[i] != ignore_label.  Labels >= K that
 *   are not ignore_label are skipped and counted in *n_bad (int64, may be NULL).
 */
SL_API int sl_confusion(const uint8_t *gt, const uint8_t *pred, long long n, int K,
                 int ignore_label, long long *cm, long long *n_bad, void *stream);

/* (a7) intersectionAndUnionGPU(output, target, K, ignore), utils/pyt_utils.py:293-305.
 *   output,target [n] int64.  Mirrors the reference's in-place side effect:
 *   output[target == ignore] = ignore.  inter/uni/tgt [K] fp32 are OVERWRITTEN with the
 *   per-call areas (the caller accumulates, ft_pop.py:333-334); ws [3*K] int64 is scratch
 *   the call zeroes itself (exact integer areas before the fp32 conversion histc implies).
 */
SL_API int sl_inter_union(long long *output, const long long *target, long long n, int K,
                   int ignore_label, float *inter, float *uni, float *tgt,
                   long long *ws, void *stream);

/* ---------------------------------------------------------------------------
 * (a9) masked_average_pooling(feature, mask), networks/pspnet.py:7-15
 *   feat [B,C,h,w] bf16; mask [B,1,H,W] fp32.
 *   mask_lr_ws [B*h*w + B*ceil(h*w/1024) + 8*B*C] fp32 scratch: the align_corners bilinear down-sample
 *           of mask, per-1024-pixel chunk sums, and up to 8 partial channel sums per image (added in
 *           index order, so the result is bit-reproducible);
 *   per_image [B,C] fp32 = sum_hw(f*m)/(sum_hw m + 1e-5); proto [C] fp32 = mean_B.
 *   h*w % 8 == 0.
 */
SL_API int sl_map_proto(const uint16_t *feat, const float *mask, int B, int C, int h, int w,
                 int H, int W, float *mask_lr_ws, float *per_image, float *proto,
                 void *stream);

/* (a8) OrthLoss.get_orth_loss, loss/criterion.py:37-43, with proto_sim built as in
 *   pspnet_pop.py:185-186 (base: Kb_other = 0, rows = base_emb) and :234-239
 *   (ft: rows = novel_emb, others = base_emb):
 *   r_hat = normalize(rows), sim = r_hat @ cat[r_hat, normalize(others)]^T  [Kr, Kr+Ko]
 *   loss = mean |sim[i][j]| over j > i.
 *   proto_sim [Kr,Kr+Ko] fp32 (may be NULL), loss [1] fp32,
 *   grad_rows [Kr,C] fp32 or NULL = d loss / d rows (others are frozen in ft mode).
 */
SL_API int sl_orth_loss(const float *rows, int Kr, const float *others, int Ko, int C,
                 float *proto_sim, float *loss, float *grad_rows, void *stream);

/* (a8) OrthLoss.get_orth_loss(proto_sim), loss/criterion.py:37-43, on a proto_sim matrix [Kr,Kc] fp32 the caller
 *   built (the model does, pspnet_pop.py:185-186,234-239): loss [1] = mean |proto_sim[i][j]| over j > i, and, unless
 *   NULL, grad_sim [Kr,Kc] = d loss / d proto_sim (sign / count on the selected entries, 0 elsewhere).
 */
SL_API int sl_orth_from_sim(const float *proto_sim, int Kr, int Kc, float *loss, float *grad_sim, void *stream);

/* (f-1) segmentation cross-entropy tail of OrthLoss.forward / CELoss.forward,
 *   loss/criterion.py:17-19,51-52:
 *     scale_pred = F.interpolate(pred, target.shape[1:], mode='bilinear', align_corners=True)
 *     loss = nn.CrossEntropyLoss(ignore_index, reduction='mean')(scale_pred, target)
 *   without materialising scale_pred.  logits_lr [B,K,h,w] fp32; target [B,H,W] int64.
 *   ws: sl_upsample_ce_ws_bytes(B,K,w,H,W) bytes, 16-byte aligned; the forward leaves the per-pixel
 *       log-sum-exp there; the backward reads it and uses the rest for its row-reduced gradients
 *       (same buffer, same shapes).
 *   fwd: loss [1] fp32 (NaN when every pixel is ignored, like torch), n_valid [1] int64.
 *        A target outside [0,K) that is not ignore_label makes nn.CrossEntropyLoss raise a device assert; here the
 *        loss (and every gradient) becomes NaN and n_valid = -(number of such targets).
 *   bwd: grad_out [1] fp32 = d(objective)/d(loss); grad_logits_lr [B,K,h,w] fp32 is OVERWRITTEN.
 *   Both directions are deterministic (no floating-point atomics).
 */
SL_API size_t sl_upsample_ce_ws_bytes(int B, int K, int w, int H, int W);
SL_API int sl_upsample_ce_fwd(const float *logits_lr, int B, int K, int h, int w, int H, int W,
                       const long long *target, int ignore_label, void *ws,
                       float *loss, long long *n_valid, void *stream);
SL_API int sl_upsample_ce_bwd(const float *logits_lr, int B, int K, int h, int w, int H, int W,
                       const long long *target, int ignore_label, void *ws,
                       const long long *n_valid, const float *grad_out,
                       float *grad_logits_lr, void *stream);

/* ---------------------------------------------------------------------------
 * (a11) fusemat.py:37-48: mats[0] + mats[1] + ... (sequential fp32 adds, in order),
 *   / divisor (fp32 IEEE division by len(fusion_list)), argmax over K (first max).
 *   mats_host: HOST array of M device pointers, each [K,HW] fp32; HW % 4 == 0.
 *   pred [HW] u8; fused [K,HW] fp32 or NULL (the divided sum);
 *   label/cm as in sl_upsample_argmax (optional mIoU of the fused map).
 */
SL_API int sl_fuse_argmax(const float *const *mats_host, int M, int K, long long HW, int divisor,
                   uint8_t *pred, float *fused,
                   const uint8_t *label, int ignore_label, long long *cm, void *stream);

/* A whole sweep in one call: mats_host[m] is [T,K,HW] (model m, all tiles, tile-major); pred [T,HW], fused [T,K,HW]
 * or NULL, label [T,HW] or NULL; every tile is fused exactly as by sl_fuse_argmax, cm accumulates over the sweep.
 */
SL_API int sl_fuse_argmax_tiles(const float *const *mats_host, int M, int T, int K, long long HW, int divisor,
                         uint8_t *pred, float *fused, const uint8_t *label, int ignore_label, long long *cm,
                         void *stream);

/* ---------------------------------------------------------------------------
 * (f-4) decoder tails: the last operators of the reference's decoders, emitting the head's bf16 NCHW features
 *   directly, so the fp32 feature tensor makes no round trip through HBM between decoder and head (the
 *   reference hands fp32 features to orthogonal_decompose, pspnet_pop.py:142-148; the bf16 conversion is this
 *   repo's input format).  x / maps are fp32 [B,C,N] (NCHW, N = h*w, N % 8 == 0, 16-byte aligned);
 *   feat_out is [B,C,N] bf16 = round-to-nearest-even of the fp32 result.
 *
 * sl_tail_layernorm: FPN_Seg_OCR_Decoder.norm, networks/convnext_pop.py:13,27 --
 *   nn.LayerNorm(C) applied over the channel axis of every pixel (biased variance, eps inside the sqrt),
 *   y = (x - mean) / sqrt(var + eps) * gamma[c] + beta[c].  C <= 1536.
 *
 * sl_tail_bn_relu_conv: PSPModule.bottleneck[1:4], networks/pspnet_pop.py:19-22 (and PSP_Plus_Decoder.fc[1:4],
 *   networks/pspplus_pop.py:44-47): inference-mode BatchNorm2d -> ReLU -> 1x1 convolution with bias,
 *   y[co] = bias[co] + sum_ci W[co][ci] * relu(bn(x)[ci]).  The convolution runs on tcgen05 as a split-bf16
 *   product (3 passes, fp32 accumulation, ~1e-5 of fp32).  bn_weight == NULL skips the normalisation
 *   (then bn_bias/mean/var are ignored); relu == 0 skips the ReLU; bias may be NULL.
 *   W_hi/W_lo [Cout][Cin] bf16 come from sl_tail_conv_prepare(W [Cout][Cin] fp32), once per weight update.
 *   Cin % 8 == 0; ws: sl_tail_bn_relu_conv_ws_bytes(B, Cin, N) bytes, 128-byte aligned.
 *
 * sl_tail_sum: torch.stack(fpn_outs, dim=-1).sum(-1), networks/swin_pop.py:169-172, lsk_pop.py:163-165:
 *   maps_host is a HOST array of M <= 8 device pointers, each n fp32 elements (n % 8 == 0); sequential fp32 adds
 *   in map order.
 */
SL_API int sl_tail_layernorm(const float *x, int B, int C, int N, const float *gamma, const float *beta, float eps,
                      uint16_t *feat_out, void *stream);
SL_API int sl_tail_conv_prepare(const float *W, int Cout, int Cin, uint16_t *W_hi, uint16_t *W_lo, void *stream);
SL_API size_t sl_tail_bn_relu_conv_ws_bytes(int B, int Cin, int N);
SL_API int sl_tail_bn_relu_conv(const float *x, int B, int Cin, int N,
                         const float *bn_weight, const float *bn_bias, const float *bn_mean, const float *bn_var,
                         float bn_eps, int relu, const uint16_t *W_hi, const uint16_t *W_lo, const float *bias,
                         int Cout, void *ws, uint16_t *feat_out, void *stream);
SL_API int sl_tail_sum(const float *const *maps_host, int M, long long n, uint16_t *feat_out, void *stream);
/* sl_tail_bn_relu: inference BatchNorm2d -> ReLU -> bf16, the `bn` + `relu` of _ConvBnReLU after _ASPP.fc's
 *   convolution (networks/deeplab_pop.py:12-29,61,66) and DoubleConv's last BN + ReLU in VGGUNet.up4
 *   (networks/vggunet_pop.py:19-20,79).  Arguments as in sl_tail_bn_relu_conv; x [B,C,N] fp32.
 * sl_tail_concat: torch.cat([x0, x1, x2, x3], 1) of HRFPN_Seg_Decoder (networks/seghr_pop.py:23-24) as bf16:
 *   maps_host[m] is fp32 [B,channels_host[m],N]; feat_out [B,sum(channels),N] bf16; M <= 16.
 */
SL_API int sl_tail_bn_relu(const float *x, int B, int C, int N,
                    const float *bn_weight, const float *bn_bias, const float *bn_mean, const float *bn_var,
                    float bn_eps, int relu, uint16_t *feat_out, void *stream);
SL_API int sl_tail_concat(const float *const *maps_host, const int *channels_host, int M, int B, int N,
                   uint16_t *feat_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGLAND_B200_H */
