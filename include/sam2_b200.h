/* libsam2b200.so -- C ABI of the B200-native (sm_100a) SAM2 memory-attention + mask-loss path.
 *
 * The reference (yangkunyi/sam2-video-training) is pure Python/PyTorch and has no FFI of its own
 * for this path: the two plug points are the nn.Module attribute `SAM2Base.memory_attention`
 * (sam2_video/model/modeling/sam2_base.py:125, invoked :695-709) and the criterion
 * `SAM2LightningModule.criterion` (sam2_video/training/trainer.py:67-94, invoked :268,303).
 * The functions below are what torch.autograd.Functions behind those plug points bind (ctypes;
 * see INTEGRATION.md).  For each entry point the reference code it replaces is cited.
 *
 * Conventions: plain device pointers + sizes + an explicit cudaStream_t; no allocation, no host
 * synchronisation, no C++ exceptions.  Return 0 on success, <0 on error (sam2b200_last_error()
 * gives a thread-local message).  All tensors are contiguous; bf16 = __nv_bfloat16 bits.
 */
#ifndef SAM2_B200_H_
#define SAM2_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sam2b200_stream_t; /* == cudaStream_t */

#define SAM2B200_OK 0
#define SAM2B200_ERR_INVALID (-1)
#define SAM2B200_ERR_CUDA (-2)
#define SAM2B200_ERR_UNSUPPORTED (-3)
#define SAM2B200_ERR_DRIVER (-4)

#define SAM2B200_DTYPE_F32 0
#define SAM2B200_DTYPE_BF16 1

#define SAM2B200_LOSS_MULTISTEP 0 /* focal + dice + IoU (MultiStepMultiMasksAndIous) */
#define SAM2B200_LOSS_BCE 1       /* BCECategoryLoss */

int sam2b200_version(void);
const char* sam2b200_last_error(void);
/* CUDA kernels launched by this library in this process so far (for bench.py's gpu_launches). */
long long sam2b200_launch_count(void);
/* Debug aid: per-CTA phase timelines of the attention kernels into a device buffer (NULL = off). */
long long sam2b200_debug_set_timeline(void* buf, long long n_u64);
/* Debug / A-B aid: choose a kernel variant at run time.  key 0 = backward of the raw-memory cross-attention
 * (0 = default: dK with resident CTAs walking over (key block, object) items when N <= 768, there are more items than SMs and
 * the gradients are bf16; 1 = experimental two-softmax-group kernels, 2 = one CTA per item, 3 = resident CTAs for dK and dQ at
 * every length); key 1 = rotation-table addressing of the gradient epilogues
 * (0 = default: axial -- rows x and y*w of a w x w grid table, 1 = full rows); key 3 = rotation table of the gradient epilogues
 * (0 = default: staged in shared memory, 1 = read from global memory per chunk); key 2 = forward of the raw-memory cross-attention
 * (0 = default: one online-softmax stream per CTA, 1 = experimental two-stream kernel, same results to bf16 rounding of the probabilities).  Returns the previous value, -1 for an unknown key. */
int sam2b200_debug_set_variant(int key, int value);
/* 0 iff CUDA device `dev` is an sm_100 part. */
int sam2b200_check_device(int dev);

/* ---- axial RoPE -------------------------------------------------------------------------
 * Replaces apply_rotary_enc (sam2_video/model/modeling/position_encoding.py:212-239) and the
 * slice write-back k[:, :, :num_k_rope] = ... (sam2_video/model/modeling/sam/transformer.py:296-302).
 * x, out: [B, L, 256]; rows [0, n_rope) of every batch item are rotated by table[row % n_tokens],
 * later rows (object-pointer keys) are copied.  table: [n_tokens, 128, 2] fp32 (cos, sin) --
 * compute_axial_cis (position_encoding.py:192-201).  inverse != 0 applies the conjugate rotation
 * (the backward of the rotation).  n_rope must be a multiple of n_tokens. */
int sam2b200_rope_apply(const void* x, int in_dtype, void* out, int out_dtype, const float* table,
                        int B, int L, int n_rope, int n_tokens, int inverse, sam2b200_stream_t stream);

/* ---- attention core ---------------------------------------------------------------------
 * Replaces F.scaled_dot_product_attention(q, k, v) for one head of width 256
 * (sam2_video/model/modeling/sam/transformer.py:306; plain variant :243), no mask, dropout 0.
 * q: [B, N, 256], k, v: [B, M, 256], out: [B, N, 256], all bf16; lse2: [B, N] fp32 =
 * log2 sum_j exp(scale * q.k_j) (kept for the backward).  nsplit > 1 splits the keys across
 * CTAs (for grids that do not fill 148 SMs) and needs the workspace below. */
int sam2b200_attn_default_nsplit(int B, int N, int M);
size_t sam2b200_attn_fwd_workspace_bytes(int B, int N, int M, int nsplit);
int sam2b200_attn_fwd(const void* q, const void* k, const void* v, void* out, float* out_f32 /* NULL or fp32 copy */,
                      float* lse2, void* workspace, size_t workspace_bytes, int B, int N, int M, float scale,
                      int nsplit, sam2b200_stream_t stream);
/* Same with attention-probability dropout (transformer.py:304-306, see "dropout" below). */
int sam2b200_attn_fwd_ex(const void* q, const void* k, const void* v, void* out, float* out_f32, float* lse2,
                         void* workspace, size_t workspace_bytes, int B, int N, int M, float scale, int nsplit,
                         float drop_p, const unsigned long long* drop_seed, unsigned drop_site, sam2b200_stream_t stream);
/* Backward (what autograd derives for transformer.py:296-306).  delta: [B, N] fp32 scratch.
 * dq: [B, N, ldq], dk: [B, M, ldk], dv: [B, M, ldv], fp32 (grad_dtype 0) or bf16 (1), first 256 columns
 * fully overwritten.  With rope_table != NULL (an AXIAL table as compute_axial_cis builds it: when rope_period is a square
 * w^2 the epilogues read pair j < 64 of row p from row p mod w and pair j >= 64 from row p - p mod w, which hold the same
 * values) the conjugate rotation is fused into the epilogue (all rows
 * of dq, rows [0, n_rope_k) of dk; table row = row % rope_period), i.e. the outputs are gradients with
 * respect to the un-rotated projections (the backward of apply_rotary_enc + the slice write-back). */
int sam2b200_attn_bwd(const void* q, const void* k, const void* v, const void* out /* bf16, or NULL if */,
                      const float* out_f32 /* the fp32 copy is given */, const void* dout, const float* lse2,
                      float* delta, void* dq, void* dk, void* dv, int grad_dtype, int ldq, int ldk, int ldv,
                      const float* rope_table, int rope_period, int n_rope_k, int B, int N, int M, float scale,
                      sam2b200_stream_t stream);

/* Same, plus the bias gradients of the q / k / v projections (transformer.py:220-222): if non-NULL, the column sums of
 * dq / dk / dv over all B x rows are ADDED to dbias_q / dbias_k / dbias_v (fp32 [256] each) inside the gradient
 * epilogues with fp32 atomics (summation order not fixed), instead of another pass over the gradient tensors.
 * parts: bit mask of the kernels to launch (0 = all): 1 Delta, 2 dV, 4 dK, 8 dQ (dK and dQ read Delta), so that the
 * key-side kernels can be enqueued on a second stream. */
int sam2b200_attn_bwd_ex(const void* q, const void* k, const void* v, const void* out, const float* out_f32,
                         const void* dout, const float* lse2, float* delta, void* dq, void* dk, void* dv,
                         int grad_dtype, int ldq, int ldk, int ldv, const float* rope_table, int rope_period,
                         int n_rope_k, int B, int N, int M, float scale, float* dbias_q, float* dbias_k, float* dbias_v,
                         int parts, float drop_p, const unsigned long long* drop_seed, unsigned drop_site,
                         sam2b200_stream_t stream);

/* ---- attention forward with the OUTPUT PROJECTION fused into the kernel's epilogue (north star; replaces the nn.Linear out_proj
 * of sam2_video/model/modeling/sam/transformer.py:308-309 that follows scaled_dot_product_attention at :306).
 * proj_out [B, N, 256] bf16 = out . w^T + bias; w [256, 256] bf16 row-major (nn.Linear layout), bias [256] fp32.  out / out_f32 /
 * lse2 as sam2b200_attn_fwd_ex (no split-KV).  The _v64 form is the raw-memory cross-attention with the folded projection
 * w = Wo Wv [256, 64], bias = Wo bv + bo (no dropout) or bo with rank1 = Wo bv multiplied by the dropped row sums. */
int sam2b200_attn_fwd_proj(const void* q, const void* k, const void* v, void* out, float* out_f32, float* lse2, const void* w,
                           const float* bias, void* proj_out, int B, int N, int M, float scale, float drop_p,
                           const unsigned long long* drop_seed, unsigned drop_site, cudaStream_t stream);
int sam2b200_attn_fwd_v64_proj(const void* q, const void* k, const void* memv, void* out64, float* out64_f32, float* lse2,
                               float* rowsum_drop, const void* w, const float* bias, const float* rank1, void* proj_out, int B, int N,
                               int M, float scale, float drop_p, const unsigned long long* drop_seed, unsigned drop_site,
                               cudaStream_t stream);

/* Cross-attention on the RAW 64-d memory features (kv_in_dim = 64: memory_attention.py:66-81 with the cross-attention of
 * configs/sam2/sam2.1_hiera_t.yaml:41-50).  softmax rows sum to 1, hence softmax(q k^T) (memv Wv^T + bv) = out64 Wv^T + bv
 * with out64 = softmax(q k^T) memv: v_proj (transformer.py:279) is applied by the caller to the [B N, 64] result instead of
 * to the [B M, 64] memory -- a quarter of the PV / dP FLOPs, no [B, M, 256] value tensor, no dV kernel.
 * fwd: q [B,N,256], k [B,M,256] (rotated), memv [B,M,64] bf16 -> out64 [B,N,64] bf16 (+ optional fp32 copy), lse2 [B,N].
 * bwd: dout64 = dO Wv [B,N,64] bf16; delta = rowsum(dout64 o out64) [B,N] fp32 (caller); parts 4 = dK, 8 = dQ; the
 * remaining arguments as sam2b200_attn_bwd_ex.
 * Attention-probability dropout (drop_p > 0 with a device seed, transformer.py:304-306): the rows of the dropped matrix do
 * not sum to 1, so the forward also returns rowsum_drop [B,N] (out = out64 Wv^T + rowsum_drop bv) and the backward takes
 * dp_bias [B,N] = dO . bv (added to dP before the mask); delta = rowsum(dout64 o out64) + dp_bias * rowsum_drop. */
int sam2b200_attn_fwd_v64(const void* q, const void* k, const void* memv, void* out64, float* out64_f32, float* lse2,
                          float* rowsum_drop, int B, int N, int M, float scale, float drop_p,
                          const unsigned long long* drop_seed, unsigned drop_site, sam2b200_stream_t stream);
int sam2b200_attn_bwd_v64(const void* q, const void* k, const void* memv, const void* dout64, const float* lse2,
                          const float* delta, void* dq, void* dk, int grad_dtype, int ldq, int ldk, const float* rope_table,
                          int rope_period, int n_rope_k, int B, int N, int M, float scale, float* dbias_q, float* dbias_k,
                          int parts, const float* dp_bias, float drop_p, const unsigned long long* drop_seed,
                          unsigned drop_site, sam2b200_stream_t stream);

/* ---- fused LayerNorm / residual / bias-gradient kernels (d_model = 256) -------------------
 * Replace nn.LayerNorm + residual add + dropout(0) + dtype casts of MemoryAttentionLayer
 * (sam2_video/model/modeling/memory_attention.py:58-99, :162) and what autograd derives for them.
 * ln_fwd: x' = x + res (res bf16, optional; x' stored to x_out if given), y = LN(x') as bf16 and/or
 * fp32, mean / rstd saved for the backward.  tr_b > 0 writes y_f32 seq-first [n][b][256] from
 * batch-first rows b*tr_n + n (the transpose at memory_attention.py:164-167). */
int sam2b200_ln_fwd(const float* x, const void* res_bf16, float* x_out, const float* gamma, const float* beta,
                    void* y_bf16, float* y_f32, float* mean, float* rstd, long long rows, float eps, int tr_b,
                    int tr_n, float drop_p, const unsigned long long* drop_seed, unsigned drop_site,
                    sam2b200_stream_t stream);
size_t sam2b200_ln_bwd_workspace_bytes(long long rows);
/* g_out = g_in + dLN/dx(dy); dgamma += sum dy*xhat; dbeta += sum dy.  Exactly one of dy_bf16 / dy_f32.
 * g_out_bf16 + dbias (optional, both or neither): bf16 copy of g_out (operand of the next GEMMs of the backward) and
 * dbias += its column sums (bias gradient of the projection whose output gradient g_out is). */
int sam2b200_ln_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* mean, const float* rstd,
                    const float* gamma, const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta,
                    float* dbias, void* workspace, long long rows, int tr_b, int tr_n, float drop_p,
                    const unsigned long long* drop_seed, unsigned drop_site, sam2b200_stream_t stream);
/* The same in two separately launchable stages (stages: bit 0 = row pass, bit 1 = fold of the per-block partial sums into dgamma / dbeta /
 * dbias): the fold feeds parameter gradients only and may be enqueued on another stream behind an event. */
int sam2b200_ln_bwd_stages(const void* dy_bf16, const float* dy_f32, const float* x, const float* mean, const float* rstd,
                           const float* gamma, const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta,
                           float* dbias, void* workspace, long long rows, int tr_b, int tr_n, float drop_p,
                           const unsigned long long* drop_seed, unsigned drop_site, int stages, sam2b200_stream_t stream);
size_t sam2b200_colsum_workspace_bytes(long long rows, int C);
/* Bias gradients (what autograd's sum over rows produces for nn.Linear):
 * mode 0: in_f32 [R,C] -> io_bf16 (cast) and colsum += column sums; mode 1: io_bf16 *= (h_bf16 > 0) in
 * place (ReLU backward, memory_attention.py:97) and colsum += sums; mode 2: colsum += sums of io_bf16. */
int sam2b200_colsum(int mode, const float* in_f32, void* io_bf16, const void* h_bf16, float* colsum, void* workspace,
                    long long rows, int C, long long ld, float scale /* mode 1: 1/(1-p) of the dropout after the ReLU */,
                    sam2b200_stream_t stream);

/* Parameter gradients of the folded cross-attention projection ca = out64 (Wo Wv)^T + (Wo bv + bo) of the raw-memory path (autograd
 * of out_proj(v_proj(.)), sam2_video/model/modeling/sam/transformer.py:279,308-309): with G = dca^T out64 [256, 64], g = colsum(dca),
 * g_rs (= g without attention dropout): dWo += G Wv^T + g_rs (x) bv, dWv += Wo^T G, dbo += g, dbv += Wo^T g_rs.  All fp32, in place. */
int sam2b200_fold_grads(const float* G, const float* g, const float* g_rs, const float* Wo, const float* Wv, const float* bv,
                        float* dWo, float* dWv, float* dbo, float* dbv, sam2b200_stream_t stream);

/* Row permutation between the module's seq-first [L, B, C] tensors and the stack's batch-first [B, L, C] rows with fused
 * add / scale / cast (C = 256 or 64; memory_attention.py:140-148, :75-76 and the transposes of the input gradients):
 * inverse 0: out[b,l] = scale * (a[l,b] + alpha2 * a2[l,b]), out2[b,l] = scale * a[l,b]; inverse 1: the other way.
 * a2 / out2 may be NULL; out_bf16 = 1 writes bf16 (memk = bf16(memory + pos), memv = bf16(memory) in one pass). */
int sam2b200_permute_rows(const float* a, const float* a2, float alpha2, void* out, void* out2, int out_bf16, int B, int L,
                          int C, int inverse, float scale, sam2b200_stream_t stream);

/* ---- dropout (train mode; memory_attention.py:58-99, transformer.py:304-306) ----------------------------------
 * Every dropout of the path is a counter-based mask: keep <=> hash(*drop_seed, drop_site, element index) >= p * 2^32
 * (csrc/dropout.cuh); the seed lives in DEVICE memory (fresh masks under CUDA-graph replay), nothing is stored, the
 * backward entry points regenerate the forward's mask from the same (seed, site).  drop_p = 0 or drop_seed = NULL: off.
 * ln_fwd: x' = x + dropout(res); ln_bwd: the bf16 branch-gradient copy is masked (not g_out); attn_fwd_ex /
 * attn_bwd_ex: attention-probability dropout, index (b N + query) M + key; dropout_inplace: the MLP's hidden
 * activation; dropout_mask: test aid, the keep mask as bytes. */
int sam2b200_dropout_inplace(void* x_bf16, long long n, float drop_p, const unsigned long long* drop_seed,
                             unsigned drop_site, sam2b200_stream_t stream);
int sam2b200_dropout_mask(unsigned char* out, long long index0, long long n, float drop_p,
                          const unsigned long long* drop_seed, unsigned drop_site, sam2b200_stream_t stream);

/* ---- fused mask loss --------------------------------------------------------------------
 * mode SAM2B200_LOSS_MULTISTEP replaces MultiStepMultiMasksAndIous._update_losses and the three
 * loss functions it calls (sam2_video/model/losses.py:20-76,143-238) for one step / one mask per
 * channel; mode SAM2B200_LOSS_BCE replaces BCECategoryLoss.forward's per-frame body (:308-372).
 * logits: HOST array of T device pointers, each [C, HW] fp32 (the per-frame tensors are used in
 * place, never stacked); targets: [T, C, HW] u8/bool; iou_pred: [T, C] fp32 (multistep only);
 * pos_weight: [C] fp32 or NULL (bce only).
 * Outputs: chan_sums [T, C, 6] (sum focal|bce, sum p*t, sum p, sum t, |pred&gt|, |pred|gt|),
 * n_valid [T] (channels with foreground; 0 => the caller raises "No valid masks"),
 * losses [4]: multistep: loss_mask, loss_dice, loss_iou, 0 summed over frames (losses.py:116-119);
 * bce: losses[0] = sum over frames of the per-frame reduced loss.
 * workspace: sam2b200_mask_loss_workspace_bytes() bytes = per-block records + the ticket counters of the in-kernel
 * finalisation.  The tickets are zeroed by a memset node on `stream` unless `mode` carries
 * SAM2B200_LOSS_TICKETS_ZEROED: the caller then guarantees a workspace that was zero-filled once and has since
 * only been used by this function on one stream at a time (the kernel leaves the tickets zeroed). */
#define SAM2B200_LOSS_TICKETS_ZEROED 0x100
size_t sam2b200_mask_loss_workspace_bytes(int T, int C, long long HW);
int sam2b200_mask_loss_fwd(const float* const* logits, const uint8_t* targets, const float* iou_pred,
                           const float* pos_weight, void* workspace, float* chan_sums, int* n_valid,
                           float* losses, int T, int C, long long HW, int mode, float alpha,
                           float gamma, float inv_temp, int iou_l1, int reduction_mean,
                           sam2b200_stream_t stream);
/* grad_losses: device [3] = d/d(loss_mask, loss_dice, loss_iou) (multistep) or [1] (bce).
 * dlogits: HOST array of T device pointers [C, HW] fp32, fully overwritten; diou: [T, C]. */
int sam2b200_mask_loss_bwd(const float* const* logits, float* const* dlogits, const uint8_t* targets,
                           const float* iou_pred, const float* pos_weight, const float* chan_sums,
                           const int* n_valid, const float* grad_losses, float* diou, int T, int C,
                           long long HW, int mode, float alpha, float gamma, float inv_temp,
                           int iou_l1, int reduction_mean, sam2b200_stream_t stream);

/* The same pair with ONE TARGET POINTER PER FRAME (target_ptrs: HOST array of T device pointers to [C, HW] bytes): the frames of
 * several clips of a step -- whose targets are different tensors (sam2_video/training/trainer.py:268 calls the criterion once per
 * batch element) -- go through one launch (up to 128 frames per launch); losses = the sum over all frames, i.e. over the clips. */
int sam2b200_mask_loss_fwd_frames(const float* const* logits, const uint8_t* const* target_ptrs, const float* iou_pred,
                                  const float* pos_weight, void* workspace, float* chan_sums, int* n_valid,
                                  float* losses, int T, int C, long long HW, int mode, float alpha,
                                  float gamma, float inv_temp, int iou_l1, int reduction_mean, sam2b200_stream_t stream);
int sam2b200_mask_loss_bwd_frames(const float* const* logits, float* const* dlogits, const uint8_t* const* target_ptrs,
                                  const float* iou_pred, const float* pos_weight, const float* chan_sums,
                                  const int* n_valid, const float* grad_losses, float* diou, int T, int C,
                                  long long HW, int mode, float alpha, float gamma, float inv_temp,
                                  int iou_l1, int reduction_mean, sam2b200_stream_t stream);

/* Backward of the six per-channel sums themselves -- what the stand-alone dice_loss / sigmoid_focal_loss of
 * sam2_video/model/losses.py:20-57 need: dlogits = coef[c][0] * d(sum_px focal)/dx + (t ? coef[c][1] : coef[c][2]) * p(1-p),
 * coef: device [T, C, 3] fp32 (coef[.][1] = d/d(sum p*t) + d/d(sum p), coef[.][2] = d/d(sum p)); no valid filter. */
int sam2b200_mask_loss_bwd_coef(const float* const* logits, float* const* dlogits, const uint8_t* targets,
                                const float* coef, int T, int C, long long HW, float alpha, float gamma, float inv_temp,
                                sam2b200_stream_t stream);

/* ---- mask loss fused with its producer side ------------------------------------------
 * Replaces, per frame, F.interpolate(low_res.float(), (S, S), "bilinear", align_corners=False)
 * (sam2_video/model/modeling/sam2_base.py:393-399), merge_object_results_to_category (sam2_video/utils/masks.py:53-212:
 * pixel-wise max over the objects of a category, IoU predictions averaged with the un-detached area weights
 * sum(sigmoid(high-res logits)); empty category -> zeros) and MultiStepMultiMasksAndIous (losses.py:112-248, one step, one
 * mask per channel, pred_obj_scores off).  S = 4 s; the high-resolution logits are never written to memory.
 * low_res: HOST array of T device pointers, each [n_obj, s, s] fp32; targets: [T, C, S, S] u8/bool (4-byte aligned);
 * obj_iou: [T, n_obj] fp32; group_offsets [C + 1] / group_members [n_obj]: DEVICE int32 CSR of the objects of each
 * category in increasing object index (every object in exactly one category, at most 255 objects per category).
 * Outputs: chan_sums [T, C, 6] and n_valid [T] as for sam2b200_mask_loss_fwd; obj_area [T, n_obj] (area weights),
 * cat_iou [T, C] (merged IoU predictions), cat_w [T, C] (sum of weights), losses [4] = loss_mask, loss_dice, loss_iou, 0.
 * T <= 64 per call.  Backward: dlow_res (HOST array of T device pointers [n_obj, s, s], every element written) and
 * d_obj_iou [T, n_obj]; grad_losses: device [3]. */
size_t sam2b200_merged_loss_workspace_bytes(int T, int C, int n_obj, int s);
int sam2b200_merged_loss_fwd(const float* const* low_res, const uint8_t* targets, const float* obj_iou,
                             const int* group_offsets, const int* group_members, void* workspace, float* chan_sums,
                             float* obj_area, float* cat_iou, float* cat_w, int* n_valid, float* losses, int T, int C,
                             int n_obj, int s, float alpha, float gamma, float inv_temp, int iou_l1,
                             sam2b200_stream_t stream);
int sam2b200_merged_loss_bwd(const float* const* low_res, float* const* dlow_res, const uint8_t* targets,
                             const float* obj_iou, const int* group_offsets, const int* group_members,
                             const float* chan_sums, const float* obj_area, const float* cat_iou, const float* cat_w,
                             const int* n_valid, const float* grad_losses, float* d_obj_iou, int T, int C, int n_obj, int s,
                             float alpha, float gamma, float inv_temp, int iou_l1, sam2b200_stream_t stream);

/* ---- memory encoder (sam2_video/model/modeling/memory_encoder.py; csrc/memenc.cu), channels-last fp32 ------------------
 * ln_gelu: y [P, C] = act(LayerNorm_C(x) * w + b), the LayerNorm2d (sam2_utils.py:141-153) + GELU pair that follows every
 * strided convolution of MaskDownSampler (memory_encoder.py:38-53); C in {4, 16, 64, 256}; act != 0 = exact GELU.
 * The backward writes dx and ADDS the parameter gradients to dw / db [C].
 * dwconv7: depth-wise 7 x 7 convolution, padding 3, of CXBlock (memory_encoder.py:84-91) on x [B, H, W, C], w [C, 7, 7];
 * flip != 0 mirrors the taps (= the data gradient when called on dy); dwconv7_bwd_w ADDS dw [C, 7, 7] and db [C] (per-block
 * partial sums in `workspace`, folded in a fixed order: deterministic). */
int sam2b200_ln_gelu_fwd(const float* x, const float* w, const float* b, float* y, long long P, int C, float eps, int act,
                         cudaStream_t stream);
int sam2b200_ln_gelu_bwd(const float* dy, const float* x, const float* w, const float* b, float* dx, float* dw, float* db, long long P, int C,
                         float eps, int act, cudaStream_t stream);
int sam2b200_dwconv7(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int C, int flip, cudaStream_t stream);
size_t sam2b200_dwconv7_bwd_w_workspace_bytes(int B, int H, int C);
int sam2b200_dwconv7_bwd_w(const float* dy, const float* x, float* dw, float* db, void* workspace, int B, int H, int W, int C,
                           cudaStream_t stream);

/* ---- weight gradients: c [Mo, ldc] fp32 += a[R, Mo]^T . b[R, No] (csrc/wgrad.cu) ------------------------------------------
 * dW = dY^T X of every nn.Linear of the stack (sam2_video/model/modeling/memory_attention.py:97, sam/transformer.py:213-216),
 * accumulated IN PLACE into the fp32 gradient (split over R, partial tiles added with vector fp32 reductions: no workspace,
 * no reduce pass).  a, b: bf16 with row strides lda / ldb (elements); Mo a multiple of 256, No = 64 or a multiple of 256.
 * dbias (nullable, fp32 [Mo]) += column sums of a = the bias gradient of the same layer (transformer.py:213-215 q/k/v_proj
 * biases), from 16 extra accumulator columns against a constant operand of ones; needs No = 64 or (No = 256 and Mo <= 768). */
int sam2b200_wgrad(float* c, long long ldc, const void* a, long long lda, const void* b, long long ldb, long long R, int Mo, int No,
                   float* dbias, cudaStream_t stream);

/* ---- dense GEMMs without a LayerNorm in front or a ReLU mask behind (csrc/gemm.cu) -------------------------------------------
 * c[R, No] (bf16, row stride ldc) = a[R, K] (bf16, row stride lda) . B (+ bias), fp32 accumulation; No = 256 or 64, K % 64 == 0.
 *   b_layout 0: b = W[No, K] (row stride ldb), c = a W^T -- nn.Linear forward: linear2 (memory_attention.py:97-98) and the
 *               memory-key projection k_proj (sam/transformer.py:278) with the axial rotation of position_encoding.py:212-239
 *               on the fp32 accumulator when table != NULL (rows with (row % rows_per_item) < n_rope_rows, table row =
 *               position % period; object-pointer rows stay un-rotated, transformer.py:296-302);
 *   b_layout 1: b = W[K, No] (row stride ldb), c = a W   -- nn.Linear input gradient dX = dY W (autograd of the same lines).
 * bias: [No] fp32 or NULL.  dot_rows [R, 64] fp32 / dot_out [R] fp32 (No = 64, both or neither): dot_out[r] = sum_c bf16(c[r, c]) *
 * dot_rows[r, c] -- Delta = rowsum(dO' o out64) of the raw-memory cross-attention backward, from the GEMM that produces dO'.
 * Resident CTAs, 4-stage TMA ring across tiles, SS-mode tcgen05.mma, two alternating TMEM
 * accumulators (the epilogue of tile i overlaps the MMAs of tile i + 1), TMA-store epilogue. */
int sam2b200_gemm(void* c, long long ldc, const void* a, long long lda, const void* b, long long ldb, int b_layout, long long R, int K,
                  int No, const float* bias, const float* table, int rows_per_item, int n_rope_rows, int period, const float* dot_rows,
                  float* dot_out, cudaStream_t stream);
/* The same kernel with the epilogues of the pre-norm block heads (behind the stand-alone LayerNorm pass sam2b200_ln_fwd): C[R, Nout],
 * Nout = 64 or a multiple of 256 up to 2048, split over n_out = Nout / out_width <= 3 outputs [R, out_width] (q | k | v); the leading
 * rope_cols columns (multiple of 256) rotated as above (transformer.py:296-302); relu != 0: max(., 0) on the biased accumulator, then
 * inverted dropout (drop_p, element index row * Nout + column) -- linear1 + activation + dropout of memory_attention.py:95-97. */
/* host only (no device): kernel variant and shared-memory layout sam2b200_gemm_ex would use on a GPU with `sms` SMs.  out[8] = {column
 * block width, row tiles per item, epilogue groups, ring slots, staging boxes per warp, grid, resident weights (0 | 1), dynamic shared
 * memory bytes}; returns 0 or the layout error. */
int sam2b200_gemm_plan(long long R, int K, int Nout, int rope_cols, int period, int sms, long long* out);
/* debug aid: per-CTA %globaltimer phase stamps of the following sam2b200_gemm* launches (32 x u64 per CTA); NULL = off */
long long sam2b200_gemm_debug_timeline(void* buf, long long n_u64);
int sam2b200_gemm_ex(void* out0, void* out1, void* out2, int out_width, long long ldc, const void* a, long long lda, const void* b,
                     long long ldb, int b_layout, long long R, int K, int Nout, const float* bias, int rope_cols, const float* table,
                     int rows_per_item, int n_rope_rows, int period, int relu, float drop_p, const unsigned long long* drop_seed,
                     unsigned drop_site, const float* dot_rows, float* dot_out, cudaStream_t stream);

/* Two such products that share the long operand, in one pass over it: c [256, ldc] += a^T b and c2 [256, ldc2] += a^T b2 with a [R, 256],
 * b, b2 [R, 64].  Used for the cross-attention key projection's weight gradient (transformer.py:214, b = the key source) together with the
 * per-segment sums of the key gradient (b2 = one-hot bank-segment indicator) that the packed bank's position tensors need
 * (sam2_video/model/modeling/sam2_base.py:618-625, 667-671: maskmem_tpos_enc / obj_ptr_tpos_proj).  dbias as above. */
int sam2b200_wgrad2(float* c, long long ldc, float* c2, long long ldc2, const void* a, long long lda, const void* b, long long ldb,
                    const void* b2, long long ldb2, long long R, float* dbias, cudaStream_t stream);

/* ---- LayerNorm + projection (+ RoPE | ReLU) in one kernel (csrc/lnproj.cu) ------------------------------------------
 * The head of every pre-norm block of MemoryAttentionLayer (sam2_video/model/modeling/memory_attention.py:58-64, 66-81,
 * 95-97 with the projections of sam2_video/model/modeling/sam/transformer.py:277-302):
 *   x_out = x + dropout(res);  y = LayerNorm(x_out) * gamma + beta;  OUT = epi(y . w^T + bias)
 * x [R, 256] fp32; res [R, 256] bf16 or NULL; y_out [R, 256] bf16 or NULL; mean / rstd [R]; w [Nout, 256] bf16;
 * n_out = Nout / out_width <= 3 outputs [R, out_width] bf16.  The leading rope_cols output columns (multiple of 128) are
 * rotated with the axial table [period, 128] (cos, sin) for rows with (row mod rows_per_item) < n_rope_rows; relu != 0
 * applies max(., 0) and then dropout (index row * Nout + column).  Replaces ln_fwd + addmm (x1..3) + rope_apply (x0..2). */
int sam2b200_ln_proj(const float* x, const void* res, float* x_out, const float* gamma, const float* beta, void* y_out,
                     float* mean, float* rstd, long long R, float eps, const void* w, const void* bias, int Nout, void* out0,
                     void* out1, void* out2, int out_width, int rope_cols, const float* table, int rows_per_item,
                     int n_rope_rows, int period, int relu, float drop_res_p, const unsigned long long* drop_res_seed,
                     unsigned drop_res_site, float drop_out_p, const unsigned long long* drop_out_seed, unsigned drop_out_site,
                     cudaStream_t stream);

/* ---- projections with fused bias + axial RoPE (transformer.py:277-279, :296-302; position_encoding.py:212-239) ------
 * Y[R, Nout] = X[R, K] . W[Nout, K]^T + bias (bf16, fp32 accumulation; K = 256 or 64), written to up to three contiguous
 * [R, 256] outputs (q | k | v); the first rope_cols output columns (multiple of 256) are rotated in the GEMM epilogue
 * for rows whose position (row % rows_per_item) < n_rope_rows, table row = position % period.  One launch instead of
 * addmm x Nout/256 + the RoPE pass. */
int sam2b200_proj_rope(const void* x, const void* w, const void* bias, void* out0, void* out1, void* out2, long long R, int K,
                       int Nout, int rope_cols, const float* table, int rows_per_item, int n_rope_rows, int period,
                       sam2b200_stream_t stream);

/* ---- MLP backward (memory_attention.py:95-98) ---------------------------------------------------------------
 * dh[R, F] = (dm[R, 256] . W2[256, F]) o (h[R, F] > 0) * scale: input gradient of linear2 with the ReLU (and hidden
 * dropout: h is the dropped activation, scale = 1/(1-p)) backward fused into the epilogue of a tcgen05 GEMM.
 * All bf16 row-major contiguous, F a multiple of 128. */
int sam2b200_mlp_dh(const void* dm, const void* w2, const void* h, void* dh, float* dbias, long long R, int F, float scale,
                    cudaStream_t stream);

/* ---- memory-bank assembly -------------------------------------------------
 * The data movement of SAM2Base._prepare_memory_conditioned_features (sam2_base.py:597-692) in one launch: for each
 * selected past frame s: memory[s*HW + t, b, :] = feats[s][b, :, t], memory_pos[...] = pos[s][b, :, t] + tpos[s][:];
 * object pointer i, chunk c: memory[n_slots*HW + i*C/64 + c, b, :] = ptrs[i][b, 64c : 64c+64], memory_pos[...] = obj_pos[i].
 * feats / pos / tpos / ptrs: HOST arrays of device pointers ([B, 64, HW], [B, 64, HW], [64] fp32, [B, C]); dtype 0 = fp32,
 * 1 = bf16; memory, memory_pos: [n_slots*HW + n_ptrs*C/64, B, 64] fp32.  Frame selection stays on the host. */
int sam2b200_bank_gather(const void* const* feats, const void* const* pos, const float* const* tpos, int n_slots,
                         int feat_dtype, const void* const* ptrs, int n_ptrs, int ptr_dtype, const float* obj_pos,
                         float* memory, float* memory_pos, int B, int HW, int mem_dim, int C, sam2b200_stream_t stream);
/* Same inputs; the bank is written directly in the layout the fused stack reads: memk = bf16(feat + pos + tpos), memv =
 * bf16(feat), both [B, M, 64] batch-first (no fp32 [M, B, 64] tensors, no re-pack pass). */
int sam2b200_bank_gather_packed(const void* const* feats, const void* const* pos, const float* const* tpos, int n_slots,
                         int feat_dtype, const void* const* ptrs, int n_ptrs, int ptr_dtype, const float* obj_pos,
                         void* memk, void* memv, int B, int HW, int mem_dim, int C, sam2b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SAM2_B200_H_ */
