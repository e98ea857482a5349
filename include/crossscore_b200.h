/* crossscore_b200 -- C ABI of the B200 (sm_100a) CrossScore inference hot path.
 *
 * The reference (ActiveVisionLab/CrossScore) has no FFI: the path sits behind a PyTorch nn.Module
 * (task/core.py:26-161, CrossScoreNet) whose arithmetic is torch / HF-transformers library calls.
 * This header is the boundary a maintainer binds instead of those calls (ctypes stub in INTEGRATION.md):
 * one entry per fused operator, plain pointers and sizes, no torch types.  Each entry cites the
 * reference call site(s) it replaces (paths relative to the reference root; $SP = site-packages of the
 * reference's pinned torch / transformers).
 *
 * Conventions
 *   - every function returns int: 0 ok, <0 invalid argument / unsupported shape, >0 a cudaError_t value;
 *     xs_last_error() returns the message of the last failure on the calling thread.  Nothing throws
 *     or exits.
 *   - all pointers are DEVICE pointers on the current CUDA device, 16-byte aligned; the caller owns all
 *     memory (PyTorch caching allocator in the shipped host code); kernels are enqueued on `stream`
 *     (a cudaStream_t) with no implicit synchronisation and are CUDA-graph capturable.
 *   - dtype selects the activation storage: XS_DTYPE_BF16 (tcgen05 tensor-core path) or
 *     XS_DTYPE_F32 (fp32 parity mode, SIMT).  The GEMM-type entries also accept XS_DTYPE_TF32: fp32
 *     operands multiplied as TF32 on the tensor cores (the bf16 product mode uses it for the ~2 % of FLOPs
 *     that dominate the bf16 error budget: patch embedding, decoder projections/FFN, head).
 *     Statistics, residual stream, tables, biases and the score map are always fp32.  Hidden size is 384
 *     (DINOv2-small) throughout.
 *   - bf16 attention operands use 64-column head slots: head h of a row lives at columns
 *     [h*64, h*64+head_dim); for head_dim 48 (decoder) the projection weights are padded so the 16
 *     trailing columns of each slot are never read.
 */
#ifndef CROSSSCORE_B200_H_
#define CROSSSCORE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XS_ABI_VERSION 1

#define XS_DTYPE_BF16 0
#define XS_DTYPE_F32 1
#define XS_DTYPE_TF32 2 /* GEMM operand mode only: fp32 in memory, multiplied as TF32 on the tensor cores */
#define XS_DTYPE_F16 3  /* attention operands (q, k, v) in fp16: xs_gemm_bias_act out_dtype, xs_flash_attn /
                           xs_attn_probs_one_head dtype.  The fp16 attention kernel keeps its logits in fp16 tensor-core
                           accumulators, as the reference's 16-mixed autocast path does (config/default_predict.yaml:25) */

#define XS_ACT_NONE 0
#define XS_ACT_GELU 1  /* exact (erf) GELU: Dinov2MLP, $SP/transformers/models/dinov2/modeling_dinov2.py:312-328 */
#define XS_ACT_RELU 2  /* decoder FFN: model/customised_transformer/transformer.py:59,208-210 */
#define XS_ACT_LEAKY 3 /* head LeakyReLU(0.01): model/cross_reference.py:47 */

#define XS_OP_PATCH_EMBED 1
#define XS_OP_SCORE_POSTPROCESS 2 /* xs_workspace_bytes(op, B, 0, 0, 0) */

typedef void* xs_stream_t; /* cudaStream_t */

int xs_version(void);
const char* xs_last_error(void);
/* 0 if the current device is compute capability 10.x (B200), <0 otherwise */
int xs_device_check(void);
/* scratch bytes an operator needs; XS_OP_PATCH_EMBED: (n_images, H, W); XS_OP_SCORE_POSTPROCESS: (B) */
size_t xs_workspace_bytes(int op, int a, int b, int c, int dtype);

/* K1  Dinov2PatchEmbeddings.projection: Conv2d(3,384,k=14,s=14) == im2col + GEMM
 *     ($SP/transformers/models/dinov2/modeling_dinov2.py:139-149).
 *     img (I,3,H,W) fp32;  w: bf16 (384,592) zero-padded K / fp32 (384,588) for F32 and TF32;
 *     tok (I*P,384) bf16 for BF16, fp32 for F32 and TF32. */
int xs_patch_embed(const float* img, const void* w, const float* bias, void* tok, void* workspace,
                   size_t workspace_bytes, int n_images, int H, int W, int dtype, xs_stream_t stream);

/* K1b CLS cat + position embedding add (modeling_dinov2.py:108-112) fused with layer-0 norm1 (:354,371).
 *     h (I*(P+1),384) fp32 residual stream out;  y = LN(h) in activation dtype. */
int xs_embed_cls_pos_ln(const void* tok, int tok_dtype, const float* cls, const float* pos, float* h,
                        const float* gamma, const float* beta, float eps, void* y, int n_images, int P, int dtype,
                        xs_stream_t stream);

/* K2  LayerNorm with fused residual add: x = res_in + delta;  res_out = x (optional);
 *     y = LN(x) (activation dtype, optional);  y32 = LN(x) (fp32, optional).
 *     DINOv2 pre-norm: modeling_dinov2.py:354,359,371-384;  decoder post-norm:
 *     model/customised_transformer/transformer.py:157-173.  res_in or delta may be NULL. */
int xs_layernorm(const float* res_in, const void* delta, float* res_out, const float* gamma, const float* beta,
                 float eps, void* y, float* y32, int rows, int dtype, xs_stream_t stream);

/* K8  final DINOv2 LayerNorm (modeling_dinov2.py:477) + CLS drop + query/reference split
 *     (task/core.py:142-153) + multi-view PE add (model/positional_encoding.py:72-74).
 *     Images are ordered "all queries, then references (b, n)": images [0,n_query_images) ->
 *     xq32 / xq (n_query*P, 384), images [n_query_images, n_images) -> mem ((n_images-n_query)*P, 384).
 *     xq32, xq, mem may each be NULL. */
int xs_final_ln_drop_cls_add_pe(const float* h, const void* delta, const float* gamma, const float* beta,
                                float eps, const float* pe, float* xq32, void* xq, void* mem, int n_images,
                                int n_query_images, int P, int dtype, xs_stream_t stream);

/* PE table (ih,iw,C) -> (oh,ow,C), bilinear align_corners=True (model/positional_encoding.py:61-69) */
int xs_pe_resample_bilinear_ac(const float* table, float* out, int ih, int iw, int oh, int ow, int channels,
                               xs_stream_t stream);
/* DINOv2 pos-emb grid (ih,iw,C) -> (oh,ow,C), bicubic align_corners=False (modeling_dinov2.py:57-95) */
int xs_pos_embed_resample_bicubic(const float* table, float* out, int ih, int iw, int oh, int ow, int channels,
                                  xs_stream_t stream);
/* Same with explicit source steps per output pixel (0 = ih/oh resp. iw/ow).  transformers 4.33.3 -- the version the
 * reference pins (environment.yaml:340) -- calls F.interpolate(scale_factor=((oh+0.1)/ih, (ow+0.1)/iw)), for which
 * ATen steps by 1/scale_factor = ih/(oh+0.1); newer transformers pass size=, i.e. ih/oh. */
int xs_pos_embed_resample_bicubic_steps(const float* table, float* out, int ih, int iw, int oh, int ow, int channels,
                                        float step_h, float step_w, xs_stream_t stream);

/* K3,K5-K7,K9-K11  out[M,N] = act(A[M,K] @ W[N,K]^T + bias[N])   (torch.nn.Linear everywhere on the path)
 *     dtype = operand type (BF16 / TF32 tensor cores, F32 SIMT); out_dtype = BF16 or F32.
 *     Tensor-core modes: N multiple of 192 or 256, row pitches multiples of 16 bytes; bf16->fp32 and
 *     tf32->bf16 support act NONE only.  fp32: any shape, lda/ldw % 4 == 0, fp32 output. */
int xs_gemm_bias_act(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc,
                     int M, int N, int K, int act, int dtype, int out_dtype, xs_stream_t stream);

/* K5,K7  h[M,N] += A[M,K] @ W[N,K]^T + bias[N]   (fp32 residual stream updated in place)
 *     Dinov2Layer's `layer_scale(attention_output) + hidden_states` and `layer_scale2(mlp(...)) + hidden_states`
 *     ($SP/transformers/models/dinov2/modeling_dinov2.py:367-386) with LayerScale folded into W / bias: the add is
 *     performed in the GEMM epilogue (the residual tile is prefetched by TMA while the tile's MMAs run), so the delta
 *     never round-trips through HBM (nor through bf16).
 *     dtype = XS_DTYPE_BF16 only; N multiple of 192; every element of h receives exactly one add (deterministic). */
int xs_gemm_bias_residual(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int ldh,
                          int M, int N, int K, int dtype, xs_stream_t stream);

/* K6+K2, K7+K2  the same residual update followed by the LayerNorm the reference applies next
 *     (modeling_dinov2.py:367-386: norm2 after the attention residual; the next layer's norm1 after the MLP residual):
 *         h (fp32, in place) += A @ W^T + bias;   y (bf16) = LayerNorm(h; gamma, beta, eps)
 *     in ONE kernel when there are enough rows to fill the GPU (the CTA that updates a 128-row block re-reads it from
 *     L2 and emits y; no LayerNorm launch, no HBM read of h), otherwise as the two launches it replaces.  N = 384. */
int xs_gemm_bias_residual_ln(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int ldh,
                             const float* gamma, const float* beta, float eps, void* y, int ldy, int M, int N, int K,
                             int dtype, xs_stream_t stream);

/* K6+K2 / K7+K2 with the LayerNorm FOLDED into the GEMM that consumes it (opt-in plan, XS_FOLD_LN=1 in the host module:
 *     measured slower than the LayerNorm kernels and less precise with outlier channels).  The same two reference steps as above (modeling_dinov2.py:367-386, 312-328, 203-234), split differently:
 *       producer:  h (fp32, in place) += A @ W^T + bias;  hb (bf16) = h;  stats[r] = (mean_r, rstd_r) of row r of h
 *       consumer:  out (bf16) = act( rstd_r * (hb @ Wf^T - mean_r * c1) + c0 )
 *     with Wf = bf16(gamma * W) (LayerNorm weight folded into the next Linear's weight), c1[n] = sum_k Wf[n,k] and
 *     c0 = W beta + b, which equals act(LayerNorm(h) @ W^T + b) up to the rounding of h (instead of LayerNorm(h)) to
 *     bf16.  No LayerNorm kernel, no second pass: the statistics come out of the residual epilogue (one CTA sees whole
 *     rows), the normalisation is two FMAs per element in the consumer's epilogue.  stats: (M, 2) fp32.  N = 384 for
 *     the producer; act = XS_ACT_NONE or XS_ACT_GELU for the consumer.  xs_row_stats is the producer's second half as a
 *     stand-alone pass (what the producer runs for row counts its fused epilogue does not cover). */
int xs_gemm_bias_residual_stats(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int ldh,
                                void* hb, int ldhb, float* stats, float eps, int M, int N, int K, int dtype,
                                xs_stream_t stream);
int xs_row_stats(const float* h, void* hb, float* stats, float eps, int rows, xs_stream_t stream);
int xs_gemm_ln_folded(const void* A, int lda, const void* W, int ldw, const float* c0, const float* c1,
                      const float* stats, void* out, int ldc, int M, int N, int K, int act, int dtype,
                      xs_stream_t stream);

/* K4,K9,K10  O = softmax(Q K^T * scale) V  per (batch, head), no mask
 *     (modeling_dinov2.py:203-234; $SP/torch/nn/functional.py:6630-6692 via transformer.py:182-205).
 *     head_slot: column pitch between heads in q/k/v rows (bf16: must be 64).  kv_shared: all batches
 *     read batch 0 of k/v (scene-level reference cache).  nsplit > 1: o is the fp32 partial buffer
 *     [nsplit][B*Lq][heads*head_dim] and lse [nsplit][B][heads][Lq] must be given (merge with xs_lse_merge).
 *     lse (optional for nsplit == 1) = ln sum_j exp(scale * q.k_j). */
int xs_flash_attn(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq,
                  int Lk, int head_dim, int head_slot, long long q_row_stride, long long q_batch_stride,
                  long long kv_row_stride, long long kv_batch_stride, int kv_shared, int nsplit, int o_is_f32,
                  float scale, int dtype, xs_stream_t stream);

/*     The bf16 kernel's first pass takes 2^(logit * log2 e) with no running maximum (exact while the row sums stay
 *     inside [2^-80, 2^100]); tiles that leave that range are redone with the online softmax in the same launch.
 *     enable = 0 sends every tile through the online-softmax pass (process-wide; default 1). */
void xs_attn_set_optimistic(int enable);
/*     layout 0: 64-key blocks, two CTAs per SM (xs_attn_tc.cu); 1: two query tiles per CTA sharing 128-key K/V blocks
 *     (xs_attn_tc2.cu; the default).  Process-wide; the results are the same up to summation order. */
void xs_attn_set_layout(int layout);

/* C2  split-KV merge: LSE = log sum_r exp(LSE_r), O = sum_r exp(LSE_r - LSE) O_r  (no reference counterpart;
 *     identity in SURVEY.md appendix B-10).  Parts may come from local splits or an NCCL all-gather:
 *     part r of O starts at o_parts + r*o_part_stride, of LSE at lse_parts + r*lse_part_stride (floats;
 *     0 = dense [n_parts][B*Lq][heads*head_dim] / [n_parts][B][heads][Lq]), so one packed all-gather
 *     buffer holding (O_r | LSE_r) per rank merges in place.  A part that saw no keys carries LSE = -inf
 *     and a finite O; it contributes nothing. */
int xs_lse_merge(const float* o_parts, const float* lse_parts, void* out, float* lse_out, int n_parts, int B,
                 int Lq, int heads, int head_dim, long long o_part_stride, long long lse_part_stride, int dtype,
                 xs_stream_t stream);

/* K12  head[2] Linear(384->196) + Sigmoid/Tanh (+pow) + jigsaw_to_image
 *      (model/cross_reference.py:45-50,82-87; model/regression_layer.py:26-62; utils/misc/image.py:8-21).
 *      A (B*ph*pw, 384);  W: (224,384) zero-padded rows for BF16 / TF32, fp32 (196,384) for F32;
 *      score (B,14ph,14pw) fp32. */
int xs_head_score_jigsaw(const void* A, int lda, const void* W, int ldw, const float* bias, float* score, int B,
                         int ph, int pw, int K, int use_tanh, float power, int dtype, xs_stream_t stream);

/* K13  attention probabilities of one head (need_attn_weights=True; transformer.py:175-178,
 *      cross_reference.py:91-93): probs[b,i,j] = exp(scale * q_i.k_j - lse[b,head,i]),  (B,Lq,Lk) fp32 */
int xs_attn_probs_one_head(const void* q, const void* k, const float* lse, float* probs, int B, int heads,
                           int head, int Lq, int Lk, int head_dim, int head_slot, long long q_row_stride,
                           long long q_batch_stride, long long kv_row_stride, long long kv_batch_stride,
                           float scale, int dtype, xs_stream_t stream);

/* C2'  the split-KV merge fused with its collective: part_ptrs is a DEVICE array of n_parts pointers, entry s = the
 *     packed fp32 buffer of rank s (peer memory mapped over NVLink, e.g. torch symmetric memory buffer_ptrs_dev):
 *     normalised partial O (B*Lq*heads*head_dim) at base_elems, its LSE (B*heads*Lq) at base_elems + lse_offset_elems.
 *     Every rank pulls the partials in place while merging: no all-gather, no gathered buffer.  The caller orders the
 *     ranks (all partials written before, all merges done before the buffers are rewritten) with a cross-rank barrier
 *     on the stream. */
int xs_lse_merge_peers(const void* const* part_ptrs, long long base_elems, long long lse_offset_elems, void* out,
                       float* lse_out, int n_parts, int B, int Lq, int heads, int head_dim, int out_dtype,
                       xs_stream_t stream);

/* ---- the steps either side of the path ---------------------------------- */

/* F2  uint8 HWC image(s) -> /255 -> antialiased bilinear resize -> ImageNet normalise -> fp32 NCHW: the dataloader's
 *     per-image CPU work (utils/io/images.py:14-29 image_read/f32; dataloading/dataset/nvs_dataset.py:218-225
 *     resize_all = torchvision T.Resize(short side, BILINEAR, antialias=True), task/predict.py:87-92;
 *     T.Normalize, task/predict.py:69-74) done on the device in front of xs_patch_embed.
 *     img (n, H0, W0, 3) uint8, out (n, 3, H1, W1) fp32; the caller computes (H1, W1) with torchvision's rule
 *     (short side -> size, long side -> int(size * long / short)); H1 == H0 && W1 == W0 skips the filter.
 *     mean_std: 6 floats in HOST memory (mean rgb, std rgb).  Down-scaling factors up to 10. */
int xs_preprocess_u8_resize_normalize(const uint8_t* img, int n, int H0, int W0, float* out, int H1, int W1,
                                      const float* mean_std, xs_stream_t stream);

/* F3  score map (B, H, W) fp32 -> any of: per-frame mean (utils/io/score_summariser.py:180-181), uint16 gray
 *     quantisation (utils/io/images.py:49-63 metric_map_write; vrange_mode 0: [0,1] -> m*65535, 1: [-1,1] ->
 *     (m+1)*32767, truncated), turbo RGB (B, H, W, 3) uint8 (utils/misc/image.py:35-49 gray2rgb with vrange
 *     (vmin, vmax)); the products utils/io/batch_writer.py:114-135,263-270 writes.  Outputs may be NULL.
 *     workspace: xs_workspace_bytes(XS_OP_SCORE_POSTPROCESS, B, 0, 0, 0) bytes when frame_mean is requested (the mean
 *     is a fixed-order fp64 reduction: deterministic). */
int xs_score_postprocess(const float* score, int B, int H, int W, float* frame_mean, uint16_t* gray16, int vrange_mode,
                         uint8_t* rgb, float vmin, float vmax, void* workspace, size_t workspace_bytes,
                         xs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CROSSSCORE_B200_H_ */
