// C-ABI entry points (include/crossscore_b200.h) and host-side plumbing shared by the kernels.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/crossscore_b200.h"
#include "xs_common.cuh"

namespace xs {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      set_last_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, Swizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return -2;
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz == SWZ_128B ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == SWZ_64B ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu] stride0 %llu box [%u,%u]",
                   (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides_bytes[0], box[0],
                   rank > 1 ? box[1] : 0);
    return -3;
  }
  return 0;
}

// kernels (defined in the other translation units)
int gemm_tc(const void*, int, const void*, int, const float*, void*, int, int, int, int, int, int, int, cudaStream_t);
int gemm_tc_residual_stats(const void*, int, const void*, int, const float*, float*, int, void*, int, float*, float, int,
                           int, int, cudaStream_t);
int gemm_tc_ln_folded(const void*, int, const void*, int, const float*, const float*, const float*, void*, int, int, int,
                      int, int, cudaStream_t);
int rows_stats(const float*, void*, float*, float, int, cudaStream_t);
int gemm_tc_residual_ln(const void*, int, const void*, int, const float*, float*, int, const float*, const float*, float,
                        void*, int, int, int, int, cudaStream_t);
int head_jigsaw_tc(const void*, int, const void*, int, const float*, float*, int, int, int, int, int, float, int,
                   cudaStream_t);
int flash_attn_bf16_tc(const void*, const void*, const void*, void*, float*, int, int, int, int, int, long long,
                       long long, long long, long long, int, int, int, float, cudaStream_t);
void attn_set_optimistic(int);
void attn_set_layout(int);
int gemm_f32(const float*, int, const float*, int, const float*, float*, int, int, int, int, int, cudaStream_t);
int preprocess_u8(const uint8_t*, int, int, int, float*, int, int, const float*, cudaStream_t);
size_t postprocess_workspace_bytes(int);
int postprocess_score(const float*, int, int, int, float*, uint16_t*, int, uint8_t*, float, float, void*, size_t,
                      cudaStream_t);
int head_jigsaw_f32(const float*, int, const float*, int, const float*, float*, int, int, int, int, int, float,
                    cudaStream_t);
int flash_attn_f32(const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, long long,
                   long long, long long, long long, int, int, float, cudaStream_t);
int rows_add_ln(const float*, const void*, float*, const float*, const float*, float, void*, float*, int, int,
                cudaStream_t);
int rows_embed_ln(const void*, int, const float*, const float*, float*, const float*, const float*, float, void*, int,
                  int, int, cudaStream_t);
int rows_final_ln_pe(const float*, const void*, const float*, const float*, float, const float*, float*, void*, void*,
                     int, int, int, int, cudaStream_t);
int rows_im2col14(const float*, void*, int, int, int, int, int, cudaStream_t);
int table_bilinear_ac(const float*, float*, int, int, int, int, int, cudaStream_t);
int table_bicubic(const float*, float*, int, int, int, int, int, float, float, cudaStream_t);
int rows_lse_merge(const float*, const float*, void*, float*, int, int, int, int, int, long long, long long, int,
                   cudaStream_t);
int rows_lse_merge_peers(const void* const*, long long, long long, void*, float*, int, int, int, int, int, int,
                         cudaStream_t);
int rows_attn_probs(const void*, const void*, const float*, float*, int, int, int, int, int, int, int, long long,
                    long long, long long, long long, float, int, cudaStream_t);

static inline int kpad_for(int dtype) { return dtype == XS_BF16 ? 592 : 588; }
static inline int elem_bytes(int dtype) { return dtype == XS_BF16 ? 2 : 4; }

}  // namespace xs

using namespace xs;

extern "C" {

int xs_version(void) { return XS_ABI_VERSION; }

const char* xs_last_error(void) { return g_err; }

int xs_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  XS_CUDA(cudaGetDevice(&dev));
  XS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  XS_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  XS_CHECK_ARG(major == 10, "crossscore_b200 is built for sm_100a (B200); device %d is sm_%d%d", dev, major, minor);
  return 0;
}

size_t xs_workspace_bytes(int op, int a, int b, int c, int dtype) {
  if (op == XS_OP_PATCH_EMBED) {
    const size_t P = (size_t)(b / 14) * (size_t)(c / 14);
    return (size_t)a * P * (size_t)kpad_for(dtype) * (size_t)elem_bytes(dtype);
  }
  if (op == XS_OP_SCORE_POSTPROCESS) return postprocess_workspace_bytes(a);
  return 0;
}

int xs_patch_embed(const float* img, const void* w, const float* bias, void* tok, void* workspace,
                   size_t workspace_bytes, int n_images, int H, int W, int dtype, xs_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ph = H / 14, pw = W / 14;
  XS_CHECK_ARG(n_images > 0 && ph > 0 && pw > 0, "patch_embed: bad dims I=%d H=%d W=%d", n_images, H, W);
  XS_CHECK_ARG(workspace_bytes >= xs_workspace_bytes(XS_OP_PATCH_EMBED, n_images, H, W, dtype),
               "patch_embed: workspace too small (%zu bytes)", workspace_bytes);
  XS_CHECK_ARG(dtype == XS_BF16 || dtype == XS_F32 || dtype == XS_TF32, "patch_embed: unknown dtype %d", dtype);
  const int Kpad = kpad_for(dtype);
  int rc = rows_im2col14(img, workspace, n_images, H, W, Kpad, dtype == XS_BF16 ? XS_BF16 : XS_F32, st);
  if (rc) return rc;
  const int M = n_images * ph * pw;
  if (dtype == XS_BF16) return gemm_tc(workspace, Kpad, w, Kpad, bias, tok, 384, M, 384, Kpad, ACT_NONE, 0, 0, st);
  if (dtype == XS_TF32) return gemm_tc(workspace, Kpad, w, Kpad, bias, tok, 384, M, 384, Kpad, ACT_NONE, 1, 1, st);
  return gemm_f32(static_cast<const float*>(workspace), Kpad, static_cast<const float*>(w), Kpad, bias,
                  static_cast<float*>(tok), 384, M, 384, Kpad, ACT_NONE, st);
}

int xs_embed_cls_pos_ln(const void* tok, int tok_dtype, const float* cls, const float* pos, float* h,
                        const float* gamma, const float* beta, float eps, void* y, int n_images, int P, int dtype,
                        xs_stream_t stream) {
  return rows_embed_ln(tok, tok_dtype, cls, pos, h, gamma, beta, eps, y, n_images, P, dtype,
                       static_cast<cudaStream_t>(stream));
}

int xs_layernorm(const float* res_in, const void* delta, float* res_out, const float* gamma, const float* beta,
                 float eps, void* y, float* y32, int rows, int dtype, xs_stream_t stream) {
  return rows_add_ln(res_in, delta, res_out, gamma, beta, eps, y, y32, rows, dtype, static_cast<cudaStream_t>(stream));
}

int xs_final_ln_drop_cls_add_pe(const float* h, const void* delta, const float* gamma, const float* beta, float eps,
                                const float* pe, float* xq32, void* xq, void* mem, int n_images,
                                int n_query_images, int P, int dtype, xs_stream_t stream) {
  return rows_final_ln_pe(h, delta, gamma, beta, eps, pe, xq32, xq, mem, n_images, n_query_images, P, dtype,
                          static_cast<cudaStream_t>(stream));
}

int xs_pe_resample_bilinear_ac(const float* table, float* out, int ih, int iw, int oh, int ow, int channels,
                               xs_stream_t stream) {
  return table_bilinear_ac(table, out, ih, iw, oh, ow, channels, static_cast<cudaStream_t>(stream));
}

int xs_pos_embed_resample_bicubic(const float* table, float* out, int ih, int iw, int oh, int ow, int channels,
                                  xs_stream_t stream) {
  return table_bicubic(table, out, ih, iw, oh, ow, channels, 0.f, 0.f, static_cast<cudaStream_t>(stream));
}

int xs_pos_embed_resample_bicubic_steps(const float* table, float* out, int ih, int iw, int oh, int ow, int channels,
                                        float step_h, float step_w, xs_stream_t stream) {
  return table_bicubic(table, out, ih, iw, oh, ow, channels, step_h, step_w, static_cast<cudaStream_t>(stream));
}

int xs_gemm_bias_act(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, int M,
                     int N, int K, int act, int dtype, int out_dtype, xs_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XS_CHECK_ARG(out_dtype == XS_BF16 || out_dtype == XS_F32 || out_dtype == XS_F16,
               "gemm: out_dtype must be bf16, fp16 or fp32");
  if (dtype == XS_BF16 || dtype == XS_TF32)
    return gemm_tc(A, lda, W, ldw, bias, out, ldc, M, N, K, act, dtype == XS_TF32,
                   out_dtype == XS_F32 ? 1 : (out_dtype == XS_F16 ? 3 : 0), st);
  if (dtype == XS_F32) {
    XS_CHECK_ARG(out_dtype == XS_F32, "gemm(fp32): output must be fp32");
    return gemm_f32(static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw, bias,
                    static_cast<float*>(out), ldc, M, N, K, act, st);
  }
  set_last_error("gemm: unknown dtype %d", dtype);
  return -1;
}

int xs_gemm_bias_residual(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int ldh, int M,
                          int N, int K, int dtype, xs_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XS_CHECK_ARG(dtype == XS_BF16, "gemm_bias_residual: bf16 operands only (the fp32 parity mode adds in xs_layernorm)");
  return gemm_tc(A, lda, W, ldw, bias, h, ldh, M, N, K, ACT_NONE, 0, 2, st);
}

int xs_gemm_bias_residual_ln(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int ldh,
                             const float* gamma, const float* beta, float eps, void* y, int ldy, int M, int N, int K,
                             int dtype, xs_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XS_CHECK_ARG(dtype == XS_BF16, "gemm_bias_residual_ln: bf16 operands only");
  XS_CHECK_ARG(N == 384 && ldy >= N && ldh >= N, "gemm_bias_residual_ln: N must be 384 (LayerNorm width), got %d", N);
  const int rc = gemm_tc_residual_ln(A, lda, W, ldw, bias, h, ldh, gamma, beta, eps, y, ldy, M, N, K, st);
  if (rc != 1) return rc;
  // shapes the fused kernel does not cover (few rows): the same two steps as separate launches
  const int r2 = gemm_tc(A, lda, W, ldw, bias, h, ldh, M, N, K, ACT_NONE, 0, 2, st);
  if (r2) return r2;
  XS_CHECK_ARG(ldh == 384 && ldy == 384, "gemm_bias_residual_ln: the unfused path needs dense rows");
  return rows_add_ln(h, nullptr, nullptr, gamma, beta, eps, y, nullptr, M, XS_BF16, st);
}

int xs_gemm_bias_residual_stats(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int ldh,
                                void* hb, int ldhb, float* stats, float eps, int M, int N, int K, int dtype,
                                xs_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XS_CHECK_ARG(dtype == XS_BF16, "gemm_bias_residual_stats: bf16 operands only");
  XS_CHECK_ARG(N == 384 && ldhb >= N && ldh >= N, "gemm_bias_residual_stats: N must be 384 (LayerNorm width), got %d", N);
  const int rc = gemm_tc_residual_stats(A, lda, W, ldw, bias, h, ldh, hb, ldhb, stats, eps, M, N, K, st);
  if (rc != 1) return rc;
  // shapes the fused epilogue does not cover (few rows): the same result in two launches
  const int r2 = gemm_tc(A, lda, W, ldw, bias, h, ldh, M, N, K, ACT_NONE, 0, 2, st);
  if (r2) return r2;
  XS_CHECK_ARG(ldh == 384 && ldhb == 384, "gemm_bias_residual_stats: the unfused path needs dense rows");
  return rows_stats(h, hb, stats, eps, M, st);
}

int xs_row_stats(const float* h, void* hb, float* stats, float eps, int rows, xs_stream_t stream) {
  return rows_stats(h, hb, stats, eps, rows, static_cast<cudaStream_t>(stream));
}

int xs_gemm_ln_folded(const void* A, int lda, const void* W, int ldw, const float* c0, const float* c1,
                      const float* stats, void* out, int ldc, int M, int N, int K, int act, int dtype,
                      xs_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XS_CHECK_ARG(dtype == XS_BF16, "gemm_ln_folded: bf16 operands only");
  return gemm_tc_ln_folded(A, lda, W, ldw, c0, c1, stats, out, ldc, M, N, K, act, st);
}

int xs_flash_attn(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq, int Lk,
                  int head_dim, int head_slot, long long q_row_stride, long long q_batch_stride,
                  long long kv_row_stride, long long kv_batch_stride, int kv_shared, int nsplit, int o_is_f32,
                  float scale, int dtype, xs_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XS_BF16) {
    XS_CHECK_ARG(head_slot == 64, "flash_attn(bf16): head_slot must be 64, got %d", head_slot);
    return flash_attn_bf16_tc(q, k, v, o, lse, B, heads, Lq, Lk, head_dim, q_row_stride, q_batch_stride,
                              kv_row_stride, kv_batch_stride, kv_shared, nsplit, o_is_f32, scale, st);
  }
  if (dtype == XS_F32) {
    XS_CHECK_ARG(o_is_f32, "flash_attn(fp32): output must be fp32");
    return flash_attn_f32(static_cast<const float*>(q), static_cast<const float*>(k), static_cast<const float*>(v),
                          static_cast<float*>(o), lse, B, heads, Lq, Lk, head_dim, head_slot, q_row_stride,
                          q_batch_stride, kv_row_stride, kv_batch_stride, kv_shared, nsplit, scale, st);
  }
  set_last_error("flash_attn: unknown dtype %d", dtype);
  return -1;
}

void xs_attn_set_optimistic(int enable) { attn_set_optimistic(enable); }
void xs_attn_set_layout(int layout) { attn_set_layout(layout); }

int xs_lse_merge(const float* o_parts, const float* lse_parts, void* out, float* lse_out, int n_parts, int B, int Lq,
                 int heads, int head_dim, long long o_part_stride, long long lse_part_stride, int dtype,
                 xs_stream_t stream) {
  return rows_lse_merge(o_parts, lse_parts, out, lse_out, n_parts, B, Lq, heads, head_dim, o_part_stride,
                        lse_part_stride, dtype, static_cast<cudaStream_t>(stream));
}

int xs_head_score_jigsaw(const void* A, int lda, const void* W, int ldw, const float* bias, float* score, int B,
                         int ph, int pw, int K, int use_tanh, float power, int dtype, xs_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XS_BF16 || dtype == XS_TF32)
    return head_jigsaw_tc(A, lda, W, ldw, bias, score, B, ph, pw, K, use_tanh, power, dtype == XS_TF32, st);
  if (dtype == XS_F32)
    return head_jigsaw_f32(static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw, bias, score, B, ph,
                           pw, K, use_tanh, power, st);
  set_last_error("head_score_jigsaw: unknown dtype %d", dtype);
  return -1;
}

int xs_lse_merge_peers(const void* const* part_ptrs, long long base_elems, long long lse_offset_elems, void* out,
                       float* lse_out, int n_parts, int B, int Lq, int heads, int head_dim, int out_dtype,
                       xs_stream_t stream) {
  return rows_lse_merge_peers(part_ptrs, base_elems, lse_offset_elems, out, lse_out, n_parts, B, Lq, heads, head_dim,
                              out_dtype, static_cast<cudaStream_t>(stream));
}

int xs_preprocess_u8_resize_normalize(const uint8_t* img, int n, int H0, int W0, float* out, int H1, int W1,
                                      const float* mean_std, xs_stream_t stream) {
  return preprocess_u8(img, n, H0, W0, out, H1, W1, mean_std, static_cast<cudaStream_t>(stream));
}

int xs_score_postprocess(const float* score, int B, int H, int W, float* frame_mean, uint16_t* gray16, int vrange_mode,
                         uint8_t* rgb, float vmin, float vmax, void* workspace, size_t workspace_bytes,
                         xs_stream_t stream) {
  return postprocess_score(score, B, H, W, frame_mean, gray16, vrange_mode, rgb, vmin, vmax, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

int xs_attn_probs_one_head(const void* q, const void* k, const float* lse, float* probs, int B, int heads, int head,
                           int Lq, int Lk, int head_dim, int head_slot, long long q_row_stride,
                           long long q_batch_stride, long long kv_row_stride, long long kv_batch_stride, float scale,
                           int dtype, xs_stream_t stream) {
  return rows_attn_probs(q, k, lse, probs, B, heads, head, Lq, Lk, head_dim, head_slot, q_row_stride, q_batch_stride,
                         kv_row_stride, kv_batch_stride, scale, dtype, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
