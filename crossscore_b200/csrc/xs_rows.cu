// Memory-bound row kernels: LayerNorm family (fused with the residual add, CLS drop and positional
// encoding), patch im2col, the two position-table resamplers, split-KV merge, attention-probability export.
// All are templated on the activation storage type AT (bf16 for the fast path, float for the fp32 parity
// mode); statistics, residual stream and tables are always fp32.  C = 384 is fixed (DINOv2-small).
//
// Reference call sites replaced: nn.LayerNorm ($SP/transformers/models/dinov2/modeling_dinov2.py:354,359,449;
// model/customised_transformer/transformer.py:78-80,159-173), CLS cat + pos-emb add (modeling_dinov2.py:108-112),
// CLS drop / query-ref split (task/core.py:142-153), MultiViewPosionalEmbeddings
// (model/positional_encoding.py:42-75), bicubic pos-emb resample (modeling_dinov2.py:57-95).
#include <cuda_fp16.h>

#include "xs_common.cuh"

namespace xs {

constexpr int C = 384;
constexpr int ROWS_PER_BLOCK = 8;  // one warp per row, 256 threads

template <typename AT>
struct Pack4;
template <>
struct Pack4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct Pack4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// each lane owns columns {4*lane + 128*i + (0..3)}, i = 0..2
struct Row12 {
  float4 v[3];
};
__device__ __forceinline__ int col_of(int lane, int i) { return 4 * lane + 128 * i; }

__device__ __forceinline__ Row12 ln_normalize(const Row12& x, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) s += x.v[i].x + x.v[i].y + x.v[i].z + x.v[i].w;
  const float mean = warp_sum(s) * (1.0f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a = x.v[i].x - mean, b = x.v[i].y - mean, c = x.v[i].z - mean, d = x.v[i].w - mean;
    ss += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / C) + eps);
  Row12 y;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col_of(lane, i)));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + col_of(lane, i)));
    y.v[i] = make_float4((x.v[i].x - mean) * rstd * g.x + bt.x, (x.v[i].y - mean) * rstd * g.y + bt.y,
                         (x.v[i].z - mean) * rstd * g.z + bt.z, (x.v[i].w - mean) * rstd * g.w + bt.w);
  }
  return y;
}

// ---------------------------------------------------------------------------------------------
// y = LN(res_in + delta);  optionally res_out = res_in + delta (pre-norm residual stream, DINOv2)
// or y32 = y (post-norm decoder: the normalised row is the next residual)
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void __launch_bounds__(256) add_ln_kernel(const float* __restrict__ res_in, const AT* __restrict__ delta,
                                                     float* __restrict__ res_out, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps,
                                                     AT* __restrict__ y, float* __restrict__ y32, int rows) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = static_cast<size_t>(row) * C;
  Row12 x;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res_in) a = *reinterpret_cast<const float4*>(res_in + base + col_of(lane, i));
    if (delta) {
      const float4 d = Pack4<AT>::load(delta + base + col_of(lane, i));
      a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
    }
    x.v[i] = a;
  }
  if (res_out) {
#pragma unroll
    for (int i = 0; i < 3; ++i) *reinterpret_cast<float4*>(res_out + base + col_of(lane, i)) = x.v[i];
  }
  const Row12 yv = ln_normalize(x, gamma, beta, eps, lane);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (y) Pack4<AT>::store(y + base + col_of(lane, i), yv.v[i]);
    if (y32) *reinterpret_cast<float4*>(y32 + base + col_of(lane, i)) = yv.v[i];
  }
}

// ---------------------------------------------------------------------------------------------
// The folded LayerNorm's producer as a stand-alone pass (shapes the fused GEMM epilogue does not cover, and its check):
// hb = bf16(h), stats[row] = (mean, rstd) of the fp32 row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ h, __nv_bfloat16* __restrict__ hb,
                                                        float2* __restrict__ stats, float eps, int rows) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = static_cast<size_t>(row) * C;
  Row12 x;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    x.v[i] = *reinterpret_cast<const float4*>(h + base + col_of(lane, i));
    s += x.v[i].x + x.v[i].y + x.v[i].z + x.v[i].w;
    Pack4<__nv_bfloat16>::store(hb + base + col_of(lane, i), x.v[i]);
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a = x.v[i].x - mean, b = x.v[i].y - mean, c = x.v[i].z - mean, d = x.v[i].w - mean;
    ss += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / C) + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

// ---------------------------------------------------------------------------------------------
// DINOv2 embeddings: h[i,0] = cls + pos[0]; h[i,1+p] = tok[i,p] + pos[1+p];  y = LN(h; layer-0 norm1)
// ---------------------------------------------------------------------------------------------
template <typename AT, typename TT>
__global__ void __launch_bounds__(256) embed_ln_kernel(const TT* __restrict__ tok, const float* __restrict__ cls,
                                                       const float* __restrict__ pos, float* __restrict__ h,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       AT* __restrict__ y, int I, int P) {
  const int lane = threadIdx.x & 31;
  const int T = P + 1;
  const long long row = static_cast<long long>(blockIdx.x) * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(I) * T) return;
  const int img = static_cast<int>(row / T);
  const int t = static_cast<int>(row - static_cast<long long>(img) * T);
  Row12 x;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = col_of(lane, i);
    const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * C + c));
    float4 a;
    if (t == 0) a = __ldg(reinterpret_cast<const float4*>(cls + c));
    else a = Pack4<TT>::load(tok + (static_cast<size_t>(img) * P + (t - 1)) * C + c);
    x.v[i] = make_float4(a.x + pe.x, a.y + pe.y, a.z + pe.z, a.w + pe.w);
  }
  const size_t base = static_cast<size_t>(row) * C;
#pragma unroll
  for (int i = 0; i < 3; ++i) *reinterpret_cast<float4*>(h + base + col_of(lane, i)) = x.v[i];
  const Row12 yv = ln_normalize(x, gamma, beta, eps, lane);
#pragma unroll
  for (int i = 0; i < 3; ++i) Pack4<AT>::store(y + base + col_of(lane, i), yv.v[i]);
}

// ---------------------------------------------------------------------------------------------
// last DINOv2 step + feature split + multi-view PE:
//   f = LN_final(h + delta);  drop CLS;  images [0, n_query) -> query tokens (fp32 residual + AT copy),
//   images [n_query, n_img) -> reference memory; every view gets the same resampled PE table added
//   (model/positional_encoding.py:72-74).  Images are ordered "all queries, then (b, ref)" so both
//   destinations are plain row-major (B*P, C) and (B*N*P, C) buffers.
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void __launch_bounds__(256)
final_ln_pe_kernel(const float* __restrict__ h, const AT* __restrict__ delta, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, const float* __restrict__ pe,
                   float* __restrict__ xq32, AT* __restrict__ xq, AT* __restrict__ mem, int n_img, int n_query,
                   int P) {
  const int lane = threadIdx.x & 31;
  const int T = P + 1;
  const long long row = static_cast<long long>(blockIdx.x) * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(n_img) * T) return;
  const int img = static_cast<int>(row / T);
  const int t = static_cast<int>(row - static_cast<long long>(img) * T);
  if (t == 0) return;  // CLS is dropped (task/core.py:142)
  const size_t base = static_cast<size_t>(row) * C;
  Row12 x;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float4 a = *reinterpret_cast<const float4*>(h + base + col_of(lane, i));
    if (delta) {
      const float4 d = Pack4<AT>::load(delta + base + col_of(lane, i));
      a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
    }
    x.v[i] = a;
  }
  Row12 f = ln_normalize(x, gamma, beta, eps, lane);
  const int p = t - 1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = col_of(lane, i);
    const float4 e = __ldg(reinterpret_cast<const float4*>(pe + static_cast<size_t>(p) * C + c));
    const float4 o = make_float4(f.v[i].x + e.x, f.v[i].y + e.y, f.v[i].z + e.z, f.v[i].w + e.w);
    if (img < n_query) {
      const size_t dst = (static_cast<size_t>(img) * P + p) * C + c;
      if (xq32) *reinterpret_cast<float4*>(xq32 + dst) = o;
      if (xq) Pack4<AT>::store(xq + dst, o);
    } else if (mem) {
      const size_t dst = (static_cast<size_t>(img - n_query) * P + p) * C + c;
      Pack4<AT>::store(mem + dst, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// im2col for the 14x14/stride-14 patch conv: out[(img*P + p), k], k = c*196 + ky*14 + kx (Conv2d weight order),
// zero padding for k in [588, Kpad)
// ---------------------------------------------------------------------------------------------
// One block = one patch row of one image: the 14 image rows of a channel are staged in shared memory with
// coalesced loads, then every patch's 196-element (c, ky, kx) segment is written contiguously (one warp per
// patch segment), so both sides of the copy run at full sector efficiency.
template <typename AT>
__global__ void __launch_bounds__(256) im2col14_kernel(const float* __restrict__ img, AT* __restrict__ out, int I,
                                                       int H, int W, int ph, int pw, int Kpad) {
  extern __shared__ float im2col_rows[];  // [14][W]
  const int im = blockIdx.x / ph, r = blockIdx.x - im * ph;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int used = 14 * pw;  // columns of the image row that belong to whole patches
  AT* orow0 = out + (static_cast<size_t>(im) * ph * pw + static_cast<size_t>(r) * pw) * Kpad;
  for (int c = 0; c < 3; ++c) {
    const float* src = img + ((static_cast<size_t>(im) * 3 + c) * H + 14 * r) * W;
    __syncthreads();  // previous channel's readers are done
    for (int e = threadIdx.x; e < 14 * used; e += 256) {
      const int ky = e / used, x = e - ky * used;
      im2col_rows[ky * W + x] = __ldg(src + static_cast<size_t>(ky) * W + x);
    }
    __syncthreads();
    for (int pc = warp; pc < pw; pc += 8) {
      AT* dst = orow0 + static_cast<size_t>(pc) * Kpad + c * 196;
      const float* s0 = im2col_rows + 14 * pc;
      if constexpr (sizeof(AT) == 4) {
        for (int k4 = lane; k4 < 49; k4 += 32) {  // 196 = 49 float4 (segment start is 16-byte aligned)
          float v[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int k = 4 * k4 + t;
            const int ky = k / 14;
            v[t] = s0[ky * W + (k - ky * 14)];
          }
          reinterpret_cast<float4*>(dst)[k4] = make_float4(v[0], v[1], v[2], v[3]);
        }
      } else {
        for (int k = lane; k < 196; k += 32) {
          const int ky = k / 14;
          dst[k] = __float2bfloat16_rn(s0[ky * W + (k - ky * 14)]);
        }
      }
      if (c == 2) {  // zero padding of the K tail
        for (int k = 588 + lane; k < Kpad; k += 32) {
          if constexpr (sizeof(AT) == 2) orow0[static_cast<size_t>(pc) * Kpad + k] = __float2bfloat16_rn(0.f);
          else orow0[static_cast<size_t>(pc) * Kpad + k] = 0.f;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// table resamplers (fp32, run once per (H, W) and cached by the host)
// ---------------------------------------------------------------------------------------------
// bilinear, align_corners=True: src = dst * (in-1)/(out-1)   (model/positional_encoding.py:61-69)
__global__ void bilinear_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int ih, int iw, int oh,
                                   int ow, int ch) {
  const long long total = static_cast<long long>(oh) * ow * ch;
  const float sy = oh > 1 ? static_cast<float>(ih - 1) / static_cast<float>(oh - 1) : 0.f;
  const float sx = ow > 1 ? static_cast<float>(iw - 1) / static_cast<float>(ow - 1) : 0.f;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % ch);
    const int ox = static_cast<int>((idx / ch) % ow);
    const int oy = static_cast<int>(idx / (static_cast<long long>(ch) * ow));
    const float fy = sy * oy, fx = sx * ox;
    const int y0 = min(static_cast<int>(fy), ih - 1), x0 = min(static_cast<int>(fx), iw - 1);
    const int y1 = min(y0 + 1, ih - 1), x1 = min(x0 + 1, iw - 1);
    const float ty = fy - y0, tx = fx - x0;
    const float v00 = in[(static_cast<size_t>(y0) * iw + x0) * ch + c], v01 = in[(static_cast<size_t>(y0) * iw + x1) * ch + c];
    const float v10 = in[(static_cast<size_t>(y1) * iw + x0) * ch + c], v11 = in[(static_cast<size_t>(y1) * iw + x1) * ch + c];
    const float top = v00 * (1.f - tx) + v01 * tx;
    const float bot = v10 * (1.f - tx) + v11 * tx;
    out[idx] = top * (1.f - ty) + bot * ty;
  }
}

__device__ __forceinline__ void cubic_w(float t, float (&w)[4]) {
  const float A = -0.75f;
  auto c1 = [&](float x) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; };
  auto c2 = [&](float x) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; };
  w[0] = c2(t + 1.f);
  w[1] = c1(t);
  w[2] = c1(1.f - t);
  w[3] = c2(2.f - t);
}
// bicubic (A=-0.75), align_corners=False, border-clamped taps ($SP/.../modeling_dinov2.py:86-91)
// sy / sx: source step per output pixel.  F.interpolate(size=...) uses in/out; F.interpolate(scale_factor=s) uses 1/s
// (ATen area_pixel_compute_scale), which is what transformers 4.33.3 -- the reference's pinned version -- asks for.
__global__ void bicubic_kernel(const float* __restrict__ in, float* __restrict__ out, int ih, int iw, int oh, int ow,
                               int ch, float sy, float sx) {
  const long long total = static_cast<long long>(oh) * ow * ch;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % ch);
    const int ox = static_cast<int>((idx / ch) % ow);
    const int oy = static_cast<int>(idx / (static_cast<long long>(ch) * ow));
    const float fy = (oy + 0.5f) * sy - 0.5f, fx = (ox + 0.5f) * sx - 0.5f;
    const float fly = floorf(fy), flx = floorf(fx);
    float wy[4], wx[4];
    cubic_w(fy - fly, wy);
    cubic_w(fx - flx, wx);
    const int iy = static_cast<int>(fly), ix = static_cast<int>(flx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), ih - 1);
      float racc = 0.f;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int xx = min(max(ix - 1 + bb, 0), iw - 1);
        racc += in[(static_cast<size_t>(yy) * iw + xx) * ch + c] * wx[bb];
      }
      acc += racc * wy[a];
    }
    out[idx] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// split-KV merge:  LSE = log sum_r exp(LSE_r);  O = sum_r exp(LSE_r - LSE) O_r      
//   o_parts  [R][rows][heads*d] fp32,  lse_parts [R][B][heads][Lq] fp32 (rows = B*Lq)
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void __launch_bounds__(256) lse_merge_kernel(const float* __restrict__ o_parts,
                                                        const float* __restrict__ lse_parts, AT* __restrict__ out,
                                                        float* __restrict__ lse_out, int R, int B, int Lq, int heads,
                                                        int d, long long part_o, long long part_l) {
  const long long total = static_cast<long long>(B) * Lq * heads * d;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(idx % d);
    const int hh = static_cast<int>((idx / d) % heads);
    const long long rowg = idx / (static_cast<long long>(d) * heads);
    const int b = static_cast<int>(rowg / Lq);
    const int r = static_cast<int>(rowg - static_cast<long long>(b) * Lq);
    const long long li = (static_cast<long long>(b) * heads + hh) * Lq + r;
    float mx = -INFINITY;
    for (int s = 0; s < R; ++s) mx = fmaxf(mx, lse_parts[s * part_l + li]);
    float den = 0.f, acc = 0.f;
    for (int s = 0; s < R; ++s) {
      const float w = expf(lse_parts[s * part_l + li] - mx);
      den += w;
      acc += w * o_parts[s * part_o + idx];
    }
    const float v = acc / den;  // a part with no keys carries lse = -inf: weight 0 (its O must be finite)
    if constexpr (sizeof(AT) == 2) out[idx] = __float2bfloat16_rn(v);
    else out[idx] = v;
    if (lse_out != nullptr && e == 0) lse_out[li] = mx + logf(den);
  }
}

// The same merge fused with its collective: parts[s] is the packed (O_s | LSE_s) buffer of rank s, read IN PLACE
// through NVLink peer pointers (symmetric memory), so the multi-GPU split-KV path needs no all-gather and no gathered
// buffer: one cross-rank barrier, then every rank pulls the R partials while it merges.
template <typename AT>
__global__ void __launch_bounds__(256) lse_merge_peers_kernel(const float* const* __restrict__ parts, long long base,
                                                              long long lse_off, AT* __restrict__ out,
                                                              float* __restrict__ lse_out, int R, int B, int Lq,
                                                              int heads, int d) {
  const long long total = static_cast<long long>(B) * Lq * heads * d;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(idx % d);
    const int hh = static_cast<int>((idx / d) % heads);
    const long long rowg = idx / (static_cast<long long>(d) * heads);
    const int b = static_cast<int>(rowg / Lq);
    const int r = static_cast<int>(rowg - static_cast<long long>(b) * Lq);
    const long long li = (static_cast<long long>(b) * heads + hh) * Lq + r;
    float mx = -INFINITY;
    for (int s = 0; s < R; ++s) mx = fmaxf(mx, parts[s][base + lse_off + li]);
    float den = 0.f, acc = 0.f;
    for (int s = 0; s < R; ++s) {
      const float* ps = parts[s] + base;
      const float w = expf(ps[lse_off + li] - mx);
      den += w;
      acc += w * ps[idx];
    }
    const float v = acc / den;
    if constexpr (sizeof(AT) == 2) out[idx] = __float2bfloat16_rn(v);
    else out[idx] = v;
    if (lse_out != nullptr && e == 0) lse_out[li] = mx + logf(den);
  }
}

// Vectorised variants (head_dim % 4 == 0, 16-byte aligned parts): one thread merges FOUR consecutive channels of a
// (row, head), so the R softmax weights are computed once per 16 bytes instead of once per element and every part is
// read with 16-byte loads -- over NVLink peer pointers a 4-byte load costs the same ~2 us round trip as a 16-byte one.
template <typename AT, bool PEERS>
__global__ void __launch_bounds__(256) lse_merge_vec4_kernel(const float* __restrict__ o_parts,
                                                             const float* __restrict__ lse_parts,
                                                             const float* const* __restrict__ parts, long long base,
                                                             long long lse_off, AT* __restrict__ out,
                                                             float* __restrict__ lse_out, int R, int B, int Lq, int heads,
                                                             int d, long long part_o, long long part_l) {
  const int d4 = d >> 2;
  const long long total4 = static_cast<long long>(B) * Lq * heads * d4;
  for (long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < total4;
       i4 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e4 = static_cast<int>(i4 % d4);
    const int hh = static_cast<int>((i4 / d4) % heads);
    const long long rowg = i4 / (static_cast<long long>(d4) * heads);
    const int b = static_cast<int>(rowg / Lq);
    const int r = static_cast<int>(rowg - static_cast<long long>(b) * Lq);
    const long long li = (static_cast<long long>(b) * heads + hh) * Lq + r;
    const long long idx = i4 * 4;
    float lse_s[16];
    float mx = -INFINITY;
#pragma unroll 4
    for (int s = 0; s < R; ++s) {
      lse_s[s & 15] = PEERS ? parts[s][base + lse_off + li] : lse_parts[s * part_l + li];
      mx = fmaxf(mx, lse_s[s & 15]);
    }
    float den = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < R; ++s) {
      const float ls = R <= 16 ? lse_s[s & 15] : (PEERS ? parts[s][base + lse_off + li] : lse_parts[s * part_l + li]);
      const float w = expf(ls - mx);
      const float4 o4 = *reinterpret_cast<const float4*>(PEERS ? parts[s] + base + idx : o_parts + s * part_o + idx);
      den += w;
      acc.x += w * o4.x;
      acc.y += w * o4.y;
      acc.z += w * o4.z;
      acc.w += w * o4.w;
    }
    const float inv = 1.0f / den;  // a part with no keys carries lse = -inf: weight 0 (its O must be finite)
    if constexpr (sizeof(AT) == 2) {
      uint2 w2;
      w2.x = pack_bf16x2(acc.x * inv, acc.y * inv);
      w2.y = pack_bf16x2(acc.z * inv, acc.w * inv);
      *reinterpret_cast<uint2*>(out + idx) = w2;
    } else {
      *reinterpret_cast<float4*>(out + idx) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
    if (lse_out != nullptr && e4 == 0) lse_out[li] = mx + logf(den);
  }
}

// ---------------------------------------------------------------------------------------------
// attention probabilities of ONE head (debug/visualisation path, need_attn_weights=True;
// model/customised_transformer/transformer.py:175-178, model/cross_reference.py:91-93):
//   probs[b, i, j] = exp(q_i . k_j * scale - lse[b, head, i])
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void __launch_bounds__(256)
attn_probs_kernel(const AT* __restrict__ q, const AT* __restrict__ k, const float* __restrict__ lse,
                  float* __restrict__ probs, int B, int heads, int head, int Lq, int Lk, int d, int head_slot,
                  long long q_row_stride, long long q_batch_stride, long long kv_row_stride,
                  long long kv_batch_stride, float scale) {
  const long long total = static_cast<long long>(B) * Lq * Lk;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(idx % Lk);
    const int i = static_cast<int>((idx / Lk) % Lq);
    const int b = static_cast<int>(idx / (static_cast<long long>(Lk) * Lq));
    const AT* qp = q + b * q_batch_stride + i * q_row_stride + head * head_slot;
    const AT* kp = k + b * kv_batch_stride + j * kv_row_stride + head * head_slot;
    float acc = 0.f;
    for (int e = 0; e < d; ++e) acc += static_cast<float>(qp[e]) * static_cast<float>(kp[e]);
    probs[idx] = expf(acc * scale - lse[(static_cast<long long>(b) * heads + head) * Lq + i]);
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers (dtype-dispatched)
// ---------------------------------------------------------------------------------------------
static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

#define XS_DISPATCH_AT(dtype, ...)                         \
  if ((dtype) == XS_BF16) {                                \
    using AT = __nv_bfloat16;                              \
    __VA_ARGS__;                                           \
  } else if ((dtype) == XS_F32) {                          \
    using AT = float;                                      \
    __VA_ARGS__;                                           \
  } else {                                                 \
    set_last_error("unknown dtype %d", (int)(dtype));      \
    return -1;                                             \
  }

int rows_add_ln(const float* res_in, const void* delta, float* res_out, const float* gamma, const float* beta,
                float eps, void* y, float* y32, int rows, int dtype, cudaStream_t stream) {
  XS_CHECK_ARG(rows > 0, "layernorm: rows=%d", rows);
  XS_CHECK_ARG(res_in || delta, "layernorm: need res_in or delta");
  const int grid = (rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
  XS_DISPATCH_AT(dtype, (add_ln_kernel<AT><<<grid, 256, 0, stream>>>(res_in, static_cast<const AT*>(delta), res_out,
                                                                     gamma, beta, eps, static_cast<AT*>(y), y32, rows)));
  XS_LAUNCH_CHECK();
  return 0;
}

int rows_stats(const float* h, void* hb, float* stats, float eps, int rows, cudaStream_t stream) {
  XS_CHECK_ARG(rows > 0, "row_stats: rows=%d", rows);
  const int grid = (rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
  row_stats_kernel<<<grid, 256, 0, stream>>>(h, static_cast<__nv_bfloat16*>(hb), reinterpret_cast<float2*>(stats), eps, rows);
  XS_LAUNCH_CHECK();
  return 0;
}

int rows_embed_ln(const void* tok, int tok_dtype, const float* cls, const float* pos, float* h, const float* gamma,
                  const float* beta, float eps, void* y, int I, int P, int dtype, cudaStream_t stream) {
  XS_CHECK_ARG(I > 0 && P > 0, "embed_ln: I=%d P=%d", I, P);
  const long long rows = static_cast<long long>(I) * (P + 1);
  const int grid = static_cast<int>((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
  if (tok_dtype == XS_F32) {
    XS_DISPATCH_AT(dtype, (embed_ln_kernel<AT, float><<<grid, 256, 0, stream>>>(
                              static_cast<const float*>(tok), cls, pos, h, gamma, beta, eps, static_cast<AT*>(y), I, P)));
  } else if (tok_dtype == XS_BF16) {
    XS_DISPATCH_AT(dtype, (embed_ln_kernel<AT, __nv_bfloat16><<<grid, 256, 0, stream>>>(
                              static_cast<const __nv_bfloat16*>(tok), cls, pos, h, gamma, beta, eps,
                              static_cast<AT*>(y), I, P)));
  } else {
    set_last_error("embed_ln: unknown tok_dtype %d", tok_dtype);
    return -1;
  }
  XS_LAUNCH_CHECK();
  return 0;
}

int rows_final_ln_pe(const float* h, const void* delta, const float* gamma, const float* beta, float eps,
                     const float* pe, float* xq32, void* xq, void* mem, int n_img, int n_query, int P, int dtype,
                     cudaStream_t stream) {
  XS_CHECK_ARG(n_img > 0 && P > 0 && n_query >= 0 && n_query <= n_img, "final_ln_pe: bad dims");
  const long long rows = static_cast<long long>(n_img) * (P + 1);
  const int grid = static_cast<int>((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
  XS_DISPATCH_AT(dtype, (final_ln_pe_kernel<AT><<<grid, 256, 0, stream>>>(
                            h, static_cast<const AT*>(delta), gamma, beta, eps, pe, xq32, static_cast<AT*>(xq),
                            static_cast<AT*>(mem), n_img, n_query, P)));
  XS_LAUNCH_CHECK();
  return 0;
}

int rows_im2col14(const float* img, void* out, int I, int H, int W, int Kpad, int dtype, cudaStream_t stream) {
  const int ph = H / 14, pw = W / 14;
  XS_CHECK_ARG(I > 0 && ph > 0 && pw > 0 && Kpad >= 588, "im2col: bad dims I=%d H=%d W=%d Kpad=%d", I, H, W, Kpad);
  const size_t smem = static_cast<size_t>(14) * W * sizeof(float);
  XS_CHECK_ARG(smem <= 200 * 1024, "im2col: image width %d too large", W);
  if (dtype == XS_BF16) {
    XS_CUDA(cudaFuncSetAttribute(im2col14_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  } else {
    XS_CUDA(cudaFuncSetAttribute(im2col14_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  XS_DISPATCH_AT(dtype, (im2col14_kernel<AT><<<I * ph, 256, smem, stream>>>(img, static_cast<AT*>(out), I, H, W, ph,
                                                                          pw, Kpad)));
  XS_LAUNCH_CHECK();
  return 0;
}

int table_bilinear_ac(const float* in, float* out, int ih, int iw, int oh, int ow, int ch, cudaStream_t stream) {
  XS_CHECK_ARG(ih > 0 && iw > 0 && oh > 0 && ow > 0 && ch > 0, "pe_resample: bad dims");
  bilinear_ac_kernel<<<grid_for(static_cast<long long>(oh) * ow * ch), 256, 0, stream>>>(in, out, ih, iw, oh, ow, ch);
  XS_LAUNCH_CHECK();
  return 0;
}

int table_bicubic(const float* in, float* out, int ih, int iw, int oh, int ow, int ch, float step_h, float step_w,
                  cudaStream_t stream) {
  XS_CHECK_ARG(ih > 0 && iw > 0 && oh > 0 && ow > 0 && ch > 0, "pos_resample: bad dims");
  XS_CHECK_ARG(step_h >= 0.f && step_w >= 0.f, "pos_resample: negative source step");
  const float sy = step_h > 0.f ? step_h : static_cast<float>(ih) / static_cast<float>(oh);
  const float sx = step_w > 0.f ? step_w : static_cast<float>(iw) / static_cast<float>(ow);
  bicubic_kernel<<<grid_for(static_cast<long long>(oh) * ow * ch), 256, 0, stream>>>(in, out, ih, iw, oh, ow, ch, sy,
                                                                                     sx);
  XS_LAUNCH_CHECK();
  return 0;
}

int rows_lse_merge(const float* o_parts, const float* lse_parts, void* out, float* lse_out, int R, int B, int Lq,
                   int heads, int d, long long o_part_stride, long long lse_part_stride, int dtype,
                   cudaStream_t stream) {
  XS_CHECK_ARG(R > 0 && B > 0 && Lq > 0 && heads > 0 && d > 0, "lse_merge: bad dims");
  const long long total = static_cast<long long>(B) * Lq * heads * d;
  const long long part_l = static_cast<long long>(B) * heads * Lq;
  if (o_part_stride == 0) o_part_stride = total;
  if (lse_part_stride == 0) lse_part_stride = part_l;
  XS_CHECK_ARG(o_part_stride >= total && lse_part_stride >= part_l, "lse_merge: part strides overlap");
  if ((d & 3) == 0 && (o_part_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(o_parts) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    XS_DISPATCH_AT(dtype, (lse_merge_vec4_kernel<AT, false><<<grid_for(total / 4), 256, 0, stream>>>(
                              o_parts, lse_parts, nullptr, 0, 0, static_cast<AT*>(out), lse_out, R, B, Lq, heads, d,
                              o_part_stride, lse_part_stride)));
    XS_LAUNCH_CHECK();
    return 0;
  }
  XS_DISPATCH_AT(dtype, (lse_merge_kernel<AT><<<grid_for(total), 256, 0, stream>>>(
                            o_parts, lse_parts, static_cast<AT*>(out), lse_out, R, B, Lq, heads, d, o_part_stride,
                            lse_part_stride)));
  XS_LAUNCH_CHECK();
  return 0;
}

int rows_lse_merge_peers(const void* const* parts, long long base, long long lse_off, void* out, float* lse_out, int R,
                         int B, int Lq, int heads, int d, int dtype, cudaStream_t stream) {
  XS_CHECK_ARG(parts != nullptr && R > 0 && B > 0 && Lq > 0 && heads > 0 && d > 0, "lse_merge_peers: bad dims");
  const long long total = static_cast<long long>(B) * Lq * heads * d;
  XS_CHECK_ARG(base >= 0 && lse_off >= total, "lse_merge_peers: LSE offset %lld overlaps O (%lld elements)", lse_off, total);
  if ((d & 3) == 0 && (base & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {  // peer buffers: cudaMalloc-aligned
    XS_DISPATCH_AT(dtype, (lse_merge_vec4_kernel<AT, true><<<grid_for(total / 4), 256, 0, stream>>>(
                              nullptr, nullptr, reinterpret_cast<const float* const*>(parts), base, lse_off,
                              static_cast<AT*>(out), lse_out, R, B, Lq, heads, d, 0, 0)));
    XS_LAUNCH_CHECK();
    return 0;
  }
  XS_DISPATCH_AT(dtype, (lse_merge_peers_kernel<AT><<<grid_for(total), 256, 0, stream>>>(
                            reinterpret_cast<const float* const*>(parts), base, lse_off, static_cast<AT*>(out), lse_out,
                            R, B, Lq, heads, d)));
  XS_LAUNCH_CHECK();
  return 0;
}

int rows_attn_probs(const void* q, const void* k, const float* lse, float* probs, int B, int heads, int head, int Lq,
                    int Lk, int d, int head_slot, long long q_row_stride, long long q_batch_stride,
                    long long kv_row_stride, long long kv_batch_stride, float scale, int dtype, cudaStream_t stream) {
  XS_CHECK_ARG(B > 0 && Lq > 0 && Lk > 0 && head >= 0 && head < heads, "attn_probs: bad dims / head id %d", head);
  const long long total = static_cast<long long>(B) * Lq * Lk;
  if (dtype == XS_F16) {  // fp16 q/k (the operands of the fp16 attention kernel)
    attn_probs_kernel<__half><<<grid_for(total), 256, 0, stream>>>(
        static_cast<const __half*>(q), static_cast<const __half*>(k), lse, probs, B, heads, head, Lq, Lk, d, head_slot,
        q_row_stride, q_batch_stride, kv_row_stride, kv_batch_stride, scale);
    XS_LAUNCH_CHECK();
    return 0;
  }
  XS_DISPATCH_AT(dtype, (attn_probs_kernel<AT><<<grid_for(total), 256, 0, stream>>>(
                            static_cast<const AT*>(q), static_cast<const AT*>(k), lse, probs, B, heads, head, Lq, Lk,
                            d, head_slot, q_row_stride, q_batch_stride, kv_row_stride, kv_batch_stride, scale)));
  XS_LAUNCH_CHECK();
  return 0;
}

}  // namespace xs
