// fp32 parity mode: the same path with fp32 storage and SIMT FMA arithmetic (no tensor cores), so the
// score map matches the fp32 reference to <= 1e-4 (BASELINE.json north_star).  This is a correctness mode:
// simple tiled kernels, not tuned.  It shares the row kernels of xs_rows.cu (AT = float).
#include "xs_common.cuh"

namespace xs {

// ---------------------------------------------------------------------------------------------
// C[M,N] = act(A[M,K] @ W[N,K]^T + bias)          128x128x16 tiles, 256 threads, 8x8 per thread
// EPI 0: row-major store; EPI 1: score activation + jigsaw scatter (N = 196)
// ---------------------------------------------------------------------------------------------
struct F32Jigsaw {
  float* score;
  int P, pw, Wout, HWout, use_tanh;
  float power;
};

template <int ACT, int EPI>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                const float* __restrict__ bias, float* __restrict__ Cout, int ldc, int M, int N, int K,
                F32Jigsaw jp) {
  constexpr int BM = 128, BN = 128, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // loader mapping: 128 rows x 16 k = 2048 floats = 512 float4; each thread loads 2 float4 per operand
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int f = tid + it * 256;  // 0..511
      const int r = f >> 2, kq = (f & 3) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk = k0 + kq;
      if (m0 + r < M) {
        const float* ap = A + static_cast<size_t>(m0 + r) * lda + kk;
        if (kk + 3 < K) a = *reinterpret_cast<const float4*>(ap);
        else {
          if (kk + 0 < K) a.x = ap[0];
          if (kk + 1 < K) a.y = ap[1];
          if (kk + 2 < K) a.z = ap[2];
        }
      }
      if (n0 + r < N) {
        const float* wp = W + static_cast<size_t>(n0 + r) * ldw + kk;
        if (kk + 3 < K) w = *reinterpret_cast<const float4*>(wp);
        else {
          if (kk + 0 < K) w.x = wp[0];
          if (kk + 1 < K) w.y = wp[1];
          if (kk + 2 < K) w.z = wp[2];
        }
      }
      As[kq + 0][r] = a.x; As[kq + 1][r] = a.y; As[kq + 2][r] = a.z; As[kq + 3][r] = a.w;
      Ws[kq + 0][r] = w.x; Ws[kq + 1][r] = w.y; Ws[kq + 2][r] = w.z; Ws[kq + 3][r] = w.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], w[8];
      // rows ty*4 + {0..3} and 64 + ty*4 + {0..3}; cols tx*4 + {0..3} and 64 + tx*4 + {0..3}
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j] + bias[n];
      if constexpr (EPI == 0) {
        if constexpr (ACT == ACT_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
        else if constexpr (ACT == ACT_RELU) v = fmaxf(v, 0.f);
        else if constexpr (ACT == ACT_LEAKY) v = v >= 0.f ? v : 0.01f * v;
        Cout[static_cast<size_t>(m) * ldc + n] = v;
      } else {
        float s = jp.use_tanh ? tanhf(v) : 1.0f / (1.0f + expf(-v));
        if (jp.power != 1.0f) s = powf(s, jp.power);
        const int b = m / jp.P, p = m - b * jp.P;
        const int r = p / jp.pw, cc = p - r * jp.pw;
        const int ii = n / 14, jj = n - ii * 14;
        jp.score[static_cast<size_t>(b) * jp.HWout + static_cast<size_t>(14 * r + ii) * jp.Wout + 14 * cc + jj] = s;
      }
    }
  }
}

int gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias, float* out, int ldc, int M, int N,
             int K, int act, cudaStream_t stream) {
  XS_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_f32: empty problem");
  XS_CHECK_ARG((lda % 4) == 0 && (ldw % 4) == 0, "gemm_f32: lda/ldw must be multiples of 4");
  dim3 grid((N + 127) / 128, (M + 127) / 128);
  F32Jigsaw jp{};
  switch (act) {
    case ACT_NONE: gemm_f32_kernel<ACT_NONE, 0><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, out, ldc, M, N, K, jp); break;
    case ACT_GELU: gemm_f32_kernel<ACT_GELU, 0><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, out, ldc, M, N, K, jp); break;
    case ACT_RELU: gemm_f32_kernel<ACT_RELU, 0><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, out, ldc, M, N, K, jp); break;
    case ACT_LEAKY: gemm_f32_kernel<ACT_LEAKY, 0><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, out, ldc, M, N, K, jp); break;
    default: set_last_error("gemm_f32: unknown activation %d", act); return -1;
  }
  XS_LAUNCH_CHECK();
  return 0;
}

int head_jigsaw_f32(const float* A, int lda, const float* W, int ldw, const float* bias, float* score, int B, int ph,
                    int pw, int K, int use_tanh, float power, cudaStream_t stream) {
  XS_CHECK_ARG(B > 0 && ph > 0 && pw > 0, "head_jigsaw_f32: empty problem");
  F32Jigsaw jp;
  jp.score = score;
  jp.P = ph * pw;
  jp.pw = pw;
  jp.Wout = 14 * pw;
  jp.HWout = 14 * ph * 14 * pw;
  jp.use_tanh = use_tanh;
  jp.power = power;
  const int M = B * ph * pw, N = 196;
  dim3 grid((N + 127) / 128, (M + 127) / 128);
  gemm_f32_kernel<ACT_NONE, 1><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, nullptr, 0, M, N, K, jp);
  XS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// fp32 flash attention: CTA = 64 queries of one (batch, head, kv split); key tiles of 64; 256 threads as a
// 16x16 grid, each thread owns a 4x4 block of S and 4 rows x D/16 columns of O.
// q/k/v rows: head h at columns [h*head_slot, h*head_slot + D).  Output layout as the tensor-core kernel.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
attn_f32_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                float* __restrict__ o, float* __restrict__ lse, int heads, int Lq, int Lk, int head_slot,
                long long q_row_stride, long long q_batch_stride, long long kv_row_stride,
                long long kv_batch_stride, int kv_shared, int nsplit, int split_len, long long o_split_stride,
                long long lse_split_stride, float scale) {
  constexpr int BQ = 64, BKV = 64, DC = D / 16;
  extern __shared__ float sm[];
  float* Qs = sm;                   // [D][BQ+1]   (transposed: Qs[e][row])
  float* Ks = Qs + D * (BQ + 1);    // [D][BKV+1]  (transposed: Ks[e][col])
  float* Vs = Ks + D * (BKV + 1);   // [BKV][D]
  float* Ps = Vs + BKV * D;         // [BQ][BKV+1]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y;
  const int b = blockIdx.z / nsplit, split = blockIdx.z - b * nsplit;
  const int kv_begin = split * split_len, kv_end = min(Lk, kv_begin + split_len);
  const float* qb = q + b * q_batch_stride + static_cast<long long>(h) * head_slot;
  const float* kb = k + (kv_shared ? 0 : b) * kv_batch_stride + static_cast<long long>(h) * head_slot;
  const float* vb = v + (kv_shared ? 0 : b) * kv_batch_stride + static_cast<long long>(h) * head_slot;

  for (int f = tid; f < BQ * D; f += 256) {
    const int r = f / D, e = f - r * D;
    Qs[e * (BQ + 1) + r] = (q0 + r < Lq) ? qb[(q0 + r) * q_row_stride + e] * scale : 0.f;
  }
  float m_run[4], l_run[4], oacc[4][DC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < DC; ++c) oacc[i][c] = 0.f;
  }

  for (int kv0 = kv_begin; kv0 < kv_end; kv0 += BKV) {
    __syncthreads();  // previous tile fully consumed (also orders the Q fill on the first pass)
    for (int f = tid; f < BKV * D; f += 256) {
      const int r = f / D, e = f - r * D;
      const bool ok = kv0 + r < kv_end;
      Ks[e * (BKV + 1) + r] = ok ? kb[(kv0 + r) * kv_row_stride + e] : 0.f;
      Vs[r * D + e] = ok ? vb[(kv0 + r) * kv_row_stride + e] : 0.f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int e = 0; e < D; ++e) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[e * (BQ + 1) + ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ks[e * (BKV + 1) + tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], w[j], s[i][j]);
    }
    // online softmax per row; the 16 threads sharing a row (same ty) are 16 consecutive lanes
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kv0 + tx * 4 + j >= kv_end) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int o2 = 8; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
      const float m_new = fmaxf(m_run[i], mx);
      const float alpha = expf(m_run[i] - m_new);  // exp(-inf) = 0 on the first tile
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pj = expf(s[i][j] - m_new);
        sum += pj;
        Ps[(ty * 4 + i) * (BKV + 1) + tx * 4 + j] = pj;
      }
#pragma unroll
      for (int o2 = 8; o2 > 0; o2 >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o2);
      l_run[i] = l_run[i] * alpha + sum;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < DC; ++c) oacc[i][c] *= alpha;
    }
    __syncthreads();
    // O += P V : thread owns rows ty*4+i, columns tx + 16*c
#pragma unroll 4
    for (int kk = 0; kk < BKV; ++kk) {
      float pv[4], vv[DC];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = Ps[(ty * 4 + i) * (BKV + 1) + kk];
#pragma unroll
      for (int c = 0; c < DC; ++c) vv[c] = Vs[kk * D + tx + 16 * c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < DC; ++c) oacc[i][c] = fmaf(pv[i], vv[c], oacc[i][c]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    if (row >= Lq) continue;
    const float inv = 1.0f / l_run[i];
    float* op = o + split * o_split_stride + (static_cast<long long>(b) * Lq + row) * (heads * D) + h * D;
#pragma unroll
    for (int c = 0; c < DC; ++c) op[tx + 16 * c] = oacc[i][c] * inv;
    if (lse != nullptr && tx == 0)
      lse[split * lse_split_stride + (static_cast<long long>(b) * heads + h) * Lq + row] = m_run[i] + logf(l_run[i]);
  }
}

int flash_attn_f32(const float* q, const float* k, const float* v, float* o, float* lse, int B, int heads, int Lq,
                   int Lk, int head_dim, int head_slot, long long q_row_stride, long long q_batch_stride,
                   long long kv_row_stride, long long kv_batch_stride, int kv_shared, int nsplit, float scale,
                   cudaStream_t stream) {
  XS_CHECK_ARG(head_dim == 64 || head_dim == 48, "flash_attn_f32: head_dim %d not supported", head_dim);
  XS_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk > 0 && nsplit > 0, "flash_attn_f32: empty problem");
  XS_CHECK_ARG(nsplit == 1 || lse != nullptr, "flash_attn_f32: split-KV needs LSE");
  const int nblk = (Lk + 63) / 64;
  const int split_len = ((nblk + nsplit - 1) / nsplit) * 64;
  XS_CHECK_ARG((long long)(nsplit - 1) * split_len < Lk, "flash_attn_f32: nsplit=%d leaves an empty kv range", nsplit);
  const long long o_split_stride = (long long)B * Lq * heads * head_dim;
  const long long lse_split_stride = (long long)B * heads * Lq;
  dim3 grid((Lq + 63) / 64, heads, B * nsplit);
  if (head_dim == 64) {
    const int smem = (2 * 64 * 65 + 64 * 64 + 64 * 65) * 4;
    XS_CUDA(cudaFuncSetAttribute(attn_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attn_f32_kernel<64><<<grid, 256, smem, stream>>>(q, k, v, o, lse, heads, Lq, Lk, head_slot, q_row_stride,
                                                     q_batch_stride, kv_row_stride, kv_batch_stride, kv_shared,
                                                     nsplit, split_len, o_split_stride, lse_split_stride, scale);
  } else {
    const int smem = (2 * 48 * 65 + 64 * 48 + 64 * 65) * 4;
    XS_CUDA(cudaFuncSetAttribute(attn_f32_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attn_f32_kernel<48><<<grid, 256, smem, stream>>>(q, k, v, o, lse, heads, Lq, Lk, head_slot, q_row_stride,
                                                     q_batch_stride, kv_row_stride, kv_batch_stride, kv_shared,
                                                     nsplit, split_len, o_split_stride, lse_split_stride, scale);
  }
  XS_LAUNCH_CHECK();
  return 0;
}

}  // namespace xs
