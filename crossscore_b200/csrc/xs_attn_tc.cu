// Flash attention (self and cross) on tcgen05 / TMEM, fed by TMA.   O = softmax(Q K^T * scale) V
//
// Replaces: Dinov2SelfAttention ($SP/transformers/models/dinov2/modeling_dinov2.py:203-234, 6 heads x 64),
// nn.MultiheadAttention's scaled-dot-product core in the decoder's self- and cross-attention blocks
// (model/customised_transformer/transformer.py:182-205 -> $SP/torch/nn/functional.py:6630-6692, 8 heads x 48).
//
// One CTA = one (batch, head, 128-query tile, kv split); two CTAs are co-resident per SM so one CTA's
// softmax overlaps the other's tensor-core work.  320 threads:
//   warp 0      TMA producer: Q tile once, then K_j / V_j tiles [128 x 64] into a 2-stage ring
//   warp 1      TMEM allocator + MMA issuer:  S = Q K_j^T  (SS, K-major operands, 128B swizzle)
//                                             O_h += P_j[:, half h] V_j[half h]  (TS: P from TMEM, V MN-major smem)
//   warps 2..9  softmax, two threads per query row: the 128 key columns of a block are split into two halves
//               that run as independent online-softmax streams (own running max / sum and own O accumulator),
//               so no per-block exchange is needed; the halves are merged once at the end with the split-KV
//               identity.  This doubles the warps available to hide MUFU/TMEM latency (4 per scheduler with
//               two CTAs) -- attention at head dim 64/48 is exp-bound (MUFU 16/clk/SM), not MMA-bound.
//               Per block: tcgen05.ld (thread == row), max, lazy O rescale (only when the running max grows
//               by > 2^8), P = exp2(S*scale - m) written back as bf16 over the thread's own S columns.
// TMEM (256 columns): S [0,128) (P_0 aliases [0,32), P_1 aliases [64,96)), O_0 [128,128+DV), O_1 [192,192+DV).
// Head dim 48 (decoder) uses 64-wide padded head slots in global memory: QK^T issues 3 K-steps (48) and
// PV uses N=48, so no padded FLOPs are executed.
// Rows / keys beyond the sequence are zero-filled by TMA (3-D tensor maps) and masked to -inf here.
// With nsplit > 1 each CTA covers one kv range and emits a normalised partial O (fp32) + LSE that
// xs_lse_merge combines (single-GPU small-batch split and the multi-GPU split-KV path).
#include "xs_common.cuh"

namespace xs {

constexpr int ATT_THREADS = 320;
constexpr uint32_t ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: [128 rows][64 bf16], 128B swizzle
constexpr uint32_t ATT_XCHG_BYTES = 2 * 128 * 8;   // (m, l) of both halves for the final merge
constexpr uint32_t ATT_SMEM_BYTES = 5 * ATT_TILE_BYTES + ATT_XCHG_BYTES + 256 + 1024;

struct AttnParams {
  void* o;
  float* lse;
  int o_is_f32;
  int Lq, Lk, heads;
  int kv_shared;
  int nsplit, split_len;
  long long o_row_stride, o_batch_stride, o_split_stride;  // elements
  long long lse_split_stride;
  float scale_log2;
};

// ---- softmax building blocks: one thread owns 64 fp32 S columns of its row, at TMEM address t_s ------------
template <bool MASKED>
__device__ __forceinline__ float chunk_max16(const uint32_t (&v)[16], int col0, int valid) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    if constexpr (MASKED) {
      m0 = fmaxf(m0, (col0 + i + 0 < valid) ? __uint_as_float(v[i + 0]) : -INFINITY);
      m1 = fmaxf(m1, (col0 + i + 1 < valid) ? __uint_as_float(v[i + 1]) : -INFINITY);
    } else {
      m0 = fmaxf(m0, __uint_as_float(v[i + 0]));
      m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
    }
  }
  return fmaxf(m0, m1);
}

// raw (unscaled) max over the 64 columns; the TMEM load of chunk c+1 is in flight while chunk c is reduced
template <bool MASKED>
__device__ __forceinline__ float half_row_max(uint32_t t_s, int col_base, int valid) {
  uint32_t va[16], vb[16];
  tmem_ld16(t_s, va);
  tmem_ld_wait16(va);
  tmem_ld16(t_s + 16, vb);
  float mx = chunk_max16<MASKED>(va, col_base, valid);
  tmem_ld_wait16(vb);
  tmem_ld16(t_s + 32, va);
  mx = fmaxf(mx, chunk_max16<MASKED>(vb, col_base + 16, valid));
  tmem_ld_wait16(va);
  tmem_ld16(t_s + 48, vb);
  mx = fmaxf(mx, chunk_max16<MASKED>(va, col_base + 32, valid));
  tmem_ld_wait16(vb);
  return fmaxf(mx, chunk_max16<MASKED>(vb, col_base + 48, valid));
}

// P = exp2(S*sl2 - m) for 16 columns -> 8 packed bf16 pairs stored to TMEM at t_dst; returns the partial sum
template <bool MASKED>
__device__ __forceinline__ float chunk_exp_store16(const uint32_t (&v)[16], uint32_t t_dst, float sl2, float m,
                                                   int col0, int valid) {
  uint32_t pk[8];
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), sl2, -m));
    float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m));
    if constexpr (MASKED) {
      p0 = (col0 + 2 * i < valid) ? p0 : 0.f;
      p1 = (col0 + 2 * i + 1 < valid) ? p1 : 0.f;
    }
    l0 += p0;
    l1 += p1;
    pk[i] = pack_bf16x2(p0, p1);
  }
  tmem_st8(t_dst, pk);
  return l0 + l1;
}

// second pass; P (32 packed columns) overwrites the first half of the thread's own S columns, always behind
// the columns already consumed: P cols [8c, 8c+8) <= S cols < 16c+16
template <bool MASKED>
__device__ __forceinline__ float half_row_exp(uint32_t t_s, float sl2, float m, int col_base, int valid) {
  uint32_t va[16], vb[16];
  tmem_ld16(t_s, va);
  tmem_ld_wait16(va);
  tmem_ld16(t_s + 16, vb);
  float l = chunk_exp_store16<MASKED>(va, t_s, sl2, m, col_base, valid);
  tmem_ld_wait16(vb);
  tmem_ld16(t_s + 32, va);
  l += chunk_exp_store16<MASKED>(vb, t_s + 8, sl2, m, col_base + 16, valid);
  tmem_ld_wait16(va);
  tmem_ld16(t_s + 48, vb);
  l += chunk_exp_store16<MASKED>(va, t_s + 16, sl2, m, col_base + 32, valid);
  tmem_ld_wait16(vb);
  l += chunk_exp_store16<MASKED>(vb, t_s + 24, sl2, m, col_base + 48, valid);
  return l;
}

template <int DQK_STEPS, int DV>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem;
  uint8_t* smK = smem + ATT_TILE_BYTES;      // 2 stages
  uint8_t* smV = smem + 3 * ATT_TILE_BYTES;  // 2 stages
  float2* xchg = reinterpret_cast<float2*>(smem + 5 * ATT_TILE_BYTES);  // [2][128] (m, l)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * ATT_TILE_BYTES + ATT_XCHG_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;    // [2]: one per column half
  uint64_t* o_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z / p.nsplit;
  const int split = blockIdx.z - b * p.nsplit;
  const int kv_begin = split * p.split_len;
  const int kv_end = min(p.Lk, kv_begin + p.split_len);
  const int nkv = (kv_end - kv_begin + 127) / 128;
  const int b_kv = p.kv_shared ? 0 : b;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&p_full[s], 128);
    }
    mbar_init(s_full, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;  // fp32 S, 128 columns

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    mbar_expect_tx(q_full, ATT_TILE_BYTES);
    tma_load_3d(smQ, &tmQ, q_full, h * 64, q0, b);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
      const int kv0 = kv_begin + j * 128;
      mbar_expect_tx(&k_full[s], ATT_TILE_BYTES);
      tma_load_3d(smK + s * ATT_TILE_BYTES, &tmK, &k_full[s], h * 64, kv0, b_kv);
      mbar_expect_tx(&v_full[s], ATT_TILE_BYTES);
      tma_load_3d(smV + s * ATT_TILE_BYTES, &tmV, &v_full[s], h * 64, kv0, b_kv);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, DV, 0, 1);  // B = V is MN-major ([kv][d], d contiguous)
    const uint32_t q_addr = smem_u32(smQ);
    mbar_wait(q_full, 0);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&k_full[s], ph);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(smK + s * ATT_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < DQK_STEPS; ++k) {
        umma_ss(tmem_S, umma_desc_sw128(q_addr + k * 32, 16, 1024), umma_desc_sw128(k_addr + k * 32, 16, 1024),
                idesc_qk, k != 0 ? 1u : 0u);
      }
      tc_commit(s_full);
      mbar_wait(&v_full[s], ph);
      const uint32_t v_addr = smem_u32(smV + s * ATT_TILE_BYTES);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        // half hf of the softmax has turned its S columns into P (and rescaled O_hf if its max moved)
        mbar_wait(&p_full[hf], j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // A: 16 bf16 of P per row = 8 TMEM columns per K-step; B: 16 kv rows x 128 B = 2048 B per K-step
          umma_ts(tmem_base + 128 + hf * 64, tmem_S + hf * 64 + k * 8,
                  umma_desc_sw128(v_addr + (hf * 4 + k) * 2048, 1024, 1024), idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
      }
      tc_commit(&kv_empty[s]);  // K_j and V_j are free once QK_j / PV_j have completed
    }
    tc_commit(o_full);
  } else if (warp >= 2) {
    // ===================== softmax / correction / epilogue =====================
    const int q = warp & 3;           // TMEM lane quarter accessible to this warp
    const int hf = (warp - 2) >> 2;   // column half handled by this thread
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_s = tmem_S + lane_off + hf * 64;
    const uint32_t t_o = tmem_base + lane_off + 128 + hf * 64;
    const int col_base = hf * 64;
    const float sl2 = p.scale_log2;
    float m = -INFINITY;  // running (possibly stale) max of this half's stream, log2 domain
    float l = 0.f;        // running sum of exp2(s - m)

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int valid = kv_end - (kv_begin + j * 128);  // columns >= valid are past the sequence end
      const bool full = valid >= 128;
      if (!full && valid <= col_base) {
        // this half of the ragged last block is entirely past the end: contributes nothing (P = 0)
        uint32_t z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st8(t_s + c * 8, z);
        if (j == 0) {  // O_hf was never written: the PV below runs with accumulate = 0 over P = 0 -> zeros
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[hf]);
        continue;
      }

      // ---- pass 1: max of this half (masking only in the ragged last block) ----
      float mx = full ? half_row_max<false>(t_s, col_base, valid) : half_row_max<true>(t_s, col_base, valid);
      mx *= sl2;

      // ---- lazy correction: rescale (l, O_hf) only when the max grew by more than 2^8 ----
      const bool need = mx > m + 8.0f;  // always true on the first block (m = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mx : m;
        const float alpha = fast_exp2(m - m_new);  // 1 when unchanged, 0 when m was -inf
        l *= alpha;
        if (j > 0) {
#pragma unroll
          for (int c = 0; c < DV / 16; ++c) {
            uint32_t v[16];
            tmem_ld16(t_o + c * 16, v);
            tmem_ld_wait16(v);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(t_o + c * 16, v);
          }
        }
        m = m_new;
      }

      // ---- pass 2: P = exp2(S*scale - m) -> bf16 pairs over this thread's own S columns ----
      l += full ? half_row_exp<false>(t_s, sl2, m, col_base, valid) : half_row_exp<true>(t_s, sl2, m, col_base, valid);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[hf]);
    }

    // ---- merge the two halves (split-KV identity) and write O / l, log-sum-exp ----
    xchg[hf * 128 + row] = make_float2(m, l);
    named_bar_sync(1, 256);
    const float2 other = xchg[(hf ^ 1) * 128 + row];
    const float m_all = fmaxf(m, other.x);  // at least one half saw a valid column, so m_all is finite
    const float w_me = fast_exp2(m - m_all), w_ot = fast_exp2(other.x - m_all);
    const float l_all = l * w_me + other.y * w_ot;
    const float inv = 1.0f / l_all;
    const float w0 = (hf == 0 ? w_me : w_ot) * inv;  // weight of O_0
    const float w1 = (hf == 0 ? w_ot : w_me) * inv;  // weight of O_1
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int row_g = q0 + row;
    const bool row_ok = row_g < p.Lq;
    // this thread writes output columns [hf*DV/2, (hf+1)*DV/2)
    constexpr int HC = DV / 2;  // 32 or 24
    const uint32_t t_o0 = tmem_base + lane_off + 128 + hf * HC;
    const uint32_t t_o1 = tmem_base + lane_off + 192 + hf * HC;
    const long long o_off = static_cast<long long>(split) * p.o_split_stride +
                            static_cast<long long>(b) * p.o_batch_stride +
                            static_cast<long long>(row_g) * p.o_row_stride + static_cast<long long>(h) * DV + hf * HC;
#pragma unroll
    for (int c = 0; c < HC / 8; ++c) {
      uint32_t a8[8], b8[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(a8[0]), "=r"(a8[1]), "=r"(a8[2]), "=r"(a8[3]), "=r"(a8[4]), "=r"(a8[5]), "=r"(a8[6]),
                     "=r"(a8[7])
                   : "r"(t_o0 + c * 8)
                   : "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(b8[0]), "=r"(b8[1]), "=r"(b8[2]), "=r"(b8[3]), "=r"(b8[4]), "=r"(b8[5]), "=r"(b8[6]),
                     "=r"(b8[7])
                   : "r"(t_o1 + c * 8)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;"
                   : "+r"(a8[0]), "+r"(a8[1]), "+r"(a8[2]), "+r"(a8[3]), "+r"(a8[4]), "+r"(a8[5]), "+r"(a8[6]),
                     "+r"(a8[7]), "+r"(b8[0]), "+r"(b8[1]), "+r"(b8[2]), "+r"(b8[3]), "+r"(b8[4]), "+r"(b8[5]),
                     "+r"(b8[6]), "+r"(b8[7])
                   :
                   : "memory");
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = __uint_as_float(a8[i]) * w0 + __uint_as_float(b8[i]) * w1;
      if (row_ok) {
        if (p.o_is_f32) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.o) + o_off + c * 8);
          dst[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
          dst[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
        } else {
          uint4 pk;
          pk.x = pack_bf16x2(o8[0], o8[1]);
          pk.y = pack_bf16x2(o8[2], o8[3]);
          pk.z = pack_bf16x2(o8[4], o8[5]);
          pk.w = pack_bf16x2(o8[6], o8[7]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.o) + o_off + c * 8) = pk;
        }
      }
    }
    if (p.lse != nullptr && row_ok && hf == 0) {
      // natural-log LSE of the scaled logits: ln sum_j exp(s_j * scale)
      p.lse[static_cast<long long>(split) * p.lse_split_stride +
            (static_cast<long long>(b) * p.heads + h) * p.Lq + row_g] = (m_all + log2f(l_all)) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// q/k/v: bf16, head h occupies 64 consecutive columns starting at h*64 of its row (d=48: 48 used + 16 pad)
// strides in elements.  o: [nsplit][B][Lq][heads*head_dim] (bf16, or fp32 when o_is_f32)
int flash_attn_bf16_tc(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq,
                       int Lk, int head_dim, long long q_row_stride, long long q_batch_stride,
                       long long kv_row_stride, long long kv_batch_stride, int kv_shared, int nsplit, int o_is_f32,
                       float scale, cudaStream_t stream) {
  XS_CHECK_ARG(head_dim == 64 || head_dim == 48, "flash_attn: head_dim %d not supported (64 or 48)", head_dim);
  XS_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk > 0 && nsplit > 0, "flash_attn: empty problem");
  XS_CHECK_ARG((q_row_stride % 8) == 0 && (kv_row_stride % 8) == 0 && (q_batch_stride % 8) == 0 &&
                   (kv_batch_stride % 8) == 0,
               "flash_attn: strides must be multiples of 8 elements");
  XS_CHECK_ARG(nsplit == 1 || (o_is_f32 && lse != nullptr), "flash_attn: split-KV needs fp32 partial O and LSE");
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)Lq, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)q_row_stride * 2, (uint64_t)(B > 1 ? q_batch_stride : (long long)Lq * q_row_stride) * 2};
    int rc = make_tmap(&tmQ, q, 2, 3, dims, strides, box, SWZ_128B);
    if (rc) return rc;
  }
  {
    const int Bkv = kv_shared ? 1 : B;
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)Lk, (uint64_t)Bkv};
    uint64_t strides[2] = {(uint64_t)kv_row_stride * 2,
                           (uint64_t)(Bkv > 1 ? kv_batch_stride : (long long)Lk * kv_row_stride) * 2};
    int rc = make_tmap(&tmK, k, 2, 3, dims, strides, box, SWZ_128B);
    if (rc) return rc;
    rc = make_tmap(&tmV, v, 2, 3, dims, strides, box, SWZ_128B);
    if (rc) return rc;
  }
  AttnParams p;
  p.o = o;
  p.lse = lse;
  p.o_is_f32 = o_is_f32;
  p.Lq = Lq;
  p.Lk = Lk;
  p.heads = heads;
  p.kv_shared = kv_shared;
  p.nsplit = nsplit;
  const int nblk = (Lk + 127) / 128;
  p.split_len = ((nblk + nsplit - 1) / nsplit) * 128;
  XS_CHECK_ARG((long long)(nsplit - 1) * p.split_len < Lk, "flash_attn: nsplit=%d leaves an empty kv range (Lk=%d)",
               nsplit, Lk);
  p.o_row_stride = (long long)heads * head_dim;
  p.o_batch_stride = (long long)Lq * p.o_row_stride;
  p.o_split_stride = (long long)B * p.o_batch_stride;
  p.lse_split_stride = (long long)B * heads * Lq;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Lq + 127) / 128, heads, B * nsplit);
  if (head_dim == 64) {
    auto kern = attn_tc_kernel<4, 64>;
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
    kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  } else {
    auto kern = attn_tc_kernel<3, 48>;
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
    kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  }
  XS_LAUNCH_CHECK();
  return 0;
}

}  // namespace xs
