// Flash attention (self and cross) on tcgen05 / TMEM, fed by TMA.   O = softmax(Q K^T * scale) V
//
// Replaces: Dinov2SelfAttention ($SP/transformers/models/dinov2/modeling_dinov2.py:203-234, 6 heads x 64),
// nn.MultiheadAttention's scaled-dot-product core in the decoder's self- and cross-attention blocks
// (model/customised_transformer/transformer.py:182-205 -> $SP/torch/nn/functional.py:6630-6692, 8 heads x 48).
//
// Persistent CTAs (two per SM) walk a static list of tiles = (batch, kv split, head, 128-query tile).  256 threads:
//   warps 0..3  softmax: thread == query row (tcgen05.ld 32x32b, TMEM lane quarter = warp id).  A 64-key block is two
//               32-column chunks, software-pipelined (chunk B's load flies during chunk A's exponentials, the next
//               block's chunk A during chunk B's).  P = exp2(S*scale - m) against a STALE reference max m: the block's
//               row sum is the overflow detector, the true max is taken (and l, O rescaled) only when it trips, so the
//               fast path has no max reduction.  P goes back to TMEM as bf16 over the upper half of its S buffer and is
//               handed to the MMA warp right after its store; final O / l and log-sum-exp in the epilogue.
//   warp 4      TMA producer: Q tile per tile, then K_j / V_j tiles [64 keys x 64] into a 5-stage ring
//   warp 5      TMEM allocator + MMA issuer:  S_b = Q K_j^T  (SS, K-major operands, 128B swizzle, N = 64)
//                                             O  += P_j V_j  (TS: P read from TMEM, V MN-major from smem)
//   warps 6..7  idle (they only lend their registers: setmaxnreg moves 4*200 + 4*56 = 8*128)
// S is TRIPLE-BUFFERED in TMEM (S_0..S_2): QK_{j+3} is issued right behind PV_j, so the softmax warps normally find
// S_{j+1} complete when they finish block j.  Barrier traffic is per warp: lane 0 polls / arrives, __syncwarp
// broadcasts.
// TMEM (256 columns, two CTAs per SM): S_b [64b, 64b+64), b = 0..2 (P_b aliases the last 32 columns of S_b),
// O [192, 192+DV).
// Head dim 48 (decoder) uses 64-wide padded head slots in global memory: QK^T issues 3 K-steps (48) and
// PV uses N=48, so no padded FLOPs are executed.
// Rows / keys beyond the sequence are zero-filled by TMA (3-D tensor maps) and masked to -inf here.
// With nsplit > 1 each CTA covers one kv range and emits a normalised partial O (fp32) + LSE that
// xs_lse_merge combines (single-GPU small-batch split and the multi-GPU split-KV path).
#include <stdlib.h>

#include "xs_common.cuh"

namespace xs {

constexpr int ATT_THREADS = 256;                        // warpgroup 0: softmax warps 0..3; warpgroup 1: TMA, MMA, 2 idle
constexpr int ATT_BKV = 64;                             // keys per block
constexpr int ATT_ST = 5;                               // K and V ring depth
// Which of the 16 column pairs of a 32-column chunk take the polynomial exponential (FMA pipe) instead of MUFU.EX2.
// Measured on the DINOv2 shape (I=48) with the eager hand-over: none 0.235 ms, 3/16 0.228 ms, 4/16 0.229 ms,
// 5/16 0.227 ms, 6/16 0.231 ms, 8/16 0.242 ms (decoder shape: 0.206 / 0.201 / 0.200 / 0.197 / 0.202 / 0.211 ms) --
// the softmax warps are paced by the hand-over chain and their own latency, not by the MUFU alone (62 % busy), so
// only a light offload pays.
#ifndef ATT_POLY_MASK
#define ATT_POLY_MASK 0x2492
#endif
// When the softmax warps hand P_g to the MMA warp: 1 = right after the P store of the block (wait::st + arrive),
// 0 = deferred behind the next block's first exponentials, 2 = at the top of the next block.  Measured (DINOv2 shape /
// decoder shape): 0: 0.241 / 0.211 ms, 2: 0.236 / 0.207 ms, 1: 0.228 / 0.201 ms -- the hand-over sits on the
// MMA -> softmax -> MMA chain that paces the kernel; the ~170 cycles of store latency it exposes are cheaper.
#ifndef ATT_EAGER_HANDOVER
#define ATT_EAGER_HANDOVER 1
#endif
constexpr int ATT_REGS_SOFTMAX = 200;                   // setmaxnreg budgets (multiples of 8): 4*200 + 4*56 = 8*128
constexpr int ATT_REGS_CTRL = 56;
constexpr int ATT_NS = 3;                               // S buffers in TMEM (QK runs ATT_NS blocks ahead of PV)
constexpr float ATT_SUM_LIMIT = 65536.0f;                // a block row-sum of P above this (vs the stale max) forces a rescale
constexpr float ATT_SUM_LIMIT_F16 = 2048.0f;             // fp16 variant: P and its packed-half partial sums stay below 65504
constexpr uint32_t ATT_Q_BYTES = 128 * 64 * 2;          // 16 KB: [128 rows][64 bf16], 128B swizzle
constexpr uint32_t ATT_KV_BYTES = ATT_BKV * 64 * 2;     // 8 KB:  [64 keys][64 bf16]
constexpr uint32_t ATT_SMEM_BYTES = ATT_Q_BYTES + 2 * ATT_ST * ATT_KV_BYTES + 256 + 1024;

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
  int nq_tiles, n_tiles;  // 128-query tiles per (batch, head, split); total tiles
  int dbg;                // PROF build only (XS_ATTN_DBG): timing experiments that break the result, see kernel
  unsigned long long* prof;  // XS_ATTN_PROF=1 (PROF instantiation only): per-phase clock totals, see flash_attn_bf16_tc
};

// development-only phase timer: lane 0 of every softmax warp / the MMA warp accumulates clock deltas
template <bool PROF>
struct PhaseClock {
  long long t;
  unsigned long long acc[8];
  __device__ __forceinline__ void start() {
    if constexpr (PROF) {
      t = clock64();
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0;
    }
  }
  __device__ __forceinline__ void lap(int i) {
    if constexpr (PROF) {
      const long long n = clock64();
      acc[i] += static_cast<unsigned long long>(n - t);
      t = n;
    }
  }
  __device__ __forceinline__ void flush(unsigned long long* out, int base, int lane) {
    if constexpr (PROF) {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(out + base + i, acc[i]);
      }
    }
  }
};

struct MaskNo { static constexpr bool value = false; };
struct MaskYes { static constexpr bool value = true; };

// columns >= valid of a 32-column chunk of logits -> -inf (ragged tail of the key sequence)
__device__ __forceinline__ void mask_tail(uint32_t (&v)[32], int valid) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i >= valid) v[i] = 0xff800000u;
}

// P = exp2(S * scale_log2 - m) for 32 logits of one row, packed to bf16x2; the two float2 accumulators collect
// the row sum.  Packed fp32x2 FMA / ADD (sm_100 FFMA2 / FADD2) halve the FMA-pipe issue slots next to the
// MUFU-bound exponentials: per pair 1 FFMA2 + 2 MUFU.EX2 + 1 FADD2 + 1 F2FP.
template <bool DBG>
__device__ __forceinline__ void exp_chunk(const uint32_t (&v)[32], float sl2, float neg_m, uint32_t (&pk)[16],
                                          float2& acc0, float2& acc1, int dbg) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = ffma2_bcast(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]), sl2, neg_m);
    float2 a;
    if (DBG && (dbg & 8)) {  // timing experiment: no MUFU
      a = x;
    } else if ((ATT_POLY_MASK >> i) & 1) {
      a = exp2_poly2(x);  // FMA-pipe exponential: takes this pair off the MUFU
    } else {
      a.x = fast_exp2(x.x);
      a.y = fast_exp2(x.y);
    }
    if (i & 1) acc1 = fadd2(acc1, a);
    else acc0 = fadd2(acc0, a);
    pk[i] = pack_bf16x2(a.x, a.y);
  }
}

// v[OFF .. OFF+N) * inv -> N consecutive outputs (16-byte vector stores)
template <int OFF, int N>
__device__ __forceinline__ void store_row_f32(float* dst, const uint32_t (&v)[32], float inv) {
#pragma unroll
  for (int i = 0; i < N / 4; ++i)
    reinterpret_cast<float4*>(dst)[i] =
        make_float4(__uint_as_float(v[OFF + 4 * i]) * inv, __uint_as_float(v[OFF + 4 * i + 1]) * inv,
                    __uint_as_float(v[OFF + 4 * i + 2]) * inv, __uint_as_float(v[OFF + 4 * i + 3]) * inv);
}
template <int OFF, int N>
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const uint32_t (&v)[32], float inv) {
#pragma unroll
  for (int i = 0; i < N / 8; ++i) {
    uint4 pk;
    pk.x = pack_bf16x2(__uint_as_float(v[OFF + 8 * i + 0]) * inv, __uint_as_float(v[OFF + 8 * i + 1]) * inv);
    pk.y = pack_bf16x2(__uint_as_float(v[OFF + 8 * i + 2]) * inv, __uint_as_float(v[OFF + 8 * i + 3]) * inv);
    pk.z = pack_bf16x2(__uint_as_float(v[OFF + 8 * i + 4]) * inv, __uint_as_float(v[OFF + 8 * i + 5]) * inv);
    pk.w = pack_bf16x2(__uint_as_float(v[OFF + 8 * i + 6]) * inv, __uint_as_float(v[OFF + 8 * i + 7]) * inv);
    reinterpret_cast<uint4*>(dst)[i] = pk;
  }
}

// Work decomposition of the persistent kernel: tile = (batch, kv split, head, 128-query tile), query tile fastest so
// the CTAs running at the same time share K/V of a few (batch, head) pairs in L2.
struct TileCoord {
  int q0, h, b, split, kv_begin, kv_end, nkv;
};
__device__ __forceinline__ TileCoord decode_tile(int tile, const AttnParams& p) {
  TileCoord t;
  const int qt = tile % p.nq_tiles;
  int r = tile / p.nq_tiles;
  t.h = r % p.heads;
  r /= p.heads;
  t.split = r % p.nsplit;
  t.b = r / p.nsplit;
  t.q0 = qt * 128;
  t.kv_begin = t.split * p.split_len;
  t.kv_end = min(p.Lk, t.kv_begin + p.split_len);
  t.nkv = (t.kv_end - t.kv_begin + ATT_BKV - 1) / ATT_BKV;
  return t;
}

// MODE 0: production; 1: phase clocks (XS_ATTN_PROF=1); 2: timing experiments that break the result (XS_ATTN_DBG=mask)
// F16: q/k/v are fp16, S = Q K^T is accumulated in FP16 (tcgen05 D format f16: one value per TMEM column, read two
// columns per register with tcgen05.ld.pack::16b at 1.75x the column rate of the fp32 load), the softmax runs on
// packed halves (HFMA2 / MUFU.EX2.F16x2 / HADD2: 3 instructions per two logits instead of 5 and no pack step), P is
// fp16.  The fp32 variant above is bound by the TMEM read of S (tcgen05.ld, ~60 % busy next to the 62 % busy MUFU);
// this one moves the bound to the MUFU alone.  Precision: logits rounded to fp16 is what the reference's own GPU path
// (16-mixed autocast, config/default_predict.yaml:25) does; the host folds scale*log2(e) into the query projection so
// that scale_log2 == 1 and the logits are small (tools/bf16_error_budget.py f16: error below the bf16-P variant).
template <int DQK_STEPS, int DV, int MODE, bool F16 = false>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, AttnParams p) {
  constexpr bool PROF = (MODE & 1) != 0;
  constexpr bool DBG = (MODE & 2) != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem;
  uint8_t* smK = smem + ATT_Q_BYTES;
  uint8_t* smV = smK + ATT_ST * ATT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smV + ATT_ST * ATT_KV_BYTES);
  const SmemBar bar0{smem_u32(bars)};
  const SmemBar q_full = bar0 + 0;             // Q tile landed
  const SmemBar q_empty = bar0 + 1;            // last QK^T of the tile complete: Q buffer free
  const SmemBar kv_full = bar0 + 2;            // [ATT_ST] K_g and V_g landed
  const SmemBar kv_empty = kv_full + ATT_ST;   // [ATT_ST] PV_g complete: slot free (also read by the O rescale)
  const SmemBar s_full = kv_empty + ATT_ST;    // [ATT_NS]
  const SmemBar p_full = s_full + ATT_NS;      // [ATT_NS]
  const SmemBar o_full = p_full + ATT_NS;      // all PV of the tile complete
  const SmemBar o_empty = o_full + 1;          // O read out by the softmax warps: next tile's PV_0 may overwrite
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * ATT_ST + 2 * ATT_NS + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < ATT_ST; ++s) {
      mbar_init(kv_full + (s), 1);
      mbar_init(kv_empty + (s), 1);
    }
    for (int s = 0; s < ATT_NS; ++s) {
      mbar_init(s_full + (s), 1);
      mbar_init(p_full + (s), 4);  // one arrival per softmax warp
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + ATT_NS * 64;
  // Register budget: the kernel is compiled for 128 registers/thread (2 CTAs x 256 threads); the control
  // warpgroup gives most of its share back and the softmax warpgroup (two 32-column chunks of logits, two packed
  // P chunks and a prefetch in flight per thread) takes it.

  // All roles walk the same static tile sequence; `g` counts K/V blocks over the CTA's lifetime (ring slot
  // g % ATT_ST, S buffer g % ATT_NS and the barrier phases follow from it), `it` counts tiles.
  if (warp >= 4) reg_dealloc<ATT_REGS_CTRL>();
  if (warp == 4) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    uint32_t g = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = decode_tile(tile, p);
      const int b_kv = p.kv_shared ? 0 : t.b;
      mbar_wait(q_empty, (it & 1) ^ 1);  // previous tile's QK^T are done with the Q buffer
      if (elect_one_sync()) {
        mbar_expect_tx(q_full, ATT_Q_BYTES);
        tma_load_3d(smQ, &tmQ, q_full, t.h * 64, t.q0, t.b);
      }
      __syncwarp();
      for (int j = 0; j < t.nkv; ++j, ++g) {
        const uint32_t s = g % ATT_ST;
        const int kv0 = t.kv_begin + j * ATT_BKV;
        mbar_wait(kv_empty + (s), ((g / ATT_ST) & 1) ^ 1);
        if (elect_one_sync()) {
          if (DBG && (p.dbg & 64) && g >= ATT_ST) {  // timing experiment: no K/V traffic after the first ring fill
            mbar_arrive(kv_full + (s));
          } else {
            mbar_expect_tx(kv_full + (s), 2 * ATT_KV_BYTES);
            tma_load_3d(smK + s * ATT_KV_BYTES, &tmK, kv_full + (s), t.h * 64, kv0, b_kv);
            tma_load_3d(smV + s * ATT_KV_BYTES, &tmV, kv_full + (s), t.h * 64, kv0, b_kv);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (converged warp, uniform operands, one elected lane issues) =========
    constexpr uint32_t idesc_qk = F16 ? umma_idesc_f16(128, ATT_BKV, 0, 0, 0) : umma_idesc_bf16(128, ATT_BKV, 0, 0);
    // B = V is MN-major ([kv][d], d contiguous)
    constexpr uint32_t idesc_pv = F16 ? umma_idesc_f16(128, DV, 0, 1, 1) : umma_idesc_bf16(128, DV, 0, 1);
    const uint32_t tb = warp_uniform(tmem_base);
    const uint32_t q_lo = umma_desc_lo(smem_u32(smQ), 16);
    const uint32_t k_lo0 = umma_desc_lo(smem_u32(smK), 16);
    const uint32_t v_lo0 = umma_desc_lo(smem_u32(smV), 1024);
    auto issue_qk = [&](uint32_t gg, bool last_of_tile) {
      const uint32_t s = gg % ATT_ST;
      mbar_wait(kv_full + (s), (gg / ATT_ST) & 1);  // K_gg (and V_gg) have landed
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t k_lo = k_lo0 + s * (ATT_KV_BYTES >> 4);
        const uint32_t d_s = tb + (gg % ATT_NS) * 64;
        if (!(DBG && (p.dbg & (256 | 2048)))) {  // (dbg 256: no MMAs at all, 2048: no QK^T MMAs)
#pragma unroll
          for (int k = 0; k < DQK_STEPS; ++k) umma_ss_lh<false>(d_s, q_lo + 2 * k, k_lo + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        }
        tc_commit(s_full + (gg % ATT_NS));
        if (last_of_tile) tc_commit(q_empty);  // Q buffer may be refilled once these MMAs have read it
      }
      __syncwarp();
    };
    PhaseClock<PROF> pc;
    pc.start();
    uint32_t g0 = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = decode_tile(tile, p);
      const int nkv = t.nkv;
      mbar_wait(q_full, it & 1);
      // the S buffers of the first blocks are free: the previous tile's PVs were issued before (the tensor pipe
      // executes in order) and the softmax warps had read those S blocks before they arrived on p_full
      for (int jj = 0; jj < ATT_NS && jj < nkv; ++jj) issue_qk(g0 + jj, jj == nkv - 1);
      pc.lap(0);  // prologue: Q + first S blocks issued
      for (int j = 0; j < nkv; ++j) {
        const uint32_t g = g0 + j;
        const uint32_t s = g % ATT_ST;
        const uint32_t sb = g % ATT_NS;
        // softmax has turned S_sb into P_g (and rescaled O if the row max moved)
        mbar_wait(p_full + (sb), (g / ATT_NS) & 1);
        if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);  // previous tile's O has been read out
        pc.lap(1);  // waiting for P_g
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t v_lo = v_lo0 + s * (ATT_KV_BYTES >> 4);
          const uint32_t a_p = tb + sb * 64 + 32;  // P_g lives in the upper half of S_sb
          if (DBG && (p.dbg & 1)) {  // timing experiment: A operand from smem (the Q tile) instead of P in TMEM
#pragma unroll
            for (int k = 0; k < ATT_BKV / 16; ++k)
              umma_ss_lh<false>(tb + ATT_NS * 64, q_lo + 2 * k, v_lo + k * 128, idesc_pv, (j | k) != 0 ? 1u : 0u);
          } else if (!(DBG && (p.dbg & (256 | 1024)))) {  // (dbg 1024: no PV MMAs)
#pragma unroll
            for (int k = 0; k < ATT_BKV / 16; ++k) {
              // A: 16 bf16 of P per row = 8 TMEM columns per K-step; B: 16 kv rows x 128 B = 2048 B per K-step
              umma_ts_lh(tb + ATT_NS * 64, a_p + k * 8, v_lo + k * 128, idesc_pv, (j | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(kv_empty + (s));  // K_g / V_g slot free; also the "PV_g complete" signal for the O rescale
          if (j == nkv - 1) tc_commit(o_full);
        }
        __syncwarp();
        pc.lap(2);  // PV issue
        if (j + ATT_NS < nkv) issue_qk(g + ATT_NS, j + ATT_NS == nkv - 1);  // overwrites S_sb behind PV_g
        pc.lap(3);  // K wait + QK issue
      }
      g0 += nkv;
    }
    pc.flush(p.prof, 8, lane);
  } else if (warp < 4) {
    reg_alloc<ATT_REGS_SOFTMAX>();
    if constexpr (F16) {
    // ===================== fp16 softmax / correction / epilogue (thread == query row) =====================
    const int q = warp;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_o = tmem_O + lane_off;
    const uint32_t sl2h = h2_bcast(p.scale_log2);
    const float sl2 = h2_lo(sl2h);  // the scale the exponentials actually use (1.0 exactly with folded weights)
    PhaseClock<PROF> pc;
    pc.start();
    uint32_t g0 = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      int nkv, tail_valid;
      {
        const TileCoord t = decode_tile(tile, p);
        nkv = t.nkv;
        tail_valid = t.kv_end - t.kv_begin - (t.nkv - 1) * ATT_BKV;
      }
      float m = -INFINITY;  // reference max (log2 domain), always exactly representable in fp16; may be stale
      float l = 0.f;
      int pending_sb = -1;
      auto flush_pending = [&]() {
        if (pending_sb >= 0) {
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full + pending_sb);
          pending_sb = -1;
        }
      };
      // Software pipeline as in the fp32 variant, on packed registers: a block is two 32-key chunks of 16 registers;
      // chunk B's load is in flight during chunk A's exponentials, the next block's chunk A during chunk B's.
      uint32_t va[16], vb[16];
      if (lane == 0) mbar_wait(s_full + (g0 % ATT_NS), (g0 / ATT_NS) & 1);
      __syncwarp();
      tc_fence_after();
      tmem_ld16_pack16(tmem_base + lane_off + (g0 % ATT_NS) * 64, va);
      pc.lap(0);

      auto block = [&](const int j, auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        const uint32_t g = g0 + j;
        const uint32_t sb = g % ATT_NS;
        const uint32_t t_s = tmem_base + lane_off + sb * 64;
        uint32_t pka[16], pkb[16];
        uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;  // packed-half partial row sums
        const uint32_t negm = h2_bcast(-m);
        // keys >= tail_valid of the last block are past the sequence end: -inf (fp16 0xFC00)
        auto mask16 = [&](uint32_t (&v)[16], int valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (2 * i >= valid) v[i] = 0xFC00FC00u;
            else if (2 * i + 1 >= valid) v[i] = (v[i] & 0x0000FFFFu) | 0xFC000000u;
          }
        };
        auto exp16 = [&](const uint32_t (&v)[16], uint32_t nm, uint32_t (&pk)[16]) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            pk[i] = h2_ex2(h2_fma(v[i], sl2h, nm)); a0 = h2_add(a0, pk[i]);
            pk[i + 1] = h2_ex2(h2_fma(v[i + 1], sl2h, nm)); a1 = h2_add(a1, pk[i + 1]);
            pk[i + 2] = h2_ex2(h2_fma(v[i + 2], sl2h, nm)); a2 = h2_add(a2, pk[i + 2]);
            pk[i + 3] = h2_ex2(h2_fma(v[i + 3], sl2h, nm)); a3 = h2_add(a3, pk[i + 3]);
          }
        };

        tmem_ld_wait16(va);
        tmem_ld16_pack16(t_s + 32, vb);  // in flight during chunk A
        pc.lap(1);
        if constexpr (MASK) mask16(va, tail_valid);
        if (j > 0) exp16(va, negm, pka);
        flush_pending();  // previous block's P
        pc.lap(2);
        tmem_ld_wait16(vb);
        pc.lap(1);
        // P_g (64 halves = 32 columns) overwrites the upper half of S_sb, all of which is in registers by now
        if (j > 0) tmem_st16(t_s + 32, pka);
        pc.lap(3);
        if (j + 1 < nkv) {  // prefetch chunk A of the next block
          const uint32_t sn = (g + 1) % ATT_NS;
          if (lane == 0) mbar_wait(s_full + (sn), ((g + 1) / ATT_NS) & 1);
          __syncwarp();
          pc.lap(0);  // pure wait for S_{g+1}
          tc_fence_after();
          tmem_ld16_pack16(tmem_base + lane_off + sn * 64, va);
        }
        pc.lap(6);  // fence + load issue
        if constexpr (MASK) mask16(vb, tail_valid - 32);
        if (j > 0) exp16(vb, negm, pkb);
        uint32_t at = h2_add(h2_add(a0, a1), h2_add(a2, a3));
        float bsum = h2_lo(at) + h2_hi(at);
        // Stale reference max, as in the fp32 variant: the block's row sum is the overflow detector (a P beyond
        // the fp16 range makes it inf).  The limit keeps every partial sum inside fp16.
        const bool need = (j == 0) || !(bsum <= ATT_SUM_LIMIT_F16);
        if (__any_sync(0xffffffffu, need)) {
          pc.lap(2);
          if (j + 1 < nkv) tmem_ld_wait16(va);  // the prefetch must land before va is reused
          tmem_ld16_pack16(t_s, va);            // S chunk A again (columns 0..31 have not been overwritten)
          tmem_ld_wait16(va);
          if constexpr (MASK) mask16(va, tail_valid);
          uint32_t mx2 = h2_max(va[0], vb[0]);
#pragma unroll
          for (int i = 1; i < 16; ++i) mx2 = h2_max(mx2, h2_max(va[i], vb[i]));
          const float mx = fmaxf(h2_lo(mx2), h2_hi(mx2)) * sl2;
          float m_new = need ? fmaxf(mx, m) : m;
          m_new = h2_lo(h2_bcast(m_new));  // keep m representable in fp16 (exact when scale_log2 == 1)
          const float alpha = fast_exp2(m - m_new);  // 1 when unchanged, 0 when m was -inf
          l *= alpha;
          if (j > 0) {
            if (lane == 0) mbar_wait(kv_empty + ((g - 1) % ATT_ST), ((g - 1) / ATT_ST) & 1);  // PV_{g-1} complete
            __syncwarp();
            tc_fence_after();
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
          const uint32_t negm2 = h2_bcast(-m);
          a0 = a1 = a2 = a3 = 0u;
          tc_wait_st();  // the early store of the stale P_A must not pass the corrected one
          exp16(va, negm2, pka);
          tmem_st16(t_s + 32, pka);
          exp16(vb, negm2, pkb);
          at = h2_add(h2_add(a0, a1), h2_add(a2, a3));
          bsum = h2_lo(at) + h2_hi(at);
          if (j + 1 < nkv) tmem_ld16_pack16(tmem_base + lane_off + ((g + 1) % ATT_NS) * 64, va);  // redo the prefetch
          pc.lap(7);
        }
        l += bsum;
        pc.lap(2);
        tmem_st16(t_s + 48, pkb);
#if ATT_EAGER_HANDOVER == 1
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + sb);
#else
        pending_sb = static_cast<int>(sb);
#endif
        pc.lap(3);
      };
      for (int j = 0; j + 1 < nkv; ++j) block(j, MaskNo{});
      block(nkv - 1, MaskYes{});
      flush_pending();
      g0 += nkv;

      // ---- epilogue: O / l and log-sum-exp (same as the fp32-logit variant) ----
      if (lane == 0) mbar_wait(o_full, it & 1);
      __syncwarp();
      pc.lap(4);
      tc_fence_after();
      uint32_t oa[32], ob[32];
      tmem_ld32(t_o, oa);
      if constexpr (DV == 64) tmem_ld32(t_o + 32, ob);
      else tmem_ld16_lo(t_o + 32, ob);
      tmem_ld_wait32(oa);
      tmem_ld_wait32(ob);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const float inv = 1.0f / l;
      const TileCoord t = decode_tile(tile, p);
      const int row_g = t.q0 + row;
      const long long o_off = static_cast<long long>(t.split) * p.o_split_stride +
                              static_cast<long long>(t.b) * p.o_batch_stride +
                              static_cast<long long>(row_g) * p.o_row_stride + static_cast<long long>(t.h) * DV;
      if (row_g < p.Lq) {
        if (p.o_is_f32) {
          float* dst = reinterpret_cast<float*>(p.o) + o_off;
          store_row_f32<0, 32>(dst, oa, inv);
          store_row_f32<0, DV - 32>(dst + 32, ob, inv);
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.o) + o_off;
          store_row_bf16<0, 32>(dst, oa, inv);
          store_row_bf16<0, DV - 32>(dst + 32, ob, inv);
        }
        if (p.lse != nullptr) {
          p.lse[static_cast<long long>(t.split) * p.lse_split_stride +
                (static_cast<long long>(t.b) * p.heads + t.h) * p.Lq + row_g] = (m + log2f(l)) * 0.6931471805599453f;
        }
      }
      pc.lap(5);
    }
    pc.flush(p.prof, 0, lane);
    pc.flush(p.prof, 16 + 8 * warp, lane);
    } else {
    // ===================== softmax / correction / epilogue (thread == query row) =====================
    const int q = warp;  // TMEM lane quarter accessible to this warp (warp id % 4)
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_o = tmem_O + lane_off;
    const float sl2 = p.scale_log2;
    PhaseClock<PROF> pc;
    pc.start();
    uint32_t g0 = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      int nkv, tail_valid;  // only these two stay live across the block loop (the tile is decoded again for the stores)
      {
        const TileCoord t = decode_tile(tile, p);
        nkv = t.nkv;
        tail_valid = t.kv_end - t.kv_begin - (t.nkv - 1) * ATT_BKV;  // valid columns of the last block
      }
      float m = -INFINITY;  // reference max of the row (log2 domain); may be stale (see below)
      float l = 0.f;        // running sum of exp2(s - m)
      // P_g's hand-over (wait for the tcgen05.st, fence, arrive on p_full) is deferred into the next block, behind
      // the first chunk's exponentials, so the store latency is never waited for
      int pending_sb = -1;
      auto flush_pending = [&]() {
        if (pending_sb >= 0) {
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full + pending_sb);
          pending_sb = -1;
        }
      };

      // Software pipeline: a block is processed as two 32-column chunks (va, vb).  While chunk A is in the
      // exponentials the tcgen05.ld of chunk B is in flight; while chunk B is in the exponentials, chunk A of
      // the NEXT block is in flight (S is triple-buffered, so S_{g+1} is normally complete long before).
      uint32_t va[32], vb[32];
      if (lane == 0) mbar_wait(s_full + (g0 % ATT_NS), (g0 / ATT_NS) & 1);
      __syncwarp();
      tc_fence_after();
      tmem_ld32(tmem_base + lane_off + (g0 % ATT_NS) * 64, va);
      pc.lap(0);

      auto block = [&](const int j, auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        const uint32_t g = g0 + j;
        const uint32_t sb = g % ATT_NS;
        const uint32_t t_s = tmem_base + lane_off + sb * 64;
        const int valid = tail_valid;  // MASK variant = last block: columns >= valid are past the sequence end
        uint32_t pka[16], pkb[16];
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        const float neg_m = -m;

        tmem_ld_wait32(va);
        if (!(DBG && (p.dbg & 4))) tmem_ld32(t_s + 32, vb);  // in flight during chunk A   (dbg 4: no S loads)
#if ATT_EAGER_HANDOVER == 2
        flush_pending();  // previous block's P: its store had the load wait above to land
#endif
        pc.lap(1);
        if constexpr (MASK) mask_tail(va, valid);
        if (j > 0) exp_chunk<DBG>(va, sl2, neg_m, pka, acc0, acc1, p.dbg);
#if ATT_EAGER_HANDOVER == 0
        flush_pending();  // previous block's P
#endif
        pc.lap(2);
        tmem_ld_wait32(vb);
        pc.lap(6);  // residual wait for chunk B
        // P_g overwrites the UPPER half of S_sb (columns 32..63, all in registers now), so chunk A's P goes out
        // while chunk B is in the exponentials and the lower half stays intact for the slow path's reload
        if (j > 0 && !(DBG && (p.dbg & 2))) tmem_st16(t_s + 32, pka);  // (dbg 2: no P stores)
        if (j + 1 < nkv) {  // prefetch chunk A of the next block (va is dead until then)
          const uint32_t sn = (g + 1) % ATT_NS;
          if (lane == 0) mbar_wait(s_full + (sn), ((g + 1) / ATT_NS) & 1);
          __syncwarp();
          tc_fence_after();
          if (!(DBG && (p.dbg & 4))) tmem_ld32(tmem_base + lane_off + sn * 64, va);
        }
        pc.lap(0);
        if constexpr (MASK) mask_tail(vb, valid - 32);
        if (j > 0) exp_chunk<DBG>(vb, sl2, neg_m, pkb, acc0, acc1, p.dbg);
        float bsum = (acc0.x + acc0.y) + (acc1.x + acc1.y);
        // The exponentials above used the STALE reference max m: any m gives the same softmax as long as
        // 2^(s-m) stays in range (bf16 P and the fp32 sums keep their relative precision at any magnitude).
        // The row max is therefore not tracked on the fast path at all; the block's row sum is the overflow
        // detector (every P <= bsum): only when it exceeds ATT_SUM_LIMIT (or is inf/NaN) is the true block max
        // taken, (l, O) rescaled and the block recomputed.  The first block of a tile always takes this path.
        const bool need = (j == 0) || (!(bsum <= ATT_SUM_LIMIT) && !(DBG && p.dbg));
        if (__any_sync(0xffffffffu, need)) {
          pc.lap(2);
          if (j + 1 < nkv) tmem_ld_wait32(va);  // the prefetch must land before va is reused
          tmem_ld32(t_s, va);                   // S chunk A again (its columns have not been overwritten)
          tmem_ld_wait32(va);
          if constexpr (MASK) mask_tail(va, valid);
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(va[i]));
            mx1 = fmaxf(mx1, __uint_as_float(va[i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(vb[i]));
            mx3 = fmaxf(mx3, __uint_as_float(vb[i + 1]));
          }
          const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
          const float m_new = need ? fmaxf(mx, m) : m;
          const float alpha = fast_exp2(m - m_new);  // 1 when unchanged, 0 when m was -inf
          l *= alpha;
          if (j > 0) {
            // O must hold PV of all earlier blocks of this tile: wait for the previous block's PV through its
            // K/V slot's kv_empty phase (the slot is refilled only ATT_ST blocks later: the parity cannot alias)
            if (lane == 0) mbar_wait(kv_empty + ((g - 1) % ATT_ST), ((g - 1) / ATT_ST) & 1);
            __syncwarp();
            tc_fence_after();
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
          acc0 = make_float2(0.f, 0.f);
          acc1 = make_float2(0.f, 0.f);
          tc_wait_st();  // the early store of the stale P_A must not pass the corrected one
          exp_chunk<false>(va, sl2, -m, pka, acc0, acc1, 0);
          tmem_st16(t_s + 32, pka);
          exp_chunk<false>(vb, sl2, -m, pkb, acc0, acc1, 0);
          bsum = (acc0.x + acc0.y) + (acc1.x + acc1.y);
          if (j + 1 < nkv) tmem_ld32(tmem_base + lane_off + ((g + 1) % ATT_NS) * 64, va);  // redo the prefetch
          pc.lap(7);  // slow path total
        }
        l += bsum;
        pc.lap(2);  // exp2 / pack
        if (!(DBG && (p.dbg & 2))) tmem_st16(t_s + 48, pkb);
#if ATT_EAGER_HANDOVER == 1
        tc_wait_st();  // hand P_g over right away: the MMA -> softmax -> MMA chain latency, not the store, is what binds
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + sb);
#else
        pending_sb = static_cast<int>(sb);
#endif
        pc.lap(3);  // tcgen05.st of P issued
      };
      for (int j = 0; j + 1 < nkv; ++j) block(j, MaskNo{});
      block(nkv - 1, MaskYes{});
      flush_pending();
      g0 += nkv;

      // ---- epilogue: read O out of TMEM (frees it for the next tile's PV_0), then O / l and log-sum-exp ----
      if (lane == 0) mbar_wait(o_full, it & 1);
      __syncwarp();
      pc.lap(4);  // waiting for the last PV
      tc_fence_after();
      tmem_ld32(t_o, va);
      if constexpr (DV == 64) tmem_ld32(t_o + 32, vb);
      else tmem_ld16_lo(t_o + 32, vb);  // DV == 48
      tmem_ld_wait32(va);
      tmem_ld_wait32(vb);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const float inv = 1.0f / l;
      const TileCoord t = decode_tile(tile, p);
      const int row_g = t.q0 + row;
      const bool row_ok = row_g < p.Lq;
      const long long o_off = static_cast<long long>(t.split) * p.o_split_stride +
                              static_cast<long long>(t.b) * p.o_batch_stride +
                              static_cast<long long>(row_g) * p.o_row_stride + static_cast<long long>(t.h) * DV;
      if (row_ok) {
        if (p.o_is_f32) {
          float* dst = reinterpret_cast<float*>(p.o) + o_off;
          store_row_f32<0, 32>(dst, va, inv);
          store_row_f32<0, DV - 32>(dst + 32, vb, inv);
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.o) + o_off;
          store_row_bf16<0, 32>(dst, va, inv);
          store_row_bf16<0, DV - 32>(dst + 32, vb, inv);
        }
        if (p.lse != nullptr) {
          // natural-log LSE of the scaled logits: ln sum_j exp(s_j * scale)
          p.lse[static_cast<long long>(t.split) * p.lse_split_stride +
                (static_cast<long long>(t.b) * p.heads + t.h) * p.Lq + row_g] = (m + log2f(l)) * 0.6931471805599453f;
        }
      }
      pc.lap(5);  // epilogue stores
    }
    pc.flush(p.prof, 0, lane);
    pc.flush(p.prof, 16 + 8 * warp, lane);  // per lane-quarter copy
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Development aid (XS_ATTN_PROF=1): run the instrumented instantiation synchronously and print the average
// clocks each softmax warp / MMA warp spent per phase (per CTA lifetime) to stderr.
static int launch_attn_prof(int head_dim, dim3 grid, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                            const CUtensorMap& tmV, AttnParams p, cudaStream_t stream) {
  static unsigned long long* buf = nullptr;
  if (buf == nullptr) XS_CUDA(cudaMalloc(&buf, 48 * sizeof(unsigned long long)));
  XS_CUDA(cudaMemsetAsync(buf, 0, 48 * sizeof(unsigned long long), stream));
  p.prof = buf;
  if (head_dim == 64 && p.dbg) {
    auto kern = attn_tc_kernel<4, 64, 3>;
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
    kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  } else if (head_dim == 64) {
    auto kern = attn_tc_kernel<4, 64, 1>;
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
    kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  } else {
    auto kern = attn_tc_kernel<3, 48, 1>;
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
    kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  }
  XS_LAUNCH_CHECK();
  XS_CUDA(cudaStreamSynchronize(stream));
  unsigned long long h[48];
  XS_CUDA(cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost));
  const double ctas = double(p.n_tiles);  // per 128-query tile
  const int nkv = (p.split_len < p.Lk ? p.split_len : p.Lk + ATT_BKV - 1) / ATT_BKV;
  fprintf(stderr, "attn prof (clk per tile, ~%d kv blocks): softmax warp: wait_S %.0f  ldA %.0f  exp %.0f  st_P %.0f  wait_O %.0f  "
                  "epi %.0f  ldB %.0f  slow %.0f | mma warp: prologue %.0f  wait_P %.0f  issue_PV %.0f  waitK+issue_QK %.0f\n",
          nkv, h[0] / ctas / 4, h[1] / ctas / 4, h[2] / ctas / 4, h[3] / ctas / 4, h[4] / ctas / 4, h[5] / ctas / 4,
          h[6] / ctas / 4, h[7] / ctas / 4, h[8] / ctas, h[9] / ctas, h[10] / ctas, h[11] / ctas);
  for (int w = 0; w < 4; ++w) {
    const unsigned long long* g = h + 16 + 8 * w;
    fprintf(stderr, "  softmax warp %d: wait_S %.0f  ldA %.0f  exp %.0f  st_P %.0f  wait_O %.0f  epi %.0f  ldB %.0f  slow %.0f\n", w,
            g[0] / ctas, g[1] / ctas, g[2] / ctas, g[3] / ctas, g[4] / ctas, g[5] / ctas, g[6] / ctas, g[7] / ctas);
  }
  return 0;
}

// q/k/v: bf16, head h occupies 64 consecutive columns starting at h*64 of its row (d=48: 48 used + 16 pad)
// strides in elements.  o: [nsplit][B][Lq][heads*head_dim] (bf16, or fp32 when o_is_f32)
// operands_f16: q/k/v are fp16 and the fp16-logit kernel runs (see attn_tc_kernel<..., F16>)
int flash_attn_bf16_tc(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq,
                       int Lk, int head_dim, long long q_row_stride, long long q_batch_stride,
                       long long kv_row_stride, long long kv_batch_stride, int kv_shared, int nsplit, int o_is_f32,
                       float scale, int operands_f16, cudaStream_t stream) {
  XS_CHECK_ARG(head_dim == 64 || head_dim == 48, "flash_attn: head_dim %d not supported (64 or 48)", head_dim);
  XS_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk > 0 && nsplit > 0, "flash_attn: empty problem");
  XS_CHECK_ARG((q_row_stride % 8) == 0 && (kv_row_stride % 8) == 0 && (q_batch_stride % 8) == 0 &&
                   (kv_batch_stride % 8) == 0,
               "flash_attn: strides must be multiples of 8 elements");
  XS_CHECK_ARG(nsplit == 1 || (o_is_f32 && lse != nullptr), "flash_attn: split-KV needs fp32 partial O and LSE");
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[3] = {64, 128, 1};
  const uint32_t box_kv[3] = {64, ATT_BKV, 1};
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
    int rc = make_tmap(&tmK, k, 2, 3, dims, strides, box_kv, SWZ_128B);
    if (rc) return rc;
    rc = make_tmap(&tmV, v, 2, 3, dims, strides, box_kv, SWZ_128B);
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
  p.nq_tiles = (Lq + 127) / 128;
  p.n_tiles = p.nq_tiles * heads * B * nsplit;
  const int max_ctas = 2 * num_sms();  // two co-resident CTAs per SM, each walking its share of the tiles
  dim3 grid(p.n_tiles < max_ctas ? p.n_tiles : max_ctas);
  p.prof = nullptr;
  p.dbg = 0;
  static int prof = -1;
  if (prof < 0) {
    const char* e = getenv("XS_ATTN_PROF");
    prof = e ? atoi(e) : 0;
  }
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("XS_ATTN_DBG");
    dbg = e ? atoi(e) : 0;
  }
  if (operands_f16) {
    if (prof) {  // phase clocks of the fp16 variant (XS_ATTN_PROF=1)
      static unsigned long long* buf = nullptr;
      if (buf == nullptr) XS_CUDA(cudaMalloc(&buf, 48 * sizeof(unsigned long long)));
      XS_CUDA(cudaMemsetAsync(buf, 0, 48 * sizeof(unsigned long long), stream));
      p.prof = buf;
      if (head_dim == 64) {
        auto kern = attn_tc_kernel<4, 64, 1, true>;
        XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
        kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
      } else {
        auto kern = attn_tc_kernel<3, 48, 1, true>;
        XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
        kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
      }
      XS_LAUNCH_CHECK();
      XS_CUDA(cudaStreamSynchronize(stream));
      unsigned long long h[48];
      XS_CUDA(cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost));
      const double ctas = double(p.n_tiles);
      fprintf(stderr, "attn f16 prof (clk per tile): softmax warp: wait_S %.0f  ld_wait %.0f  exp %.0f  st_P %.0f  wait_O %.0f  "
                      "epi %.0f  fence+ld_issue %.0f  slow %.0f | mma warp: prologue %.0f  wait_P %.0f  issue_PV %.0f  waitK+issue_QK %.0f\n",
              h[0] / ctas / 4, h[1] / ctas / 4, h[2] / ctas / 4, h[3] / ctas / 4, h[4] / ctas / 4, h[5] / ctas / 4,
              h[6] / ctas / 4, h[7] / ctas / 4, h[8] / ctas, h[9] / ctas, h[10] / ctas, h[11] / ctas);
      return 0;
    }
    if (head_dim == 64) {
      auto kern = attn_tc_kernel<4, 64, 0, true>;
      XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
      kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
    } else {
      auto kern = attn_tc_kernel<3, 48, 0, true>;
      XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
      kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
    }
    XS_LAUNCH_CHECK();
    return 0;
  }
  if (head_dim == 64) p.dbg = dbg;
  if (dbg && !prof && head_dim == 64) {  // development: timing experiments on the d=64 shape
    auto kern = attn_tc_kernel<4, 64, 2>;
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
    kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
    XS_LAUNCH_CHECK();
    return 0;
  }
  if (prof) return launch_attn_prof(head_dim, grid, tmQ, tmK, tmV, p, stream);
  if (head_dim == 64) {
    auto kern = attn_tc_kernel<4, 64, 0>;
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
    kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  } else {
    auto kern = attn_tc_kernel<3, 48, 0>;
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
    kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  }
  XS_LAUNCH_CHECK();
  return 0;
}

}  // namespace xs
