// Flash attention (self and cross) on tcgen05 / TMEM, fed by TMA.   O = softmax(Q K^T * scale) V
//
// Replaces: Dinov2SelfAttention ($SP/transformers/models/dinov2/modeling_dinov2.py:203-234, 6 heads x 64),
// nn.MultiheadAttention's scaled-dot-product core in the decoder's self- and cross-attention blocks
// (model/customised_transformer/transformer.py:182-205 -> $SP/torch/nn/functional.py:6630-6692, 8 heads x 48).
//
// Persistent CTAs (two per SM) walk a static list of tiles = (batch, kv split, head, 128-query tile).  384 threads:
//   warps 0..7  softmax.  Warp w owns 16 query rows: TMEM lane quarter w % 4, lower / upper half by w / 4, read with
//               tcgen05.ld.16x256b (thread t holds rows t/4 and t/4+8, two adjacent columns of every 8-column group,
//               so the bf16x2 packing of P needs no shuffles and goes back with tcgen05.st.16x128b).  Sixteen softmax
//               warps per SM (four per scheduler) hide each other's TMEM / mbarrier latencies; round 1 ran eight, one
//               thread per row, and was latency-bound at 32 % tensor-pipe activity.
//   warp 8      TMA producer: Q tile per tile, then K_j / V_j tiles [64 keys x 64] into a 5-stage ring
//   warp 9      TMEM allocator + PV issuer:  O += P_j V_j  (TS: P read from TMEM, V MN-major from smem)
//   warp 10     QK issuer:  S_b = Q K_j^T  (SS, K-major operands, 128B swizzle, N = 64), up to 3 blocks ahead
//   warp 11     idle (it only lends registers: setmaxnreg moves 8*96 + 4*40 <= 12*80)
// S is TRIPLE-BUFFERED in TMEM (S_0..S_2): QK_{j+3} is issued as soon as PV_j has completed (s_free).
// TMEM (256 columns per CTA): S_b [64b, 64b+64), b = 0..2 (P_b overwrites the upper 32 columns of S_b), O [192, 192+DV).
//
// Softmax arithmetic.  Pass 1 ("optimistic") uses NO running maximum: P = 2^(S * scale * log2 e) directly, row sums in
// fp32.  Softmax is invariant to the reference point as long as nothing overflows or flushes, and bf16 / fp32 carry
// 8 exponent bits, so this is exact for |logit * log2 e| < ~100 -- per element it costs [MUFU.EX2 | cubic on the FMA
// pipe] + 1/2 FADD2 + 1/2 F2FP and no max / subtract / rescale.  A tile whose final row sum leaves [2^-80, 2^100] (or is
// not finite) is marked in shared memory and REDONE in pass 2 by the same CTA with a conventional online softmax
// (per-block row max, O rescaled in TMEM when it moves), so any input finite in fp32 gets the right answer.
// A fraction of the exponentials (ATT_POLY_MASK) runs as a cubic polynomial on the FMA pipe: the MUFU (16 / clk / SM)
// is the binding unit at head dim 64.
// Head dim 48 (decoder) uses 64-wide padded head slots in global memory: QK^T issues 3 K-steps (48) and PV uses N=48,
// so no padded FLOPs are executed.  Rows / keys beyond the sequence are zero-filled by TMA and masked to -inf here.
// With nsplit > 1 each tile covers one kv range and emits a normalised partial O (fp32) + LSE that xs_lse_merge
// combines (single-GPU small-batch split and the multi-GPU split-KV path).
#include <atomic>

#include "xs_common.cuh"

namespace xs {

constexpr int ATT_THREADS = 384;
constexpr int ATT_BKV = 64;                             // keys per block
constexpr int ATT_ST = 5;                               // K and V ring depth
constexpr int ATT_NS = 3;                               // S buffers in TMEM (QK runs ATT_NS blocks ahead of PV)
// Which of the 16 register pairs of a block take the polynomial exponential (FMA pipe) instead of MUFU.EX2.
// In-run A/B on the DINOv2 shape (192 images, profiles/r2_attn_experiments.txt): 0/16 0.857 ms, 3/16 0.741, 4/16 0.699,
// 5/16 0.699, 6/16 +2 %: without the offload the MUFU binds, beyond 5/16 the issue slots do (a polynomial pair costs 14
// instructions against 4 for a MUFU pair).
// The decoder shape (d = 48: 3 + 3 MMAs per block instead of 4 + 4, so the softmax weighs more) prefers 5/16 (0.633 vs 0.649 ms).
#ifndef ATT_POLY_MASK
#define ATT_POLY_MASK(DV) ((DV) == 48 ? 0x4924 : 0x1248)
#endif
#ifndef ATT_REGS_SOFTMAX
#define ATT_REGS_SOFTMAX 96                             // setmaxnreg budgets (multiples of 8): 8*96 + 4*40 <= 12*80
#endif
#ifndef ATT_REGS_CTRL
#define ATT_REGS_CTRL 40
#endif
// ATT_DBG (development builds only, XS_BUILD_TAG + XS_BUILD_DEFS=-DATT_DBG=mask): timing experiments that BREAK the
// result.  1: no K/V TMA traffic after the first ring fill; 2: no exponentials; 4: no S loads; 8: no P stores;
// 16: no QK^T MMAs; 32: no PV MMAs.  The product library is built without it.
#ifndef ATT_DBG
#define ATT_DBG 0
#endif
// ATT_PROF (development builds only): per-phase clock64 totals of each role (lane 0), printed by the launcher.
#ifndef ATT_PROF
#define ATT_PROF 0
#endif
struct PhaseClock {
#if ATT_PROF
  long long t;
  unsigned long long acc[8];
  __device__ __forceinline__ void start() {
    t = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0;
  }
  __device__ __forceinline__ void lap(int i) {
    const long long n = clock64();
    acc[i] += static_cast<unsigned long long>(n - t);
    t = n;
  }
  __device__ __forceinline__ void flush(unsigned long long* out, int base, int lane) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(out + base + i, acc[i]);
    }
  }
#else
  __device__ __forceinline__ void start() {}
  __device__ __forceinline__ void lap(int) {}
  __device__ __forceinline__ void flush(unsigned long long*, int, int) {}
#endif
};
constexpr int ATT_MAX_TILES_PER_CTA = 1024;             // redo bitmap capacity
constexpr float ATT_L_MIN = 8.2718061e-25f;             // 2^-80
constexpr float ATT_L_MAX = 1.2676506e30f;              // 2^100
constexpr uint32_t ATT_Q_BYTES = 128 * 64 * 2;          // 16 KB: [128 rows][64 bf16], 128B swizzle
constexpr uint32_t ATT_KV_BYTES = ATT_BKV * 64 * 2;     // 8 KB:  [64 keys][64 bf16]
constexpr uint32_t ATT_SMEM_BYTES = ATT_Q_BYTES + 2 * ATT_ST * ATT_KV_BYTES + 512 + 1024;

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
  int all_safe;           // skip the optimistic pass (xs_attn_set_optimistic(0), or more tiles per CTA than the bitmap)
  unsigned long long* prof;  // ATT_PROF builds only
};

struct TagNo { static constexpr bool value = false; };
struct TagYes { static constexpr bool value = true; };

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

// 2^x for a pair on the FMA / ALU pipes.  x is clamped to [-125, 125] with ONE instruction per element
// (min.xorsign.abs -> FMNMX.XORSIGN: sign(x) * min(|x|, 125)): below, the result is ~2^-125 (nothing next to row sums
// >= 2^-80); above, it is ~2^125, which pushes the row sum past ATT_L_MAX so the tile is redone -- an unclamped exponent
// insert would wrap into the sign bit and go unnoticed.  Round-to-nearest split x = n + r with the 1.5 * 2^23 magic
// constant, cubic minimax polynomial for 2^r on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P).
__device__ __forceinline__ float clamp_sym125(float x) {
  float d;
  asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(x), "f"(125.0f));
  return d;
}
__device__ __forceinline__ float2 exp2_poly2_clamped(float2 x) {
  x.x = clamp_sym125(x.x);
  x.y = clamp_sym125(x.y);
  const float2 t = fadd2(x, make_float2(12582912.0f, 12582912.0f));
  const float2 nf = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = ffma2(nf, make_float2(-1.0f, -1.0f), x);
  float2 p = ffma2(make_float2(0.055171460f, 0.055171460f), r, make_float2(0.24261086f, 0.24261086f));
  p = ffma2(p, r, make_float2(0.69326097f, 0.69326097f));
  p = ffma2(p, r, make_float2(0.99992812f, 0.99992812f));
  float2 y;
  y.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  y.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return y;
}

template <int DQK_STEPS, int DV, bool SCALE1>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, AttnParams p) {
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
  const SmemBar s_free = p_full + ATT_NS;      // [ATT_NS] PV_g complete: S buffer g % ATT_NS may be overwritten
  const SmemBar o_full = s_free + ATT_NS;      // all PV of the tile complete
  const SmemBar o_empty = o_full + 1;          // O read out by the softmax warps: next tile's PV_0 may overwrite
  constexpr int N_BARS = 2 + 2 * ATT_ST + 3 * ATT_NS + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  uint32_t* redo = tmem_slot + 2;  // [ATT_MAX_TILES_PER_CTA / 32] bitmap over this CTA's tile sequence

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
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
      mbar_init(p_full + (s), 8);  // one arrival per softmax warp
      mbar_init(s_free + (s), 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    fence_mbar_init();
  }
  if (warp == 10) redo[lane] = 0u;  // 32 words = 1024 tiles
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + ATT_NS * 64;

  // All roles walk the same tile sequence twice: pass 0 handles every tile optimistically (no running max), pass 1
  // handles the tiles pass 0 marked in `redo` with the online softmax.  `g` counts K/V blocks over the CTA's lifetime
  // (ring slot g % ATT_ST, S buffer g % ATT_NS and the barrier phases follow from it), `n_done` counts processed tiles.
  auto selected = [&](int pass, int idx) -> bool {
    if (p.all_safe) return pass == 1;
    return pass == 0 || ((redo[idx >> 5] >> (idx & 31)) & 1u) != 0u;
  };

  if (warp >= 8) reg_dealloc<ATT_REGS_CTRL>();
  if (warp == 8) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    uint32_t g = 0, n_done = 0;
    PhaseClock pc;
    pc.start();
    for (int pass = 0; pass < 2; ++pass) {
      int idx = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++idx) {
        if (!selected(pass, idx)) continue;
        const TileCoord t = decode_tile(tile, p);
        const int b_kv = p.kv_shared ? 0 : t.b;
        pc.lap(3);
        mbar_wait(q_empty, (n_done & 1) ^ 1);  // previous tile's QK^T are done with the Q buffer
        pc.lap(0);
        if (elect_one_sync()) {
          mbar_expect_tx(q_full, ATT_Q_BYTES);
          tma_load_3d(smQ, &tmQ, q_full, t.h * 64, t.q0, t.b);
        }
        __syncwarp();
        uint32_t s = g % ATT_ST, ph_k = (g / ATT_ST) & 1;
        for (int j = 0; j < t.nkv; ++j, ++g) {
          const int kv0 = t.kv_begin + j * ATT_BKV;
          pc.lap(3);
          mbar_wait(kv_empty + (s), ph_k ^ 1);
          pc.lap(1);
          if (elect_one_sync()) {
            if ((ATT_DBG & 1) && g >= ATT_ST) {
              mbar_arrive(kv_full + (s));
            } else {
              mbar_expect_tx(kv_full + (s), 2 * ATT_KV_BYTES);
              tma_load_3d(smK + s * ATT_KV_BYTES, &tmK, kv_full + (s), t.h * 64, kv0, b_kv);
              tma_load_3d(smV + s * ATT_KV_BYTES, &tmV, kv_full + (s), t.h * 64, kv0, b_kv);
            }
          }
          __syncwarp();
          pc.lap(2);
          if (++s == ATT_ST) { s = 0; ph_k ^= 1; }
        }
        ++n_done;
      }
      if (pass == 0) named_bar_sync(1, ATT_THREADS);
    }
    pc.flush(p.prof, 16, lane);
  } else if (warp == 9) {
    // ===================== PV issuer (converged warp, uniform operands, one elected lane issues) =========
    // O += P_g V_g as soon as the softmax warps have handed P_g over.  Its completion frees the K/V ring slot (TMA
    // warp), the S buffer P_g lived in (QK issuer) and, for the last block, publishes O (softmax epilogue).
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, DV, 0, 1);  // B = V is MN-major ([kv][d], d contiguous)
    const uint32_t tb = warp_uniform(tmem_base);
    const uint32_t v_lo0 = umma_desc_lo(smem_u32(smV), 1024);
    PhaseClock pc;
    pc.start();
    uint32_t g = 0, n_done = 0;
    for (int pass = 0; pass < 2; ++pass) {
      int idx = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++idx) {
        if (!selected(pass, idx)) continue;
        const int nkv = decode_tile(tile, p).nkv;
        uint32_t s = g % ATT_ST, sb = g % ATT_NS, ph_s = (g / ATT_NS) & 1;  // running ring / buffer indices
        for (int j = 0; j < nkv; ++j, ++g) {
          pc.lap(6);
          mbar_wait(p_full + (sb), ph_s);  // softmax has turned S_sb into P_g
          pc.lap(1);
          if (j == 0) mbar_wait(o_empty, (n_done & 1) ^ 1);  // previous tile's O has been read out
          pc.lap(2);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t v_lo = v_lo0 + s * (ATT_KV_BYTES >> 4);
            const uint32_t a_p = tb + sb * 64 + 32;  // P_g lives in the upper half of S_sb
            if (!(ATT_DBG & 32)) {
#pragma unroll
              for (int k = 0; k < ATT_BKV / 16; ++k) {
                // A: 16 bf16 of P per row = 8 TMEM columns per K-step; B: 16 kv rows x 128 B = 2048 B per K-step
                umma_ts_lh(tb + ATT_NS * 64, a_p + k * 8, v_lo + k * 128, idesc_pv, (j | k) != 0 ? 1u : 0u);
              }
            }
            tc_commit(kv_empty + (s));  // K_g / V_g slot free; also the "PV_g complete" signal for the O rescale
            tc_commit(s_free + (sb));   // S_sb may be overwritten by QK_{g+ATT_NS}
            if (j == nkv - 1) tc_commit(o_full);
          }
          __syncwarp();
          pc.lap(3);
          if (++s == ATT_ST) s = 0;
          if (++sb == ATT_NS) { sb = 0; ph_s ^= 1; }
        }
        ++n_done;
      }
      if (pass == 0) named_bar_sync(1, ATT_THREADS);
    }
    pc.flush(p.prof, 8, lane);
  } else if (warp == 10) {
    // ===================== QK issuer: S_sb = Q K_g^T, running up to ATT_NS blocks ahead of the softmax ==========
    // Two issuing warps instead of one: every mbarrier wait costs 100-200 clk even when the barrier completed long
    // ago and a tcgen05.mma issue blocks until the (shallow) tensor queue takes it, so a single warp that serialises
    // [wait P, 4 PV, commits, wait K, 4 QK, commit] needed ~1200 clk per block and paced the whole CTA (phase clocks,
    // profiles/r2_attn_phase_clocks.txt).
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, ATT_BKV, 0, 0);
    const uint32_t tb = warp_uniform(tmem_base);
    const uint32_t q_lo = umma_desc_lo(smem_u32(smQ), 16);
    const uint32_t k_lo0 = umma_desc_lo(smem_u32(smK), 16);
    PhaseClock pc;
    pc.start();
    uint32_t g = 0, n_done = 0;
    for (int pass = 0; pass < 2; ++pass) {
      int idx = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++idx) {
        if (!selected(pass, idx)) continue;
        const int nkv = decode_tile(tile, p).nkv;
        pc.lap(6);
        mbar_wait(q_full, n_done & 1);
        pc.lap(0);
        uint32_t s = g % ATT_ST, ph_k = (g / ATT_ST) & 1, sb = g % ATT_NS, ph_s = (g / ATT_NS) & 1;
        for (int j = 0; j < nkv; ++j, ++g) {
          // both polls are issued before either result is used, so their latencies overlap.  S_sb is free once
          // PV_{g-ATT_NS} has completed (the wait on the barrier's previous phase passes at once for g < ATT_NS)
          const bool k_ok = mbar_try_wait(kv_full + (s), ph_k);
          const bool s_ok = mbar_try_wait(s_free + (sb), ph_s ^ 1);
          if (!k_ok) mbar_wait(kv_full + (s), ph_k);
          if (!s_ok) mbar_wait(s_free + (sb), ph_s ^ 1);
          pc.lap(4);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t k_lo = k_lo0 + s * (ATT_KV_BYTES >> 4);
            const uint32_t d_s = tb + sb * 64;
            if (!(ATT_DBG & 16)) {
#pragma unroll
              for (int k = 0; k < DQK_STEPS; ++k) umma_ss_lh<false>(d_s, q_lo + 2 * k, k_lo + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
            }
            tc_commit(s_full + (sb));
            if (j == nkv - 1) tc_commit(q_empty);  // Q buffer may be refilled once these MMAs have read it
          }
          __syncwarp();
          pc.lap(5);
          if (++s == ATT_ST) { s = 0; ph_k ^= 1; }
          if (++sb == ATT_NS) { sb = 0; ph_s ^= 1; }
        }
        ++n_done;
      }
      if (pass == 0) named_bar_sync(1, ATT_THREADS);
    }
    pc.flush(p.prof, 20, lane);
  } else if (warp == 11) {
    named_bar_sync(1, ATT_THREADS);
  } else {
    // ===================== softmax / epilogue: warp w owns rows [32 (w%4) + 16 (w/4), +16) of the tile ==========
    reg_alloc<ATT_REGS_SOFTMAX>();
    const int lane_base = (warp & 3) * 32 + (warp >> 2) * 16;
    const uint32_t lane_off = static_cast<uint32_t>(lane_base) << 16;
    const int rA = lane_base + (lane >> 2);  // this thread's rows: rA and rA + 8
    const int cq = (lane & 3) * 2;           // its two columns inside every 8-column group
    const uint32_t t_o = tmem_O + lane_off;
    const float sl2 = p.scale_log2;
    uint32_t g0 = 0, n_done = 0;
    PhaseClock pc;
    pc.start();

    auto tile_fn = [&](const int tile, const int idx, auto safe_tag) {
      constexpr bool SAFE = decltype(safe_tag)::value;
      int nkv, tail_valid;
      {
        const TileCoord t = decode_tile(tile, p);
        nkv = t.nkv;
        tail_valid = t.kv_end - t.kv_begin - (t.nkv - 1) * ATT_BKV;  // valid columns of the last block
      }
      float mA = -INFINITY, mB = -INFINITY;  // SAFE: running row maxima (log2 domain)
      float2 lA = make_float2(0.f, 0.f), lB = make_float2(0.f, 0.f);  // per-thread partial row sums

      uint32_t sb = g0 % ATT_NS, ph_s = (g0 / ATT_NS) & 1;  // running S buffer index and barrier phase of the block
      auto block = [&](const int j, auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        const uint32_t t_s = tmem_base + lane_off + sb * 64;
        uint32_t v[32], pk[16];
        pc.lap(7);
        mbar_wait(s_full + (sb), ph_s);
        pc.lap(0);
        tc_fence_after();
        if (!(ATT_DBG & 4) || j == 0) tmem_ld_16x256b_x8(t_s, v);
        tmem_ld_wait32(v);
        pc.lap(1);
        if constexpr (MASK) {  // last block: columns >= tail_valid are past the sequence end
#pragma unroll
          for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (8 * k + cq + e >= tail_valid) {
                v[4 * k + e] = 0xff800000u;
                v[4 * k + 2 + e] = 0xff800000u;
              }
            }
          }
        }
        if constexpr (!SAFE) {
          // register pair i = (v[2i], v[2i+1]): even i belongs to row A, odd i to row B
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 x = make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            if constexpr (!SCALE1) x = ffma2(x, make_float2(sl2, sl2), make_float2(0.f, 0.f));
            float2 a;
            if (ATT_DBG & 2) {
              a = x;
            } else if (!MASK && ((ATT_POLY_MASK(DV) >> i) & 1)) {
              a = exp2_poly2_clamped(x);
            } else {
              a.x = fast_exp2(x.x);
              a.y = fast_exp2(x.y);
            }
            if (i & 1) lB = fadd2(lB, a);
            else lA = fadd2(lA, a);
            pk[i] = pack_bf16x2(a.x, a.y);
          }
        } else {
          float xa = -INFINITY, xb = -INFINITY;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            xa = fmaxf(xa, fmaxf(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1])));
            xb = fmaxf(xb, fmaxf(__uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3])));
          }
          xa = fmaxf(xa, __shfl_xor_sync(0xffffffffu, xa, 1));
          xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, 1));
          xa = fmaxf(xa, __shfl_xor_sync(0xffffffffu, xa, 2));
          xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, 2));
          const float nA = fmaxf(mA, xa * sl2), nB = fmaxf(mB, xb * sl2);
          const float alA = fast_exp2(mA - nA), alB = fast_exp2(mB - nB);  // 1 when unchanged, 0 when m was -inf
          if (j > 0 && __any_sync(0xffffffffu, (nA != mA) || (nB != mB))) {
            // O must hold PV of all earlier blocks of this tile: wait for the previous block's PV through its
            // K/V slot's kv_empty phase (the slot is refilled only ATT_ST blocks later: the parity cannot alias)
            const uint32_t g = g0 + j;
            mbar_wait(kv_empty + ((g - 1) % ATT_ST), ((g - 1) / ATT_ST) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < DV / 16; ++c) {
              uint32_t o[8];
              tmem_ld_16x256b_x2(t_o + c * 16, o);
              tmem_ld_wait8(o);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                o[i] = __float_as_uint(__uint_as_float(o[i]) * ((i & 2) ? alB : alA));
              tmem_st_16x256b_x2(t_o + c * 16, o);
            }
          }
          lA.x *= alA;
          lA.y *= alA;
          lB.x *= alB;
          lB.y *= alB;
          mA = nA;
          mB = nB;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = ffma2_bcast(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]), sl2, (i & 1) ? -mB : -mA);
            float2 a;
            a.x = fast_exp2(x.x);
            a.y = fast_exp2(x.y);
            if (i & 1) lB = fadd2(lB, a);
            else lA = fadd2(lA, a);
            pk[i] = pack_bf16x2(a.x, a.y);
          }
        }
        // P_g overwrites the upper half of S_sb (the whole block is in registers by now)
        pc.lap(2);
        if (!(ATT_DBG & 8)) tmem_st_16x128b_x8(t_s + 32, pk);
        tc_wait_st();
        pc.lap(3);
        tc_fence_before();
        __syncwarp();
        if (elect_one_sync()) mbar_arrive(p_full + sb);
        pc.lap(4);
        if (++sb == ATT_NS) {
          sb = 0;
          ph_s ^= 1;
        }
      };
      for (int j = 0; j + 1 < nkv; ++j) block(j, TagNo{});
      block(nkv - 1, TagYes{});
      g0 += nkv;

      // ---- epilogue: read O out of TMEM (frees it for the next tile's PV_0), then O / l and log-sum-exp ----
      pc.lap(7);
      mbar_wait(o_full, n_done & 1);
      pc.lap(5);
      tc_fence_after();
      uint32_t o[32];
      if constexpr (DV == 64) {
        tmem_ld_16x256b_x8(t_o, o);
      } else {  // DV == 48: groups 0..3, then 4..5
        tmem_ld_16x256b_x4(t_o, o);
        tmem_ld_16x256b_x2_hi(t_o + 32, o);
      }
      tmem_ld_wait32(o);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      ++n_done;
      float la = lA.x + lA.y, lb = lB.x + lB.y;
      la += __shfl_xor_sync(0xffffffffu, la, 1);
      lb += __shfl_xor_sync(0xffffffffu, lb, 1);
      la += __shfl_xor_sync(0xffffffffu, la, 2);
      lb += __shfl_xor_sync(0xffffffffu, lb, 2);
      if constexpr (!SAFE) {
        // every row sum must sit well inside the fp32 range, else 2^S overflowed / flushed somewhere: redo the tile
        const bool bad = !ATT_DBG && (!(la >= ATT_L_MIN && la <= ATT_L_MAX) || !(lb >= ATT_L_MIN && lb <= ATT_L_MAX));
        if (__any_sync(0xffffffffu, bad)) {
          if (lane == 0) atomicOr(&redo[idx >> 5], 1u << (idx & 31));
          return;  // pass 2 writes this tile
        }
      }
      const float invA = 1.0f / la, invB = 1.0f / lb;
      const TileCoord t = decode_tile(tile, p);
      const int rowA = t.q0 + rA, rowB = rowA + 8;
      const long long base = static_cast<long long>(t.split) * p.o_split_stride +
                             static_cast<long long>(t.b) * p.o_batch_stride + static_cast<long long>(t.h) * DV + cq;
      const long long offA = base + static_cast<long long>(rowA) * p.o_row_stride;
      const long long offB = base + static_cast<long long>(rowB) * p.o_row_stride;
      if (p.o_is_f32) {
        float* dst = reinterpret_cast<float*>(p.o);
#pragma unroll
        for (int k = 0; k < DV / 8; ++k) {
          if (rowA < p.Lq)
            *reinterpret_cast<float2*>(dst + offA + 8 * k) =
                make_float2(__uint_as_float(o[4 * k]) * invA, __uint_as_float(o[4 * k + 1]) * invA);
          if (rowB < p.Lq)
            *reinterpret_cast<float2*>(dst + offB + 8 * k) =
                make_float2(__uint_as_float(o[4 * k + 2]) * invB, __uint_as_float(o[4 * k + 3]) * invB);
        }
      } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.o);
#pragma unroll
        for (int k = 0; k < DV / 8; ++k) {
          if (rowA < p.Lq)
            *reinterpret_cast<uint32_t*>(dst + offA + 8 * k) =
                pack_bf16x2(__uint_as_float(o[4 * k]) * invA, __uint_as_float(o[4 * k + 1]) * invA);
          if (rowB < p.Lq)
            *reinterpret_cast<uint32_t*>(dst + offB + 8 * k) =
                pack_bf16x2(__uint_as_float(o[4 * k + 2]) * invB, __uint_as_float(o[4 * k + 3]) * invB);
        }
      }
      if (p.lse != nullptr && (lane & 3) == 0) {
        // natural-log LSE of the scaled logits: ln sum_j exp(s_j * scale)
        float* lse = p.lse + static_cast<long long>(t.split) * p.lse_split_stride +
                     (static_cast<long long>(t.b) * p.heads + t.h) * p.Lq;
        const float m0A = SAFE ? mA : 0.f, m0B = SAFE ? mB : 0.f;
        if (rowA < p.Lq) lse[rowA] = (m0A + log2f(la)) * 0.6931471805599453f;
        if (rowB < p.Lq) lse[rowB] = (m0B + log2f(lb)) * 0.6931471805599453f;
      }
      pc.lap(6);
    };

    if (!p.all_safe) {
      int idx = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++idx) tile_fn(tile, idx, TagNo{});
    }
    named_bar_sync(1, ATT_THREADS);
    {
      int idx = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++idx)
        if (selected(1, idx)) tile_fn(tile, idx, TagYes{});
    }
    pc.flush(p.prof, 0, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// xs_attn_set_optimistic(0) forces every tile through the online-softmax pass (tests exercise it on benign inputs)
static std::atomic<int> g_attn_optimistic{1};
void attn_set_optimistic(int enable) { g_attn_optimistic.store(enable ? 1 : 0); }
int attn_optimistic_enabled() { return g_attn_optimistic.load(); }
// 0: this kernel (64-key blocks, two CTAs per SM); 1: the pair kernel of xs_attn_tc2.cu where it applies
static std::atomic<int> g_attn_layout{1};  // 1: pair kernel (xs_attn_tc2.cu), the default; 0: this file's kernel
void attn_set_layout(int layout) { g_attn_layout.store(layout); }
int flash_attn_bf16_pair(const void*, const void*, const void*, void*, float*, int, int, int, int, int, long long, long long,
                         long long, long long, int, int, int, float, cudaStream_t);

template <int DQK, int DV, bool SCALE1>
static int launch_attn(dim3 grid, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                       const AttnParams& p, cudaStream_t stream) {
  auto kern = attn_tc_kernel<DQK, DV, SCALE1>;
  XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM_BYTES));
#if ATT_PROF
  static unsigned long long* buf = nullptr;
  if (buf == nullptr) XS_CUDA(cudaMalloc(&buf, 32 * sizeof(unsigned long long)));
  XS_CUDA(cudaMemsetAsync(buf, 0, 32 * sizeof(unsigned long long), stream));
  AttnParams pp = p;
  pp.prof = buf;
  kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, pp);
  XS_LAUNCH_CHECK();
  XS_CUDA(cudaStreamSynchronize(stream));
  unsigned long long h[32];
  XS_CUDA(cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost));
  const double nb = double(p.n_tiles) * ((p.split_len < p.Lk ? p.split_len : p.Lk) + ATT_BKV - 1) / ATT_BKV;  // blocks
  fprintf(stderr, "attn prof, clk per 64-key block | softmax warp (avg of 8): wait_S %.0f  ld %.0f  math %.0f  st+wait %.0f  "
                  "arrive %.0f  wait_O %.0f  epilogue %.0f  other %.0f | pv warp: wait_P %.0f  wait_Oempty %.0f  issue_PV %.0f  "
                  "other %.0f | qk warp: wait_Q %.0f  wait_KV+Sfree %.0f  issue_QK %.0f  other %.0f | tma warp: wait_Qempty %.0f  "
                  "wait_KVempty %.0f  issue %.0f  other %.0f\n",
          h[0] / nb / 8, h[1] / nb / 8, h[2] / nb / 8, h[3] / nb / 8, h[4] / nb / 8, h[5] / nb / 8, h[6] / nb / 8, h[7] / nb / 8,
          h[9] / nb, h[10] / nb, h[11] / nb, h[14] / nb, h[20] / nb, h[24] / nb, h[25] / nb, h[26] / nb, h[16] / nb, h[17] / nb,
          h[18] / nb, h[19] / nb);
#else
  kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  XS_LAUNCH_CHECK();
#endif
  return 0;
}

// q/k/v: bf16, head h occupies 64 consecutive columns starting at h*64 of its row (d=48: 48 used + 16 pad)
// strides in elements.  o: [nsplit][B][Lq][heads*head_dim] (bf16, or fp32 when o_is_f32)
int flash_attn_bf16_tc(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq,
                       int Lk, int head_dim, long long q_row_stride, long long q_batch_stride,
                       long long kv_row_stride, long long kv_batch_stride, int kv_shared, int nsplit, int o_is_f32,
                       float scale, cudaStream_t stream) {
  XS_CHECK_ARG(head_dim == 64 || head_dim == 48, "flash_attn: head_dim %d not supported (64 or 48)", head_dim);
  XS_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk > 0 && nsplit > 0, "flash_attn: empty problem");
  XS_CHECK_ARG(scale > 0.f, "flash_attn: scale must be positive");
  XS_CHECK_ARG((q_row_stride % 8) == 0 && (kv_row_stride % 8) == 0 && (q_batch_stride % 8) == 0 &&
                   (kv_batch_stride % 8) == 0,
               "flash_attn: strides must be multiples of 8 elements");
  XS_CHECK_ARG(nsplit == 1 || (o_is_f32 && lse != nullptr), "flash_attn: split-KV needs fp32 partial O and LSE");
  if (g_attn_layout.load() == 1) {
    const int rc = flash_attn_bf16_pair(q, k, v, o, lse, B, heads, Lq, Lk, head_dim, q_row_stride, q_batch_stride,
                                        kv_row_stride, kv_batch_stride, kv_shared, nsplit, o_is_f32, scale, stream);
    if (rc != 1) return rc;
  }
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
  p.prof = nullptr;
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
  const long long n_tiles = (long long)p.nq_tiles * heads * B * nsplit;
  XS_CHECK_ARG(n_tiles < (1ll << 31), "flash_attn: too many tiles");
  p.n_tiles = (int)n_tiles;
  const int max_ctas = 2 * num_sms();  // two co-resident CTAs per SM, each walking its share of the tiles
  dim3 grid(p.n_tiles < max_ctas ? p.n_tiles : max_ctas);
  const int per_cta = (p.n_tiles + (int)grid.x - 1) / (int)grid.x;
  p.all_safe = (g_attn_optimistic.load() == 0 || per_cta > ATT_MAX_TILES_PER_CTA) ? 1 : 0;
  // scale * log2(e) == 1: the caller folded the softmax scale into the query projection (logits arrive in the log2 domain)
  const bool scale1 = fabsf(p.scale_log2 - 1.0f) < 1e-6f;
  if (scale1) p.scale_log2 = 1.0f;
  if (head_dim == 64)
    return scale1 ? launch_attn<4, 64, true>(grid, tmQ, tmK, tmV, p, stream)
                  : launch_attn<4, 64, false>(grid, tmQ, tmK, tmV, p, stream);
  return scale1 ? launch_attn<3, 48, true>(grid, tmQ, tmK, tmV, p, stream)
                : launch_attn<3, 48, false>(grid, tmQ, tmK, tmV, p, stream);
}

}  // namespace xs
