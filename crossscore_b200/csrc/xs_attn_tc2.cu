// Flash attention, "pair" layout (the default, xs_attn_set_layout(1)): ONE CTA per SM works on TWO 128-row query tiles
// at a time, in 128-key blocks -- normally two adjacent tiles of a (batch, head), which share every K/V tile; when the
// tile count is odd, the last tiles of two different heads, each with its own K/V stages (struct Unit below).
// O = softmax(Q K^T * scale) V, same contract as xs_attn_tc.cu (which keeps the 64-key / two-CTAs-per-SM layout,
// selectable with xs_attn_set_layout(0)); reference call sites: modeling_dinov2.py:203-234 (DINOv2 SDPA),
// torch/nn/functional.py:6630-6692 via model/customised_transformer/transformer.py:182-205 (decoder).
//
// Why: the 64-key kernel is bounded by issue slots and the MUFU together (ncu: issue 58 %, MUFU 67 %, tensor 42 %), and
// 28 % of its instructions are per-block overhead -- three control warps (TMA, PV issuer, QK issuer) at ~135
// instructions per 64-key block each, plus ~30 per softmax warp and block.  With 128-key blocks shared by two query tiles
// the TMA work per (query tile x key) drops 4x, the MMA-issue and softmax hand-over work 2x, the L2 -> SM K/V traffic
// 2x, and QK^T runs at N = 128, where the operand fetch from shared memory no longer limits the MMA (48 clk per N = 64
// MMA instead of 32).
//
// 640 threads:
//   warps 0-7   softmax of query tile A, warps 8-15 of tile B: warp w owns 16 rows (TMEM lane quarter w % 4, lower / upper
//               half by (w / 4) % 2), tcgen05.ld.16x256b fragments, 64 logits per thread and block in two halves
//   warp 16     TMA producer: Q_A, Q_B per unit; K / V 128-key tiles through a 5-stage ring
//   warp 17     PV issuer  (O_t += P_t V, 8 K-steps per block and tile; P read from TMEM)
//   warp 18     QK issuer  (S_t = Q_t K^T, N = 128)
//   warp 19     TMEM allocator
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
// S is single-buffered per tile but P has its own columns: a softmax warp releases S_t as soon as both halves of the
// block are in registers (s_free), so QK_{j+1} of that tile runs under the exponentials of block j, and PV_j reads P_t
// while S_t is already being overwritten.  Both tiles run the same chain, in phase (a forced offset measured slower:
// profiles/r2_attn_experiments.txt item 12).
// Softmax arithmetic, redo pass, masking, split-KV outputs: as in xs_attn_tc.cu (max-free first pass, online-softmax
// redo of tiles whose row sums leave [2^-80, 2^100]).
#include "xs_common.cuh"

namespace xs {

constexpr int AT2_THREADS = 640;
constexpr int AT2_BKV = 128;                              // keys per block
#ifndef AT2_STAGES
#define AT2_STAGES 5
#endif
constexpr int AT2_ST = AT2_STAGES;                        // K/V ring depth
#ifndef AT2_POLY_MASK
#define AT2_POLY_MASK(DV) ((DV) == 48 ? 0x4924 : 0x4924)  // per 32-logit half: pairs on the FMA-pipe polynomial
#endif
constexpr int AT2_MAX_UNITS_PER_CTA = 1024;
// Launched at 96 registers per thread; setmaxnreg then moves registers from the control warpgroup to the 16 softmax warps
// inside the CTA's own pool of 640 x 96 = 61 440: 512 x 104 + 128 x 64 = 61 440 (112 / 32 spills in the control warps).
#ifndef AT2_REGS_SOFTMAX
#define AT2_REGS_SOFTMAX 104
#define AT2_REGS_CTRL 64
#endif
static_assert(512 * AT2_REGS_SOFTMAX + 128 * AT2_REGS_CTRL <= 640 * 96, "setmaxnreg budget exceeds the CTA's register pool");
constexpr float AT2_L_MIN = 8.2718061e-25f;               // 2^-80
constexpr float AT2_L_MAX = 1.2676506e30f;                // 2^100
constexpr uint32_t AT2_Q_BYTES = 128 * 64 * 2;            // 16 KB per query tile
constexpr uint32_t AT2_KV_BYTES = AT2_BKV * 64 * 2;       // 16 KB: [128 keys][64 bf16]
constexpr uint32_t AT2_SMEM_BYTES = 2 * AT2_Q_BYTES + 2 * AT2_ST * AT2_KV_BYTES + 1024 + 1024;

struct Attn2Params {
  void* o;
  float* lse;
  int o_is_f32;
  int Lq, Lk, heads;
  int kv_shared;
  int nsplit, split_len;
  long long o_row_stride, o_batch_stride, o_split_stride;  // elements
  long long lse_split_stride;
  float scale_log2;
  int n_pairs_full;      // full pairs of 128-query tiles per (batch, head)
  int units_per_bs;      // units per (batch, split): heads * n_pairs_full shared + ceil(heads / 2) mixed when the tile count is odd
  int n_units;           // total units
  int all_safe;
};

// A unit = two 128-row query tiles processed together by one CTA.  "Shared" units pair adjacent tiles of ONE (batch,
// head): both read the same K/V stages.  When the number of query tiles is odd (T = 1370 -> 11), the last tiles of two
// DIFFERENT heads form a "mixed" unit: each tile then streams its own head's K/V through its own ring stages (no
// sharing, but no idle half either -- a lone tile would leave half of the softmax warps without work for a sixth of
// the time).  A mixed unit's second tile is absent when the head count is odd.
struct Unit {
  int b, split, kv_begin, kv_end, nkb;
  int q0a, q0b, ha, hb;  // per tile (scalars: a runtime-indexed array would live in local memory)
  bool valid_b;          // tile A is always present
  bool shared;
  __device__ __forceinline__ int q0(int tt) const { return tt ? q0b : q0a; }
  __device__ __forceinline__ int h(int tt) const { return tt ? hb : ha; }
  __device__ __forceinline__ bool valid(int tt) const { return tt ? valid_b : true; }
};
__device__ __forceinline__ Unit decode_unit(int u, const Attn2Params& p) {
  Unit t;
  const int per_bs = p.units_per_bs;
  const int loc = u % per_bs;
  int r = u / per_bs;
  t.split = r % p.nsplit;
  t.b = r / p.nsplit;
  const int n_shared = p.heads * p.n_pairs_full;
  if (loc < n_shared) {
    const int pr = loc % p.n_pairs_full;
    t.ha = t.hb = loc / p.n_pairs_full;
    t.q0a = pr * 256;
    t.q0b = pr * 256 + 128;
    t.valid_b = true;
    t.shared = true;
  } else {
    const int m = loc - n_shared;
    t.ha = 2 * m;
    t.hb = 2 * m + 1;
    t.q0a = t.q0b = p.n_pairs_full * 256;  // the odd last tile of each head
    t.valid_b = t.hb < p.heads;
    if (!t.valid_b) t.hb = t.ha;
    t.shared = false;
  }
  t.kv_begin = t.split * p.split_len;
  t.kv_end = min(p.Lk, t.kv_begin + p.split_len);
  t.nkb = (t.kv_end - t.kv_begin + AT2_BKV - 1) / AT2_BKV;
  return t;
}

struct T2No { static constexpr bool value = false; };
struct T2Yes { static constexpr bool value = true; };

__device__ __forceinline__ float clamp_sym125_2(float x) {
  float d;
  asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(x), "f"(125.0f));
  return d;
}
// 2^x for a pair on the FMA / ALU pipes (see xs_attn_tc.cu: symmetric clamp, magic-constant split, cubic polynomial)
__device__ __forceinline__ float2 exp2_poly2_c2(float2 x) {
  x.x = clamp_sym125_2(x.x);
  x.y = clamp_sym125_2(x.y);
  const float2 t = fadd2(x, make_float2(12582912.0f, 12582912.0f));
  const float2 nf = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = ffma2(nf, make_float2(-1.0f, -1.0f), x);
  float2 pl = ffma2(make_float2(0.055171460f, 0.055171460f), r, make_float2(0.24261086f, 0.24261086f));
  pl = ffma2(pl, r, make_float2(0.69326097f, 0.69326097f));
  pl = ffma2(pl, r, make_float2(0.99992812f, 0.99992812f));
  float2 y;
  y.x = __int_as_float(__float_as_int(pl.x) + (__float_as_int(t.x) << 23));
  y.y = __int_as_float(__float_as_int(pl.y) + (__float_as_int(t.y) << 23));
  return y;
}

template <int DQK_STEPS, int DV, bool SCALE1>
__global__ void __launch_bounds__(AT2_THREADS, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem;                                  // [2 tiles][16 KB]
  uint8_t* smK = smem + 2 * AT2_Q_BYTES;                // [AT2_ST][16 KB]
  uint8_t* smV = smK + AT2_ST * AT2_KV_BYTES;           // [AT2_ST][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smV + AT2_ST * AT2_KV_BYTES);
  const SmemBar bar0{smem_u32(bars)};
  // per tile t: index + t
  const SmemBar q_full = bar0 + 0;     // [2] Q tile landed
  const SmemBar q_empty = bar0 + 2;    // [2] last QK^T of the unit complete
  const SmemBar s_full = bar0 + 4;     // [2] S_t complete
  const SmemBar s_free = bar0 + 6;     // [2] both halves of S_t are in the softmax warps' registers (8 arrivals)
  const SmemBar p_full = bar0 + 8;     // [2] P_t stored (8 arrivals)
  const SmemBar p_free = bar0 + 10;    // [2] PV_t complete: P_t may be overwritten (also "O_t holds all earlier blocks")
  const SmemBar o_full = bar0 + 12;    // [2] all PV of the unit complete
  const SmemBar o_empty = bar0 + 14;   // [2] O_t read out (8 arrivals)
  const SmemBar kv_full = bar0 + 16;   // [AT2_ST]
  const SmemBar kv_empty = kv_full + AT2_ST;
  constexpr int N_BARS = 16 + 2 * AT2_ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  uint32_t* redo = tmem_slot + 2;  // [AT2_MAX_UNITS_PER_CTA / 32]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int t = 0; t < 2; ++t) {
      mbar_init(q_full + t, 1);
      mbar_init(q_empty + t, 1);
      mbar_init(s_full + t, 1);
      mbar_init(s_free + t, 8);
      mbar_init(p_full + t, 8);
      mbar_init(p_free + t, 1);
      mbar_init(o_full + t, 1);
      mbar_init(o_empty + t, 8);
    }
    for (int s = 0; s < AT2_ST; ++s) {
      mbar_init(kv_full + s, 1);
      mbar_init(kv_empty + s, 1);
    }
    fence_mbar_init();
  }
  if (warp == 18) redo[lane] = 0u;
  if (warp == 19) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // column map
  constexpr uint32_t COL_S = 0, COL_O = 256, COL_P = 384;

  auto selected = [&](int pass, int idx) -> bool {
    if (p.all_safe) return pass == 1;
    return pass == 0 || ((redo[idx >> 5] >> (idx & 31)) & 1u) != 0u;
  };

  if (warp >= 16) reg_dealloc<AT2_REGS_CTRL>();
  if (warp == 16) {
    // ===================== TMA producer =====================
    uint32_t g = 0;          // 128-key blocks over the CTA's lifetime
    uint32_t nu[2] = {0, 0};  // units processed per tile (tile B only counts units in which it is valid)
    for (int pass = 0; pass < 2; ++pass) {
      int idx = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++idx) {
        if (!selected(pass, idx)) continue;
        const Unit t = decode_unit(u, p);
        const int b_kv = p.kv_shared ? 0 : t.b;
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          if (!t.valid(tt)) continue;
          mbar_wait(q_empty + tt, (nu[tt] & 1) ^ 1);  // previous unit's QK^T are done with this Q buffer
          if (elect_one_sync()) {
            mbar_expect_tx(q_full + tt, AT2_Q_BYTES);
            tma_load_3d(smQ + tt * AT2_Q_BYTES, &tmQ, q_full + tt, t.h(tt) * 64, t.q0(tt), t.b);
          }
          __syncwarp();
          ++nu[tt];
        }
        uint32_t s = g % AT2_ST, ph = (g / AT2_ST) & 1;
        for (int j = 0; j < t.nkb; ++j) {
          const int kv0 = t.kv_begin + j * AT2_BKV;
          for (int tt = 0; tt < 2; ++tt) {  // shared: one stage for both tiles; mixed: one stage per tile
            if (!t.valid(tt) || (t.shared && tt == 1)) continue;
            mbar_wait(kv_empty + s, ph ^ 1);
            if (elect_one_sync()) {
              mbar_expect_tx(kv_full + s, 2 * AT2_KV_BYTES);
              tma_load_3d(smK + s * AT2_KV_BYTES, &tmK, kv_full + s, t.h(tt) * 64, kv0, b_kv);
              tma_load_3d(smV + s * AT2_KV_BYTES, &tmV, kv_full + s, t.h(tt) * 64, kv0, b_kv);
            }
            __syncwarp();
            ++g;
            if (++s == AT2_ST) { s = 0; ph ^= 1; }
          }
        }
      }
      if (pass == 0) named_bar_sync(1, AT2_THREADS);
    }
  } else if (warp == 17) {
    // ===================== PV issuer =====================
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, DV, 0, 1);  // B = V is MN-major ([kv][d], d contiguous)
    const uint32_t tb = warp_uniform(tmem_base);
    const uint32_t v_lo0 = umma_desc_lo(smem_u32(smV), 1024);
    uint32_t g = 0;
    uint32_t cb[2] = {0, 0};  // blocks processed per tile
    uint32_t nu[2] = {0, 0};
    for (int pass = 0; pass < 2; ++pass) {
      int idx = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++idx) {
        if (!selected(pass, idx)) continue;
        const Unit t = decode_unit(u, p);
        uint32_t s = g % AT2_ST;
        const int last_tt = t.valid_b ? 1 : 0;
        for (int j = 0; j < t.nkb; ++j) {
          for (int tt = 0; tt < 2; ++tt) {
            if (!t.valid(tt)) continue;
            mbar_wait(p_full + tt, cb[tt] & 1);  // softmax has written P_t of this block
            if (j == 0) mbar_wait(o_empty + tt, (nu[tt] & 1) ^ 1);  // previous unit's O_t has been read out
            tc_fence_after();
            const bool stage_done = !t.shared || tt == last_tt;  // last reader of this K/V stage
            if (elect_one_sync()) {
              const uint32_t v_lo = v_lo0 + s * (AT2_KV_BYTES >> 4);
              const uint32_t a_p = tb + COL_P + tt * 64;
              const uint32_t d_o = tb + COL_O + tt * 64;
#pragma unroll
              for (int k = 0; k < AT2_BKV / 16; ++k)  // A: 16 bf16 of P per row = 8 TMEM columns per K-step
                umma_ts_lh(d_o, a_p + k * 8, v_lo + k * 128, idesc_pv, (j | k) != 0 ? 1u : 0u);
              tc_commit(p_free + tt);
              if (j == t.nkb - 1) tc_commit(o_full + tt);
              // the stage is free once its last PV has completed (the QK^T that read K finished long before)
              if (stage_done) tc_commit(kv_empty + s);
            }
            __syncwarp();
            ++cb[tt];
            if (stage_done) {
              ++g;
              if (++s == AT2_ST) s = 0;
            }
          }
        }
        ++nu[0];
        if (t.valid_b) ++nu[1];
      }
      if (pass == 0) named_bar_sync(1, AT2_THREADS);
    }
  } else if (warp == 18) {
    // ===================== QK issuer =====================
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, AT2_BKV, 0, 0);
    const uint32_t tb = warp_uniform(tmem_base);
    const uint32_t q_lo0 = umma_desc_lo(smem_u32(smQ), 16);
    const uint32_t k_lo0 = umma_desc_lo(smem_u32(smK), 16);
    uint32_t g = 0;
    uint32_t cb[2] = {0, 0};
    uint32_t nu[2] = {0, 0};
    for (int pass = 0; pass < 2; ++pass) {
      int idx = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++idx) {
        if (!selected(pass, idx)) continue;
        const Unit t = decode_unit(u, p);
        uint32_t s = g % AT2_ST, ph = (g / AT2_ST) & 1;
        const int last_tt = t.valid_b ? 1 : 0;
        for (int j = 0; j < t.nkb; ++j) {
          for (int tt = 0; tt < 2; ++tt) {
            if (!t.valid(tt)) continue;
            if (!t.shared || tt == 0) mbar_wait(kv_full + s, ph);  // K (and V) of this stage have landed
            if (j == 0) mbar_wait(q_full + tt, nu[tt] & 1);
            mbar_wait(s_free + tt, (cb[tt] & 1) ^ 1);  // the previous block's S_t has been read into registers
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t k_lo = k_lo0 + s * (AT2_KV_BYTES >> 4);
              const uint32_t q_lo = q_lo0 + tt * (AT2_Q_BYTES >> 4);
              const uint32_t d_s = tb + COL_S + tt * 128;
#pragma unroll
              for (int k = 0; k < DQK_STEPS; ++k) umma_ss_lh<false>(d_s, q_lo + 2 * k, k_lo + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
              tc_commit(s_full + tt);
              if (j == t.nkb - 1) tc_commit(q_empty + tt);
            }
            __syncwarp();
            ++cb[tt];
            if (!t.shared || tt == last_tt) {
              ++g;
              if (++s == AT2_ST) { s = 0; ph ^= 1; }
            }
          }
        }
        ++nu[0];
        if (t.valid_b) ++nu[1];
      }
      if (pass == 0) named_bar_sync(1, AT2_THREADS);
    }
  } else if (warp == 19) {
    named_bar_sync(1, AT2_THREADS);
  } else {
    // ===================== softmax / epilogue: tile tt = warp / 8, 16 rows per warp ==========
    reg_alloc<AT2_REGS_SOFTMAX>();
    const int tt = warp >> 3;
    const int w8 = warp & 7;
    const int lane_base = (w8 & 3) * 32 + (w8 >> 2) * 16;
    const uint32_t lane_off = static_cast<uint32_t>(lane_base) << 16;
    const int rA = lane_base + (lane >> 2);
    const int cq = (lane & 3) * 2;
    const uint32_t t_s = tmem_base + lane_off + COL_S + tt * 128;
    const uint32_t t_p = tmem_base + lane_off + COL_P + tt * 64;
    const uint32_t t_o = tmem_base + lane_off + COL_O + tt * 64;
    const float sl2 = p.scale_log2;
    uint32_t cb = 0, nu = 0;  // blocks / units processed by this tile

    // 32 logits per thread (one 64-column half: pairs i even -> row A, odd -> row B) -> 16 packed bf16x2 words
    auto exp_half = [&](const uint32_t (&v)[32], uint32_t (&pk)[16], float2& lA, float2& lB, float nmA, float nmB,
                        auto safe_tag, auto mask_tag) {
      constexpr bool SAFE = decltype(safe_tag)::value;
      constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 x = make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        if constexpr (SAFE) x = ffma2_bcast(x.x, x.y, sl2, (i & 1) ? nmB : nmA);
        else if constexpr (!SCALE1) x = ffma2(x, make_float2(sl2, sl2), make_float2(0.f, 0.f));
        float2 a;
        if (!SAFE && !MASK && ((AT2_POLY_MASK(DV) >> i) & 1)) {
          a = exp2_poly2_c2(x);
        } else {
          a.x = fast_exp2(x.x);
          a.y = fast_exp2(x.y);
        }
        if (i & 1) lB = fadd2(lB, a);
        else lA = fadd2(lA, a);
        pk[i] = pack_bf16x2(a.x, a.y);
      }
    };
    // columns >= valid of a 64-column half -> -inf
    auto mask_half = [&](uint32_t (&v)[32], int valid) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (8 * k + cq + e >= valid) {
            v[4 * k + e] = 0xff800000u;
            v[4 * k + 2 + e] = 0xff800000u;
          }
        }
      }
    };

    auto unit_fn = [&](const int u, const int idx, auto safe_tag) {
      constexpr bool SAFE = decltype(safe_tag)::value;
      const Unit t = decode_unit(u, p);
      if (!t.valid(tt)) return;
      const int nkb = t.nkb;
      const int tail_valid = t.kv_end - t.kv_begin - (nkb - 1) * AT2_BKV;  // valid columns of the last block (1..128)
      float mA = -INFINITY, mB = -INFINITY;
      float2 lA = make_float2(0.f, 0.f), lB = make_float2(0.f, 0.f);

      auto block = [&](const int j, auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        uint32_t v[32], pk[16];
        mbar_wait(s_full + tt, cb & 1);
        tc_fence_after();
        float nmA = 0.f, nmB = 0.f;
        if constexpr (SAFE) {
          // first sweep: row maxima of the whole 128-key block (S stays in TMEM until s_free)
          float xa = -INFINITY, xb = -INFINITY;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            tmem_ld_16x256b_x8(t_s + hf * 64, v);
            tmem_ld_wait32(v);
            if constexpr (MASK) mask_half(v, tail_valid - hf * 64);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              xa = fmaxf(xa, fmaxf(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1])));
              xb = fmaxf(xb, fmaxf(__uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3])));
            }
          }
          xa = fmaxf(xa, __shfl_xor_sync(0xffffffffu, xa, 1));
          xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, 1));
          xa = fmaxf(xa, __shfl_xor_sync(0xffffffffu, xa, 2));
          xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, 2));
          const float nA = fmaxf(mA, xa * sl2), nB = fmaxf(mB, xb * sl2);
          const float alA = fast_exp2(mA - nA), alB = fast_exp2(mB - nB);  // 1 when unchanged, 0 when m was -inf
          // PV of the previous block must be complete before O is rescaled AND before P is overwritten below
          mbar_wait(p_free + tt, (cb & 1) ^ 1);
          if (j > 0 && __any_sync(0xffffffffu, (nA != mA) || (nB != mB))) {
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < DV / 16; ++c) {
              uint32_t o[8];
              tmem_ld_16x256b_x2(t_o + c * 16, o);
              tmem_ld_wait8(o);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * ((i & 2) ? alB : alA));
              tmem_st_16x256b_x2(t_o + c * 16, o);
            }
          }
          lA.x *= alA;
          lA.y *= alA;
          lB.x *= alB;
          lB.y *= alB;
          mA = nA;
          mB = nB;
          nmA = -mA;
          nmB = -mB;
        }
        if constexpr (SAFE) {
          // ---- first half: columns [0, 64) ----
          tmem_ld_16x256b_x8(t_s, v);
          tmem_ld_wait32(v);
          if constexpr (MASK) mask_half(v, tail_valid);
          exp_half(v, pk, lA, lB, nmA, nmB, safe_tag, mask_tag);
          tmem_st_16x128b_x8(t_p, pk);
          // ---- second half: columns [64, 128); S_t is free once it is in registers ----
          tmem_ld_16x256b_x8(t_s + 64, v);
          tmem_ld_wait32(v);
          tc_fence_before();
          __syncwarp();
          if (elect_one_sync()) mbar_arrive(s_free + tt);
          if constexpr (MASK) mask_half(v, tail_valid - 64);
          exp_half(v, pk, lA, lB, nmA, nmB, safe_tag, mask_tag);
          tmem_st_16x128b_x8(t_p + 32, pk);
        } else {
          // Both halves of S_t go to registers first and S_t is released at once: QK^T of the next block then runs
          // under ALL of this block's exponentials, and no tcgen05.ld latency sits between the two halves.
          uint32_t v1[32];
          tmem_ld_16x256b_x8(t_s, v);
          tmem_ld_16x256b_x8(t_s + 64, v1);
          tmem_ld_wait32(v);
          tmem_ld_wait32(v1);
          tc_fence_before();
          __syncwarp();
          if (elect_one_sync()) mbar_arrive(s_free + tt);
          if constexpr (MASK) mask_half(v, tail_valid);
          exp_half(v, pk, lA, lB, nmA, nmB, safe_tag, mask_tag);
          mbar_wait(p_free + tt, (cb & 1) ^ 1);  // PV of the previous block has read P_t
          tmem_st_16x128b_x8(t_p, pk);
          if constexpr (MASK) mask_half(v1, tail_valid - 64);
          exp_half(v1, pk, lA, lB, nmA, nmB, safe_tag, mask_tag);
          tmem_st_16x128b_x8(t_p + 32, pk);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (elect_one_sync()) mbar_arrive(p_full + tt);
        ++cb;
      };
      for (int j = 0; j + 1 < nkb; ++j) block(j, T2No{});
      block(nkb - 1, T2Yes{});

      // ---- epilogue ----
      mbar_wait(o_full + tt, nu & 1);
      tc_fence_after();
      uint32_t o[32];
      if constexpr (DV == 64) {
        tmem_ld_16x256b_x8(t_o, o);
      } else {
        tmem_ld_16x256b_x4(t_o, o);
        tmem_ld_16x256b_x2_hi(t_o + 32, o);
#pragma unroll
        for (int i = 24; i < 32; ++i) o[i] = 0u;
      }
      tmem_ld_wait32(o);
      tc_fence_before();
      __syncwarp();
      if (elect_one_sync()) mbar_arrive(o_empty + tt);
      ++nu;
      float la = lA.x + lA.y, lb = lB.x + lB.y;
      la += __shfl_xor_sync(0xffffffffu, la, 1);
      lb += __shfl_xor_sync(0xffffffffu, lb, 1);
      la += __shfl_xor_sync(0xffffffffu, la, 2);
      lb += __shfl_xor_sync(0xffffffffu, lb, 2);
      if constexpr (!SAFE) {
        const bool bad = !(la >= AT2_L_MIN && la <= AT2_L_MAX) || !(lb >= AT2_L_MIN && lb <= AT2_L_MAX);
        if (__any_sync(0xffffffffu, bad)) {
          if (lane == 0) atomicOr(&redo[idx >> 5], 1u << (idx & 31));
          return;  // pass 2 redoes the whole unit (both tiles)
        }
      }
      const float invA = 1.0f / la, invB = 1.0f / lb;
      const int rowA = t.q0(tt) + rA, rowB = rowA + 8;
      const long long base = static_cast<long long>(t.split) * p.o_split_stride +
                             static_cast<long long>(t.b) * p.o_batch_stride + static_cast<long long>(t.h(tt)) * DV + cq;
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
        float* lse = p.lse + static_cast<long long>(t.split) * p.lse_split_stride +
                     (static_cast<long long>(t.b) * p.heads + t.h(tt)) * p.Lq;
        const float m0A = SAFE ? mA : 0.f, m0B = SAFE ? mB : 0.f;
        if (rowA < p.Lq) lse[rowA] = (m0A + log2f(la)) * 0.6931471805599453f;
        if (rowB < p.Lq) lse[rowB] = (m0B + log2f(lb)) * 0.6931471805599453f;
      }
    };

    if (!p.all_safe) {
      int idx = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++idx) unit_fn(u, idx, T2No{});
    }
    named_bar_sync(1, AT2_THREADS);
    {
      int idx = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++idx)
        if (selected(1, idx)) unit_fn(u, idx, T2Yes{});
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 19) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attn_optimistic_enabled();  // xs_attn_tc.cu

template <int DQK, int DV, bool SCALE1>
static int launch_attn2(dim3 grid, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                        const Attn2Params& p, cudaStream_t stream) {
  auto kern = attn_pair_kernel<DQK, DV, SCALE1>;
  XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT2_SMEM_BYTES));
  kern<<<grid, AT2_THREADS, AT2_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  XS_LAUNCH_CHECK();
  return 0;
}

// Same contract as flash_attn_bf16_tc.  Returns 1 when the shape is better served by the 64-key kernel (nothing launched).
int flash_attn_bf16_pair(const void* q, const void* k, const void* v, void* o, float* lse, int B, int heads, int Lq,
                         int Lk, int head_dim, long long q_row_stride, long long q_batch_stride,
                         long long kv_row_stride, long long kv_batch_stride, int kv_shared, int nsplit, int o_is_f32,
                         float scale, cudaStream_t stream) {
  if (!(head_dim == 64 || head_dim == 48) || B <= 0 || heads <= 0 || Lq <= 0 || Lk <= 0 || nsplit <= 0 || !(scale > 0.f))
    return 1;
  if (nsplit > 1 && !(o_is_f32 && lse != nullptr)) return 1;
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
  Attn2Params p;
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
  if ((long long)(nsplit - 1) * p.split_len >= Lk) return 1;
  p.o_row_stride = (long long)heads * head_dim;
  p.o_batch_stride = (long long)Lq * p.o_row_stride;
  p.o_split_stride = (long long)B * p.o_batch_stride;
  p.lse_split_stride = (long long)B * heads * Lq;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int nq_tiles = (Lq + 127) / 128;
  p.n_pairs_full = nq_tiles / 2;
  p.units_per_bs = heads * p.n_pairs_full + ((nq_tiles & 1) ? (heads + 1) / 2 : 0);
  const long long n_units = (long long)p.units_per_bs * B * nsplit;
  if (n_units >= (1ll << 31)) return 1;
  p.n_units = (int)n_units;
  const int max_ctas = num_sms();
  dim3 grid(p.n_units < max_ctas ? p.n_units : max_ctas);
  const int per_cta = (p.n_units + (int)grid.x - 1) / (int)grid.x;
  p.all_safe = (attn_optimistic_enabled() == 0 || per_cta > AT2_MAX_UNITS_PER_CTA) ? 1 : 0;
  const bool scale1 = fabsf(p.scale_log2 - 1.0f) < 1e-6f;
  if (scale1) p.scale_log2 = 1.0f;
  if (head_dim == 64)
    return scale1 ? launch_attn2<4, 64, true>(grid, tmQ, tmK, tmV, p, stream)
                  : launch_attn2<4, 64, false>(grid, tmQ, tmK, tmV, p, stream);
  return scale1 ? launch_attn2<3, 48, true>(grid, tmQ, tmK, tmV, p, stream)
                : launch_attn2<3, 48, false>(grid, tmQ, tmK, tmV, p, stream);
}

}  // namespace xs
