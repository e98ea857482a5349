// Linear layers on tcgen05 tensor cores:  out[M,N] = act(A[M,K] @ W[N,K]^T + bias[N])
// Operands are bf16 (kind::f16, K=16 per instruction) or fp32 read as TF32 (kind::tf32, K=8; used for the
// small precision-critical GEMMs: patch embedding, decoder projections/FFN and the head, ~2 % of the FLOPs);
// the output tile is stored as bf16 or fp32.
//
// One persistent, warp-specialised kernel (one CTA per SM):
//   warp 0      TMA producer   A tile [128 x 64] and W tile [BN x 64] (128B-swizzled) into a smem ring
//   warp 1      MMA issuer     one thread issues tcgen05.mma (M=128, N=BN, K=16) into a TMEM accumulator;
//                              two accumulator stages so tile i+1's mainloop overlaps tile i's epilogue
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       tcgen05.ld (thread == row; two warps share a TMEM lane quarter and split the
//                              columns) -> +bias -> activation -> bf16/fp32 -> 128B-swizzled smem staging ->
//                              TMA store (coalesced, clips the M tail)
// A second epilogue (EPI_JIGSAW) fuses the regression head's last Linear with the score activation and
// the jigsaw scatter (reference: model/cross_reference.py:45-50,82-87, model/regression_layer.py:26-62,
// utils/misc/image.py:8-21): out[b, 14r+i, 14c+j] = act(z[b, r*pw+c, 14i+j]).
//
// Reference call sites replaced: every torch.nn.Linear / Conv2d(k=s=14) on the path
// ($SP/transformers/models/dinov2/modeling_dinov2.py:139-149,199-213,246-252,317-327;
//  model/customised_transformer/transformer.py:68-75,208-210; $SP/torch/nn/functional.py:5849-5858,6692).
#include <stdlib.h>

#include "xs_common.cuh"

namespace xs {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_THREADS = 256;
constexpr int ACC_STAGE_COLS = 256;  // TMEM column offset between the two accumulator stages
constexpr int GEMM_KB_MAX = 6;       // A-stationary mode: K <= 6 * 64
constexpr int JIG_LD = 197;          // padded row length of the fp32 score staging tile

enum Epi : int { EPI_STORE = 0, EPI_JIGSAW = 1 };
enum InT : int { IN_BF16 = 0, IN_TF32 = 1 };
// ADD: out (fp32) += tile, in place; F16: fp16 output (q|k|v operands of the fp16 attention kernel)
enum OutT : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ADD = 2, OUT_F16 = 3 };
template <int OUT>
__device__ __forceinline__ uint32_t pack_out16(float lo, float hi) {
  if constexpr (OUT == OUT_F16) return pack_f16x2(lo, hi);
  else return pack_bf16x2(lo, hi);
}

struct JigsawParams {
  float* score;  // (B, 14*ph, 14*pw) fp32
  int P;         // tokens per map = ph*pw
  int pw;
  int Wout;      // 14*pw
  int HWout;     // 14*ph*14*pw
  int use_tanh;
  float power;
};

template <int BN, int STAGES, int EPI, int CTA2 = 0, int OUT = 0, int LNF = 0>
struct GemmSmem {
  // one 32-row x 32-column staging tile of an epilogue warp (64-byte rows for bf16 output, 128-byte rows for fp32)
  static constexpr uint32_t WARP_STAGE_BYTES = 32 * 32 * ((OUT == 1 || OUT == 2) ? 4 : 2);
  static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = (CTA2 ? BN / 2 : BN) * GEMM_BK * 2;  // a CTA pair splits the W tile rows
  // jigsaw epilogue: the fp32 score tile + one 64-bit output offset per token of the tile
  static constexpr uint32_t STAGING_BYTES = (EPI == EPI_STORE) ? 8 * 2 * WARP_STAGE_BYTES : GEMM_BM * JIG_LD * 4 + GEMM_BM * 8;
  static constexpr int A_SLOTS = (CTA2 == 2) ? GEMM_KB_MAX : STAGES;  // A-stationary: one slot per k-block of the m-block
  static constexpr uint32_t OFF_A = 0;
  static constexpr uint32_t OFF_B = OFF_A + A_SLOTS * A_BYTES;
  static constexpr uint32_t OFF_STAGING = OFF_B + STAGES * B_BYTES;
  static constexpr uint32_t OFF_BAR = (OFF_STAGING + STAGING_BYTES + 15u) & ~15u;
  static constexpr uint32_t BAR_BYTES = (2 * STAGES + 4 + 2 * GEMM_KB_MAX + 16) * 8 + 16;  // +16: residual-tile barriers
  static constexpr uint32_t OFF_LNX = (OFF_BAR + BAR_BYTES + 15u) & ~15u;  // LNF: [2 row blocks][8 warps][32 rows] float2
  static constexpr uint32_t LNX_BYTES = LNF ? 2 * 8 * 32 * 8 : 0;
  static constexpr uint32_t TOTAL = OFF_LNX + LNX_BYTES + 1024;  // +1024: manual 1 KB alignment of the base
  static_assert(TOTAL <= 232448, "shared memory layout exceeds 227 KB");
};

// CTA2 = 1: CTA pairs (cluster of 2, tcgen05 cta_group::2).  One MMA covers a 256 x BN tile: CTA r of the pair
// stages rows [128 r, 128 r + 128) of A and rows [BN/2 r, BN/2 r + BN/2) of the W tile in ITS shared memory and
// holds its 128 accumulator rows in ITS TMEM; the leader CTA (rank 0) issues the MMAs, whose completion is
// multicast to both CTAs' barriers.  Shared-memory fill traffic per output element falls by a third, which is
// what bounds these K = 384 GEMMs (L2 -> SM bandwidth), see DESIGN.md.
// CTA2 = 2: CTA pairs with a STATIONARY A block (K <= 384).  A pair owns 256 rows of A at a time: its six k-blocks
// stay in shared memory while the pair walks ALL n-tiles of that row block, so only W (half a tile per CTA) is
// streamed.  Per 128 x BN output tile a CTA then pulls BN/2 x 384 x 2 bytes from L2 (74 KB at BN = 192) instead
// of 240 KB, which moves the K = 384 GEMMs from the L2 -> SM bandwidth bound to the MMA / epilogue bound.
// LNF = 1 (OUT_F32_ADD, N = 2 n-tiles): the LayerNorm that follows the residual add is fused into the epilogue.  A CTA
// walks BOTH n-tiles of its 128 rows (row-block-major order in every operand mode), so after the second tile's
// residual update it knows the statistics of the full 384-wide rows; it then re-reads the rows it has just stored
// (its own TMA stores, L2 hits) tile by tile, normalises and emits y = LN(h) in bf16 through a second tensor map.
// The LayerNorm kernel launch and its HBM read of h disappear (modeling_dinov2.py:367-386: norm2 after the attention
// residual, the next layer's norm1 after the MLP residual).
// LNF = 2 (same shapes): the epilogue only MEASURES the rows -- it emits a bf16 copy of the updated residual stream and
// (mean, rstd) per row; the LayerNorm itself is folded into the GEMM that consumes it (ACT_LN_*: gamma folded into the
// weights, rstd * (acc - mean * c1[n]) + c0[n] in that GEMM's epilogue), so no second pass and no LayerNorm kernel.
struct LnParams {
  const float* gamma;   // LNF = 1: LayerNorm weight.  ACT_LN_* consumer: c1[n] = sum_k bf16(gamma[k] W[n,k])
  const float* beta;
  float eps;
  const float* h;       // the residual stream the epilogue has just updated (row pitch ldh floats)
  __nv_bfloat16* y;     // LNF = 1: LayerNorm output; LNF = 2: bf16 copy of h (row pitch ldy elements)
  int ldh, ldy;
  float2* stats;        // LNF = 2: (mean, rstd) per row, written; ACT_LN_* consumer: read
};

template <int BN, int STAGES, int EPI, int ACT, int IN, int OUT, int CTA2, int LNF = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmY,
               const float* __restrict__ bias, int M, int N, int K, JigsawParams jp, LnParams lnp) {
  using L = GemmSmem<BN, STAGES, EPI, CTA2, OUT, LNF>;
  static_assert(!LNF || (OUT == OUT_F32_ADD && EPI == EPI_STORE && BN == 192), "LN fusion: residual epilogue only");
  static_assert(!CTA2 || EPI == EPI_STORE, "CTA pairs: store epilogue only");
  static_assert(CTA2 != 2 || IN == IN_BF16, "A-stationary pairs: bf16 operands (K <= 384 must fit six 64-element k-blocks)");
  constexpr int NC = CTA2 ? 2 : 1;                       // CTAs per tile
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;   // 0 = leader
  const int worker = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;
  const int n_workers = CTA2 ? (gridDim.x >> 1) : gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint8_t* smA = smem + L::OFF_A;
  uint8_t* smB = smem + L::OFF_B;
  uint8_t* staging = smem + L::OFF_STAGING;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* a_full = tmem_empty + 2;            // [GEMM_KB_MAX] A-stationary mode only
  uint64_t* a_empty = a_full + GEMM_KB_MAX;     // [GEMM_KB_MAX]
  uint64_t* res_full = a_empty + GEMM_KB_MAX;   // [8 epilogue warps][2] OUT_F32_ADD only: residual tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 16);
  constexpr bool ASTAT = CTA2 == 2;
  constexpr bool ROWMAJOR = ASTAT || LNF;  // a worker walks whole row blocks (all n-tiles) instead of single tiles

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + NC * GEMM_BM - 1) / (NC * GEMM_BM);
  const int num_n = N / BN;
  constexpr int BKE = (IN == IN_TF32) ? 32 : 64;  // elements per 128-byte smem row
  const int num_k = (K + BKE - 1) / BKE;
  const int num_tiles = num_m * num_n;
  // tile sequence of this worker: streaming modes walk tiles worker, worker + n_workers, ... (n fastest);
  // the A-stationary mode walks row blocks worker, worker + n_workers, ... and all n-tiles inside each
  const int my_rounds = worker < num_m ? (num_m - worker + n_workers - 1) / n_workers : 0;
  const int my_count = ROWMAJOR ? my_rounds * num_n : (worker < num_tiles ? (num_tiles - worker + n_workers - 1) / n_workers : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if constexpr (EPI == EPI_STORE) tma_prefetch_desc(&tmC);
    if constexpr (LNF == 1) tma_prefetch_desc(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8 * NC);  // one elected lane per epilogue warp (of both CTAs of a pair)
    }
    for (int s = 0; s < GEMM_KB_MAX; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 16; ++s) mbar_init(&res_full[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // barrier inits and TMEM of both CTAs are in place
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    uint32_t stage = 0, phase = 0;
    if constexpr (ASTAT) {
      for (int r = 0; r < my_rounds; ++r) {
        const int m_blk = worker + r * n_workers;
        for (int n_blk = 0; n_blk < num_n; ++n_blk) {
          for (int kb = 0; kb < num_k; ++kb) {
            if (n_blk == 0) {
              // slot kb was last read by the previous row block's last n-tile; A and the first W tile are issued
              // k-block by k-block so neither waits behind the other
              mbar_wait(&a_empty[kb], (r & 1) ^ 1);
              if (elect_one_sync()) {
                const uint32_t a_leader = map_to_cta(smem_u32(&a_full[kb]), 0);
                if (rank == 0) mbar_expect_tx(&a_full[kb], 2 * L::A_BYTES);
                tma_load_2d_pair(smA + kb * L::A_BYTES, &tmA, a_leader, kb * BKE,
                                 (m_blk * 2 + static_cast<int>(rank)) * GEMM_BM);
              }
              __syncwarp();
            }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one_sync()) {
              const uint32_t full_leader = map_to_cta(smem_u32(&full_bar[stage]), 0);
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::B_BYTES);
              tma_load_2d_pair(smB + stage * L::B_BYTES, &tmW, full_leader, kb * BKE,
                               n_blk * BN + static_cast<int>(rank) * (BN / 2));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else
    for (int t = 0; t < my_count; ++t) {
      const int tile = ROWMAJOR ? 0 : worker + t * n_workers;
      const int m_blk = ROWMAJOR ? worker + (t / num_n) * n_workers : tile / num_n;
      const int n_blk = ROWMAJOR ? t % num_n : tile % num_n;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if constexpr (CTA2) {
            // both CTAs' bytes are credited to the LEADER's full barrier (its expect_tx covers the pair)
            const uint32_t full_leader = map_to_cta(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (L::A_BYTES + L::B_BYTES));
            tma_load_2d_pair(smA + stage * L::A_BYTES, &tmA, full_leader, kb * BKE,
                             (m_blk * 2 + static_cast<int>(rank)) * GEMM_BM);
            tma_load_2d_pair(smB + stage * L::B_BYTES, &tmW, full_leader, kb * BKE,
                             n_blk * BN + static_cast<int>(rank) * (BN / 2));
          } else {
            mbar_expect_tx(&full_bar[stage], L::A_BYTES + L::B_BYTES);
            tma_load_2d(smA + stage * L::A_BYTES, &tmA, &full_bar[stage], kb * BKE, m_blk * GEMM_BM);
            tma_load_2d(smB + stage * L::B_BYTES, &tmW, &full_bar[stage], kb * BKE, n_blk * BN);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (converged warp, uniform operands, one elected lane issues) =========
    constexpr uint32_t idesc =
        (IN == IN_TF32) ? umma_idesc_tf32(NC * GEMM_BM, BN) : umma_idesc_bf16(NC * GEMM_BM, BN, 0, 0);
    const uint32_t tb = warp_uniform(tmem_base);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smA), 16);
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(smB), 16);
    uint32_t stage = 0, phase = 0, acc_stage = 0, acc_phase = 0;
    if constexpr (ASTAT) {
      for (int r = 0; r < my_rounds; ++r) {
        for (int n_blk = 0; n_blk < num_n; ++n_blk) {
          mbar_wait(&tmem_empty[acc_stage], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tb + acc_stage * ACC_STAGE_COLS;
          for (int kb = 0; kb < num_k; ++kb) {
            if (n_blk == 0) mbar_wait(&a_full[kb], r & 1);  // this row block's k-block kb has landed in both CTAs
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t a_lo = a_lo0 + kb * (L::A_BYTES >> 4);
              const uint32_t b_lo = b_lo0 + stage * (L::B_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss_lh_pair(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              tc_commit_pair(&empty_bar[stage]);
              if (n_blk == num_n - 1) tc_commit_pair(&a_empty[kb]);  // last reader of A slot kb in this row block
              if (kb == num_k - 1) tc_commit_pair(&tmem_full[acc_stage]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          acc_stage ^= 1;
          if (acc_stage == 0) acc_phase ^= 1;
        }
      }
    } else
    for (int t = 0; t < my_count; ++t) {
      mbar_wait(&tmem_empty[acc_stage], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tb + acc_stage * ACC_STAGE_COLS;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + stage * (L::A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (L::B_BYTES >> 4);
          if constexpr (CTA2) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss_lh_pair<IN == IN_TF32>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit_pair(&empty_bar[stage]);  // frees the slot in BOTH CTAs
            if (kb == num_k - 1) tc_commit_pair(&tmem_full[acc_stage]);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 4 x 32 bytes of K per 128-byte row (16 bf16 or 8 tf32 each)
              umma_ss_lh<IN == IN_TF32>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
            if (kb == num_k - 1) tc_commit(&tmem_full[acc_stage]);  // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int epi_tid = threadIdx.x - 128;  // 0..255
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;       // which half of each column chunk this warp handles
    const int row = q * 32 + lane;          // row of the tile owned by this thread
    uint32_t acc_stage = 0, acc_phase = 0;
    uint32_t chunk_counter = 0;
    // (row block, n-tile) of this worker's t-th tile
    auto tile_coords = [&](int t, int& m_blk_o, int& n_blk_o) {
      int m_tile;
      if constexpr (ROWMAJOR) {
        m_tile = worker + (t / num_n) * n_workers;
        n_blk_o = t % num_n;
      } else {
        const int tile = worker + t * n_workers;
        m_tile = tile / num_n;
        n_blk_o = tile % num_n;
      }
      m_blk_o = m_tile * NC + static_cast<int>(rank);
    };
    // OUT_F32_ADD (out += tile): every epilogue warp fetches the 32 x 32 fp32 tile of the residual stream its next
    // step will update with a TMA load into the step's staging buffer, ONE STEP AHEAD (across tile boundaries: the
    // first load of a tile is in flight during that tile's mainloop), adds in shared memory and stores the tile
    // back.  (A TMA reduce-store (UTMAREDG.ADD) does the same add at the L2 without the load, but its throughput
    // capped these GEMMs: proj 0.083 -> 0.174 ms, fc2 0.230 -> 0.305 ms, measured.)
    uint64_t* my_res_full = res_full + (warp - 4) * 2;
    auto issue_res_load = [&](uint32_t step_idx) {  // one lane
      if constexpr (OUT == OUT_F32_ADD && EPI == EPI_STORE) {
        constexpr int MY_STEPS_ = BN / 64;
        int t, sidx;
        if constexpr (LNF == 1) {
          // per row block: pass 1 = the residual tiles of its num_n * MY_STEPS_ steps, pass 2 = the same tiles again
          // (now holding h + delta, written by this warp's own TMA stores)
          const int spr = num_n * MY_STEPS_;
          const int rb = static_cast<int>(step_idx) / (2 * spr), s2 = static_cast<int>(step_idx) % spr;
          t = rb * num_n + s2 / MY_STEPS_;
          sidx = s2 % MY_STEPS_;
        } else {
          t = static_cast<int>(step_idx / MY_STEPS_);
          sidx = static_cast<int>(step_idx % MY_STEPS_);
        }
        if (t >= my_count) return;
        int mb, nb;
        tile_coords(t, mb, nb);
        const uint32_t b = step_idx & 1;
        uint8_t* dst = staging + (warp - 4) * (2 * L::WARP_STAGE_BYTES) + b * L::WARP_STAGE_BYTES;
        mbar_expect_tx(&my_res_full[b], L::WARP_STAGE_BYTES);
        tma_load_2d(dst, &tmC, &my_res_full[b], nb * BN + (2 * sidx + half) * 32, mb * GEMM_BM + q * 32);
      }
    };
    if constexpr (OUT == OUT_F32_ADD && EPI == EPI_STORE) {
      if (lane == 0) issue_res_load(0);
    }
    // LNF: shifted sums of this thread's half of the row (192 of 384 columns) over the row block's n-tiles
    float ln_c = 0.f, ln_s1 = 0.f, ln_s2 = 0.f;
    float2* ln_x = reinterpret_cast<float2*>(smem + L::OFF_LNX);
    for (int i = 0; i < my_count; ++i) {
      int m_blk, n_blk;
      tile_coords(i, m_blk, n_blk);
      const int row_block = i / num_n;  // (the step loop below reuses the name i)
      (void)row_block;
      mbar_wait(&tmem_full[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc_stage * ACC_STAGE_COLS + (static_cast<uint32_t>(q * 32) << 16);

      if constexpr (EPI == EPI_STORE) {
        // Every epilogue warp works on its own: 32 accumulator columns per step (steps alternate between the two
        // warps of a lane quarter), own double-buffered 32-row staging tile, own TMA stores.  No block-level
        // barrier; the tcgen05.ld of the next step is in flight while the current one is converted and stored.
        constexpr int NSTEP = BN / 32;
        constexpr int MY_STEPS = NSTEP / 2;
        static_assert(NSTEP % 2 == 0, "BN must be a multiple of 64");
        uint8_t* my_staging = staging + (warp - 4) * (2 * L::WARP_STAGE_BYTES);
        const int m0 = m_blk * GEMM_BM + q * 32;
        [[maybe_unused]] float ln_rstd = 0.f, ln_rm = 0.f;
        if constexpr (ACT == ACT_LN_NONE || ACT == ACT_LN_GELU) {  // thread == row: its (mean, rstd), once per tile
          if (m0 + lane < M) {
            const float2 st = __ldg(lnp.stats + m0 + lane);
            ln_rstd = st.y;
            ln_rm = -st.y * st.x;
          }
        }
        uint32_t v[2][32];
        tmem_ld32(taddr0 + half * 32, v[0]);
#pragma unroll
        for (int i = 0; i < MY_STEPS; ++i) {
          const int c = 2 * i + half;  // this warp's i-th 32-column step
          tmem_ld_wait32(v[i & 1]);
          if (i + 1 < MY_STEPS) {
            tmem_ld32(taddr0 + (c + 2) * 32, v[(i + 1) & 1]);
          } else {  // accumulator fully read by this warp: hand the TMEM stage back to the (leader's) MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CTA2) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc_stage]), 0));
              else mbar_arrive(&tmem_empty[acc_stage]);
            }
          }
          const uint32_t buf = chunk_counter & 1;
          if constexpr (OUT == OUT_F32_ADD) {
            // the other buffer was last read by the previous step's store: once that has drained, fetch the next
            // step's residual tile into it; then wait for this step's tile (requested one step ago)
            if (lane == 0) {
              tma_store_wait_read<0>();
              if constexpr (LNF == 1) {
                // the next load may be a pass-2 load (last pass-1 step of the row block): the tile it re-reads was
                // stored 2 * num_n * MY_STEPS - 1 steps ago; at most 4 younger store groups may still be pending
                if (n_blk == num_n - 1 && i == MY_STEPS - 1) tma_store_wait_all<4>();
              }
              issue_res_load(chunk_counter + 1);
            }
            mbar_wait(&my_res_full[buf], (chunk_counter >> 1) & 1);
          } else {
            // the TMA store that last read this staging buffer (two steps ago) must have drained
            if (lane == 0) tma_store_wait_read<1>();
          }
          __syncwarp();
          ++chunk_counter;
          const int n0 = n_blk * BN + c * 32;
          uint8_t* srow = my_staging + buf * L::WARP_STAGE_BYTES;
          const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
          const uint32_t(&vv)[32] = v[i & 1];
          if constexpr (OUT == OUT_BF16 || OUT == OUT_F16) {
            // 64-byte rows, 64B swizzle: 16-byte chunk index XOR ((row >> 1) & 3); conflict-free for thread == row
            uint8_t* rp = srow + lane * 64;
            [[maybe_unused]] const float4* c14 = reinterpret_cast<const float4*>(lnp.gamma + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 ba = __ldg(b4 + j * 2), bb = __ldg(b4 + j * 2 + 1);
              const int o = j * 8;
              uint4 pk;
              float2 r0, r1, r2, r3;
              if constexpr (ACT == ACT_LN_NONE || ACT == ACT_LN_GELU) {
                const float4 ca = __ldg(c14 + j * 2), cb = __ldg(c14 + j * 2 + 1);
                r0 = lnfold_act2<ACT>(__uint_as_float(vv[o + 0]), __uint_as_float(vv[o + 1]), ln_rstd, ln_rm, ca.x, ca.y, ba.x, ba.y);
                r1 = lnfold_act2<ACT>(__uint_as_float(vv[o + 2]), __uint_as_float(vv[o + 3]), ln_rstd, ln_rm, ca.z, ca.w, ba.z, ba.w);
                r2 = lnfold_act2<ACT>(__uint_as_float(vv[o + 4]), __uint_as_float(vv[o + 5]), ln_rstd, ln_rm, cb.x, cb.y, bb.x, bb.y);
                r3 = lnfold_act2<ACT>(__uint_as_float(vv[o + 6]), __uint_as_float(vv[o + 7]), ln_rstd, ln_rm, cb.z, cb.w, bb.z, bb.w);
              } else {
                r0 = bias_act2<ACT>(__uint_as_float(vv[o + 0]), __uint_as_float(vv[o + 1]), ba.x, ba.y);
                r1 = bias_act2<ACT>(__uint_as_float(vv[o + 2]), __uint_as_float(vv[o + 3]), ba.z, ba.w);
                r2 = bias_act2<ACT>(__uint_as_float(vv[o + 4]), __uint_as_float(vv[o + 5]), bb.x, bb.y);
                r3 = bias_act2<ACT>(__uint_as_float(vv[o + 6]), __uint_as_float(vv[o + 7]), bb.z, bb.w);
              }
              pk.x = pack_out16<OUT>(r0.x, r0.y);
              pk.y = pack_out16<OUT>(r1.x, r1.y);
              pk.z = pack_out16<OUT>(r2.x, r2.y);
              pk.w = pack_out16<OUT>(r3.x, r3.y);
              *reinterpret_cast<uint4*>(rp + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
            }
          } else {
            // 128-byte rows, 128B swizzle: 16-byte chunk index XOR (row & 7)
            uint8_t* rp = srow + lane * 128;
            [[maybe_unused]] uint32_t hb_lo0 = 0u, hb_lo1 = 0u;
            [[maybe_unused]] uint4* hb_row = nullptr;
            if constexpr (LNF == 2)
              hb_row = reinterpret_cast<uint4*>(lnp.y + static_cast<size_t>(m0 + lane) * lnp.ldy + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 ba = __ldg(b4 + j);
              float4 o4;
              o4.x = apply_act<ACT>(__uint_as_float(vv[j * 4 + 0]) + ba.x);
              o4.y = apply_act<ACT>(__uint_as_float(vv[j * 4 + 1]) + ba.y);
              o4.z = apply_act<ACT>(__uint_as_float(vv[j * 4 + 2]) + ba.z);
              o4.w = apply_act<ACT>(__uint_as_float(vv[j * 4 + 3]) + ba.w);
              float4* sp = reinterpret_cast<float4*>(rp + ((j ^ (lane & 7)) << 4));
              if constexpr (OUT == OUT_F32_ADD) {  // the residual tile sits in the same swizzled layout
                const float4 r4 = *sp;
                o4.x += r4.x; o4.y += r4.y; o4.z += r4.z; o4.w += r4.w;
              }
              if constexpr (LNF) {
                if (n_blk == 0 && i == 0 && j == 0) {  // first value of the row block: the shift of the running sums
                  ln_c = o4.x;
                  ln_s1 = 0.f;
                  ln_s2 = 0.f;
                }
                const float d0 = o4.x - ln_c, d1 = o4.y - ln_c, d2 = o4.z - ln_c, d3 = o4.w - ln_c;
                ln_s1 += (d0 + d1) + (d2 + d3);
                ln_s2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
              }
              if constexpr (LNF == 2) {
                // bf16 copy of the updated rows, straight from registers: 64 contiguous bytes per thread and step
                if (j & 1) {
                  if (m0 + lane < M)
                    hb_row[j >> 1] = make_uint4(hb_lo0, hb_lo1, pack_bf16x2(o4.x, o4.y), pack_bf16x2(o4.z, o4.w));
                } else {
                  hb_lo0 = pack_bf16x2(o4.x, o4.y);
                  hb_lo1 = pack_bf16x2(o4.z, o4.w);
                }
              }
              *sp = o4;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // box = 32 columns x 32 rows; clips the M tail
            tma_store_2d(&tmC, srow, n0, m0);
            tma_store_commit();
          }
        }
        if constexpr (LNF) {
          if (n_blk == num_n - 1) {
            // ---- statistics of the full rows: this warp saw COLS_HALF columns of each of its 32 rows, the other warp
            //      of the lane quarter the rest.  Chan's parallel update of (mean, M2) merges the two halves. ----
            const float nh = static_cast<float>(num_n * (BN / 2));
            const float mean_t = ln_c + ln_s1 / nh;
            const float m2_t = ln_s2 - ln_s1 * ln_s1 / nh;
            const int rbp = row_block & 1;
            ln_x[(rbp * 8 + (warp - 4)) * 32 + lane] = make_float2(mean_t, m2_t);
            named_bar_sync(2 + q, 64);  // the two warps of this lane quarter
            const float2 other = ln_x[(rbp * 8 + ((warp - 4) ^ 4)) * 32 + lane];
            const float dm = other.x - mean_t;
            const float mean = mean_t + 0.5f * dm;
            const float var = (m2_t + other.y + dm * dm * (0.5f * nh)) / (2.0f * nh);
            const float rstd = rsqrtf(var + lnp.eps);
            if constexpr (LNF == 2) {
              if (half == 0 && m0 + lane < M) lnp.stats[m0 + lane] = make_float2(mean, rstd);
            }
            // ---- pass 2: re-read the updated tiles (this warp's own TMA stores: L2 hits) through the same
            //      double-buffered TMA pipeline, normalise, store y (bf16) through the second tensor map.
            //      (A variant with plain per-thread vector loads instead of TMA was 2x slower: thread == row makes every
            //      lane touch its own cache line.  profiles/r2_experiments.txt) ----
#pragma unroll 1
            for (int s2 = 0; LNF == 1 && s2 < num_n * MY_STEPS; ++s2) {
              const int nb2 = s2 / MY_STEPS, c2 = 2 * (s2 % MY_STEPS) + half;
              const uint32_t buf = chunk_counter & 1;
              if (lane == 0) {
                tma_store_wait_read<0>();
                // next load: another pass-2 tile (its store has at most 4 younger groups) or the next row block's
                // first residual tile (no dependence)
                if (s2 + 1 < num_n * MY_STEPS) tma_store_wait_all<4>();
                issue_res_load(chunk_counter + 1);
              }
              mbar_wait(&my_res_full[buf], (chunk_counter >> 1) & 1);
              __syncwarp();
              ++chunk_counter;
              const int n0 = nb2 * BN + c2 * 32;
              uint8_t* srow = my_staging + buf * L::WARP_STAGE_BYTES;
              const float4* g4 = reinterpret_cast<const float4*>(lnp.gamma + n0);
              const float4* b4l = reinterpret_cast<const float4*>(lnp.beta + n0);
              uint32_t pk[16];
              const uint8_t* rp = srow + lane * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 x4 = *reinterpret_cast<const float4*>(rp + ((j ^ (lane & 7)) << 4));
                const float4 g = __ldg(g4 + j), bb = __ldg(b4l + j);
                pk[2 * j] = pack_bf16x2((x4.x - mean) * rstd * g.x + bb.x, (x4.y - mean) * rstd * g.y + bb.y);
                pk[2 * j + 1] = pack_bf16x2((x4.z - mean) * rstd * g.z + bb.z, (x4.w - mean) * rstd * g.w + bb.w);
              }
              __syncwarp();  // every lane has read its fp32 row before the bf16 rows overwrite the buffer
              uint8_t* wp = srow + lane * 64;  // 64-byte rows, 64B swizzle (the layout of the bf16 output tensor map)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(wp + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmY, srow, n0, m0);
                tma_store_commit();
              }
            }
          }
        }
      } else {
        // ---- regression head: activation + jigsaw scatter (fp32 score map) ----
        float* stile = reinterpret_cast<float*>(staging);
        constexpr int NCHUNK = BN / 32;
        constexpr int LAST = ((NCHUNK - 1) & 1);  // parity of the last chunk
#pragma unroll 1
        for (int c = half; c < NCHUNK; c += 2) {
          uint32_t v[32];
          tmem_ld32(taddr0 + c * 32, v);
          tc_wait_ld();
          if (c + 2 >= NCHUNK) {  // this warp's last chunk
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc_stage]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int k = c * 32 + j;
            if (k < 196) {
              const float z = __uint_as_float(v[j]) + __ldg(bias + k);
              float s = jp.use_tanh ? tanhf(z) : __fdividef(1.0f, 1.0f + __expf(-z));
              if (jp.power != 1.0f) s = powf(s, jp.power);
              stile[row * JIG_LD + k] = s;
            }
          }
        }
        (void)LAST;
        // where each token's 14 x 14 patch starts in the score map (utils/misc/image.py:8-21): the divisions once per token
        const int m0 = m_blk * GEMM_BM;
        const int rows_valid = min(GEMM_BM, M - m0);
        long long* tok_off = reinterpret_cast<long long*>(stile + GEMM_BM * JIG_LD);
        if (epi_tid < rows_valid) {
          const int t = m0 + epi_tid;
          const int b = t / jp.P;
          const int p = t - b * jp.P;
          const int r = p / jp.pw;
          const int cc = p - r * jp.pw;
          tok_off[epi_tid] = static_cast<long long>(b) * jp.HWout + static_cast<long long>(14 * r) * jp.Wout + 14 * cc;
        }
        named_bar_sync(1, EPI_THREADS);
        // one (patch row i, token) pair per thread and step: 14 contiguous floats as seven 8-byte stores (every offset is
        // even: Wout = 14 pw).  Consecutive threads take consecutive tokens, i.e. consecutive 56-byte pieces of an image row.
        for (int e = epi_tid; e < 14 * GEMM_BM; e += EPI_THREADS) {
          const int i = e / GEMM_BM;
          const int tt = e - i * GEMM_BM;
          if (tt < rows_valid) {
            const float* src = stile + tt * JIG_LD + i * 14;
            float2* dst = reinterpret_cast<float2*>(jp.score + tok_off[tt] + static_cast<long long>(i) * jp.Wout);
#pragma unroll
            for (int j = 0; j < 7; ++j) dst[j] = make_float2(src[2 * j], src[2 * j + 1]);
          }
        }
        named_bar_sync(1, EPI_THREADS);  // staging tile is reused by the next tile
      }
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
    if constexpr (EPI == EPI_STORE) {
      if (lane == 0) tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // the peer's smem / TMEM / barriers stay alive until both CTAs are done
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES, int EPI, int ACT, int IN, int OUT, int CTA2 = 0, int LNF = 0>
static int launch_gemm(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, int M,
                       int N, int K, JigsawParams jp, cudaStream_t stream, void* y = nullptr, int ldy = 0,
                       LnParams lnp = LnParams{nullptr, nullptr, 0.f, nullptr, nullptr, 0, 0, nullptr}) {
  using L = GemmSmem<BN, STAGES, EPI, CTA2, OUT, LNF>;
  constexpr int IN_B = (IN == IN_TF32) ? 4 : 2;
  constexpr int OUT_B = (OUT == OUT_F32 || OUT == OUT_F32_ADD) ? 4 : 2;
  constexpr uint32_t BKE = 128 / IN_B;
  CUtensorMap tmA, tmW, tmC;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda * IN_B};
    uint32_t box[2] = {BKE, GEMM_BM};
    int rc = make_tmap(&tmA, A, IN_B, 2, dims, strides, box, SWZ_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldw * IN_B};
    uint32_t box[2] = {BKE, (uint32_t)(CTA2 ? BN / 2 : BN)};
    int rc = make_tmap(&tmW, W, IN_B, 2, dims, strides, box, SWZ_128B);
    if (rc) return rc;
  }
  if (EPI == EPI_STORE) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ldc * OUT_B};
    uint32_t box[2] = {32, 32};  // one epilogue-warp step: 32 columns x 32 rows (64-byte rows for bf16, 128 for fp32)
    int rc = make_tmap(&tmC, out, OUT_B, 2, dims, strides, box, OUT_B == 2 ? SWZ_64B : SWZ_128B);
    if (rc) return rc;
  } else {
    tmC = tmA;  // unused
  }
  CUtensorMap tmY = tmC;  // unused unless LNF == 1
  if constexpr (LNF == 1) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ldy * 2};
    uint32_t box[2] = {32, 32};
    int rc = make_tmap(&tmY, y, 2, 2, dims, strides, box, SWZ_64B);
    if (rc) return rc;
  }
  auto kern = gemm_tc_kernel<BN, STAGES, EPI, ACT, IN, OUT, CTA2, LNF>;
  XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL));
  if constexpr (CTA2) {
    const int num_m2 = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    const int units = (CTA2 == 2 || LNF) ? num_m2 : num_m2 * (N / BN);  // row-block-major modes walk whole row blocks
    const int pairs = units < num_sms() / 2 ? units : num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    XS_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, tmC, tmY, bias, M, N, K, jp, lnp));
  } else {
    const int num_tiles = LNF ? (M + GEMM_BM - 1) / GEMM_BM : ((M + GEMM_BM - 1) / GEMM_BM) * (N / BN);
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmW, tmC, tmY, bias, M, N, K, jp, lnp);
  }
  XS_LAUNCH_CHECK();
  return 0;
}

#define XS_GEMM_ARGS A, lda, W, ldw, bias, out, ldc, M, N, K, jp, stream

template <int BN, int STAGES, int IN, int OUT, int CTA2 = 0>
static int dispatch_act(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, int M,
                        int N, int K, int act, cudaStream_t stream) {
  JigsawParams jp{};
  switch (act) {
    case ACT_NONE: return launch_gemm<BN, STAGES, EPI_STORE, ACT_NONE, IN, OUT, CTA2>(XS_GEMM_ARGS);
    case ACT_GELU: return launch_gemm<BN, STAGES, EPI_STORE, ACT_GELU, IN, OUT, CTA2>(XS_GEMM_ARGS);
    case ACT_RELU: return launch_gemm<BN, STAGES, EPI_STORE, ACT_RELU, IN, OUT, CTA2>(XS_GEMM_ARGS);
    case ACT_LEAKY: return launch_gemm<BN, STAGES, EPI_STORE, ACT_LEAKY, IN, OUT, CTA2>(XS_GEMM_ARGS);
  }
  set_last_error("xs_gemm_bias_act: unknown activation %d", act);
  return -1;
}

// A-stationary CTA pairs walk whole 256-row blocks: with num_m2 blocks on SMs/2 pairs the last wave may be nearly empty
// (172 blocks on 74 pairs = 2.3 waves -> 3).  Below this wave efficiency the streaming pair kernel, whose unit is one
// 256 x BN tile, is used instead.  Measured (profiles/r2_experiments.txt): at 43 840 rows (32 query images, cfg 3) the
// streaming pairs win by 19 % (q|k|v), 26 % (fc1), 12 % (N = 384); at 65 000 rows (efficiency 0.86) still by 7-12 %; at the
// headline's 263 040 rows (0.99) the A-stationary form is the faster one.
#ifndef XS_ASTAT_MIN_EFF
#define XS_ASTAT_MIN_EFF 0.92f
#endif
static bool astat_pays(int num_m2) {
  const int pairs = num_sms() / 2;
  if (num_m2 < pairs) return false;
  const int waves = (num_m2 + pairs - 1) / pairs;
  return static_cast<float>(num_m2) >= XS_ASTAT_MIN_EFF * static_cast<float>(waves * pairs);
}

// Tensor-core GEMM entry used by xs_api.cu.  in_tf32: A/W are fp32 (TF32 multiply), else bf16.
// out_f32: 0 bf16 output, 1 fp32 output, 2 fp32 output accumulated in place (out += ...), 3 fp16 output.
// N must be a multiple of 192 or 256; row pitches must be multiples of 16 bytes.
int gemm_tc(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, int M, int N, int K,
            int act, int in_tf32, int out_f32, cudaStream_t stream) {
  const int al = in_tf32 ? 4 : 8;
  const int cl = out_f32 ? 4 : 8;
  XS_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  XS_CHECK_ARG((K % al) == 0 && (lda % al) == 0 && (ldw % al) == 0 && (ldc % cl) == 0,
               "gemm: K/lda/ldw/ldc must be multiples of 16 bytes for TMA, got K=%d lda=%d ldw=%d ldc=%d", K, lda, ldw,
               ldc);
  XS_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
               "gemm: pointers must be 16-byte aligned");
  const bool use192 = (N % 192 == 0) && (N % 256 != 0 || N < 1024);
  XS_CHECK_ARG(use192 || N % 256 == 0, "gemm: N=%d must be a multiple of 192 or 256 (pad the weight rows)", N);
  if (!in_tf32 && out_f32 == 3) {
    // bf16 operands, fp16 result (fused q|k|v and the decoder K/V cache feeding the fp16 attention kernel): same
    // kernel selection as the bf16-output path below
    XS_CHECK_ARG(act == ACT_NONE, "gemm: fp16 output supports act=NONE only");
    JigsawParams jp{};
    const int bn = use192 ? 192 : 256;
    const int num_m2 = (M + 255) / 256;
    const bool fits = num_m2 * (N / bn) >= num_sms() / 2;
    if (K <= GEMM_KB_MAX * GEMM_BK && num_m2 >= num_sms() / 2) {
      if (use192) return launch_gemm<192, 8, EPI_STORE, ACT_NONE, IN_BF16, OUT_F16, 2>(XS_GEMM_ARGS);
      return launch_gemm<256, 6, EPI_STORE, ACT_NONE, IN_BF16, OUT_F16, 2>(XS_GEMM_ARGS);
    }
    if (fits && K >= 1024) {
      if (use192) return launch_gemm<192, 6, EPI_STORE, ACT_NONE, IN_BF16, OUT_F16, 1>(XS_GEMM_ARGS);
      return launch_gemm<256, 6, EPI_STORE, ACT_NONE, IN_BF16, OUT_F16, 1>(XS_GEMM_ARGS);
    }
    if (use192) return launch_gemm<192, 4, EPI_STORE, ACT_NONE, IN_BF16, OUT_F16>(XS_GEMM_ARGS);
    return launch_gemm<256, 4, EPI_STORE, ACT_NONE, IN_BF16, OUT_F16>(XS_GEMM_ARGS);
  }
  if (in_tf32 && out_f32 == 3) {  // tf32 operands, fp16 result (decoder Q/K/V projections)
    XS_CHECK_ARG(act == ACT_NONE, "gemm: tf32->fp16 supports act=NONE only");
    JigsawParams jp{};
    if (use192) return launch_gemm<192, 4, EPI_STORE, ACT_NONE, IN_TF32, OUT_F16>(XS_GEMM_ARGS);
    return launch_gemm<256, 3, EPI_STORE, ACT_NONE, IN_TF32, OUT_F16>(XS_GEMM_ARGS);
  }
  if (!in_tf32 && !out_f32) {
    // CTA pairs for the long-K GEMM (fc2, K = 1536: 1.35 PF vs 1.13 PF single-CTA, measured); the K = 384 GEMMs
    // are bound by their epilogue / operand refill per tile and run faster as independent CTAs.
    const int bn = use192 ? 192 : 256;
    const int num_m2 = (M + 255) / 256;
    const bool fits = num_m2 * (N / bn) >= num_sms() / 2;
    const bool pair = fits && K >= 1024;
    // K <= 384 (qkv, proj, fc1): A-stationary pairs once every pair of SMs has a 256-row block of its own
    const bool astat = K <= GEMM_KB_MAX * GEMM_BK && astat_pays(num_m2);
    if (astat) {
      if (use192) return dispatch_act<192, 8, IN_BF16, OUT_BF16, 2>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
      return dispatch_act<256, 6, IN_BF16, OUT_BF16, 2>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
    }
    if (pair || (fits && num_m2 >= num_sms() / 2)) {
      if (use192) return dispatch_act<192, 6, IN_BF16, OUT_BF16, 1>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
      return dispatch_act<256, 6, IN_BF16, OUT_BF16, 1>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
    }
    if (use192) return dispatch_act<192, 4, IN_BF16, OUT_BF16>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
    return dispatch_act<256, 4, IN_BF16, OUT_BF16>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
  }
  JigsawParams jp{};
  if (!in_tf32 && out_f32 == 2) {
    // bf16 operands, out (fp32) += A W^T + bias: the residual add of the DINOv2 blocks (attention.output.dense and
    // mlp.fc2 with LayerScale folded in, modeling_dinov2.py:367-386) done in the epilogue (residual tile prefetched
    // by TMA), so the delta never exists in HBM and the LayerNorm that follows reads the residual stream only
    XS_CHECK_ARG(act == ACT_NONE, "gemm: residual accumulate supports act=NONE only");
    XS_CHECK_ARG(use192, "gemm: residual accumulate needs N %% 192 == 0 (N=%d)", N);
    const int num_m2 = (M + 255) / 256;
    const bool fits = num_m2 * (N / 192) >= num_sms() / 2;
    if (K <= GEMM_KB_MAX * GEMM_BK && astat_pays(num_m2))
      return launch_gemm<192, 5, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32_ADD, 2>(XS_GEMM_ARGS);
    if (fits && (K >= 1024 || num_m2 >= num_sms() / 2))
      return launch_gemm<192, 5, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32_ADD, 1>(XS_GEMM_ARGS);
    return launch_gemm<192, 4, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32_ADD>(XS_GEMM_ARGS);
  }
  XS_CHECK_ARG(out_f32 != 2, "gemm: residual accumulate needs bf16 operands");
  if (!in_tf32 && out_f32) {  // bf16 operands, fp32 result (residual deltas kept unrounded)
    XS_CHECK_ARG(act == ACT_NONE, "gemm: bf16->fp32 supports act=NONE only");
    if (use192) return launch_gemm<192, 4, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32>(XS_GEMM_ARGS);
    return launch_gemm<256, 3, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32>(XS_GEMM_ARGS);
  }
  if (in_tf32 && out_f32) {
    // CTA pairs once there are enough 256-row tiles for every pair of SMs (patch embedding and the decoder GEMMs of a
    // full batch): each CTA stages half of the W tile, which is what bounds these fp32-operand GEMMs (L2 -> SM).
    // In-run A/B: patch embedding 0.68 -> 0.61 ms per step, decoder GEMMs 0.59 -> 0.58 (profiles/r2_experiments.txt).
    const int num_m2 = (M + 255) / 256;
    if (use192 && num_m2 * (N / 192) >= num_sms() / 2)
      return dispatch_act<192, 5, IN_TF32, OUT_F32, 1>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
    if (use192) return dispatch_act<192, 4, IN_TF32, OUT_F32>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
    return dispatch_act<256, 3, IN_TF32, OUT_F32>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
  }
  // tf32 operands, bf16 result (decoder Q/K/V projections feeding the bf16 attention kernel)
  XS_CHECK_ARG(act == ACT_NONE, "gemm: tf32->bf16 supports act=NONE only");
  if (use192) return launch_gemm<192, 4, EPI_STORE, ACT_NONE, IN_TF32, OUT_BF16>(XS_GEMM_ARGS);
  return launch_gemm<256, 3, EPI_STORE, ACT_NONE, IN_TF32, OUT_BF16>(XS_GEMM_ARGS);
}

// h (fp32, in place) += A W^T + bias, then y (bf16) = LayerNorm(h; gamma, beta, eps) in the same kernel.
// Returns 1 when the shape is not covered by the fused kernel (the caller then runs the two steps separately).
int gemm_tc_residual_ln(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int ldh,
                        const float* gamma, const float* beta, float eps, void* y, int ldy, int M, int N, int K,
                        cudaStream_t stream) {
  XS_CHECK_ARG(M > 0 && K > 0, "gemm_residual_ln: empty problem");
  XS_CHECK_ARG((K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0 && (ldh % 4) == 0 && (ldy % 8) == 0,
               "gemm_residual_ln: row pitches must be multiples of 16 bytes");
  XS_CHECK_ARG(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(h) |
                 reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(gamma) |
                 reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
               "gemm_residual_ln: pointers must be 16-byte aligned");
  if (N != 384) return 1;  // the fused epilogue needs the whole row in one CTA: two 192-wide n-tiles
  const int num_m2 = (M + 255) / 256;
  JigsawParams jp{};
  void* out = h;
  const int ldc = ldh;
  const LnParams lnp{gamma, beta, eps, h, static_cast<__nv_bfloat16*>(y), ldh, ldy, nullptr};
  if (K <= GEMM_KB_MAX * GEMM_BK && num_m2 >= num_sms() / 2)
    return launch_gemm<192, 5, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32_ADD, 2, 1>(XS_GEMM_ARGS, y, ldy, lnp);
  if (K >= 1024 && num_m2 >= num_sms() / 2)
    return launch_gemm<192, 5, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32_ADD, 1, 1>(XS_GEMM_ARGS, y, ldy, lnp);
  return 1;
}

// h (fp32, in place) += A W^T + bias; hb (bf16) = h; stats[r] = (mean, rstd) of row r of h.  The producer half of the
// folded LayerNorm (LNF = 2).  Returns 1 when the shape is not covered (same coverage as gemm_tc_residual_ln).
int gemm_tc_residual_stats(const void* A, int lda, const void* W, int ldw, const float* bias, float* h, int ldh,
                           void* hb, int ldhb, float* stats, float eps, int M, int N, int K, cudaStream_t stream) {
  XS_CHECK_ARG(M > 0 && K > 0, "gemm_residual_stats: empty problem");
  XS_CHECK_ARG((K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0 && (ldh % 4) == 0 && (ldhb % 8) == 0,
               "gemm_residual_stats: row pitches must be multiples of 16 bytes");
  XS_CHECK_ARG(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(h) |
                 reinterpret_cast<uintptr_t>(hb) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(stats)) & 15) == 0,
               "gemm_residual_stats: pointers must be 16-byte aligned");
  if (N != 384) return 1;
  const int num_m2 = (M + 255) / 256;
  JigsawParams jp{};
  void* out = h;
  const int ldc = ldh;
  const LnParams lnp{nullptr, nullptr, eps, h, static_cast<__nv_bfloat16*>(hb), ldh, ldhb, reinterpret_cast<float2*>(stats)};
  if (K <= GEMM_KB_MAX * GEMM_BK && num_m2 >= num_sms() / 2)
    return launch_gemm<192, 5, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32_ADD, 2, 2>(XS_GEMM_ARGS, hb, ldhb, lnp);
  if (K >= 1024 && num_m2 >= num_sms() / 2)
    return launch_gemm<192, 5, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32_ADD, 1, 2>(XS_GEMM_ARGS, hb, ldhb, lnp);
  return 1;
}

// out (bf16) = act(LayerNorm(h) W^T + b) with the LayerNorm folded in: A = bf16 copy of h, W = bf16(gamma * W), and per
// row rstd * (acc - mean * c1[n]) + c0[n] in the epilogue (c1[n] = sum_k W[n,k] of the folded bf16 weight, c0 = W beta + b).
// The consumer half of the folded LayerNorm (A-stationary CTA pairs for K <= 384 and many rows, else independent CTAs).
int gemm_tc_ln_folded(const void* A, int lda, const void* W, int ldw, const float* c0, const float* c1,
                      const float* stats, void* out, int ldc, int M, int N, int K, int act, cudaStream_t stream) {
  XS_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_ln_folded: empty problem");
  XS_CHECK_ARG(act == ACT_NONE || act == ACT_GELU, "gemm_ln_folded: act must be NONE or GELU, got %d", act);
  XS_CHECK_ARG((K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0 && (ldc % 8) == 0,
               "gemm_ln_folded: row pitches must be multiples of 16 bytes");
  XS_CHECK_ARG(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out) |
                 reinterpret_cast<uintptr_t>(c0) | reinterpret_cast<uintptr_t>(c1) | reinterpret_cast<uintptr_t>(stats)) & 15) == 0,
               "gemm_ln_folded: pointers must be 16-byte aligned");
  const bool use192 = (N % 192 == 0) && (N % 256 != 0 || N < 1024);
  XS_CHECK_ARG(use192 || N % 256 == 0, "gemm_ln_folded: N=%d must be a multiple of 192 or 256", N);
  const int num_m2 = (M + 255) / 256;
  JigsawParams jp{};
  const float* bias = c0;
  const LnParams lnp{c1, nullptr, 0.f, nullptr, nullptr, 0, 0,
                     const_cast<float2*>(reinterpret_cast<const float2*>(stats))};
  if (!(K <= GEMM_KB_MAX * GEMM_BK && num_m2 >= num_sms() / 2)) {  // few rows or long K: independent CTAs
    if (act == ACT_GELU) {
      if (use192) return launch_gemm<192, 4, EPI_STORE, ACT_LN_GELU, IN_BF16, OUT_BF16>(XS_GEMM_ARGS, nullptr, 0, lnp);
      return launch_gemm<256, 4, EPI_STORE, ACT_LN_GELU, IN_BF16, OUT_BF16>(XS_GEMM_ARGS, nullptr, 0, lnp);
    }
    if (use192) return launch_gemm<192, 4, EPI_STORE, ACT_LN_NONE, IN_BF16, OUT_BF16>(XS_GEMM_ARGS, nullptr, 0, lnp);
    return launch_gemm<256, 4, EPI_STORE, ACT_LN_NONE, IN_BF16, OUT_BF16>(XS_GEMM_ARGS, nullptr, 0, lnp);
  }
  if (act == ACT_GELU) {
    if (use192) return launch_gemm<192, 8, EPI_STORE, ACT_LN_GELU, IN_BF16, OUT_BF16, 2>(XS_GEMM_ARGS, nullptr, 0, lnp);
    return launch_gemm<256, 6, EPI_STORE, ACT_LN_GELU, IN_BF16, OUT_BF16, 2>(XS_GEMM_ARGS, nullptr, 0, lnp);
  }
  if (use192) return launch_gemm<192, 8, EPI_STORE, ACT_LN_NONE, IN_BF16, OUT_BF16, 2>(XS_GEMM_ARGS, nullptr, 0, lnp);
  return launch_gemm<256, 6, EPI_STORE, ACT_LN_NONE, IN_BF16, OUT_BF16, 2>(XS_GEMM_ARGS, nullptr, 0, lnp);
}

// head.2 Linear (384 -> 196, weight rows padded to 224) + sigmoid/tanh (+pow) + jigsaw scatter
int head_jigsaw_tc(const void* A, int lda, const void* W, int ldw, const float* bias, float* score, int B, int ph,
                   int pw, int K, int use_tanh, float power, int in_tf32, cudaStream_t stream) {
  const int al = in_tf32 ? 4 : 8;
  XS_CHECK_ARG(B > 0 && ph > 0 && pw > 0, "head_jigsaw: empty problem");
  XS_CHECK_ARG((K % al) == 0 && (lda % al) == 0 && (ldw % al) == 0, "head_jigsaw: K/lda/ldw must be 16-byte multiples");
  XS_CHECK_ARG((reinterpret_cast<uintptr_t>(score) & 7) == 0, "head_jigsaw: the score map must be 8-byte aligned");
  JigsawParams jp;
  jp.score = score;
  jp.P = ph * pw;
  jp.pw = pw;
  jp.Wout = 14 * pw;
  jp.HWout = 14 * ph * 14 * pw;
  jp.use_tanh = use_tanh;
  jp.power = power;
  void* out = nullptr;
  const int ldc = 0, M = B * ph * pw, N = 224;
  if (in_tf32) return launch_gemm<224, 2, EPI_JIGSAW, ACT_NONE, IN_TF32, OUT_F32>(XS_GEMM_ARGS);
  return launch_gemm<224, 2, EPI_JIGSAW, ACT_NONE, IN_BF16, OUT_F32>(XS_GEMM_ARGS);
}

}  // namespace xs
