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
#include "xs_common.cuh"

namespace xs {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_THREADS = 256;
constexpr int ACC_STAGE_COLS = 256;  // TMEM column offset between the two accumulator stages
constexpr int JIG_LD = 197;          // padded row length of the fp32 score staging tile

enum Epi : int { EPI_STORE = 0, EPI_JIGSAW = 1 };
enum InT : int { IN_BF16 = 0, IN_TF32 = 1 };
enum OutT : int { OUT_BF16 = 0, OUT_F32 = 1 };

struct JigsawParams {
  float* score;  // (B, 14*ph, 14*pw) fp32
  int P;         // tokens per map = ph*pw
  int pw;
  int Wout;      // 14*pw
  int HWout;     // 14*ph*14*pw
  int use_tanh;
  float power;
};

template <int BN, int STAGES, int EPI>
struct GemmSmem {
  static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = BN * GEMM_BK * 2;
  static constexpr uint32_t STAGING_BYTES = (EPI == EPI_STORE) ? 2 * GEMM_BM * 128 : GEMM_BM * JIG_LD * 4;
  static constexpr uint32_t OFF_A = 0;
  static constexpr uint32_t OFF_B = OFF_A + STAGES * A_BYTES;
  static constexpr uint32_t OFF_STAGING = OFF_B + STAGES * B_BYTES;
  static constexpr uint32_t OFF_BAR = (OFF_STAGING + STAGING_BYTES + 15u) & ~15u;
  static constexpr uint32_t BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr uint32_t TOTAL = OFF_BAR + BAR_BYTES + 1024;  // +1024: manual 1 KB alignment of the base
};

template <int BN, int STAGES, int EPI, int ACT, int IN, int OUT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, int M, int N, int K,
               JigsawParams jp) {
  using L = GemmSmem<BN, STAGES, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint8_t* smA = smem + L::OFF_A;
  uint8_t* smB = smem + L::OFF_B;
  uint8_t* staging = smem + L::OFF_STAGING;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = N / BN;
  constexpr int BKE = (IN == IN_TF32) ? 32 : 64;  // elements per 128-byte smem row
  const int num_k = (K + BKE - 1) / BKE;
  const int num_tiles = num_m * num_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if constexpr (EPI == EPI_STORE) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);  // one elected lane per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[stage], L::A_BYTES + L::B_BYTES);
          tma_load_2d(smA + stage * L::A_BYTES, &tmA, &full_bar[stage], kb * BKE, m_blk * GEMM_BM);
          tma_load_2d(smB + stage * L::B_BYTES, &tmW, &full_bar[stage], kb * BKE, n_blk * BN);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, uniform operands, one elected lane issues) =========
    constexpr uint32_t idesc = (IN == IN_TF32) ? umma_idesc_tf32(GEMM_BM, BN) : umma_idesc_bf16(GEMM_BM, BN, 0, 0);
    const uint32_t tb = warp_uniform(tmem_base);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smA), 16);
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(smB), 16);
    uint32_t stage = 0, phase = 0, acc_stage = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc_stage], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tb + acc_stage * ACC_STAGE_COLS;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + stage * (L::A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (L::B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x 32 bytes of K per 128-byte row (16 bf16 or 8 tf32 each)
            umma_ss_lh<IN == IN_TF32>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (kb == num_k - 1) tc_commit(&tmem_full[acc_stage]);  // accumulator complete -> epilogue
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      mbar_wait(&tmem_full[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc_stage * ACC_STAGE_COLS + (static_cast<uint32_t>(q * 32) << 16);

      if constexpr (EPI == EPI_STORE) {
        constexpr int CW = (OUT == OUT_F32) ? 32 : 64;  // columns per 128-byte staging row
        constexpr int HV = CW / 2;                       // accumulator values per thread per chunk
        constexpr int NCHUNK = BN / CW;
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          const uint32_t buf = chunk_counter & 1;
          ++chunk_counter;
          // the TMA store that last read this staging buffer (two chunks ago) must have drained
          if (epi_tid == 0) tma_store_wait_read<1>();
          named_bar_sync(1, EPI_THREADS);
          uint32_t v[HV];
          if constexpr (HV == 32) tmem_ld32(taddr0 + c * CW + half * HV, v);
          else tmem_ld16(taddr0 + c * CW + half * HV, v);
          tc_wait_ld();
          if (c == NCHUNK - 1) {  // accumulator fully read: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc_stage]);
          }
          const int n0 = n_blk * BN + c * CW;
          uint8_t* srow = staging + buf * (GEMM_BM * 128) + row * 128;
          const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + half * HV);
          // 128B swizzle: 16-byte chunk index XOR (row mod 8); conflict-free for thread==row writes
          if constexpr (OUT == OUT_BF16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 ba = __ldg(b4 + j * 2), bb = __ldg(b4 + j * 2 + 1);
              const int o = j * 8;
              const float x0 = apply_act<ACT>(__uint_as_float(v[o + 0]) + ba.x);
              const float x1 = apply_act<ACT>(__uint_as_float(v[o + 1]) + ba.y);
              const float x2 = apply_act<ACT>(__uint_as_float(v[o + 2]) + ba.z);
              const float x3 = apply_act<ACT>(__uint_as_float(v[o + 3]) + ba.w);
              const float x4 = apply_act<ACT>(__uint_as_float(v[o + 4]) + bb.x);
              const float x5 = apply_act<ACT>(__uint_as_float(v[o + 5]) + bb.y);
              const float x6 = apply_act<ACT>(__uint_as_float(v[o + 6]) + bb.z);
              const float x7 = apply_act<ACT>(__uint_as_float(v[o + 7]) + bb.w);
              uint4 pk;
              pk.x = pack_bf16x2(x0, x1);
              pk.y = pack_bf16x2(x2, x3);
              pk.z = pack_bf16x2(x4, x5);
              pk.w = pack_bf16x2(x6, x7);
              *reinterpret_cast<uint4*>(srow + (((half * 4 + j) ^ (row & 7)) << 4)) = pk;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 ba = __ldg(b4 + j);
              float4 o4;
              o4.x = apply_act<ACT>(__uint_as_float(v[j * 4 + 0]) + ba.x);
              o4.y = apply_act<ACT>(__uint_as_float(v[j * 4 + 1]) + ba.y);
              o4.z = apply_act<ACT>(__uint_as_float(v[j * 4 + 2]) + ba.z);
              o4.w = apply_act<ACT>(__uint_as_float(v[j * 4 + 3]) + ba.w);
              *reinterpret_cast<float4*>(srow + (((half * 4 + j) ^ (row & 7)) << 4)) = o4;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, EPI_THREADS);
          if (epi_tid == 0) {
            tma_store_2d(&tmC, staging + buf * (GEMM_BM * 128), n0, m_blk * GEMM_BM);
            tma_store_commit();
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
              float s = jp.use_tanh ? tanhf(z) : 1.0f / (1.0f + expf(-z));
              if (jp.power != 1.0f) s = powf(s, jp.power);
              stile[row * JIG_LD + k] = s;
            }
          }
        }
        (void)LAST;
        named_bar_sync(1, EPI_THREADS);
        const int m0 = m_blk * GEMM_BM;
        const int rows_valid = min(GEMM_BM, M - m0);
        // (i, token, j) order: consecutive tokens of one grid row are contiguous in the image row
        for (int e = epi_tid; e < 14 * GEMM_BM * 14; e += EPI_THREADS) {
          const int i = e / (GEMM_BM * 14);
          const int rem = e - i * (GEMM_BM * 14);
          const int tt = rem / 14;
          const int j = rem - tt * 14;
          if (tt < rows_valid) {
            const int t = m0 + tt;
            const int b = t / jp.P;
            const int p = t - b * jp.P;
            const int r = p / jp.pw;
            const int cc = p - r * jp.pw;
            jp.score[static_cast<size_t>(b) * jp.HWout + static_cast<size_t>(14 * r + i) * jp.Wout + 14 * cc + j] =
                stile[tt * JIG_LD + i * 14 + j];
          }
        }
        named_bar_sync(1, EPI_THREADS);  // staging tile is reused by the next tile
      }
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
    if constexpr (EPI == EPI_STORE) {
      if (epi_tid == 0) tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES, int EPI, int ACT, int IN, int OUT>
static int launch_gemm(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, int M,
                       int N, int K, JigsawParams jp, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, EPI>;
  constexpr int IN_B = (IN == IN_TF32) ? 4 : 2;
  constexpr int OUT_B = (OUT == OUT_F32) ? 4 : 2;
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
    uint32_t box[2] = {BKE, (uint32_t)BN};
    int rc = make_tmap(&tmW, W, IN_B, 2, dims, strides, box, SWZ_128B);
    if (rc) return rc;
  }
  if (EPI == EPI_STORE) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ldc * OUT_B};
    uint32_t box[2] = {128 / OUT_B, GEMM_BM};
    int rc = make_tmap(&tmC, out, OUT_B, 2, dims, strides, box, SWZ_128B);
    if (rc) return rc;
  } else {
    tmC = tmA;  // unused
  }
  auto kern = gemm_tc_kernel<BN, STAGES, EPI, ACT, IN, OUT>;
  XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL));
  const int num_tiles = ((M + GEMM_BM - 1) / GEMM_BM) * (N / BN);
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmW, tmC, bias, M, N, K, jp);
  XS_LAUNCH_CHECK();
  return 0;
}

#define XS_GEMM_ARGS A, lda, W, ldw, bias, out, ldc, M, N, K, jp, stream

template <int BN, int STAGES, int IN, int OUT>
static int dispatch_act(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldc, int M,
                        int N, int K, int act, cudaStream_t stream) {
  JigsawParams jp{};
  switch (act) {
    case ACT_NONE: return launch_gemm<BN, STAGES, EPI_STORE, ACT_NONE, IN, OUT>(XS_GEMM_ARGS);
    case ACT_GELU: return launch_gemm<BN, STAGES, EPI_STORE, ACT_GELU, IN, OUT>(XS_GEMM_ARGS);
    case ACT_RELU: return launch_gemm<BN, STAGES, EPI_STORE, ACT_RELU, IN, OUT>(XS_GEMM_ARGS);
    case ACT_LEAKY: return launch_gemm<BN, STAGES, EPI_STORE, ACT_LEAKY, IN, OUT>(XS_GEMM_ARGS);
  }
  set_last_error("xs_gemm_bias_act: unknown activation %d", act);
  return -1;
}

// Tensor-core GEMM entry used by xs_api.cu.  in_tf32: A/W are fp32 (TF32 multiply), else bf16.
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
  if (!in_tf32 && !out_f32) {
    if (use192) return dispatch_act<192, 4, IN_BF16, OUT_BF16>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
    return dispatch_act<256, 3, IN_BF16, OUT_BF16>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
  }
  JigsawParams jp{};
  if (!in_tf32 && out_f32) {  // bf16 operands, fp32 result (residual deltas kept unrounded)
    XS_CHECK_ARG(act == ACT_NONE, "gemm: bf16->fp32 supports act=NONE only");
    if (use192) return launch_gemm<192, 4, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32>(XS_GEMM_ARGS);
    return launch_gemm<256, 3, EPI_STORE, ACT_NONE, IN_BF16, OUT_F32>(XS_GEMM_ARGS);
  }
  if (in_tf32 && out_f32) {
    if (use192) return dispatch_act<192, 4, IN_TF32, OUT_F32>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
    return dispatch_act<256, 3, IN_TF32, OUT_F32>(A, lda, W, ldw, bias, out, ldc, M, N, K, act, stream);
  }
  // tf32 operands, bf16 result (decoder Q/K/V projections feeding the bf16 attention kernel)
  XS_CHECK_ARG(act == ACT_NONE, "gemm: tf32->bf16 supports act=NONE only");
  if (use192) return launch_gemm<192, 4, EPI_STORE, ACT_NONE, IN_TF32, OUT_BF16>(XS_GEMM_ARGS);
  return launch_gemm<256, 3, EPI_STORE, ACT_NONE, IN_TF32, OUT_BF16>(XS_GEMM_ARGS);
}

// head.2 Linear (384 -> 196, weight rows padded to 224) + sigmoid/tanh (+pow) + jigsaw scatter
int head_jigsaw_tc(const void* A, int lda, const void* W, int ldw, const float* bias, float* score, int B, int ph,
                   int pw, int K, int use_tanh, float power, int in_tf32, cudaStream_t stream) {
  const int al = in_tf32 ? 4 : 8;
  XS_CHECK_ARG(B > 0 && ph > 0 && pw > 0, "head_jigsaw: empty problem");
  XS_CHECK_ARG((K % al) == 0 && (lda % al) == 0 && (ldw % al) == 0, "head_jigsaw: K/lda/ldw must be 16-byte multiples");
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
