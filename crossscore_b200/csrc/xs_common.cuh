// Common device/host helpers for the crossscore_b200 sm_100a kernels.
//
// Everything Blackwell-specific (mbarrier, TMA, tcgen05, TMEM) is wrapped here as thin inline-PTX
// helpers so the kernels read as plain CUDA.  sm_100a only; there is no fallback path.
#pragma once

#include <cuda.h>          // CUtensorMap (types only; the driver is reached through cudart)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace xs {

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI: every export returns int, 0 = ok, <0 = bad argument, >0 = cudaError_t)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define XS_CHECK_ARG(cond, ...)                    \
  do {                                             \
    if (!(cond)) {                                 \
      ::xs::set_last_error(__VA_ARGS__);           \
      return -1;                                   \
    }                                              \
  } while (0)

#define XS_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::xs::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                           __LINE__);                                                         \
      return (int)_e;                                                                         \
    }                                                                                         \
  } while (0)

#define XS_LAUNCH_CHECK() XS_CUDA(cudaGetLastError())

enum DType : int { XS_BF16 = 0, XS_F32 = 1, XS_TF32 = 2, XS_F16 = 3 };
enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_LEAKY = 3,
                 // LayerNorm folded into the GEMM (xs_gemm_ln_folded): per row rstd * (acc - mean * c1[n]) + c0[n], then the activation
                 ACT_LN_NONE = 4, ACT_LN_GELU = 5 };

int num_sms();

// host: build TMA descriptors through the driver entry point obtained from cudart (no -lcuda)
// rank-2: dims {inner, outer}; rank-3: dims {inner, mid, outer}; strides in BYTES for dims 1..rank-1
enum Swizzle : int { SWZ_NONE = 0, SWZ_64B = 2, SWZ_128B = 3 };
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, Swizzle swz);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// generic device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
  return r;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2, two fp32 lanes per issue slot)
__device__ __forceinline__ float2 ffma2_bcast(float a0, float a1, float b, float c) {  // (a0, a1) * b + c
  uint64_t d;
  asm("{\n"
      ".reg .b64 a, b, c;\n"
      "mov.b64 a, {%1, %2};\n"
      "mov.b64 b, {%3, %3};\n"
      "mov.b64 c, {%4, %4};\n"
      "fma.rn.f32x2 %0, a, b, c;\n"
      "}\n"
      : "=l"(d)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
  return make_float2(__uint_as_float(static_cast<uint32_t>(d)), __uint_as_float(static_cast<uint32_t>(d >> 32)));
}
__device__ __forceinline__ float2 fadd2(float2 x, float2 y) {
  uint64_t d;
  asm("{\n"
      ".reg .b64 a, b;\n"
      "mov.b64 a, {%1, %2};\n"
      "mov.b64 b, {%3, %4};\n"
      "add.rn.f32x2 %0, a, b;\n"
      "}\n"
      : "=l"(d)
      : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y));
  return make_float2(__uint_as_float(static_cast<uint32_t>(d)), __uint_as_float(static_cast<uint32_t>(d >> 32)));
}
__device__ __forceinline__ float2 ffma2(float2 x, float2 y, float2 z) {  // x * y + z, elementwise
  uint64_t d;
  asm("{\n"
      ".reg .b64 a, b, c;\n"
      "mov.b64 a, {%1, %2};\n"
      "mov.b64 b, {%3, %4};\n"
      "mov.b64 c, {%5, %6};\n"
      "fma.rn.f32x2 %0, a, b, c;\n"
      "}\n"
      : "=l"(d)
      : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y), "f"(z.x), "f"(z.y));
  return make_float2(__uint_as_float(static_cast<uint32_t>(d)), __uint_as_float(static_cast<uint32_t>(d >> 32)));
}
// 2^x for a pair of fp32 values on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + r with the
// 1.5 * 2^23 magic constant, cubic minimax polynomial for 2^r on [-0.5, 0.5] (max relative error 7.5e-5, far below
// the bf16 rounding of the softmax numerators it feeds), exponent inserted by an integer shift-add.
// Inputs are clamped at -126 (result flushes towards 0); inputs up to +127 are exact in range.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
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
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// GELU (erf form) for the bf16 path's GEMM epilogue:  0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))), coefficients
// fitted to the exact-erf GELU (max abs deviation 2.5e-5 over R, far below the bf16 rounding of the result);
// 7 FMA-pipe ops + one MUFU.TANH per element, so the fc1 epilogue (393k GELUs per 128x256 tile ... per SM)
// stays under the tile's MMA time.  x^2 is clamped at 64 (tanh is saturated there; keeps the quintic monotone).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float t = fmaf(-3.51516783e-04f, x2, 3.70056460e-02f);
  t = fmaf(t, x2, 7.97507884e-01f);
  const float th = fast_tanh(x * t);
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}

// the same GELU on a pair of values with packed fp32x2 FMAs (bit-identical to gelu_erf_fast: same operations, same
// roundings; 13 issue slots per pair instead of 20 -- the fc1 epilogue is issue-bound next to its tile's MMAs)
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
  const float2 zero = make_float2(0.f, 0.f);
  float2 x2 = ffma2(x, x, zero);
  x2.x = fminf(x2.x, 64.0f);
  x2.y = fminf(x2.y, 64.0f);
  float2 t = ffma2(make_float2(-3.51516783e-04f, -3.51516783e-04f), x2, make_float2(3.70056460e-02f, 3.70056460e-02f));
  t = ffma2(t, x2, make_float2(7.97507884e-01f, 7.97507884e-01f));
  const float2 u = ffma2(x, t, zero);
  const float2 th = make_float2(fast_tanh(u.x), fast_tanh(u.y));
  const float2 hx = ffma2(x, make_float2(0.5f, 0.5f), zero);
  return ffma2(hx, th, hx);
}

template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
  if constexpr (ACT == ACT_GELU) return gelu_erf_fast(v);
  if constexpr (ACT == ACT_RELU) return fmaxf(v, 0.0f);
  if constexpr (ACT == ACT_LEAKY) return v >= 0.0f ? v : 0.01f * v;
  return v;
}

// act(v0 + b0), act(v1 + b1) for a pair of accumulator values (packed path for GELU)
template <int ACT>
__device__ __forceinline__ float2 bias_act2(float v0, float v1, float b0, float b1) {
  const float2 x = fadd2(make_float2(v0, v1), make_float2(b0, b1));  // one FADD2 for the two bias adds
  if constexpr (ACT == ACT_GELU) return gelu_erf_fast2(x);
  else return make_float2(apply_act<ACT>(x.x), apply_act<ACT>(x.y));
}

// act(acc * rstd + (rm * c1 + c0)) for a pair: the LayerNorm-folded epilogue (rm = -rstd * mean of the row)
template <int ACT>
__device__ __forceinline__ float2 lnfold_act2(float v0, float v1, float rstd, float rm, float c1a, float c1b, float c0a,
                                              float c0b) {
  const float2 t = ffma2(make_float2(rm, rm), make_float2(c1a, c1b), make_float2(c0a, c0b));
  const float2 x = ffma2(make_float2(v0, v1), make_float2(rstd, rstd), t);
  if constexpr (ACT == ACT_LN_GELU) return gelu_erf_fast2(x);
  else return x;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
// A barrier named by its 32-bit shared-window address (kernels whose hot loops wait on many barriers keep one
// base register instead of re-deriving shared addresses from generic pointers at every use).
struct SmemBar {
  uint32_t addr;
  __device__ __forceinline__ SmemBar operator+(int i) const { return SmemBar{addr + 8u * static_cast<uint32_t>(i)}; }
  __device__ __forceinline__ SmemBar operator+(uint32_t i) const { return SmemBar{addr + 8u * i}; }
};
__device__ __forceinline__ uint32_t bar_u32(const uint64_t* bar) { return smem_u32(bar); }
__device__ __forceinline__ uint32_t bar_u32(SmemBar bar) { return bar.addr; }

template <class B>
__device__ __forceinline__ void mbar_init(B bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
template <class B>
__device__ __forceinline__ void mbar_expect_tx(B bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u32(bar)), "r"(bytes)
               : "memory");
}
template <class B>
__device__ __forceinline__ void mbar_arrive(B bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_u32(bar)) : "memory");
}
template <class B>
__device__ __forceinline__ bool mbar_try_wait(B bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin on the phase bit.  try_wait parks the warp in hardware for up to the suspend-time hint, so the loop
// issues few instructions.  A pipeline bug would otherwise hang the GPU box until the job limit, so the wait
// traps after a few seconds (clock checked every 4096 polls) and the host sees a launch failure instead.
#ifndef XS_MBAR_HINT_NS
#define XS_MBAR_HINT_NS 20000
#endif
template <class B>
__device__ __forceinline__ bool mbar_try_wait_hint(B bar, uint32_t parity) {
#if XS_MBAR_HINT_NS == 0
  return mbar_try_wait(bar, parity);  // default (implementation-defined, short) suspend
#else
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar_u32(bar)), "r"(parity), "r"((uint32_t)XS_MBAR_HINT_NS)
      : "memory");
  return ok != 0;
#endif
}
// Fully inline (no calls: a call inside a setmaxnreg-raised region would pin that region to the kernel-wide
// register cap).  A pipeline bug traps after a few seconds instead of hanging the GPU box.
template <class B>
__device__ __forceinline__ void mbar_wait(B bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  // Slow path: every poll parks the warp in hardware (suspend-time hint), so the loop issues three instructions per
  // wake-up.  The watchdog reads the clock only once per 2^16 polls (>= 1 s of parked time) and traps ~4 s later, so a
  // pipeline bug fails the launch instead of hanging the GPU box; nothing of it is on the common path.
  long long t0 = 0;
#pragma unroll 1
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try_wait_hint(bar, parity)) return;
    if ((spins & 0xFFFFu) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
template <class B>
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, B bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
template <class B>
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, B bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (TMA store / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// per-warpgroup register re-budgeting (all warps of the warpgroup execute it)
template <int REGS>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp as alloc
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
template <class B>
__device__ __forceinline__ void tc_commit(B bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor (UMMA).  128-byte swizzle, version 1 (Blackwell).
//   K-major  operand tile [rows][64 bf16]: SBO = 1024 B (8 rows x 128 B), LBO unused (=1)
//   MN-major operand tile [k rows][64 bf16]: SBO = 1024 B (8 k-rows), LBO = stride between 64-wide MN chunks
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

// Split form used by the MMA-issuing warps: the high word is a constant, the low word is (addr >> 4) | LBO and
// is advanced by plain 32-bit adds (K-step inside a 128B-swizzled row: +32 B -> +2; 16 MN-major rows: +2048 B -> +128).
constexpr uint32_t UMMA_DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO=1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}

// One lane of a converged warp (elect.sync).  The MMA / TMA issuing warps stay converged and keep their
// operands warp-uniform, so the tcgen05 / TMA instructions take uniform-register operands directly; a
// `lane == 0` branch instead makes the compiler wrap every such instruction in an elect/R2UR waterfall loop
// (~80 cycles per MMA, measured) that starves the tensor pipe when the MMAs are small.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// Instruction descriptor, kind::f16 with bf16 A/B and fp32 accumulate.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                                 // D format f32
         | (1u << 7)                               // A format bf16
         | (1u << 10)                              // B format bf16
         | (static_cast<uint32_t>(a_mn_major) << 15)
         | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(N >> 3) << 17)
         | (static_cast<uint32_t>(M >> 4) << 24);
}

// kind::f16 with fp16 A/B; D is fp16 (d_f32 = 0) or fp32.  (fp16 accumulators are only legal with fp16 operands, and
// the A / B formats cannot be mixed: both measured, tools/ubench_f16acc.cu.)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major, int d_f32) {
  return (static_cast<uint32_t>(d_f32) << 4)
         | (static_cast<uint32_t>(a_mn_major) << 15)
         | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(N >> 3) << 17)
         | (static_cast<uint32_t>(M >> 4) << 24);
}

__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, kind::tf32: A/B are fp32 in smem, read as TF32 (10-bit mantissa), K = 8 per instruction
__device__ __forceinline__ void umma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// lo/hi descriptor forms (see umma_desc_lo)
template <bool TF32>
__device__ __forceinline__ void umma_ss_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %5};\n"
        "mov.b64 db, {%2, %5};\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %5};\n"
        "mov.b64 db, {%2, %5};\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128)
        : "memory");
  }
}
__device__ __forceinline__ void umma_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "mov.b64 db, {%2, %5};\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): one MMA spans two SMs, each CTA stages its own A rows and
// half of the B rows, so the shared-memory fill traffic per output tile drops by a third.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Arrive on a (possibly remote) barrier of the cluster.  RELAXED: the callers order their tcgen05 / TMEM accesses
// with tcgen05.wait + tcgen05.fence::before_thread_sync and publish no generic-proxy data through this barrier; a
// release at cluster scope costs a MEMBAR + ERRBAR per arrive (30 % of the epilogue warps' time, measured).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {  // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into THIS CTA's shared memory whose completion bytes are credited to `bar_cluster_addr`
// (the leader CTA's mbarrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs once all prior MMAs of this thread completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .b16 m;\n"
      "mov.b16 m, 3;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
      "}\n" ::"r"(bar_u32(bar))
      : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]; issued by the leader CTA only
template <bool TF32 = false>
__device__ __forceinline__ void umma_ss_lh_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %5};\n"
        "mov.b64 db, {%2, %5};\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %5};\n"
        "mov.b64 db, {%2, %5};\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128)
        : "memory");
  }
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TMEM -> registers, 32 lanes x 32 columns (thread i of the warp gets lane base+i, 32 consecutive columns)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 16 columns into the low half of a 32-register buffer
__device__ __forceinline__ void tmem_ld16_lo(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 8 columns into r[16..23] of a 32-register buffer
__device__ __forceinline__ void tmem_ld8_at16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15])
               :
               : "memory");
}
// tcgen05.wait::ld that also names the destination registers, so the compiler cannot move their consumers
// above the wait (the loads are asynchronous until this point)
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// registers -> TMEM, 32 lanes x 16 columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 64 TMEM columns holding one 16-bit value each (fp16 accumulators) -> 32 registers: .pack::16b puts column 2i in the
// low and column 2i+1 in the high half of register i (1.75x the column rate of the unpacked load, measured)
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 columns of 16-bit values -> 16 registers
__device__ __forceinline__ void tmem_ld16_pack16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM, 32 lanes x 32 columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// ---- 16-lane fragment shapes (two warps share a 32-lane TMEM quarter: lower / upper 16 lanes) ----
// tcgen05.ld.16x256b: thread t of the warp receives rows (lanes) base + t/4 and base + t/4 + 8; per 8-column group k:
// r[4k], r[4k+1] = row t/4, columns 8k + 2(t%4) + {0,1};  r[4k+2], r[4k+3] = row t/4 + 8, same columns.
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr)
               : "memory");
}
// 32 columns into r[0..15]
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
// 16 columns into r[16..23]
__device__ __forceinline__ void tmem_ld_16x256b_x2_hi(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// tcgen05.st.16x128b: per 4-column group k: r[2k] = row t/4, column 4k + t%4;  r[2k+1] = row t/4 + 8, same column
// (column = 32-bit TMEM column, i.e. one bf16x2 pair of P: the packed pairs of a 16x256b fragment map 1:1)
__device__ __forceinline__ void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}

// packed half (f16x2) arithmetic on raw 32-bit registers
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2_max(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2_ex2(uint32_t a) {  // one MUFU instruction, two exponentials
  uint32_t d;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ float h2_lo(uint32_t a) {
  float f;
  asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f) : "r"(a));
  return f;
}
__device__ __forceinline__ float h2_hi(uint32_t a) {
  float f;
  asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(f) : "r"(a));
  return f;
}
__device__ __forceinline__ uint32_t h2_bcast(float v) { return pack_f16x2(v, v); }
#endif  // __CUDACC__

}  // namespace xs
