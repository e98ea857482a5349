// Hardware probe (development tool): tcgen05.mma kind::f16 with an fp16 ACCUMULATOR (D format f16), how those
// accumulators sit in TMEM (tcgen05.ld 32x32b plain vs .pack::16b), and whether a TMEM A operand in fp16 may be
// multiplied with a bf16 B operand from shared memory (A/B formats are separate idesc fields).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I crossscore_b200/csrc tools/ubench_f16acc.cu -o tools/ubench_bin_f16acc
#include <cstdio>
#include <cuda_fp16.h>
#include "xs_common.cuh"
namespace xs { void set_last_error(const char*, ...) {} int num_sms() { return 148; }
int make_tmap(CUtensorMap*, const void*, int, int, const uint64_t*, const uint64_t*, const uint32_t*, Swizzle) { return 0; } }
using namespace xs;

__device__ __forceinline__ void ld_pack16_x32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  tmem_st16(taddr, reinterpret_cast<const uint32_t (&)[16]>(r[0]));
  tmem_st16(taddr + 16, reinterpret_cast<const uint32_t (&)[16]>(r[16]));
}

__host__ __device__ inline int qv(int r, int c) { return ((r + c) % 5) - 2; }
__host__ __device__ inline int kv(int j, int c) { return ((j * 3 + c) % 7) - 3; }
__host__ __device__ inline int pv(int r, int j) { return (r + 2 * j) % 3; }
__host__ __device__ inline int vv(int j, int d) { return ((j + d) % 5) - 2; }

// byte offset of element (r, c) of a [rows][64 x 16-bit] tile in the 128-byte swizzled layout (8-row atoms)
__device__ __forceinline__ uint32_t sw128(int r, int c) {
  return (r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 3) ^ (r & 7)) & 7) << 4) + (c & 7) * 2;
}

__host__ __device__ constexpr uint32_t idesc(int M, int N, int dfmt, int afmt, int bfmt, int a_mn, int b_mn) {
  return (uint32_t(dfmt) << 4) | (uint32_t(afmt) << 7) | (uint32_t(bfmt) << 10) | (uint32_t(a_mn) << 15) |
         (uint32_t(b_mn) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// out: [0] mismatches S via pack16 (lo=even col), [1] (lo=odd col), [2] mismatches S plain low-half, [3..6] raw words,
// [8] mismatches O with A=f16 x B=bf16, [9] mismatches O with A=bf16 x B=bf16
__global__ void __launch_bounds__(128) probe(int* out, float* dump, int mode) {
  __shared__ __align__(1024) uint8_t smQ[128 * 128];
  __shared__ __align__(1024) uint8_t smK[64 * 128];
  __shared__ __align__(1024) uint8_t smV[64 * 128];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 128 * 64; e += 128) {
    const int r = e >> 6, c = e & 63;
    if (mode & 24) *reinterpret_cast<__half*>(smQ + sw128(r, c)) = __float2half((float)qv(r, c));
    else *reinterpret_cast<__nv_bfloat16*>(smQ + sw128(r, c)) = __float2bfloat16((float)qv(r, c));
  }
  for (int e = tid; e < 64 * 64; e += 128) {
    const int r = e >> 6, c = e & 63;
    if (mode & 24) {
      *reinterpret_cast<__half*>(smK + sw128(r, c)) = __float2half((float)kv(r, c));
      *reinterpret_cast<__half*>(smV + sw128(r, c)) = __float2half((float)vv(r, c));
    } else {
      *reinterpret_cast<__nv_bfloat16*>(smK + sw128(r, c)) = __float2bfloat16((float)kv(r, c));
      *reinterpret_cast<__nv_bfloat16*>(smV + sw128(r, c)) = __float2bfloat16((float)vv(r, c));
    }
  }
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 256);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  const uint32_t lane_off = uint32_t(warp * 32) << 16;
  const int row = warp * 32 + lane;
  // ---- S = Q K^T with fp16 accumulators at columns [0, 64)
  if (mode & 9) {
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t q_lo = umma_desc_lo(smem_u32(smQ), 16), k_lo = umma_desc_lo(smem_u32(smK), 16);
      for (int k = 0; k < 4; ++k) umma_ss_lh<false>(tb, q_lo + 2 * k, k_lo + 2 * k, idesc(128, 64, 0, (mode & 8) ? 0 : 1, (mode & 8) ? 0 : 1, 0, 0), k != 0);
      tc_commit(&bar[0]);
    }
    __syncwarp();
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();
  uint32_t raw[32], pk[32];
  tmem_ld32(tb + lane_off, raw);
  tmem_ld_wait32(raw);
  ld_pack16_x32(tb + lane_off, pk);
  tmem_ld_wait32(pk);
  int bad_even = 0, bad_odd = 0, bad_plain = 0;
  for (int i = 0; i < 32; ++i) {
    int e0 = 0, e1 = 0;
    for (int c = 0; c < 64; ++c) { e0 += qv(row, c) * kv(2 * i, c); e1 += qv(row, c) * kv(2 * i + 1, c); }
    const float lo = __half2float(__ushort_as_half((unsigned short)(pk[i] & 0xffff)));
    const float hi = __half2float(__ushort_as_half((unsigned short)(pk[i] >> 16)));
    bad_even += (lo != (float)e0) + (hi != (float)e1);
    bad_odd += (lo != (float)e1) + (hi != (float)e0);
    int ei = 0;
    for (int c = 0; c < 64; ++c) ei += qv(row, c) * kv(i, c);
    bad_plain += (__half2float(__ushort_as_half((unsigned short)(raw[i] & 0xffff))) != (float)ei);
  }
  atomicAdd(&out[0], bad_even); atomicAdd(&out[1], bad_odd); atomicAdd(&out[2], bad_plain);
  if (tid == 5) { for (int i = 0; i < 4; ++i) out[3 + i] = (int)raw[i]; }
  }
  // ---- O = P V:  P (A operand, TMEM columns [64, 96), 64 x 16-bit per row) in fp16, then in bf16; V bf16 MN-major
  int nb = 0;
  for (int variant = 0; variant < 2; ++variant) {
    if (!(mode & (2 << variant)) && !(variant == 0 && (mode & 16))) continue;
    uint32_t pp[32];
    for (int i = 0; i < 32; ++i) {
      const float a = (float)pv(row, 2 * i), b = (float)pv(row, 2 * i + 1);
      if (variant == 0) {
        pp[i] = (uint32_t)__half_as_ushort(__float2half(a)) | ((uint32_t)__half_as_ushort(__float2half(b)) << 16);
      } else {
        pp[i] = pack_bf16x2(a, b);
      }
    }
    st32(tb + lane_off + 64, pp);
    tc_wait_st();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) {
      if (elect_one_sync()) {
        const uint32_t v_lo = umma_desc_lo(smem_u32(smV), 1024);
        const uint32_t id = idesc(128, 64, 1, variant == 0 ? 0 : 1, (mode & 16) ? 0 : 1, 0, 1);
        for (int k = 0; k < 4; ++k) umma_ts_lh(tb + 128, tb + 64 + k * 8, v_lo + k * 128, id, k != 0);
        tc_commit(&bar[1]);
      }
      __syncwarp();
    }
    mbar_wait(&bar[1], nb & 1); ++nb;
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld32(tb + lane_off + 128, o0);
    tmem_ld32(tb + lane_off + 160, o1);
    tmem_ld_wait32(o0); tmem_ld_wait32(o1);
    int bad = 0;
    for (int d = 0; d < 64; ++d) {
      int e = 0;
      for (int j = 0; j < 64; ++j) e += pv(row, j) * vv(j, d);
      const float got = __uint_as_float(d < 32 ? o0[d] : o1[d - 32]);
      bad += (got != (float)e);
      if (variant == 0 && row == 5 && d < 8) dump[d] = got, dump[8 + d] = (float)e;
    }
    atomicAdd(&out[8 + variant], bad);
    tc_fence_before(); __syncthreads(); tc_fence_after();
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 7;
  int* d; float* dump; cudaMalloc(&d, 64); cudaMalloc(&dump, 64); cudaMemset(d, 0, 64); cudaMemset(dump, 0, 64);
  probe<<<1, 128>>>(d, dump, mode);
  cudaError_t e = cudaDeviceSynchronize();
  int h[16]; float f[16]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost); cudaMemcpy(f, dump, 64, cudaMemcpyDeviceToHost);
  printf("mode %d status: %s\n", mode, cudaGetErrorString(e));
  printf("S fp16 accumulators (8192 elements): mismatches pack16[lo=even col] %d, pack16[lo=odd col] %d, plain low half (first 32 cols) %d\n", h[0], h[1], h[2]);
  printf("raw 32-bit words row 5 cols 0..3: %08x %08x %08x %08x\n", h[3], h[4], h[5], h[6]);
  printf("O = P V (8192 elements): mismatches A=f16 x B=bf16 %d, A=bf16 x B=bf16 %d\n", h[8], h[9]);
  printf("row 5 got:"); for (int i = 0; i < 8; ++i) printf(" %g", f[i]); printf("\nrow 5 exp:"); for (int i = 0; i < 8; ++i) printf(" %g", f[8 + i]); printf("\n");
  return 0;
}
