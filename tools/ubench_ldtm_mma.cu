// Does tcgen05.ld bandwidth survive a busy tensor pipe?  (development tool)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I crossscore_b200/csrc tools/ubench_ldtm_mma.cu -o tools/ubench_bin_ldtm_mma
// One CTA per SM: warp 1 streams tcgen05.mma (M=128, N=NMMA, SS) into TMEM columns [256, 256+NMMA) while W warps
// (4..4+W-1) stream tcgen05.ld 32x32b.x32 pairs from columns [0, 128).  Reports both rates, alone and together.
#include <cstdio>
#include "xs_common.cuh"
namespace xs { void set_last_error(const char*, ...) {} int num_sms() { return 148; }
int make_tmap(CUtensorMap*, const void*, int, int, const uint64_t*, const uint64_t*, const uint32_t*, Swizzle) { return 0; } }
using namespace xs;

template <int NMMA>
__global__ void __launch_bounds__(384) k(int n_mma, int ld_iters, int ld_warps, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  float acc = 0.f;
  if (warp == 1 && n_mma > 0) {
    const uint32_t tbu = warp_uniform(tb);
    constexpr uint32_t idesc = umma_idesc_bf16(128, NMMA, 0, 0);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16);
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem + 16 * 1024), 16);
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss_lh<false>(tbu + 256, a_lo + 2 * (kk & 3), b_lo + 2 * (kk & 3), idesc, 1);
      }
      __syncwarp();
    }
    if (elect_one_sync()) tc_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x * 16] = t1 - t0;
  } else if (warp >= 4 && warp < 4 + ld_warps && ld_iters > 0) {
    const uint32_t t = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t va[32], vb[32];
    const long long t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) {
      const uint32_t col = (i & 1) * 64;
      tmem_ld32(t + col, va); tmem_ld32(t + col + 32, vb);
      tmem_ld_wait32(va); tmem_ld_wait32(vb);
      acc += __uint_as_float(va[i & 31]) + __uint_as_float(vb[i & 31]);
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp - 3] = t1 - t0;
  }
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int NMMA>
void run(int n_mma, int ld_iters, int ld_warps) {
  long long* d; float* sink; cudaMalloc(&d, 148 * 16 * 8); cudaMalloc(&sink, 4);
  cudaMemset(d, 0, 148 * 16 * 8);
  auto kern = k<NMMA>;
  const int smem = 64 * 1024 + 1024 + 64;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  kern<<<148, 384, smem>>>(n_mma, ld_iters, ld_warps, d, sink); cudaDeviceSynchronize();
  cudaMemset(d, 0, 148 * 16 * 8);
  kern<<<148, 384, smem>>>(n_mma, ld_iters, ld_warps, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  static long long h[148 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double mma = 0, ld = 0;
  for (int b = 0; b < 148; ++b) { mma = h[b * 16] > mma ? h[b * 16] : mma; for (int w = 1; w <= ld_warps; ++w) ld = h[b * 16 + w] > ld ? h[b * 16 + w] : ld; }
  printf("N=%3d  mma stream %6d  ld warps %d x %5d blocks : ", NMMA, n_mma, ld_warps, ld_iters);
  if (n_mma) printf("MMA %.1f clk each (ideal %d)  ", mma / n_mma, 128 * NMMA / 256);
  if (ld_iters) printf("LDTM %.1f B/clk/SM", 32.0 * 64 * 4 * ld_warps * ld_iters / ld);
  printf(" %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}
int main() {
  run<256>(8192, 0, 0);
  run<256>(0, 2048, 8);
  run<256>(16384, 2048, 8);   // sized so both streams run ~ the same time
  run<256>(16384, 2048, 4);
  run<64>(32768, 0, 0);
  run<64>(32768, 2048, 8);
  return 0;
}
