// Tensor-pipe throughput of tcgen05.mma shapes used by the attention kernel (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I crossscore_b200/csrc tools/ubench_mma_rate.cu -o tools/ubench_bin_mma
// A converged warp issues (elect.sync, uniform operands) a long stream of MMAs; reports clk per MMA to completion.
//   SS: A and B from shared memory (K-major, 128B swizzle);  TS: A from TMEM, B from smem (MN-major)
// CTAS = co-resident CTAs per SM (each with its own TMEM slice), to see how two CTAs share the pipe.
#include <cstdio>
#include "xs_common.cuh"
namespace xs { void set_last_error(const char*, ...) {} int num_sms() { return 148; }
int make_tmap(CUtensorMap*, const void*, int, int, const uint64_t*, const uint64_t*, const uint32_t*, Swizzle) { return 0; } }
using namespace xs;

template <int N, int MODE>  // MODE 0: SS, 1: TS, 2: alternate 4 SS (QK-like, fresh accumulator) + 4 TS (PV-like)
__global__ void __launch_bounds__(128) ubench(int n_mma, long long* out, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int SM_BYTES = 64 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM_BYTES);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < SM_BYTES / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(slot, tmem_cols);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  if (warp == 1) {
    const uint32_t tbu = warp_uniform(tb);
    constexpr uint32_t idesc_ss = umma_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t idesc_ts = umma_idesc_bf16(128, N, 0, 1);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16);
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem + 16 * 1024), 16);
    const uint32_t v_lo = umma_desc_lo(smem_u32(smem + 16 * 1024), 1024);
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
      if (elect_one_sync()) {
        if (MODE == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss_lh<false>(tbu, a_lo + 2 * (k & 3), b_lo + 2 * (k & 3), idesc_ss, 1);
        } else if (MODE == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ts_lh(tbu, tbu + N + (k & 3) * 8, v_lo + (k & 3) * 128, idesc_ts, 1);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_lh<false>(tbu, a_lo + 2 * k, b_lo + 2 * k, idesc_ss, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ts_lh(tbu + N, tbu + k * 8, v_lo + k * 128, idesc_ts, 1);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one_sync()) tc_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 32) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, tmem_cols); }
}

template <int N, int MODE>
void run(const char* name, int ctas_per_sm) {
  long long* d; cudaMalloc(&d, 148 * 4 * 16);
  auto k = ubench<N, MODE>;
  const int smem = 64 * 1024 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 4096;
  const int grid = 148 * ctas_per_sm;
  const int cols = ctas_per_sm == 1 ? 512 : 256;
  k<<<grid, 128, smem>>>(n, d, cols);
  cudaDeviceSynchronize();
  k<<<grid, 128, smem>>>(n, d, cols);
  cudaError_t e = cudaDeviceSynchronize();
  static long long h[148 * 4 * 2];
  cudaMemcpy(h, d, grid * 16, cudaMemcpyDeviceToHost);
  double mx = 0, mi = 0;
  for (int i = 0; i < grid; ++i) { mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx; mi = h[2 * i] > mi ? h[2 * i] : mi; }
  printf("%-6s N=%3d CTAs/SM=%d : issue %.1f, complete %.1f clk/MMA per CTA -> %.1f clk/MMA per SM (ideal %d) %s\n", name, N, ctas_per_sm,
         mi / n, mx / n, mx / n / ctas_per_sm, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int c : {1, 2}) {
    run<64, 0>("SS", c); run<128, 0>("SS", c); run<256, 0>("SS", c);
    run<48, 1>("TS", c); run<64, 1>("TS", c); run<128, 1>("TS", c);
    run<64, 2>("QK+PV", c); run<128, 2>("QK+PV", c);
  }
  return 0;
}
