// TMEM load/store throughput as the attention softmax warps use it (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I crossscore_b200/csrc tools/ubench_ldtm.cu -o tools/ubench_bin_ldtm
// W warps per CTA (warp w reads lane quarter w%4), C CTAs per SM; every warp streams ITER x (2 x tcgen05.ld 32x32b.x32)
// = one 64-column fp32 S block per iteration.  MODE 0: ld + wait per block; 1: two loads in flight (wait once per block);
// 2: ld + 64 ex2 per thread per block (the softmax mix: does LDTM overlap MUFU?); 3: st x16 x2 + wait::st per block.
#include <cstdio>
#include "xs_common.cuh"
namespace xs { void set_last_error(const char*, ...) {} int num_sms() { return 148; }
int make_tmap(CUtensorMap*, const void*, int, int, const uint64_t*, const uint64_t*, const uint32_t*, Swizzle) { return 0; } }
using namespace xs;

template <int MODE>
__global__ void __launch_bounds__(512) ubench(int iters, long long* out, int tmem_cols, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, tmem_cols);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t va[32], vb[32];
  float acc = 0.f;
  const long long t0 = clock64();
  if (MODE == 1) tmem_ld32(tb, va);
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (i & 1) * 64;
    if (MODE == 0) {
      tmem_ld32(tb + col, va); tmem_ld32(tb + col + 32, vb);
      tmem_ld_wait32(va); tmem_ld_wait32(vb);
      acc += __uint_as_float(va[i & 31]) + __uint_as_float(vb[i & 31]);
    } else if (MODE == 1) {
      tmem_ld_wait32(va); tmem_ld32(tb + col + 32, vb);
      acc += __uint_as_float(va[i & 31]);
      tmem_ld_wait32(vb); tmem_ld32(tb + (col ^ 64), va);
      acc += __uint_as_float(vb[i & 31]);
    } else if (MODE == 2) {
      tmem_ld32(tb + col, va); tmem_ld32(tb + col + 32, vb);
      tmem_ld_wait32(va); tmem_ld_wait32(vb);
#pragma unroll
      for (int k = 0; k < 32; ++k) acc += fast_exp2(__uint_as_float(va[k]) * 1e-9f) + fast_exp2(__uint_as_float(vb[k]) * 1e-9f);
    } else {
      uint32_t pk[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) pk[k] = i + k;
      tmem_st16(tb + col, pk); tmem_st16(tb + col + 16, pk);
      tc_wait_st();
    }
  }
  if (MODE == 1) tmem_ld_wait32(va);
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc + __uint_as_float(va[0]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, tmem_cols); }
}

template <int MODE>
void run(const char* name, int warps, int ctas) {
  long long* d; float* sink; cudaMalloc(&d, 148 * 2 * 16 * 8); cudaMalloc(&sink, 4);
  const int iters = 4096, grid = 148 * ctas, cols = ctas == 1 ? 512 : 256;
  ubench<MODE><<<grid, warps * 32>>>(iters, d, cols, sink);
  cudaDeviceSynchronize();
  ubench<MODE><<<grid, warps * 32>>>(iters, d, cols, sink);
  cudaError_t e = cudaDeviceSynchronize();
  static long long h[148 * 2 * 16];
  cudaMemcpy(h, d, grid * 16 * 8, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int b = 0; b < grid; ++b) for (int w = 0; w < warps; ++w) mx = h[b * 16 + w] > mx ? h[b * 16 + w] : mx;
  const double per_blk = mx / iters;  // clk per warp-iteration (32 rows x 64 cols)
  const double bytes = MODE == 3 ? 32.0 * 32 * 4 : 32.0 * 64 * 4;
  printf("%-28s warps/CTA=%2d CTAs/SM=%d : %.1f clk per warp-block, %.1f B/clk/SM %s\n", name, warps, ctas, per_blk,
         bytes * warps * ctas / per_blk, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}

int main() {
  for (int c : {1, 2}) for (int w : {1, 4, 8}) run<0>("ld x32 x2 + wait", w, c);
  for (int c : {1, 2}) for (int w : {4, 8}) run<1>("ld pipelined", w, c);
  for (int c : {1, 2}) for (int w : {4, 8}) run<2>("ld + 64 ex2", w, c);
  for (int c : {1, 2}) for (int w : {4, 8}) run<3>("st x16 x2 + wait::st", w, c);
  return 0;
}
