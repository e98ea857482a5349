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


__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_16x128b_x16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x128b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[32], int o) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]),
                 "=r"(r[o + 7])
               : "r"(taddr)
               : "memory");
}

// 32x32b.x32 with .pack::16b: 64 TMEM columns (low 16 bits of each) -> 32 registers
__device__ __forceinline__ void ld_32x32b_x32_pack16(uint32_t taddr, uint32_t (&r)[32]) {
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
    } else if (MODE == 4) {  // 16x256b.x8: 16 lanes x 64 columns per instruction, two instructions per 32-lane block
      ld_16x256b_x8(tb + col, va); ld_16x256b_x8(tb + col + (16u << 16), vb);
      tmem_ld_wait32(va); tmem_ld_wait32(vb);
      acc += __uint_as_float(va[i & 31]) + __uint_as_float(vb[i & 31]);
    } else if (MODE == 5) {  // 16x128b.x16
      ld_16x128b_x16(tb + col, va); ld_16x128b_x16(tb + col + (16u << 16), vb);
      tmem_ld_wait32(va); tmem_ld_wait32(vb);
      acc += __uint_as_float(va[i & 31]) + __uint_as_float(vb[i & 31]);
    } else if (MODE == 6) {  // 8 x (32x32b.x8)
#pragma unroll
      for (int c = 0; c < 4; ++c) { ld_32x32b_x8(tb + col + 8 * c, va, 8 * c); ld_32x32b_x8(tb + col + 32 + 8 * c, vb, 8 * c); }
      tmem_ld_wait32(va); tmem_ld_wait32(vb);
      acc += __uint_as_float(va[i & 31]) + __uint_as_float(vb[i & 31]);
    } else if (MODE == 8) {  // one packed load covers the whole 64-column block (are 16-bit S accumulators cheaper to read?)
      ld_32x32b_x32_pack16(tb + col, va);
      tmem_ld_wait32(va);
      acc += __uint_as_float(va[i & 31]);
    } else if (MODE == 9) {  // packed load + 64 ex2
      ld_32x32b_x32_pack16(tb + col, va);
      tmem_ld_wait32(va);
#pragma unroll
      for (int k = 0; k < 32; ++k) acc += fast_exp2(__uint_as_float(va[k]) * 1e-9f) + fast_exp2(__uint_as_float(va[k] << 16) * 1e-9f);
    } else if (MODE == 7) {  // 16x256b + 64 ex2
      ld_16x256b_x8(tb + col, va); ld_16x256b_x8(tb + col + (16u << 16), vb);
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
  for (int c : {1, 2}) for (int w : {1, 4, 8}) run<4>("ld 16x256b.x8 x2 + wait", w, c);
  for (int c : {1, 2}) for (int w : {4, 8}) run<5>("ld 16x128b.x16 x2 + wait", w, c);
  for (int c : {1, 2}) for (int w : {4, 8}) run<6>("ld 32x32b.x8 x8 + wait", w, c);
  for (int c : {1, 2}) for (int w : {4, 8}) run<7>("ld 16x256b + 64 ex2", w, c);
  for (int c : {1, 2}) for (int w : {1, 4, 8}) run<8>("ld 32x32b.x32.pack16 (64 col)", w, c);
  for (int c : {1, 2}) for (int w : {4, 8}) run<9>("ld pack16 + 64 ex2", w, c);
  return 0;
}
