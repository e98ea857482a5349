// Micro-benchmark of the per-SM throughput of the instructions the softmax / GELU inner loops are made of
// (development tool, not part of the product):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench_mufu.cu -o gpurun_out/ubench_mufu
// Each test runs W warps per SM (one CTA per SM) executing ITER x 16 independent copies of one instruction;
// the result is warp-instructions / clk / SM and elements / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

constexpr int ITER = 2048;
constexpr int U = 16;

enum Op { EX2_F32, EX2_F16X2, EX2_BF16X2, TANH_F32, TANH_F16X2, CVT_F16X2, CVT_BF16X2, FMNMX, HMNMX2, FFMA, HFMA2,
          EX2_F32_PLUS_FFMA, EX2_F16X2_MIX };

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, long long* clk, uint32_t seed) {
  uint32_t r[U];
#pragma unroll
  for (int i = 0; i < U; ++i) r[i] = seed + threadIdx.x * 17 + i * 3;
  float f2[U];
#pragma unroll
  for (int i = 0; i < U; ++i) f2[i] = __uint_as_float(0x3f000000u + i + threadIdx.x);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if constexpr (OP == EX2_F32) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
      if constexpr (OP == EX2_F16X2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if constexpr (OP == EX2_BF16X2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if constexpr (OP == TANH_F32) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(r[i]));
      if constexpr (OP == TANH_F16X2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if constexpr (OP == CVT_F16X2) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f2[i]), "f"(__uint_as_float(r[i])));
      if constexpr (OP == CVT_BF16X2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f2[i]), "f"(__uint_as_float(r[i])));
      if constexpr (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f2[i]) : "f"(__uint_as_float(r[i])));
      if constexpr (OP == HMNMX2) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(__float_as_uint(f2[i])));
      if constexpr (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f2[i]) : "f"(__uint_as_float(r[i])));
      if constexpr (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(r[i]) : "r"(__float_as_uint(f2[i])));
      if constexpr (OP == EX2_F32_PLUS_FFMA) {  // softmax-like mix: FFMA -> MUFU -> FADD
        asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f2[i]) : "f"(__uint_as_float(r[i])));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(r[i]) : "f"(f2[i]));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(f2[(i + 1) % U]) : "f"(__uint_as_float(r[i])));
      }
      if constexpr (OP == EX2_F16X2_MIX) {  // packed variant: 2 FFMA + cvt.f16x2 + max.f16x2 + ex2.f16x2 per 2 elements
        asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f2[i]) : "f"(__uint_as_float(r[i])));
        asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f2[(i + 1) % U]) : "f"(__uint_as_float(r[i])));
        uint32_t pk;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(f2[i]), "f"(f2[(i + 1) % U]));
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[(i + 2) % U]) : "r"(pk));
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r[i]) : "r"(pk));
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < U; ++i) acc ^= r[i] ^ __float_as_uint(f2[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int elems_per_instr, int instr_per_iter) {
  uint32_t* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  for (int warps : {4, 8, 16, 32}) {
    k<OP><<<148, warps * 32>>>(out, clk, 12345);
    cudaDeviceSynchronize();
    k<OP><<<148, warps * 32>>>(out, clk, 12345);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double winstr = double(warps) * ITER * U * instr_per_iter;
    printf("%-20s warps/SM=%2d : %.3f warp-instr/clk/SM, %.2f elements/clk/SM, %.2f clk per warp-iter-group %s\n", name, warps,
           winstr / mx, double(warps) * ITER * U * 32 * elems_per_instr / mx, mx / (ITER * U),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<EX2_F32>("ex2.f32", 1, 1);
  run<EX2_F16X2>("ex2.f16x2", 2, 1);
  run<EX2_BF16X2>("ex2.bf16x2", 2, 1);
  run<TANH_F32>("tanh.f32", 1, 1);
  run<TANH_F16X2>("tanh.f16x2", 2, 1);
  run<CVT_F16X2>("cvt.f16x2.f32", 2, 1);
  run<CVT_BF16X2>("cvt.bf16x2.f32", 2, 1);
  run<FMNMX>("max.f32", 1, 1);
  run<HMNMX2>("max.f16x2", 2, 1);
  run<FFMA>("fma.f32", 1, 1);
  run<HFMA2>("fma.f16x2", 2, 1);
  run<EX2_F32_PLUS_FFMA>("ffma+ex2.f32+fadd", 1, 3);
  run<EX2_F16X2_MIX>("packed softmax mix", 2, 5);
  return 0;
}
