// Throughput / latency of the packed fp32x2 instructions (FFMA2, FADD2) vs scalar FFMA / FADD (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench_f32x2.cu -o tools/ubench_bin_f32x2
#include <cstdio>
#include <cstdint>
constexpr int ITER = 4096;
template <int OP, int CHAINS>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* clk, float seed) {
  float a[2 * CHAINS];
#pragma unroll
  for (int i = 0; i < 2 * CHAINS; ++i) a[i] = seed + threadIdx.x + i;
  const float b = seed * 0.5f, c = seed * 0.25f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) {  // FFMA2 (dependent per chain)
        uint64_t d;
        asm volatile("{.reg .b64 x, y, z;\n mov.b64 x, {%1, %2};\n mov.b64 y, {%3, %3};\n mov.b64 z, {%4, %4};\n fma.rn.f32x2 %0, x, y, z;}\n"
                     : "=l"(d) : "f"(a[2 * i]), "f"(a[2 * i + 1]), "f"(b), "f"(c));
        a[2 * i] = __uint_as_float((uint32_t)d); a[2 * i + 1] = __uint_as_float((uint32_t)(d >> 32));
      } else if (OP == 1) {  // FADD2
        uint64_t d;
        asm volatile("{.reg .b64 x, y;\n mov.b64 x, {%1, %2};\n mov.b64 y, {%3, %3};\n add.rn.f32x2 %0, x, y;}\n"
                     : "=l"(d) : "f"(a[2 * i]), "f"(a[2 * i + 1]), "f"(b));
        a[2 * i] = __uint_as_float((uint32_t)d); a[2 * i + 1] = __uint_as_float((uint32_t)(d >> 32));
      } else if (OP == 2) {  // 2 x FFMA
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[2 * i]) : "f"(b), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[2 * i + 1]) : "f"(b), "f"(c));
      } else {  // 2 x FADD
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[2 * i]) : "f"(b));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[2 * i + 1]) : "f"(b));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * CHAINS; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP, int CHAINS>
void run(const char* name) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  for (int warps : {1, 4, 8, 16}) {
    k<OP, CHAINS><<<148, warps * 32>>>(out, clk, 1.0001f); cudaDeviceSynchronize();
    k<OP, CHAINS><<<148, warps * 32>>>(out, clk, 1.0001f); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-10s chains=%2d warps/SM=%2d : %.2f clk per step per warp (each step = %d fp32 elements/lane), %.1f fp32 lane-ops/clk/SM\n", name, CHAINS, warps,
           mx / ITER / CHAINS, 2, double(warps) * 32 * 2 * CHAINS * ITER / mx);
  }
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<0, 1>("FFMA2"); run<0, 8>("FFMA2"); run<1, 1>("FADD2"); run<1, 8>("FADD2");
  run<2, 1>("2xFFMA"); run<2, 8>("2xFFMA"); run<3, 1>("2xFADD"); run<3, 8>("2xFADD");
  return 0;
}
