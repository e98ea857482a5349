// Micro-benchmark of tcgen05.mma issue/latency behaviour on B200 (development tool, not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I crossscore_b200/csrc tools/ubench_tcgen05.cu -o gpurun_out/ubench
#include <cstdio>
#include "xs_common.cuh"
namespace xs { void set_last_error(const char*, ...) {} int num_sms() { return 148; }
int make_tmap(CUtensorMap*, const void*, int, int, const uint64_t*, const uint64_t*, const uint32_t*, Swizzle) { return 0; } }
using namespace xs;

// mode 0: dependent chain into one accumulator; 1: alternate 2 accumulators; 2: alternate 4 accumulators
// commit_every: issue a tcgen05.commit (to a scratch barrier) after every c MMAs (0 = only at the end)
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) ubench(int n_mma, int mode, int commit_every, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint64_t* scratch = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(scratch, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
    const int nacc = mode == 0 ? 1 : (mode == 1 ? 2 : 4);
    // warm-up
    umma_ss(tb, umma_desc_sw128(a_addr, 16, 1024), umma_desc_sw128(b_addr, 16, 1024), idesc, 0);
    tc_commit(bar); mbar_wait(bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tb + (i % nacc) * (N <= 128 ? 128 : 256) % 512;
      const int k = i & 3;
      if (TS) umma_ts(d, tb + 448 + k * 8, umma_desc_sw128(b_addr + k * 2048, 1024, 1024), umma_idesc_bf16(128, N, 0, 1), 1);
      else umma_ss(d, umma_desc_sw128(a_addr + k * 32, 16, 1024), umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc, 1);
      if (commit_every > 0 && (i + 1) % commit_every == 0) tc_commit(scratch);
    }
    const long long t1 = clock64();
    tc_commit(bar); mbar_wait(bar, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, bool TS>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 16);
  auto k = ubench<N, TS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n = 256;
  for (int mode = 0; mode < 3; ++mode)
    for (int ce : {0, 4, 1}) {
      k<<<grid, 128, 100 * 1024>>>(n, mode, ce, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%s N=%3d grid=%3d accs=%d commit_every=%d : issue %.1f clk/mma, complete %.1f clk/mma (ideal %d) %s\n", name, N, grid,
             mode == 0 ? 1 : (mode == 1 ? 2 : 4), ce, double(h[0]) / n, double(h[1]) / n, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  cudaFree(d);
}
int main() {
  run<64, false>("SS", 1);  run<128, false>("SS", 1); run<256, false>("SS", 1);
  run<64, true>("TS", 1);   run<128, true>("TS", 1);
  run<64, false>("SS", 148); run<128, false>("SS", 148); run<64, true>("TS", 148);
  return 0;
}
