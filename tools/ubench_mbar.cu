// mbarrier hand-off latency between two warps of a CTA (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I crossscore_b200/csrc tools/ubench_mbar.cu -o tools/ubench_bin_mbar
// Warp 0 arrives on A and waits on B; warp 1 waits on A and arrives on B: one iteration = two hand-offs.
// MODE 0: try_wait (default suspend) loop; 1: try_wait with 20 us suspend hint; 2: test_wait spin; 3: test_wait spin
// with the waiting done by lane 0 only + __syncwarp; 4: like 0 but 6 extra warps per CTA run FFMA loops (busy SM).
// MODE 5: tcgen05.commit -> mbarrier -> waiting warp (MMA completion hand-off): warp 0 issues one tiny MMA + commit per
// iteration and warp 1 waits for it and arrives back.
#include <cstdio>
#include "xs_common.cuh"
namespace xs { void set_last_error(const char*, ...) {} int num_sms() { return 148; }
int make_tmap(CUtensorMap*, const void*, int, int, const uint64_t*, const uint64_t*, const uint32_t*, Swizzle) { return 0; } }
using namespace xs;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int MODE>
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity, int lane) {
  if (MODE == 0 || MODE == 4 || MODE == 5 || MODE == 6) { while (!mbar_try_wait(bar, parity)) {} }
  else if (MODE == 1) { while (!mbar_try_wait_hint(bar, parity)) {} }
  else if (MODE == 2) { while (!mbar_test_wait(bar, parity)) {} }
  else { if (lane == 0) { while (!mbar_test_wait(bar, parity)) {} } __syncwarp(); }
}

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, long long* out, float* sink) {
  __shared__ __align__(1024) uint8_t tile[16384 + 2048];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], MODE >= 5 ? 1 : 32); mbar_init(&bars[1], 32); fence_mbar_init(); }
  for (int i = threadIdx.x; i < (16384 + 2048) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tile)[i] = 0;
  if (MODE >= 5 && warp == 0) tmem_alloc(&slot, 32);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  float acc = threadIdx.x;
  if (warp == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (MODE >= 5) {
        if (elect_one_sync()) {
          if (MODE == 5)
            umma_ss_lh<false>(slot, umma_desc_lo(smem_u32(tile), 16), umma_desc_lo(smem_u32(tile + 16384), 16), umma_idesc_bf16(128, 16, 0, 0), 0);
          tc_commit(&bars[0]);
        }
        __syncwarp();
      } else {
        mbar_arrive(&bars[0]);
      }
      wait<MODE>(&bars[1], i & 1, lane);
    }
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 1) {
    for (int i = 0; i < iters; ++i) {
      wait<MODE>(&bars[0], i & 1, lane);
      if (MODE >= 5) tc_fence_after();
      mbar_arrive(&bars[1]);
    }
  } else if (MODE == 4) {
    for (int i = 0; i < iters * 200; ++i) acc = fmaf(acc, 1.0001f, 0.5f);
  }
  if (acc == 12345.678f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (MODE >= 5 && warp == 0) { tc_fence_after(); tmem_dealloc(slot, 32); }
}

template <int MODE>
void run(const char* name, int threads) {
  long long* d; float* sink; cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(iters, d, sink); cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(iters, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0, mn = 1e30; for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  printf("%-44s: %.0f .. %.0f clk per round trip (2 hand-offs) %s\n", name, mn / iters, mx / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}
int main() {
  run<0>("try_wait loop", 64);
  run<1>("try_wait + 20us suspend hint", 64);
  run<2>("test_wait spin (all lanes)", 64);
  run<3>("test_wait spin (lane 0) + syncwarp", 64);
  run<4>("try_wait loop, 6 busy FFMA warps", 256);
  run<5>("1 MMA + tcgen05.commit -> try_wait -> arrive -> try_wait", 64);
  run<6>("tcgen05.commit (no MMA) -> try_wait -> arrive -> try_wait", 64);
  return 0;
}
