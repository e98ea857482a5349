// Image pre-processing in front of the model and score-map post-processing behind it.
//
// preprocess:  uint8 HWC image -> /255 -> antialiased bilinear resize -> ImageNet normalise -> fp32 NCHW
//   replaces the dataloader's per-image CPU work: utils/io/images.py:14-29 (image_read / f32),
//   dataloading/dataset/nvs_dataset.py:428-446 (permute), :218-225 (resize_all = torchvision
//   T.Resize(short side, BILINEAR, antialias=True), task/predict.py:87-92 -> ATen _upsample_bilinear2d_aa),
//   :242-279 + task/predict.py:69-74 (T.Normalize(ImageNet mean/std)).
// postprocess: fp32 score map -> per-frame mean (utils/io/score_summariser.py:180-181), uint16 gray quantisation
//   (utils/io/images.py:49-63, metric_map_write) and turbo RGB (utils/misc/image.py:35-49, gray2rgb; u8 at
//   utils/io/images.py:20-23), as written by utils/io/batch_writer.py:114-135,263-270.
//
// Both are HBM-bound byte/pixel work: one pass over the input, one over the output; the resize stages its input
// window in shared memory with coalesced loads and runs the two separable passes out of shared memory.
#include "xs_common.cuh"
#include "xs_turbo.h"

namespace xs {

// ---------------------------------------------------------------------------------------------------------------
// antialiased bilinear resize (triangle filter whose support scales with the down-scaling factor)
// ---------------------------------------------------------------------------------------------------------------
constexpr int RS_TW = 32;       // output tile width
constexpr int RS_TH = 16;       // output tile height
constexpr int RS_MAXTAP = 24;   // taps per output index: 2 * ceil(scale) + 2 (scale <= 11)
constexpr int RS_THREADS = 256;

struct ResizeParams {
  const uint8_t* img;  // (n, H0, W0, 3)
  float* out;          // (n, 3, H1, W1)
  int n, H0, W0, H1, W1;
  float scale_y, scale_x;  // in / out (fp32 division, as ATen's area_pixel_compute_scale)
  float mean[3], std[3];
  int in_rows_max, in_cols_max;  // bound of the input window of a tile (host-computed, sizes the shared memory)
  int raw_pitch;                 // bytes per staged row (multiple of 4)
  long long total_bytes;         // n * H0 * W0 * 3
};

// first input index and tap count of output index i (ATen _compute_indices_min_size_weights_aa, align_corners=False)
__device__ __forceinline__ void aa_window(int i, float scale, int in_size, int& xmin, int& xsize, float& center,
                                          float& invscale) {
  const float support = scale >= 1.0f ? scale : 1.0f;
  invscale = scale >= 1.0f ? __fdiv_rn(1.0f, scale) : 1.0f;
  center = __fmul_rn(scale, __fadd_rn(static_cast<float>(i), 0.5f));
  xmin = max(static_cast<int>(__fadd_rn(__fsub_rn(center, support), 0.5f)), 0);
  xsize = min(static_cast<int>(__fadd_rn(__fadd_rn(center, support), 0.5f)), in_size) - xmin;
}
// normalised weights of one output index into w[0..xsize)
__device__ __forceinline__ void aa_weights(int xmin, int xsize, float center, float invscale, float* w) {
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) {
    const float x = __fmul_rn(__fadd_rn(__fsub_rn(static_cast<float>(j + xmin), center), 0.5f), invscale);
    const float v = fmaxf(0.f, __fsub_rn(1.0f, fabsf(x)));
    w[j] = v;
    total = __fadd_rn(total, v);
  }
  if (total != 0.f) {
    for (int j = 0; j < xsize; ++j) w[j] = __fdiv_rn(w[j], total);
  }
}

__global__ void __launch_bounds__(RS_THREADS) resize_aa_norm_kernel(ResizeParams p) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  // layout: wx[RS_TW][MAXTAP] | wy[RS_TH][MAXTAP] | x0[RS_TW], nx[RS_TW], y0[RS_TH], ny[RS_TH] | lut[256] |
  //         rshift[in_rows_max] | tmp[in_rows_max][RS_TW][3] fp32 | raw[in_rows_max][raw_pitch] u8
  float* wx = reinterpret_cast<float*>(rs_smem);
  float* wy = wx + RS_TW * RS_MAXTAP;
  int* x0s = reinterpret_cast<int*>(wy + RS_TH * RS_MAXTAP);
  int* nxs = x0s + RS_TW;
  int* y0s = nxs + RS_TW;
  int* nys = y0s + RS_TH;
  float* lut = reinterpret_cast<float*>(nys + RS_TH);
  int* rshift = reinterpret_cast<int*>(lut + 256);  // [in_rows_max] byte offset of the window inside its first word
  float* tmp = reinterpret_cast<float*>(rshift + p.in_rows_max);
  uint8_t* raw = reinterpret_cast<uint8_t*>(tmp + static_cast<size_t>(p.in_rows_max) * RS_TW * 3);

  const int tid = threadIdx.x;
  const int img = blockIdx.z;
  const int ox0 = blockIdx.x * RS_TW, oy0 = blockIdx.y * RS_TH;
  const int tw = min(RS_TW, p.W1 - ox0), th = min(RS_TH, p.H1 - oy0);

  lut[tid] = __fdiv_rn(static_cast<float>(tid), 255.0f);  // utils/io/images.py:14-17, RS_THREADS == 256
  if (tid < tw) {
    int xmin, xsize;
    float center, inv;
    aa_window(ox0 + tid, p.scale_x, p.W0, xmin, xsize, center, inv);
    x0s[tid] = xmin;
    nxs[tid] = xsize;
    aa_weights(xmin, xsize, center, inv, wx + tid * RS_MAXTAP);
  } else if (tid >= 64 && tid < 64 + th) {
    const int t = tid - 64;
    int ymin, ysize;
    float center, inv;
    aa_window(oy0 + t, p.scale_y, p.H0, ymin, ysize, center, inv);
    y0s[t] = ymin;
    nys[t] = ysize;
    aa_weights(ymin, ysize, center, inv, wy + t * RS_MAXTAP);
  }
  __syncthreads();
  // input window of the tile (windows are monotone in the output index)
  const int xlo = x0s[0], xhi = x0s[tw - 1] + nxs[tw - 1];
  const int ylo = y0s[0], yhi = y0s[th - 1] + nys[th - 1];
  const int ncol = xhi - xlo, nrow = yhi - ylo;
  const int row_bytes = ncol * 3;
  const int raw_pitch = p.raw_pitch;
  // 1. stage the window: one warp per row, aligned 32-bit words (the window starts `shift` bytes into its first word)
  const long long off0 = (static_cast<long long>(img) * p.H0 + ylo) * p.W0 * 3 + static_cast<long long>(xlo) * 3;
  const long long full_words = p.total_bytes >> 2;  // words that lie entirely inside the image buffer
  for (int r = tid >> 5; r < nrow; r += RS_THREADS / 32) {
    const long long off = off0 + static_cast<long long>(r) * p.W0 * 3;
    const int shift = static_cast<int>(off & 3);
    const long long w0 = (off - shift) >> 2;
    const int nwords = (row_bytes + shift + 3) >> 2;
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(p.img) + w0;
    uint32_t* d4 = reinterpret_cast<uint32_t*>(raw + r * raw_pitch);
    for (int w = tid & 31; w < nwords; w += 32) {
      uint32_t v;
      if (w0 + w < full_words) {
        v = __ldg(s4 + w);
      } else {  // the last, partial word of the whole buffer: byte loads
        v = 0;
        for (int b = 0; b < 4; ++b) {
          const long long a = ((w0 + w) << 2) + b;
          if (a < p.total_bytes) v |= static_cast<uint32_t>(p.img[a]) << (8 * b);
        }
      }
      d4[w] = v;
    }
    if ((tid & 31) == 0) rshift[r] = shift;
  }
  __syncthreads();
  // 2. width pass: tmp[r][tx][c] = sum_k wx[tx][k] * lut[raw[r][x0 - xlo + k][c]]   (fp32, taps in ascending order)
  //    thread = (row r = tid / 32 + 8 i, column tx = tid % 32): no integer divisions, three channels per thread
  {
    const int tx = tid & (RS_TW - 1);
    if (tx < tw) {
      const int xs = (x0s[tx] - xlo) * 3, n = nxs[tx];
      const float* w = wx + tx * RS_MAXTAP;
      for (int r = tid >> 5; r < nrow; r += RS_THREADS / RS_TW) {
        const uint8_t* rp = raw + r * raw_pitch + rshift[r] + xs;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int k = 0; k < n; ++k) {
          const float wk = w[k];
          a0 = __fadd_rn(a0, __fmul_rn(wk, lut[rp[3 * k + 0]]));
          a1 = __fadd_rn(a1, __fmul_rn(wk, lut[rp[3 * k + 1]]));
          a2 = __fadd_rn(a2, __fmul_rn(wk, lut[rp[3 * k + 2]]));
        }
        float* t = tmp + (r * RS_TW + tx) * 3;
        t[0] = a0;
        t[1] = a1;
        t[2] = a2;
      }
    }
  }
  __syncthreads();
  // 3. height pass + normalise + NCHW store: thread = (ty = tid / 32 + 8 i, tx = tid % 32), three channels per
  //    thread (one weight load per tap), tx fastest so every warp writes 128 contiguous bytes per channel row
  {
    const int tx = tid & (RS_TW - 1);
    if (tx < tw) {
      const float m0 = p.mean[0], m1 = p.mean[1], m2 = p.mean[2], s0 = p.std[0], s1 = p.std[1], s2 = p.std[2];
      const size_t plane = static_cast<size_t>(p.H1) * p.W1;
      for (int ty = tid >> 5; ty < th; ty += RS_THREADS / RS_TW) {
        const int ys = y0s[ty] - ylo, n = nys[ty];
        const float* w = wy + ty * RS_MAXTAP;
        const float* t = tmp + (ys * RS_TW + tx) * 3;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int k = 0; k < n; ++k) {
          const float wk = w[k];
          a0 = __fadd_rn(a0, __fmul_rn(wk, t[k * RS_TW * 3 + 0]));
          a1 = __fadd_rn(a1, __fmul_rn(wk, t[k * RS_TW * 3 + 1]));
          a2 = __fadd_rn(a2, __fmul_rn(wk, t[k * RS_TW * 3 + 2]));
        }
        // torchvision Normalize: sub_(mean).div_(std)
        float* o = p.out + static_cast<size_t>(img) * 3 * plane + static_cast<size_t>(oy0 + ty) * p.W1 + ox0 + tx;
        o[0] = __fdiv_rn(__fsub_rn(a0, m0), s0);
        o[plane] = __fdiv_rn(__fsub_rn(a1, m1), s1);
        o[2 * plane] = __fdiv_rn(__fsub_rn(a2, m2), s2);
      }
    }
  }
}

// same-size path: no filtering at all (T.Resize is skipped / an identity), one thread per output element
__global__ void __launch_bounds__(256) u8_norm_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int n,
                                                      int H, int W, float m0, float m1, float m2, float s0, float s1,
                                                      float s2) {
  const long long total = static_cast<long long>(n) * 3 * H * W;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int c = static_cast<int>((idx / (static_cast<long long>(W) * H)) % 3);
    const long long im = idx / (3LL * W * H);
    const float v = __fdiv_rn(static_cast<float>(img[((im * H + y) * W + x) * 3 + c]), 255.0f);
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    out[idx] = __fdiv_rn(__fsub_rn(v, mean), sd);
  }
}

// same-size path, 4 pixels per thread: three aligned words in, one float4 per channel plane out (H*W % 4 == 0)
__global__ void __launch_bounds__(256) u8_norm_vec4_kernel(const uint32_t* __restrict__ img, float* __restrict__ out,
                                                           long long n_img, long long hw, float m0, float m1, float m2,
                                                           float s0, float s1, float s2) {
  const long long groups = n_img * (hw >> 2);
  for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long im = g / (hw >> 2);
    const long long pix = (g - im * (hw >> 2)) << 2;
    const uint32_t* s = img + (im * hw + pix) * 3 / 4;
    const uint32_t w0 = __ldg(s), w1 = __ldg(s + 1), w2 = __ldg(s + 2);
    const uint8_t b[12] = {uint8_t(w0), uint8_t(w0 >> 8), uint8_t(w0 >> 16), uint8_t(w0 >> 24),
                           uint8_t(w1), uint8_t(w1 >> 8), uint8_t(w1 >> 16), uint8_t(w1 >> 24),
                           uint8_t(w2), uint8_t(w2 >> 8), uint8_t(w2 >> 16), uint8_t(w2 >> 24)};
    auto nrm = [](uint8_t v, float mean, float sd) {
      return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), mean), sd);
    };
    float* o = out + im * 3 * hw + pix;
    *reinterpret_cast<float4*>(o) = make_float4(nrm(b[0], m0, s0), nrm(b[3], m0, s0), nrm(b[6], m0, s0), nrm(b[9], m0, s0));
    *reinterpret_cast<float4*>(o + hw) =
        make_float4(nrm(b[1], m1, s1), nrm(b[4], m1, s1), nrm(b[7], m1, s1), nrm(b[10], m1, s1));
    *reinterpret_cast<float4*>(o + 2 * hw) =
        make_float4(nrm(b[2], m2, s2), nrm(b[5], m2, s2), nrm(b[8], m2, s2), nrm(b[11], m2, s2));
  }
}

static inline int aa_span(int in_size, int out_size, int tile) {
  // upper bound of the input window of `tile` consecutive output indices
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  const float support = scale >= 1.0f ? scale : 1.0f;
  return static_cast<int>(scale * tile + 2.0f * support) + 3;
}

int preprocess_u8(const uint8_t* img, int n, int H0, int W0, float* out, int H1, int W1, const float* mean_std,
                  cudaStream_t stream) {
  XS_CHECK_ARG(n > 0 && H0 > 0 && W0 > 0 && H1 > 0 && W1 > 0, "preprocess: empty problem");
  XS_CHECK_ARG(mean_std != nullptr, "preprocess: mean_std (6 floats, host memory) is required");
  if (H0 == H1 && W0 == W1) {
    const long long cap = static_cast<long long>(num_sms()) * 32;
    const long long hw = static_cast<long long>(H0) * W0;
    if ((hw & 3) == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      const long long groups = static_cast<long long>(n) * (hw >> 2);
      const long long gv = (groups + 255) / 256;
      u8_norm_vec4_kernel<<<static_cast<int>(gv < cap ? gv : cap), 256, 0, stream>>>(
          reinterpret_cast<const uint32_t*>(img), out, n, hw, mean_std[0], mean_std[1], mean_std[2], mean_std[3],
          mean_std[4], mean_std[5]);
      XS_LAUNCH_CHECK();
      return 0;
    }
    const long long total = static_cast<long long>(n) * 3 * H0 * W0;
    long long g = (total + 255) / 256;
    u8_norm_kernel<<<static_cast<int>(g < cap ? g : cap), 256, 0, stream>>>(img, out, n, H0, W0, mean_std[0], mean_std[1],
                                                                         mean_std[2], mean_std[3], mean_std[4],
                                                                         mean_std[5]);
    XS_LAUNCH_CHECK();
    return 0;
  }
  ResizeParams p;
  p.img = img;
  p.out = out;
  p.n = n;
  p.H0 = H0;
  p.W0 = W0;
  p.H1 = H1;
  p.W1 = W1;
  p.scale_y = static_cast<float>(H0) / static_cast<float>(H1);
  p.scale_x = static_cast<float>(W0) / static_cast<float>(W1);
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean_std[c];
    p.std[c] = mean_std[3 + c];
  }
  const float smax = p.scale_x > p.scale_y ? p.scale_x : p.scale_y;
  XS_CHECK_ARG(2 * static_cast<int>(smax + 1.0f) + 2 <= RS_MAXTAP, "preprocess: down-scaling factor %.2f too large (max 10)",
               smax);
  p.in_rows_max = aa_span(H0, H1, RS_TH);
  p.in_cols_max = aa_span(W0, W1, RS_TW);
  p.raw_pitch = (p.in_cols_max * 3 + 3 + 3) & ~3;
  p.total_bytes = static_cast<long long>(n) * H0 * W0 * 3;
  XS_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 3) == 0, "preprocess: image pointer must be 4-byte aligned");
  const size_t smem = sizeof(float) * (RS_TW * RS_MAXTAP + RS_TH * RS_MAXTAP + 256) + sizeof(int) * 2 * (RS_TW + RS_TH) +
                     sizeof(int) * p.in_rows_max + sizeof(float) * static_cast<size_t>(p.in_rows_max) * RS_TW * 3 +
                     static_cast<size_t>(p.in_rows_max) * p.raw_pitch + 16;
  XS_CHECK_ARG(smem <= 200 * 1024, "preprocess: input window of one tile needs %zu bytes of shared memory", smem);
  XS_CUDA(cudaFuncSetAttribute(resize_aa_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  // many small blocks per SM: ask for the largest shared-memory carve-out so that threads, not shared memory, cap occupancy
  XS_CUDA(cudaFuncSetAttribute(resize_aa_norm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  dim3 grid((W1 + RS_TW - 1) / RS_TW, (H1 + RS_TH - 1) / RS_TH, n);
  resize_aa_norm_kernel<<<grid, RS_THREADS, smem, stream>>>(p);
  XS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// score-map post-processing
// ---------------------------------------------------------------------------------------------------------------
constexpr int PP_BLOCKS_PER_MAP = 64;  // partial sums per map (fixed: the mean is deterministic)

// one pass: partial sums (fp64) per block, uint16 quantisation, turbo RGB.  4 pixels per thread (float4 / ushort4).
__global__ void __launch_bounds__(256)
score_post_kernel(const float* __restrict__ score, long long hw, double* __restrict__ partial,
                  uint16_t* __restrict__ gray, int vrange11, uint8_t* __restrict__ rgb, float vmin, float vspan) {
  __shared__ uint8_t lut[768];
  if (rgb) {
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = g_turbo_u8[i];
    __syncthreads();
  }
  const int map = blockIdx.y;
  const float* s = score + static_cast<long long>(map) * hw;
  const long long n4 = hw >> 2;
  double acc = 0.0;
  auto quant = [&](float m) -> uint16_t {
    // utils/io/images.py:55-61: fp32 scale, astype(int32) truncates; the PNG keeps the low 16 bits
    const float q = vrange11 ? __fmul_rn(__fadd_rn(m, 1.0f), 32767.0f) : __fmul_rn(m, 65535.0f);
    return static_cast<uint16_t>(static_cast<int>(q));
  };
  auto colour = [&](float m, uint8_t* dst) {
    // plt.Normalize in fp32, Colormap.__call__: x*N, x == N -> N-1, <0 -> first, >= N -> last, NaN -> (0,0,0)
    float xa = __fmul_rn(__fdiv_rn(__fsub_rn(m, vmin), vspan), 256.0f);
    if (xa == 256.0f) xa = 255.0f;
    if (xa != xa) {
      dst[0] = dst[1] = dst[2] = 0;
      return;
    }
    const int idx = xa < 0.f ? 0 : (xa >= 256.0f ? 255 : static_cast<int>(xa));
    dst[0] = lut[3 * idx];
    dst[1] = lut[3 * idx + 1];
    dst[2] = lut[3 * idx + 2];
  };
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(s + 4 * i);
    if (partial) acc += (static_cast<double>(v.x) + v.y) + (static_cast<double>(v.z) + v.w);
    if (gray) {
      ushort4 q;
      q.x = quant(v.x); q.y = quant(v.y); q.z = quant(v.z); q.w = quant(v.w);
      *reinterpret_cast<ushort4*>(gray + static_cast<long long>(map) * hw + 4 * i) = q;
    }
    if (rgb) {
      uint8_t c[12];
      colour(v.x, c); colour(v.y, c + 3); colour(v.z, c + 6); colour(v.w, c + 9);
      uint32_t* d = reinterpret_cast<uint32_t*>(rgb + (static_cast<long long>(map) * hw + 4 * i) * 3);
      d[0] = c[0] | (c[1] << 8) | (c[2] << 16) | (static_cast<uint32_t>(c[3]) << 24);
      d[1] = c[4] | (c[5] << 8) | (c[6] << 16) | (static_cast<uint32_t>(c[7]) << 24);
      d[2] = c[8] | (c[9] << 8) | (c[10] << 16) | (static_cast<uint32_t>(c[11]) << 24);
    }
  }
  // tail (hw not a multiple of 4): block 0, first threads
  if (blockIdx.x == 0 && threadIdx.x < (hw & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float v = s[i];
    if (partial) acc += v;
    if (gray) gray[static_cast<long long>(map) * hw + i] = quant(v);
    if (rgb) colour(v, rgb + (static_cast<long long>(map) * hw + i) * 3);
  }
  if (partial) {
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      partial[static_cast<long long>(map) * gridDim.x + blockIdx.x] = t;
    }
  }
}

__global__ void score_mean_finalize_kernel(const double* __restrict__ partial, int nblk, long long hw,
                                           float* __restrict__ mean, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double t = 0.0;
  for (int k = 0; k < nblk; ++k) t += partial[static_cast<long long>(b) * nblk + k];  // fixed order
  mean[b] = static_cast<float>(t / static_cast<double>(hw));
}

size_t postprocess_workspace_bytes(int B) { return static_cast<size_t>(B) * PP_BLOCKS_PER_MAP * sizeof(double); }

int postprocess_score(const float* score, int B, int H, int W, float* frame_mean, uint16_t* gray16, int vrange_mode,
                      uint8_t* rgb, float vmin, float vmax, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  XS_CHECK_ARG(B > 0 && H > 0 && W > 0, "postprocess: empty problem");
  XS_CHECK_ARG(vrange_mode == 0 || vrange_mode == 1, "postprocess: vrange_mode must be 0 ([0,1]) or 1 ([-1,1])");
  const long long hw = static_cast<long long>(H) * W;
  XS_CHECK_ARG(((reinterpret_cast<uintptr_t>(score) & 15) == 0) && (hw % 4 == 0 || B == 1),
               "postprocess: maps must be 16-byte aligned (H*W %% 4 == 0 for B > 1)");
  XS_CHECK_ARG(rgb == nullptr || vmax != vmin, "postprocess: empty colour range");
  double* partial = nullptr;
  if (frame_mean) {
    XS_CHECK_ARG(workspace != nullptr && ws_bytes >= postprocess_workspace_bytes(B),
                 "postprocess: workspace of %zu bytes required for the frame means", postprocess_workspace_bytes(B));
    partial = static_cast<double*>(workspace);
  }
  dim3 grid(PP_BLOCKS_PER_MAP, B);
  score_post_kernel<<<grid, 256, 0, stream>>>(score, hw, partial, gray16, vrange_mode, rgb, vmin,
                                              static_cast<float>(static_cast<double>(vmax) - static_cast<double>(vmin)));
  XS_LAUNCH_CHECK();
  if (frame_mean) {
    score_mean_finalize_kernel<<<(B + 127) / 128, 128, 0, stream>>>(partial, PP_BLOCKS_PER_MAP, hw, frame_mean, B);
    XS_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace xs
