// metrics.cuh — on-device SSIM per frame.
//
// Reference: calculate_ssim / _ssim (BasicSR/basicsr/metrics/psnr_ssim.py:49-128): per channel, the five
// local moments under an 11x11 Gaussian window (sigma 1.5, cv2.getGaussianKernel) at every position where
// the window fits (the [5:-5, 5:-5] slice of cv2.filter2D), float64 arithmetic,
//   ssim = (2 mu1 mu2 + c1)(2 s12 + c2) / ((mu1^2 + mu2^2 + c1)(s1 + s2 + c2)),  c1 = (0.01 L)^2, c2 = (0.03 L)^2,
// averaged over positions, then over channels.  The reference applies it to [0,255] images (L = 255); on
// [0,1] floats with L = 1 the value is the same number.  Called once per validation frame, not hot:
// a direct 121-tap window in double precision from a shared-memory tile, deterministic two-pass reduction.
#pragma once
#include <cuda_runtime.h>

namespace bsvd {

constexpr int kSsimWin = 11;
constexpr int kSsimTile = 16;                         // 16 x 16 window positions per block
constexpr int kSsimPatch = kSsimTile + kSsimWin - 1;  // 26

struct SsimWindow { double w[kSsimWin]; };            // 1-D Gaussian taps (the window is their outer product)

static __global__ void __launch_bounds__(kSsimTile * kSsimTile)
ssim_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, int H, int W, int cb,
                    double L, SsimWindow g, int tiles_x, int tiles_y, double* __restrict__ part) {
  __shared__ float pa[kSsimPatch][kSsimPatch + 1], pb[kSsimPatch][kSsimPatch + 1];
  __shared__ double red[kSsimTile * kSsimTile / 32];
  const int tile = blockIdx.x, c = blockIdx.y, t = blockIdx.z;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int hh = H - 2 * cb, ww = W - 2 * cb;             // cropped image
  const int oh = hh - (kSsimWin - 1), ow = ww - (kSsimWin - 1);   // window positions
  const long long base = (static_cast<long long>(t) * C + c) * H * W;
  const int y0 = ty * kSsimTile, x0 = tx * kSsimTile;
  for (int i = threadIdx.x; i < kSsimPatch * kSsimPatch; i += blockDim.x) {
    const int py = i / kSsimPatch, px = i - py * kSsimPatch;
    const int y = y0 + py, x = x0 + px;
    float va = 0.f, vb = 0.f;
    if (y < hh && x < ww) {
      const long long o = base + static_cast<long long>(y + cb) * W + (x + cb);
      va = a[o]; vb = b[o];
    }
    pa[py][px] = va; pb[py][px] = vb;
  }
  __syncthreads();
  const int ly = threadIdx.x / kSsimTile, lx = threadIdx.x - ly * kSsimTile;
  double val = 0.0;
  if (y0 + ly < oh && x0 + lx < ow) {
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int dy = 0; dy < kSsimWin; ++dy) {
      double r1 = 0, r2 = 0, r11 = 0, r22 = 0, r12 = 0;
#pragma unroll
      for (int dx = 0; dx < kSsimWin; ++dx) {
        const double u = pa[ly + dy][lx + dx], v = pb[ly + dy][lx + dx], w = g.w[dx];
        r1 += w * u; r2 += w * v; r11 += w * u * u; r22 += w * v * v; r12 += w * u * v;
      }
      const double w = g.w[dy];
      m1 += w * r1; m2 += w * r2; s11 += w * r11; s22 += w * r22; s12 += w * r12;
    }
    const double c1 = (0.01 * L) * (0.01 * L), c2 = (0.03 * L) * (0.03 * L);
    const double v1 = s11 - m1 * m1, v2 = s22 - m2 * m2, cov = s12 - m1 * m2;
    val = ((2 * m1 * m2 + c1) * (2 * cov + c2)) / ((m1 * m1 + m2 * m2 + c1) * (v1 + v2 + c2));
  }
  for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < kSsimTile * kSsimTile / 32; ++i) s += red[i];
    part[(static_cast<long long>(t) * C + c) * (tiles_x * tiles_y) + tile] = s;
  }
}

static __global__ void ssim_final_kernel(const double* __restrict__ part, int n_per_frame, double count,
                                         float* __restrict__ ssim) {
  const int t = blockIdx.x;
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_per_frame; i += blockDim.x) s += part[static_cast<long long>(t) * n_per_frame + i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    ssim[t] = static_cast<float>(tot / count);
  }
}

}  // namespace bsvd
