// split_edge.cuh — first and last conv of the fp32-grade mode (BSVD_PREC_FP32X3), on the CUDA cores in fp32.
//
// In that mode every activation is stored as a (hi, lo) fp16 pair ([hi(C) | lo(C)] per pixel) and the 64- to
// 256-channel convs run three tensor-core products (conv_tc.cuh, EPI_SPLIT).  The two edge convs have 4 input /
// 3 output channels: 0.2 % of the FLOPs, no tensor-core shape worth building for a mode that runs at a third
// of the speed anyway — plain fp32 FMAs, weights broadcast from shared memory.
//   first: fp32 NCHW frames (+ noise map / constant sigma, reflect padding of the fused caller entry)
//          -> conv 3x3 (<=4 -> 64) + bias + ReLU6/ReLU -> [T][H][W][hi 64 | lo 64] fp16
//          (reference: torch.cat + InputCvBlock's first conv, bsvd_arch.py:492-493, 207-209)
//   final: [T][H][W][hi 64 | lo 64] -> conv 3x3 (64 -> 3) + bias, out = skip1 - conv (none_minus, :408-414),
//          optional clamp / crop / uint8 -> fp32 NCHW
#pragma once
#include "conv_tc.cuh"

namespace bsvd {

constexpr int kEdgeThreads = 128;

// w: fp32 [64][4][9] (zeros for channels the model does not have), bias fp32 [64]
static __global__ void __launch_bounds__(kEdgeThreads)
first_conv_split_kernel(const float* __restrict__ in, const float* __restrict__ nmap, int in_c,
                        const float* __restrict__ w, const float* __restrict__ bias, uint16_t* __restrict__ out,
                        int T, int H, int W, int src_H, int src_W, int use_sigma, float sigma_const, int relu6) {
  __shared__ float ws[36][64];      // [k = tap*4 + ci][co]
  __shared__ float bs[64];
  for (int i = threadIdx.x; i < 36 * 64; i += kEdgeThreads) {
    const int k = i / 64, co = i % 64, tap = k / 4, ci = k % 4;
    ws[k][co] = w[(co * 4 + ci) * 9 + tap];
  }
  if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int x = blockIdx.x * kEdgeThreads + threadIdx.x, y = blockIdx.y, t = blockIdx.z;
  if (x >= W) return;
  const int sH = src_H ? src_H : H, sW = src_W ? src_W : W;
  const long long plane = static_cast<long long>(sH) * sW;
  float patch[36];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int yy = y + dy - 1, xx = x + dx - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;      // zero padding of the (padded) network input
      const long long o = static_cast<long long>(reflect_src(yy, sH)) * sW + reflect_src(xx, sW);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v = 0.f;
        if (ok) {
          if (c < in_c) v = __ldg(in + (static_cast<long long>(t) * in_c + c) * plane + o);
          else if (c == in_c) v = nmap ? __ldg(nmap + static_cast<long long>(t) * plane + o) : (use_sigma ? sigma_const : 0.f);
        }
        patch[(dy * 3 + dx) * 4 + c] = v;
      }
    }
  uint16_t* dst = out + ((static_cast<long long>(t) * H + y) * W + x) * 128;
#pragma unroll 1
  for (int cb = 0; cb < 64; cb += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bs[cb + i];
#pragma unroll
    for (int k = 0; k < 36; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[k][cb]), w1 = *reinterpret_cast<const float4*>(&ws[k][cb + 4]);
      const float pv = patch[k];
      acc[0] = fmaf(pv, w0.x, acc[0]); acc[1] = fmaf(pv, w0.y, acc[1]); acc[2] = fmaf(pv, w0.z, acc[2]); acc[3] = fmaf(pv, w0.w, acc[3]);
      acc[4] = fmaf(pv, w1.x, acc[4]); acc[5] = fmaf(pv, w1.y, acc[5]); acc[6] = fmaf(pv, w1.z, acc[6]); acc[7] = fmaf(pv, w1.w, acc[7]);
    }
    uint4 hi, lo;
    uint32_t* hp = reinterpret_cast<uint32_t*>(&hi);
    uint32_t* lp = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = acc[2 * i], b = acc[2 * i + 1];
      a = relu6 ? relu6f(a) : fmaxf(a, 0.f);
      b = relu6 ? relu6f(b) : fmaxf(b, 0.f);
      hp[i] = pack2<false>(a, b);
      const float2 h = unpack2<false>(hp[i]);
      lp[i] = pack2<false>(a - h.x, b - h.y);
    }
    *reinterpret_cast<uint4*>(dst + cb) = hi;
    *reinterpret_cast<uint4*>(dst + 64 + cb) = lo;
  }
}

// in: [T][H][W][hi 64 | lo 64]; w: fp32 [3][64][9]; skip: fp32 [T][H][W][4] (temp1 output channels 0..3)
static __global__ void __launch_bounds__(kEdgeThreads)
final_conv_split_kernel(const uint16_t* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                        const float* __restrict__ skip, void* __restrict__ out, int T, int H, int W, int src_H, int src_W,
                        int clamp01, int out_u8, int u8_bgr) {
  __shared__ float ws[9][64][4];    // [tap][ci][co (3 used)]
  for (int i = threadIdx.x; i < 9 * 64 * 4; i += kEdgeThreads) {
    const int co = i & 3, ci = (i >> 2) & 63, tap = i >> 8;
    ws[tap][ci][co] = co < 3 ? w[(co * 64 + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  const int x = blockIdx.x * kEdgeThreads + threadIdx.x, y = blockIdx.y, t = blockIdx.z;
  const int sH = src_H ? src_H : H, sW = src_W ? src_W : W;
  if (x >= W || x >= sW || y >= sH) return;
  float a0 = bias[0], a1 = bias[1], a2 = bias[2];
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const uint16_t* px = in + ((static_cast<long long>(t) * H + yy) * W + xx) * 128;
#pragma unroll
    for (int c8 = 0; c8 < 64; c8 += 8) {
      const uint4 hv = __ldg(reinterpret_cast<const uint4*>(px + c8)), lv = __ldg(reinterpret_cast<const uint4*>(px + 64 + c8));
      const uint32_t* hp = reinterpret_cast<const uint32_t*>(&hv);
      const uint32_t* lp = reinterpret_cast<const uint32_t*>(&lv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 h = unpack2<false>(hp[i]), l = unpack2<false>(lp[i]);
        const float v0 = h.x + l.x, v1 = h.y + l.y;
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[tap][c8 + 2 * i][0]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[tap][c8 + 2 * i + 1][0]);
        a0 = fmaf(v0, w0.x, a0); a1 = fmaf(v0, w0.y, a1); a2 = fmaf(v0, w0.z, a2);
        a0 = fmaf(v1, w1.x, a0); a1 = fmaf(v1, w1.y, a1); a2 = fmaf(v1, w1.z, a2);
      }
    }
  }
  const float4 sk = __ldg(reinterpret_cast<const float4*>(skip) + (static_cast<long long>(t) * H + y) * W + x);
  float r[3] = {sk.x - a0, sk.y - a1, sk.z - a2};
  const long long plane = static_cast<long long>(sH) * sW;
#pragma unroll
  for (int co = 0; co < 3; ++co) {
    float v = r[co];
    if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
    if (out_u8)
      reinterpret_cast<uint8_t*>(out)[(static_cast<long long>(t) * plane + static_cast<long long>(y) * sW + x) * 3 + (u8_bgr ? 2 - co : co)] =
          static_cast<uint8_t>(__float2uint_rn(v * 255.0f));
    else
      reinterpret_cast<float*>(out)[(static_cast<long long>(t) * 3 + co) * plane + static_cast<long long>(y) * sW + x] = v;
  }
}

}  // namespace bsvd
