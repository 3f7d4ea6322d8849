// final_conv.cuh — temp2.outc.convblock.3 (64 -> 3 channels) + residual, fp32 NCHW output
// (reference: OutputCvBlock's last nn.Conv2d, bsvd_arch.py:295-298, and none_minus :408-414).
//
// With only 3 output channels an N=16 MMA per (tap, k-step) is dominated by the fixed cost of
// fetching its 128x16 A operand from shared memory.  This kernel stacks the three vertical taps
// into the N dimension instead: for every haloed INPUT row hr it accumulates
//     D_hr[p, dy*3+co] = sum_dx sum_ci X[hr][p+dx][ci] * W[co][ci][dy][dx]          (12 MMAs/row)
// and the epilogue finishes the 3x3 sum with three fp32 adds per output:
//     out[r][p][co] = skip - (D_r[p,co] + D_{r+1}[p,3+co] + D_{r+2}[p,6+co] + bias[co]).
// Per 4 output rows that is 72 MMAs instead of 144.  Same TMA halo tile, same descriptor-offset
// trick, same warp roles as conv_tc.cuh; single CTA (there is no filter slab worth sharing).
#pragma once
#include "conv_tc.cuh"

namespace bsvd {

constexpr int kFinalR = 4;                       // output rows per tile
constexpr int kFinalHaloRows = kFinalR + 2;
constexpr int kFinalN = 16;                      // GEMM N: 9 used columns (dy, co)
constexpr uint32_t kFinalATx = kFinalHaloRows * kRowBytes;              // 99840
constexpr uint32_t kFinalAStage = (kFinalATx + 1023u) & ~1023u;         // 100352
constexpr uint32_t kFinalWStage = kFinalN * 128;                        // one dx slab: 2 KB
constexpr size_t kFinalSmem = 1024 + 2 * kFinalAStage + 3 * kFinalWStage;
// PAIR instances (c32 configurations): the 32-channel input is read as pixel PAIRS ([H][W/2][64]); one GEMM
// row is a pair, the 3 pair taps (x-1, x, x+1) replace dx, and GEMM column n = dy*8 + a*4 + co is output
// channel co of pixel 2x + a (24 of 32 columns used; filter rows for |2*dxp + b - a| > 1 are zero).
constexpr int kFinalNPair = 32;
constexpr uint32_t kFinalWStagePair = kFinalNPair * 128;                // 4 KB per pair-tap slab
constexpr size_t kFinalSmemPair = 1024 + 2 * kFinalAStage + 3 * kFinalWStagePair;

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

template <bool BF16, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1)
final_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ ConvParams p) {
  constexpr int kN = PAIR ? kFinalNPair : kFinalN;         // GEMM N = TMEM columns per halo row
  constexpr uint32_t kWStage = PAIR ? kFinalWStagePair : kFinalWStage;
  constexpr int kAccStride = PAIR ? 256 : 128;   // TMEM columns between the two accumulator sets
  constexpr int kTmemCols = PAIR ? 512 : 256;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ float bias_s[4];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = a_base + 2 * kFinalAStage;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (4 + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (6 + b); };
  const uint32_t w_full = bar0 + 8u * 8;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), kEpiThreads / 32);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 3) bias_s[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, 3 * kWStage);
      bulk_load(w_base, p.wpack, 3 * kWStage, w_full);
      uint32_t sa = 0, pa = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile<kFinalR>(p, tile);
        mbar_wait(a_empty(sa), pa ^ 1);
        mbar_expect_tx(a_full(sa), kFinalATx);
        tma_load_4d(a_base + sa * kFinalAStage, &map_a, a_full(sa), 0, tc.x0 - 1, tc.y0 - 1, tc.t);
        if (++sa == 2) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(kN, BF16 ? 1 : 0);
    const uint32_t leader = elect_one();
    constexpr uint32_t kDescHi = 0x40000000u | (1u << 14) | (1024u >> 4);
    const uint64_t desc_hi = static_cast<uint64_t>(kDescHi) << 32;
    const uint32_t b_lo0 = ((w_base & 0x3FFFFu) >> 4) | (1u << 16);
    uint32_t sa = 0, pa = 0, it = 0;
    mbar_wait(w_full, 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(acc_empty(buf), acc_phase ^ 1);
      mbar_wait(a_full(sa), pa);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * kAccStride;
      const uint32_t a_lo0 = (((a_base + sa * kFinalAStage) & 0x3FFFFu) >> 4) | (1u << 16);
      if (leader) {
#pragma unroll
        for (int hr = 0; hr < kFinalHaloRows; ++hr) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = desc_hi | (a_lo0 + static_cast<uint32_t>((hr * kHaloPx + dx) * 8) + k * 2u);
              const uint64_t bd = desc_hi | (b_lo0 + static_cast<uint32_t>(dx * (kWStage >> 4)) + k * 2u);
              umma_f16(tmem_acc + hr * kN, ad, bd, idesc, (dx > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(a_empty(sa));
        umma_commit(acc_full(buf));
      }
      __syncwarp();
      if (++sa == 2) { sa = 0; pa ^= 1; }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const int sH = p.src_H ? p.src_H : p.H, sW = p.src_W ? p.src_W : (PAIR ? 2 * p.W : p.W);   // crop back to the raw image
    const long long plane = static_cast<long long>(sH) * sW;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const TileCoord tc = decode_tile<kFinalR>(p, tile);
      const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
      const int x = tc.x0 + quad * 32 + lane;              // pixel, or pixel pair (PAIR)
      if constexpr (PAIR) {
        // skip operand of the pair: [T][H][2W][4] 16-bit -> 16 bytes per pair
        uint4 skp[2] = {};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int y = tc.y0 + half + 2 * i;
          if (x < p.W && y < p.H)
            skp[i] = __ldg(reinterpret_cast<const uint4*>(
                reinterpret_cast<const uint16_t*>(p.skip) + tc.t * p.skip_frame_stride +
                (static_cast<long long>(y) * (2 * p.W) + 2 * x) * p.skip_C));
        }
        mbar_wait(acc_full(buf), acc_phase);
        tc_fence_after();
        const uint32_t tacc = tmem_base + lane_base + buf * kAccStride;
        uint32_t d[2][3][8];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
            tmem_ld8(tacc + (half + 2 * i + dy) * kN + 8 * dy, d[i][dy]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(buf));
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int y = tc.y0 + half + 2 * i;
          if (x < p.W && y < p.H && y < sH) {
            const float2 s0 = unpack2<BF16>(skp[i].x), s1 = unpack2<BF16>(skp[i].y),
                         s2 = unpack2<BF16>(skp[i].z), s3 = unpack2<BF16>(skp[i].w);
            const float sv[2][3] = {{s0.x, s0.y, s1.x}, {s2.x, s2.y, s3.x}};
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              const int xp = 2 * x + a;
              if (xp >= sW) continue;
              float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(tc.t) * 3 * plane +
                         static_cast<long long>(y) * sW + xp;
              uint8_t* o8 = reinterpret_cast<uint8_t*>(p.out) +
                            (static_cast<long long>(tc.t) * plane + static_cast<long long>(y) * sW + xp) * 3;
#pragma unroll
              for (int co = 0; co < 3; ++co) {
                const float conv = __uint_as_float(d[i][0][4 * a + co]) + __uint_as_float(d[i][1][4 * a + co]) +
                                   __uint_as_float(d[i][2][4 * a + co]) + bias_s[co];
                float r = sv[a][co] - conv;
                if (p.clamp01) r = fminf(fmaxf(r, 0.f), 1.f);
                if (p.out_u8) o8[p.u8_bgr ? 2 - co : co] = static_cast<uint8_t>(__float2uint_rn(r * 255.0f));
                else o[co * plane] = r;
              }
            }
          }
        }
        continue;
      }
      // skip operand (temp1 output, channels 0..2) of this warp's two rows: in flight during the MMAs
      uint2 sk[2] = {};
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int y = tc.y0 + half + 2 * i;
        if (x < p.W && y < p.H)
          sk[i] = __ldg(reinterpret_cast<const uint2*>(
              reinterpret_cast<const uint16_t*>(p.skip) + tc.t * p.skip_frame_stride +
              (static_cast<long long>(y) * p.W + x) * p.skip_C));
      }
      mbar_wait(acc_full(buf), acc_phase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + buf * kAccStride;
      // output row r = half + 2i needs columns [3dy, 3dy+3) of halo row r + dy
      uint32_t d[2][3][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
          tmem_ld4(tacc + (half + 2 * i + dy) * kFinalN + 3 * dy, d[i][dy]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int y = tc.y0 + half + 2 * i;
        if (x < p.W && y < p.H) {
          const float2 s01 = unpack2<BF16>(sk[i].x), s23 = unpack2<BF16>(sk[i].y);
          const float sv[3] = {s01.x, s01.y, s23.x};
          float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(tc.t) * 3 * plane +
                     static_cast<long long>(y) * sW + x;
          if (y < sH && x < sW) {
            uint8_t* o8 = reinterpret_cast<uint8_t*>(p.out) +
                          (static_cast<long long>(tc.t) * plane + static_cast<long long>(y) * sW + x) * 3;
#pragma unroll
            for (int co = 0; co < 3; ++co) {
              const float conv = __uint_as_float(d[i][0][co]) + __uint_as_float(d[i][1][co]) +
                                 __uint_as_float(d[i][2][co]) + bias_s[co];
              float r = sv[co] - conv;
              if (p.clamp01) r = fminf(fmaxf(r, 0.f), 1.f);      // temp_denoise: torch.clamp(out, 0, 1)
              if (p.out_u8)      // tensor2img (img_util.py): (img * 255.0).round() as uint8, HWC
                o8[p.u8_bgr ? 2 - co : co] = static_cast<uint8_t>(__float2uint_rn(r * 255.0f));
              else
                o[co * plane] = r;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace bsvd
