// first_conv.cuh — temp1.inc.convblock.0 (4 -> 64 channels, ReLU6) fused with the input staging.
//
// Reference: the torch.cat([input, noise_map]) of BSVD.forward (bsvd_arch.py:492-493) and the first
// nn.Conv2d of InputCvBlock (bsvd_arch.py:207-209).  With only 4 input channels the 3x3x4 patch of a
// pixel (36 values) fits one 128-byte K row, so the conv is a single-tap GEMM with K = 64 (36 used).
// Instead of materialising the patches in HBM (a 663 MB round trip per 10-frame clip) eight producer
// warps build them directly in shared memory, in the SWIZZLE_128B K-major layout tcgen05.mma reads:
//   warps 0-7  patch producers, two groups of 4 warps taking alternate tiles: fp32 NCHW (+ noise
//              map) -> 16-bit swizzled rows, fence.proxy.async
//   warp  8    MMA issuer (8 MMAs per 2-row tile, filter bank resident)
//   warps 9-16 epilogue (shared with conv_tc.cuh: bias, ReLU6, staged coalesced NHWC stores)
//   RAW instances use another split of the same roles: 4 producer warps, the MMA warp, 16 epilogue warps and a
//   TMA loader warp for the raw fp32 tiles (see kFirstThreadsRaw)
// RAW = true (plain fp32 NCHW input, no reflection): the haloed raw tile [4 planes][4 rows][136 px] of
// every tile is fetched by TMA into a 6-deep shared-memory ring, several tiles ahead (out-of-bounds = the
// conv's zero padding), and the producers read their 48 values with conflict-free LDS.  With direct global
// loads (RAW = false: uint8 / reflect-padded callers) each producer thread sits on 48 scalar LDGs of DRAM
// latency per tile with only two tiles in flight per SM — measured 0.40 of the HBM roofline.
#pragma once
#include "conv_tc.cuh"

namespace bsvd {

constexpr int kFirstR = 2;
constexpr int kFirstThreads = 32 * 17;        // direct-load instances: 8 producer + 1 MMA + 8 epilogue warps
constexpr int kFirstProducers = 128;          // threads of one producer group
constexpr int kFirstStages = 4;
constexpr uint32_t kFirstAStage = kFirstR * kRunPx * 128;   // 32 KB
constexpr uint32_t kFirstW = 64 * 128;                      // 8 KB
// epilogue staging: two 2 KB tiles per warp (units leave through TMA stores, see conv_tc.cuh)
constexpr size_t kFirstSmem = 1024 + kFirstStages * kFirstAStage + kFirstW + 8 * 2 * kStageBytesPerWarp;
// raw-tile ring of the RAW instances
constexpr int kRawStages = 6;
constexpr int kRawPx = 136;                                  // columns x0-4 .. x0+131: a TMA box must START on a 16-byte
                                                             // boundary of the innermost dimension (x0-1 on fp32 raises
                                                             // 'illegal instruction', tools/probes/tma_probe.cu), and 136
                                                             // makes a plane (4 rows) a multiple of 128 B, the alignment a
                                                             // TMA destination needs (the noise-map plane is its own box)
constexpr int kRawLead = 4;                                  // columns in front of x0 (3 unused + the halo column)
constexpr int kRawRows = kFirstR + 2;
constexpr uint32_t kRawPlane = kRawRows * kRawPx * 4;        // 2176 B
constexpr uint32_t kRawStage = 4 * kRawPlane;                // 8704 B (4 channel planes)
static_assert(kRawPlane % 128 == 0, "TMA destinations must be 128-byte aligned");
// RAW instances: the producers no longer wait on DRAM, so ONE group of 4 warps keeps up; the warps saved go to
// the epilogue — 16 epilogue warps, one unit each per tile — which is what the kernel was waiting for (the ncu
// source page showed the producers parked 35 % of all samples behind the eight epilogue warps).
//   warps 0-3 producers | warp 4 MMA | warps 5-20 epilogue | warp 21 raw-tile loader
constexpr int kFirstStagesRaw = 3;
constexpr int kFirstEpiWarpsRaw = 16;
constexpr size_t kFirstSmemRaw = 1024 + kFirstStagesRaw * kFirstAStage + kFirstW +
                                 kFirstEpiWarpsRaw * kStageBytesPerWarp + kRawStages * kRawStage;
constexpr int kFirstThreadsRaw = 32 * (4 + 1 + kFirstEpiWarpsRaw + 1);

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// U8: the raw input is uint8 [T][H][W][3] (decoded frames, HWC; ConvParams::u8_bgr = channel order
// B,G,R as cv2 delivers them): normalised with /255 on the fly (img2tensor,
// BasicSR/basicsr/utils/img_util.py), and the normalised RGB planes of every pixel are also written
// to `norm_out` (fp32 [T][3][H][W]) for the temp1 residual, which needs the raw input again.
// ACT6 (RAW instances): the stage applies ReLU6 (act 'relu6', the BSVD-64 yml) — the epilogue is compiled without
// the plain-ReLU path and the fp16 range guard (epilogue_unit, kOnly6).  This kernel is bound by the instructions
// its epilogue warps issue (ncu: 83 % of all warp instructions, issue slots 55 % busy).
template <bool BF16, bool U8 = false, bool RAW = false, bool ACT6 = false>
__global__ void __launch_bounds__(RAW ? kFirstThreadsRaw : kFirstThreads, 1)
first_conv_kernel(const float* __restrict__ in, const float* __restrict__ nmap, int in_c,
                  const __grid_constant__ CUtensorMap map_o, const __grid_constant__ ConvParams p,
                  float* __restrict__ norm_out, const __grid_constant__ CUtensorMap map_raw,
                  const __grid_constant__ CUtensorMap map_rawnm) {
  static_assert(!(U8 && RAW), "the raw-tile TMA path takes fp32 planes");
  constexpr int NT = 64;
  constexpr int kAccCols = kFirstR * NT;
  constexpr int kTmemCols = 2 * kAccCols;
  constexpr int kStages = RAW ? kFirstStagesRaw : kFirstStages;       // patch stages
  constexpr int kProdWarps = RAW ? 4 : 8, kGroups = kProdWarps / 4;    // producer groups of 4 warps
  constexpr int kMmaWarp = kProdWarps;
  constexpr int kEpiWarps = RAW ? kFirstEpiWarpsRaw : 8;
  constexpr int kEpi0 = kMmaWarp + 1;                                  // first epilogue warp
  constexpr int kLoaderWarp = kEpi0 + kEpiWarps;                       // RAW only
  constexpr uint32_t kStgPerWarp = RAW ? kStageBytesPerWarp : 2 * kStageBytesPerWarp;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kFirstStages + 5 + 2 * kRawStages];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = a_base + kStages * kFirstAStage;
  const uint32_t stg_base = w_base + kFirstW;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kFirstStages + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * kFirstStages + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * kFirstStages + 2 + b); };
  const uint32_t w_full = bar0 + 8u * (2 * kFirstStages + 4);
  auto raw_full = [&](int s) { return bar0 + 8u * (2 * kFirstStages + 5 + s); };
  auto raw_empty = [&](int s) { return bar0 + 8u * (2 * kFirstStages + 5 + kRawStages + s); };
  const uint32_t raw_base = stg_base + kEpiWarps * kStgPerWarp;
  constexpr int kNumThreads = RAW ? kFirstThreadsRaw : kFirstThreads;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(a_full(s), kFirstProducers);
      mbar_init(a_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), kEpiWarps);
    }
    mbar_init(w_full, 1);
    if constexpr (RAW) {
      for (int s = 0; s < kRawStages; ++s) {
        mbar_init(raw_full(s), 1);
        mbar_init(raw_empty(s), kFirstProducers);
      }
    }
    fence_barrier_init();
  }
  // K slots 36..63 of every patch row are zero for the whole kernel: clear the stages once
  for (uint32_t off = threadIdx.x * 16; off < kStages * kFirstAStage; off += kNumThreads * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_base + off), "r"(0) : "memory");
  if constexpr (RAW) {
    // a channel plane no TMA box ever fills (blind 3-channel input) must read as zero, not as garbage
    for (uint32_t off = threadIdx.x * 16; off < kRawStages * kRawStage; off += kNumThreads * 16)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(raw_base + off), "r"(0) : "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp < kProdWarps) {
    // =================================== patch producers ===================================
    // thread = one x position of the 128-px run; it builds the patch rows of both tile rows
    // (they share two of their three input rows).  Group g = warp/4 handles every second tile, so
    // the global-load latency of one group's tile overlaps the stores of the other's.
    const int i = threadIdx.x & 127;             // 0..127
    const int grp = threadIdx.x >> 7;            // 0 / 1
    const int sH = p.src_H ? p.src_H : p.H, sW = p.src_W ? p.src_W : p.W;   // raw image (reflected beyond)
    const long long plane = static_cast<long long>(sH) * sW;
    uint32_t it = grp;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.total_tiles; tile += kGroups * gridDim.x, it += kGroups) {
      const uint32_t sa = it % kStages, pa = (it / kStages) & 1;
      const TileCoord tc = decode_tile<kFirstR>(p, tile);
      const int x = tc.x0 + i;
      // 4 input rows (y0-1 .. y0+2) x 3 columns x 4 channels.  Row / column offsets (with the
      // reflection of the fused caller entry) and validity are computed once per tile: 7 instead of 12
      // coordinate computations, none inside the channel loop.
      float v[4][3][4];
      int ro[4], co[3];
      bool rok[4], cok[3];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int yy = tc.y0 - 1 + rr;
        rok[rr] = (yy >= 0) && (yy < p.H);
        ro[rr] = reflect_src(yy, sH) * sW;
      }
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x - 1 + dx;
        cok[dx] = (xx >= 0) && (xx < p.W);
        co[dx] = reflect_src(xx, sW);
      }
      // per-channel plane pointers of this frame (channel in_c.. = noise map, or a constant fill)
      const float* pc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        pc[c] = (c < in_c) ? in + (static_cast<long long>(tc.t) * in_c + c) * plane
                           : (nmap && c == in_c ? nmap + static_cast<long long>(tc.t) * plane : nullptr);
      const float fill = p.use_sigma ? p.sigma_const : 0.f;      // constant noise map (inside the image)
      if constexpr (RAW) {
        // the loader warp's TMA box [4 planes][4 rows][136 px] starts at (x0 - 4, y0 - 1): zero outside
        const uint32_t rs = it % kRawStages, rp = (it / kRawStages) & 1;
        mbar_wait(raw_full(rs), rp);
        const uint32_t src = raw_base + rs * kRawStage + (i + kRawLead - 1) * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              uint32_t u;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(src + c * kRawPlane + (rr * kRawPx + dx) * 4) : "memory");
              v[rr][dx][c] = __uint_as_float(u);
            }
        mbar_arrive_release(raw_empty(rs));
      } else {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const bool ok = rok[rr] && cok[dx];
          const int o = ro[rr] + co[dx];
          if constexpr (U8) {
            const uint8_t* px = reinterpret_cast<const uint8_t*>(in) +
                                (static_cast<long long>(tc.t) * plane + o) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c)
              v[rr][dx][c] = ok ? static_cast<float>(__ldg(px + (p.u8_bgr ? 2 - c : c))) / 255.0f : 0.f;
            v[rr][dx][3] = (ok && in_c == 3) ? fill : 0.f;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float f = 0.f;
              if (ok) f = pc[c] ? __ldg(pc[c] + o) : (c == in_c ? fill : 0.f);
              v[rr][dx][c] = f;
            }
          }
        }
      }
      }
      if constexpr (U8) {
        // this thread's own two pixels (rows y0, y0+1 at column x): normalised RGB for the residual
        if (norm_out && x < sW) {
#pragma unroll
          for (int r = 0; r < kFirstR; ++r) {
            const int y = tc.y0 + r;
            if (y < sH) {
#pragma unroll
              for (int c = 0; c < 3; ++c)
                norm_out[(static_cast<long long>(tc.t) * 3 + c) * plane + static_cast<long long>(y) * sW + x] =
                    v[r + 1][1][c];
            }
          }
        }
      }
      mbar_wait(a_empty(sa), pa ^ 1);
      const uint32_t stage = a_base + sa * kFirstAStage;
#pragma unroll
      for (int r = 0; r < kFirstR; ++r) {
        // k = tap*4 + c, tap = dy*3 + dx  ->  chunk j holds taps 2j, 2j+1 (chunk 4: tap 8 + zeros)
        const uint32_t row = stage + r * (kRunPx * 128) + i * 128;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          float f[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int tap = 2 * j + h;
#pragma unroll
            for (int c = 0; c < 4; ++c) f[4 * h + c] = (tap < 9) ? v[r + tap / 3][tap % 3][c] : 0.f;
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(row + ((j ^ (i & 7)) << 4)), "r"(pack2<BF16>(f[0], f[1])),
                         "r"(pack2<BF16>(f[2], f[3])), "r"(pack2<BF16>(f[4], f[5])),
                         "r"(pack2<BF16>(f[6], f[7]))
                       : "memory");
        }
      }
      fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async proxy
      mbar_arrive_release(a_full(sa));
    }
  } else if (RAW && warp == kLoaderWarp) {
    // ===================================== raw-tile loader =====================================
    if (lane == 0) {
      const int nm_planes = (nmap != nullptr) ? 1 : 0;
      const uint32_t tx = static_cast<uint32_t>(in_c + nm_planes) * kRawPlane;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t rs = it % kRawStages, rp = (it / kRawStages) & 1;
        const TileCoord tc = decode_tile<kFirstR>(p, tile);
        mbar_wait(raw_empty(rs), rp ^ 1);
        mbar_expect_tx(raw_full(rs), tx);
        const uint32_t dst = raw_base + rs * kRawStage;
        tma_load_3d(dst, &map_raw, raw_full(rs), tc.x0 - kRawLead, tc.y0 - 1, tc.t * in_c);
        if (nm_planes) tma_load_3d(dst + in_c * kRawPlane, &map_rawnm, raw_full(rs), tc.x0 - kRawLead, tc.y0 - 1, tc.t);
      }
    }
  } else if (warp == kMmaWarp) {
    // ====================================== MMA issuer ======================================
    // 32 output channels (c32 configurations): N = 32, the accumulator keeps its 64-column row pitch
    const uint32_t idesc = make_idesc(p.out_C == 32 ? 32 : NT, BF16 ? 1 : 0);
    const uint32_t leader = elect_one();
    constexpr uint32_t kDescHi = 0x40000000u | (1u << 14) | (1024u >> 4);
    const uint64_t desc_hi = static_cast<uint64_t>(kDescHi) << 32;
    const uint32_t b_lo0 = ((w_base & 0x3FFFFu) >> 4) | (1u << 16);
    if (leader) {
      mbar_expect_tx(w_full, kFirstW);
      bulk_load(w_base, p.wpack, kFirstW, w_full);
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    uint32_t sa = 0, pa = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(acc_empty(buf), acc_phase ^ 1);
      mbar_wait(a_full(sa), pa);
      tc_fence_after();
      const uint32_t a_lo0 = (((a_base + sa * kFirstAStage) & 0x3FFFFu) >> 4) | (1u << 16);
      if (leader) {
#pragma unroll
        for (int r = 0; r < kFirstR; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + buf * kAccCols + r * NT,
                     desc_hi | (a_lo0 + static_cast<uint32_t>(r * (kRunPx * 8)) + k * 2u),
                     desc_hi | (b_lo0 + k * 2u), idesc, k > 0 ? 1u : 0u);
        umma_commit(a_empty(sa));
        umma_commit(acc_full(buf));
      }
      __syncwarp();
      if (++sa == kStages) { sa = 0; pa ^= 1; }
    }
  } else {
    // ======================================= epilogue =======================================
    const int ew = warp - kEpi0;               // 0 .. kEpiWarps-1
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    constexpr int G = NT / 32;
    constexpr int kEpiMask = (ACT6 ? EPI_RELU6 : (EPI_RELU6 | EPI_RELU)) | EPI_TMA_OUT;
    // which of the kEpiWarps / 4 warps of this quadrant: they split the R * G units of a tile
    int sub = 0;
    for (int w = kEpi0; w < warp; ++w) sub += ((w & 3) == quad) ? 1 : 0;
    uint32_t it = 0;
    if constexpr (RAW) {
      // 16 epilogue warps: one unit (row u / G, column group u % G) per warp and tile, one staging tile per warp
      const uint32_t stg = stg_base + ew * kStgPerWarp;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const TileCoord tc = decode_tile<kFirstR>(p, tile);
        EpiLane el = epi_lane<EPI_RELU6 | EPI_RELU>(p, lane);
        el.nvalid = min(32, p.W - (tc.x0 + quad * 32));
        const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
        const int units = (p.out_C == 32) ? kFirstR : kFirstR * G;      // 32 stored channels: one unit per row
        const int row = (p.out_C == 32) ? sub : sub / G, grp32 = (p.out_C == 32) ? 0 : sub % G;
        mbar_wait(acc_full(buf), acc_phase);
        tc_fence_after();
        const uint32_t tacc = tmem_base + lane_base + buf * kAccCols;
        uint32_t va[32];
        if (sub < units) {
          tmem_ld32(tacc + row * NT + grp32 * 32, va);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(buf));
        if (sub < units) {
          const uint4 nosk[4] = {};
          const float norin[3] = {};
          float bv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) bv[i] = p.bias_c[grp32 * 32 + i];
          epilogue_unit<BF16, kEpiMask>(p, tc, el, tc.y0 + row, grp32 * 32, va, nosk, bv, stg, quad, lane, norin, false, &map_o);
        }
      }
    } else {
    const int half = sub;
    const uint32_t stg = stg_base + ew * kStgPerWarp;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const TileCoord tc = decode_tile<kFirstR>(p, tile);
      EpiLane el = epi_lane<EPI_RELU6 | EPI_RELU>(p, lane);
      el.nvalid = min(32, p.W - (tc.x0 + quad * 32));
      const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(acc_full(buf), acc_phase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + buf * kAccCols;
      const uint4 nosk[4] = {};
      const float norin[3] = {};
      if (p.out_C == 32) {
        // 32 output channels: one unit per image row, one row per warp half; the two staging tiles alternate
        // per tile (one bulk group per tile: the group awaited before a tile is rewritten is two tiles old)
        uint32_t va[32];
        tmem_ld32(tacc + half * NT, va);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(buf));
        float bv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) bv[i] = p.bias_c[i];
        epilogue_unit<BF16, kEpiMask>(p, tc, el, tc.y0 + half, 0, va, nosk, bv, stg + (it & 1u) * kStageBytesPerWarp, quad, lane,
                                      norin, false, &map_o);
        continue;
      }
      const int u0 = half * 2;                 // 4 units (2 rows x 2 column groups), 2 per warp half
      uint32_t va[32], vb[32];
      tmem_ld32(tacc + (u0 / G) * NT + (u0 % G) * 32, va);
      tmem_ld32(tacc + ((u0 + 1) / G) * NT + ((u0 + 1) % G) * 32, vb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      float bv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) bv[i] = p.bias_c[(u0 % G) * 32 + i];
      epilogue_unit<BF16, kEpiMask>(p, tc, el, tc.y0 + u0 / G, (u0 % G) * 32, va, nosk, bv, stg, quad, lane, norin, false, &map_o);
#pragma unroll
      for (int i = 0; i < 32; ++i) bv[i] = p.bias_c[((u0 + 1) % G) * 32 + i];
      epilogue_unit<BF16, kEpiMask>(p, tc, el, tc.y0 + (u0 + 1) / G, ((u0 + 1) % G) * 32, vb, nosk, bv, stg + kStageBytesPerWarp, quad, lane, norin, false, &map_o);
    }
    }
    if (lane == 0) bulk_wait_group_all();      // staging must outlive the last TMA reads
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace bsvd
