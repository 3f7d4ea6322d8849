// bsvd_capi.cu — host side of the B200-native BSVD-64 path: weight prepack, TMA tensor maps,
// the 32-stage clip/stream schedules and the extern "C" boundary declared in include/bsvd_b200.h.
//
// Reference being replaced: Experimental_root/archs/bsvd_arch.py (BSVD.forward :490-499,
// streaming_forward :501-552, DenBlock.forward :374-396, BiBufferConv :53-114, MemSkip :308-322).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bsvd_b200.h"
#include "conv_tc.cuh"
#include "stage_launch.cuh"
#include "final_conv.cuh"
#include "first_conv.cuh"
#include "metrics.cuh"
#include "split_edge.cuh"

namespace bsvd {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
// ------------------------------------------------------------------------------------------------
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Stride-1 view: 16-bit NHWC [T][H][W][C]; box = [1][R+2][130][64] (haloed tile of one chunk).
static int make_map_halo(CUtensorMap* m, const void* base, int T, int H, int W, int C, int R) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kChunk, (cuuint32_t)kHaloPx, (cuuint32_t)(R + 2), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides,
                   box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(halo) failed: %d", (int)r);
  return 0;
}
// Stride-2 view of the same tensor: [T][H/2][2][W/2][2*C]; box = [1][R][1][box_px][64]: one tap
// (box_px 128, generic pipeline) or one input sub-plane (box_px 129, R or R+1 rows; conv_tc.cuh PIPE 4).
static int make_map_s2(CUtensorMap* m, const void* base, int T, int H, int W, int C, int R, int box_px = kRunPx) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)T};
  cuuint64_t strides[4] = {(cuuint64_t)2 * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)2 * W * C * 2,
                           (cuuint64_t)H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)kChunk, (cuuint32_t)box_px, 1, (cuuint32_t)R, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<void*>(base), dims, strides,
                   box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(s2) failed: %d", (int)r);
  return 0;
}

// [T][Ho][Wo][C] seen as [T][Ho][Wo/2][2*C] (PixelShuffle: the two sub-pixel columns of a source pixel
// are adjacent in memory), or as is (ps = 0).  box_c x box_px elements per box; `swz` 128 B for MMA
// operands (skip blocks [128 px][64 ch]), 64 B for epilogue units ([32 px][32 ch] TMA stores).
// T may be a ring of slots with its own stride (streaming mode).
static int make_map_pix(CUtensorMap* m, const void* base, int T, long long t_stride_bytes, int Ho, int Wo,
                        int C, int ps, int box_c, int box_px, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
  const int cw = ps ? 2 * C : C, wp = ps ? Wo / 2 : Wo;
  cuuint64_t dims[4] = {(cuuint64_t)cw, (cuuint64_t)wp, (cuuint64_t)Ho, (cuuint64_t)T};
  cuuint64_t strides[3] = {(cuuint64_t)cw * 2, (cuuint64_t)Wo * C * 2, (cuuint64_t)t_stride_bytes};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_px, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(pixel view) failed: %d", (int)r);
  return 0;
}

// Raw network input for the first conv's TMA loader: fp32 planes [planes][H][W]; box = [box_planes][4][136].
static int make_map_raw(CUtensorMap* m, const void* base, long long planes, int H, int W, int box_planes) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {(cuuint32_t)kRawPx, (cuuint32_t)kRawRows, (cuuint32_t)box_planes};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(raw input) failed: %d", (int)r);
  return 0;
}
// The TMA path of the first conv needs 16-byte aligned planes (W % 4 == 0 holds for every network input).
static bool raw_tma_ok(const void* in, const void* nmap, int W) {
  static const int off = [] { const char* e = getenv("BSVD_B200_NO_RAW_TMA"); return (e && e[0] == '1') ? 1 : 0; }();
  return !off && (W % 4 == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)nmap % 16 == 0);
}

// Packed weights as a 2-D tensor [rows][64] (rows of 128 B, already swizzled by the host): the CTA-pair
// kernels fetch half of a filter slab per CTA with a TMA that can signal the leader CTA's barrier.
static int make_map_w(CUtensorMap* m, const void* base, size_t rows, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)kChunk, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kChunk * 2};
  cuuint32_t box[2] = {(cuuint32_t)kChunk, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides,
                   box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// 16-bit conversion on the host (round to nearest even, like the device epilogue)
// ------------------------------------------------------------------------------------------------
static uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static uint16_t f32_to_f16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return (uint16_t)(sign | (x > 0x7f800000u ? 0x7e00u : 0x7c00u));
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);   // rounds to >= 65520 -> inf
  if (x < 0x33000001u) return (uint16_t)sign;                // < 2^-25 -> 0
  int e = (int)(x >> 23) - 127;
  uint32_t m = (x & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t base;
  if (e >= -14) { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
  else { shift = 13 + (-14 - e); base = 0; }
  uint32_t q = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) ++q;
  return (uint16_t)(sign | (base + q));
}
static inline uint16_t to16(float f, int bf16) { return bf16 ? f32_to_bf16(f) : f32_to_f16(f); }
static float f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  uint32_t u;
  if (e == 0) {
    if (m == 0) u = sign;
    else { int sh = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; ++sh; } u = sign | ((uint32_t)(113 - sh) << 23) | ((mm & 0x3ffu) << 13); }
  } else if (e == 31) u = sign | 0x7f800000u | (m << 13);
  else u = sign | ((e + 112) << 23) | (m << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// ------------------------------------------------------------------------------------------------
// one fused conv stage: static description + packed weights + launch
// ------------------------------------------------------------------------------------------------
struct StageSpec {
  int cin = 64, cout = 64;   // conv channels as the kernels see them (multiples of 64 except the
                             // first conv's input and the last conv's output)
  int cin_l = 0, cout_l = 0; // the reference conv's channels when smaller (c32 configurations run
                             // zero-padded to the 64-channel kernels); 0 = same as cin / cout
  int stride = 1;
  bool relu6 = false, relu = false, pixshuf = false, skip = false, shift = false;
  bool resid_in = false, final_out = false, first_im2col = false;
  bool stacked = false;   // 64->64 stride-1 stage with the vertical taps stacked in N (conv_tc.cuh mode 2)
  // c32 configurations, native layout: a 32-channel full-resolution tensor [H][W][32] is the same memory as
  // [H][W/2][64], so a 3x3 conv over 32 channels runs as a 3x3 conv over pixel PAIRS with 64 "channels"
  // (column n = a*32 + co of pixel 2x+a, K index b*32 + ci of pixel 2x'+b, weight W[dy][2 dxp + b - a] or 0):
  // half of that GEMM multiplies structural zeros, but these stages are HBM-bound and now move 32 channels
  // per pixel instead of 64 zero-padded ones.
  bool pairx = false;       // stride-1 pair conv (physical 64 -> 64 on an image of W/2 pairs)
  bool pair_s2 = false;     // stride-2 conv reading pairs (conv_tc.cuh PIPE 5): 6 (dy, pair) taps
  bool pair_final = false;  // last conv (32 -> 3) on pairs (final_conv.cuh PAIR instances)
  int store_c = 0;          // first conv: channels actually stored when fewer than the GEMM's 64 columns
  // fp32-grade mode (BSVD_PREC_FP32X3, ConvParams::phys_chunks): activations stored as [hi(C) | lo(C)] fp16,
  // K walks [x_hi | x_lo | x_hi] against [W_hi | W_hi | W_lo]; the first / last conv run in fp32 on the CUDA
  // cores (split_edge.cuh)
  bool split = false;
  // derived
  int gemm_n = 64, ntile = 64, rows = 2, cin_chunks = 1, tap_begin = 0, tap_end = 9;
  void derive() {
    if (!cin_l) cin_l = cin;
    if (!cout_l) cout_l = cout;
    gemm_n = final_out ? 16 : cout;
    static const int ntile_max = [] { const char* e = getenv("BSVD_B200_NTILE_MAX"); return e ? atoi(e) : 256; }();
    ntile = gemm_n >= 256 ? (ntile_max >= 256 ? 256 : 128) : gemm_n;
    static const int r256 = [] { const char* e = getenv("BSVD_B200_R256"); return e ? atoi(e) : 1; }();
    rows = final_out ? kFinalR : ((ntile == 256) ? r256 : 2);
    cin_chunks = first_im2col ? 1 : cin / kChunk;
    tap_begin = first_im2col ? 4 : 0;
    tap_end = first_im2col ? 5 : 9;
    static const int stack_on = [] { const char* e = getenv("BSVD_B200_STACK"); return e ? atoi(e) : 1; }();
    static const int cta2_on = [] { const char* e = getenv("BSVD_B200_NO_CTA2"); return (e && e[0] == '1') ? 0 : 1; }();
    stacked = stack_on && cta2_on && cin == 64 && cout == 64 && stride == 1 && !first_im2col &&
              !final_out && !pixshuf && !skip && !split;
    if (split && !first_im2col && !final_out) cin_chunks = 3 * cin / kChunk;     // virtual chunks
    if (stacked) { tap_begin = 0; tap_end = 3; }   // three dx slabs of [192 rows per CTA][64]
    if (pair_s2) { tap_begin = 0; tap_end = 6; }
  }
  int ntaps() const { return tap_end - tap_begin; }
  int n_tiles() const { return gemm_n / ntile; }
  size_t pack_elems() const {
    if (split && first_im2col) return (size_t)64 * 4 * 9 * 2;     // fp32 [64][4][9] (counted in 16-bit units)
    if (split && final_out) return (size_t)3 * 64 * 9 * 2;        // fp32 [3][64][9]
    if (pair_final) return (size_t)3 * kFinalNPair * kChunk;   // [pair tap][dy*8+a*4+co][64]
    if (final_out) return (size_t)3 * kFinalN * kChunk;   // [dx][dy*3+co][64]
    if (stacked) return (size_t)3 * 2 * 192 * kChunk;     // [dx][cta rank][192 rows][64]
    return (size_t)n_tiles() * cin_chunks * ntaps() * ntile * kChunk;
  }
};

// GEMM column -> reference output channel (PixelShuffle permutes so that one sub-pixel's channels
// are contiguous: column q*Cq + c  <-  conv channel c*4 + q, nn.PixelShuffle semantics).
static inline int col_to_cout(const StageSpec& s, int col) {
  if (!s.pixshuf) return col < s.cout_l ? col : -1;
  const int cq = s.cout / 4;
  const int q = col / cq, c = col % cq;
  return c < s.cout_l / 4 ? c * 4 + q : -1;     // padded channels of a sub-pixel carry zeros
}

// Repack OIHW fp32 -> [n_tile][chunk][tap][ntile rows][64 k] 16-bit, each 128-byte row stored with
// its eight 16-byte chunks XOR-swizzled by (row & 7) (SWIZZLE_128B K-major canonical layout).
static void pack_weights(const StageSpec& s, const float* w, const float* b, int bf16,
                         std::vector<uint16_t>& pack, std::vector<float>& bias) {
  if (s.split && (s.first_im2col || s.final_out)) {
    // fp32 weights for the CUDA-core edge convs (split_edge.cuh)
    pack.assign(s.pack_elems(), 0);
    float* wf = reinterpret_cast<float*>(pack.data());
    if (s.first_im2col) {
      bias.assign(64, 0.f);
      for (int co = 0; co < s.cout_l; ++co) {
        bias[co] = b ? b[co] : 0.f;
        for (int ci = 0; ci < s.cin_l && ci < 4; ++ci)
          for (int t = 0; t < 9; ++t) wf[(co * 4 + ci) * 9 + t] = w[((size_t)co * s.cin_l + ci) * 9 + t];
      }
    } else {
      bias.assign(16, 0.f);
      for (int co = 0; co < s.cout_l; ++co) {
        bias[co] = b ? b[co] : 0.f;
        for (int ci = 0; ci < s.cin_l && ci < 64; ++ci)
          for (int t = 0; t < 9; ++t) wf[(co * 64 + ci) * 9 + t] = w[((size_t)co * s.cin_l + ci) * 9 + t];
      }
    }
    return;
  }
  if (s.split && s.cin_chunks == 3 * s.cin / kChunk && s.cin_l <= s.cin) {
    // [W_hi | W_hi | W_lo] over the virtual input channels [x_hi | x_lo | x_hi]; every entry is exactly
    // representable in fp16, so the 16-bit packing below is lossless
    const int C = s.cin;
    std::vector<float> w3((size_t)s.cout_l * 3 * C * 9, 0.f);
    for (int co = 0; co < s.cout_l; ++co)
      for (int ci = 0; ci < s.cin_l; ++ci)
        for (int t = 0; t < 9; ++t) {
          const float v = w[((size_t)co * s.cin_l + ci) * 9 + t];
          const float hi = f16_to_f32(f32_to_f16(v)), lo = f16_to_f32(f32_to_f16(v - hi));
          w3[((size_t)co * 3 * C + ci) * 9 + t] = hi;
          w3[((size_t)co * 3 * C + C + ci) * 9 + t] = hi;
          w3[((size_t)co * 3 * C + 2 * C + ci) * 9 + t] = lo;
        }
    StageSpec d = s;
    d.split = false; d.cin = 3 * C; d.cin_l = 3 * C;          // plain stage with 3C input channels
    pack_weights(d, w3.data(), b, 0, pack, bias);
    return;
  }
  if (s.pairx) {
    // expand to the dense 64 -> 64 pair conv and pack that like any other 64 -> 64 stage
    std::vector<float> wp((size_t)64 * 64 * 9, 0.f), bp(64, 0.f);
    for (int a = 0; a < 2; ++a)
      for (int co = 0; co < s.cout_l; ++co) {
        bp[a * 32 + co] = b ? b[co] : 0.f;
        for (int bb = 0; bb < 2; ++bb)
          for (int ci = 0; ci < s.cin_l; ++ci)
            for (int dy = 0; dy < 3; ++dy)
              for (int dxp = -1; dxp <= 1; ++dxp) {
                const int dx = 2 * dxp + bb - a;
                if (dx < -1 || dx > 1) continue;
                wp[((size_t)(a * 32 + co) * 64 + (bb * 32 + ci)) * 9 + dy * 3 + (dxp + 1)] =
                    w[((size_t)co * s.cin_l + ci) * 9 + dy * 3 + (dx + 1)];
              }
      }
    StageSpec d = s;
    d.pairx = false; d.cin_l = 64; d.cout_l = 64;
    pack_weights(d, wp.data(), bp.data(), bf16, pack, bias);
    return;
  }
  pack.assign(s.pack_elems(), 0);
  bias.assign(s.gemm_n, 0.f);
  for (int col = 0; col < s.gemm_n; ++col) {
    const int co = col_to_cout(s, col);
    if (co >= 0 && co < s.cout_l) bias[col] = b ? b[co] : 0.f;
  }
  if (s.pair_final) {
    // final_conv.cuh PAIR layout: slab = pair tap dxp, row n = dy*8 + a*4 + co, K = b*32 + ci
    for (int dxp = -1; dxp <= 1; ++dxp)
      for (int dy = 0; dy < 3; ++dy)
        for (int a = 0; a < 2; ++a)
          for (int co = 0; co < s.cout_l; ++co) {
            const int n = dy * 8 + a * 4 + co;
            for (int bb = 0; bb < 2; ++bb) {
              const int dx = 2 * dxp + bb - a;
              if (dx < -1 || dx > 1) continue;
              for (int ci = 0; ci < s.cin_l && ci < 32; ++ci) {
                const int k = bb * 32 + ci;
                const float v = w[((size_t)co * s.cin_l + ci) * 9 + dy * 3 + (dx + 1)];
                pack[((size_t)(dxp + 1) * kFinalNPair + n) * kChunk + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = to16(v, bf16);
              }
            }
          }
    return;
  }
  if (s.pair_s2) {
    // conv_tc.cuh PIPE 5: six slabs in consumption order (dy0,x-1) (dy0,x) (dy2,x-1) (dy2,x) (dy1,x-1) (dy1,x);
    // output pixel x reads input columns 2x-1 (pair x-1, b=1), 2x (pair x, b=0), 2x+1 (pair x, b=1)
    static const int kDy[6] = {0, 0, 2, 2, 1, 1};
    for (int j = 0; j < 6; ++j) {
      const int dy = kDy[j], dxp = (j & 1) ? 0 : -1;
      uint16_t* blk = pack.data() + (size_t)j * s.ntile * kChunk;
      for (int n = 0; n < s.ntile; ++n) {
        const int co = col_to_cout(s, n);
        if (co < 0 || co >= s.cout_l) continue;
        for (int bb = 0; bb < 2; ++bb) {
          const int dx = (dxp == -1) ? (bb == 1 ? 0 : -1) : 1 + bb;    // tap column index 0..2, -1 = none
          if (dx < 0) continue;
          for (int ci = 0; ci < s.cin_l && ci < 32; ++ci) {
            const int k = bb * 32 + ci;
            const float v = w[((size_t)co * s.cin_l + ci) * 9 + dy * 3 + dx];
            blk[(size_t)n * kChunk + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = to16(v, bf16);
          }
        }
      }
    }
    return;
  }
  if (s.final_out) {
    // final_conv.cuh layout: slab dx, row n = dy*3 + co, 64 input channels, SW128-swizzled rows
    for (int dx = 0; dx < 3; ++dx)
      for (int dy = 0; dy < 3; ++dy)
        for (int co = 0; co < s.cout_l; ++co) {
          const int n = dy * 3 + co;
          for (int k = 0; k < kChunk && k < s.cin_l; ++k) {
            const float v = w[((size_t)co * s.cin_l + k) * 9 + dy * 3 + dx];
            pack[((size_t)dx * kFinalN + n) * kChunk + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = to16(v, bf16);
          }
        }
    return;
  }
  if (s.stacked) {
    // conv_tc.cuh mode 2.  Per dx, per CTA of the pair, four slabs (row ranges inside the CTA's 24 KB):
    //   [0,32)    hr 0: dy0, couts [32*rank, +32)          -> acc0
    //   [32,96)   hr 1: rank 0 = dy1 (acc0), rank 1 = dy0 (acc1)
    //   [96,160)  hr 2: rank 0 = dy2 (acc0), rank 1 = dy1 (acc1)
    //   [160,192) hr 3: dy2, couts [32*rank, +32)          -> acc1
    for (int dx = 0; dx < 3; ++dx)
      for (int rank = 0; rank < 2; ++rank) {
        uint16_t* cta = pack.data() + ((size_t)dx * 2 + rank) * 192 * kChunk;
        for (int row = 0; row < 192; ++row) {
          int dy, co, local;   // local = row index inside its slab (swizzle phase)
          if (row < 32) { dy = 0; co = rank * 32 + row; local = row; }
          else if (row < 96) { dy = rank == 0 ? 1 : 0; co = row - 32; local = row - 32; }
          else if (row < 160) { dy = rank == 0 ? 2 : 1; co = row - 96; local = row - 96; }
          else { dy = 2; co = rank * 32 + (row - 160); local = row - 160; }
          if (co >= s.cout_l) continue;
          for (int k = 0; k < kChunk && k < s.cin_l; ++k) {
            const float v = w[((size_t)co * s.cin_l + k) * 9 + dy * 3 + dx];
            cta[(size_t)row * kChunk + (((k >> 3) ^ (local & 7)) << 3) + (k & 7)] = to16(v, bf16);
          }
        }
      }
    return;
  }
  const int nt_count = s.n_tiles();
  for (int nt = 0; nt < nt_count; ++nt)
    for (int c = 0; c < s.cin_chunks; ++c)
      for (int tap = s.tap_begin; tap < s.tap_end; ++tap) {
        uint16_t* blk = pack.data() +
            ((size_t)(nt * s.cin_chunks + c) * s.ntaps() + (tap - s.tap_begin)) * s.ntile * kChunk;
        for (int n = 0; n < s.ntile; ++n) {
          const int co = col_to_cout(s, nt * s.ntile + n);
          if (co < 0 || co >= s.cout_l) continue;
          for (int k = 0; k < kChunk; ++k) {
            float v = 0.f;
            if (s.first_im2col) {
              // K index = tap'*4 + ci: first_conv.cuh builds 4 slots per tap (slot 3 is the noise
              // map, or zero for the blind 3-channel variant)
              const int tp = k / 4, ci = k % 4;
              if (tp < 9 && ci < s.cin_l) v = w[((size_t)co * s.cin_l + ci) * 9 + tp];
            } else {
              const int ci = c * kChunk + k;
              if (ci < s.cin_l) v = w[((size_t)co * s.cin_l + ci) * 9 + tap];
            }
            const int chunk16 = (k >> 3) ^ (n & 7);
            blk[(size_t)n * kChunk + chunk16 * 8 + (k & 7)] = to16(v, bf16);
          }
        }
      }
}

struct StageDev {
  StageSpec spec;
  void* wpack = nullptr;
  float* bias = nullptr;
  std::vector<float> bias_h;   // host copy: travels in the kernel parameter bank (ConvParams::bias_c)
  bool loaded = false;
};

struct StageIO {
  const void* in = nullptr;    // 16-bit NHWC input tensor (full tensor base)
  int T = 1, H = 0, W = 0;     // input frames / rows / cols
  void* out = nullptr;
  void* out_prev = nullptr;
  void* out_next = nullptr;
  int ring_mode = 0;
  int zero_future = 0;
  const void* skip = nullptr;
  int skip_C = 0;
  long long skip_frame_stride = 0;
  const float* resid_in = nullptr;
  int resid_C = 0;
  void* aux_out = nullptr;
  unsigned* overflow = nullptr;
  // streaming: `skip` / `out` are ring bases of skip_T / out_T slots (0 = plain [T] tensors)
  int skip_T = 0, out_T = 0;
  long long skip_T_stride = 0, out_T_stride = 0;   // bytes
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// reciprocals for decode_tile (ConvParams::div_nt ...): umulhi(n, ceil(2^32 / d)) == n / d while n * d < 2^32
static int set_tile_div(ConvParams& p) {
  auto magic = [](int d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); };
  // the three dividends of decode_tile: tile index, pixel-tile position (one past the end for the padding tile
  // of an odd CTA pair), position / xblocks
  const uint64_t n1 = (uint64_t)p.total_tiles + 1, n2 = (uint64_t)p.positions + 2,
                 n3 = n2 / (uint64_t)std::max(p.xblocks, 1) + 1;
  const uint64_t lim = 1ull << 32;
  if (n1 * (uint64_t)p.n_tiles >= lim || n2 * (uint64_t)p.xblocks >= lim || n3 * (uint64_t)p.yblocks >= lim)
    return fail("too many tiles for one launch (%d frames of %d x %d tile rows / columns): split the clip", p.T,
                p.yblocks, p.xblocks);
  p.div_nt = magic(p.n_tiles); p.div_xb = magic(p.xblocks); p.div_yb = magic(p.yblocks);
  return 0;
}
static size_t staging_bytes(int ew) { return (size_t)ew * kStageBytesPerWarp; }   // epilogue staging

// conv3x3_tc_kernel instances live in their own translation units (conv_inst.cu, compiled once per
// (NTILE, R) pair in parallel); this one only calls them
extern template int launch_one<64, 2>(const StageLaunch&, cudaStream_t);
extern template int launch_one<128, 2>(const StageLaunch&, cudaStream_t);
extern template int launch_one<256, 1>(const StageLaunch&, cudaStream_t);
extern template int launch_one<256, 2>(const StageLaunch&, cudaStream_t);

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Build the launch record (tensor map + params) of one stage for concrete tensors.
static int use_cta2_default() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BSVD_B200_NO_CTA2");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v;
}

static int plan_stage(const StageDev& sd, const StageIO& io, int bf16, int desc_variant,
                      StageLaunch* L) {
  const StageSpec& s = sd.spec;
  ConvParams& p = L->p;
  memset(&p, 0, sizeof(p));
  if (sd.bias_h.size() > (size_t)kMaxBias) return fail("bias does not fit the parameter bank");
  std::copy(sd.bias_h.begin(), sd.bias_h.end(), p.bias_c);
  // CTA pairs (cta_group::2): one M=256 MMA drives two pixel tiles and each CTA stages only half
  // of the filter slab.  The 3-channel output stage (N=16) stays single-CTA.
  const int cta2 = (s.ntile != 16 && use_cta2_default() && !(desc_variant & 32)) ? 1 : 0;
  L->cta2 = cta2;
  // 8 epilogue warps: 16 were measured (kernel template parameter EW) and bring nothing — the
  // gap between a full stage and its epilogue-less run is shared-memory/L2 contention, not latency
  // (re-measured in round 2 on upc1.convblock.0, the stage with the longest epilogue — 8 units of ~175
  // instructions per warp and tile: 0.61 -> 0.63-0.67 ms with 16 warps and one staging tile each)
  L->ew = 8;
  // PixelShuffle + skip stages on the CTA-pair <256,1> kernel: the skip tensor is accumulated by an
  // identity MMA (TMA-loaded like an extra K chunk); without a temporal shift on the output the
  // epilogue units also leave through TMA stores (double-buffered staging).
  static const int up_on = [] { const char* e = getenv("BSVD_B200_NO_SKIP_MMA"); return (e && e[0] == '1') ? 0 : 1; }();
  // (upc2.convblock.0, with its temporal shift, keeps the skip add in the epilogue unless
  // BSVD_B200_SKIP_MMA_SHIFT=1: its MMA time is the bound, not its epilogue)
  static const int up_shift_on = [] { const char* e = getenv("BSVD_B200_SKIP_MMA_SHIFT"); return (e && e[0] == '1') ? 1 : 0; }();
  const int skip_blocks = s.ntile / 64 * (s.split ? 2 : 1);      // split: hi and lo halves of the skip tensor
  const bool skip_mma = (up_on || s.split) && cta2 && s.skip && s.pixshuf && s.ntile == 256 && s.rows == 1 &&
                        !s.first_im2col && !s.final_out && (!s.shift || up_shift_on || s.split) &&
                        s.cin_chunks * s.ntaps() >= 4 * skip_blocks;   // one skip block per four slabs
  if (s.split && s.skip && !skip_mma) return fail("fp32-grade mode: the skip add of this stage must run on the tensor core");
  static const int tma_on = [] { const char* e = getenv("BSVD_B200_NO_TMA_OUT"); return (e && e[0] == '1') ? 0 : 1; }();
  static const int tma_shift_on = [] { const char* e = getenv("BSVD_B200_TMA_SHIFT"); return e ? atoi(e) : 0; }();   // measured: -3 % when on
  const bool tma_out = tma_on && cta2 && (!s.shift || (tma_shift_on && !s.pixshuf)) &&
                       !(s.skip && !skip_mma) && !s.first_im2col && s.stride == 1 &&
                       !s.final_out && !(desc_variant & ~(2 | 4 | 16 | 128));
  // two staging tiles per warp when they fit; the stacked 64->64 stages (resident 72 KB bank) have
  // room for one, whose TMA read is awaited right before it is rewritten
  const int stg_bufs = ((tma_out && !s.stacked) || s.split) ? 2 : 1;   // split: hi and lo tiles
  const size_t kStagingBytes = staging_bytes(L->ew) * stg_bufs + (skip_mma ? 4096 : 0);
  const int Ho = io.H / s.stride, Wo = io.W / s.stride;
  if (s.stride == 2 && ((io.H & 1) || (io.W & 1))) return fail("stride-2 stage needs even H, W");
  p.T = io.T; p.H = Ho; p.W = Wo;
  p.overflow = io.overflow;
  p.pair_px = s.pairx ? 1 : 0;
  p.cin_chunks = s.cin_chunks;
  p.n_tiles = s.n_tiles();
  p.tap_begin = s.tap_begin; p.tap_end = s.tap_end;
  p.xblocks = (Wo + kRunPx - 1) / kRunPx;
  p.yblocks = (Ho + s.rows - 1) / s.rows;
  L->split = s.split ? 1 : 0;
  p.out_pitch_C = 0;
  if (s.split && (s.first_im2col || s.final_out)) {
    // fp32 CUDA-core edge convs of the fp32-grade mode (split_edge.cuh): one thread per pixel
    p.wpack = sd.wpack; p.bias = sd.bias;
    p.flags = s.first_im2col ? (s.relu ? EPI_RELU : EPI_RELU6) : EPI_FINAL;
    p.out = io.out; p.skip = io.skip;
    if (s.final_out && !io.skip) return fail("stage needs a skip tensor");
    L->first_in = nullptr;
    L->in_ptr = io.in;
    L->cta2 = 0;
    L->map = CUtensorMap(); L->map_w = L->map; L->map_s = L->map; L->map_o = L->map; L->map_raw = L->map; L->map_rawnm = L->map;
    L->grid = 0; L->smem = 0;
    L->ntile = s.first_im2col ? -2 : 18; L->rows = 1;
    return 0;
  }
  if (s.first_im2col) {
    // dedicated kernel (first_conv.cuh): patches are built in shared memory from the raw input
    p.n_tiles = 1;
    p.yblocks = (Ho + kFirstR - 1) / kFirstR;
    p.positions = p.T * p.yblocks * p.xblocks;
    p.total_tiles = p.positions;
    if (set_tile_div(p)) return 1;
    p.wpack = sd.wpack; p.bias = sd.bias;
    p.flags = (s.relu ? EPI_RELU : EPI_RELU6) | (bf16 ? EPI_BF16 : 0);
    const int oc = s.store_c ? s.store_c : s.cout;     // 32: only the first unit of every row is stored
    p.out = io.out; p.out_C = oc; p.out_H = Ho; p.out_W = Wo; p.out_C_log2 = (oc == 32) ? 5 : 6;
    p.out_pitch_C = oc;
    p.out_frame_stride = (long long)Ho * Wo * oc;
    L->cta2 = 0;
    if (make_map_pix(&L->map_o, io.out, io.out_T ? io.out_T : io.T,
                     io.out_T ? io.out_T_stride : p.out_frame_stride * 2, Ho, Wo, oc, 0, 32, 32,
                     CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    p.tma_out = 1; p.stg_bytes_per_warp = 2 * kStageBytesPerWarp;
    L->map = L->map_o; L->map_w = L->map_o; L->map_s = L->map_o;
    L->map_raw = L->map_o; L->map_rawnm = L->map_o; L->first_raw_tma = 0;   // set per call (input pointer)
    L->grid = std::min(p.total_tiles, num_sms());
    L->smem = kFirstSmem;
    L->ntile = -1; L->rows = kFirstR;
    return 0;
  }
  if (s.final_out) {
    // dedicated kernel (final_conv.cuh): tiles of 4 output rows, filter bank resident
    p.n_tiles = 1;
    p.positions = p.T * p.yblocks * p.xblocks;
    p.total_tiles = p.positions;
    if (set_tile_div(p)) return 1;
    p.wpack = sd.wpack; p.bias = sd.bias;
    p.flags = EPI_FINAL | (bf16 ? EPI_BF16 : 0);
    p.out = io.out; p.out_C = 3; p.out_H = Ho; p.out_W = Wo;
    p.skip = io.skip; p.skip_C = io.skip_C; p.skip_frame_stride = io.skip_frame_stride;
    if (!io.skip) return fail("stage needs a skip tensor");
    L->cta2 = 0;
    if (make_map_halo(&L->map, io.in, io.T, io.H, io.W, s.cin, s.rows)) return 1;
    L->map_w = L->map;
    L->grid = std::min(p.total_tiles, num_sms());
    L->smem = s.pair_final ? kFinalSmemPair : kFinalSmem;
    L->ntile = s.pair_final ? 17 : 16; L->rows = s.rows;     // 16 / 17: final_conv_kernel, per pixel / per pair
    return 0;
  }
  p.positions = p.T * p.yblocks * p.xblocks;
  p.total_tiles = (cta2 ? (p.positions + 1) / 2 : p.positions) * p.n_tiles;
  if (set_tile_div(p)) return 1;
  p.mode = (s.stride == 2) ? 1 : 0;
  p.cin_total = s.split ? 2 * s.cin : s.cin;          // channels per pixel as stored (stride-2 column parity offset)
  p.phys_chunks = s.split ? 2 * s.cin / kChunk : 0;
  p.w_stage_bytes = (uint32_t)s.ntile * 128u / (cta2 ? 2u : 1u);
  p.w_rows_cta = s.ntile / 2;
  const bool stacked = s.stacked;
  if (stacked && !cta2) return fail("stacked filter layout needs the CTA-pair kernel");
  if (stacked) { p.w_stage_bytes = 192u * 128u; p.w_rows_cta = 192; }
  const size_t budget = kSmemOptIn - 1024 - kStagingBytes;   // minus alignment slack and staging
  if (stacked) p.mode = 2;
  // stride 2: sub-plane boxes (conv_tc.cuh PIPE 4) unless BSVD_B200_S2_TAPS=1 asks for the per-tap boxes
  static const int s2_taps = [] { const char* e = getenv("BSVD_B200_S2_TAPS"); return (e && e[0] == '1') ? 1 : 0; }();
  const bool s2_boxes = p.mode == 1 && cta2 && ((!s2_taps && !desc_variant) || s.pair_s2);
  if (s.pair_s2 && !cta2) return fail("the pair stride-2 stage needs the CTA-pair kernel");
  if (p.mode != 1) {
    p.a_tx_bytes = (uint32_t)(s.rows + 2) * kRowBytes;
    p.a_stage_bytes = (uint32_t)align_up(p.a_tx_bytes, 1024);
    p.a_stages = 2;
    const int total_w = s.cin_chunks * s.ntaps();
    const size_t left = budget - (size_t)p.a_stages * p.a_stage_bytes;
    if (s.n_tiles() == 1 && (size_t)total_w * p.w_stage_bytes <= left && total_w <= kMaxStages) {
      p.w_resident = 1;
      p.w_stages = total_w;
    } else {
      p.w_resident = 0;
      p.w_stages = (int)std::min<size_t>(kMaxStages, left / p.w_stage_bytes);
      if (p.w_stages < 2) return fail("shared memory budget too small for weight stages");
    }
  } else if (s2_boxes) {
    // one box per input sub-plane: ring slots sized for the largest, (R+1) rows x 129 px
    p.mode = s.pair_s2 ? 5 : 4;
    p.a_tx_bytes = (uint32_t)(s.rows + 1) * kS2BoxPx * 128u;
    p.a_stage_bytes = (uint32_t)align_up(p.a_tx_bytes, 1024);
    p.a_stages = 3;
    const size_t left = budget - (size_t)p.a_stages * p.a_stage_bytes;
    p.w_stages = (int)std::min<size_t>(kMaxStages, left / p.w_stage_bytes);
    p.w_resident = 0;
    if (p.w_stages < 3) return fail("shared memory budget too small for stride-2 stages");
  } else {
    p.a_tx_bytes = (uint32_t)s.rows * kRunPx * 128u;
    p.a_stage_bytes = p.a_tx_bytes;
    const size_t per = (size_t)p.a_stage_bytes + p.w_stage_bytes;
    int st = (int)std::min<size_t>(kMaxStages, budget / per);
    if (st < 2) return fail("shared memory budget too small for stride-2 stages");
    p.a_stages = st; p.w_stages = st; p.w_resident = 0;
  }
  p.desc_variant = desc_variant;
  static const int no_skip_pf = [] { const char* e = getenv("BSVD_B200_NO_SKIP_PF"); return (e && e[0] == '1') ? 1 : 0; }();
  if (no_skip_pf) p.desc_variant |= 256;   // debug: no L2 warm-up of the skip operand
  p.wpack = sd.wpack;
  p.bias = sd.bias;
  int flags = 0;
  if (s.relu6) flags |= EPI_RELU6;
  if (s.relu) flags |= EPI_RELU;
  if (s.pixshuf) flags |= EPI_PIXSHUF;
  if (s.skip && !skip_mma) flags |= EPI_SKIP;
  if (s.shift) flags |= EPI_SHIFT;
  if (s.resid_in) flags |= EPI_RESID_IN;
  if (s.final_out) flags |= EPI_FINAL;
  if (bf16) flags |= EPI_BF16;
  if (io.zero_future) flags |= EPI_ZERO_FUTURE;
  p.flags = flags;
  p.out = io.out; p.out_prev = io.out_prev; p.out_next = io.out_next; p.ring_mode = io.ring_mode;
  if (s.pixshuf) { p.out_C = s.cout / 4; p.out_H = 2 * Ho; p.out_W = 2 * Wo; }
  else { p.out_C = s.final_out ? 3 : s.cout; p.out_H = Ho; p.out_W = Wo; }
  p.out_pitch_C = s.split ? 2 * p.out_C : p.out_C;
  p.out_frame_stride = (long long)p.out_H * p.out_W * p.out_pitch_C;
  p.out_C_log2 = 0;
  while ((1 << p.out_C_log2) < p.out_C) ++p.out_C_log2;
  if ((1 << p.out_C_log2) != p.out_C || p.out_C < 32) return fail("output channels must be a power of two >= 32");
  if (s.skip && io.skip && io.skip_C != p.out_C) return fail("skip tensor must have the output's channel count");
  p.skip = io.skip; p.skip_C = io.skip_C; p.skip_frame_stride = io.skip_frame_stride;
  p.resid_in = io.resid_in; p.resid_C = io.resid_C; p.aux_out = io.aux_out;
  p.fold = p.out_C / 8;
  p.stg_bytes_per_warp = kStageBytesPerWarp * stg_bufs;
  p.skip_mma = skip_mma ? skip_blocks : 0;
  p.skip_blocks = s.ntile / 64;
  p.tma_out = tma_out ? 1 : 0;
  if ((s.skip || s.final_out) && !io.skip) return fail("stage needs a skip tensor");
  if (s.resid_in && !io.resid_in) return fail("stage needs the raw input for the residual");

  // pair stride-2: the [T][H/2][2][W/2][2*32] view of the 32-channel tensor has the pairs as its pixels
  const int cin_map = s.first_im2col ? kChunk : (s.pair_s2 ? 32 : (s.split ? 2 * s.cin : s.cin));
  int rc = (s.stride != 2) ? make_map_halo(&L->map, io.in, io.T, io.H, io.W, cin_map, s.rows)
           : (p.mode >= 4) ? make_map_s2(&L->map, io.in, io.T, io.H, io.W, cin_map, s.rows + 1, kS2BoxPx)
                           : make_map_s2(&L->map, io.in, io.T, io.H, io.W, cin_map, s.rows);
  if (rc) return rc;
  if (cta2) {
    rc = make_map_w(&L->map_w, sd.wpack, s.pack_elems() / kChunk, p.w_rows_cta);
    if (rc) return rc;
  } else {
    L->map_w = L->map;   // unused
  }
  L->map_s = L->map; L->map_o = L->map;   // unused unless set below
  if (p.mode >= 4) {   // even input rows: R-row boxes (the skip map slot is free: stride-2 stages have no skip)
    rc = make_map_s2(&L->map_s, io.in, io.T, io.H, io.W, cin_map, s.rows, kS2BoxPx);
    if (rc) return rc;
  }
  const long long frame_bytes = p.out_frame_stride * 2;
  if (skip_mma) {
    if (p.w_stage_bytes != 16384u || p.w_resident || s.cin_chunks * s.ntaps() < 4 * skip_blocks)
      return fail("skip blocks need 16 KB filter-ring slots and at least four slabs per block");
    rc = make_map_pix(&L->map_s, io.skip, io.skip_T ? io.skip_T : io.T,
                      io.skip_T ? io.skip_T_stride : frame_bytes, p.out_H, p.out_W, p.out_pitch_C, 1, kChunk,
                      kRunPx, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (tma_out) {
    rc = make_map_pix(&L->map_o, io.out, io.out_T ? io.out_T : io.T, io.out_T ? io.out_T_stride : frame_bytes,
                      p.out_H, p.out_W, p.out_pitch_C, s.pixshuf ? 1 : 0, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  L->grid = cta2 ? 2 * std::min(p.total_tiles, num_sms() / 2) : std::min(p.total_tiles, num_sms());
  L->smem = 1024 + (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.w_stages * p.w_stage_bytes +
            kStagingBytes;
  L->ntile = s.ntile; L->rows = s.rows;
  return 0;
}

static int launch_first(const StageLaunch& L, cudaStream_t st) {
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(first_conv_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirstSmem));
    CUDA_TRY(cudaFuncSetAttribute(first_conv_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirstSmem));
    CUDA_TRY(cudaFuncSetAttribute(first_conv_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirstSmem));
    CUDA_TRY(cudaFuncSetAttribute(first_conv_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirstSmem));
    CUDA_TRY(cudaFuncSetAttribute(first_conv_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirstSmemRaw));
    CUDA_TRY(cudaFuncSetAttribute(first_conv_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirstSmemRaw));
    CUDA_TRY(cudaFuncSetAttribute(first_conv_kernel<false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirstSmemRaw));
    CUDA_TRY(cudaFuncSetAttribute(first_conv_kernel<true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirstSmemRaw));
    attr_done[dev & 63] = true;
  }
  if (!L.first_in) return fail("first stage launched without an input pointer");
  const bool raw = L.first_raw_tma && !L.first_u8;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(raw ? kFirstThreadsRaw : kFirstThreads);
  cfg.dynamicSmemBytes = raw ? kFirstSmemRaw : L.smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  float* norm = L.first_u8 ? L.first_norm : nullptr;
  const bool bf = (L.p.flags & EPI_BF16) != 0;
  ConvParams pr = L.p;
  if (raw) pr.stg_bytes_per_warp = kStageBytesPerWarp;     // RAW instances: one staging tile per epilogue warp
  const bool act6 = (L.p.flags & EPI_RELU6) && !(L.p.flags & EPI_RELU);
#define BSVD_FIRST(B, U, R, ...) \
  CUDA_TRY(cudaLaunchKernelEx(&cfg, first_conv_kernel<B, U, R, ##__VA_ARGS__>, L.first_in, L.first_nmap, L.first_inc, L.map_o, pr, norm, L.map_raw, L.map_rawnm))
  if (raw && act6) { if (bf) BSVD_FIRST(true, false, true, true); else BSVD_FIRST(false, false, true, true); }
  else if (raw) { if (bf) BSVD_FIRST(true, false, true); else BSVD_FIRST(false, false, true); }
  else if (L.first_u8) { if (bf) BSVD_FIRST(true, true, false); else BSVD_FIRST(false, true, false); }
  else { if (bf) BSVD_FIRST(true, false, false); else BSVD_FIRST(false, false, false); }
#undef BSVD_FIRST
  return 0;
}

static int launch_stage(const StageLaunch& L, cudaStream_t st) {
  if (L.ntile == -2 || L.ntile == 18) {
    // fp32-grade mode, first / last conv on the CUDA cores (split_edge.cuh)
    const ConvParams& p = L.p;
    const dim3 grid((p.W + kEdgeThreads - 1) / kEdgeThreads, p.H, p.T);
    if (L.ntile == -2) {
      if (!L.first_in) return fail("first stage launched without an input pointer");
      if (L.first_u8) return fail("fp32-grade mode: uint8 frame input is not implemented");
      first_conv_split_kernel<<<grid, kEdgeThreads, 0, st>>>(
          L.first_in, L.first_nmap, L.first_inc, reinterpret_cast<const float*>(p.wpack), p.bias,
          reinterpret_cast<uint16_t*>(p.out), p.T, p.H, p.W, p.src_H, p.src_W, p.use_sigma, p.sigma_const,
          (p.flags & EPI_RELU6) ? 1 : 0);
    } else {
      final_conv_split_kernel<<<grid, kEdgeThreads, 0, st>>>(
          reinterpret_cast<const uint16_t*>(L.in_ptr), reinterpret_cast<const float*>(p.wpack), p.bias,
          reinterpret_cast<const float*>(p.skip), p.out, p.T, p.H, p.W, p.src_H, p.src_W, p.clamp01, p.out_u8, p.u8_bgr);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  if (L.ntile == -1) return launch_first(L, st);
  if (L.ntile == 16 || L.ntile == 17) {
    static std::atomic<bool> attr_done[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(final_conv_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinalSmem));
      CUDA_TRY(cudaFuncSetAttribute(final_conv_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinalSmem));
      CUDA_TRY(cudaFuncSetAttribute(final_conv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinalSmemPair));
      CUDA_TRY(cudaFuncSetAttribute(final_conv_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinalSmemPair));
      attr_done[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L.smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const bool bf = (L.p.flags & EPI_BF16) != 0;
    if (L.ntile == 17) {
      if (bf) CUDA_TRY(cudaLaunchKernelEx(&cfg, final_conv_kernel<true, true>, L.map, L.p));
      else CUDA_TRY(cudaLaunchKernelEx(&cfg, final_conv_kernel<false, true>, L.map, L.p));
    } else {
      if (bf) CUDA_TRY(cudaLaunchKernelEx(&cfg, final_conv_kernel<true, false>, L.map, L.p));
      else CUDA_TRY(cudaLaunchKernelEx(&cfg, final_conv_kernel<false, false>, L.map, L.p));
    }
    return 0;
  }
  if (L.ntile == 64 && L.rows == 2) return launch_one<64, 2>(L, st);
  if (L.ntile == 128 && L.rows == 2) return launch_one<128, 2>(L, st);
  if (L.ntile == 256 && L.rows == 1) return launch_one<256, 1>(L, st);
  if (L.ntile == 256 && L.rows == 2) return launch_one<256, 2>(L, st);
  return fail("no kernel instance for NTILE=%d R=%d", L.ntile, L.rows);
}

static int upload_stage(StageDev& sd, const float* w, const float* b, int bf16) {
  std::vector<uint16_t> pack;
  std::vector<float> bias;
  pack_weights(sd.spec, w, b, bf16, pack, bias);
  if (!sd.wpack) CUDA_TRY(cudaMalloc(&sd.wpack, pack.size() * 2));
  if (!sd.bias) CUDA_TRY(cudaMalloc((void**)&sd.bias, bias.size() * 4));
  CUDA_TRY(cudaMemcpy(sd.wpack, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(sd.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  sd.bias_h = bias;
  sd.loaded = true;
  return 0;
}
static void free_stage(StageDev& sd) {
  if (sd.wpack) cudaFree(sd.wpack);
  if (sd.bias) cudaFree(sd.bias);
  sd.wpack = nullptr; sd.bias = nullptr; sd.loaded = false;
}

}  // namespace bsvd

using namespace bsvd;

// ================================================================================================
// network handle
// ================================================================================================
// Streaming ring buffers of one DenBlock (sizes in frames; the reference keeps the same state in
// BiBufferConv.left/center (bsvd_arch.py:72-74,112-113) and the MemSkip FIFOs (:308-322)):
//   shifted tensors feeding a BiBufferConv live in 3-slot rings (frames f-1, f, f+1),
//   skip2 (x0) and skip1 (block input) wait 8 steps -> 9 slots, skip3 (x1) waits 4 steps -> 5 slots.
enum { kRingP = 0, kRingA, kRingX0, kRingB2, kRingB3, kRingX1, kRingB5, kRingB6, kRingB7, kRingB8,
       kRingB9, kRingB10, kRingB11, kRingB12, kRingB13, kRingB14, kRingM, kNumRings };
static const int kRingSlots[kNumRings] = {1, 1, 9, 3, 3, 5, 3, 3, 3, 3, 1, 3, 3, 1, 1, 1, 9};
static const int kRingRes[kNumRings] = {1, 1, 1, 2, 2, 2, 4, 4, 4, 4, 4, 2, 2, 2, 1, 1, 1};  // 1 full, 2 half, 4 quarter
// per layer: delay (steps) inside a DenBlock, input ring, output ring
static const int kLayerDelay[16] = {0, 0, 0, 1, 2, 2, 3, 4, 5, 6, 6, 7, 8, 8, 8, 8};
static const int kLayerIn[16] = {kRingP, kRingA, kRingX0, kRingB2, kRingB3, kRingX1, kRingB5, kRingB6,
                                 kRingB7, kRingB8, kRingB9, kRingB10, kRingB11, kRingB12, kRingB13,
                                 kRingB14};
static const int kLayerOut[16] = {kRingA, kRingX0, kRingB2, kRingB3, kRingX1, kRingB5, kRingB6, kRingB7,
                                  kRingB8, kRingB9, kRingB10, kRingB11, kRingB12, kRingB13, kRingB14,
                                  kRingM};
static const int kBlockDelay = 8;   // latency of one DenBlock in steps (8 BiBufferConv each)
struct bsvd_handle {
  bsvd_config cfg;
  int bf16 = 0;
  int pair32 = 0;               // c32 configurations in their native layout (32 channels at full resolution)
  int split = 0;                // fp32-grade mode: every activation tensor holds [hi | lo], 2x the channels below
  int cp[3] = {64, 128, 256};   // channels of the full / half / quarter resolution tensors as stored
  StageDev stages[BSVD_NUM_LAYERS];
  // ---- clip-mode workspace / plan (rebuilt when T,H,W change) ----
  int pT = 0, pH = 0, pW = 0, p_inc = 0;
  const float* p_in = nullptr; const float* p_nmap = nullptr; float* p_out = nullptr;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  uint16_t *bufP = nullptr, *bufA = nullptr, *bufX0 = nullptr, *bufM = nullptr;
  uint16_t *bufH0 = nullptr, *bufH1 = nullptr, *bufX1 = nullptr, *bufQ0 = nullptr, *bufQ1 = nullptr;
  uint16_t* bufS = nullptr;   // compact [T][H][W][4] copy of temp1's output channels 0..3 (skip1 of temp2)
  std::vector<StageLaunch> plan;
  int last_launches = 0;
  unsigned long long weights_epoch = 0;   // bumped by bsvd_set_weights (cached CUDA graphs are rebuilt)
  int device = 0;                     // CUDA device the handle (weights, workspaces, tensor maps) lives on
  unsigned* d_overflow = nullptr;     // sticky flag: a stored 16-bit activation was not finite
  float* d_norm = nullptr;            // normalised fp32 planes of the uint8 input (temp1 residual)
  size_t d_norm_bytes = 0;
  // per-stage event timing
  int profiling = 0;
  std::vector<std::vector<cudaEvent_t>> ev_sets;   // each: BSVD_NUM_STAGES + 1 events
  int ev_used = 0;
  // pinned/device staging for the host entry
  float* d_in = nullptr; float* d_nmap = nullptr; float* d_out = nullptr;
  size_t d_in_bytes = 0, d_nmap_bytes = 0, d_out_bytes = 0;
  // ---- pipelined host entry (bsvd_forward_clip_host_async): 2-deep staging ----
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[2] = {}, ev_comp[2] = {}, ev_d2h[2] = {};
  float* pin[2] = {}; float* pnm[2] = {}; float* pout[2] = {};
  size_t pin_bytes[2] = {}, pnm_bytes[2] = {}, pout_bytes[2] = {};
  unsigned long long host_calls = 0;
  // ---- streaming mode (feedin_one_element) ----
  struct StreamLayer {
    StageLaunch tmpl;                 // planned for T=1 on ring slot 0
    std::vector<CUtensorMap> maps;    // one per input ring slot
    std::vector<CUtensorMap> maps2;   // stride-2 sub-plane pipeline: the second (R-row) box map per slot
    std::vector<CUtensorMap> raw_maps;   // first conv: one raw-plane map per raw ring slot
  };
  struct Stream {
    int H = 0, W = 0;
    long long step = 0;               // pushes so far in this stream
    long long n_in = 0;               // real frames fed
    bool ended = false;               // a NULL frame has been pushed
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // rings[blk][k] : device base of ring k of DenBlock blk; see kRing* below
    uint8_t* ring[2][kNumRings] = {};
    float* raw = nullptr;             // fp32 [9][4][H][W] ring of the raw network input
    uint8_t* aux = nullptr;           // 16-bit [9][H][W][4] ring: temp1 output channels 0..3 (skip1 of temp2)
    StreamLayer layers[BSVD_NUM_LAYERS];
    // steady-state pushes replay one CUDA graph per ring phase (all ring sizes divide kPhases)
    static constexpr int kPhases = 45;          // lcm(1, 3, 5, 9)
    cudaGraphExec_t gexec[kPhases] = {};
    unsigned long long gepoch = 0;              // weights_epoch the graphs were captured with
    cudaStream_t cap = nullptr;                 // capture stream (the caller's may be the legacy stream)
    float* out_slot = nullptr;                  // [3][H][W] fp32: where a replayed graph leaves its frame
    long long graph_replays = 0;
  } stream;
};

// Options of ONE call (the fused caller entries bsvd_denoise_clip / _u8 set them); they travel as an
// argument so that no call-scoped state is ever parked on the handle.
struct CallOpts {
  int src_H = 0, src_W = 0;           // raw-image view of the first / residual / last kernels (0 = same)
  int use_sigma = 0, clamp01 = 0;
  float sigma_const = 0.f;
  int u8_io = 0, u8_bgr = 0;          // uint8 HWC frames in and out
  int seg_T = 0;                      // > 0: the T frames are independent clips of seg_T frames each
};

// Every entry point runs on the device the handle was created on (weights, workspaces and tensor maps
// live there); a call made with another device current is an error, never a silent remote access.
static int check_device(const bsvd_handle* h) {
  int dev = -1;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev != h->device)
    return fail("handle belongs to CUDA device %d but device %d is current (create one handle per "
                "device; bsvd_b200.arch.BSVD does this when the module is moved)", h->device, dev);
  return 0;
}
static int check_dev_ptr(const bsvd_handle* h, const void* p, const char* what) {
  if (!p) return 0;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  if ((a.type == cudaMemoryTypeDevice) && a.device != h->device)
    return fail("%s lives on CUDA device %d, the handle on device %d", what, a.device, h->device);
  if (a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeUnregistered)
    return fail("%s is a host pointer; this entry point takes device memory", what);
  return 0;
}

static inline int pad64(int c) { return (c + 63) / 64 * 64; }
static void build_specs(bsvd_handle* h) {
  // DenBlock (bsvd_arch.py:325-396).  BSVD-64: chns 64/128/256, temp1 4->64, temp2 64->3.  The c32
  // configurations (chns 32/64/128, interm_ch 30, mid_ch 32) run on the same kernels with every
  // channel count below 64 zero-padded to 64 (weights, bias and hence activations of the padded
  // channels are exactly zero; ReLU keeps them there).
  for (int blk = 0; blk < 2; ++blk) {
    auto S = [&](int l) -> StageSpec& { return h->stages[blk * 16 + l].spec; };
    const int c0 = h->cfg.chns[0], c1 = h->cfg.chns[1], c2 = h->cfg.chns[2], ci = h->cfg.interm_ch;
    const int in_ch = blk == 0 ? h->cfg.in_ch : h->cfg.mid_ch;
    const int out_ch = blk == 0 ? h->cfg.mid_ch : h->cfg.out_ch;
    const bool r6 = h->cfg.act_relu6 != 0;
    // (layer, logical cin, logical cout): physical = padded to 64, PixelShuffle convs per sub-pixel
    auto set = [&](int l, int cin, int cout, bool act, bool ps = false) {
      StageSpec& s = S(l);
      s = StageSpec();
      s.cin_l = cin; s.cout_l = cout;
      s.cin = pad64(cin);
      s.cout = ps ? 4 * pad64(cout / 4) : pad64(cout);
      s.relu6 = act && r6; s.relu = act && !r6; s.pixshuf = ps;
    };
    set(0, in_ch, ci, true);
    if (blk == 0) { S(0).first_im2col = true; S(0).cin = in_ch; }     // raw 3/4-channel input
    set(1, ci, c0, true);
    set(2, c0, c1, true); S(2).stride = 2; S(2).shift = true;
    set(3, c1, c1, true); S(3).shift = true;
    set(4, c1, c1, true);
    set(5, c1, c2, true); S(5).stride = 2; S(5).shift = true;
    set(6, c2, c2, true); S(6).shift = true;
    set(7, c2, c2, true); S(7).shift = true;
    set(8, c2, c2, true); S(8).shift = true;
    set(9, c2, c2, true);
    set(10, c2, c1 * 4, false, true); S(10).skip = true; S(10).shift = true;
    set(11, c1, c1, true); S(11).shift = true;
    set(12, c1, c1, true);
    set(13, c1, c0 * 4, false, true); S(13).skip = true;
    set(14, c0, c0, true);
    set(15, c0, out_ch, false);
    S(15).resid_in = (blk == 0); S(15).final_out = (blk == 1);
    if (blk == 1) S(15).cout = out_ch;                                // 3 output planes, no padding
    if (h->split)
      for (int l = 0; l < 16; ++l) S(l).split = true;
    if (h->pair32) {
      // native layout of the c32 configurations (see StageSpec::pairx): full-resolution tensors hold 32
      // channels per pixel and are processed as pixel pairs; nothing below full resolution is padded
      S(0).cout = 64; S(0).store_c = 32;                       // first conv: 64-column GEMM, 32 channels stored
      if (blk == 1) { S(0).pairx = true; S(0).store_c = 0; }
      S(1).pairx = true;
      S(2).pair_s2 = true;                                     // cin 64 = one pair, cout = c1
      S(13).cout = 4 * (c0);                                   // PixelShuffle to 32 channels: no padding
      S(14).pairx = true;
      if (blk == 0) S(15).pairx = true; else S(15).pair_final = true;
    }
    for (int l = 0; l < 16; ++l) S(l).derive();
  }
  h->cp[0] = h->pair32 ? 32 : pad64(h->cfg.chns[0]);
  h->cp[1] = pad64(h->cfg.chns[1]); h->cp[2] = pad64(h->cfg.chns[2]);
  if (h->split)            // [hi | lo]: twice the channels per pixel in every activation tensor
    for (int i = 0; i < 3; ++i) h->cp[i] *= 2;
}

extern "C" { static void free_stream(bsvd_handle* h); }
static void free_workspace(bsvd_handle* h) {
  if (h->ws) cudaFree(h->ws);
  h->ws = nullptr; h->ws_bytes = 0; h->plan.clear();
  h->pT = h->pH = h->pW = 0;
}

// Clip-mode schedule: layer by layer over all T frames (the TSN order; identical arithmetic to the
// streaming order), temporal folds exchanged through shifted stores.
static int build_clip_plan(bsvd_handle* h, const float* in, const float* nmap, float* out, int T,
                           int in_c, int H, int W) {
  const bool same_shape = (h->pT == T && h->pH == H && h->pW == W && h->ws);
  if (same_shape && !h->plan.empty()) return 0;   // in/out pointers are patched at launch time
  if (!same_shape) {
    free_workspace(h);
    const size_t full = (size_t)T * H * W * h->cp[0] * 2;
    const size_t half = (size_t)T * (H / 2) * (W / 2) * h->cp[1] * 2;
    const size_t quar = (size_t)T * (H / 4) * (W / 4) * h->cp[2] * 2;
    const size_t fa = align_up(full, 1024), ha = align_up(half, 1024), qa = align_up(quar, 1024);
    const size_t sa = align_up((size_t)T * H * W * 4 * (h->split ? 4 : 2), 1024);   // compact skip1 copy (fp32 when split)
    h->ws_bytes = 4 * fa + 3 * ha + 2 * qa + sa;
    CUDA_TRY(cudaMalloc(&h->ws, h->ws_bytes));
    uint8_t* b = reinterpret_cast<uint8_t*>(h->ws);
    h->bufP = (uint16_t*)b; b += fa;
    h->bufA = (uint16_t*)b; b += fa;
    h->bufX0 = (uint16_t*)b; b += fa;
    h->bufM = (uint16_t*)b; b += fa;
    h->bufH0 = (uint16_t*)b; b += ha;
    h->bufH1 = (uint16_t*)b; b += ha;
    h->bufX1 = (uint16_t*)b; b += ha;
    h->bufQ0 = (uint16_t*)b; b += qa;
    h->bufQ1 = (uint16_t*)b; b += qa;
    h->bufS = (uint16_t*)b; b += sa;
    h->pT = T; h->pH = H; h->pW = W;
  }
  h->p_in = in; h->p_nmap = nmap; h->p_out = out; h->p_inc = in_c;
  h->plan.assign(BSVD_NUM_LAYERS, StageLaunch());
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  const long long fs_full = (long long)H * W * h->cp[0], fs_half = (long long)H2 * W2 * h->cp[1];
  for (int blk = 0; blk < 2; ++blk) {
    auto plan = [&](int l, const void* src, int sh, int sw, void* dst, const void* skip = nullptr,
                    int skip_C = 0, long long skip_fs = 0) -> int {
      StageIO io;
      const StageSpec& sp = h->stages[blk * 16 + l].spec;
      // pair stages see an image of W/2 pixel pairs (same memory)
      io.in = src; io.T = T; io.H = sh; io.W = (sp.pairx || sp.pair_final) ? sw / 2 : sw; io.out = dst;
      io.skip = skip; io.skip_C = skip_C; io.skip_frame_stride = skip_fs;
      io.overflow = h->d_overflow;
      if (blk == 0 && l == 15) { io.resid_in = in; io.resid_C = in_c; io.aux_out = h->bufS; }
      return plan_stage(h->stages[blk * 16 + l], io, h->bf16, 0, &h->plan[blk * 16 + l]);
    };
    const void* src0 = (blk == 0) ? (const void*)h->bufP : (const void*)h->bufM;
    int rc = 0;
    rc |= plan(0, src0, H, W, h->bufA);
    rc |= plan(1, h->bufA, H, W, h->bufX0);
    rc |= plan(2, h->bufX0, H, W, h->bufH0);
    rc |= plan(3, h->bufH0, H2, W2, h->bufH1);
    rc |= plan(4, h->bufH1, H2, W2, h->bufX1);
    rc |= plan(5, h->bufX1, H2, W2, h->bufQ0);
    rc |= plan(6, h->bufQ0, H4, W4, h->bufQ1);
    rc |= plan(7, h->bufQ1, H4, W4, h->bufQ0);
    rc |= plan(8, h->bufQ0, H4, W4, h->bufQ1);
    rc |= plan(9, h->bufQ1, H4, W4, h->bufQ0);
    const int lc0 = h->split ? h->cp[0] / 2 : h->cp[0], lc1 = h->split ? h->cp[1] / 2 : h->cp[1];   // logical channels
    rc |= plan(10, h->bufQ0, H4, W4, h->bufH0, h->bufX1, lc1, fs_half);
    rc |= plan(11, h->bufH0, H2, W2, h->bufH1);
    rc |= plan(12, h->bufH1, H2, W2, h->bufH0);
    rc |= plan(13, h->bufH0, H2, W2, h->bufA, h->bufX0, lc0, fs_full);
    rc |= plan(14, h->bufA, H, W, h->bufP);
    if (blk == 0) rc |= plan(15, h->bufP, H, W, h->bufM);
    else rc |= plan(15, h->bufP, H, W, out, h->bufS, 4, (long long)H * W * 4);
    if (rc) { h->plan.clear(); return 1; }
  }
  return 0;
}

static int check_hw(int T, int in_c, int H, int W, bool has_nmap, int net_in_ch = 4) {
  if (net_in_ch == 3) {
    if (in_c != 3 || has_nmap)
      return fail("blind model: input must have 3 channels and no noise map (got in_c=%d, noise_map=%d)",
                  in_c, (int)has_nmap);
    has_nmap = true;   // fall through the shape checks below
  }
  if (T < 1) return fail("T must be >= 1");
  if (H < 4 || W < 4 || (H % 4) || (W % 4))
    return fail("H and W must be multiples of 4 (got %dx%d); the reference fails at the skip add "
                "(bsvd_arch.py:402-406)", H, W);
  if (!((in_c == 4 && !has_nmap) || (in_c == 3 && has_nmap)))
    return fail("input must have 4 channels, or 3 channels plus a noise map (got in_c=%d, "
                "noise_map=%d)", in_c, (int)has_nmap);
  return 0;
}

extern "C" {

const char* bsvd_last_error(void) { return g_err.c_str(); }
const char* bsvd_version(void) { return "bsvd_b200 0.1 (sm_100a, tcgen05/TMEM/TMA)"; }

int bsvd_create(const bsvd_config* cfg, bsvd_handle** out) {
  if (!cfg || !out) return fail("null argument");
  const bool c64 = cfg->chns[0] == 64 && cfg->chns[1] == 128 && cfg->chns[2] == 256;
  const bool c32 = cfg->chns[0] == 32 && cfg->chns[1] == 64 && cfg->chns[2] == 128;
  if (!((c64 || c32) && cfg->mid_ch >= 3 && cfg->mid_ch <= 64 && cfg->interm_ch >= 1 &&
        cfg->interm_ch <= 64 && (cfg->in_ch == 4 || cfg->in_ch == 3) && cfg->out_ch == 3 &&
        (cfg->act_relu6 == 0 || cfg->act_relu6 == 1) && cfg->norm_none == 1))
    return fail("implemented on the GPU path: chns=[64,128,256] (options/test/bsvd_c64.yml) or "
                "[32,64,128] (options/train/0402_*_c32.yml, zero-padded to the 64-channel kernels), "
                "mid_ch<=64, interm_ch<=64, in_ch=4 (or 3 = blind), out_ch=3, norm='none', "
                "act='relu6'|'relu'; there is no CPU fallback");
  if (cfg->precision != BSVD_PREC_FP16 && cfg->precision != BSVD_PREC_BF16 && cfg->precision != BSVD_PREC_FP32X3)
    return fail("unknown precision %d", cfg->precision);
  if (cfg->precision == BSVD_PREC_FP32X3 && !(c64 && cfg->act_relu6 == 1))
    return fail("the fp32-grade mode (BSVD_PREC_FP32X3) implements the BSVD-64 configuration (options/test/bsvd_c64.yml)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: this library only runs on a B200 (sm_100a); no CPU fallback");
  if (cfg->device >= 0) CUDA_TRY(cudaSetDevice(cfg->device));
  int dev = 0, major = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail("device compute capability %d.x is not sm_100a (B200)", major);
  bsvd_handle* h = new bsvd_handle();
  h->cfg = *cfg;
  h->device = dev;
  h->bf16 = (cfg->precision == BSVD_PREC_BF16);
  h->split = (cfg->precision == BSVD_PREC_FP32X3) ? 1 : 0;
  {
    // BSVD_B200_C32_PADDED=1: the round-1 path (every channel count below 64 zero-padded to 64)
    const char* e = getenv("BSVD_B200_C32_PADDED");
    h->pair32 = (c32 && !(e && e[0] == '1') && use_cta2_default()) ? 1 : 0;
  }
  build_specs(h);
  if (cudaMalloc((void**)&h->d_overflow, sizeof(unsigned)) != cudaSuccess ||
      cudaMemset(h->d_overflow, 0, sizeof(unsigned)) != cudaSuccess) {
    delete h;
    return fail("cudaMalloc of the overflow flag failed");
  }
  *out = h;
  return 0;
}

int bsvd_destroy(bsvd_handle* h) {
  if (!h) return 0;
  free_workspace(h);
  free_stream(h);
  for (auto& s : h->stages) free_stage(s);
  for (auto& set : h->ev_sets)
    for (auto& e : set) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->pin[i]) cudaFree(h->pin[i]);
    if (h->pnm[i]) cudaFree(h->pnm[i]);
    if (h->pout[i]) cudaFree(h->pout[i]);
    if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
    if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
    if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
  }
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  if (h->d_norm) cudaFree(h->d_norm);
  if (h->d_overflow) cudaFree(h->d_overflow);
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_nmap) cudaFree(h->d_nmap);
  if (h->d_out) cudaFree(h->d_out);
  delete h;
  return 0;
}

int bsvd_layer_shape(const bsvd_handle* h, int layer, int* out_ch, int* in_ch, int* stride) {
  if (!h || layer < 0 || layer >= BSVD_NUM_LAYERS) return fail("bad layer index %d", layer);
  const StageSpec& s = h->stages[layer].spec;
  if (out_ch) *out_ch = s.cout_l;
  if (in_ch) *in_ch = s.cin_l;
  if (stride) *stride = s.stride;
  return 0;
}

int bsvd_set_weights(bsvd_handle* h, int layer, const float* w, const float* bias, int out_ch,
                     int in_ch) {
  if (!h || layer < 0 || layer >= BSVD_NUM_LAYERS) return fail("bad layer index %d", layer);
  if (!w) return fail("null weight pointer");
  StageDev& sd = h->stages[layer];
  if (sd.spec.cout_l != out_ch || sd.spec.cin_l != in_ch)
    return fail("layer %d expects weight [%d,%d,3,3], got [%d,%d,3,3]", layer, sd.spec.cout_l,
                sd.spec.cin_l, out_ch, in_ch);
  if (check_device(h)) return 1;
  // A forward enqueued earlier (on any stream) may still read the packed weights: every kernel loads
  // them before griddepcontrol.wait.  Weight updates are rare, so simply drain the device first.
  if (sd.loaded) CUDA_TRY(cudaDeviceSynchronize());
  if (upload_stage(sd, w, bias, h->bf16)) return 1;
  // The bias also travels in the kernel-parameter bank of every cached launch record (clip plan,
  // streaming templates): refresh those copies, or a reload after a forward would pair new weights
  // with old biases.
  auto refresh = [&](StageLaunch& L) {
    std::copy(sd.bias_h.begin(), sd.bias_h.end(), L.p.bias_c);
  };
  if (!h->plan.empty()) refresh(h->plan[layer]);
  if (h->stream.ws) refresh(h->stream.layers[layer].tmpl);
  ++h->weights_epoch;
  return 0;
}

int bsvd_set_profiling(bsvd_handle* h, int on) {
  if (!h) return fail("null handle");
  h->profiling = on ? 1 : 0;
  if (on) h->ev_used = 0;
  return 0;
}

int bsvd_get_stage_ms(bsvd_handle* h, float* ms, int n, int* passes) {
  if (!h || !ms || n < BSVD_NUM_STAGES) return fail("bad arguments");
  for (int i = 0; i < BSVD_NUM_STAGES; ++i) ms[i] = 0.f;
  for (int s = 0; s < h->ev_used; ++s) {
    auto& set = h->ev_sets[s];
    CUDA_TRY(cudaEventSynchronize(set[BSVD_NUM_STAGES]));
    for (int i = 0; i < BSVD_NUM_STAGES; ++i) {
      float t = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&t, set[i], set[i + 1]));
      ms[i] += t;
    }
  }
  if (passes) *passes = h->ev_used;
  return 0;
}

int bsvd_stage_info(const bsvd_handle* h, int stage, int* cin, int* cout, int* stride, int* ntile,
                    int* rows) {
  if (!h || stage < 1 || stage > BSVD_NUM_LAYERS) return fail("bad stage index %d", stage);
  const StageSpec& s = h->stages[stage - 1].spec;
  if (cin) *cin = s.cin_l;
  if (cout) *cout = s.cout_l;
  if (stride) *stride = s.stride;
  if (ntile) *ntile = s.ntile;
  if (rows) *rows = s.rows;
  return 0;
}

int bsvd_last_launch_count(const bsvd_handle* h) { return h ? h->last_launches : 0; }

int bsvd_overflow_flag(bsvd_handle* h, int* flag, int reset) {
  if (!h || !flag) return fail("null argument");
  if (check_device(h)) return 1;
  unsigned v = 0;
  CUDA_TRY(cudaMemcpy(&v, h->d_overflow, sizeof(v), cudaMemcpyDeviceToHost));   // waits for prior work
  if (reset && v) CUDA_TRY(cudaMemset(h->d_overflow, 0, sizeof(v)));
  *flag = v ? 1 : 0;
  return 0;
}
size_t bsvd_workspace_bytes(const bsvd_handle* h) { return h ? h->ws_bytes : 0; }

static int forward_clip_impl(bsvd_handle* h, const float* in, const float* noise_map, float* out, int T,
                             int in_c, int H, int W, void* stream, const CallOpts& o) {
  if (!h || !in || !out) return fail("null argument");
  if (check_device(h)) return 1;
  if (check_hw(T, in_c, H, W, noise_map != nullptr || o.use_sigma, h->cfg.in_ch)) return 1;
  for (int l = 0; l < BSVD_NUM_LAYERS; ++l)
    if (!h->stages[l].loaded) return fail("weights of layer %d were never set", l);
  if (build_clip_plan(h, in, noise_map, out, T, in_c, H, W)) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::vector<cudaEvent_t>* evs = nullptr;
  if (h->profiling) {
    if (h->ev_used == (int)h->ev_sets.size()) {
      std::vector<cudaEvent_t> set(BSVD_NUM_STAGES + 1);
      for (auto& e : set) CUDA_TRY(cudaEventCreate(&e));
      h->ev_sets.push_back(set);
    }
    evs = &h->ev_sets[h->ev_used++];
    CUDA_TRY(cudaEventRecord((*evs)[0], st));
  }
  // stage 0 (input staging) is fused into temp1.inc.convblock.0 (first_conv.cuh)
  h->plan[0].first_in = in; h->plan[0].first_nmap = noise_map; h->plan[0].first_inc = in_c;
  h->plan[0].first_u8 = o.u8_io; h->plan[0].first_norm = o.u8_io ? h->d_norm : nullptr;
  {
    // raw fp32 planes through TMA unless the caller view is reflect-padded / uint8 / misaligned
    StageLaunch& L0 = h->plan[0];
    const bool same_view = (!o.src_H || o.src_H == H) && (!o.src_W || o.src_W == W);
    L0.first_raw_tma = (!o.u8_io && !o.use_sigma && same_view && raw_tma_ok(in, noise_map, W)) ? 1 : 0;
    if (L0.first_raw_tma) {
      if (make_map_raw(&L0.map_raw, in, (long long)T * in_c, H, W, in_c)) return 1;
      if (noise_map && make_map_raw(&L0.map_rawnm, noise_map, T, H, W, 1)) return 1;
    }
  }
  if (evs) CUDA_TRY(cudaEventRecord((*evs)[1], st));
  int launches = 0;
  h->plan[15].p.resid_in = o.u8_io ? h->d_norm : in;   // temp1 residual reads the raw input (skip1)
  h->plan[15].p.resid_C = o.u8_io ? 3 : in_c;
  h->plan[BSVD_NUM_LAYERS - 1].p.out = out;
  // raw-image view of the first / residual / last kernels (the fused caller entries)
  for (int l : {0, 15, BSVD_NUM_LAYERS - 1}) {
    ConvParams& q = h->plan[l].p;
    q.src_H = o.src_H; q.src_W = o.src_W;
    q.use_sigma = o.use_sigma; q.sigma_const = o.sigma_const; q.clamp01 = o.clamp01;
    q.u8_bgr = o.u8_bgr; q.out_u8 = o.u8_io;
  }
  for (int l = 0; l < BSVD_NUM_LAYERS; ++l) {
    h->plan[l].p.seg_T = o.seg_T;
    if (launch_stage(h->plan[l], st)) return 1;
    if (evs) CUDA_TRY(cudaEventRecord((*evs)[l + 2], st));
    ++launches;
  }
  h->last_launches = launches;
  return 0;
}

int bsvd_forward_clip(bsvd_handle* h, const float* in, const float* noise_map, float* out, int T,
                      int in_c, int H, int W, void* stream) {
  if (!h || !in || !out) return fail("null argument");
  if (check_dev_ptr(h, in, "in") || check_dev_ptr(h, noise_map, "noise_map") || check_dev_ptr(h, out, "out"))
    return 1;
  return forward_clip_impl(h, in, noise_map, out, T, in_c, H, W, stream, CallOpts());
}

int bsvd_forward_clips(bsvd_handle* h, const float* in, const float* noise_map, float* out, int N, int T,
                       int in_c, int H, int W, void* stream) {
  if (!h || !in || !out) return fail("null argument");
  if (N < 1 || T < 1) return fail("N and T must be >= 1 (got N=%d T=%d)", N, T);
  if (check_dev_ptr(h, in, "in") || check_dev_ptr(h, noise_map, "noise_map") || check_dev_ptr(h, out, "out"))
    return 1;
  CallOpts o;
  o.seg_T = T;
  return forward_clip_impl(h, in, noise_map, out, N * T, in_c, H, W, stream, o);
}

static int denoise_clip_impl(bsvd_handle* h, const float* in, float sigma, float* out, int T, int H, int W,
                             void* stream, CallOpts o) {
  if (!h || !in || !out) return fail("null argument");
  if (H < 2 || W < 2) return fail("reflect padding needs H, W >= 2 (got %dx%d)", H, W);
  const bool blind = h->cfg.in_ch == 3;
  if (blind != (sigma < 0.f))
    return fail(blind ? "blind model: pass sigma < 0" : "non-blind model needs sigma >= 0");
  const int Hp = (H + 3) / 4 * 4, Wp = (W + 3) / 4 * 4;
  if (Hp - H >= H || Wp - W >= W) return fail("image too small to reflect-pad to a multiple of 4");
  o.src_H = H; o.src_W = W; o.use_sigma = blind ? 0 : 1; o.sigma_const = sigma; o.clamp01 = 1;
  // a 3-channel raw input: the 4th (noise-map) slot is synthesised by the first kernel
  return forward_clip_impl(h, in, nullptr, out, T, 3, Hp, Wp, stream, o);
}

int bsvd_denoise_clip(bsvd_handle* h, const float* in, float sigma, float* out, int T, int H, int W,
                      void* stream) {
  if (!h || !in || !out) return fail("null argument");
  if (check_dev_ptr(h, in, "in") || check_dev_ptr(h, out, "out")) return 1;
  return denoise_clip_impl(h, in, sigma, out, T, H, W, stream, CallOpts());
}

int bsvd_denoise_clip_u8(bsvd_handle* h, const uint8_t* in, float sigma, uint8_t* out, int T, int H,
                         int W, int bgr, void* stream) {
  if (!h || !in || !out) return fail("null argument");
  if (check_device(h) || check_dev_ptr(h, in, "in") || check_dev_ptr(h, out, "out")) return 1;
  const size_t need = (size_t)T * 3 * H * W * sizeof(float);
  if (h->d_norm_bytes < need) {
    if (h->d_norm) cudaFree(h->d_norm);
    h->d_norm = nullptr; h->d_norm_bytes = 0;
    CUDA_TRY(cudaMalloc((void**)&h->d_norm, need));
    h->d_norm_bytes = need;
  }
  CallOpts o;
  o.u8_io = 1; o.u8_bgr = bgr ? 1 : 0;
  return denoise_clip_impl(h, reinterpret_cast<const float*>(in), sigma, reinterpret_cast<float*>(out), T, H,
                           W, stream, o);
}

// ---- PSNR per frame (calculate_psnr_float, BasicSR/basicsr/metrics/psnr_ssim.py:130-168) -----------
constexpr int kPsnrBlocks = 64;
static __global__ void psnr_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int C,
                                           int H, int W, int cb, double* __restrict__ part) {
  const int t = blockIdx.y;
  const int hh = H - 2 * cb, ww = W - 2 * cb;
  const long long n = static_cast<long long>(C) * hh * ww;
  const long long base = static_cast<long long>(t) * C * H * W;
  double acc = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % ww);
    const long long r = i / ww;
    const int y = static_cast<int>(r % hh), c = static_cast<int>(r / hh);
    const long long o = base + (static_cast<long long>(c) * H + (y + cb)) * W + (x + cb);
    const float d = a[o] - b[o];
    acc += static_cast<double>(d) * d;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  __shared__ double ws[32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) part[t * gridDim.x + blockIdx.x] = acc;
  }
}
static __global__ void psnr_final_kernel(const double* __restrict__ part, int nb, double n, float* psnr) {
  const int t = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += 32) acc += part[t * nb + i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) {
    const double mse = acc / n;
    psnr[t] = mse == 0.0 ? __int_as_float(0x7f800000) : static_cast<float>(-10.0 * log10(mse));
  }
}
int bsvd_psnr(const float* a, const float* b, int T, int C, int H, int W, int crop_border, float* psnr,
              void* stream) {
  if (!a || !b || !psnr) return fail("null argument");
  if (T < 1 || C < 1 || crop_border < 0 || H - 2 * crop_border < 1 || W - 2 * crop_border < 1)
    return fail("bad PSNR shape T=%d C=%d %dx%d crop_border=%d", T, C, H, W, crop_border);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* part = nullptr;
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&part), sizeof(double) * T * kPsnrBlocks, st));
  psnr_partial_kernel<<<dim3(kPsnrBlocks, T), 256, 0, st>>>(a, b, C, H, W, crop_border, part);
  const double n = static_cast<double>(C) * (H - 2 * crop_border) * (W - 2 * crop_border);
  psnr_final_kernel<<<T, 32, 0, st>>>(part, kPsnrBlocks, n, psnr);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaFreeAsync(part, st));
  return 0;
}

// ---- SSIM per frame (calculate_ssim, BasicSR/basicsr/metrics/psnr_ssim.py:49-128) ------------------------
int bsvd_ssim(const float* a, const float* b, int T, int C, int H, int W, int crop_border, float data_range,
              float* ssim, void* stream) {
  if (!a || !b || !ssim) return fail("null argument");
  const int hh = H - 2 * crop_border, ww = W - 2 * crop_border;
  if (T < 1 || C < 1 || crop_border < 0 || hh < kSsimWin || ww < kSsimWin || !(data_range > 0.f))
    return fail("bad SSIM arguments T=%d C=%d %dx%d crop_border=%d data_range=%g (the cropped image must hold an "
                "11x11 window)", T, C, H, W, crop_border, (double)data_range);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int oh = hh - (kSsimWin - 1), ow = ww - (kSsimWin - 1);
  const int tiles_x = (ow + kSsimTile - 1) / kSsimTile, tiles_y = (oh + kSsimTile - 1) / kSsimTile;
  SsimWindow g;   // cv2.getGaussianKernel(11, 1.5): exp(-(i - 5)^2 / (2 sigma^2)), normalised
  double sum = 0.0;
  for (int i = 0; i < kSsimWin; ++i) { g.w[i] = exp(-((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); sum += g.w[i]; }
  for (int i = 0; i < kSsimWin; ++i) g.w[i] /= sum;
  double* part = nullptr;
  const size_t n_per_frame = (size_t)C * tiles_x * tiles_y;
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&part), sizeof(double) * T * n_per_frame, st));
  ssim_partial_kernel<<<dim3(tiles_x * tiles_y, C, T), kSsimTile * kSsimTile, 0, st>>>(
      a, b, C, H, W, crop_border, (double)data_range, g, tiles_x, tiles_y, part);
  ssim_final_kernel<<<T, 256, 0, st>>>(part, (int)n_per_frame, (double)C * oh * ow, ssim);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaFreeAsync(part, st));
  return 0;
}

static int ensure_dev(float** p, size_t* cur, size_t need) {
  if (*cur >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cur = 0;
  CUDA_TRY(cudaMalloc((void**)p, need));
  *cur = need;
  return 0;
}

int bsvd_forward_clip_host(bsvd_handle* h, const float* in_host, const float* nmap_host,
                           float* out_host, int T, int in_c, int H, int W, void* stream) {
  if (!h || !in_host || !out_host) return fail("null argument");
  if (check_device(h)) return 1;
  if (check_hw(T, in_c, H, W, nmap_host != nullptr, h->cfg.in_ch)) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t plane = (size_t)H * W * sizeof(float);
  if (ensure_dev(&h->d_in, &h->d_in_bytes, plane * T * in_c)) return 1;
  if (ensure_dev(&h->d_out, &h->d_out_bytes, plane * T * 3)) return 1;
  if (nmap_host && ensure_dev(&h->d_nmap, &h->d_nmap_bytes, plane * T)) return 1;
  CUDA_TRY(cudaMemcpyAsync(h->d_in, in_host, plane * T * in_c, cudaMemcpyHostToDevice, st));
  if (nmap_host)
    CUDA_TRY(cudaMemcpyAsync(h->d_nmap, nmap_host, plane * T, cudaMemcpyHostToDevice, st));
  if (bsvd_forward_clip(h, h->d_in, nmap_host ? h->d_nmap : nullptr, h->d_out, T, in_c, H, W,
                        stream))
    return 1;
  CUDA_TRY(cudaMemcpyAsync(out_host, h->d_out, plane * T * 3, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// Pipelined variant: returns after enqueueing.  The H2D copy of call i+1 (own copy stream) overlaps
// the forward of call i (caller's stream) and the D2H copy of call i-1 (second copy stream); device
// staging is double-buffered, ordering is by events.  bsvd_host_sync() waits for everything.
int bsvd_forward_clip_host_async(bsvd_handle* h, const float* in_host, const float* nmap_host,
                                 float* out_host, int T, int in_c, int H, int W, void* stream) {
  if (!h || !in_host || !out_host) return fail("null argument");
  if (check_device(h)) return 1;
  if (check_hw(T, in_c, H, W, nmap_host != nullptr, h->cfg.in_ch)) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!h->s_h2d) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
    }
  }
  const int k = (int)(h->host_calls & 1);
  const size_t plane = (size_t)H * W * sizeof(float);
  for (int b = 0; b < 2; ++b) {   // both staging sets up front: cudaMalloc synchronises the device
    if (ensure_dev(&h->pin[b], &h->pin_bytes[b], plane * T * in_c)) return 1;
    if (ensure_dev(&h->pout[b], &h->pout_bytes[b], plane * T * 3)) return 1;
    if (nmap_host && ensure_dev(&h->pnm[b], &h->pnm_bytes[b], plane * T)) return 1;
  }
  if (h->host_calls >= 2) {
    CUDA_TRY(cudaStreamWaitEvent(h->s_h2d, h->ev_comp[k], 0));   // forward i-2 has consumed pin[k]
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_d2h[k], 0));          // copy-out i-2 has drained pout[k]
  }
  CUDA_TRY(cudaMemcpyAsync(h->pin[k], in_host, plane * T * in_c, cudaMemcpyHostToDevice, h->s_h2d));
  if (nmap_host)
    CUDA_TRY(cudaMemcpyAsync(h->pnm[k], nmap_host, plane * T, cudaMemcpyHostToDevice, h->s_h2d));
  CUDA_TRY(cudaEventRecord(h->ev_h2d[k], h->s_h2d));
  CUDA_TRY(cudaStreamWaitEvent(st, h->ev_h2d[k], 0));
  if (bsvd_forward_clip(h, h->pin[k], nmap_host ? h->pnm[k] : nullptr, h->pout[k], T, in_c, H, W,
                        stream))
    return 1;
  CUDA_TRY(cudaEventRecord(h->ev_comp[k], st));
  CUDA_TRY(cudaStreamWaitEvent(h->s_d2h, h->ev_comp[k], 0));
  CUDA_TRY(cudaMemcpyAsync(out_host, h->pout[k], plane * T * 3, cudaMemcpyDeviceToHost, h->s_d2h));
  CUDA_TRY(cudaEventRecord(h->ev_d2h[k], h->s_d2h));
  ++h->host_calls;
  return 0;
}

int bsvd_host_last_output(bsvd_handle* h, float** dev_out) {
  if (!h || !dev_out) return fail("null argument");
  if (!h->host_calls) return fail("no bsvd_forward_clip_host_async call has been made on this handle");
  *dev_out = h->pout[(h->host_calls - 1) & 1];
  return 0;
}

int bsvd_host_sync(bsvd_handle* h) {
  if (!h) return fail("null handle");
  if (h->s_d2h) {
    CUDA_TRY(cudaStreamSynchronize(h->s_h2d));
    CUDA_TRY(cudaStreamSynchronize(h->s_d2h));
  }
  return 0;
}

static void drop_stream_graphs(bsvd_handle* h) {
  for (auto& g : h->stream.gexec)
    if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}
static void free_stream(bsvd_handle* h) {
  auto& S = h->stream;
  drop_stream_graphs(h);
  if (S.cap) cudaStreamDestroy(S.cap);
  if (S.ws) cudaFree(S.ws);
  S = bsvd_handle::Stream();
}

static size_t ring_slot_bytes(const bsvd_handle* h, int ring, int H, int W) {
  const int r = kRingRes[ring];
  const int C = (r == 1) ? h->cp[0] : (r == 2 ? h->cp[1] : h->cp[2]);
  return align_up((size_t)(H / r) * (W / r) * C * 2, 1024);
}

// Allocate the rings and pre-encode one tensor map per (layer, input ring slot).
static int build_stream(bsvd_handle* h, int H, int W) {
  auto& S = h->stream;
  free_stream(h);
  S.H = H; S.W = W;
  size_t total = 0;
  for (int b = 0; b < 2; ++b)
    for (int k = 0; k < kNumRings; ++k) total += ring_slot_bytes(h, k, H, W) * kRingSlots[k];
  const size_t raw_bytes = align_up((size_t)9 * 4 * H * W * sizeof(float), 1024);
  total += raw_bytes;
  const size_t aux_slot = align_up((size_t)H * W * 4 * (h->split ? 4 : 2), 1024);
  total += 9 * aux_slot;
  const size_t out_bytes = align_up((size_t)3 * H * W * sizeof(float), 1024);
  total += out_bytes;
  CUDA_TRY(cudaMalloc(&S.ws, total));
  S.ws_bytes = total;
  uint8_t* p = reinterpret_cast<uint8_t*>(S.ws);
  S.raw = reinterpret_cast<float*>(p); p += raw_bytes;
  S.aux = p; p += 9 * aux_slot;
  S.out_slot = reinterpret_cast<float*>(p); p += out_bytes;
  for (int b = 0; b < 2; ++b)
    for (int k = 0; k < kNumRings; ++k) { S.ring[b][k] = p; p += ring_slot_bytes(h, k, H, W) * kRingSlots[k]; }
  for (int b = 0; b < 2; ++b)
    for (int l = 0; l < 16; ++l) {
      auto& SL = S.layers[b * 16 + l];
      const StageDev& sd = h->stages[b * 16 + l];
      // block 1 reads the previous block's output ring as its input
      const int in_ring = (b == 1 && l == 0) ? kRingM : kLayerIn[l];
      const int in_blk = (b == 1 && l == 0) ? 0 : b;
      const int r = kRingRes[in_ring];
      const size_t in_bytes = ring_slot_bytes(h, in_ring, H, W);
      StageIO io;
      io.T = 1; io.H = H / r; io.W = W / r;
      if (sd.spec.pairx || sd.spec.pair_final) io.W /= 2;      // pixel pairs (same memory)
      io.overflow = h->d_overflow;
      io.in = S.ring[in_blk][in_ring];
      io.out = S.ring[b][kLayerOut[l]];          // patched per step
      io.ring_mode = sd.spec.shift ? 1 : 0;
      io.zero_future = sd.spec.shift ? 1 : 0;
      io.out_next = io.out;                      // placeholders so plan_stage's checks pass
      io.skip = S.ring[b][kRingX0]; io.skip_C = h->split ? h->cp[0] / 2 : h->cp[0];
      io.resid_in = S.raw; io.resid_C = 4;
      io.skip_T = kRingSlots[kRingX0]; io.skip_T_stride = (long long)ring_slot_bytes(h, kRingX0, H, W);
      if (l == 10) {
        io.skip = S.ring[b][kRingX1]; io.skip_C = h->split ? h->cp[1] / 2 : h->cp[1];
        io.skip_T = kRingSlots[kRingX1]; io.skip_T_stride = (long long)ring_slot_bytes(h, kRingX1, H, W);
      }
      io.out_T = kRingSlots[kLayerOut[l]]; io.out_T_stride = (long long)ring_slot_bytes(h, kLayerOut[l], H, W);
      if (l == 15 && b == 1) { io.skip = S.aux; io.skip_C = 4; }
      if (l == 15 && b == 0) io.aux_out = S.aux;
      if (plan_stage(sd, io, h->bf16, 0, &SL.tmpl)) return 1;
      const int nslots = kRingSlots[in_ring];
      SL.maps.resize(nslots);
      const int cin_map = sd.spec.first_im2col ? kChunk
                          : (sd.spec.pair_s2 ? 32 : (sd.spec.split ? 2 * sd.spec.cin : sd.spec.cin));   // channels per pixel as stored
      if (b == 0 && l == 0 && raw_tma_ok(S.raw, nullptr, W)) {
        SL.raw_maps.resize(9);
        for (int k = 0; k < 9; ++k)
          if (make_map_raw(&SL.raw_maps[k], S.raw + (size_t)k * 4 * H * W, h->cfg.in_ch, H, W, h->cfg.in_ch)) return 1;
      }
      const bool s2b = SL.tmpl.p.mode >= 4;
      if (s2b) SL.maps2.resize(nslots);
      for (int k = 0; k < nslots; ++k) {
        const void* base = S.ring[in_blk][in_ring] + (size_t)k * in_bytes;
        int rc = (sd.spec.stride != 2) ? make_map_halo(&SL.maps[k], base, 1, io.H, io.W, cin_map, sd.spec.rows)
                 : s2b ? make_map_s2(&SL.maps[k], base, 1, io.H, io.W, cin_map, sd.spec.rows + 1, kS2BoxPx)
                       : make_map_s2(&SL.maps[k], base, 1, io.H, io.W, cin_map, sd.spec.rows);
        if (!rc && s2b) rc = make_map_s2(&SL.maps2[k], base, 1, io.H, io.W, cin_map, sd.spec.rows, kS2BoxPx);
        if (rc) return rc;
      }
    }
  return 0;
}

// The stage launches of one push (step s, F frames known): every stage whose delayed frame exists runs
// on one frame; ring slots are selected by frame index.  `st` may be a capturing stream.
static int run_stream_layers(bsvd_handle* h, long long s, long long F, cudaStream_t st, float* out,
                             int* produced, int* launches_out) {
  auto& S = h->stream;
  const int H = S.H, W = S.W;
  const size_t plane = (size_t)H * W;
  int launches = 0;
  for (int b = 0; b < 2; ++b)
    for (int l = 0; l < 16; ++l) {
      const long long f = s - (kLayerDelay[l] + b * kBlockDelay);
      if (f < 0 || f >= F) continue;   // None propagation (bsvd_arch.py:135,139,219-224,...)
      auto& SL = S.layers[b * 16 + l];
      StageLaunch L = SL.tmpl;
      const int in_ring = (b == 1 && l == 0) ? kRingM : kLayerIn[l];
      L.map = SL.maps[f % kRingSlots[in_ring]];
      if (!SL.maps2.empty()) L.map_s = SL.maps2[f % kRingSlots[in_ring]];
      const int oring = kLayerOut[l];
      const size_t ob = ring_slot_bytes(h, oring, H, W);
      auto oslot = [&](long long ff) { return S.ring[b][oring] + (size_t)(ff % kRingSlots[oring]) * ob; };
      ConvParams& p = L.p;
      p.out = oslot(f);
      if (p.flags & EPI_SHIFT) {
        p.out_prev = (f > 0) ? oslot(f - 1) : nullptr;
        p.out_next = oslot(f + 1);
      }
      p.out_t0 = (int)(f % kRingSlots[oring]);     // TMA stores address the ring through map_o
      if (l == 10) {
        p.skip_t0 = (int)(f % kRingSlots[kRingX1]);   // skip add on the tensor core: ring slot of map_s
        p.skip = S.ring[b][kRingX1] + (size_t)p.skip_t0 * ring_slot_bytes(h, kRingX1, H, W);
      }
      if (l == 13) {
        p.skip_t0 = (int)(f % kRingSlots[kRingX0]);
        p.skip = S.ring[b][kRingX0] + (size_t)p.skip_t0 * ring_slot_bytes(h, kRingX0, H, W);
      }
      const size_t aux_slot = align_up((size_t)H * W * 4 * (h->split ? 4 : 2), 1024);
      if (l == 0 && b == 0) {
        L.first_in = S.raw + (size_t)(f % 9) * 4 * plane;   // raw ring slot holds all 4 channels
        L.first_nmap = nullptr; L.first_inc = h->cfg.in_ch;   // blind model: planes 0..2 only
        if (!SL.raw_maps.empty()) { L.first_raw_tma = 1; L.map_raw = SL.raw_maps[f % 9]; }
      }
      if (l == 15 && b == 0) {
        p.resid_in = S.raw + (size_t)(f % 9) * 4 * plane;
        p.aux_out = S.aux + (size_t)(f % 9) * aux_slot;
      }
      if (l == 15 && b == 1) {
        p.skip = S.aux + (size_t)(f % 9) * aux_slot;
        p.out = out;
        if (produced) *produced = 1;
      }
      if (launch_stage(L, st)) return 1;
      ++launches;
    }
  if (launches_out) *launches_out = launches;
  return 0;
}

int bsvd_stream_push(bsvd_handle* h, const float* frame, const float* noise_map, float* out,
                     int in_c, int H, int W, int* produced, void* stream) {
  if (produced) *produced = 0;
  if (!h || !out) return fail("null argument");
  if (check_device(h) || check_dev_ptr(h, frame, "frame") || check_dev_ptr(h, noise_map, "noise_map") ||
      check_dev_ptr(h, out, "out"))
    return 1;
  if (check_hw(1, frame ? in_c : h->cfg.in_ch, H, W, frame ? noise_map != nullptr : false, h->cfg.in_ch)) return 1;
  for (int l = 0; l < BSVD_NUM_LAYERS; ++l)
    if (!h->stages[l].loaded) return fail("weights of layer %d were never set", l);
  auto& S = h->stream;
  if (!S.ws || S.H != H || S.W != W) {
    if (S.ws && S.step != 0) return fail("frame size changed mid-stream (call bsvd_reset first)");
    if (build_stream(h, H, W)) return 1;
  }
  if (frame && S.ended)
    return fail("a frame was pushed after the end-of-stream marker (call bsvd_reset first)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long s = S.step;
  const size_t plane = (size_t)H * W;
  int launches = 0;
  if (frame) {
    // keep the raw frame for the temp1 residual 8 steps later (skip1, bsvd_arch.py:380,394)
    float* slot = S.raw + (size_t)(s % 9) * 4 * plane;
    CUDA_TRY(cudaMemcpyAsync(slot, frame, plane * in_c * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (noise_map)
      CUDA_TRY(cudaMemcpyAsync(slot + 3 * plane, noise_map, plane * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    ++S.n_in;
  } else {
    S.ended = true;
  }
  const long long F = S.n_in;          // frames known so far; final once ended
  // Steady state (every one of the 32 stages has a frame to work on): the launches of one push depend
  // only on the ring phase s mod 45, so they are captured once per phase into a CUDA graph and replayed —
  // one submission instead of 32 cudaLaunchKernelEx calls (the PDL edges between the stages are kept).
  // The graph leaves its frame in an internal slot; a D2D copy hands it to the caller's buffer.
  static const int graphs_on = [] { const char* e = getenv("BSVD_B200_NO_STREAM_GRAPH"); return (e && e[0] == '1') ? 0 : 1; }();
  const bool steady = frame && s > 2 * kBlockDelay;   // every stage has run once through the plain path
  if (steady && graphs_on) {
    if (S.gepoch != h->weights_epoch) { drop_stream_graphs(h); S.gepoch = h->weights_epoch; }
    const int ph = (int)(s % bsvd_handle::Stream::kPhases);
    if (!S.gexec[ph]) {
      if (!S.cap) CUDA_TRY(cudaStreamCreateWithFlags(&S.cap, cudaStreamNonBlocking));
      CUDA_TRY(cudaStreamBeginCapture(S.cap, cudaStreamCaptureModeThreadLocal));
      int n = 0, prod = 0;
      const int rc = run_stream_layers(h, s, F, S.cap, S.out_slot, &prod, &n);
      cudaGraph_t g = nullptr;
      const cudaError_t e = cudaStreamEndCapture(S.cap, &g);
      if (rc || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return rc ? 1 : fail("capturing the streaming step failed: %s", cudaGetErrorString(e));
      }
      const cudaError_t ei = cudaGraphInstantiate(&S.gexec[ph], g, 0);
      cudaGraphDestroy(g);
      if (ei != cudaSuccess) { S.gexec[ph] = nullptr; return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(ei)); }
    }
    CUDA_TRY(cudaGraphLaunch(S.gexec[ph], st));
    CUDA_TRY(cudaMemcpyAsync(out, S.out_slot, (size_t)3 * plane * sizeof(float), cudaMemcpyDeviceToDevice, st));
    ++S.graph_replays;
    launches = BSVD_NUM_LAYERS;
    if (produced) *produced = 1;
  } else {
    if (run_stream_layers(h, s, F, st, out, produced, &launches)) return 1;
  }
  ++S.step;
  h->last_launches = launches;
  return 0;
}

long long bsvd_stream_graph_replays(const bsvd_handle* h) { return h ? h->stream.graph_replays : 0; }

int bsvd_reset(bsvd_handle* h) {
  if (!h) return fail("null handle");
  h->stream.step = 0;
  h->stream.n_in = 0;
  h->stream.ended = false;
  return 0;
}

// ---- single-stage hook -------------------------------------------------------------------------
static thread_local float g_last_stage_ms = 0.f;
float bsvd_last_stage_ms(void) { return g_last_stage_ms; }
int bsvd_conv_stage(const bsvd_conv_desc* d, const void* in, const float* w, const float* bias,
                    const void* skip, void* out, void* stream) {
  if (!d || !in || !w || !out) return fail("null argument");
  if (d->cin % 64 || d->cin > 256) return fail("cin must be 64, 128 or 256");
  if (!(d->cout == 64 || d->cout == 128 || d->cout == 256 || d->cout == 512))
    return fail("cout must be 64, 128, 256 or 512");
  StageDev sd;
  StageSpec& s = sd.spec;
  s.cin = d->cin; s.cout = d->cout;
  s.stride = (d->flags & BSVD_EPI_STRIDE2) ? 2 : 1;
  s.relu6 = d->flags & BSVD_EPI_RELU6;
  s.pixshuf = d->flags & BSVD_EPI_PIXSHUF;
  s.skip = d->flags & BSVD_EPI_SKIP_ADD;
  s.shift = d->flags & BSVD_EPI_SHIFT_STORE;
  if (s.pixshuf && s.stride == 2) return fail("pixel shuffle and stride 2 cannot be combined");
  s.derive();
  if (d->debug_variant & (32 | (1 << 20))) {   // debug: single-CTA kernels / plain (unstacked) 64->64
    s.stacked = false; s.tap_begin = 0; s.tap_end = 9;
  }
  const int bf16 = d->precision == BSVD_PREC_BF16;
  if (upload_stage(sd, w, bias, bf16)) return 1;
  StageIO io;
  io.in = in; io.T = d->T; io.H = d->H; io.W = d->W; io.out = out;
  const int Ho = d->H / s.stride * (s.pixshuf ? 2 : 1), Wo = d->W / s.stride * (s.pixshuf ? 2 : 1);
  const int Co = s.pixshuf ? s.cout / 4 : s.cout;
  io.skip = skip; io.skip_C = Co; io.skip_frame_stride = (long long)Ho * Wo * Co;
  StageLaunch L;
  int rc = plan_stage(sd, io, bf16, (d->debug_variant & 0xff) | ((d->debug_variant & (1 << 21)) ? 256 : 0), &L);
  const int a_over = (d->debug_variant >> 8) & 0xf;
  if (!rc && a_over) {   // debug: override the number of A stages (smem permitting)
    L.smem += (size_t)(a_over - L.p.a_stages) * L.p.a_stage_bytes;
    L.p.a_stages = a_over;
    if (L.smem > kSmemOptIn) rc = fail("debug a_stages override exceeds shared memory");
  }
  const int w_over = (d->debug_variant >> 16) & 0xf;
  if (!rc && w_over && !L.p.w_resident) {   // debug: override the number of W stages
    L.smem += (size_t)(w_over - L.p.w_stages) * L.p.w_stage_bytes;
    L.p.w_stages = w_over;
    if (L.smem > kSmemOptIn) rc = fail("debug w_stages override exceeds shared memory");
  }
  cudaStream_t cst = reinterpret_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = ((d->debug_variant >> 12) & 0xf) + 1;
  if (!rc) rc = launch_stage(L, cst);                 // warm-up / the result
  cudaEventRecord(e0, cst);
  for (int i = 0; i < reps && !rc; ++i) rc = launch_stage(L, cst);
  cudaEventRecord(e1, cst);
  cudaError_t e = cudaStreamSynchronize(cst);
  float ms = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  g_last_stage_ms = ms / reps;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  free_stage(sd);
  if (!rc && e != cudaSuccess) return fail("conv stage failed: %s", cudaGetErrorString(e));
  return rc;
}

}  // extern "C"
