// stage_launch.cuh — launch record of one fused conv stage and the dispatch from a plan's run-time
// description (tile shape, operand type, epilogue feature set, pipeline shape) to the matching
// conv3x3_tc_kernel instance.  Shared by bsvd_capi.cu (host schedules) and conv_inst.cu (the
// translation units that instantiate the kernels, one per (NTILE, R) pair so they build in parallel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstring>

#include "common.cuh"
#include "conv_tc.cuh"

namespace bsvd {


// dynamic shared memory: 227 KB opt-in limit minus the kernel's static shared (barriers, bias)
constexpr size_t kSmemOptIn = 232448 - 3072;

struct StageLaunch;
template <int NTILE, int R, bool BF16, bool CTA2, int MASK, int EW>
static int launch_inst(const StageLaunch& L, cudaStream_t st);
template <int NTILE, int R, bool BF16, bool CTA2, int MASK, int EW, int PIPE>
static int launch_pipe(const StageLaunch& L, cudaStream_t st);
struct StageLaunch {
  CUtensorMap map;      // activations
  CUtensorMap map_w;    // packed weights (CTA-pair kernels)
  CUtensorMap map_s;    // skip tensor, PixelShuffle view (skip add on the tensor core)
  CUtensorMap map_o;    // output tensor (TMA stores)
  ConvParams p;
  int grid = 0;
  size_t smem = 0;
  int ntile = 0, rows = 0;
  int cta2 = 0;         // 1 = cta_group::2 kernel, launched as clusters of 2 CTAs
  int split = 0;        // 1 = fp32-grade mode (EPI_SPLIT instances, fp16 pieces)
  const void* in_ptr = nullptr;   // fp32-grade edge convs (split_edge.cuh): the stage's input tensor
  int ew = 8;           // epilogue warps (8 or 16)
  // fused first stage (first_conv.cuh): raw network input, patched per call
  const float* first_in = nullptr;
  const float* first_nmap = nullptr;
  int first_inc = 4;
  int first_u8 = 0;               // raw input is uint8 HWC (bsvd_denoise_clip_u8)
  float* first_norm = nullptr;    // fp32 [T][3][H][W] copy of the normalised frames (temp1 residual)
  // raw fp32 planes through TMA (first_conv.cuh RAW instances): [planes][H][W] views of the input / noise map
  int first_raw_tma = 0;
  CUtensorMap map_raw, map_rawnm;
};

// pick the compile-time pipeline shape (conv_tc.cuh PIPE) the plan asks for
template <int NTILE, int R, bool BF16, bool CTA2, int MASK, int EW>
static int launch_inst(const StageLaunch& L, cudaStream_t st) {
  if constexpr (CTA2) {
    const int mode = L.p.mode, res = L.p.w_resident;
    if (mode == 4) {
      // stride 2 with sub-plane boxes: only this pipeline implements it
      if constexpr ((MASK & (EPI_PIXSHUF | EPI_RESID_IN | EPI_TMA_OUT)) == 0)
        return launch_pipe<NTILE, R, BF16, CTA2, MASK, EW, 4>(L, st);
      return fail("no stride-2 kernel instance for this stage (NTILE=%d R=%d)", NTILE, R);
    }
    if (mode == 5) {
      // the same on pixel pairs (c32 configurations: 32 -> 64 channels, full -> half resolution)
      if constexpr (NTILE == 64 && R == 2 && (MASK & (EPI_PIXSHUF | EPI_RESID_IN | EPI_TMA_OUT)) == 0)
        return launch_pipe<NTILE, R, BF16, CTA2, MASK, EW, 5>(L, st);
      return fail("no pair stride-2 kernel instance for this stage (NTILE=%d R=%d)", NTILE, R);
    }
    if (L.p.desc_variant != 0 || L.p.tap_begin != 0 || L.p.tap_end != (mode == 2 ? 3 : 9)) {
      // debug switches / partial tap ranges only exist in the generic pipeline
    } else if constexpr (NTILE == 64 && R == 2 && (MASK & EPI_SPLIT) == 0) {
      if (mode == 2 && res) return launch_pipe<NTILE, R, BF16, CTA2, MASK, EW, 2>(L, st);
    } else {
      if (mode == 0 && !res) return launch_pipe<NTILE, R, BF16, CTA2, MASK, EW, 0>(L, st);
    }
  }
  return launch_pipe<NTILE, R, BF16, CTA2, MASK, EW, 3>(L, st);
}
template <int NTILE, int R, bool BF16, bool CTA2, int MASK, int EW, int PIPE>
static int launch_pipe(const StageLaunch& L, cudaStream_t st) {
  static std::atomic<bool> attr_done[64];      // per device: the attribute is per (function, device)
  auto kern = conv3x3_tc_kernel<NTILE, R, BF16, CTA2, MASK, EW, PIPE>;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    attr_done[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(64 + 32 * EW);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL, see conv_tc.cuh
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (CTA2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, L.map, L.map_w, L.map_s, L.map_o, L.p));
  return 0;
}
// Epilogue feature sets the kernels are specialised for; a stage runs on the smallest set that
// covers its flags (the single-CTA debug path only has the general instance).
constexpr int kMaskAct = EPI_RELU6 | EPI_RELU;       // the activation is a run-time flag inside every instance
constexpr int kMaskPlain = kMaskAct;
constexpr int kMaskShift = kMaskAct | EPI_SHIFT;
constexpr int kMaskResid = kMaskAct | EPI_RESID_IN;
// PixelShuffle stages: the skip add runs on the tensor core (ConvParams::skip_mma), so the
// epilogue instances carry no skip path
constexpr int kMaskUpTma = EPI_PIXSHUF | EPI_TMA_OUT;   // upc1.convblock.0: units leave through TMA stores
constexpr int kMaskUpShift = EPI_PIXSHUF | EPI_SHIFT;   // upc2.convblock.0 with the skip on the tensor core (opt-in)
constexpr int kMaskUpSkipShift = EPI_PIXSHUF | EPI_SKIP | EPI_SHIFT;   // upc2.convblock.0: skip added in the epilogue
constexpr int kMaskUpSkip = EPI_PIXSHUF | EPI_SKIP;                    // c32 upc1.convblock.0 (N tile 128)
constexpr int kMaskAll = kMaskAct | EPI_SHIFT | EPI_PIXSHUF | EPI_SKIP | EPI_RESID_IN;
// ReLU6-only sets (the BSVD-64 yml: act 'relu6'): the clamp is unconditional, no fp16 range guard, no plain-ReLU
// path (see epilogue_unit, kOnly6); and temp1's last conv, which has no activation at all
constexpr int kMaskPlain6 = EPI_RELU6;
constexpr int kMaskShift6 = EPI_RELU6 | EPI_SHIFT;
constexpr int kMaskResidOnly = EPI_RESID_IN;
template <int NTILE, int R, bool BF16>
static int launch_dtype(const StageLaunch& L, cudaStream_t st) {
  if (L.split) {
    // fp32-grade mode: every stored value leaves as a (hi, lo) fp16 pair (EPI_SPLIT); the skip add always
    // runs on the tensor core, so the instance set is small
    if constexpr (!BF16) {
      const int f = L.p.flags & kMaskAll;
      if (!L.cta2) return fail("the fp32-grade mode needs the CTA-pair kernels");
      if (L.p.tma_out) {
        if ((f & ~kMaskPlain) == 0) return launch_inst<NTILE, R, false, true, kMaskPlain | EPI_TMA_OUT | EPI_SPLIT, 8>(L, st);
        if ((f & ~kMaskResid) == 0) return launch_inst<NTILE, R, false, true, kMaskResid | EPI_TMA_OUT | EPI_SPLIT, 8>(L, st);
        if constexpr (NTILE == 256 && R == 1)
          if ((f & ~kMaskUpTma) == 0) return launch_inst<NTILE, R, false, true, kMaskUpTma | EPI_SPLIT, 8>(L, st);
      }
      if constexpr (NTILE != 64)
        if ((f & ~kMaskShift) == 0) return launch_inst<NTILE, R, false, true, kMaskShift | EPI_SPLIT, 8>(L, st);
      if constexpr (NTILE == 256 && R == 1)
        if ((f & ~kMaskUpShift) == 0) return launch_inst<NTILE, R, false, true, kMaskUpShift | EPI_SPLIT, 8>(L, st);
    }
    return fail("no fp32-grade kernel instance for this stage (NTILE=%d R=%d flags=%d)", NTILE, R, L.p.flags);
  }
  if (!L.cta2) return launch_inst<NTILE, R, BF16, false, kMaskAll, 8>(L, st);
  const int f = L.p.flags & kMaskAll;
  {
    if constexpr (!(NTILE == 256 && R == 2)) {
      const bool only6 = (f & EPI_RELU6) && !(f & EPI_RELU);
      if (L.p.tma_out) {
        if (only6 && (f & ~kMaskPlain6) == 0) return launch_inst<NTILE, R, BF16, true, kMaskPlain6 | EPI_TMA_OUT, 8>(L, st);
        if constexpr (NTILE == 64)
          if (f == kMaskResidOnly) return launch_inst<NTILE, R, BF16, true, kMaskResidOnly | EPI_TMA_OUT, 8>(L, st);
      } else if (only6 && (f & ~kMaskShift6) == 0) {
        return launch_inst<NTILE, R, BF16, true, kMaskShift6, 8>(L, st);
      }
    }
    if (L.p.tma_out) {     // no temporal shift on the output: units leave through TMA stores
      if ((f & ~kMaskPlain) == 0) return launch_inst<NTILE, R, BF16, true, kMaskPlain | EPI_TMA_OUT, 8>(L, st);
      if ((f & ~kMaskResid) == 0) return launch_inst<NTILE, R, BF16, true, kMaskResid | EPI_TMA_OUT, 8>(L, st);
      if constexpr (NTILE != 64)
        if ((f & ~kMaskShift) == 0) return launch_inst<NTILE, R, BF16, true, kMaskShift | EPI_TMA_OUT, 8>(L, st);
    }
    if ((f & ~kMaskPlain) == 0) return launch_inst<NTILE, R, BF16, true, kMaskPlain, 8>(L, st);
    if ((f & ~kMaskShift) == 0) return launch_inst<NTILE, R, BF16, true, kMaskShift, 8>(L, st);
    if ((f & ~kMaskResid) == 0) return launch_inst<NTILE, R, BF16, true, kMaskResid, 8>(L, st);
    if constexpr (NTILE == 128 && R == 2) {
      // c32 upc1.convblock.0 (64 -> 128, PixelShuffle to 32 channels + skip added in the epilogue)
      if ((f & ~kMaskUpSkip) == 0) return launch_inst<NTILE, R, BF16, true, kMaskUpSkip, 8>(L, st);
    }
    if constexpr (NTILE == 256 && R == 1) {
      if (L.p.tma_out && (f & ~kMaskUpTma) == 0) return launch_inst<NTILE, R, BF16, true, kMaskUpTma, 8>(L, st);
      if ((f & ~kMaskUpShift) == 0) return launch_inst<NTILE, R, BF16, true, kMaskUpShift, 8>(L, st);
      if ((f & ~kMaskUpSkipShift) == 0) return launch_inst<NTILE, R, BF16, true, kMaskUpSkipShift, 8>(L, st);
    }
  }
  return launch_inst<NTILE, R, BF16, true, kMaskAll, 8>(L, st);
}
// external linkage: explicitly instantiated in conv_inst.cu, declared `extern template` by callers
template <int NTILE, int R>
int launch_one(const StageLaunch& L, cudaStream_t st) {
  return (L.p.flags & EPI_BF16) ? launch_dtype<NTILE, R, true>(L, st)
                                : launch_dtype<NTILE, R, false>(L, st);
}


}  // namespace bsvd
