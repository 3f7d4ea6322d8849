// conv_tc.cuh — fused 3x3 convolution stage on sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// One kernel family covers every conv of the BSVD DenBlocks (reference: nn.Conv2d call sites
// bsvd_arch.py:31-38, 208-213, 238-239, 265, 295-298) together with what the reference runs as
// separate elementwise kernels around it: bias, ReLU6 / ReLU (:185-192), PixelShuffle (:266), the
// skip add (:402-406), the residual (:408-414) and the bidirectional-buffer channel shift (:42-50).
//
// Formulation: implicit GEMM, im2col-free.
//   M = 128 consecutive output pixels of one image row per CTA; CTA pairs (cta_group::2) issue one
//       M=256 tcgen05.mma over two pixel tiles and share each filter slab (half per CTA)
//   N = NTILE output channels
//   K = 9 taps x Cin, walked as (64-channel chunk) x (tap) x (4 k-steps of 16)
// Activations live in HBM as NHWC 16-bit, so a pixel's 64-channel chunk is one 128-byte row of
// the SWIZZLE_128B K-major canonical layout.  Stride-1 convs stage ONE haloed tile
// [(R+2) rows][130 px][64 ch] per chunk with a single TMA box (out-of-bounds = conv zero padding)
// and form the nine shifted A operands purely by offsetting the UMMA descriptor start address by
// (dy*130+dx)*128 bytes — no im2col, no re-load.  Stride-2 convs load one box per tap from a 5-D
// space-to-depth view of the same NHWC tensor.  Weights are pre-swizzled on the host and streamed
// through a ring of filter slabs (or kept resident when the whole bank fits: the 64->64 stages,
// which also stack the vertical taps in N, see the PIPE == 2 branch of the MMA warp).
//
// Warp roles (320 threads, 1 CTA/SM, persistent over tiles):
//   warp 0      TMA producer          warp 1      MMA issuer + TMEM allocator (leader CTA issues)
//   warps 2..9  epilogue: TMEM -> registers -> bias/act/skip/residual in fp32 -> 16-bit ->
//               per-warp staging tile in SWIZZLE_64B layout -> TMA store of the unit, or (stages
//               whose output carries the temporal shift) transposed read-back and 128-bit global
//               stores routed to frame t-1 / t / t+1
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of i+1.
// PixelShuffle + skip stages accumulate the skip tensor on the tensor core (identity MMA on
// TMA-loaded skip blocks, ConvParams::skip_mma) instead of adding it in the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <type_traits>

// the PIPE == 4 branches leave their tile loops with `continue`: the generic loops behind them are
// unreachable in those instances by design
#pragma nv_diag_suppress 128

namespace bsvd {

constexpr int kRunPx = 128;        // pixels per MMA row-run (UMMA M)
constexpr int kChunk = 64;         // channels per K chunk (128 B of 16-bit)
constexpr int kHaloPx = kRunPx + 2;
constexpr int kRowBytes = kHaloPx * 128;   // one haloed image row of one chunk: 16640 B
constexpr int kMaxStages = 12;
constexpr int kThreads = 320;        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kEpiThreads = 256;
constexpr int kStageBytesPerWarp = 2048;   // epilogue staging: [32 px][32 ch] 16-bit per warp
constexpr int kMaxBias = 512;
// Stride-2 stages (PIPE 4): the 3x3 taps of a stride-2 conv touch only four sub-planes (row parity py,
// column parity px) of the input.  One TMA box per sub-plane from the space-to-depth view
// [T][H/2][2][W/2][2C], consumed by the taps that read it in the order below (tap = dy*3+dx):
//   box 0 (py1,px1): taps 0,2,6,8   box 1 (py1,px0): taps 1,7   box 2 (py0,px1): taps 3,5   box 3 (py0,px0): tap 4
// A tap's operand is its box offset by (row, col) 128-byte pixels, exactly like the halo tile.
constexpr int kS2BoxPx = kRunPx + 1;                       // 129 pixels per box row
// PIPE 5 is the same pipeline for a 32-channel input read as pixel PAIRS ([H][W/2][64], see the c32
// configurations in bsvd_capi.cu): the stride-2 conv then strides only in y, output pixel x reads pairs x-1
// and x, so there are two boxes (odd rows: R+1 rows, even rows: R rows, both starting at pair x0-1) and six
// (dy, pair) taps whose filter slabs the host packs in consumption order:
//   box 0 (py1): (dy0,x-1) (dy0,x) (dy2,x-1) (dy2,x)     box 1 (py0): (dy1,x-1) (dy1,x)
template <int PIPE> __host__ __device__ constexpr int s2_ntaps() { return PIPE == 5 ? 6 : 9; }
template <int PIPE> __host__ __device__ constexpr int s2_slab(int j) {     // filter slab of the j-th tap consumed
  if (PIPE == 5) return j;
  return j == 0 ? 0 : j == 1 ? 2 : j == 2 ? 6 : j == 3 ? 8 : j == 4 ? 1 : j == 5 ? 7 : j == 6 ? 3 : j == 7 ? 5 : 4;
}
template <int PIPE> __host__ __device__ constexpr int s2_box(int j) {
  if (PIPE == 5) return j < 4 ? 0 : 1;
  return j < 4 ? 0 : j < 6 ? 1 : j < 8 ? 2 : 3;
}
template <int PIPE> __host__ __device__ constexpr bool s2_box_first(int j) { return j == 0 || s2_box<PIPE>(j) != s2_box<PIPE>(j - 1); }
template <int PIPE> __host__ __device__ constexpr bool s2_box_last(int j) {
  return j == s2_ntaps<PIPE>() - 1 || s2_box<PIPE>(j) != s2_box<PIPE>(j + 1);
}
template <int PIPE> __host__ __device__ constexpr int s2_box_py(int b) { return PIPE == 5 ? (b == 0) : (b < 2); }
template <int PIPE> __host__ __device__ constexpr int s2_box_px(int b) { return PIPE == 5 ? 0 : (b == 0 || b == 2); }
template <int PIPE> __host__ __device__ constexpr int s2_row_off(int j) {   // box row of output row 0's operand
  if (PIPE == 5) return (j == 2 || j == 3) ? 1 : 0;
  return (s2_slab<4>(j) / 3 == 2) ? 1 : 0;
}
template <int PIPE> __host__ __device__ constexpr int s2_col_off(int j) {   // box column of output pixel 0's operand
  if (PIPE == 5) return j & 1;
  return (s2_slab<4>(j) % 3 == 2) ? 1 : 0;
}

// masks of epilogue features a kernel instance is compiled with
enum : int {
  EPI_RELU6 = 1,
  EPI_PIXSHUF = 2,
  EPI_SKIP = 4,
  EPI_SHIFT = 8,
  EPI_RESID_IN = 16,    // temp1 outc.3: out[:, :3] = raw_in[:, :3] - out[:, :3]
  EPI_FINAL = 32,       // temp2 outc.3 (final_conv.cuh): fp32 NCHW out = skip[:, :3] - conv[:, :3]
  EPI_BF16 = 64,        // 16-bit storage type is bf16 (else fp16)
  EPI_ZERO_FUTURE = 128, // streaming: always zero own [0:fold) (overwritten when t+1 arrives)
  EPI_TMA_OUT = 256,    // compile-time only: units leave through cp.async.bulk.tensor stores
  EPI_RELU = 512,       // nn.ReLU (act='relu', the c32 configurations) instead of nn.ReLU6
  EPI_SPLIT = 1024      // compile-time only: fp32-grade mode, every stored value leaves as a (hi, lo) fp16 pair
};

struct ConvParams {
  // ---- problem geometry (OUTPUT grid of the conv, before PixelShuffle) ----
  int T, H, W;
  int cin_chunks;       // Cin / 64
  int n_tiles;          // GEMM N / NTILE
  int tap_begin, tap_end;
  int xblocks, yblocks; // ceil(W/128), ceil(H/R)
  int total_tiles;      // 1-CTA: positions*n_tiles; CTA pair: ceil(positions/2)*n_tiles
  int positions;        // T*yblocks*xblocks pixel tiles
  // decode_tile runs once per tile in every warp of every role: its three divisions by run-time values were
  // ~80 of the ~330 instructions an epilogue warp issues per unit (ncu, first conv).  The host supplies
  // ceil(2^32 / d) for d = n_tiles, xblocks, yblocks (0 for d == 1): umulhi(n, magic) == n / d is exact while
  // n * d < 2^32, which plan_stage checks for each of the three dividends (it fails otherwise: such a launch
  // would need a workspace far beyond the 180 GB of the device).  The instances with the skip add in their
  // epilogue (EPI_SKIP) keep the plain divisions: they sit on the 168-register ceiling of a 320-thread CTA, and
  // with the shorter coordinate code ptxas schedules their skip loads earlier and spills (measured: the c32
  // upc1.convblock.0 stage 0.22 -> 0.45 ms); their epilogues are not instruction-bound anyway.
  uint32_t div_nt, div_xb, div_yb;
  int mode;             // 0 = halo (stride 1), 1 = per-tap boxes (stride 2, generic pipeline only),
                        // 2 = halo with the vertical taps stacked in N (64->64 stages, see below),
                        // 4 = stride 2 with one box per input sub-plane (see s2_slab)
  int w_rows_cta;       // CTA pair: filter rows one CTA stages per W stage
  int cin_total;        // Cin (stride-2 coordinate math)
  // ---- pipeline ----
  int a_stages, w_stages;
  uint32_t a_stage_bytes, w_stage_bytes;   // stage strides in smem (multiples of 1024)
  uint32_t a_tx_bytes;                     // bytes one A-stage TMA box delivers
  int w_resident;
  int desc_variant;     // 0 production; debug bits: 2 skip MMA issue, 4 skip epilogue stores,
                        // 8 tap offsets forced to 0 (aligned A), 16 no TMA loads at all,
                        // 64 no skip loads, 128 no global stores
  // ---- operands ----
  const void* wpack;    // [n_tile][chunk][tap][NTILE][64] 16-bit, rows pre-swizzled (SW128)
  const float* bias;    // [GEMM N]
  // the same bias values inside the kernel-parameter (constant) bank: the epilogue reads them with
  // constant loads instead of shared-memory loads, which compete with tcgen05.mma operand reads for
  // the shared-memory data pipe (ncu: 8 of the 16 LDS/STS per epilogue unit were bias reads)
  float bias_c[kMaxBias];
  // ---- epilogue ----
  int flags;
  void* out;            // 16-bit NHWC, or fp32 NCHW for EPI_FINAL
  void* out_prev;       // streaming: frame t-1 / t+1 slots (clip mode: unused)
  void* out_next;
  int ring_mode;        // 1 = use out_prev/out_next pointers (T must be 1)
  int out_C, out_H, out_W;
  int out_C_log2;       // out_C is a power of two: PixelShuffle column -> (sub-pixel, channel) by shift/mask
  long long out_frame_stride;   // elements
  const void* skip;     // 16-bit NHWC with the output's shape (EPI_SKIP) / temp1 output (EPI_FINAL)
  long long skip_frame_stride;
  int skip_C;
  const float* resid_in;   // fp32 NCHW raw network input (EPI_RESID_IN)
  int resid_C;
  void* aux_out;        // EPI_RESID_IN: compact 16-bit [T][H][W][4] copy of output channels 0..3
                        // (the skip1 operand of temp2's residual: 8 B/px instead of a 128 B row)
  int fold;             // shift fold size in output channels (EPI_SHIFT)
  // ---- skip add on the tensor core / TMA stores (CTA-pair kernels, see conv3x3_tc_kernel) ----
  int skip_mma;         // > 0: number of 64-column blocks per tile whose skip operand is accumulated
                        // by an identity MMA (PixelShuffle + skip stages); the epilogue sees no skip
  int tma_out;          // 1: the epilogue stores its units with TMA (EPI_TMA_OUT instances)
  int skip_t0, out_t0;  // frame coordinate offsets into map_s / map_o (streaming: ring slot)
  uint32_t stg_bytes_per_warp;   // epilogue staging per warp: 2 KB, or 4 KB (double-buffered) with TMA stores
  // ---- callers either side of the path folded into the first / last kernels (bsvd_denoise_clip) ----
  // The network runs on H x W (multiples of 4); the raw input / final output are src_H x src_W images
  // (0 = same): reads beyond them are reflected (DenoisingModel.padding_input, F.pad 'reflect'),
  // the final store crops back and optionally clamps to [0,1] (temp_denoise).
  int src_H, src_W;
  int use_sigma;        // 1: the 4th input channel is the constant sigma_const (no noise-map tensor)
  float sigma_const;
  int clamp01;
  int u8_bgr;           // uint8 HWC frame I/O (bsvd_denoise_clip_u8): 1 = channel order B,G,R
  int out_u8;           // last kernel: store round(clamp(x) * 255) as uint8 [T][H][W][3] (tensor2img)
  // fp16 storage has no headroom beyond 65504: stages whose output is not clamped by ReLU6 (the
  // PixelShuffle + skip convs, temp1's output, every stage of an act='relu' model) set this sticky flag
  // when a value they store is inf/NaN, so an overflow can never pass silently (bsvd_overflow_flag)
  unsigned* overflow;
  // c32 configurations: a 32-channel full-resolution tensor is processed as pixel PAIRS ([H][W/2][64]); the
  // stage's W then counts pairs and GEMM column n = a*32 + co is channel co of pixel 2*x + a.  Only the
  // residual / compact-copy code of temp1's last conv needs to know (EPI_RESID_IN).
  int pair_px;
  int seg_T;            // clip mode: frames per independent clip when the T frames are several clips (0 = one clip)
  // ---- fp32-grade mode (BSVD_PREC_FP32X3): x = hi + lo, W = hi + lo in fp16, y = W_hi x_hi + W_hi x_lo + W_lo x_hi.
  // A tensor with C logical channels is stored with 2C channels per pixel, [hi(C) | lo(C)]; the K loop walks
  // 3C "virtual" input channels [x_hi | x_lo | x_hi] against the packed weights [W_hi | W_hi | W_lo]:
  // virtual chunk c reads physical chunk c % phys_chunks.  The epilogue stores hi at channel n, lo at n + out_C.
  int phys_chunks;      // physical 64-channel chunks per pixel of the input tensor (0 = cin_chunks: no wrap)
  int out_pitch_C;      // channels per pixel of the OUTPUT tensor as stored (out_C, or 2 * out_C in split mode)
  int skip_blocks;      // skip_mma: blocks per accumulator set (= skip_mma without split; skip_mma = 2x with it)
};
// reflected source coordinate of padded coordinate v (v < n_pad), source extent n (bottom/right pad)
__device__ __forceinline__ int reflect_src(int v, int n) { return v < n ? v : 2 * n - 2 - v; }

// --------------------------------------------------------------------------------------------
// PTX helpers
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  // relaxed (see mbar_arrive_cluster): no memory fence wanted on the accumulator hand-back
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((++spins & 0x3ff) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();   // 4 s
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes,
                                          uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// TMA store of one epilogue unit: shared memory [32 px][32 ch] (SWIZZLE_64B) -> global tensor.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_group() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: both CTAs issue their own TMA loads but signal the LEADER
// CTA's mbarrier (peer bit 24 of the shared::cluster address cleared); the leader issues one MMA
// that drives both SMs' tensor cores and multicasts the commit to both CTAs' barriers.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
// Arrive on the barrier at the same offset in CTA `rank` of the cluster.  RELAXED on purpose: the
// only thing this arrival publishes is "my tcgen05.ld of the accumulator has completed", which
// tcgen05.fence::before_thread_sync orders.  A .release arrive makes ptxas emit MEMBAR.ALL.GPU +
// ERRBAR, i.e. the warp would first wait for every global store of its previous units to drain
// (measured: 20 % of all stall samples, and the whole epilogue store/skip traffic serialised
// behind the MMAs).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
// Programmatic dependent launch: let the next stage's CTAs start their prologue (barrier init, TMEM
// allocation, bias staging) on SMs this grid has already left; `pdl_wait` then blocks until the
// previous stage has completed and its stores are visible.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(0) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr, int variant) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);          // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                          // [16,30) LBO (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // [32,46) SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;                          // [46,48) descriptor version (sm_100)
  if (variant == 1) d |= static_cast<uint64_t>((saddr >> 7) & 7u) << 49;   // [49,52) base offset
  d |= static_cast<uint64_t>(2) << 61;                          // [61,64) SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=fp16|bf16, both K-major, M=128, N=n.
__host__ __device__ constexpr uint32_t make_idesc(int n, int bf16, int m = kRunPx) {
  return (1u << 4) | (static_cast<uint32_t>(bf16) << 7) | (static_cast<uint32_t>(bf16) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ float relu6f(float x) { return fminf(fmaxf(x, 0.f), 6.f); }

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if constexpr (BF16) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
}
template <bool BF16>
__device__ __forceinline__ float load16(const void* p, long long idx) {
  if constexpr (BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
  else return __half2float(reinterpret_cast<const __half*>(p)[idx]);
}

// --------------------------------------------------------------------------------------------
// tile bookkeeping
// --------------------------------------------------------------------------------------------
struct TileCoord {
  int nt, t, y0, x0;
};
// CTA2: `tile` indexes (pair of neighbouring pixel tiles, n tile); CTA `rank` takes position
// 2*pair+rank.  A position past the end (odd count) yields t == T: every TMA box is then fully
// out of bounds (zero fill) and the epilogue stores nothing.
// n / d for a divisor whose reciprocal the host prepared (see ConvParams::div_nt)
__device__ __forceinline__ int tile_div(int n, uint32_t magic) {
  return magic ? static_cast<int>(__umulhi(static_cast<uint32_t>(n), magic)) : n;
}
template <int R, bool FAST = true>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile, int cta2 = 0,
                                                 int rank = 0) {
  TileCoord c;
  if constexpr (FAST) {
    int s = tile_div(tile, p.div_nt);
    c.nt = tile - s * p.n_tiles;
    if (cta2) s = 2 * s + rank;
    const int q = tile_div(s, p.div_xb);
    const int xb = s - q * p.xblocks;
    c.t = tile_div(q, p.div_yb);
    const int yb = q - c.t * p.yblocks;
    c.y0 = yb * R;
    c.x0 = xb * kRunPx;
  } else {
    c.nt = tile % p.n_tiles;
    int s = tile / p.n_tiles;
    if (cta2) s = 2 * s + rank;
    int xb = s % p.xblocks;
    s /= p.xblocks;
    int yb = s % p.yblocks;
    c.t = s / p.yblocks;
    c.y0 = yb * R;
    c.x0 = xb * kRunPx;
  }
  return c;
}

// --------------------------------------------------------------------------------------------
// Epilogue of one unit = 32 GEMM columns x 32 pixels of one warp.
//   phase 1 (lane = pixel): fp32 bias / skip add / ReLU6 / residual, round to 16 bit, park the
//            pixel's 64 bytes in the warp's staging tile (XOR-swizzled: conflict-free both ways)
//   phase 2 (lane = 16-byte chunk): read back transposed so that 4 lanes cover one pixel's 64
//            contiguous bytes, route to the destination frame/sub-pixel and store.
// --------------------------------------------------------------------------------------------
// packed ReLU6 on two 16-bit values: clamp(round(x)) == round(clamp(x)) because 0 and 6 are exact
template <bool BF16>
__device__ __forceinline__ uint32_t relu_packed(uint32_t u) {
  if constexpr (BF16) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    h = __hmax2(h, __float2bfloat162_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    h = __hmax2(h, __float2half2_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool BF16>
__device__ __forceinline__ uint32_t relu6_packed(uint32_t u) {
  if constexpr (BF16) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    h = __hmin2(__hmax2(h, __float2bfloat162_rn(0.f)), __float2bfloat162_rn(6.f));
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    h = __hmin2(__hmax2(h, __float2half2_rn(0.f)), __float2half2_rn(6.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

// Skip-tensor operand of one unit (lane = pixel): 64 contiguous bytes at the (PixelShuffle-
// scattered) output location.  Issued one unit ahead of its use so the L2 round trip overlaps the
// arithmetic of the previous unit.
// Register-resident copy of the fields the epilogue touches per unit (ConvParams lives in the
// kernel parameter bank; re-reading it per unit costs constant-cache latency on the epilogue's
// critical path).  skip_prefetch / epilogue_unit are duck-typed on it.
struct EpiParams {
  int flags, T, H, W, out_C, out_W, out_C_log2, fold, ring_mode, skip_C, resid_C, desc_variant, out_t0, pair_px, seg_T,
      out_pitch_C;
  uint32_t stg_bytes_per_warp;
  int src_H, src_W;
  void* out; void* out_prev; void* out_next; void* aux_out;
  const void* skip; const float* resid_in;
  unsigned* overflow;
  long long out_frame_stride, skip_frame_stride;
  __device__ __forceinline__ explicit EpiParams(const ConvParams& p)
      : flags(p.flags), T(p.T), H(p.H), W(p.W), out_C(p.out_C), out_W(p.out_W),
        out_C_log2(p.out_C_log2), fold(p.fold),
        ring_mode(p.ring_mode), skip_C(p.skip_C), resid_C(p.resid_C), desc_variant(p.desc_variant),
        out_t0(p.out_t0), pair_px(p.pair_px), seg_T(p.seg_T), out_pitch_C(p.out_pitch_C), stg_bytes_per_warp(p.stg_bytes_per_warp),
        src_H(p.src_H ? p.src_H : p.H), src_W(p.src_W ? p.src_W : (p.pair_px ? 2 * p.W : p.W)),
        out(p.out), out_prev(p.out_prev), out_next(p.out_next), aux_out(p.aux_out), skip(p.skip),
        resid_in(p.resid_in), overflow(p.overflow), out_frame_stride(p.out_frame_stride),
        skip_frame_stride(p.skip_frame_stride) {}
};

// Where one epilogue unit (32 GEMM columns x 32 pixels of one image row) lives in the output / skip
// tensor.  Everything here is warp-uniform (the compiler keeps it in uniform registers): a unit's 32
// columns share one PixelShuffle sub-pixel q because out_C is a power of two >= 32.
struct UnitPos {
  long long base;   // element offset inside a frame of (pixel 0 of the warp's quadrant, channel cb)
  int cb;           // first output channel of the unit
};
template <int MASK, class P>
__device__ __forceinline__ UnitPos unit_pos(const P& p, const TileCoord& tc, int y, int nbase, int quad) {
  UnitPos u;
  const int xb = tc.x0 + quad * 32;
  if (((MASK & EPI_PIXSHUF) != 0) && (p.flags & EPI_PIXSHUF)) {
    const int q = nbase >> p.out_C_log2;
    u.cb = nbase & (p.out_C - 1);
    u.base = (static_cast<long long>(2 * y + (q >> 1)) * p.out_W + 2 * xb + (q & 1)) * p.out_pitch_C + u.cb;
  } else {
    u.cb = nbase;
    u.base = (static_cast<long long>(y) * p.out_W + xb) * p.out_pitch_C + u.cb;
  }
  return u;
}
// Per-lane constants of the coalesced access pattern shared by the skip fetch and the store phase:
// lane = (pixel group lane>>2, 16-byte chunk lane&3), so four lanes cover the 64 contiguous bytes
// of one pixel; a lane serves pixels (lane>>2) + 8 i, i = 0..3.
struct EpiLane {
  uint32_t off;     // element offset of (pixel lane>>2, chunk lane&3) relative to UnitPos::base
  uint32_t step;    // element stride between a lane's four pixels
  int nvalid;       // valid pixels of the warp's quadrant in this tile (<= 0: none); per tile
};
template <int MASK, class P>
__device__ __forceinline__ EpiLane epi_lane(const P& p, int lane) {
  EpiLane l;
  const bool ps = ((MASK & EPI_PIXSHUF) != 0) && (p.flags & EPI_PIXSHUF);
  l.off = static_cast<uint32_t>((ps ? 2 : 1) * (lane >> 2) * p.out_pitch_C + 8 * (lane & 3));
  l.step = static_cast<uint32_t>((ps ? 16 : 8) * p.out_pitch_C);
  l.nvalid = 0;
  return l;
}

// Skip-tensor operand of one unit: 64 contiguous bytes per pixel at the (PixelShuffle-scattered)
// output location.  Issued one unit ahead of its use so the L2 round trip overlaps the arithmetic
// of the previous unit.  (The skip tensor has the output's shape: skip_C == out_C.)
template <int MASK, class P>
__device__ __forceinline__ void skip_prefetch(const P& p, const TileCoord& tc, const EpiLane& el, int y,
                                              int nbase, int quad, int lane, uint4 (&sk)[4]) {
  if constexpr ((MASK & EPI_SKIP) != 0) {
    if ((p.flags & EPI_SKIP) && tc.t < p.T && y < p.H && !(p.desc_variant & 64)) {
      const UnitPos up = unit_pos<MASK>(p, tc, y, nbase, quad);
      const uint16_t* src = reinterpret_cast<const uint16_t*>(p.skip) + tc.t * p.skip_frame_stride + up.base + el.off;
      const int pg = lane >> 2;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (pg + 8 * i < el.nvalid) sk[i] = __ldg(reinterpret_cast<const uint4*>(src + i * el.step));
    }
  }
}

// MASK = set of EPI_* features compiled into this instance (runtime flags are a subset of it).
// Store phase of one unit: park the lane's 64 bytes in the staging tile, then either one TMA store of the
// tile or the transposed read-back with 128-bit global stores routed to frame t-1 / t / t+1.  `coff` is
// the channel offset of this piece inside the pixel (0, or out_C for the lo half in split mode).
template <bool BF16, int MASK, class P>
__device__ __forceinline__ void epilogue_emit(const P& p, const TileCoord& tc, const EpiLane& el, int y, int nbase,
                                              const uint4 (&o)[4], uint32_t stg, int quad, int lane, int coff,
                                              const CUtensorMap* map_o) {
  const int flags = p.flags & MASK;
  {
    const uint32_t row = stg + lane * 64;
    const uint32_t swz = (lane >> 1) & 3;
    if constexpr ((MASK & EPI_TMA_OUT) != 0) {
      // the TMA store that last read this staging tile (the previous unit's, or with two tiles per
      // warp the one before) must have drained it; awaited as late as possible
      if (lane == 0) {
        if (p.stg_bytes_per_warp > kStageBytesPerWarp) bulk_wait_group_read<1>();
        else bulk_wait_group_read<0>();
      }
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((j ^ swz) << 4)),
                   "r"(o[j].x), "r"(o[j].y), "r"(o[j].z), "r"(o[j].w) : "memory");
  }
  if constexpr ((MASK & EPI_TMA_OUT) != 0) {
    // ------------------------------ phase 2 (TMA) ------------------------------
    // The staging tile [32 px][64 B] with its XOR pattern IS the SWIZZLE_64B box layout: one
    // cp.async.bulk.tensor store moves the unit; the tensor map clips pixels beyond the image.
    // With a temporal shift only the unit(s) holding the two folds take the routed LSU path below.
    bool routed = false;
    if constexpr ((MASK & EPI_SHIFT) != 0) {
      const int cb = (p.flags & EPI_PIXSHUF) ? (nbase & (p.out_C - 1)) : nbase;
      routed = (flags & EPI_SHIFT) && cb < 2 * p.fold;
    }
    if (routed) {
      if (lane == 0) bulk_commit_group();     // empty group: keeps the per-unit group count exact
    } else {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (el.nvalid > 0 && !(p.desc_variant & 128)) {
        int c0 = nbase + coff, c2 = y;
        if (p.flags & EPI_PIXSHUF) {
          const int q = nbase >> p.out_C_log2;
          c0 = (q & 1) * p.out_pitch_C + (nbase & (p.out_C - 1)) + coff;
          c2 = 2 * y + (q >> 1);
        }
        tma_store_4d(map_o, stg, c0, tc.x0 + quad * 32, c2, tc.t + p.out_t0);
      }
      bulk_commit_group();      // one group per unit, empty or not: keeps the wait_group count exact
    }
    return;
    }
  }
  __syncwarp();
  // ------------------------------ phase 2 ------------------------------
  {
    const int j = lane & 3;
    const UnitPos up = unit_pos<MASK>(p, tc, y, nbase, quad);
    // fold routing of this 8-channel group (ShiftConv.forward, bsvd_arch.py:42-50): the consumer
    // conv of frame u reads channels [0,f) of frame u+1 and [f,2f) of frame u-1, so the producer
    // of frame t stores those folds straight into the tensors of frames t-1 / t+1.
    uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + tc.t * p.out_frame_stride;
    uint16_t* zdst = nullptr;       // own-frame location that must read as zero (clip ends)
    if constexpr ((MASK & EPI_SHIFT) != 0) {
      if ((flags & EPI_SHIFT) && up.cb < 2 * p.fold) {     // (uniform) only the unit holding the folds
        const int c0 = up.cb + 8 * j;
        uint16_t* own = dst;
        // the T frames may be several independent clips of seg_T frames each (TSN train-mode shift,
        // temporal_shift.py:27-49; bsvd_forward_clips): folds never cross a clip boundary
        const int seg = p.seg_T > 0 ? p.seg_T : p.T;
        const int ts = tc.t % seg;
        if (c0 < p.fold) {
          if (p.ring_mode) {
            dst = reinterpret_cast<uint16_t*>(p.out_prev);
            if (p.flags & EPI_ZERO_FUTURE) zdst = own;
          } else {
            dst = (ts > 0) ? own - p.out_frame_stride : nullptr;
            if (ts == seg - 1) zdst = own;
          }
        } else if (c0 < 2 * p.fold) {
          if (p.ring_mode) {
            dst = reinterpret_cast<uint16_t*>(p.out_next);
            if (!p.out_prev) zdst = own;
          } else {
            dst = (ts + 1 < seg) ? own + p.out_frame_stride : nullptr;
            if (ts == 0) zdst = own;
          }
        }
      }
    }
    const long long off0 = up.base + el.off + coff;
    const bool row_ok = (y < p.H) && !(p.desc_variant & 128);
    const int pg = lane >> 2;
    const uint32_t a0 = stg + pg * 64;
    uint4 ob[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pl = 8 * i + pg;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(ob[i].x), "=r"(ob[i].y), "=r"(ob[i].z), "=r"(ob[i].w)
                   : "r"(a0 + i * 512 + ((j ^ ((pl >> 1) & 3)) << 4)) : "memory");
    }
    if (dst) dst += off0;
    if constexpr ((MASK & EPI_SHIFT) != 0) {
      if (zdst) zdst += off0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (row_ok && pg + 8 * i < el.nvalid) {
        if (dst) *reinterpret_cast<uint4*>(dst + i * el.step) = ob[i];
        if constexpr ((MASK & EPI_SHIFT) != 0) {
          if (zdst) *reinterpret_cast<uint4*>(zdst + i * el.step) = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  __syncwarp();
}

// MASK = set of EPI_* features compiled into this instance (runtime flags are a subset of it).
template <bool BF16, int MASK, class P>
__device__ __forceinline__ void epilogue_unit(const P& p, const TileCoord& tc, const EpiLane& el, int y,
                                              int nbase, const uint32_t (&v)[32],
                                              const uint4 (&sk)[4],
                                              const float (&bv)[32], uint32_t stg, int quad,
                                              int lane, const float (&rin)[3], bool use_rin,
                                              const CUtensorMap* map_o = nullptr) {
  const int flags = p.flags & MASK;
  constexpr bool kSplit = (MASK & EPI_SPLIT) != 0;
  static_assert(!(kSplit && BF16), "the fp32-grade split uses fp16 pieces");
  // Instances compiled with EPI_RELU6 but without EPI_RELU serve only stages that DO apply ReLU6 (the
  // dispatch in stage_launch.cuh guarantees it): the clamp is unconditional, and the fp16 range guard and the
  // plain-ReLU path — predicated off but still issued in the general instances — are not compiled at all.
  constexpr bool kOnly6 = (MASK & EPI_RELU6) != 0 && (MASK & EPI_RELU) == 0 && !kSplit;

  // ------------------------------ phase 0 ------------------------------
  // hand the coalesced skip operand over to the pixel-owning lanes through the staging tile
  if constexpr ((MASK & EPI_SKIP) != 0) {
    if (flags & EPI_SKIP) {
      const int j = lane & 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pl = 8 * i + (lane >> 2);
        const uint32_t a = stg + pl * 64 + ((j ^ ((pl >> 1) & 3)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(sk[i].x), "r"(sk[i].y),
                     "r"(sk[i].z), "r"(sk[i].w) : "memory");
      }
      __syncwarp();
    }
  }
  // ------------------------------ phase 1 ------------------------------
  // (shared-memory accesses are batched: all loads of a phase are issued before their first use)
  uint4 o[4];
  uint4 olo[kSplit ? 4 : 1];
  {
    const int x = tc.x0 + quad * 32 + lane;
    const bool valid = (lane < el.nvalid) && (y < p.H);
    const uint32_t row = stg + lane * 64;
    const uint32_t swz = (lane >> 1) & 3;
    uint4 sv[4];
    bool add_skip = false;
    if constexpr ((MASK & EPI_SKIP) != 0) {
      add_skip = (flags & EPI_SKIP) && valid;
      if (flags & EPI_SKIP) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(sv[j].x), "=r"(sv[j].y), "=r"(sv[j].z), "=r"(sv[j].w)
                       : "r"(row + ((j ^ swz) << 4)) : "memory");
      }
    }
    // range guard: running maximum of what this lane stores (|x|, or x under ReLU, which clamps the
    // negative side); ReLU6 stages are bounded and skip it
    const bool chk_range = !BF16 && !kOnly6 && !(flags & EPI_RELU6) && p.overflow != nullptr;
    const bool chk_signed = (flags & EPI_RELU) != 0;
    float vmax = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[8 * j + i]) + bv[8 * j + i];
      if constexpr ((MASK & EPI_SKIP) != 0) {
        if (add_skip) {
          const float2 a = unpack2<BF16>(sv[j].x), b = unpack2<BF16>(sv[j].y),
                       c = unpack2<BF16>(sv[j].z), d = unpack2<BF16>(sv[j].w);
          f[0] += a.x; f[1] += a.y; f[2] += b.x; f[3] += b.y;
          f[4] += c.x; f[5] += c.y; f[6] += d.x; f[7] += d.y;
        }
      }
      if constexpr ((MASK & EPI_RESID_IN) != 0) {
        // pair mode: unit 0 / 1 of a row hold pixel 2x / 2x+1, each with its own channels 0..2
        const bool resid_unit = p.pair_px ? ((nbase & 31) == 0) : (nbase == 0);
        if ((flags & EPI_RESID_IN) && j == 0 && resid_unit && valid) {
          // temp1 residual (bsvd_arch.py:394, 408-414): out[:, :3] = in[:, :3] - out[:, :3]
          if (use_rin) {
#pragma unroll
            for (int i = 0; i < 3; ++i) f[i] = rin[i] - f[i];
          } else {
            const int xr = p.pair_px ? 2 * x + (nbase >> 5) : x;
            const int sH = p.src_H ? p.src_H : p.H, sW = p.src_W ? p.src_W : (p.pair_px ? 2 * p.W : p.W);
            const long long plane = static_cast<long long>(sH) * sW;
            const float* r = p.resid_in + (static_cast<long long>(tc.t) * p.resid_C) * plane +
                             static_cast<long long>(reflect_src(y, sH)) * sW + reflect_src(xr, sW);
#pragma unroll
            for (int i = 0; i < 3; ++i) f[i] = __ldg(r + i * plane) - f[i];
          }
        }
      }
      if constexpr (kSplit) {
        // fp32-grade mode: the activation is applied in fp32, then the value leaves as hi + lo
        if (flags & EPI_RELU6) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = relu6f(f[i]);
        } else if (flags & EPI_RELU) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
        }
      }
      if constexpr (!BF16) {
        if (chk_range) {
          if (chk_signed) {
            vmax = fmaxf(fmaxf(vmax, f[0]), fmaxf(f[1], f[2]));
            vmax = fmaxf(fmaxf(vmax, f[3]), fmaxf(f[4], f[5]));
            vmax = fmaxf(vmax, fmaxf(f[6], f[7]));
          } else {
            vmax = fmaxf(fmaxf(vmax, fabsf(f[0])), fmaxf(fabsf(f[1]), fabsf(f[2])));
            vmax = fmaxf(fmaxf(vmax, fabsf(f[3])), fmaxf(fabsf(f[4]), fabsf(f[5])));
            vmax = fmaxf(vmax, fmaxf(fabsf(f[6]), fabsf(f[7])));
          }
        }
      }
      o[j].x = pack2<BF16>(f[0], f[1]); o[j].y = pack2<BF16>(f[2], f[3]);
      o[j].z = pack2<BF16>(f[4], f[5]); o[j].w = pack2<BF16>(f[6], f[7]);
      if constexpr (kSplit) {
        const float2 h0 = unpack2<false>(o[j].x), h1 = unpack2<false>(o[j].y),
                     h2 = unpack2<false>(o[j].z), h3 = unpack2<false>(o[j].w);
        olo[j].x = pack2<false>(f[0] - h0.x, f[1] - h0.y); olo[j].y = pack2<false>(f[2] - h1.x, f[3] - h1.y);
        olo[j].z = pack2<false>(f[4] - h2.x, f[5] - h2.y); olo[j].w = pack2<false>(f[6] - h3.x, f[7] - h3.y);
      }
      if constexpr ((MASK & EPI_RESID_IN) != 0) {
        const bool aux_unit = p.pair_px ? ((nbase & 31) == 0) : (nbase == 0);
        if ((flags & EPI_RESID_IN) && j == 0 && aux_unit && valid && p.aux_out) {
          const long long pix = p.pair_px
              ? (static_cast<long long>(tc.t) * p.H + y) * (2 * p.W) + 2 * x + (nbase >> 5)
              : (static_cast<long long>(tc.t) * p.H + y) * p.W + x;
          if constexpr (kSplit) reinterpret_cast<float4*>(p.aux_out)[pix] = make_float4(f[0], f[1], f[2], f[3]);
          else reinterpret_cast<uint2*>(p.aux_out)[pix] = make_uint2(o[j].x, o[j].y);
        }
      }
      if constexpr (!kSplit) {
        if (kOnly6 || (flags & EPI_RELU6)) {
          o[j].x = relu6_packed<BF16>(o[j].x); o[j].y = relu6_packed<BF16>(o[j].y);
          o[j].z = relu6_packed<BF16>(o[j].z); o[j].w = relu6_packed<BF16>(o[j].w);
        } else if (flags & EPI_RELU) {
          o[j].x = relu_packed<BF16>(o[j].x); o[j].y = relu_packed<BF16>(o[j].y);
          o[j].z = relu_packed<BF16>(o[j].z); o[j].w = relu_packed<BF16>(o[j].w);
        }
      }
    }
    if constexpr (!BF16) {
      // fp16 range guard (see ConvParams::overflow): a value that rounds to inf (|x| >= 65520).  A NaN can
      // only come from an inf stored by an earlier stage, which raised the (sticky) flag there.
      if (chk_range && __any_sync(0xffffffffu, valid && !(vmax < 65520.f)) && lane == 0) atomicOr(p.overflow, 1u);
    }
  }
  if constexpr (kSplit) {
    // two staging tiles per warp: hi through the first, lo through the second
    epilogue_emit<BF16, MASK>(p, tc, el, y, nbase, o, stg, quad, lane, 0, map_o);
    epilogue_emit<BF16, MASK>(p, tc, el, y, nbase, olo, stg + kStageBytesPerWarp, quad, lane, p.out_C, map_o);
  } else {
    epilogue_emit<BF16, MASK>(p, tc, el, y, nbase, o, stg, quad, lane, 0, map_o);
  }
}

// --------------------------------------------------------------------------------------------
// The kernel
// --------------------------------------------------------------------------------------------
// PIPE: compile-time pipeline shape, so the producer / MMA loops carry no per-tap parameter tests:
//   0 halo tile + streamed filter slabs   1 stride-2 per-tap boxes + streamed slabs (generic only)
//   2 stacked 64->64 (resident bank)      3 generic: mode / w_resident read from ConvParams
//   4 stride-2 sub-plane boxes + streamed slabs (map_s = the R-row box map of the even input rows)
//   5 the same on a 32-channel input read as pixel pairs (stride 2 in y only; c32 configurations)
template <int NTILE, int R, bool BF16, bool CTA2, int MASK, int EW, int PIPE>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_o,
                  const __grid_constant__ ConvParams p) {
  static_assert(NTILE == 64 || NTILE == 128 || NTILE == 256, "unsupported NTILE");
  constexpr bool kFastDiv = (MASK & EPI_SKIP) == 0;   // see ConvParams::div_nt
  constexpr int kAccCols = R * NTILE;                 // TMEM columns of one accumulator set
  // two accumulator sets (epilogue of tile i overlaps the MMAs of tile i+1) when they fit in the
  // 512 TMEM columns, otherwise one set (bigger tile: every filter slab is reused for R rows)
  constexpr int kNumAcc = (2 * kAccCols <= 512) ? 2 : 1;
  constexpr int kTmemCols = (kNumAcc * kAccCols < 32) ? 32 : kNumAcc * kAccCols;
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;      // 0 = leader of the CTA pair
  const int tile0 = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;
  const int tstep = CTA2 ? (gridDim.x >> 1) : gridDim.x;

  // 1024-byte aligned operand area (swizzle pattern anchor)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + p.a_stages * p.a_stage_bytes;
  const uint32_t stg_base = w_base + p.w_stages * p.w_stage_bytes;
  const uint32_t id_base = stg_base + EW * p.stg_bytes_per_warp;   // 64x64 identity (skip_mma), 4 KB per CTA

  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (3 * kMaxStages + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (4 * kMaxStages + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (4 * kMaxStages + 2 + b); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), CTA2 ? 2 * EW : EW);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if constexpr (CTA2 && R == 1) {
    if (p.skip_mma) {
      // This CTA's half of the 64x64 identity B operand (N rows [32 rank, +32), K-major, SW128):
      // the skip tensor goes through the tensor core as D += S * I (exact: 16-bit x 1.0 into fp32).
      for (int i = threadIdx.x; i < 256; i += 64 + 32 * EW) {
        const int row = i >> 3, cphys = i & 7;
        const int n = 32 * static_cast<int>(rank) + row;
        const uint32_t one = BF16 ? 0x3F80u : 0x3C00u;
        const bool hit = (cphys ^ (row & 7)) == (n >> 3);
        const uint32_t v = one << (16 * (n & 1));
        const int wi = (n & 7) >> 1;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(id_base + i * 16),
                     "r"((hit && wi == 0) ? v : 0u), "r"((hit && wi == 1) ? v : 0u),
                     "r"((hit && wi == 2) ? v : 0u), "r"((hit && wi == 3) ? v : 0u) : "memory");
      }
      fence_proxy_async();
    }
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if constexpr (NTILE == 64 && R == 2 && CTA2) {
    if ((PIPE == 3 ? p.mode : PIPE) == 2) {
      // stacked mode accumulates from the first MMA on: start from zeroed accumulators
      if (warp >= 2) {
        const uint32_t lb = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int part = (warp - 2) >> 2;
        for (int c = part * 32; c < kTmemCols; c += 32 * (EW / 4)) tmem_st32_zero(tmem_base + lb + c);
        tmem_st_wait();
      }
      tc_fence_before();
      cluster_sync_all();
      tc_fence_after();
    }
  }
  pdl_launch_dependents();
  pdl_wait();            // everything above touched only weights/bias; activations come next

  // Local copy of the pipeline parameters for the producer / MMA loops.  NOTE (measured): forcing these
  // into general registers (opaque asm) makes every non-stacked stage 15-25 % slower, because the MMA
  // issue loop then leaves the uniform datapath (R2UR in front of every descriptor); left alone, ptxas
  // re-materialises them from the constant bank with uniform loads, which is the faster of the two.
  struct {
    int a_stages, w_stages, cin_chunks, tap_begin, tap_end, mode, w_resident, total_tiles, desc_variant,
        skip_mma, w_rows_cta, cin_total, out_C_log2, out_C, skip_t0, phys_chunks, out_pitch_C, skip_blocks;
    uint32_t a_stage_bytes, w_stage_bytes, a_tx_bytes;
  } const pp = {p.a_stages, p.w_stages, p.cin_chunks, PIPE == 3 ? p.tap_begin : 0,
                PIPE == 3 ? p.tap_end : (PIPE == 2 ? 3 : PIPE == 5 ? 6 : 9),
                PIPE == 3 ? p.mode : PIPE, PIPE == 3 ? p.w_resident : (PIPE == 2 ? 1 : 0),
                p.total_tiles, p.desc_variant, p.skip_mma, p.w_rows_cta, p.cin_total, p.out_C_log2, p.out_C,
                p.skip_t0, p.phys_chunks > 0 ? p.phys_chunks : p.cin_chunks, p.out_pitch_C,
                p.skip_blocks > 0 ? p.skip_blocks : (p.skip_mma > 0 ? p.skip_mma : 1),
                p.a_stage_bytes, p.w_stage_bytes, p.a_tx_bytes};
  const int ntaps = pp.tap_end - pp.tap_begin;
  // bytes one stage receives in total (both CTAs of a pair signal the leader's barrier)
  const uint32_t a_tx = CTA2 ? 2 * pp.a_tx_bytes : pp.a_tx_bytes;
  const uint32_t w_tx = CTA2 ? 2 * pp.w_stage_bytes : pp.w_stage_bytes;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0 && !(pp.desc_variant & 16)) {
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      bool first = true;
      for (int tile = tile0; tile < pp.total_tiles; tile += tstep) {
        const TileCoord tc = decode_tile<R, kFastDiv>(p, tile, CTA2, rank);
        int nskip = 0;
        if constexpr ((PIPE == 4 || PIPE == 5) && CTA2) {
          // stride 2, sub-plane boxes: the A boxes and filter slabs of a chunk, issued in consumption order
          for (int c = 0; c < pp.cin_chunks; ++c) {
#pragma unroll
            for (int j = 0; j < s2_ntaps<PIPE>(); ++j) {
              if (s2_box_first<PIPE>(j)) {
                const int b = s2_box<PIPE>(j);
                const int py = s2_box_py<PIPE>(b), px = s2_box_px<PIPE>(b);
                const uint32_t bytes = static_cast<uint32_t>((py ? R + 1 : R) * kS2BoxPx * 128);
                mbar_wait(a_empty(sa), pa ^ 1);
                if (rank == 0) mbar_expect_tx(a_full(sa), 2 * bytes);
                tma_load_5d_2sm(a_base + sa * pp.a_stage_bytes, py ? &map_a : &map_s, a_full(sa),
                                px * pp.cin_total + (c % pp.phys_chunks) * kChunk, (px || PIPE == 5) ? tc.x0 - 1 : tc.x0, py,
                                py ? tc.y0 - 1 : tc.y0, tc.t);
                if (++sa == (uint32_t)pp.a_stages) { sa = 0; pa ^= 1; }
              }
              mbar_wait(w_empty(sw), pw ^ 1);
              if (rank == 0) mbar_expect_tx(w_full(sw), w_tx);
              const size_t blk = static_cast<size_t>(tc.nt * pp.cin_chunks + c) * s2_ntaps<PIPE>() + s2_slab<PIPE>(j);
              tma_load_2d_2sm(w_base + sw * pp.w_stage_bytes, &map_w, w_full(sw), 0,
                              (static_cast<int>(blk) * 2 + static_cast<int>(rank)) * pp.w_rows_cta);
              if (++sw == (uint32_t)pp.w_stages) { sw = 0; pw ^= 1; }
            }
          }
          continue;
        }
        for (int c = 0; c < pp.cin_chunks; ++c) {
          if (pp.mode != 1) {
            mbar_wait(a_empty(sa), pa ^ 1);
            if (rank == 0) mbar_expect_tx(a_full(sa), a_tx);
            if constexpr (CTA2)
              tma_load_4d_2sm(a_base + sa * pp.a_stage_bytes, &map_a, a_full(sa), (c % pp.phys_chunks) * kChunk,
                              tc.x0 - 1, tc.y0 - 1, tc.t);
            else
              tma_load_4d(a_base + sa * pp.a_stage_bytes, &map_a, a_full(sa), (c % pp.phys_chunks) * kChunk, tc.x0 - 1,
                          tc.y0 - 1, tc.t);
            if (++sa == (uint32_t)pp.a_stages) { sa = 0; pa ^= 1; }
          }
          for (int tap = pp.tap_begin; tap < pp.tap_end; ++tap) {
            if (pp.mode == 1) {
              // stride 2: input (2y+dy-1, 2x+dx-1) seen through the [T][H/2][2][W/2][2*C] view
              const int dy = tap / 3, dx = tap - dy * 3;
              const int px = (dx == 1) ? 0 : 1, x2 = tc.x0 + ((dx == 0) ? -1 : 0);
              const int py = (dy == 1) ? 0 : 1, y2 = tc.y0 + ((dy == 0) ? -1 : 0);
              mbar_wait(a_empty(sa), pa ^ 1);
              if (rank == 0) mbar_expect_tx(a_full(sa), a_tx);
              if constexpr (CTA2)
                tma_load_5d_2sm(a_base + sa * pp.a_stage_bytes, &map_a, a_full(sa),
                                px * pp.cin_total + c * kChunk, x2, py, y2, tc.t);
              else
                tma_load_5d(a_base + sa * pp.a_stage_bytes, &map_a, a_full(sa),
                            px * pp.cin_total + c * kChunk, x2, py, y2, tc.t);
              if (++sa == (uint32_t)pp.a_stages) { sa = 0; pa ^= 1; }
            }
            if (!pp.w_resident || first) {
              mbar_wait(w_empty(sw), pw ^ 1);
              if (rank == 0) mbar_expect_tx(w_full(sw), w_tx);
              const size_t blk = static_cast<size_t>(tc.nt * pp.cin_chunks + c) * ntaps + (tap - pp.tap_begin);
              if constexpr (CTA2) {
                // each CTA of the pair stages half of the N rows of this (chunk, tap) filter slab
                tma_load_2d_2sm(w_base + sw * pp.w_stage_bytes, &map_w, w_full(sw), 0,
                                (static_cast<int>(blk) * 2 + static_cast<int>(rank)) * pp.w_rows_cta);
              } else {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wpack) + blk * pp.w_stage_bytes;
                bulk_load(w_base + sw * pp.w_stage_bytes, src, pp.w_stage_bytes, w_full(sw));
              }
              if (++sw == (uint32_t)pp.w_stages) { sw = 0; pw ^= 1; }
            }
            if constexpr (CTA2 && R == 1 && (MASK & EPI_PIXSHUF) != 0) {
              // Skip operand (ConvParams::skip_mma): block bb = GEMM columns [64 bb, +64) = 64 channels
              // of one sub-pixel; in the [T][2H][W][2][C] view of the skip tensor that is a plain
              // [128 px][64 ch] box, exactly one filter-ring slot.  The blocks ride in the FILTER ring,
              // one after every fourth slab: its slots turn over every few hundred cycles and it is
              // five deep, so a stage with almost no MMA work does not let the pipeline run dry
              // (in the 2-deep activation ring it did).
              if (nskip < pp.skip_mma && (((c * ntaps + tap - pp.tap_begin) + 1) & 3) == 0) {
                mbar_wait(w_empty(sw), pw ^ 1);
                if (rank == 0) mbar_expect_tx(w_full(sw), w_tx);
                // split mode: blocks [0, skip_blocks) are the hi halves, the next skip_blocks the lo halves
                const int n = tc.nt * NTILE + (nskip % pp.skip_blocks) * 64;
                const int q = n >> pp.out_C_log2, ch = n & (pp.out_C - 1);
                tma_load_4d_2sm(w_base + sw * pp.w_stage_bytes, &map_s, w_full(sw),
                                (q & 1) * pp.out_pitch_C + ch + (nskip >= pp.skip_blocks ? pp.out_C : 0),
                                tc.x0, 2 * tc.y0 + (q >> 1), tc.t + pp.skip_t0);
                ++nskip;
                if (++sw == (uint32_t)pp.w_stages) { sw = 0; pw ^= 1; }
              }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    // The whole warp walks the (uniform) tile/chunk/tap loops and waits on the barriers; one
    // elected lane issues the tcgen05.mma / tcgen05.commit instructions.  Descriptors are built
    // once; per MMA only the 32-bit start-address word changes by a compile-time constant.
    // In a CTA pair only the leader's warp issues: one M=256 instruction covers both pixel tiles.
    if (rank == 0) {
      const uint32_t idesc = make_idesc(NTILE, BF16 ? 1 : 0, CTA2 ? 256 : 128);
      const uint32_t leader = elect_one();
      constexpr uint32_t kDescHi = 0x40000000u | (1u << 14) | (1024u >> 4);  // SW128, version 1, SBO 1024
      const uint64_t desc_hi = static_cast<uint64_t>(kDescHi) << 32;
      const bool no_mma = (pp.desc_variant & 2) != 0;
      const bool no_load = (pp.desc_variant & 16) != 0;
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      uint32_t it = 0;
      bool stacked = false;
      if constexpr (NTILE == 64 && R == 2 && CTA2) stacked = (pp.mode == 2);
      if (stacked) {
        // ---- 64->64 stages, vertical taps stacked in N -------------------------------------------
        // The two output-row accumulators sit side by side in TMEM (columns [0,64) and [64,128)).
        // Haloed INPUT row hr feeds output row hr-dy with filter row dy, so one MMA per (dx, k-step)
        // with the filter blocks ordered to match covers all output rows it touches:
        //   hr 0: [dy0]      -> acc0          N=64      hr 2: [dy2|dy1] -> acc0|acc1   N=128
        //   hr 1: [dy1|dy0]  -> acc0|acc1     N=128     hr 3: [dy2]     -> acc1        N=64
        // 48 MMAs per tile instead of 72 and a quarter fewer operand bytes, same FLOPs.  Every MMA
        // accumulates; the epilogue hands the accumulators back zeroed.
        const uint32_t idesc64 = make_idesc(64, BF16 ? 1 : 0, 256);
        const uint32_t idesc128 = make_idesc(128, BF16 ? 1 : 0, 256);
        const uint32_t b_lo0 = ((w_base & 0x3FFFFu) >> 4) | (1u << 16);
        for (int tile = tile0; tile < pp.total_tiles; tile += tstep, ++it) {
          const uint32_t buf = it & 1, acc_phase = (it >> 1) & 1;
          mbar_wait(acc_empty(buf), acc_phase ^ 1);
          if (!no_load) {
            mbar_wait(a_full(sa), pa);
            if (it == 0) { mbar_wait(w_full(0), 0); mbar_wait(w_full(1), 0); mbar_wait(w_full(2), 0); }
          }
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + buf * kAccCols;
          const uint32_t a_lo0 = (((a_base + sa * pp.a_stage_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          if (leader && !no_mma) {
#pragma unroll
            for (int hr = 0; hr < 4; ++hr) {
              const uint32_t col0 = (hr == 3) ? 64u : 0u;
              const uint32_t boff = (hr == 0 ? 0u : hr == 1 ? 4096u : hr == 2 ? 12288u : 20480u) >> 4;
              const uint32_t idesc_h = (hr == 0 || hr == 3) ? idesc64 : idesc128;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t ad = desc_hi | (a_lo0 + static_cast<uint32_t>((hr * kHaloPx + dx) * 8) + k * 2u);
                  const uint64_t bd = desc_hi | (b_lo0 + static_cast<uint32_t>(dx * (24576 >> 4)) + boff + k * 2u);
                  umma_f16_2sm(tmem_acc + col0, ad, bd, idesc_h, 1u);
                }
              }
            }
          }
          if (leader) {
            umma_commit_2sm(a_empty(sa));
            umma_commit_2sm(acc_full(buf));
          }
          __syncwarp();
          if (++sa == (uint32_t)pp.a_stages) { sa = 0; pa ^= 1; }
        }
      } else {
      // With a compile-time pipeline shape (PIPE != 3) the loop carries no debug switches and the
      // tap range is 0..9: everything per tap is one barrier wait, the MMAs, one commit.
      constexpr bool kGen = (PIPE == 3);
      const int tap_b = kGen ? pp.tap_begin : 0, tap_e = kGen ? pp.tap_end : 9;
      const bool dbg_no_load = kGen && no_load, dbg_no_mma = kGen && no_mma;
      const bool dbg_flat_a = kGen && (pp.desc_variant & 8);
      const uint32_t a_lo_base = ((a_base & 0x3FFFFu) >> 4) | (1u << 16), a_lo_step = pp.a_stage_bytes >> 4;
      const uint32_t w_lo_base = ((w_base & 0x3FFFFu) >> 4) | (1u << 16), w_lo_step = pp.w_stage_bytes >> 4;
      for (int tile = tile0; tile < pp.total_tiles; tile += tstep, ++it) {
        const uint32_t buf = (kNumAcc == 2) ? (it & 1) : 0u;
        const uint32_t acc_phase = (kNumAcc == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(acc_empty(buf), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * kAccCols;
        int nskip = 0;
        if constexpr ((PIPE == 4 || PIPE == 5) && CTA2) {
          for (int c = 0; c < pp.cin_chunks; ++c) {
            uint32_t a_lo0 = 0;
#pragma unroll
            for (int j = 0; j < s2_ntaps<PIPE>(); ++j) {
              if (s2_box_first<PIPE>(j)) {
                mbar_wait(a_full(sa), pa);
                a_lo0 = a_lo_base + sa * a_lo_step;
              }
              mbar_wait(w_full(sw), pw);
              tc_fence_after();
              const int ro = s2_row_off<PIPE>(j), co = s2_col_off<PIPE>(j);   // offset inside the sub-plane box
              const uint32_t b_lo0 = w_lo_base + sw * w_lo_step;
              const uint32_t first = (c == 0 && j == 0) ? 0u : 1u;
              if (leader) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const uint32_t a_off = static_cast<uint32_t>(((r + ro) * kS2BoxPx + co) * 8);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(tmem_acc + r * NTILE, desc_hi | (a_lo0 + a_off + k * 2u), desc_hi | (b_lo0 + k * 2u),
                                 idesc, (k > 0) ? 1u : first);
                }
                umma_commit_2sm(w_empty(sw));
                if (s2_box_last<PIPE>(j)) umma_commit_2sm(a_empty(sa));
              }
              __syncwarp();
              if (++sw == (uint32_t)pp.w_stages) { sw = 0; pw ^= 1; }
              if (s2_box_last<PIPE>(j)) {
                if (++sa == (uint32_t)pp.a_stages) { sa = 0; pa ^= 1; }
              }
            }
          }
          if (leader) umma_commit_2sm(acc_full(buf));
          __syncwarp();
          continue;
        }
        for (int c = 0; c < pp.cin_chunks; ++c) {
          if (pp.mode == 0 && !dbg_no_load) {
            mbar_wait(a_full(sa), pa);
            tc_fence_after();
          }
          uint32_t a_lo0 = a_lo_base + sa * a_lo_step;     // mode 1: recomputed per tap below
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (kGen && (tap < tap_b || tap >= tap_e)) continue;
            if (pp.mode == 1) {
              if (!dbg_no_load) mbar_wait(a_full(sa), pa);
              a_lo0 = a_lo_base + sa * a_lo_step;
            }
            if ((!pp.w_resident || it == 0) && !dbg_no_load) mbar_wait(w_full(sw), pw);
            tc_fence_after();
            const int dy = tap / 3, dx = tap % 3;
            const uint32_t b_lo0 = w_lo_base + sw * w_lo_step;
            const uint32_t first = (c == 0 && tap == tap_b) ? 0u : 1u;
            if (leader && !dbg_no_mma) {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const uint32_t a_off = (pp.mode == 0 && !dbg_flat_a)
                    ? static_cast<uint32_t>(((r + dy) * kHaloPx + dx) * 8)
                    : static_cast<uint32_t>(r * (kRunPx * 8));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t ad = desc_hi | (a_lo0 + a_off + k * 2u);
                  const uint64_t bd = desc_hi | (b_lo0 + k * 2u);
                  if constexpr (CTA2)
                    umma_f16_2sm(tmem_acc + r * NTILE, ad, bd, idesc, (k > 0) ? 1u : first);
                  else
                    umma_f16(tmem_acc + r * NTILE, ad, bd, idesc, (k > 0) ? 1u : first);
                }
              }
            }
            if (leader) {
              if constexpr (CTA2) {
                if (!pp.w_resident) umma_commit_2sm(w_empty(sw));
                if (pp.mode == 1) umma_commit_2sm(a_empty(sa));
              } else {
                if (!pp.w_resident) umma_commit(w_empty(sw));
                if (pp.mode == 1) umma_commit(a_empty(sa));
              }
            }
            __syncwarp();
            if (++sw == (uint32_t)pp.w_stages) { sw = 0; pw ^= 1; }
            if (pp.mode == 1) {
              if (++sa == (uint32_t)pp.a_stages) { sa = 0; pa ^= 1; }
            }
            if constexpr (CTA2 && R == 1 && (MASK & EPI_PIXSHUF) != 0) {
              // skip block in the filter ring (see the producer): D[:, 64 bb .. +64) += S_bb * I
              if (nskip < pp.skip_mma && (((c * (tap_e - tap_b) + tap - tap_b) + 1) & 3) == 0) {
                if (!dbg_no_load) mbar_wait(w_full(sw), pw);
                tc_fence_after();
                const uint32_t s_lo0 = w_lo_base + sw * w_lo_step;
                const uint32_t id_lo0 = ((id_base & 0x3FFFFu) >> 4) | (1u << 16);
                if (leader && !dbg_no_mma) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(tmem_acc + (nskip % pp.skip_blocks) * 64, desc_hi | (s_lo0 + k * 2u), desc_hi | (id_lo0 + k * 2u),
                                 make_idesc(64, BF16 ? 1 : 0, 256), 1u);
                }
                if (leader) umma_commit_2sm(w_empty(sw));
                __syncwarp();
                ++nskip;
                if (++sw == (uint32_t)pp.w_stages) { sw = 0; pw ^= 1; }
              }
            }
          }
          if (pp.mode == 0) {
            if (leader) { if constexpr (CTA2) umma_commit_2sm(a_empty(sa)); else umma_commit(a_empty(sa)); }
            __syncwarp();
            if (++sa == (uint32_t)pp.a_stages) { sa = 0; pa ^= 1; }
          }
        }
        if (leader) { if constexpr (CTA2) umma_commit_2sm(acc_full(buf)); else umma_commit(acc_full(buf)); }
        __syncwarp();
      }
      }
    }
  } else {
    // ======================================= epilogue =======================================
    // warp index through a shuffle: tells the compiler it is warp-uniform, so everything derived
    // from it (unit index, bias offset) lives in uniform registers / uniform constant loads
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int ew = warp_u - 2;                 // 0..7
    const int quad = warp_u & 3;               // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                  // EW/4 warps share a TMEM lane quadrant and split the units
    const uint32_t stg0 = stg_base + ew * p.stg_bytes_per_warp;
    uint32_t ucount = 0;                       // TMA stores: units alternate between two staging tiles
    auto next_stg = [&]() -> uint32_t {
      if constexpr ((MASK & EPI_SPLIT) != 0) return stg0;      // a unit uses both tiles itself (hi, lo)
      else if constexpr ((MASK & EPI_TMA_OUT) != 0)
        return stg0 + ((p.stg_bytes_per_warp > kStageBytesPerWarp) ? (ucount++ & 1u) * kStageBytesPerWarp : 0u);
      else return stg0;
    };
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    // register copy of the epilogue parameters, except in the register-starved general instance
    using EP = typename std::conditional<(MASK & EPI_SKIP) != 0, const ConvParams&, const EpiParams>::type;
    EP e(p);
    EpiLane el = epi_lane<MASK>(e, lane);
    uint32_t it = 0;
    // temp1 residual operand (raw fp32 network input, channels 0..2) of this warp's first unit,
    // fetched one whole tile ahead: a first-touch DRAM read takes longer than one tile's MMAs
    float rin_next[3] = {0.f, 0.f, 0.f};
    auto load_rin = [&](int tile_idx) {
      if constexpr ((MASK & EPI_RESID_IN) != 0) {
        constexpr int G = NTILE / 32;
        const int u0 = (ew >> 2) * ((R * G) / (EW / 4));
        if ((e.flags & EPI_RESID_IN) && tile_idx < p.total_tiles) {
          const TileCoord tn = decode_tile<R, kFastDiv>(p, tile_idx, CTA2, rank);
          const int y = tn.y0 + u0 / G, x = tn.x0 + quad * 32 + lane;
          if (tn.nt * NTILE + (u0 % G) * 32 == 0 && tn.t < e.T && y < e.H && x < e.W) {
            const long long plane = static_cast<long long>(e.src_H) * e.src_W;
            const int xr = e.pair_px ? 2 * x : x;          // unit 0 of a pair row is pixel 2x
            const float* r = e.resid_in + (static_cast<long long>(tn.t) * e.resid_C) * plane +
                             static_cast<long long>(reflect_src(y, e.src_H)) * e.src_W + reflect_src(xr, e.src_W);
#pragma unroll
            for (int i = 0; i < 3; ++i) rin_next[i] = __ldg(r + i * plane);
          }
        }
      }
    };
    load_rin(tile0);
    for (int tile = tile0; tile < p.total_tiles; tile += tstep, ++it) {
      const TileCoord tc = decode_tile<R, kFastDiv>(p, tile, CTA2, rank);
      const uint32_t buf = (kNumAcc == 2) ? (it & 1) : 0u;
      const uint32_t acc_phase = (kNumAcc == 2) ? ((it >> 1) & 1) : (it & 1);
      constexpr int G = NTILE / 32;            // 32-column groups per row
      constexpr int kUnits = R * G;
      static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
      constexpr int kMine = kUnits / (EW / 4);
      static_assert(kMine >= 1 && kMine * (EW / 4) == kUnits, "units must split evenly over the warps of a quadrant");
      const int u0 = half * kMine;
      const int nb0 = tc.nt * NTILE;           // global GEMM column of this tile's first column
      el.nvalid = (tc.t < e.T) ? min(32, e.W - (tc.x0 + quad * 32)) : 0;
      // skip operand of the first unit: issued before waiting for the accumulator, so its latency
      // is covered by the MMAs of this very tile; later units are prefetched one unit ahead
      uint4 ska[4] = {}, skb[4] = {};
      skip_prefetch<MASK>(e, tc, el, tc.y0 + u0 / G, nb0 + (u0 % G) * 32, quad, lane, ska);
      float rin[3] = {rin_next[0], rin_next[1], rin_next[2]};
      bool rin_ok = false;
      if constexpr ((MASK & EPI_RESID_IN) != 0) {
        rin_ok = (e.flags & EPI_RESID_IN) && nb0 + (u0 % G) * 32 == 0;
        load_rin(tile + tstep);
      }
      mbar_wait(acc_full(buf), acc_phase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + buf * kAccCols;
      // accumulator drained: the (leader's) MMA warp may reuse the buffer
      bool stacked = false;
      if constexpr (NTILE == 64 && R == 2 && CTA2) stacked = ((PIPE == 3 ? p.mode : PIPE) == 2);
      auto release_acc = [&]() {
        if (stacked) tmem_st_wait();           // the zeros written behind the loads have landed
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTA2) mbar_arrive_cluster(acc_empty(buf), 0); else mbar_arrive(acc_empty(buf));
        }
      };
      const bool live = tc.t < e.T;            // false only for the padding tile of an odd pair
      const bool work = live && !(e.desc_variant & 4);
      uint32_t va[32], vb[32];
      tmem_ld32(tacc + (u0 / G) * NTILE + (u0 % G) * 32, va);
      // two units per iteration (register double buffering); NOT unrolled further: the epilogue
      // body is large and the three warp roles already compete for the instruction cache
#pragma unroll 1
      for (int k = 0; k < kMine; k += 2) {
        tmem_ld_wait();
        if (stacked) { const int u = u0 + k; tmem_st32_zero(tacc + (u / G) * NTILE + (u % G) * 32); }
        if (k + 1 < kMine) {
          const int u = u0 + k + 1;
          tmem_ld32(tacc + (u / G) * NTILE + (u % G) * 32, vb);
          skip_prefetch<MASK>(e, tc, el, tc.y0 + u / G, nb0 + (u % G) * 32, quad, lane, skb);
        } else {
          release_acc();
        }
        if (work) {
          const int u = u0 + k;
          float bv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) bv[i] = p.bias_c[nb0 + (u % G) * 32 + i];
          epilogue_unit<BF16, MASK>(e, tc, el, tc.y0 + u / G, nb0 + (u % G) * 32, va, ska, bv, next_stg(), quad, lane,
                                    rin, k == 0 && rin_ok, &map_o);
        }
        if (k + 1 < kMine) {
          tmem_ld_wait();
          if (stacked) { const int u = u0 + k + 1; tmem_st32_zero(tacc + (u / G) * NTILE + (u % G) * 32); }
          if (k + 2 < kMine) {
            const int u = u0 + k + 2;
            tmem_ld32(tacc + (u / G) * NTILE + (u % G) * 32, va);
            skip_prefetch<MASK>(e, tc, el, tc.y0 + u / G, nb0 + (u % G) * 32, quad, lane, ska);
          } else {
            release_acc();
          }
          if (work) {
            const int u = u0 + k + 1;
            float bv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) bv[i] = p.bias_c[nb0 + (u % G) * 32 + i];
            epilogue_unit<BF16, MASK>(e, tc, el, tc.y0 + u / G, nb0 + (u % G) * 32, vb, skb, bv, next_stg(), quad, lane,
                                      rin, false, &map_o);
          }
        }
      }
    }
    if constexpr ((MASK & EPI_TMA_OUT) != 0) {
      if (lane == 0) bulk_wait_group_all();    // staging must outlive the last TMA reads
      __syncwarp();
    }
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace bsvd
