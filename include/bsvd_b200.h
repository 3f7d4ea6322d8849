/*
 * bsvd_b200.h — C ABI of the B200-native BSVD-64 inference path.
 *
 * This is the drop-in boundary for ONE hot path of ChenyangQiQi/BSVD:
 *   Experimental_root/archs/bsvd_arch.py::BSVD.forward          (bsvd_arch.py:490-499)
 *   -> streaming_forward / feedin_one_element / DenBlock.forward (bsvd_arch.py:501-552, 485-488, 374-396)
 * Plain pointers and sizes only; no torch types.  The reference-side binding is the
 * ctypes stub in bsvd_b200/capi.py; the nn.Module that registers under
 * ARCH_REGISTRY['BSVD'] (bsvd_b200/arch.py) calls nothing but these entry points.
 *
 * All device pointers are plain CUDA device pointers in the current context.  `stream` is a
 * cudaStream_t passed as void* (0 = legacy default stream).  Every call returns 0 on success,
 * non-zero on failure; bsvd_last_error() returns a description of the last failure of the
 * calling thread.  Nothing here ever falls back to a CPU path.
 *
 * Threading and devices: a handle is bound to the CUDA device it was created on (weights, workspaces,
 * tensor maps); every entry point fails if another device is current or a device pointer lives
 * elsewhere.  A handle caches its launch plan and workspaces, so calls on ONE handle must not run
 * concurrently from several host threads (use one handle per thread); different handles are independent.
 */
#ifndef BSVD_B200_H_
#define BSVD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bsvd_handle bsvd_handle;

/* Operand precision of the tensor-core contractions (accumulation is always fp32,
 * bias/activation/residual arithmetic is fp32, activations are stored in this type). */
enum {
  BSVD_PREC_FP16 = 0,
  BSVD_PREC_BF16 = 1,
  /* fp32-grade: for callers that run the reference with val.fp16 False and TF32 off
   * (Experimental_root/models/denoising_model.py:204).  Every activation and weight is carried as a pair of
   * fp16 numbers (x = hi + lo) and every contraction is three tensor-core products
   * (W_hi x_hi + W_hi x_lo + W_lo x_hi, fp32 accumulation), the first / last conv run in fp32 on the CUDA
   * cores: ~1e-5 from the fp32 reference at about a third of the fp16 mode's speed.  BSVD-64 only. */
  BSVD_PREC_FP32X3 = 2
};

/* Mirrors the constructor kwargs of BSVD (bsvd_arch.py:446-447) as passed by
 * options/test/bsvd_c64.yml:85-93: chns={64,128,256}, mid_ch=64, interm_ch=64, in_ch=4 (3 = blind),
 * out_ch=3, norm='none', act='relu6'.  Also accepted: the c32 configurations
 * (options/train/0402_*_blind_c32.yml:63-68: chns={32,64,128}, mid_ch=32, interm_ch=30, act='relu'),
 * which run on the same kernels with channel counts below 64 zero-padded to 64.
 * Anything else makes bsvd_create fail (there is no CPU fallback). */
typedef struct bsvd_config {
  int chns[3];
  int mid_ch;
  int interm_ch;
  int in_ch;
  int out_ch;
  int act_relu6;   /* 1 = relu6 (bsvd_c64.yml), 0 = relu (the c32 configurations) */
  int norm_none;   /* 1 = norm 'none' (only supported value) */
  int precision;   /* BSVD_PREC_* */
  int device;      /* CUDA device ordinal, -1 = current */
} bsvd_config;

#define BSVD_NUM_LAYERS 32 /* 16 convs per DenBlock x 2 DenBlocks (bsvd_arch.py:325-396) */

/* -- lifetime ---------------------------------------------------------------------------- */
/* replaces BSVD.__init__ (bsvd_arch.py:446-456) */
int bsvd_create(const bsvd_config* cfg, bsvd_handle** out);
/* replaces nn.Module teardown */
int bsvd_destroy(bsvd_handle* h);
const char* bsvd_last_error(void);
/* library / build information, e.g. "bsvd_b200 0.1 sm_100a" */
const char* bsvd_version(void);

/* -- weights -------------------------------------------------------------------------------
 * replaces BSVD.load / DenBlock.load_from (bsvd_arch.py:462-474, 349-355).
 * layer = block*16 + l, block 0 = temp1, 1 = temp2, l in reference execution order:
 *   0 inc.convblock.0      1 inc.convblock.3        2 downc0.convblock.0 (stride 2)
 *   3 downc0.memconv.c1    4 downc0.memconv.c2      5 downc1.convblock.0 (stride 2)
 *   6 downc1.memconv.c1    7 downc1.memconv.c2      8 upc2.memconv.c1
 *   9 upc2.memconv.c2     10 upc2.convblock.0 (+PixelShuffle)
 *  11 upc1.memconv.c1     12 upc1.memconv.c2       13 upc1.convblock.0 (+PixelShuffle)
 *  14 outc.convblock.0    15 outc.convblock.3
 * w is the nn.Conv2d weight, HOST pointer, fp32, OIHW contiguous; bias HOST fp32 [O].
 * The call repacks into the tensor-core layout and uploads (synchronous; when the layer already held
 * weights the device is drained first, so a forward still in flight on any stream never sees a mix). */
int bsvd_set_weights(bsvd_handle* h, int layer, const float* w_oihw, const float* bias,
                     int out_ch, int in_ch);
/* expected (out_ch, in_ch) of a layer; returns non-zero for a bad index */
int bsvd_layer_shape(const bsvd_handle* h, int layer, int* out_ch, int* in_ch, int* stride);

/* -- clip mode: BSVD.forward on one stream of T frames (bsvd_arch.py:490-552) ---------------
 * in:        device fp32 [T, in_c, H, W] contiguous, in_c = 4, or in_c = 3 with noise_map
 * noise_map: device fp32 [T, 1, H, W] or NULL   (the torch.cat of bsvd_arch.py:492-493 is folded in)
 * out:       device fp32 [T, 3, H, W]
 * H and W must be multiples of 4 (the reference raises at the skip add otherwise).
 * Work is enqueued on `stream`; no host synchronisation. Equivalent to feeding the T frames
 * through streaming_forward (same zero folds at both clip ends). */
int bsvd_forward_clip(bsvd_handle* h, const float* in, const float* noise_map, float* out,
                      int T, int in_c, int H, int W, void* stream);

/* N independent clips of T frames each in ONE pass of the 32 stages (in / noise_map / out hold N*T frames,
 * clip after clip): the temporal folds never cross a clip boundary.  This is the batch-of-clips
 * configuration on one GPU and the arithmetic of the training twin's train-mode shift
 * (TemporalShift.forward -> shift(x, n_segment), Experimental_root/archs/temporal_shift_ops/
 * temporal_shift.py:27-49, with n_segment = T); bsvd_forward_clip is the N = 1 case. */
int bsvd_forward_clips(bsvd_handle* h, const float* in, const float* noise_map, float* out, int N, int T,
                       int in_c, int H, int W, void* stream);

/* -- the callers either side of the path, fused -----------------------------------
 * replaces temp_denoise (Experimental_root/models/validation_seq_infer.py:10-31) together with
 * DenoisingModel.padding_input / crop_output (Experimental_root/models/denoising_model.py:133-168):
 *   in:    device fp32 [T, 3, H, W] noisy frames in [0,1], ANY H, W >= 2
 *   sigma: noise standard deviation in [0,1]; the constant noise-map channel the reference builds
 *          with torch.ones(...) * sigma is synthesised inside the first kernel (pass < 0 for a
 *          blind model)
 *   out:   device fp32 [T, 3, H, W], clamped to [0,1]
 * H and W are reflect-padded (bottom/right) to multiples of 4 inside the first kernel's loads and
 * cropped by the last kernel's stores: no padded copy, no concat, no clamp pass, no crop copy. */
int bsvd_denoise_clip(bsvd_handle* h, const float* in, float sigma, float* out, int T, int H, int W,
                      void* stream);

/* -- frame I/O ----------------------------------------------------------------------
 * The same call on decoded frames: in/out are device uint8 [T, H, W, 3] (HWC, bgr != 0: channel
 * order B,G,R as cv2.imread delivers it).  The first kernel normalises with /255 while it builds
 * its patches (img2tensor, BasicSR/basicsr/utils/img_util.py; ValFolderDataset,
 * Experimental_root/data/video_dali_dataset.py:199-249), the last kernel stores
 * round(clamp(x, 0, 1) * 255) (tensor2img, DenoisingModel save path denoising_model.py:276-310):
 * bit-identical to uint8 -> float -> bsvd_denoise_clip -> uint8 done with separate passes. */
int bsvd_denoise_clip_u8(bsvd_handle* h, const uint8_t* in, float sigma, uint8_t* out, int T, int H,
                         int W, int bgr, void* stream);

/* -- on-device PSNR ----------------------------------------------------------------
 * replaces calculate_psnr_float (BasicSR/basicsr/metrics/psnr_ssim.py:130-168) applied per frame:
 * a, b: device fp32 [T, C, H, W] in [0,1]; crop_border pixels are ignored on every edge;
 * psnr: device fp32 [T] <- -10 log10(mean((a-b)^2)) (+inf when identical).  Enqueued on `stream`. */
int bsvd_psnr(const float* a, const float* b, int T, int C, int H, int W, int crop_border,
              float* psnr, void* stream);

/* -- on-device SSIM ----------------------------------------------------------------
 * replaces calculate_ssim (BasicSR/basicsr/metrics/psnr_ssim.py:49-128) applied per frame: 11x11 Gaussian
 * window (sigma 1.5) at every position where it fits inside the crop_border-cropped image, the five local
 * moments and the SSIM map in double precision, mean over positions and channels.
 * a, b: device fp32 [T, C, H, W]; data_range = L of c1 = (0.01 L)^2, c2 = (0.03 L)^2: 1 for [0,1] floats,
 * 255 for [0,255] values (the reference's uint8 images); ssim: device fp32 [T].  Enqueued on `stream`. */
int bsvd_ssim(const float* a, const float* b, int T, int C, int H, int W, int crop_border, float data_range,
              float* ssim, void* stream);

/* Same call with HOST buffers (pinned or pageable): H2D copy, forward, D2H copy, then waits for
 * completion.  This is the end-to-end entry the bench's `e2e` number goes through. */
int bsvd_forward_clip_host(bsvd_handle* h, const float* in_host, const float* noise_map_host,
                           float* out_host, int T, int in_c, int H, int W, void* stream);

/* Pipelined form of the host entry: enqueues H2D (own copy stream), forward (caller's stream) and
 * D2H (second copy stream) and returns; consecutive calls overlap copy-in of clip i+1, compute of
 * clip i and copy-out of clip i-1 (double-buffered device staging).  in_host/out_host must stay
 * valid (and should be pinned) until bsvd_host_sync() returns. */
int bsvd_forward_clip_host_async(bsvd_handle* h, const float* in_host, const float* noise_map_host,
                                 float* out_host, int T, int in_c, int H, int W, void* stream);
int bsvd_host_sync(bsvd_handle* h);
/* Device staging buffer ([T,3,H,W] fp32) holding the result of the most recent
 * bsvd_forward_clip_host_async call, valid for consumers ordered behind that call on its stream and
 * until the call after next reuses it (multi-GPU: the gather reads it from here, see bsvd_peer_*). */
int bsvd_host_last_output(bsvd_handle* h, float** dev_out);

/* -- streaming mode: BSVD.feedin_one_element / reset (bsvd_arch.py:485-488, 459-461) -------
 * frame:     device fp32 [in_c, H, W] or NULL (NULL = the reference's feedin_one_element(None))
 * noise_map: device fp32 [1, H, W] or NULL
 * out:       device fp32 [3, H, W]; written iff *produced is set to 1
 * The first 16 pushes of a stream produce nothing (count_shift, bsvd_arch.py:554-560). */
int bsvd_stream_push(bsvd_handle* h, const float* frame, const float* noise_map, float* out,
                     int in_c, int H, int W, int* produced, void* stream);
int bsvd_reset(bsvd_handle* h);
/* Steady-state pushes (from the 18th frame of a stream on) replay one CUDA graph per ring phase instead of
 * issuing 32 launches; this counts the replays so far (BSVD_B200_NO_STREAM_GRAPH=1 disables them). */
long long bsvd_stream_graph_replays(const bsvd_handle* h);

/* -- multi-GPU plumbing: one-sided exchange over NVLink through peer-mapped memory ---------
 * replaces the reference's DataParallel scatter/gather (BasicSR/basicsr/models/base_model.py:74-75) for
 * the two ways this path shards: independent clips, one per GPU (gathering the denoised clips), and 4K
 * frames cut into spatial tiles (the input strips a tile needs from its neighbours).
 * One process per GPU on one node.  Every rank owns a symmetric buffer of `bytes` data bytes and `nflags`
 * 32-bit flags; peers map it through CUDA IPC (exchange the handles with any out-of-band channel, e.g.
 * torch.distributed).  Transfers are copy-engine copies enqueued on `stream` — no SM is used, so they
 * overlap the compute kernels entirely; a signal is a flag write ordered behind the puts of the same
 * stream, a wait makes `stream` block until the LOCAL flag is >= value.  Signal values must increase. */
typedef struct bsvd_peer_group bsvd_peer_group;
int bsvd_peer_create(int rank, int world, size_t bytes, int nflags, bsvd_peer_group** out);
int bsvd_peer_handle_bytes(void);                               /* size of one exported handle */
int bsvd_peer_get_handle(bsvd_peer_group* g, void* handle_out); /* this rank's handle */
int bsvd_peer_open(bsvd_peer_group* g, const void* handles);    /* world x handle_bytes, rank order */
void* bsvd_peer_local_data(bsvd_peer_group* g);                 /* device pointer of the local data area */
int bsvd_peer_put(bsvd_peer_group* g, int dst_rank, size_t dst_off, const void* src, size_t bytes, void* stream);
int bsvd_peer_put2d(bsvd_peer_group* g, int dst_rank, size_t dst_off, size_t dst_pitch, const void* src,
                    size_t src_pitch, size_t width_bytes, size_t rows, void* stream);
/* planes x rows x width_bytes block between two pitched volumes (pitch in bytes, plane height in rows):
 * e.g. the strip [T*C][rows][cols] of an NCHW clip into a neighbour's enlarged tile, one transfer */
int bsvd_peer_put3d(bsvd_peer_group* g, int dst_rank, size_t dst_off, size_t dst_pitch, size_t dst_plane_rows,
                    const void* src, size_t src_pitch, size_t src_plane_rows, size_t width_bytes, size_t rows,
                    size_t planes, void* stream);
int bsvd_peer_signal(bsvd_peer_group* g, int dst_rank, int flag, unsigned value, void* stream);
int bsvd_peer_wait(bsvd_peer_group* g, int flag, unsigned value, void* stream);
int bsvd_peer_read_flag(bsvd_peer_group* g, int flag, unsigned* value);   /* synchronous, for tests */
int bsvd_peer_destroy(bsvd_peer_group* g);

/* -- introspection used by bench.py / tests -------------------------------------------------- */
/* number of kernel launches enqueued by the last forward/push on this handle */
int bsvd_last_launch_count(const bsvd_handle* h);
/* fp16 range guard.  Activations are stored in 16 bits; with BSVD_PREC_FP16 a stage whose output is not
 * clamped by ReLU6 (the PixelShuffle + skip convs, temp1's output, every stage of an act='relu' model)
 * could exceed 65504 where the reference's fp32 tensors (bsvd_arch.py:402-414) would not.  Such a store
 * sets a sticky device flag; this call waits for the device, returns it through *flag (0 / 1) and
 * clears it when reset != 0.  bf16 has fp32's range and never sets it. */
int bsvd_overflow_flag(bsvd_handle* h, int* flag, int reset);
/* bytes of device workspace currently held */
size_t bsvd_workspace_bytes(const bsvd_handle* h);

/* Per-stage device timing (CUDA events on the caller's stream, recorded between the stage launches
 * of bsvd_forward_clip).  on=1 starts a fresh accumulation.  bsvd_get_stage_ms synchronises the
 * recorded events and writes, for stage 0 (input staging) and stages 1..32 (the 32 conv stages),
 * the device milliseconds summed over the forwards recorded since profiling was switched on;
 * returns the number of forwards accumulated through *passes. */
#define BSVD_NUM_STAGES 33
int bsvd_set_profiling(bsvd_handle* h, int on);
int bsvd_get_stage_ms(bsvd_handle* h, float* ms, int n, int* passes);
/* Static description of stage i (1..32) for roofline arithmetic: kernel instance name,
 * algorithmic FLOPs and activation bytes per frame-pixel-independent unit are computed by the
 * caller from (cin, cout, stride, ntile, rows). */
int bsvd_stage_info(const bsvd_handle* h, int stage, int* cin, int* cout, int* stride, int* ntile,
                    int* rows);

/* -- single fused conv stage (test / micro-benchmark hook) ------------------------------------
 * Runs ONE fused 3x3 conv stage exactly as the network does (same kernels), on NHWC
 * 16-bit activations.  Used by tests/ to check every kernel variant against the oracle.
 * Flags describe the prologue/epilogue that the reference expresses as separate ops. */
enum {
  BSVD_EPI_RELU6 = 1,        /* nn.ReLU6                               (bsvd_arch.py:185-192)  */
  BSVD_EPI_PIXSHUF = 2,      /* nn.PixelShuffle(2)                     (bsvd_arch.py:266)      */
  BSVD_EPI_SKIP_ADD = 4,     /* DenBlock.none_add                      (bsvd_arch.py:402-406)  */
  BSVD_EPI_SHIFT_STORE = 8,  /* BiBufferConv/ShiftConv channel folds   (bsvd_arch.py:42-50)    */
  BSVD_EPI_STRIDE2 = 16      /* DownBlock conv stride 2                (bsvd_arch.py:238-239)  */
};
typedef struct bsvd_conv_desc {
  int T, H, W;        /* INPUT frames / rows / cols */
  int cin, cout;      /* conv channels (cout counts conv outputs, i.e. before PixelShuffle) */
  int flags;          /* BSVD_EPI_* */
  int precision;      /* BSVD_PREC_* */
  int debug_variant;  /* 0 = production; other values select descriptor experiments */
} bsvd_conv_desc;
/* in:  device 16-bit [T,H,W,cin]; w/bias: HOST fp32 OIHW / [cout];
 * skip: device 16-bit, shape of the output, or NULL; out: device 16-bit NHWC:
 *   [T,H,W,cout] | stride2: [T,H/2,W/2,cout] | pixshuf: [T,2H,2W,cout/4]
 * With SHIFT_STORE the output holds the *shifted* tensor a following BiBufferConv reads:
 *   out[t][..., 0:f] = y[t+1][..., 0:f], out[t][..., f:2f] = y[t-1][..., f:2f] (zeros off-clip),
 *   f = channels/8. */
int bsvd_conv_stage(const bsvd_conv_desc* d, const void* in, const float* w_oihw,
                    const float* bias, const void* skip, void* out, void* stream);
/* device milliseconds of the (re-)launches timed inside the last bsvd_conv_stage call */
float bsvd_last_stage_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* BSVD_B200_H_ */
