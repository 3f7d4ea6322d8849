// peer.cu — one-sided exchange between the one-process-per-GPU ranks of a node over NVLink / NVSwitch.
//
// Where the reference runs multi-GPU it wraps the network in DataParallel (base_model.py:74-75); here
// independent clips shard one per GPU and the only traffic is (a) gathering the denoised clips and
// (b) for 4K frames cut into spatial tiles, the input strips a tile needs from its neighbours
// (SURVEY §8e).  Both are bulk copies with no arithmetic, so they run on the COPY ENGINES: every rank
// owns a symmetric device buffer that its peers map through CUDA IPC, a transfer is a
// cudaMemcpyAsync / cudaMemcpy2DAsync between peer-mapped pointers on a side stream, and completion is a
// 4-byte flag written behind the data on the same stream; the consumer's stream waits for the flag with
// cuStreamWaitValue32.  No SM is used, so the persistent conv kernels (1 CTA per SM, all 148 SMs) never
// share the machine with a collective's CTAs, and a transfer overlaps the next clip's compute entirely.
// NCCL stays the control plane (rendezvous, barriers, the timing reduction) in bsvd_b200/peer.py.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/bsvd_b200.h"
#include "common.cuh"   // fail(), CUDA_TRY

namespace bsvd {

typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamWaitValue32Fn get_wait_fn() {
  static StreamWaitValue32Fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StreamWaitValue32Fn>(p);
  });
  return fn;
}

// A signal is a 4-byte copy out of a device pool of values: seq[i] = seq_base + i (4 MB).  The pool covers a
// window of 2^20 consecutive values; a value outside it (once per ~10^6 steps of a long-running service) drains
// the device and rewrites the pool around the new value, so step counters are good for the full 32 bits.
constexpr int kSeqLen = 1 << 20;
constexpr unsigned kSeqSlack = 4096;   // values a little below the newest one stay in the window (late releases)
static __global__ void fill_seq_kernel(uint32_t* seq, int n, uint32_t base) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) seq[i] = base + (uint32_t)i;
}

}  // namespace bsvd

using namespace bsvd;

// Layout of every rank's symmetric allocation: [ flags: nflags x uint32 (padded to 4 KB) | data: bytes ].
struct bsvd_peer_group {
  int rank = 0, world = 1, device = 0;
  size_t bytes = 0, flag_bytes = 0;
  int nflags = 0;
  uint8_t* local = nullptr;                 // base of the local allocation (flags first)
  std::vector<uint8_t*> base;               // base[r]: rank r's allocation as mapped here (base[rank] == local)
  uint32_t* seq = nullptr;                  // device pool of signal values seq_base .. seq_base + kSeqLen - 1
  uint32_t seq_base = 0;
  bool opened = false;
};

extern "C" {

int bsvd_peer_create(int rank, int world, size_t bytes, int nflags, bsvd_peer_group** out) {
  if (!out || world < 1 || rank < 0 || rank >= world || nflags < 0) return fail("bad peer group arguments");
  bsvd_peer_group* g = new bsvd_peer_group();
  g->rank = rank; g->world = world; g->bytes = bytes; g->nflags = nflags;
  g->flag_bytes = ((size_t)nflags * 4 + 4095) / 4096 * 4096;
  if (cudaGetDevice(&g->device) != cudaSuccess) { delete g; return fail("no CUDA device"); }
  // plain cudaMalloc: the legacy allocator's blocks are what cudaIpcGetMemHandle can export
  if (cudaMalloc((void**)&g->local, g->flag_bytes + bytes) != cudaSuccess ||
      cudaMemset(g->local, 0, g->flag_bytes) != cudaSuccess ||
      cudaMalloc((void**)&g->seq, sizeof(uint32_t) * kSeqLen) != cudaSuccess) {
    cudaGetLastError();
    if (g->local) cudaFree(g->local);
    if (g->seq) cudaFree(g->seq);
    delete g;
    return fail("cudaMalloc of the symmetric peer buffer (%zu bytes) failed", bytes);
  }
  fill_seq_kernel<<<64, 256>>>(g->seq, kSeqLen, 0u);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    cudaGetLastError();
    cudaFree(g->local); cudaFree(g->seq);
    delete g;
    return fail("peer group initialisation failed");
  }
  g->base.assign(world, nullptr);
  g->base[rank] = g->local;
  if (world == 1) g->opened = true;
  *out = g;
  return 0;
}

int bsvd_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int bsvd_peer_get_handle(bsvd_peer_group* g, void* handle_out) {
  if (!g || !handle_out) return fail("null argument");
  cudaIpcMemHandle_t hdl;
  CUDA_TRY(cudaIpcGetMemHandle(&hdl, g->local));
  memcpy(handle_out, &hdl, sizeof(hdl));
  return 0;
}

int bsvd_peer_open(bsvd_peer_group* g, const void* handles) {
  if (!g || !handles) return fail("null argument");
  if (g->opened) return 0;
  const uint8_t* hb = reinterpret_cast<const uint8_t*>(handles);
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank) continue;
    cudaIpcMemHandle_t hdl;
    memcpy(&hdl, hb + (size_t)r * sizeof(hdl), sizeof(hdl));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return fail("cudaIpcOpenMemHandle for rank %d failed: %s (peers must be GPUs of one node with P2P "
                  "access, one process per GPU)", r, cudaGetErrorString(e));
    g->base[r] = reinterpret_cast<uint8_t*>(p);
  }
  g->opened = true;
  return 0;
}

void* bsvd_peer_local_data(bsvd_peer_group* g) { return g ? g->local + g->flag_bytes : nullptr; }

static int check_range(const bsvd_peer_group* g, int dst_rank, size_t off, size_t span) {
  if (!g->opened) return fail("peer group is not opened yet (bsvd_peer_open)");
  if (dst_rank < 0 || dst_rank >= g->world) return fail("bad destination rank %d", dst_rank);
  if (off > g->bytes || span > g->bytes - off) return fail("peer put out of range (offset %zu + %zu > %zu)", off, span, g->bytes);
  return 0;
}

int bsvd_peer_put(bsvd_peer_group* g, int dst_rank, size_t dst_off, const void* src, size_t bytes, void* stream) {
  if (!g || !src) return fail("null argument");
  if (check_range(g, dst_rank, dst_off, bytes)) return 1;
  CUDA_TRY(cudaMemcpyAsync(g->base[dst_rank] + g->flag_bytes + dst_off, src, bytes, cudaMemcpyDeviceToDevice,
                           reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int bsvd_peer_put2d(bsvd_peer_group* g, int dst_rank, size_t dst_off, size_t dst_pitch, const void* src,
                    size_t src_pitch, size_t width_bytes, size_t rows, void* stream) {
  if (!g || !src) return fail("null argument");
  if (rows == 0 || width_bytes == 0) return 0;
  if (width_bytes > dst_pitch || width_bytes > src_pitch) return fail("2-D put: width exceeds a pitch");
  if (check_range(g, dst_rank, dst_off, (rows - 1) * dst_pitch + width_bytes)) return 1;
  CUDA_TRY(cudaMemcpy2DAsync(g->base[dst_rank] + g->flag_bytes + dst_off, dst_pitch, src, src_pitch, width_bytes,
                             rows, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

// planes x rows x width_bytes block between two pitched volumes (src: local device memory; dst: data area
// of dst_rank): one copy-engine transfer for e.g. the [T*C planes][rows][cols] strip of an NCHW clip.
int bsvd_peer_put3d(bsvd_peer_group* g, int dst_rank, size_t dst_off, size_t dst_pitch, size_t dst_plane_rows,
                    const void* src, size_t src_pitch, size_t src_plane_rows, size_t width_bytes, size_t rows,
                    size_t planes, void* stream) {
  if (!g || !src) return fail("null argument");
  if (rows == 0 || width_bytes == 0 || planes == 0) return 0;
  if (width_bytes > dst_pitch || width_bytes > src_pitch || rows > dst_plane_rows || rows > src_plane_rows)
    return fail("3-D put: block exceeds a pitch / plane height");
  if (check_range(g, dst_rank, dst_off, ((planes - 1) * dst_plane_rows + rows - 1) * dst_pitch + width_bytes)) return 1;
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src), src_pitch, src_pitch, src_plane_rows);
  p.dstPtr = make_cudaPitchedPtr(g->base[dst_rank] + g->flag_bytes + dst_off, dst_pitch, dst_pitch, dst_plane_rows);
  p.extent = make_cudaExtent(width_bytes, rows, planes);
  p.kind = cudaMemcpyDeviceToDevice;
  CUDA_TRY(cudaMemcpy3DAsync(&p, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int bsvd_peer_signal(bsvd_peer_group* g, int dst_rank, int flag, unsigned value, void* stream) {
  if (!g) return fail("null argument");
  if (!g->opened) return fail("peer group is not opened yet (bsvd_peer_open)");
  if (dst_rank < 0 || dst_rank >= g->world || flag < 0 || flag >= g->nflags) return fail("bad flag %d on rank %d", flag, dst_rank);
  if (value < g->seq_base || value - g->seq_base >= (unsigned)kSeqLen) {
    // every pending signal copy (on any stream) still reads the pool: drain the device before rewriting it
    CUDA_TRY(cudaDeviceSynchronize());
    g->seq_base = value > kSeqSlack ? value - kSeqSlack : 0u;
    fill_seq_kernel<<<64, 256>>>(g->seq, kSeqLen, g->seq_base);
    CUDA_TRY(cudaDeviceSynchronize());
  }
  // a 4-byte copy behind the data copies of the same stream: ordered after them, still no SM involved
  CUDA_TRY(cudaMemcpyAsync(g->base[dst_rank] + (size_t)flag * 4, g->seq + (value - g->seq_base), 4, cudaMemcpyDeviceToDevice,
                           reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int bsvd_peer_wait(bsvd_peer_group* g, int flag, unsigned value, void* stream) {
  if (!g) return fail("null argument");
  if (flag < 0 || flag >= g->nflags) return fail("bad flag %d", flag);
  StreamWaitValue32Fn fn = get_wait_fn();
  if (!fn) return fail("cuStreamWaitValue32 entry point not available");
  CUresult r = fn(reinterpret_cast<CUstream>(stream), (CUdeviceptr)(uintptr_t)(g->local + (size_t)flag * 4), value,
                  CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) return fail("cuStreamWaitValue32 failed: %d", (int)r);
  return 0;
}

int bsvd_peer_read_flag(bsvd_peer_group* g, int flag, unsigned* value) {
  if (!g || !value || flag < 0 || flag >= g->nflags) return fail("bad argument");
  CUDA_TRY(cudaMemcpy(value, g->local + (size_t)flag * 4, 4, cudaMemcpyDeviceToHost));
  return 0;
}

int bsvd_peer_destroy(bsvd_peer_group* g) {
  if (!g) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < g->world; ++r)
    if (r != g->rank && g->base[r]) cudaIpcCloseMemHandle(g->base[r]);
  if (g->local) cudaFree(g->local);
  if (g->seq) cudaFree(g->seq);
  delete g;
  return 0;
}

}  // extern "C"
