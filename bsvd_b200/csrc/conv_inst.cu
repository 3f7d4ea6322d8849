// conv_inst.cu — explicit instantiation of the conv3x3_tc_kernel family for one (NTILE, R) pair.
// Compiled once per pair (-DBSVD_INST=0..3, see __graft_entry__.build) so the four groups of
// kernel instances build in parallel; bsvd_capi.cu declares them `extern template`.
#include "stage_launch.cuh"

namespace bsvd {
#if BSVD_INST == 0
template int launch_one<64, 2>(const StageLaunch&, cudaStream_t);
#elif BSVD_INST == 1
template int launch_one<128, 2>(const StageLaunch&, cudaStream_t);
#elif BSVD_INST == 2
template int launch_one<256, 1>(const StageLaunch&, cudaStream_t);
#elif BSVD_INST == 3
template int launch_one<256, 2>(const StageLaunch&, cudaStream_t);
#else
#error "BSVD_INST must be 0..3"
#endif
}  // namespace bsvd
