// Operand-switching scan kernels: every query of a batch scores its own candidate list in one launch.
#include "scan_launch_impl.cuh"

namespace vrag {
cudaError_t scan_launch_bsw(int QP, bool packed, const ScanLaunch& L) {
  if (QP == 32) return packed ? scan_launch_t<32, true, true>(L) : scan_launch_t<32, false, true>(L);
  if (QP == 64) return packed ? scan_launch_t<64, true, true>(L) : scan_launch_t<64, false, true>(L);
  return cudaErrorInvalidValue;
}
}  // namespace vrag
