// Sub-query scan kernels (QP == 128): 128 single-row queries or 4 x 32-row queries share every document tile.
#include "scan_launch_impl.cuh"

namespace vrag {
cudaError_t scan_launch_multi(int QS, bool packed, const ScanLaunch& L) {
  if (QS == kMultiQs8x32) return packed ? cudaErrorInvalidValue : scan_launch_t<128, false, false, 32, 2>(L);   // 8 x 32 columns, plain fp16
  if (QS == 1) return packed ? scan_launch_t<128, true, false, 1>(L) : scan_launch_t<128, false, false, 1>(L);
  if (QS == 32) return packed ? scan_launch_t<128, true, false, 32>(L) : scan_launch_t<128, false, false, 32>(L);
  return cudaErrorInvalidValue;
}
}  // namespace vrag
