// Single-query scan kernels (QS == QP): exhaustive scans, stage-1 pooled scans, candidate gathers.
#include "scan_launch_impl.cuh"

namespace vrag {
cudaError_t scan_launch_single(int QP, bool packed, const ScanLaunch& L) {
#define VRAG_CASE(QPV) \
  case QPV:            \
    return packed ? scan_launch_t<QPV, true>(L) : scan_launch_t<QPV, false>(L);
  switch (QP) {
    VRAG_CASE(8)
    VRAG_CASE(16)
    case 24:
      return packed ? cudaErrorInvalidValue : scan_launch_t<24, false>(L);
    VRAG_CASE(32)
    VRAG_CASE(64)
    VRAG_CASE(128)
    default:
      return cudaErrorInvalidValue;
  }
#undef VRAG_CASE
}
}  // namespace vrag
