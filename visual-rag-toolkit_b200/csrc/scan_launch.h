// Launch interface of the MaxSim scan kernels. The kernel templates (maxsim_scan.cuh) are instantiated in
// scan_kernels_{single,bsw,multi}.cu so that the host side of the library compiles without them.
#pragma once
#include "scan_params.h"

namespace vrag {

struct ScanLaunch {
  const CUtensorMap* tm_rows;     // 128-row boxes (or the padded 3-D view when p.pad_rows > 0)
  const RowMaps* tm_small;        // boxes of 4, 8, ..., 32 rows
  const CUtensorMap* tm_scale128;
  const CUtensorMap* tm_scale32;
  ScanParams p;
  long long n_units;              // work units (LARGE: items, PACKED: tiles), all groups
  int num_sms;
  cudaStream_t stream;
};

// Variant = (QP, QS, PACKED, BSW). Returns cudaSuccess, a CUDA error, or cudaErrorInvalidValue for a variant that is
// not built. Sets the opt-in shared-memory attribute once per device and variant.
cudaError_t scan_launch_single(int QP, bool packed, const ScanLaunch& L);          // QS == QP, one query
cudaError_t scan_launch_bsw(int QP, bool packed, const ScanLaunch& L);             // operand switching (QP 32 / 64)
cudaError_t scan_launch_multi(int QS, bool packed, const ScanLaunch& L);           // QP == 128, QS in {1, 32}; kMultiQs8x32 (LARGE):
constexpr int kMultiQs8x32 = 3208;                                                 //   eight plain-fp16 queries per image (first pass)

}  // namespace vrag
