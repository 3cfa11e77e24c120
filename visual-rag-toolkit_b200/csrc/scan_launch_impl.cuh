// Shared by the scan_kernels_*.cu translation units.
#pragma once
#include <algorithm>

#include "maxsim_scan.cuh"
#include "scan_launch.h"

namespace vrag {

struct PerDeviceOnceK {
  bool done[64] = {false};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

template <int QP, bool PACKED, bool BSW = false, int QS = QP, int NBLK = 1>
static cudaError_t scan_launch_t(const ScanLaunch& L) {
  auto kern = maxsim_scan_kernel<QP, QS, PACKED, BSW, NBLK>;
  const size_t smem = ScanCfg<QP>::smem_bytes(PACKED, BSW, QS < QP);
  static PerDeviceOnceK once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  const unsigned grid = static_cast<unsigned>(std::min<long long>(L.num_sms, L.n_units));
  kern<<<grid, ScanCfg<QP>::threads(PACKED, QS < QP, BSW), smem, L.stream>>>(*L.tm_rows, *L.tm_small, *L.tm_scale128,
                                                                       *L.tm_scale32, L.p);
  return cudaGetLastError();
}

}  // namespace vrag
