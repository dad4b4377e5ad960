// halo.cuh -- ghost-cell exchange across block connections (K12).
// Replaces the reference's slice -> MPI_Pack -> MPI_Sendrecv_replace -> InsertSlice path
// (ref: include/multiArray3d.hpp:790-926,1440-1550; src/boundaryConditions.cpp:1016-1150,
// 3006-3181; src/utility.cpp:400-423). Same-GPU connections are one gather/scatter kernel;
// cross-GPU connections pack on device and travel with ncclSend/ncclRecv.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/aither_gpu.h"
#include "layout.cuh"

namespace aither {

enum HaloField { kHaloState = 0, kHaloUpdate = 1 };

struct HaloPlan {
  int nConn = 0;
};

inline std::string &HaloErrorRef() {
  static thread_local std::string e;
  return e;
}
inline std::string HaloError() { return HaloErrorRef(); }

inline int HaloBuild(HaloPlan &plan, const std::vector<aither_conn> &conns,
                     const std::vector<const BlockDev *> &devs, const std::vector<int> &globalPos,
                     int neq, int g, int rank, int nRanks, void *ncclComm) {
  (void)devs; (void)globalPos; (void)neq; (void)g; (void)rank; (void)nRanks; (void)ncclComm;
  plan.nConn = static_cast<int>(conns.size());
  if (!conns.empty()) {
    HaloErrorRef() = "block connections (interblock/periodic) are not built yet";
    return 1;
  }
  return 0;
}
inline int HaloExchange(HaloPlan &plan, const std::vector<const BlockDev *> &devs, int which,
                        cudaStream_t stream, long long *launches) {
  (void)plan; (void)devs; (void)which; (void)stream; (void)launches;
  return 0;
}
inline void HaloDestroy(HaloPlan &plan) { (void)plan; }

}  // namespace aither
