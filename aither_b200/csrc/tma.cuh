// tma.cuh -- Tensor Memory Accelerator plumbing for the plane-marching kernels (sm_100a).
//
// Every field of a block is carved from ONE allocation with the same index function
// (layout.cuh), so the whole block is a single 4-D tensor  (i, j, k, field)  with strides
// (1, sj, sk, fs): one CUtensorMap per box shape addresses any field of the block, and one
// cp.async.bulk.tensor instruction brings a (bx, by, 1, nf) tile -- a plane tile of nf consecutive
// fields, halo included -- into shared memory and signals an mbarrier when the bytes have landed.
// Out-of-range coordinates are zero-filled by the hardware, so tiles that stick out of the
// allocation need no clamping.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "layout.cuh"

namespace aither {

// ---- host: tensor-map encoding through the driver entry point (no link-time libcuda) -----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn GetEncodeTiled(std::string *err) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    if (err) *err = "cuTensorMapEncodeTiled is not available from this driver";
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// tensor map over all `nFields` scalar fields of block `b` (allocation base `base`) with a
// (bx, by, 1, nf) box of doubles
inline int EncodeBlockMap(CUtensorMap *out, const BlockDev &b, void *base, int nFields, int bx,
                          int by, int nf, std::string *err) {
  EncodeTiledFn enc = GetEncodeTiled(err);
  if (!enc) return 1;
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(b.sj),
                              static_cast<cuuint64_t>(b.nj + 2 * b.g + 1),
                              static_cast<cuuint64_t>(b.nk + 2 * b.g + 1),
                              static_cast<cuuint64_t>(nFields)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(b.sj) * 8, static_cast<cuuint64_t>(b.sk) * 8,
                                 static_cast<cuuint64_t>(b.fs) * 8};
  const cuuint32_t box[4] = {static_cast<cuuint32_t>(bx), static_cast<cuuint32_t>(by), 1u,
                             static_cast<cuuint32_t>(nf)};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return 1;
  }
  return 0;
}

// ---- device: mbarrier + bulk tensor copy -------------------------------------------------------
__device__ __forceinline__ uint32_t SmemAddr(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void MbarInit(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(SmemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void MbarInitFence() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(SmemAddr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(SmemAddr(bar)),
      "r"(parity)
      : "memory");
}
// (c0, c1, c2, c3) = (i, j, k, field) element coordinates of the tile's first element
__device__ __forceinline__ void TmaLoad4D(void *smemDst, const CUtensorMap *map, int c0, int c1,
                                          int c2, int c3, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(SmemAddr(smemDst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(SmemAddr(bar))
      : "memory");
}

}  // namespace aither
