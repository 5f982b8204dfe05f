// TMA (cp.async.bulk.tensor) + mbarrier helpers for sm_100a, and the host-side tensor-map encoder.
// The driver entry point cuTensorMapEncodeTiled is looked up at run time through the CUDA runtime
// (cudaGetDriverEntryPoint), so the library does not link against libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace rg {
namespace tma {

// ---- host ---------------------------------------------------------------------------------------
// Tensor map of a scratch array laid out [comp][plane][j][i] (i fastest, element size esize) with a
// box of (bx, by, 1, ncomp) elements: ONE bulk-tensor copy brings a (bx x by) tile of one plane for
// all components into shared memory as [comp][by][bx]; out-of-range elements are zero-filled.
// Returns false when the layout cannot be described (row pitch not a multiple of 16 bytes, ...).
inline bool encodeTile4D(CUtensorMap* map, const void* base, int esize, int isize, int jsize, int planes, int ncomp,
                         int bx, int by) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
    fn = reinterpret_cast<EncodeFn>(p);
  }
  if (((size_t)isize * esize) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  if (((size_t)bx * esize) % 16 != 0 || bx > 256 || by > 256 || ncomp > 256) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)isize, (cuuint64_t)jsize, (cuuint64_t)planes, (cuuint64_t)ncomp};
  const cuuint64_t strides[3] = {(cuuint64_t)isize * esize, (cuuint64_t)isize * jsize * esize,
                                 (cuuint64_t)isize * jsize * planes * esize};
  const cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1u, (cuuint32_t)ncomp};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapDataType dt = (esize == 8) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  return fn(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// ---- device -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void fenceBarrierInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smemAddr(bar)),
      "r"(parity)
      : "memory");
}
// one (bx x by x 1 x ncomp) box at element coordinates (i, j, plane, comp) -> dst, completion on bar
__device__ __forceinline__ void loadTile4D(void* dst, const CUtensorMap* map, uint64_t* bar, int i, int j, int plane,
                                           int comp) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smemAddr(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smemAddr(bar)), "r"(i), "r"(j), "r"(plane), "r"(comp)
      : "memory");
}

}  // namespace tma
}  // namespace rg
