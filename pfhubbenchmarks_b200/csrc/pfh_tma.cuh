// Shared TMA / mbarrier plumbing of the TMA-staged kernels (pfh_scalar3_tma.cu, pfh_stencil3_tma.cu).
#pragma once
#include <cuda.h>

#include "pfh_internal.h"

namespace tma {

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"   // suspends the thread up to the hint instead of spinning
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680)
      : "memory");
}
static __device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// 8-byte asynchronous copy (LDGSTS); src_bytes = 0 zero-fills the destination
static __device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// the executing thread arrives on `bar` once all its earlier cp.async copies have landed (the arrival is part of the
// barrier's expected count: initialise it with the number of arriving threads)
static __device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

static constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    PFH_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    PFH_CHECK(q == cudaDriverEntryPointSuccess && p != nullptr, "cuTensorMapEncodeTiled is not available in this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp64 tensor (nxv, nyv, nz) at `ptr`, box (bx, by, 1), zero fill outside
static inline CUtensorMap make_map(const double* ptr, int nxv, int nyv, int nz, int bx, int by) {
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)nxv, (cuuint64_t)nyv, (cuuint64_t)nz};
  const cuuint64_t strides[2] = {(cuuint64_t)nxv * 8, (cuuint64_t)nxv * nyv * 8};
  const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult rc = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(ptr), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) throw PfhError("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
  return map;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
}  // namespace tma

// make_map through a per-ctx cache keyed on (pointer, shape, box): the operands of the solver's launches are a handful
// of fixed work vectors, so the driver call is paid once per operand instead of five times per launch
CUtensorMap pfh_tensor_map(pfh_ctx* c, const double* ptr, int nxv, int nyv, int nz, int bx, int by);
