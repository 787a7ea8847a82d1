// tmap.cuh — 2-D TMA tensor maps over the f32 residual stream (sm_100a).
//
// The (B,T,K,N) stream seen from the band axis is a row-major matrix: row = (b,t), row stride K*N floats, and the 128 rows
// x N floats of a (sequence tile, band) block form a rectangular box of it.  One cp.async.bulk.tensor copy moves the whole
// box; the 1-D bulk copies it replaces needed one copy per 784-byte row, and the TMA unit's rate of ~1 small copy per ~90
// cycles per SM — not HBM — then paced the band-axis Linear+skip epilogue, BandSplit's store-only GEMM and norm_cast
// (profiles/r02 call50-52).  Rows past the end of the matrix are clipped (stores) / zero-filled (loads) by the hardware.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsrnn {

// rows x inner floats, row stride row_stride_bytes (multiple of 16), box = box_rows x box_inner; false if the driver
// entry point is missing or the encode fails (callers then keep the 1-D bulk copies)
static inline bool make_tmap_2d_f32(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                                    uint32_t box_inner, uint32_t box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
    else
      cudaGetLastError();
  }
  if (!fn || (reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15) || box_inner > 256 || box_rows > 256 ||
      (box_inner * 4) % 16 != 0)
    return false;
  const cuuint64_t gdim[2] = {inner, rows};
  const cuuint64_t gstride[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t tm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// box at (col c0, row c1) -> dense shared-memory box (128-byte aligned); completion bytes on a CTA mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(tm_smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(tm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(tm_smem_u32(smem_src)) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
#endif

}  // namespace bsrnn
