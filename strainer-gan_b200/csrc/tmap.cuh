// Host-side CUtensorMap encoding through the driver entry point resolved in sg_init (no link-time libcuda).
#pragma once
#include "common.cuh"

namespace sg {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// dims / box: innermost first; strides: bytes of dims 1..rank-1 (multiples of 16).  Out-of-range box elements
// (negative coordinates included) are filled with zeros.
static inline int encode_tmap(CUtensorMap* m, int rank, const void* ptr, const cuuint64_t* dims, const cuuint64_t* strides,
                              const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
                              CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                              const cuuint32_t* element_strides = nullptr) {
  // element_strides: traversal stride per dimension (a box of extent E with stride s delivers ceil(E / s) elements);
  // used by the stride-2 convolutions to fetch one filter tap of a tile of output pixels as a dense operand
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (element_strides)
    for (int i = 0; i < rank; ++i) estr[i] = element_strides[i];
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(state().encode_tiled);
  CUresult r = fn(m, dtype, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return SG_ECUDA;
  }
  return SG_OK;
}

}  // namespace sg
