// Host-side construction of TMA tensor maps for (B, NH, S, DH) activation tensors with
// arbitrary B/NH/S strides (innermost DH contiguous).  The driver entry point is resolved at
// run time through cudaGetDriverEntryPoint, so the library does not link against libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlstm {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  return fn;
}

// bf16 tensor (B, NH, S, DH); strides in elements.  Box = 64 (DH) x box_rows (S) x 1 x 1,
// 128-byte swizzle, out-of-bounds rows read as zero / are clipped on store.
// Returns 0 on success, else the CUresult (or -1 if the encoder is unavailable).
inline int make_act_tmap(CUtensorMap* out, const void* ptr, int B, int NH, int S, int DH, int64_t stride_b,
                         int64_t stride_h, int64_t stride_s, int box_rows) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)DH, (cuuint64_t)S, (cuuint64_t)NH, (cuuint64_t)B};
  // size-1 dimensions may come with arbitrary strides; give them a legal one
  int64_t ss = (S > 1) ? stride_s : (int64_t)DH;
  int64_t sh = (NH > 1) ? stride_h : (int64_t)DH * S;
  int64_t sb = (B > 1) ? stride_b : (int64_t)DH * S * NH;
  cuuint64_t strides[3] = {(cuuint64_t)ss * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

}  // namespace mlstm
