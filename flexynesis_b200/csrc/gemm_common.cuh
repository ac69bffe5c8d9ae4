// Pieces shared by the two GEMM translation units (gemm_umma.cu: one tile per CTA; gemm_umma2.cu: persistent CTA pairs).
#pragma once
#include "fxn_internal.h"
#include "ptx.cuh"

namespace fxn {

__device__ __forceinline__ float epi_activation(float x, int act) {
  switch (act) {
    case 1: return fmaxf(x, 0.f);
    case 3: return 1.f / (1.f + __expf(-x));
    case 6: return x > 0.f ? x : 0.2f * x;
    default: return x;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// 2D bf16 row-major array [rows x cols], leading dimension ld (elements). Box = {64 cols, box_rows}.
inline int make_tensor_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(FXN_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return set_error(FXN_ERR_ARG, "bf16 plane must be 16B aligned with ld %% 8 == 0");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(FXN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return 0;
}

// persistent CTA-pair kernel (gemm_umma2.cu)
int gemm2_dispatch(const fxn_gemm_desc* d, cudaStream_t stream);

}  // namespace fxn
