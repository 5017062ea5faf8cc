// Host-side construction of TMA tensor maps for the [B, L, 256] bf16 row-major operands of the
// attention kernels.  cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint so that
// libsam2b200.so does not link against libcuda directly.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "abi_common.cuh"

namespace sam2b200 {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// x[B][L][256] (bf16, contiguous) as a 3-D tensor (256 elems, L rows, B) with a (64, rows, 1) box:
// one box is a [rows][128 B] slab with the 128-byte swizzle, i.e. one 64-column chunk of a tile in
// the canonical UMMA SW128 layout (K-major when the 256-dim is the contraction, MN-major when
// the rows are).  A [rows x 256] tile is four such boxes, chunk-major in shared memory.
// Rows beyond L are zero-filled by the TMA unit.
// `width` = 256 (projected q / k / v / dO) or 64 (raw memory features, dO Wv): row pitch = width * 2 bytes.
inline int make_rows256_map(CUtensorMap* map, const void* base, int B, int L, int box_rows, int width = 256) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(SAM2B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)L * (cuuint64_t)width * 2};  // bytes, dims 1..2
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(last_error_buffer(), 512, "cuTensorMapEncodeTiled failed (%d) B=%d L=%d rows=%d", (int)r,
             B, L, box_rows);
    return SAM2B200_ERR_DRIVER;
  }
  return SAM2B200_OK;
}

// Output tensor y[B][L][ld] (bf16 or fp32; first 256 columns used) as a 3-D tensor with a box of
// [box_rows][128 bytes] (64 bf16 or 32 fp32 columns), 128-byte swizzle: the staging layout of the TMA-store
// epilogues (one box per warp: its 32 accumulator rows).  Rows beyond L are clipped by the TMA unit.
inline int make_out_map(CUtensorMap* map, void* base, int is_bf16, int B, int L, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(SAM2B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  const cuuint64_t es = is_bf16 ? 2 : 4;
  cuuint64_t dims[3] = {256, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * es, (cuuint64_t)L * (cuuint64_t)ld * es};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(last_error_buffer(), 512, "cuTensorMapEncodeTiled (output) failed (%d) B=%d L=%d ld=%lld", (int)r, B, L, ld);
    return SAM2B200_ERR_DRIVER;
  }
  return SAM2B200_OK;
}

}  // namespace sam2b200
