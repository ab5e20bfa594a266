// tmap.h — host-side TMA tensor-map construction.
// cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint so that
// libb200ltx.so has no link-time dependency on libcuda (it must dlopen on a
// CPU-only box for the symbol-export test).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (PFN_encodeTiled)p;
  }
  return fn;
}

enum TmapDType { TM_BF16 = 0, TM_F32 = 1 };

// Row-major tensor of up to 3 dims. dims[0] is the innermost (contiguous) extent in elements,
// strides_bytes[i] is the byte stride of dims[i+1].  box[] in elements.  swizzle128: box[0]*elem
// must be <= 128 bytes.  Returns 0 on success.
inline int make_tmap(CUtensorMap* out, const void* base, TmapDType dt, int rank,
                     const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                     bool swizzle128) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return -100;
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3];
  cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, dt == TM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)(1000 + r);
}

// 2-D bf16 row-major [rows, cols] with row pitch ld (elements); box = [box_rows, box_cols].
inline int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                             uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 2};
  uint32_t box[2] = {box_cols, box_rows};
  return make_tmap(out, base, TM_BF16, 2, dims, str, box, true);
}

}  // namespace b200
