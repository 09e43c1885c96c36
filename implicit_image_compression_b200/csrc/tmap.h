// Host-side TMA descriptor helper.  cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint so the library does not link against libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace sb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Row-major [rows, cols] tensor of 2-byte elements; box = {64 cols, box_rows}, 128B swizzle.
// Returns 0 on success.
inline int make_tmap_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                           uint32_t box_rows, bool bf16) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  memset(out, 0, sizeof(*out));
  CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -int(r);
}

// The same row-major [rows, cols] tensor seen as [cols / 64 chunks][rows][64 columns]: ONE box {64, box_rows, chunks}
// lands in shared memory as `chunks` consecutive {64 x box_rows} tiles, each in the 128B-swizzled layout the 2-D map
// above produces - a whole multi-chunk operand with a single TMA instruction.  Coordinates: (0, row, first chunk).
inline int make_tmap_16bit_chunks(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                                  uint32_t box_rows, uint32_t box_chunks) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return -1;
  cuuint64_t dims[3] = {64, rows, cols / 64};
  cuuint64_t strides[2] = {cols * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  memset(out, 0, sizeof(*out));
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -int(r);
}

}  // namespace sb
