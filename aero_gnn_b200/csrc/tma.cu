// tma.cu -- host side of tma.cuh: tensor maps (cuTensorMapEncodeTiled resolved through the runtime at first use, so the
// library has no link-time dependency on libcuda).
#include "common.cuh"
#include "tma.cuh"

namespace aero {

// ---- tensor maps ------------------------------------------------------------------------------------
namespace tma {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
int make_rows_map(const void* base, int64_t rows, CUtensorMap* out) { return make_rows_map_ld(base, rows, 128, 128, out); }
int make_rows_map_ld(const void* base, int64_t rows, int64_t cols, int64_t ld, CUtensorMap* out) {
  EncodeTiledFn fn = resolve_encode();
  if (!fn || rows <= 0 || cols <= 0 || (cols % 64) || (ld % 8) || ld < cols) return 1;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};          // bytes between rows
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}
int make_rows_map_f32(const void* base, int64_t rows, CUtensorMap* out) {
  EncodeTiledFn fn = resolve_encode();
  if (!fn || rows <= 0 || ((uintptr_t)base & 15)) return 1;
  const cuuint64_t dims[2] = {128, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {512};                         // bytes between rows
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}
int make_ids_map(const int32_t* base, int64_t n, uint32_t box, CUtensorMap* out) {
  EncodeTiledFn fn = resolve_encode();
  if (!fn || n <= 0 || box == 0 || box > 256 || (box % 4) || ((uintptr_t)base & 15)) return 1;
  const cuuint64_t dims[1] = {(cuuint64_t)n};
  const cuuint64_t strides[1] = {0};                           // unused for a rank-1 map
  const cuuint32_t bx[1] = {box};
  const cuuint32_t estr[1] = {1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_INT32, 1, const_cast<int32_t*>(base), dims, strides, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}
}  // namespace tma

}  // namespace aero
