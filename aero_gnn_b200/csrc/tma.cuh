// tma.cuh -- Tensor Memory Accelerator helpers: tensor maps over [rows,128] bf16 row matrices and the
// cp.async.bulk.tensor instructions that move one 64-column x 128-row panel of a row tile (umma.cuh) between
// global memory and the SWIZZLE_128B shared-memory tile format, asynchronously and without any thread touching
// the data.  Out-of-range rows are zero-filled on load and clipped on store, so the ragged last tile needs no code.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace aero {
namespace tma {

// host: tensor map of a row-major [rows,128] bf16 matrix, box = 64 columns (128 bytes) x 128 rows, SWIZZLE_128B.
// Returns 0 on success; the driver entry point is resolved at run time (no link-time dependency on libcuda).
int make_rows_map(const void* base, int64_t rows, CUtensorMap* out);
// same for a matrix with `cols` (a multiple of 64) columns and a row stride of `ld` elements: panel p = columns [64p, 64p+64)
int make_rows_map_ld(const void* base, int64_t rows, int64_t cols, int64_t ld, CUtensorMap* out);

// host: tensor map of a row-major [rows,128] FP32 matrix, box = 32 columns (128 bytes) x 128 rows, SWIZZLE_128B: an fp32
// row tile is four 16 KB panels with the same chunk swizzle as a bf16 panel (16-byte chunk c of row r at c ^ (r & 7))
int make_rows_map_f32(const void* base, int64_t rows, CUtensorMap* out);
// host: 1-D tensor map over n int32 ids, box = `box` ids (box * 4 a multiple of 16 bytes), no swizzle; coordinates
// outside [0, n) -- negative ones included -- are zero-filled
int make_ids_map(const int32_t* base, int64_t n, uint32_t box, CUtensorMap* out);

__device__ __forceinline__ void prefetch_map(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(tm) : "memory");
}
// one panel: columns [64*panel, 64*panel+64) of rows [row0, row0+128) -> dst (16 KB, 1024-byte aligned)
__device__ __forceinline__ void load_panel(uint32_t dst_saddr, const CUtensorMap* tm, int panel, int row0, uint32_t mbar_saddr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst_saddr),
      "l"(tm), "r"(panel * 64), "r"(row0), "r"(mbar_saddr)
      : "memory");
}
// ids [x0, x0 + box) of a make_ids_map tensor map -> dst (16-byte aligned)
__device__ __forceinline__ void load_ids(uint32_t dst_saddr, const CUtensorMap* tm, int x0, uint32_t mbar_saddr) {
  asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];\n" ::"r"(dst_saddr),
               "l"(tm), "r"(x0), "r"(mbar_saddr)
               : "memory");
}
// whole 128x128 row tile (two panels); the caller has armed the mbarrier with 32768 bytes
__device__ __forceinline__ void load_tile(uint32_t dst_saddr, const CUtensorMap* tm, int row0, uint32_t mbar_saddr) {
  load_panel(dst_saddr, tm, 0, row0, mbar_saddr);
  load_panel(dst_saddr + 16384u, tm, 1, row0, mbar_saddr);
}
// hint: bring both panels of the row tile at row0 into L2
__device__ __forceinline__ void prefetch_tile_l2(const CUtensorMap* tm, int row0) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(tm), "r"(0), "r"(row0) : "memory");
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(tm), "r"(64), "r"(row0) : "memory");
}
__device__ __forceinline__ void store_tile(const CUtensorMap* tm, uint32_t src_saddr, int row0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tm), "r"(0), "r"(row0),
               "r"(src_saddr)
               : "memory");
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tm), "r"(64), "r"(row0),
               "r"(src_saddr + 16384u)
               : "memory");
}
// same into columns [col0, col0 + 128) of a wider matrix (a make_rows_map_ld map)
__device__ __forceinline__ void store_tile_at(const CUtensorMap* tm, uint32_t src_saddr, int col0, int row0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tm), "r"(col0), "r"(row0),
               "r"(src_saddr)
               : "memory");
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tm), "r"(col0 + 64),
               "r"(row0), "r"(src_saddr + 16384u)
               : "memory");
}
// one fp32 panel (32 columns x 128 rows, 16 KB) -> columns [col0, col0 + 32) of a make_rows_map_f32 matrix
__device__ __forceinline__ void store_panel_f32(const CUtensorMap* tm, uint32_t src_saddr, int col0, int row0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tm), "r"(col0), "r"(row0),
               "r"(src_saddr)
               : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// the issuing thread's bulk stores have finished READING shared memory (the tiles may be overwritten)
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// ... have completed entirely (before the CTA exits)
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

}  // namespace tma
}  // namespace aero
