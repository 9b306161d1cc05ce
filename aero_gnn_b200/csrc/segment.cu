// segment.cu -- row gathers and deterministic segmented reductions (HBM-bound kernels).
// One warp owns one output row; a lane owns 4 consecutive columns per 128-column slab, so every
// global access is a 128-bit (fp32) or 64-bit (bf16) vector and a warp reads whole 512 B / 256 B rows.
#include "common.cuh"

namespace aero {

template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ in, const int32_t* __restrict__ idx,
                                                          const T* __restrict__ add, T* __restrict__ out,
                                                          int64_t n_out, int width) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_out) return;
  int lane = threadIdx.x & 31;
  int64_t srow = idx ? (int64_t)idx[row] : row;
  const bool hole = srow < 0;   // negative index = a zero row (Unpool: fine nodes that were not selected)
  const T* ip = in + (hole ? 0 : srow) * width;
  T* op = out + row * width;
  const T* ap = add ? add + row * width : nullptr;
  if ((width & 3) == 0) {
    for (int c = lane * 4; c < width; c += 128) {
      float4 v = hole ? make_float4(0.f, 0.f, 0.f, 0.f) : load4(ip + c);
      if (ap) {
        float4 a = load4(ap + c);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      store4(op + c, v);
    }
  } else {
    for (int c = lane; c < width; c += 32) {
      float v = hole ? 0.f : load1(ip + c);
      if (ap) v += load1(ap + c);
      store1(op + c, v);
    }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) segment_reduce_kernel(const TI* __restrict__ in, const int32_t* __restrict__ ptr,
                                                             const int32_t* __restrict__ list, TO* __restrict__ out,
                                                             int64_t n_seg, int width, int64_t ld_out, int mean) {
  int64_t seg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seg >= n_seg) return;
  int lane = threadIdx.x & 31;
  int b = ptr[seg], e = ptr[seg + 1];
  float scale = 1.f;
  if (mean) scale = 1.f / (float)(e - b > 1 ? e - b : 1);
  TO* op = out + seg * ld_out;
  if ((width & 3) == 0) {
    for (int c = lane * 4; c < width; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int k = b;
      // two rows in flight to overlap the index -> row dependency
      for (; k + 1 < e; k += 2) {
        int64_t r0 = list ? (int64_t)list[k] : (int64_t)k;
        int64_t r1 = list ? (int64_t)list[k + 1] : (int64_t)k + 1;
        float4 v0 = load4(in + r0 * width + c);
        float4 v1 = load4(in + r1 * width + c);
        acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
        acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
      }
      if (k < e) {
        int64_t r0 = list ? (int64_t)list[k] : (int64_t)k;
        float4 v0 = load4(in + r0 * width + c);
        acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
      }
      if (mean) {
        // divide (not multiply by reciprocal) to match scatter_mean = sum / count
        float cnt = (float)(e - b > 1 ? e - b : 1);
        acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt;
      }
      store4(op + c, acc);
    }
  } else {
    for (int c = lane; c < width; c += 32) {
      float acc = 0.f;
      for (int k = b; k < e; ++k) {
        int64_t r = list ? (int64_t)list[k] : (int64_t)k;
        acc += load1(in + r * width + c);
      }
      if (mean) acc /= (float)(e - b > 1 ? e - b : 1);
      store1(op + c, acc);
    }
  }
  (void)scale;
}

template <typename T>
__global__ void __launch_bounds__(256) segment_bcast_kernel(const T* __restrict__ g_out, const int32_t* __restrict__ seg_of_row,
                                                            const int32_t* __restrict__ ptr, T* __restrict__ g_in,
                                                            int64_t n_rows, int width, int mean) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  int lane = threadIdx.x & 31;
  int seg = seg_of_row[row];
  float cnt = 1.f;
  if (mean) {
    int d = ptr[seg + 1] - ptr[seg];
    cnt = (float)(d > 1 ? d : 1);
  }
  const T* ip = g_out + (int64_t)seg * width;
  T* op = g_in + row * width;
  if ((width & 3) == 0) {
    for (int c = lane * 4; c < width; c += 128) {
      float4 v = load4(ip + c);
      if (mean) { v.x /= cnt; v.y /= cnt; v.z /= cnt; v.w /= cnt; }
      store4(op + c, v);
    }
  } else {
    for (int c = lane; c < width; c += 32) {
      float v = load1(ip + c);
      if (mean) v /= cnt;
      store1(op + c, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Receiver sums that straddle row tiles.  The block kernels write, per tile t, two fp32 partial
// rows: part[t][0] = sum of the leading run when it is incomplete in t (started earlier or runs
// past the end), part[t][1] = sum of the trailing run when it starts inside t and continues.
// Node n with CSR range [b,e) spanning tiles t0 < t1 gets
//     agg[n] = (b % R == 0 ? part[t0][0] : part[t0][1]) + part[t0+1..t1][0]   in that order.
// One block per tile boundary; the block acts only if the run crossing the boundary starts in
// the tile left of it (so each straddling node is handled exactly once).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fixup_store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void fixup_store(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(p);
  o[0] = __floats2bfloat162_rn(v.x, v.y);
  o[1] = __floats2bfloat162_rn(v.z, v.w);
}

// One warp per item, lane = 4 columns.  Blocks [0, n_bound_blocks): item = tile boundary t0 | t0 + 1 -- the receiver
// whose run of rows crosses it gets the sum of its per-tile partial rows, in tile order, from the warp of the tile the
// run starts in.  Further blocks (launched only when the rows of receivers without any edge row must be defined): item
// = 32 receivers; the ones with an empty run are written as zeros -- no tile ever wrote them, so no memset pass is needed.
// Shared by the forward aggregate (fp32 rows) and the receiver sums aero_wgrad takes per tile (bf16 rows, stride ld).
template <typename OutT>
__global__ void __launch_bounds__(256) run_fixup_kernel(const float* __restrict__ part, const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ dst, OutT* __restrict__ out, int64_t ld,
                                                        int64_t rows, int64_t n_nodes, int64_t n_bound_blocks) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if ((int64_t)blockIdx.x >= n_bound_blocks) {
    const int64_t n0 = (w - n_bound_blocks * 8) * 32;
    const int64_t n = n0 + lane;
    uint32_t empty = __ballot_sync(0xffffffffu, n < n_nodes && rowptr[n + 1] == rowptr[n]);
    for (; empty; empty &= empty - 1)
      fixup_store(out + (size_t)(n0 + __ffs(empty) - 1) * ld + lane * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const int64_t t0 = w;
  const int64_t last_row = (t0 + 1) * 128 - 1;
  if (last_row + 1 >= rows) return;       // no boundary after the final tile
  const int n = dst[last_row];
  const int b = rowptr[n], e = rowptr[n + 1];
  if (e <= last_row + 1) return;          // run ends at the boundary: complete
  if (b / 128 != t0) return;              // run started in an earlier tile: another warp owns it
  const int64_t t1 = ((int64_t)e - 1) / 128;
  float4 acc = *reinterpret_cast<const float4*>(part + ((size_t)t0 * 2 + ((b % 128) == 0 ? 0 : 1)) * 128 + lane * 4);
  for (int64_t t = t0 + 1; t <= t1; ++t) {
    const float4 v = *reinterpret_cast<const float4*>(part + ((size_t)t * 2) * 128 + lane * 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  fixup_store(out + (size_t)n * ld + lane * 4, acc);
}

// tile sizes other than 128 rows (the CUDA-core forward): one block per boundary, thread = column
__global__ void __launch_bounds__(128) agg_fixup_kernel(const float* __restrict__ part, const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ dst, float* __restrict__ agg,
                                                        int64_t rows, int tile_rows) {
  int64_t t0 = blockIdx.x;
  int64_t last_row = (t0 + 1) * tile_rows - 1;
  if (last_row + 1 >= rows) return;  // no boundary after the final tile
  int n = dst[last_row];
  int b = rowptr[n], e = rowptr[n + 1];
  if (e <= last_row + 1) return;          // run ends at the boundary: complete
  if (b / tile_rows != t0) return;        // run started in an earlier tile: another block owns it
  int64_t t1 = ((int64_t)e - 1) / tile_rows;
  int c = threadIdx.x;
  float acc = part[((size_t)t0 * 2 + ((b % tile_rows) == 0 ? 0 : 1)) * 128 + c];
  for (int64_t t = t0 + 1; t <= t1; ++t) acc += part[((size_t)t * 2) * 128 + c];
  agg[(size_t)n * 128 + c] = acc;
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int n_parts, size_t stride,
                                                              float* __restrict__ out, size_t n) {
  // 64 outputs per block; the partials are split into four contiguous groups summed by four thread groups, then
  // combined in a fixed order (deterministic for a given n_parts)
  __shared__ float red[4][64];
  const int j = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const size_t o = (size_t)blockIdx.x * 64 + j;
  const int per = (n_parts + 3) / 4;
  const int p0 = grp * per, p1 = min(n_parts, p0 + per);
  float acc = 0.f;
  if (o < n)
    for (int p = p0; p < p1; ++p) acc += part[(size_t)p * stride + o];
  red[grp][j] = acc;
  __syncthreads();
  if (grp == 0 && o < n) out[o] = (red[0][j] + red[1][j]) + (red[2][j] + red[3][j]);
}

int launch_agg_fixup(const float* part, const int32_t* rowptr, float* agg, int64_t rows, int64_t n_nodes,
                     int tile_rows, const int32_t* dst, cudaStream_t st) {
  (void)n_nodes;
  int64_t tiles = cdiv(rows, tile_rows);
  if (tiles <= 1) return AERO_OK;
  agg_fixup_kernel<<<(unsigned)(tiles - 1), 128, 0, st>>>(part, rowptr, dst, agg, rows, tile_rows);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

template <typename OutT>
static int launch_run_fixup(const float* part, const int32_t* rowptr, const int32_t* dst, OutT* out, int64_t ld, int64_t rows,
                            int64_t n_nodes, bool zero_empty, cudaStream_t st) {
  const int64_t tiles = cdiv(rows > 0 ? rows : 1, 128);
  const int64_t n_bound = cdiv(tiles - 1, 8), n_zero = zero_empty ? cdiv(n_nodes, 256) : 0;
  if (n_bound + n_zero <= 0) return AERO_OK;
  run_fixup_kernel<OutT><<<(unsigned)(n_bound + n_zero), 256, 0, st>>>(part, rowptr, dst, out, ld, rows, n_nodes, n_bound);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

int launch_agg_fixup128(const float* part, const int32_t* rowptr, const int32_t* dst, float* agg, int64_t rows,
                        int64_t n_nodes, bool zero_empty, cudaStream_t st) {
  return launch_run_fixup<float>(part, rowptr, dst, agg, 128, rows, n_nodes, zero_empty, st);
}

int launch_gpd_fixup(const float* part, const int32_t* rowptr, const int32_t* dst, __nv_bfloat16* out, int64_t ld,
                     int64_t rows, int64_t n_nodes, cudaStream_t st) {
  return launch_run_fixup<__nv_bfloat16>(part, rowptr, dst, out, ld, rows, n_nodes, true, st);
}

int launch_reduce_partials(const float* part, int n_parts, size_t stride, float* out, size_t n, cudaStream_t st) {
  if (n == 0) return AERO_OK;
  reduce_partials_kernel<<<(unsigned)cdiv((int64_t)n, 64), 256, 0, st>>>(part, n_parts, stride, out, n);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

}  // namespace aero

using namespace aero;

extern "C" int aero_gather_rows(const void* in, const int32_t* idx, const void* add, void* out, int64_t n_out,
                                int64_t width, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(n_out >= 0 && width > 0 && width < (1 << 20), "aero_gather_rows: bad sizes");
  if (n_out == 0) return AERO_OK;
  AERO_CHECK_ARG(in && out, "aero_gather_rows: null pointer");
  unsigned blocks = (unsigned)cdiv(n_out, 8);
  if (dtype == AERO_F32)
    gather_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)in, idx, (const float*)add, (float*)out, n_out, (int)width);
  else if (dtype == AERO_BF16)
    gather_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in, idx, (const __nv_bfloat16*)add,
                                                               (__nv_bfloat16*)out, n_out, (int)width);
  else {
    set_error("aero_gather_rows: unsupported dtype %d", dtype);
    return AERO_EUNSUPPORTED;
  }
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" int aero_segment_reduce_ld(const void* in, const int32_t* ptr, const int32_t* list, void* out, int64_t n_seg,
                                      int64_t width, int64_t ld_out, int in_dtype, int out_dtype, int mean,
                                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(n_seg >= 0 && width > 0 && width < (1 << 20), "aero_segment_reduce: bad sizes");
  AERO_CHECK_ARG(ld_out >= width && (((width & 3) != 0) || (ld_out & 3) == 0),
                 "aero_segment_reduce: ld_out must be >= width (and a multiple of 4 for vector stores)");
  if (n_seg == 0) return AERO_OK;
  AERO_CHECK_ARG(ptr && out, "aero_segment_reduce: null pointer");
  unsigned blocks = (unsigned)cdiv(n_seg, 8);
  int w = (int)width;
  if (in_dtype == AERO_F32 && out_dtype == AERO_F32)
    segment_reduce_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)in, ptr, list, (float*)out, n_seg, w, ld_out, mean);
  else if (in_dtype == AERO_BF16 && out_dtype == AERO_F32)
    segment_reduce_kernel<__nv_bfloat16, float><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in, ptr, list, (float*)out, n_seg, w, ld_out, mean);
  else if (in_dtype == AERO_BF16 && out_dtype == AERO_BF16)
    segment_reduce_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in, ptr, list, (__nv_bfloat16*)out, n_seg, w, ld_out, mean);
  else if (in_dtype == AERO_F32 && out_dtype == AERO_BF16)
    segment_reduce_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>((const float*)in, ptr, list, (__nv_bfloat16*)out, n_seg, w, ld_out, mean);
  else {
    set_error("aero_segment_reduce: unsupported dtypes %d -> %d", in_dtype, out_dtype);
    return AERO_EUNSUPPORTED;
  }
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" int aero_segment_reduce(const void* in, const int32_t* ptr, const int32_t* list, void* out, int64_t n_seg,
                                   int64_t width, int in_dtype, int out_dtype, int mean, void* stream) {
  return aero_segment_reduce_ld(in, ptr, list, out, n_seg, width, width, in_dtype, out_dtype, mean, stream);
}

extern "C" int aero_segment_bcast(const void* g_out, const int32_t* seg_of_row, const int32_t* ptr, void* g_in,
                                  int64_t n_rows, int64_t width, int dtype, int mean, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(n_rows >= 0 && width > 0 && width < (1 << 20), "aero_segment_bcast: bad sizes");
  if (n_rows == 0) return AERO_OK;
  AERO_CHECK_ARG(g_out && seg_of_row && g_in && (!mean || ptr), "aero_segment_bcast: null pointer");
  unsigned blocks = (unsigned)cdiv(n_rows, 8);
  if (dtype == AERO_F32)
    segment_bcast_kernel<float><<<blocks, 256, 0, st>>>((const float*)g_out, seg_of_row, ptr, (float*)g_in, n_rows, (int)width, mean);
  else if (dtype == AERO_BF16)
    segment_bcast_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)g_out, seg_of_row, ptr, (__nv_bfloat16*)g_in, n_rows, (int)width, mean);
  else {
    set_error("aero_segment_bcast: unsupported dtype %d", dtype);
    return AERO_EUNSUPPORTED;
  }
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

// ---------------------------------------------------------------------------------------------
// Multi-segment strided copy with dtype conversion: one launch packs all parameters of a processor step into the
// images the block kernels read (forward) or scatters the packed gradients back into per-parameter buffers (backward).
// ---------------------------------------------------------------------------------------------
namespace aero {

struct CopySegs {
  aero_copy_seg s[AERO_MAX_COPY_SEGS];
};

template <typename TS, typename TD>
__device__ __forceinline__ void copy_seg_elems(const aero_copy_seg& g) {
  const TS* src = reinterpret_cast<const TS*>(g.src);
  TD* dst = reinterpret_cast<TD*>(g.dst);
  const int64_t n = g.rows * g.cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / g.cols, c = i - r * g.cols;
    store1(dst + r * g.dst_ld + c, src ? load1(src + r * g.src_ld + c) : 0.f);
  }
}

__global__ void __launch_bounds__(256) multi_copy_kernel(const __grid_constant__ CopySegs segs) {
  const aero_copy_seg& g = segs.s[blockIdx.y];
  if (g.src_dtype == AERO_F32) {
    if (g.dst_dtype == AERO_F32) copy_seg_elems<float, float>(g);
    else copy_seg_elems<float, __nv_bfloat16>(g);
  } else {
    if (g.dst_dtype == AERO_F32) copy_seg_elems<__nv_bfloat16, float>(g);
    else copy_seg_elems<__nv_bfloat16, __nv_bfloat16>(g);
  }
}

}  // namespace aero

extern "C" int aero_multi_copy(const aero_copy_seg* segs, int n_segs, void* stream) {
  g_launch_count = 0;
  AERO_CHECK_ARG(n_segs >= 0 && n_segs <= AERO_MAX_COPY_SEGS && (n_segs == 0 || segs),
                 "aero_multi_copy: 0 <= n_segs <= %d", AERO_MAX_COPY_SEGS);
  if (n_segs == 0) return AERO_OK;
  CopySegs cs;
  memset(&cs, 0, sizeof(cs));
  int64_t max_elems = 1;
  for (int i = 0; i < n_segs; ++i) {
    const aero_copy_seg& g = segs[i];
    AERO_CHECK_ARG(g.rows >= 0 && g.cols >= 0 && (g.rows * g.cols == 0 || g.dst) && g.src_ld >= 0 && g.dst_ld >= 0,
                   "aero_multi_copy: bad segment %d", i);
    if ((g.src_dtype != AERO_F32 && g.src_dtype != AERO_BF16) || (g.dst_dtype != AERO_F32 && g.dst_dtype != AERO_BF16)) {
      set_error("aero_multi_copy: unsupported dtype in segment %d", i);
      return AERO_EUNSUPPORTED;
    }
    cs.s[i] = g;
    if (g.rows * g.cols > max_elems) max_elems = g.rows * g.cols;
  }
  int64_t bx = cdiv(max_elems, 256 * 4);
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, (unsigned)n_segs);
  multi_copy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cs);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}
