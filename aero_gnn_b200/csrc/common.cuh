// common.cuh -- shared helpers for libaero_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/aero_gnn.h"

namespace aero {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local int g_launch_count;

#define AERO_CHECK_ARG(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::aero::set_error(__VA_ARGS__);                              \
      return AERO_EINVAL;                                          \
    }                                                              \
  } while (0)

#define AERO_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::aero::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return AERO_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define AERO_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    ::aero::g_launch_count++;                                                             \
    if (_e != cudaSuccess) {                                                              \
      ::aero::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return AERO_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// carve sub-buffers out of a caller workspace (256-byte aligned pieces)
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    T* r = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T), 256);
    return r;
  }
};

int sm_count();

// ---- row element access -----------------------------------------------------------------------
// 4 consecutive columns of a 128-wide (or any 4-aligned) row, as fp32.
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  float4 r;
  r.x = __uint_as_float(v.x << 16);
  r.y = __uint_as_float(v.x & 0xffff0000u);
  r.z = __uint_as_float(v.y << 16);
  r.w = __uint_as_float(v.y & 0xffff0000u);
  return r;
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
template <typename T>
__device__ __forceinline__ float load1(const T* p);
template <>
__device__ __forceinline__ float load1<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void store1(T* p, float v);
template <>
__device__ __forceinline__ void store1<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

template <typename T>
__device__ __forceinline__ float round_to(float v);
template <>
__device__ __forceinline__ float round_to<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_to<__nv_bfloat16>(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

// ---- activations (derivative expressed through the activation output) -----------------------
__device__ __forceinline__ float act_fwd(float v, int act) {
  switch (act) {
    case AERO_ACT_RELU: return fmaxf(v, 0.f);
    case AERO_ACT_TANH: return tanhf(v);
    case AERO_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case AERO_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case AERO_ACT_LEAKY_RELU: return v > 0.f ? v : 0.01f * v;
  }
  return v;
}
__device__ __forceinline__ float act_grad_from_out(float h, int act) {
  switch (act) {
    case AERO_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case AERO_ACT_TANH: return 1.f - h * h;
    case AERO_ACT_SIGMOID: return h * (1.f - h);
    case AERO_ACT_ELU: return h > 0.f ? 1.f : h + 1.f;
    case AERO_ACT_LEAKY_RELU: return h > 0.f ? 1.f : 0.01f;
  }
  return 1.f;
}

// ---- packed block weights ---------------------------------------------------------------------
struct PackedLayout {
  int L;
  __host__ __device__ size_t w_main() const { return 0; }
  __host__ __device__ size_t w_hidden(int l) const { return (size_t)(1 + l) * 16384; }  // l = 0..L-1
  __host__ __device__ size_t w_out() const { return (size_t)(1 + L) * 16384; }
  __host__ __device__ size_t b_hidden(int l) const { return (size_t)(2 + L) * 16384 + (size_t)l * 128; }
  __host__ __device__ size_t b_out() const { return (size_t)(2 + L) * 16384 + (size_t)L * 128; }
  __host__ __device__ size_t gamma() const { return b_out() + 128; }
  __host__ __device__ size_t beta() const { return b_out() + 256; }
  // gradient-only slot: column sums of dL/d(first pre-activation) = gradient of the first Linear's bias, which the
  // caller folds into P (unused, zero, in the weight vector itself)
  __host__ __device__ size_t bias0() const { return b_out() + 384; }
  __host__ __device__ size_t total() const { return b_out() + 512; }
};

}  // namespace aero

// entry points implemented per translation unit
namespace aero {
int simt_block_fwd(const aero_block_desc* d, cudaStream_t st);
int simt_block_bwd(const aero_block_desc* d, cudaStream_t st);
size_t simt_block_workspace_bytes(const aero_block_desc* d, int backward);
size_t simt_prepared_bytes(int L);
int simt_prepare(const float* w, int L, void* prepared, cudaStream_t st);

int umma_block_fwd(const aero_block_desc* d, cudaStream_t st);
int umma_block_bwd(const aero_block_desc* d, cudaStream_t st);
size_t umma_block_workspace_bytes(const aero_block_desc* d, int backward);
size_t umma_prepared_bytes(int L);
int umma_prepare(const float* w, int L, void* prepared, cudaStream_t st);

// agg partial fix-up shared by both paths: complete the receiver sums that straddle row tiles
int launch_agg_fixup(const float* part, const int32_t* rowptr, float* agg, int64_t rows,
                     int64_t n_nodes, int tile_rows, const int32_t* dst, cudaStream_t st);
// 128-row tiles (the tcgen05 forward): one warp per boundary; zero_empty also writes zeros into the rows of receivers
// without any row, which replaces the memset of the whole aggregate
int launch_agg_fixup128(const float* part, const int32_t* rowptr, const int32_t* dst, float* agg, int64_t rows,
                        int64_t n_nodes, bool zero_empty, cudaStream_t st);
// receiver sums taken tile by tile by aero_wgrad (bf16 rows, stride ld): complete the runs that straddle 128-row tiles
// from the per-tile fp32 partial rows, and zero the rows of receivers without any row
int launch_gpd_fixup(const float* part, const int32_t* rowptr, const int32_t* dst, __nv_bfloat16* out, int64_t ld,
                     int64_t rows, int64_t n_nodes, cudaStream_t st);
// deterministic reduction of per-CTA weight-gradient partials: out[j] = sum_c part[c*stride + j]
int launch_reduce_partials(const float* part, int n_parts, size_t stride, float* out, size_t n,
                           cudaStream_t st);
// inclusive prefix sum of n int32 (sort_plan.cu); `out` may alias `in`; ws needs scan_ws_bytes(n)
size_t scan_ws_bytes(int64_t n);
int inclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, cudaStream_t st);
}  // namespace aero
