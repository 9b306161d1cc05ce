// thin_linear.cu -- the encoders' first Linear on raw features: out[rows,128] = x[rows,K] W^T + b with K <= 16
// (mgn.py:123-124: node_encoder / edge_encoder start with Linear(6 | 4, 128); mlp.py:40-44).  K is far too small for
// a tensor-core tile; the op is bound by writing (forward) / reading (backward) the 128-wide rows, so it is a plain
// streaming kernel: a warp per row, lane = 4 output columns with its 4 x K weights in registers, the row's K inputs
// broadcast by shuffles.  The backward takes d(W) and d(b) in ONE pass over the gradient rows (the library path was a
// GEMM plus a separate column-sum reduction), per-block partials summed in block order (deterministic).
#include "common.cuh"

namespace aero {

constexpr int TL_THREADS = 256;

template <typename T, int KP>
__global__ void __launch_bounds__(TL_THREADS) thin_linear_fwd_kernel(const T* __restrict__ x, int64_t ldx,
                                                                     const T* __restrict__ W, const T* __restrict__ b,
                                                                     T* __restrict__ out, int64_t rows, int K) {
  const int lane = threadIdx.x & 31, c = lane * 4;
  float w[4][KP], bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bias[j] = b ? load1(b + c + j) : 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) w[j][k] = k < K ? load1(W + (size_t)(c + j) * K + k) : 0.f;
  }
  const int64_t warps = (int64_t)gridDim.x * (TL_THREADS / 32);
  for (int64_t r = (int64_t)blockIdx.x * (TL_THREADS / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float xl = lane < K ? load1(x + r * ldx + lane) : 0.f;
    float4 acc = make_float4(bias[0], bias[1], bias[2], bias[3]);
#pragma unroll
    for (int k = 0; k < KP; ++k) {   // k ascending: the order of a dot product over the input features
      const float xk = __shfl_sync(0xffffffffu, xl, k);
      acc.x = fmaf(w[0][k], xk, acc.x); acc.y = fmaf(w[1][k], xk, acc.y);
      acc.z = fmaf(w[2][k], xk, acc.z); acc.w = fmaf(w[3][k], xk, acc.w);
    }
    store4(out + r * 128 + c, acc);
  }
}

// part[block][col][K + 1]: d(W)[col][k] = sum_r g[r][col] x[r][k], slot K = d(b)[col] = sum_r g[r][col]
template <typename T, int KP>
__global__ void __launch_bounds__(TL_THREADS) thin_linear_bwd_kernel(const T* __restrict__ g, const T* __restrict__ x,
                                                                     int64_t ldx, float* __restrict__ part, int64_t rows,
                                                                     int K) {
  __shared__ float red[128 * (KP + 1)];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, c = lane * 4;
  float dw[4][KP], db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < KP; ++k) dw[j][k] = 0.f;
  const int64_t warps = (int64_t)gridDim.x * (TL_THREADS / 32);
  for (int64_t r = (int64_t)blockIdx.x * (TL_THREADS / 32) + wid; r < rows; r += warps) {
    const float xl = lane < K ? load1(x + r * ldx + lane) : 0.f;
    const float4 gv = load4(g + r * 128 + c);
    db[0] += gv.x; db[1] += gv.y; db[2] += gv.z; db[3] += gv.w;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const float xk = __shfl_sync(0xffffffffu, xl, k);
      dw[0][k] = fmaf(gv.x, xk, dw[0][k]); dw[1][k] = fmaf(gv.y, xk, dw[1][k]);
      dw[2][k] = fmaf(gv.z, xk, dw[2][k]); dw[3][k] = fmaf(gv.w, xk, dw[3][k]);
    }
  }
  // the 8 warps of the block add their sums in warp order (fixed order -> deterministic)
  for (int w = 0; w < TL_THREADS / 32; ++w) {
    if (wid == w) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          float* p = red + (c + j) * (KP + 1) + k;
          *p = (w == 0 ? 0.f : *p) + dw[j][k];
        }
        float* p = red + (c + j) * (KP + 1) + KP;
        *p = (w == 0 ? 0.f : *p) + db[j];
      }
    }
    __syncthreads();
  }
  float* po = part + (size_t)blockIdx.x * 128 * (K + 1);
  for (int i = threadIdx.x; i < 128 * (K + 1); i += TL_THREADS) {
    const int col = i / (K + 1), k = i - col * (K + 1);
    po[i] = red[col * (KP + 1) + (k == K ? KP : k)];
  }
}

static int tl_grid(int64_t rows) {
  int64_t full = cdiv(rows > 0 ? rows : 1, TL_THREADS / 32);
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(full < cap ? full : cap);
}

template <typename T>
static int tl_fwd(const void* x, int64_t ldx, const void* W, const void* b, void* out, int64_t rows, int K, cudaStream_t st) {
  const int grid = tl_grid(rows);
  const T *xp = (const T*)x, *wp = (const T*)W, *bp = (const T*)b;
  if (K <= 4) thin_linear_fwd_kernel<T, 4><<<grid, TL_THREADS, 0, st>>>(xp, ldx, wp, bp, (T*)out, rows, K);
  else if (K <= 8) thin_linear_fwd_kernel<T, 8><<<grid, TL_THREADS, 0, st>>>(xp, ldx, wp, bp, (T*)out, rows, K);
  else thin_linear_fwd_kernel<T, 16><<<grid, TL_THREADS, 0, st>>>(xp, ldx, wp, bp, (T*)out, rows, K);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

template <typename T>
static int tl_bwd(const void* g, const void* x, int64_t ldx, float* part, int grid, int64_t rows, int K, cudaStream_t st) {
  const T *gp = (const T*)g, *xp = (const T*)x;
  if (K <= 4) thin_linear_bwd_kernel<T, 4><<<grid, TL_THREADS, 0, st>>>(gp, xp, ldx, part, rows, K);
  else if (K <= 8) thin_linear_bwd_kernel<T, 8><<<grid, TL_THREADS, 0, st>>>(gp, xp, ldx, part, rows, K);
  else thin_linear_bwd_kernel<T, 16><<<grid, TL_THREADS, 0, st>>>(gp, xp, ldx, part, rows, K);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

}  // namespace aero

using namespace aero;

extern "C" int aero_thin_linear_fwd(const void* x, int64_t ldx, const void* W, const void* b, void* out, int64_t rows,
                                    int K, int dtype, void* stream) {
  g_launch_count = 0;
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(K >= 1 && K <= 16 && rows >= 0 && ldx >= K, "aero_thin_linear_fwd: 1 <= K <= 16, ldx >= K");
  if (rows == 0) return AERO_OK;
  AERO_CHECK_ARG(x && W && out, "aero_thin_linear_fwd: null pointer");
  if (dtype == AERO_F32) return tl_fwd<float>(x, ldx, W, b, out, rows, K, st);
  if (dtype == AERO_BF16) return tl_fwd<__nv_bfloat16>(x, ldx, W, b, out, rows, K, st);
  set_error("aero_thin_linear_fwd: unsupported dtype %d", dtype);
  return AERO_EUNSUPPORTED;
}

extern "C" size_t aero_thin_linear_workspace_bytes(int64_t rows, int K) {
  return align_up((size_t)tl_grid(rows) * 128 * (K + 1) * sizeof(float), 256);
}

extern "C" int aero_thin_linear_bwd(const void* g, const void* x, int64_t ldx, float* dwb, int64_t rows, int K, int dtype,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  g_launch_count = 0;
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(K >= 1 && K <= 16 && rows >= 0 && ldx >= K && dwb, "aero_thin_linear_bwd: 1 <= K <= 16, ldx >= K");
  const size_t n = (size_t)128 * (K + 1);
  if (rows == 0) {
    AERO_CUDA(cudaMemsetAsync(dwb, 0, n * sizeof(float), st));
    return AERO_OK;
  }
  AERO_CHECK_ARG(g && x && workspace && workspace_bytes >= aero_thin_linear_workspace_bytes(rows, K),
                 "aero_thin_linear_bwd: null pointer or workspace too small");
  const int grid = tl_grid(rows);
  float* part = reinterpret_cast<float*>(workspace);
  int rc;
  if (dtype == AERO_F32) rc = tl_bwd<float>(g, x, ldx, part, grid, rows, K, st);
  else if (dtype == AERO_BF16) rc = tl_bwd<__nv_bfloat16>(g, x, ldx, part, grid, rows, K, st);
  else {
    set_error("aero_thin_linear_bwd: unsupported dtype %d", dtype);
    return AERO_EUNSUPPORTED;
  }
  if (rc) return rc;
  return launch_reduce_partials(part, grid, n, dwb, n, st);
}
