// block_simt.cu -- fused MGN block (forward and backward) on CUDA cores, fp32 math.
//
// This is the exact-fp32 path (north_star: "fp32 path <= 1e-5 relative per layer output") and the
// on-device cross-check of the tcgen05 path.  Storage of latent rows is templated (fp32 | bf16);
// all arithmetic, LayerNorm statistics and reductions are fp32.
//
// A CTA of 128 threads owns a tile of TM rows.  Activations live in shared memory as fp32
// [TM][132]; every Linear(128,128) is a register-tiled SGEMM (thread tile (TM/8) x 8) whose
// weight operand streams through a double-buffered 16x128 slab fed by cp.async.  Tiles are
// assigned round-robin to CTAs in increasing order, and every cross-row reduction (receiver
// sums, bias / LayerNorm / weight gradients) runs in a fixed order, so results are bit-stable
// from run to run.
#include "common.cuh"

namespace aero {

constexpr int LDS_ = 132;           // padded smem row stride (floats)
constexpr int SLAB_K = 16;          // weight slab depth
constexpr int SLAB_FLOATS = SLAB_K * 128;
constexpr int MAX_L = 4;

struct SimtArgs {
  int L, act, use_ln, main_f32, has_resid_grad;
  int64_t rows, n_nodes, ldp, poff0, poff1;
  const void* main;
  const float* main_scale;
  const void* resid;
  const void* P;
  const int32_t* idx0;
  const int32_t* idx1;
  const int32_t* rowptr;
  const float* prep;  // prepared weights
  void* out;
  float* agg;
  float* agg_part;    // [tiles][2][128]
  const void* g_out;
  const float* g_agg;
  void* g_main;
  void* g_h0;
  float* w_part;      // [grid][PackedLayout.total()]
};

__device__ __forceinline__ const float* prep_mat(const float* prep, int m, bool transposed) {
  return prep + (size_t)(2 * m + (transposed ? 1 : 0)) * 16384;
}
__device__ __forceinline__ const float* prep_vec(const float* prep, int L, int v) {
  return prep + (size_t)(L + 2) * 2 * 16384 + (size_t)v * 128;  // v: 0..L-1 hidden bias, L out bias, L+1 gamma, L+2 beta
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void load_slab(float* Ws, const float* __restrict__ Bk, int s) {
  // 16 x 128 floats = 512 x 16 B, 4 per thread
  const int tid = threadIdx.x;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int id = tid + q * 128;
    int r = id >> 5, c4 = id & 31;
    cp_async16(Ws + r * 128 + c4 * 4, Bk + (size_t)(s * SLAB_K + r) * 128 + c4 * 4);
  }
}

// acc[i][j] (+)= sum_k As[(i*8+tr)*LDS_ + k] * Bk[k*128 + tc*8 + j], k = 0..127
// Bk is a global [128][128] matrix indexed [k][n].  Must be called by all 128 threads.
template <int TM>
__device__ __forceinline__ void gemm128(const float* __restrict__ As, const float* __restrict__ Bk, float* Ws,
                                        float (&acc)[TM / 8][8]) {
  constexpr int RPT = TM / 8;
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  __syncthreads();  // As complete; slab buffers free
  load_slab(Ws, Bk, 0);
  cp_async_commit();
  for (int s = 0; s < 128 / SLAB_K; ++s) {
    if (s + 1 < 128 / SLAB_K) {
      load_slab(Ws + ((s + 1) & 1) * SLAB_FLOATS, Bk, s + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* W = Ws + (s & 1) * SLAB_FLOATS;
#pragma unroll
    for (int kk = 0; kk < SLAB_K; ++kk) {
      float4 b0 = *reinterpret_cast<const float4*>(W + kk * 128 + tc * 8);
      float4 b1 = *reinterpret_cast<const float4*>(W + kk * 128 + tc * 8 + 4);
      float a[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) a[i] = As[(i * 8 + tr) * LDS_ + s * SLAB_K + kk];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        acc[i][0] = fmaf(a[i], b0.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b0.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b0.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b0.w, acc[i][3]);
        acc[i][4] = fmaf(a[i], b1.x, acc[i][4]);
        acc[i][5] = fmaf(a[i], b1.y, acc[i][5]);
        acc[i][6] = fmaf(a[i], b1.z, acc[i][6]);
        acc[i][7] = fmaf(a[i], b1.w, acc[i][7]);
      }
    }
    __syncthreads();
  }
}

// dst[r][c] = f(acc[r][c] + bias[c]) for the thread's fragment
template <int TM, typename F>
__device__ __forceinline__ void store_frag(float* dst, const float (&acc)[TM / 8][8], F f) {
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
#pragma unroll
  for (int i = 0; i < TM / 8; ++i) {
    int r = i * 8 + tr;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = f(r, tc * 8 + j, acc[i][j]);
    *reinterpret_cast<float4*>(dst + r * LDS_ + tc * 8) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + r * LDS_ + tc * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <typename T>
__device__ __forceinline__ float4 load_main4(const SimtArgs& a, int64_t row, int c) {
  float4 v = a.main_f32 ? load4(reinterpret_cast<const float*>(a.main) + row * 128 + c)
                        : load4(reinterpret_cast<const T*>(a.main) + row * 128 + c);
  if (a.main_scale) {
    float s = a.main_scale[row];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
  }
  return v;
}

// stage `main` rows of the tile (zero padded) and the gather indices
template <typename T, int TM>
__device__ __forceinline__ void stage_tile(const SimtArgs& a, int64_t row0, int nrows, float* Mb, int* sidx0,
                                           int* sidx1) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  for (int r = w; r < TM; r += 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows) v = load_main4<T>(a, row0 + r, lane * 4);
    *reinterpret_cast<float4*>(Mb + r * LDS_ + lane * 4) = v;
  }
  if (tid < TM) {
    int64_t row = row0 + tid;
    bool ok = tid < nrows;
    sidx0[tid] = ok ? (a.idx0 ? a.idx0[row] : (int)row) : 0;
    sidx1[tid] = ok ? (a.idx1 ? a.idx1[row] : -1) : -1;
  }
}

// first layer epilogue: act(acc + P[idx0] (+ P[idx1]))
template <typename T, int TM>
__device__ __forceinline__ void first_layer_epilogue(const SimtArgs& a, const float (&acc)[TM / 8][8], float* dst,
                                                     const int* sidx0, const int* sidx1, int nrows) {
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  const T* P = reinterpret_cast<const T*>(a.P);
#pragma unroll
  for (int i = 0; i < TM / 8; ++i) {
    int r = i * 8 + tr;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (r < nrows) {
      const T* p0 = P + (int64_t)sidx0[r] * a.ldp + a.poff0 + tc * 8;
      float4 q0 = load4(p0), q1 = load4(p0 + 4);
      v[0] = acc[i][0] + q0.x; v[1] = acc[i][1] + q0.y; v[2] = acc[i][2] + q0.z; v[3] = acc[i][3] + q0.w;
      v[4] = acc[i][4] + q1.x; v[5] = acc[i][5] + q1.y; v[6] = acc[i][6] + q1.z; v[7] = acc[i][7] + q1.w;
      if (sidx1[r] >= 0) {
        const T* p1 = P + (int64_t)sidx1[r] * a.ldp + a.poff1 + tc * 8;
        float4 s0 = load4(p1), s1 = load4(p1 + 4);
        v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
        v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_fwd(v[j], a.act);
    }
    *reinterpret_cast<float4*>(dst + r * LDS_ + tc * 8) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + r * LDS_ + tc * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// =============================================================================================
// forward
// =============================================================================================
template <typename T, int TM>
__global__ void __launch_bounds__(128) simt_block_fwd_kernel(SimtArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;
  float* bufB = bufA + TM * LDS_;
  float* Ws = bufB + TM * LDS_;
  int* sidx0 = reinterpret_cast<int*>(Ws + 2 * SLAB_FLOATS);
  int* sidx1 = sidx0 + TM;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int L = a.L;
  const int64_t tiles = (a.rows + TM - 1) / TM;
  float acc[TM / 8][8];

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = t * TM;
    const int nrows = (int)((a.rows - row0) < TM ? (a.rows - row0) : TM);
    __syncthreads();  // previous tile fully consumed
    stage_tile<T, TM>(a, row0, nrows, bufA, sidx0, sidx1);
    // layer 0
    gemm128<TM>(bufA, prep_mat(a.prep, 0, true), Ws, acc);
    first_layer_epilogue<T, TM>(a, acc, bufB, sidx0, sidx1, nrows);
    float* cur = bufB;
    float* nxt = bufA;
    for (int l = 0; l < L; ++l) {
      gemm128<TM>(cur, prep_mat(a.prep, 1 + l, true), Ws, acc);
      const float* b = prep_vec(a.prep, L, l);
      const int act = a.act;
      store_frag<TM>(nxt, acc, [&](int, int c, float v) { return act_fwd(v + b[c], act); });
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    gemm128<TM>(cur, prep_mat(a.prep, 1 + L, true), Ws, acc);
    {
      const float* b = prep_vec(a.prep, L, L);
      store_frag<TM>(nxt, acc, [&](int, int c, float v) { return v + b[c]; });
    }
    float* Y = nxt;
    __syncthreads();
    // LayerNorm + residual, one warp per row
    {
      const float4 g4 = *reinterpret_cast<const float4*>(prep_vec(a.prep, L, L + 1) + lane * 4);
      const float4 b4 = *reinterpret_cast<const float4*>(prep_vec(a.prep, L, L + 2) + lane * 4);
      for (int r = w; r < nrows; r += 4) {
        float4 y = *reinterpret_cast<const float4*>(Y + r * LDS_ + lane * 4);
        float4 u = y;
        if (a.use_ln) {
          float mean = warp_sum(y.x + y.y + y.z + y.w) * (1.f / 128.f);
          float dx = y.x - mean, dy = y.y - mean, dz = y.z - mean, dw = y.w - mean;
          float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f);
          float rstd = rsqrtf(var + 1e-5f);
          u.x = dx * rstd * g4.x + b4.x;
          u.y = dy * rstd * g4.y + b4.y;
          u.z = dz * rstd * g4.z + b4.z;
          u.w = dw * rstd * g4.w + b4.w;
        }
        float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.resid) res = load4(reinterpret_cast<const T*>(a.resid) + (row0 + r) * 128 + lane * 4);
        float4 o = make_float4(res.x + u.x, res.y + u.y, res.z + u.z, res.w + u.w);
        T* op = reinterpret_cast<T*>(a.out) + (row0 + r) * 128 + lane * 4;
        store4(op, o);
        if (a.agg) {
          // aggregate what was stored (rounded to the storage type)
          float4 q = make_float4(round_to<T>(o.x), round_to<T>(o.y), round_to<T>(o.z), round_to<T>(o.w));
          *reinterpret_cast<float4*>(Y + r * LDS_ + lane * 4) = q;
        }
      }
    }
    if (a.agg) {
      __syncthreads();
      // receiver sums: thread = column, rows walked in CSR order
      const int c = tid;
      const int64_t tile_end = row0 + nrows;
      int r = 0;
      while (r < nrows) {
        int n = sidx1[r];
        int b = a.rowptr[n], e = a.rowptr[n + 1];
        int re = (int)(((int64_t)e < tile_end ? (int64_t)e : tile_end) - row0);
        float s = 0.f;
        for (int q = r; q < re; ++q) s += Y[q * LDS_ + c];
        bool complete = (b >= row0) && (e <= tile_end);
        if (complete) a.agg[(size_t)n * 128 + c] = s;
        else a.agg_part[((size_t)t * 2 + (r == 0 ? 0 : 1)) * 128 + c] = s;
        r = re;
      }
    }
  }
}

// =============================================================================================
// backward
// =============================================================================================
// part[m][n] += sum_r G[r][m] * H[r][n]   (128x128 output, K = TM rows), CTA-private read-modify-write
template <int TM>
__device__ __forceinline__ void dw_accumulate(const float* __restrict__ G, const float* __restrict__ H,
                                              float* __restrict__ part) {
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int m0 = half * 64 + tr * 8;
#pragma unroll 4
    for (int r = 0; r < TM; ++r) {
      float4 g0 = *reinterpret_cast<const float4*>(G + r * LDS_ + m0);
      float4 g1 = *reinterpret_cast<const float4*>(G + r * LDS_ + m0 + 4);
      float4 h0 = *reinterpret_cast<const float4*>(H + r * LDS_ + tc * 8);
      float4 h1 = *reinterpret_cast<const float4*>(H + r * LDS_ + tc * 8 + 4);
      float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(g[i], h[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4* p = reinterpret_cast<float4*>(part + (size_t)(m0 + i) * 128 + tc * 8);
      float4 o0 = p[0], o1 = p[1];
      o0.x += acc[i][0]; o0.y += acc[i][1]; o0.z += acc[i][2]; o0.w += acc[i][3];
      o1.x += acc[i][4]; o1.y += acc[i][5]; o1.z += acc[i][6]; o1.w += acc[i][7];
      p[0] = o0; p[1] = o1;
    }
  }
}

template <typename T, int TM>
__global__ void __launch_bounds__(128) simt_block_bwd_kernel(SimtArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int L = a.L;
  float* Mb = smem;
  float* H[MAX_L + 1];
#pragma unroll
  for (int l = 0; l <= MAX_L; ++l) H[l] = Mb + (size_t)(1 + (l <= L ? l : 0)) * TM * LDS_;
  float* G0 = Mb + (size_t)(L + 2) * TM * LDS_;
  float* G1 = G0 + TM * LDS_;
  float* Ws = G1 + TM * LDS_;
  int* sidx0 = reinterpret_cast<int*>(Ws + 2 * SLAB_FLOATS);
  int* sidx1 = sidx0 + TM;
  float* red = reinterpret_cast<float*>(sidx1 + TM);  // [4][128] cross-warp scratch

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t tiles = (a.rows + TM - 1) / TM;
  const PackedLayout pl{L};
  float* part = a.w_part + (size_t)blockIdx.x * pl.total();
  float acc[TM / 8][8];
  float db[MAX_L + 1];
#pragma unroll
  for (int l = 0; l <= MAX_L; ++l) db[l] = 0.f;
  float dgam[4] = {0.f, 0.f, 0.f, 0.f}, dbet[4] = {0.f, 0.f, 0.f, 0.f};
  float db0 = 0.f;
  const float4 gam4 = *reinterpret_cast<const float4*>(prep_vec(a.prep, L, L + 1) + lane * 4);
  const int act = a.act;

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t row0 = t * TM;
    const int nrows = (int)((a.rows - row0) < TM ? (a.rows - row0) : TM);
    __syncthreads();
    stage_tile<T, TM>(a, row0, nrows, Mb, sidx0, sidx1);
    // ---- recompute forward ----
    gemm128<TM>(Mb, prep_mat(a.prep, 0, true), Ws, acc);
    first_layer_epilogue<T, TM>(a, acc, H[0], sidx0, sidx1, nrows);
    for (int l = 0; l < L; ++l) {
      gemm128<TM>(H[l], prep_mat(a.prep, 1 + l, true), Ws, acc);
      const float* b = prep_vec(a.prep, L, l);
      store_frag<TM>(H[l + 1], acc, [&](int r, int c, float v) { return r < nrows ? act_fwd(v + b[c], act) : 0.f; });
    }
    gemm128<TM>(H[L], prep_mat(a.prep, 1 + L, true), Ws, acc);
    {
      const float* b = prep_vec(a.prep, L, L);
      store_frag<TM>(G1, acc, [&](int, int c, float v) { return v + b[c]; });
    }
    __syncthreads();
    // ---- LayerNorm backward, one warp per row; G0 <- dL/dy ----
    for (int r = w; r < TM; r += 4) {
      float4 gy = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) {
        float4 go = load4(reinterpret_cast<const T*>(a.g_out) + (row0 + r) * 128 + lane * 4);
        if (a.g_agg) {
          float4 ga = *reinterpret_cast<const float4*>(a.g_agg + (size_t)sidx1[r] * 128 + lane * 4);
          go.x += ga.x; go.y += ga.y; go.z += ga.z; go.w += ga.w;
        }
        if (a.use_ln) {
          float4 y = *reinterpret_cast<const float4*>(G1 + r * LDS_ + lane * 4);
          float mean = warp_sum(y.x + y.y + y.z + y.w) * (1.f / 128.f);
          float dx = y.x - mean, dy = y.y - mean, dz = y.z - mean, dw = y.w - mean;
          float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f);
          float rstd = rsqrtf(var + 1e-5f);
          float hx = dx * rstd, hy = dy * rstd, hz = dz * rstd, hw = dw * rstd;
          dgam[0] += go.x * hx; dgam[1] += go.y * hy; dgam[2] += go.z * hz; dgam[3] += go.w * hw;
          dbet[0] += go.x; dbet[1] += go.y; dbet[2] += go.z; dbet[3] += go.w;
          float gx = go.x * gam4.x, gyy = go.y * gam4.y, gz = go.z * gam4.z, gw = go.w * gam4.w;
          float m1 = warp_sum(gx + gyy + gz + gw) * (1.f / 128.f);
          float m2 = warp_sum(gx * hx + gyy * hy + gz * hz + gw * hw) * (1.f / 128.f);
          gy.x = rstd * (gx - m1 - hx * m2);
          gy.y = rstd * (gyy - m1 - hy * m2);
          gy.z = rstd * (gz - m1 - hz * m2);
          gy.w = rstd * (gw - m1 - hw * m2);
        } else {
          gy = go;
        }
      }
      *reinterpret_cast<float4*>(G0 + r * LDS_ + lane * 4) = gy;
    }
    __syncthreads();
    // ---- output layer and hidden layers, last to first ----
    float* Gc = G0;
    float* Gn = G1;
    for (int l = L; l >= 0; --l) {
      __syncthreads();  // Gc complete
      // layer index l here: l == L is the output Linear (input H[L]); l < L is hidden Linear l (input H[l])
      // weight matrix id: out -> 1+L, hidden l -> 1+l ; both read H[l] as input
      const int mat = 1 + l;
      const size_t woff = (l == L) ? pl.w_out() : pl.w_hidden(l);
      dw_accumulate<TM>(Gc, H[l], part + woff);
      {
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s += Gc[r * LDS_ + tid];
        db[l] += s;
      }
      gemm128<TM>(Gc, prep_mat(a.prep, mat, false), Ws, acc);
      const float* Hin = H[l];
      store_frag<TM>(Gn, acc, [&](int r, int c, float v) { return v * act_grad_from_out(Hin[r * LDS_ + c], act); });
      float* tmp = Gc; Gc = Gn; Gn = tmp;
    }
    __syncthreads();
    // Gc = dL/d(pre-activation of h0)
    {
      float s0 = 0.f;
      for (int r = 0; r < TM; ++r) s0 += Gc[r * LDS_ + tid];
      db0 += s0;
      T* gh = reinterpret_cast<T*>(a.g_h0);
      for (int r = w; r < nrows; r += 4) {
        float4 v = *reinterpret_cast<const float4*>(Gc + r * LDS_ + lane * 4);
        store4(gh + (row0 + r) * 128 + lane * 4, v);
      }
    }
    gemm128<TM>(Gc, prep_mat(a.prep, 0, false), Ws, acc);
    store_frag<TM>(Gn, acc, [&](int, int, float v) { return v; });
    __syncthreads();
    for (int r = w; r < nrows; r += 4) {
      float4 v = *reinterpret_cast<const float4*>(Gn + r * LDS_ + lane * 4);
      if (a.main_scale) {
        float s = a.main_scale[row0 + r];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      }
      if (a.has_resid_grad) {
        float4 go = load4(reinterpret_cast<const T*>(a.g_out) + (row0 + r) * 128 + lane * 4);
        if (a.g_agg) {
          float4 ga = *reinterpret_cast<const float4*>(a.g_agg + (size_t)sidx1[r] * 128 + lane * 4);
          go.x += ga.x; go.y += ga.y; go.z += ga.z; go.w += ga.w;
        }
        v.x += go.x; v.y += go.y; v.z += go.z; v.w += go.w;
      }
      if (a.main_f32) store4(reinterpret_cast<float*>(a.g_main) + (row0 + r) * 128 + lane * 4, v);
      else store4(reinterpret_cast<T*>(a.g_main) + (row0 + r) * 128 + lane * 4, v);
    }
  }
  // ---- flush per-CTA vector partials ----
  __syncthreads();
  for (int l = 0; l < L; ++l) part[pl.b_hidden(l) + tid] = db[l];
  part[pl.b_out() + tid] = db[L];
  part[pl.bias0() + tid] = db0;
#pragma unroll
  for (int j = 0; j < 4; ++j) red[w * 128 + lane * 4 + j] = dgam[j];
  __syncthreads();
  part[pl.gamma() + tid] = red[tid] + red[128 + tid] + red[256 + tid] + red[384 + tid];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) red[w * 128 + lane * 4 + j] = dbet[j];
  __syncthreads();
  part[pl.beta() + tid] = red[tid] + red[128 + tid] + red[256 + tid] + red[384 + tid];
}

// =============================================================================================
// weight preparation: [W | W^T] per matrix, then the vectors
// =============================================================================================
__global__ void simt_prepare_kernel(const float* __restrict__ w, int L, float* __restrict__ prep) {
  const PackedLayout pl{L};
  const int nm = L + 2;
  size_t total = (size_t)nm * 16384;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int m = (int)(i / 16384);
    int o = (int)((i % 16384) / 128), in = (int)(i % 128);
    float v = w[(size_t)m * 16384 + (size_t)o * 128 + in];
    prep[(size_t)(2 * m) * 16384 + (size_t)o * 128 + in] = v;
    prep[(size_t)(2 * m + 1) * 16384 + (size_t)in * 128 + o] = v;
  }
  size_t nv = (size_t)(L + 3) * 128;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x)
    prep[(size_t)nm * 2 * 16384 + i] = w[pl.b_hidden(0) + i];
}

// ---- host side ------------------------------------------------------------------------------
constexpr int FWD_TM = 64;
constexpr int BWD_TM = 32;

static size_t fwd_smem_bytes() { return (size_t)(2 * FWD_TM * LDS_ + 2 * SLAB_FLOATS) * 4 + 2 * FWD_TM * 4; }
static size_t bwd_smem_bytes(int L) {
  return (size_t)((L + 4) * BWD_TM * LDS_ + 2 * SLAB_FLOATS) * 4 + 2 * BWD_TM * 4 + 4 * 128 * 4;
}
static int fwd_grid(int64_t rows) {
  int64_t tiles = cdiv(rows, FWD_TM);
  int64_t g = 2LL * sm_count();
  return (int)(tiles < g ? tiles : g);
}
static int bwd_grid(int64_t rows) {
  int64_t tiles = cdiv(rows, BWD_TM);
  int64_t g = sm_count();
  return (int)(tiles < g ? tiles : g);
}

size_t simt_prepared_bytes(int L) { return ((size_t)(L + 2) * 2 * 16384 + (size_t)(L + 3) * 128) * sizeof(float); }

int simt_prepare(const float* w, int L, void* prepared, cudaStream_t st) {
  simt_prepare_kernel<<<64, 256, 0, st>>>(w, L, reinterpret_cast<float*>(prepared));
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

size_t simt_block_workspace_bytes(const aero_block_desc* d, int backward) {
  if (!backward) {
    int64_t tiles = cdiv(d->rows > 0 ? d->rows : 1, FWD_TM);
    return d->agg ? align_up((size_t)tiles * 2 * 128 * sizeof(float), 256) : 256;
  }
  PackedLayout pl{d->L};
  return align_up((size_t)bwd_grid(d->rows > 0 ? d->rows : 1) * pl.total() * sizeof(float), 256);
}

static SimtArgs make_args(const aero_block_desc* d) {
  SimtArgs a;
  a.L = d->L; a.act = d->act; a.use_ln = d->use_ln; a.main_f32 = d->main_f32; a.has_resid_grad = d->has_resid_grad;
  a.rows = d->rows; a.n_nodes = d->n_nodes; a.ldp = d->ldp; a.poff0 = d->poff0; a.poff1 = d->poff1;
  a.main = d->main; a.main_scale = d->main_scale; a.resid = d->resid; a.P = d->P;
  a.idx0 = d->idx0; a.idx1 = d->idx1; a.rowptr = d->rowptr;
  a.prep = reinterpret_cast<const float*>(d->prepared);
  a.out = d->out; a.agg = d->agg; a.agg_part = nullptr;
  a.g_out = d->g_out; a.g_agg = d->g_agg; a.g_main = d->g_main; a.g_h0 = d->g_h0; a.w_part = nullptr;
  return a;
}

int simt_block_fwd(const aero_block_desc* d, cudaStream_t st) {
  if (d->L > MAX_L) {
    set_error("simt_block_fwd: L=%d > %d", d->L, MAX_L);
    return AERO_EUNSUPPORTED;
  }
  SimtArgs a = make_args(d);
  if (d->agg) {
    a.agg_part = reinterpret_cast<float*>(d->workspace);
    if (!(d->flags & AERO_BLOCK_AGG_NO_CLEAR))
      AERO_CUDA(cudaMemsetAsync(d->agg, 0, (size_t)d->n_nodes * 128 * sizeof(float), st));
  }
  if (d->rows == 0) return AERO_OK;
  size_t smem = fwd_smem_bytes();
  int grid = fwd_grid(d->rows);
  if (d->dtype == AERO_F32) {
    AERO_CUDA(cudaFuncSetAttribute(simt_block_fwd_kernel<float, FWD_TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    simt_block_fwd_kernel<float, FWD_TM><<<grid, 128, smem, st>>>(a);
  } else {
    AERO_CUDA(cudaFuncSetAttribute(simt_block_fwd_kernel<__nv_bfloat16, FWD_TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    simt_block_fwd_kernel<__nv_bfloat16, FWD_TM><<<grid, 128, smem, st>>>(a);
  }
  AERO_LAUNCH_CHECK();
  if (d->agg) return launch_agg_fixup(a.agg_part, d->rowptr, d->agg, d->rows, d->n_nodes, FWD_TM, d->idx1, st);
  return AERO_OK;
}

int simt_block_bwd(const aero_block_desc* d, cudaStream_t st) {
  if (d->L > MAX_L) {
    set_error("simt_block_bwd: L=%d > %d", d->L, MAX_L);
    return AERO_EUNSUPPORTED;
  }
  SimtArgs a = make_args(d);
  PackedLayout pl{d->L};
  int grid = bwd_grid(d->rows > 0 ? d->rows : 1);
  a.w_part = reinterpret_cast<float*>(d->workspace);
  size_t part_bytes = (size_t)grid * pl.total() * sizeof(float);
  AERO_CUDA(cudaMemsetAsync(a.w_part, 0, part_bytes, st));
  if (d->rows > 0) {
    size_t smem = bwd_smem_bytes(d->L);
    if (d->dtype == AERO_F32) {
      AERO_CUDA(cudaFuncSetAttribute(simt_block_bwd_kernel<float, BWD_TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      simt_block_bwd_kernel<float, BWD_TM><<<grid, 128, smem, st>>>(a);
    } else {
      AERO_CUDA(cudaFuncSetAttribute(simt_block_bwd_kernel<__nv_bfloat16, BWD_TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      simt_block_bwd_kernel<__nv_bfloat16, BWD_TM><<<grid, 128, smem, st>>>(a);
    }
    AERO_LAUNCH_CHECK();
  }
  // g_w[W_1 .. beta] = sum over CTAs (W_main slot untouched)
  size_t off = pl.w_hidden(0);
  return launch_reduce_partials(a.w_part + off, grid, pl.total(), d->g_w + off, pl.total() - off, st);
}

}  // namespace aero
