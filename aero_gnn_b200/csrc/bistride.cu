// bistride.cu -- BFS-bistride pooling index kernels and the WeightedEdgeConv message passing of the reference's
// bistride_ops module (models/__pycache__/bistride_ops.cpython-311.pyc: BistridePooling :13-94, Unpool :96,
// WeightedEdgeConv :131-209) and of MultiScaleGraphPreprocessor.create_multiscale_graph
// (models/__pycache__/bsms_mgn.cpython-311.pyc :32).
//
// Integer kernels (BFS levels, even-level selection, coarse edge filter) are bit-exact: BFS distances do not depend
// on the visiting order, selections and surviving edges keep ascending / caller order through prefix sums.
// The WeightedEdgeConv kernels are HBM-bound gathers: one warp owns one receiver (forward, backward pass 1) or one
// sender (backward pass 2), walks its CSR segment in a fixed order and accumulates in fp32 registers -- no atomics.
#include "common.cuh"

namespace aero {

// =============================================================================================
// BFS levels (frontier queues, one launch per level; distances are order-independent)
// =============================================================================================
__global__ void bfs_init_kernel(int64_t* __restrict__ dist, int64_t N, int64_t start, int32_t* __restrict__ q0,
                                int32_t* __restrict__ cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) dist[i] = (i == start) ? 0 : -1;
  if (i == 0) {
    q0[0] = (int32_t)start;
    cnt[0] = 1;
    cnt[1] = 0;
    cnt[2] = 0;
  }
}

// level `lv`: frontier = q_in[0 .. cnt[lv % 3]); discovered nodes are appended to q_out, counted in cnt[(lv+1) % 3];
// cnt[(lv+2) % 3] (the output counter of the next launch) is cleared here.
__global__ void __launch_bounds__(256) bfs_level_kernel(const int32_t* __restrict__ sptr, const int32_t* __restrict__ sperm,
                                                        const int32_t* __restrict__ dst, int64_t* __restrict__ dist,
                                                        const int32_t* __restrict__ q_in, int32_t* __restrict__ q_out,
                                                        int32_t* __restrict__ cnt, int lv) {
  const int n = cnt[lv % 3];
  int32_t* out_cnt = cnt + (lv + 1) % 3;
  if (blockIdx.x == 0 && threadIdx.x == 0) cnt[(lv + 2) % 3] = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int u = q_in[i];
    const int b = sptr[u], e = sptr[u + 1];
    for (int j = b; j < e; ++j) {
      const int v = dst[sperm[j]];
      unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(dist + v), ~0ull,
                                         (unsigned long long)(lv + 1));
      if (old == ~0ull) q_out[atomicAdd(out_cnt, 1)] = v;
    }
  }
}

__global__ void bfs_status_kernel(const int32_t* __restrict__ cnt, int lv_next, int64_t* __restrict__ status) {
  status[0] = cnt[lv_next % 3];   // frontier still to expand
  status[1] = lv_next;
}

// =============================================================================================
// even-level selection with the 30 % fallback, index map
// =============================================================================================
__global__ void select_flags_kernel(const int64_t* __restrict__ dist, int64_t N, int32_t* __restrict__ f_even,
                                    int32_t* __restrict__ f_reach) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t d = dist[i];
  f_reach[i] = d >= 0;
  f_even[i] = (d >= 0) && ((d & 1) == 0);
}

__global__ void select_scatter_kernel(const int64_t* __restrict__ dist, int64_t N, const int32_t* __restrict__ s_even,
                                      const int32_t* __restrict__ s_reach, int64_t* __restrict__ selected,
                                      int64_t* __restrict__ index_map, int64_t* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t n_even = s_even[N - 1];
  // Python: len(selected) < num_nodes * 0.3  (int compared with a double product)
  const bool fallback = (double)n_even < (double)N * 0.3;
  const int64_t d = dist[i];
  const bool keep = fallback ? (d >= 0) : ((d >= 0) && ((d & 1) == 0));
  const int32_t* sc = fallback ? s_reach : s_even;
  if (keep) {
    const int64_t k = (int64_t)sc[i] - 1;
    selected[k] = i;
    index_map[i] = k;
  } else {
    index_map[i] = -1;
  }
  if (i == 0) {
    counts[0] = sc[N - 1];
    counts[1] = fallback ? 1 : 0;
  }
}

// =============================================================================================
// coarse edges: both endpoints selected, remapped, self-loops dropped, caller order kept
// =============================================================================================
__global__ void filter_flags_kernel(const int64_t* __restrict__ ei, int64_t E, const int64_t* __restrict__ index_map,
                                    int64_t N, int32_t* __restrict__ flag, unsigned long long* __restrict__ bad) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  const int64_t s = ei[k], d = ei[E + k];
  if (s < 0 || s >= N || d < 0 || d >= N) {
    atomicAdd(bad, 1ull);
    flag[k] = 0;
    return;
  }
  const int64_t ms = index_map[s], md = index_map[d];
  flag[k] = (ms >= 0 && md >= 0 && ms != md) ? 1 : 0;
}

__global__ void filter_scatter_kernel(const int64_t* __restrict__ ei, int64_t E, const int64_t* __restrict__ index_map,
                                      const int32_t* __restrict__ flag, const int32_t* __restrict__ scan,
                                      int64_t* __restrict__ out, int32_t* __restrict__ kept, int64_t* __restrict__ counts) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  if (flag[k]) {
    const int64_t j = (int64_t)scan[k] - 1;
    out[j] = index_map[ei[k]];
    out[E + j] = index_map[ei[E + k]];
    if (kept) kept[j] = (int32_t)k;
  }
  if (k == 0) counts[0] = scan[E - 1];
}

// =============================================================================================
// WeightedEdgeConv
// =============================================================================================
template <typename T>
__device__ __forceinline__ float2 load2(const T* p);
template <>
__device__ __forceinline__ float2 load2<float>(const float* p) {
  return *reinterpret_cast<const float2*>(p);
}
template <>
__device__ __forceinline__ float2 load2<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint32_t v = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void store2(T* p, float2 v);
template <>
__device__ __forceinline__ void store2<float>(float* p, float2 v) {
  *reinterpret_cast<float2*>(p) = v;
}
template <>
__device__ __forceinline__ void store2<__nv_bfloat16>(__nv_bfloat16* p, float2 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&a);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;   // identical on every lane (butterfly)
}

constexpr int WEC_HID = 64;        // hidden width of edge_weight_mlp (bistride_ops pyc :136, const 64)
constexpr int WEC_SMALL = 2 * WEC_HID + 1;   // d(w1_len)[64] | d(W2)[64] | d(b2)
constexpr int WEC_PART_LD = 132;
constexpr int WEC_THREADS = 256;

struct WecArgs {
  int64_t N, E, ldq;
  int out_dim, toff, pos_dim, mean;
  const void* Q;
  const float* pos;
  const float* w1_len;
  const float* w2;
  const float* b2;
  const int32_t *rowptr, *src, *dst, *perm, *sptr, *sperm;
  void* w;
  void* out;
  const void* g_out;
  const void* g_w_ext;
  void* dQ;
  void* g_w;
  float* ds;      // [E] fp32, CSR slot order
  float* part;    // [grid, WEC_PART_LD]
};

__device__ __forceinline__ float edge_len(const float* __restrict__ pos, int pos_dim, int64_t a, int64_t b) {
  float s = 0.f;
  for (int d = 0; d < pos_dim; ++d) {
    float v = pos[b * pos_dim + d] - pos[a * pos_dim + d];
    s = fmaf(v, v, s);
  }
  return sqrtf(s);
}

// All three kernels walk a CSR segment in chunks of 32 slots: every lane first fetches the indices, the edge length,
// the stored weight ... of ITS slot of the chunk (coalesced index loads, one gather latency for the whole chunk), then
// the warp processes the slots in order, WEC_U at a time: the row gathers of WEC_U edges are issued together before
// any of them is consumed, so a warp keeps WEC_U rows in flight instead of one (the kernels are latency-bound
// gathers; the accumulation order stays slot by slot, i.e. deterministic).
constexpr int WEC_U = 4;

// logistic function with the hardware exp2 / reciprocal approximations (relative error ~1e-6, far inside the fp32
// tolerance of 1e-5); the forward and the backward recompute use the same expression, so they agree bit for bit
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& t) {
  acc.x = fmaf(w, t.x, acc.x); acc.y = fmaf(w, t.y, acc.y); acc.z = fmaf(w, t.z, acc.z); acc.w = fmaf(w, t.w, acc.w);
}

// edge weights, one LANE per edge (CSR slot k): w_k = sigmoid(W2 . relu(A[src_k] + B[dst_k] + w1_len * len_k) + b2).
// The 64-wide dot product runs inside one thread (2 x 16 vector loads of the two half rows, 64 x 4 arithmetic
// instructions) instead of being spread over a warp with a 5-step shuffle reduction per edge: ~9 warp instructions
// per edge instead of ~25.  Consecutive slots share their receiver, so the B half-row loads of neighbouring lanes
// coalesce; the A half rows are L2 / L1 gathers.  The weight is rounded to the storage dtype and stored in caller
// order, exactly what the fused forward below stores when it computes the weights itself.
template <typename T>
__global__ void __launch_bounds__(256) wec_weight_kernel(WecArgs a) {
  __shared__ float s_wl[WEC_HID], s_w2[WEC_HID];
  if (threadIdx.x < WEC_HID) {
    s_wl[threadIdx.x] = a.w1_len[threadIdx.x];
    s_w2[threadIdx.x] = a.w2[threadIdx.x];
  }
  __syncthreads();
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= a.E) return;
  const T* __restrict__ Q = reinterpret_cast<const T*>(a.Q);
  const int64_t s = a.src[k], n = a.dst[k];
  const float len = edge_len(a.pos, a.pos_dim, s, n);
  const T* __restrict__ A = Q + s * a.ldq;
  const T* __restrict__ B = Q + n * a.ldq + WEC_HID;
  float sc = 0.f;
#pragma unroll 4
  for (int j = 0; j < WEC_HID; j += 4) {
    const float4 av = load4(A + j), bv = load4(B + j);
    sc = fmaf(fmaxf(av.x + bv.x + s_wl[j] * len, 0.f), s_w2[j], sc);
    sc = fmaf(fmaxf(av.y + bv.y + s_wl[j + 1] * len, 0.f), s_w2[j + 1], sc);
    sc = fmaf(fmaxf(av.z + bv.z + s_wl[j + 2] * len, 0.f), s_w2[j + 2], sc);
    sc = fmaf(fmaxf(av.w + bv.w + s_wl[j + 3] * len, 0.f), s_w2[j + 3], sc);
  }
  store1(reinterpret_cast<T*>(a.w) + a.perm[k], round_to<T>(sigmoid_fast(sc + a.b2[0])));
}

// forward: out[n] = sum_k w_k * T[src_k] over the receiver's CSR segment; w_k from the edge-weight MLP (COMPUTE) or given
template <typename T, bool COMPUTE>
__global__ void __launch_bounds__(WEC_THREADS) wec_fwd_kernel(WecArgs a) {
  // lane-0 broadcasts tell the compiler these values are warp-uniform: the loops below then need no per-shuffle
  // reconvergence barriers
  const int64_t n = (int64_t)blockIdx.x * (WEC_THREADS / 32) + __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (n >= a.N) return;
  const int lane = threadIdx.x & 31;
  const T* __restrict__ Q = reinterpret_cast<const T*>(a.Q);
  T* __restrict__ w = reinterpret_cast<T*>(a.w);
  const int b = __shfl_sync(0xffffffffu, a.rowptr[n], 0), e = __shfl_sync(0xffffffffu, a.rowptr[n + 1], 0);
  const int c = lane * 4;
  const bool colok = c < a.out_dim;
  float2 Bn = make_float2(0.f, 0.f), wl = Bn, w2v = Bn;
  float b2v = 0.f;
  if (COMPUTE) {
    Bn = load2(Q + n * a.ldq + WEC_HID + 2 * lane);
    wl = *reinterpret_cast<const float2*>(a.w1_len + 2 * lane);
    w2v = *reinterpret_cast<const float2*>(a.w2 + 2 * lane);
    b2v = a.b2[0];
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = b; base < e; base += 32) {
    const int cnt = e - base < 32 ? e - base : 32;
    int my_s = 0, my_p = 0;
    float my_x = 0.f;   // COMPUTE: edge length of my slot; otherwise: the given weight of my slot
    if (lane < cnt) {
      my_s = a.src[base + lane];
      my_p = a.perm[base + lane];
      my_x = COMPUTE ? edge_len(a.pos, a.pos_dim, my_s, n) : load1(w + my_p);
    }
    float my_w = my_x;
    for (int j0 = 0; j0 < cnt; j0 += WEC_U) {
      float4 t[WEC_U];
      float2 As[WEC_U];
#pragma unroll
      for (int u = 0; u < WEC_U; ++u) {
        const int64_t s = __shfl_sync(0xffffffffu, my_s, (j0 + u) & 31);
        t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        As[u] = make_float2(0.f, 0.f);
        if (j0 + u < cnt) {
          if (colok) t[u] = load4(Q + s * a.ldq + a.toff + c);
          if (COMPUTE) As[u] = load2(Q + s * a.ldq + 2 * lane);
        }
      }
#pragma unroll
      for (int u = 0; u < WEC_U; ++u) {
        // slots past the end of the chunk carry zero rows (t = As = 0): they add nothing, so no branch is needed and
        // the shuffles below stay in straight-line code
        const float xv = __shfl_sync(0xffffffffu, my_x, (j0 + u) & 31);
        float wgt = xv;
        if (COMPUTE) {
          const float h0 = fmaxf(As[u].x + Bn.x + wl.x * xv, 0.f);
          const float h1 = fmaxf(As[u].y + Bn.y + wl.y * xv, 0.f);
          const float sc = warp_sum(h0 * w2v.x + h1 * w2v.y) + b2v;
          wgt = round_to<T>(sigmoid_fast(sc));
          if (lane == j0 + u) my_w = wgt;
        }
        fma4(acc, wgt, t[u]);
      }
    }
    if (COMPUTE && lane < cnt) store1(w + my_p, my_w);
  }
  if (a.mean) {
    const float cnt = (float)(e - b > 1 ? e - b : 1);
    acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt;
  }
  if (colok) store4(reinterpret_cast<T*>(a.out) + n * a.out_dim + c, acc);
}

// backward pass 1 (receiver side): dw_k = <T[src_k], g_out[n]> (+ external gradient of the returned weights);
//   COMPUTE : ds_k = dw_k w_k (1 - w_k) kept per CSR slot, dB[n] = sum_k dz_k, partial sums of d(w1_len), d(W2), d(b2)
//   !COMPUTE: g_w[perm_k] = dw_k (gradient of the given edge weights)
template <typename T, bool COMPUTE>
__global__ void __launch_bounds__(WEC_THREADS) wec_bwd_recv_kernel(WecArgs a) {
  __shared__ float red[WEC_THREADS / 32][WEC_PART_LD];
  const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const T* __restrict__ Q = reinterpret_cast<const T*>(a.Q);
  const T* __restrict__ G = reinterpret_cast<const T*>(a.g_out);
  const T* __restrict__ gwx = reinterpret_cast<const T*>(a.g_w_ext);
  const int c = lane * 4;
  const bool colok = c < a.out_dim;
  float2 wl = make_float2(0.f, 0.f), w2v = wl;
  float b2v = 0.f;
  if (COMPUTE) {
    wl = *reinterpret_cast<const float2*>(a.w1_len + 2 * lane);
    w2v = *reinterpret_cast<const float2*>(a.w2 + 2 * lane);
    b2v = a.b2[0];
  }
  float2 dwl = make_float2(0.f, 0.f), dw2 = dwl;
  float db2 = 0.f;
  const int64_t warps = (int64_t)gridDim.x * (WEC_THREADS / 32);
  for (int64_t n = (int64_t)blockIdx.x * (WEC_THREADS / 32) + wid; n < a.N; n += warps) {
    const int b = __shfl_sync(0xffffffffu, a.rowptr[n], 0), e = __shfl_sync(0xffffffffu, a.rowptr[n + 1], 0);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (colok) g = load4(G + n * a.out_dim + c);
    if (a.mean) {
      const float cnt = (float)(e - b > 1 ? e - b : 1);
      g.x /= cnt; g.y /= cnt; g.z /= cnt; g.w /= cnt;
    }
    float2 Bn = make_float2(0.f, 0.f), dB = Bn;
    if (COMPUTE) Bn = load2(Q + n * a.ldq + WEC_HID + 2 * lane);
    for (int base = b; base < e; base += 32) {
      const int cnt = e - base < 32 ? e - base : 32;
      int my_s = 0, my_p = 0;
      float my_len = 0.f, my_gw = 0.f, my_out = 0.f;
      if (lane < cnt) {
        my_s = a.src[base + lane];
        my_p = a.perm[base + lane];
        if (COMPUTE) my_len = edge_len(a.pos, a.pos_dim, my_s, n);
        if (gwx) my_gw = load1(gwx + my_p);
      }
      for (int j0 = 0; j0 < cnt; j0 += WEC_U) {
        float4 t[WEC_U];
        float2 As[WEC_U];
#pragma unroll
        for (int u = 0; u < WEC_U; ++u) {
          const int64_t s = __shfl_sync(0xffffffffu, my_s, (j0 + u) & 31);
          t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          As[u] = make_float2(0.f, 0.f);
          if (j0 + u < cnt) {
            if (colok) t[u] = load4(Q + s * a.ldq + a.toff + c);
            if (COMPUTE) As[u] = load2(Q + s * a.ldq + 2 * lane);
          }
        }
#pragma unroll
        for (int u = 0; u < WEC_U; ++u) {
          // slots past the end of the chunk: t = As = 0 and gw = 0 give dw = ds = 0, every sum below gains 0
          const float len = __shfl_sync(0xffffffffu, my_len, (j0 + u) & 31);
          const float gw = __shfl_sync(0xffffffffu, my_gw, (j0 + u) & 31);
          const float dw = warp_sum(t[u].x * g.x + t[u].y * g.y + t[u].z * g.z + t[u].w * g.w) + gw;
          if (COMPUTE) {
            const float h0 = fmaxf(As[u].x + Bn.x + wl.x * len, 0.f);
            const float h1 = fmaxf(As[u].y + Bn.y + wl.y * len, 0.f);
            const float sc = warp_sum(h0 * w2v.x + h1 * w2v.y) + b2v;
            const float wgt = sigmoid_fast(sc);
            const float ds = dw * wgt * (1.f - wgt);
            if (lane == j0 + u) my_out = ds;
            const float dz0 = h0 > 0.f ? ds * w2v.x : 0.f;
            const float dz1 = h1 > 0.f ? ds * w2v.y : 0.f;
            dB.x += dz0; dB.y += dz1;
            dwl.x = fmaf(dz0, len, dwl.x); dwl.y = fmaf(dz1, len, dwl.y);
            dw2.x = fmaf(ds, h0, dw2.x); dw2.y = fmaf(ds, h1, dw2.y);
            db2 += ds;
          } else {
            if (lane == j0 + u) my_out = dw;
          }
        }
      }
      if (lane < cnt) {
        if (COMPUTE) a.ds[base + lane] = my_out;                       // coalesced, CSR slot order
        else store1(reinterpret_cast<T*>(a.g_w) + my_p, my_out);
      }
    }
    if (COMPUTE) store2(reinterpret_cast<T*>(a.dQ) + n * a.ldq + WEC_HID + 2 * lane, dB);
  }
  if (COMPUTE) {
    red[wid][2 * lane] = dwl.x;
    red[wid][2 * lane + 1] = dwl.y;
    red[wid][WEC_HID + 2 * lane] = dw2.x;
    red[wid][WEC_HID + 2 * lane + 1] = dw2.y;
    if (lane == 0) red[wid][2 * WEC_HID] = db2;
    __syncthreads();
    if (threadIdx.x < WEC_SMALL) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < WEC_THREADS / 32; ++q) s += red[q][threadIdx.x];
      a.part[(size_t)blockIdx.x * WEC_PART_LD + threadIdx.x] = s;
    }
  }
}

// backward pass 2 (sender side): dT[s] = sum_k w_k g_out[dst_k];  COMPUTE: dA[s] = sum_k dz_k (dz recomputed from ds_k)
template <typename T, bool COMPUTE>
__global__ void __launch_bounds__(WEC_THREADS) wec_bwd_send_kernel(WecArgs a) {
  const int64_t s = (int64_t)blockIdx.x * (WEC_THREADS / 32) + __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (s >= a.N) return;
  const int lane = threadIdx.x & 31;
  const T* __restrict__ Q = reinterpret_cast<const T*>(a.Q);
  const T* __restrict__ G = reinterpret_cast<const T*>(a.g_out);
  const T* __restrict__ w = reinterpret_cast<const T*>(a.w);
  T* __restrict__ dQ = reinterpret_cast<T*>(a.dQ);
  const int c = lane * 4;
  const bool colok = c < a.out_dim;
  float2 As = make_float2(0.f, 0.f), wl = As, w2v = As, dA = As;
  if (COMPUTE) {
    As = load2(Q + s * a.ldq + 2 * lane);
    wl = *reinterpret_cast<const float2*>(a.w1_len + 2 * lane);
    w2v = *reinterpret_cast<const float2*>(a.w2 + 2 * lane);
  }
  float4 dT = make_float4(0.f, 0.f, 0.f, 0.f);
  const int b = __shfl_sync(0xffffffffu, a.sptr[s], 0), e = __shfl_sync(0xffffffffu, a.sptr[s + 1], 0);
  for (int base = b; base < e; base += 32) {
    const int cnt = e - base < 32 ? e - base : 32;
    int my_n = 0;
    float my_w = 0.f, my_len = 0.f, my_ds = 0.f;
    if (lane < cnt) {
      const int k = a.sperm[base + lane];
      my_n = a.dst[k];
      my_w = load1(w + a.perm[k]);
      if (COMPUTE) {
        my_ds = a.ds[k];
        my_len = edge_len(a.pos, a.pos_dim, s, my_n);
      }
      if (a.mean) {   // d out[n] / d msg = 1 / max(deg(n), 1): folded into the slot's weight, one division per slot
        const int deg = a.rowptr[my_n + 1] - a.rowptr[my_n];
        my_w /= (float)(deg > 1 ? deg : 1);
      }
    }
    for (int j0 = 0; j0 < cnt; j0 += WEC_U) {
      float4 g[WEC_U];
      float2 Bn[WEC_U];
#pragma unroll
      for (int u = 0; u < WEC_U; ++u) {
        const int64_t n = __shfl_sync(0xffffffffu, my_n, (j0 + u) & 31);
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        Bn[u] = make_float2(0.f, 0.f);
        if (j0 + u < cnt) {
          if (colok) g[u] = load4(G + n * a.out_dim + c);
          if (COMPUTE) Bn[u] = load2(Q + n * a.ldq + WEC_HID + 2 * lane);
        }
      }
#pragma unroll
      for (int u = 0; u < WEC_U; ++u) {
        const float wgt = __shfl_sync(0xffffffffu, my_w, (j0 + u) & 31);
        const float len = __shfl_sync(0xffffffffu, my_len, (j0 + u) & 31);
        const float ds = __shfl_sync(0xffffffffu, my_ds, (j0 + u) & 31);
        // slots past the end of the chunk: g = 0, wgt = ds = 0 -> nothing is added
        fma4(dT, wgt, g[u]);
        if (COMPUTE) {
          if (As.x + Bn[u].x + wl.x * len > 0.f) dA.x += ds * w2v.x;
          if (As.y + Bn[u].y + wl.y * len > 0.f) dA.y += ds * w2v.y;
        }
      }
    }
  }
  if (COMPUTE) store2(dQ + s * a.ldq + 2 * lane, dA);
  if (colok) store4(dQ + s * a.ldq + a.toff + c, dT);
}

static int wec_grid_recv(int64_t N) {
  int64_t full = cdiv(N > 0 ? N : 1, WEC_THREADS / 32);
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(full < cap ? full : cap);
}

static int wec_validate(const aero_wec_desc* d, int backward) {
  AERO_CHECK_ARG(d != nullptr, "aero_wec: null descriptor");
  AERO_CHECK_ARG(d->N >= 0 && d->E >= 0 && d->N < (1ll << 31) && d->E < (1ll << 31), "aero_wec: bad sizes");
  if (d->dtype != AERO_F32 && d->dtype != AERO_BF16) {
    set_error("aero_wec: unsupported dtype %d", d->dtype);
    return AERO_EUNSUPPORTED;
  }
  if (d->out_dim <= 0 || d->out_dim > 128 || (d->out_dim & 3)) {
    set_error("aero_wec: out_dim=%lld outside the kernel's range (multiple of 4, <= 128)", (long long)d->out_dim);
    return AERO_EUNSUPPORTED;
  }
  const int64_t toff = d->compute_w ? 2 * WEC_HID : 0;
  AERO_CHECK_ARG(d->ldq >= toff + d->out_dim && (d->ldq & 3) == 0, "aero_wec: ldq too small or not a multiple of 4");
  AERO_CHECK_ARG(d->N == 0 || (d->Q && d->rowptr), "aero_wec: null pointer");
  AERO_CHECK_ARG(d->E == 0 || (d->src && d->dst && d->perm && d->w), "aero_wec: null edge arrays");
  if (d->compute_w) {
    AERO_CHECK_ARG(d->pos && d->w1_len && d->w2 && d->b2 && d->pos_dim >= 1 && d->pos_dim <= 8,
                   "aero_wec: edge-weight MLP inputs missing (pos, w1_len, w2, b2; 1 <= pos_dim <= 8)");
  }
  if (backward) {
    AERO_CHECK_ARG(d->N == 0 || (d->g_out && d->dQ && d->sptr), "aero_wec_bwd: null pointer");
    AERO_CHECK_ARG(d->E == 0 || d->sperm, "aero_wec_bwd: null sperm");
    if (d->compute_w) AERO_CHECK_ARG(d->g_small != nullptr, "aero_wec_bwd: g_small missing");
    else AERO_CHECK_ARG(d->E == 0 || d->g_w != nullptr, "aero_wec_bwd: g_w missing");
  } else {
    AERO_CHECK_ARG(d->out != nullptr, "aero_wec_fwd: out missing");
  }
  return AERO_OK;
}

static WecArgs wec_args(const aero_wec_desc* d) {
  WecArgs a;
  a.N = d->N; a.E = d->E; a.ldq = d->ldq;
  a.out_dim = (int)d->out_dim;
  a.toff = d->compute_w ? 2 * WEC_HID : 0;
  a.pos_dim = d->pos_dim; a.mean = d->mean;
  a.Q = d->Q; a.pos = d->pos; a.w1_len = d->w1_len; a.w2 = d->w2; a.b2 = d->b2;
  a.rowptr = d->rowptr; a.src = d->src; a.dst = d->dst; a.perm = d->perm; a.sptr = d->sptr; a.sperm = d->sperm;
  a.w = d->w; a.out = d->out; a.g_out = d->g_out; a.g_w_ext = d->g_w_ext; a.dQ = d->dQ; a.g_w = d->g_w;
  a.ds = nullptr; a.part = nullptr;
  return a;
}

}  // namespace aero

using namespace aero;

static inline dim3 grid1(int64_t n, int threads = 256) { return dim3((unsigned)cdiv(n > 0 ? n : 1, threads)); }

// ---------------------------------------------------------------------------------------------
extern "C" size_t aero_bfs_levels_workspace_bytes(int64_t N) {
  return 2 * align_up((size_t)(N > 0 ? N : 1) * 4, 256) + 256;
}

extern "C" int aero_bfs_levels(const int32_t* sptr, const int32_t* sperm, const int32_t* dst, int64_t N, int64_t E,
                               int64_t start, int64_t level_begin, int64_t level_count, int64_t* dist, int64_t* status,
                               void* workspace, size_t workspace_bytes, void* stream) {
  g_launch_count = 0;
  AERO_CHECK_ARG(sptr && dist && status && workspace && (E == 0 || (sperm && dst)), "aero_bfs_levels: null pointer");
  AERO_CHECK_ARG(N > 0 && N < (1ll << 31) && E >= 0 && start >= 0 && start < N && level_begin >= 0 && level_count >= 0,
                 "aero_bfs_levels: bad sizes (N=%lld, start=%lld)", (long long)N, (long long)start);
  if (workspace_bytes < aero_bfs_levels_workspace_bytes(N)) {
    set_error("aero_bfs_levels: workspace %zu < %zu", workspace_bytes, aero_bfs_levels_workspace_bytes(N));
    return AERO_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace);
  int32_t* qa = cv.take<int32_t>(N);
  int32_t* qb = cv.take<int32_t>(N);
  int32_t* cnt = cv.take<int32_t>(4);
  if (level_begin == 0) {
    bfs_init_kernel<<<grid1(N), 256, 0, st>>>(dist, N, start, qa, cnt);
    AERO_LAUNCH_CHECK();
  }
  const int grid = sm_count() * 4;
  for (int64_t lv = level_begin; lv < level_begin + level_count; ++lv) {
    const int32_t* qi = (lv & 1) ? qb : qa;
    int32_t* qo = (lv & 1) ? qa : qb;
    bfs_level_kernel<<<grid, 256, 0, st>>>(sptr, sperm, dst, dist, qi, qo, cnt, (int)lv);
    AERO_LAUNCH_CHECK();
  }
  bfs_status_kernel<<<1, 1, 0, st>>>(cnt, (int)(level_begin + level_count), status);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" size_t aero_bistride_select_workspace_bytes(int64_t N) {
  size_t n = (size_t)(N > 0 ? N : 1);
  return 2 * align_up(n * 4, 256) + scan_ws_bytes(N) + 256;
}

extern "C" int aero_bistride_select(const int64_t* dist, int64_t N, int64_t* selected, int64_t* index_map,
                                    int64_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  g_launch_count = 0;
  AERO_CHECK_ARG(dist && selected && index_map && counts && workspace, "aero_bistride_select: null pointer");
  AERO_CHECK_ARG(N > 0 && N < (1ll << 31), "aero_bistride_select: bad N=%lld", (long long)N);
  if (workspace_bytes < aero_bistride_select_workspace_bytes(N)) {
    set_error("aero_bistride_select: workspace too small");
    return AERO_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace);
  int32_t* fe = cv.take<int32_t>(N);
  int32_t* fr = cv.take<int32_t>(N);
  void* sws = cv.take<char>(scan_ws_bytes(N));
  select_flags_kernel<<<grid1(N), 256, 0, st>>>(dist, N, fe, fr);
  AERO_LAUNCH_CHECK();
  int rc;
  if ((rc = inclusive_scan_i32(fe, fe, N, sws, st))) return rc;
  if ((rc = inclusive_scan_i32(fr, fr, N, sws, st))) return rc;
  select_scatter_kernel<<<grid1(N), 256, 0, st>>>(dist, N, fe, fr, selected, index_map, counts);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" size_t aero_filter_edges_workspace_bytes(int64_t E) {
  size_t n = (size_t)(E > 0 ? E : 1);
  return 2 * align_up(n * 4, 256) + scan_ws_bytes(E) + 256;
}

extern "C" int aero_filter_edges(const int64_t* edge_index, int64_t E, const int64_t* index_map, int64_t N,
                                 int64_t* out_edge_index, int32_t* kept_ids, int64_t* counts, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  g_launch_count = 0;
  AERO_CHECK_ARG(index_map && out_edge_index && counts && workspace, "aero_filter_edges: null pointer");
  AERO_CHECK_ARG(E >= 0 && E < (1ll << 31) && N >= 0, "aero_filter_edges: bad sizes");
  if (workspace_bytes < aero_filter_edges_workspace_bytes(E)) {
    set_error("aero_filter_edges: workspace too small");
    return AERO_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int64_t), st));
  if (E == 0) return AERO_OK;
  AERO_CHECK_ARG(edge_index != nullptr, "aero_filter_edges: null edge_index");
  Carver cv(workspace);
  int32_t* flag = cv.take<int32_t>(E);
  int32_t* scan = cv.take<int32_t>(E);
  void* sws = cv.take<char>(scan_ws_bytes(E));
  filter_flags_kernel<<<grid1(E), 256, 0, st>>>(edge_index, E, index_map, N, flag,
                                                reinterpret_cast<unsigned long long*>(counts + 1));
  AERO_LAUNCH_CHECK();
  int rc;
  if ((rc = inclusive_scan_i32(flag, scan, E, sws, st))) return rc;
  filter_scatter_kernel<<<grid1(E), 256, 0, st>>>(edge_index, E, index_map, flag, scan, out_edge_index, kept_ids, counts);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" size_t aero_wec_workspace_bytes(const aero_wec_desc* d, int backward) {
  if (!d || !backward || !d->compute_w) return 256;
  return align_up((size_t)(d->E > 0 ? d->E : 1) * 4, 256) +
         align_up((size_t)wec_grid_recv(d->N) * WEC_PART_LD * 4, 256) + 256;
}

template <typename T>
static int wec_fwd_t(const aero_wec_desc* d, cudaStream_t st) {
  WecArgs a = wec_args(d);
  dim3 grid((unsigned)cdiv(d->N, WEC_THREADS / 32));
  // compute_w: the weights come from a lane-per-edge pass, the aggregation then reads them like given weights
  // (AERO_WEC_FUSED=1: the earlier single kernel that spreads each edge's 64-wide dot product over a warp)
  static const bool fused = [] { const char* e = getenv("AERO_WEC_FUSED"); return e && e[0] == '1'; }();
  if (d->compute_w && !fused) {
    if (d->E > 0) {
      wec_weight_kernel<T><<<grid1(d->E), 256, 0, st>>>(a);
      AERO_LAUNCH_CHECK();
    }
    wec_fwd_kernel<T, false><<<grid, WEC_THREADS, 0, st>>>(a);
  } else if (d->compute_w) wec_fwd_kernel<T, true><<<grid, WEC_THREADS, 0, st>>>(a);
  else wec_fwd_kernel<T, false><<<grid, WEC_THREADS, 0, st>>>(a);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" int aero_wec_fwd(const aero_wec_desc* d, void* stream) {
  g_launch_count = 0;
  int rc = wec_validate(d, 0);
  if (rc) return rc;
  if (d->N == 0) return AERO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return d->dtype == AERO_F32 ? wec_fwd_t<float>(d, st) : wec_fwd_t<__nv_bfloat16>(d, st);
}

template <typename T>
static int wec_bwd_t(const aero_wec_desc* d, cudaStream_t st) {
  WecArgs a = wec_args(d);
  const int g1 = wec_grid_recv(d->N);
  if (d->compute_w) {
    if (d->workspace == nullptr || d->workspace_bytes < aero_wec_workspace_bytes(d, 1)) {
      set_error("aero_wec_bwd: workspace %zu < %zu", d->workspace_bytes, aero_wec_workspace_bytes(d, 1));
      return AERO_EWORKSPACE;
    }
    Carver cv(d->workspace);
    a.ds = cv.take<float>(d->E > 0 ? d->E : 1);
    a.part = cv.take<float>((size_t)g1 * WEC_PART_LD);
    wec_bwd_recv_kernel<T, true><<<g1, WEC_THREADS, 0, st>>>(a);
    AERO_LAUNCH_CHECK();
    int rc = launch_reduce_partials(a.part, g1, WEC_PART_LD, d->g_small, WEC_SMALL, st);
    if (rc) return rc;
  } else {
    wec_bwd_recv_kernel<T, false><<<g1, WEC_THREADS, 0, st>>>(a);
    AERO_LAUNCH_CHECK();
  }
  dim3 grid((unsigned)cdiv(d->N, WEC_THREADS / 32));
  if (d->compute_w) wec_bwd_send_kernel<T, true><<<grid, WEC_THREADS, 0, st>>>(a);
  else wec_bwd_send_kernel<T, false><<<grid, WEC_THREADS, 0, st>>>(a);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" int aero_wec_bwd(const aero_wec_desc* d, void* stream) {
  g_launch_count = 0;
  int rc = wec_validate(d, 1);
  if (rc) return rc;
  if (d->N == 0) return AERO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return d->dtype == AERO_F32 ? wec_bwd_t<float>(d, st) : wec_bwd_t<__nv_bfloat16>(d, st);
}
