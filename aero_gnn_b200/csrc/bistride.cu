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

// forward: out[n] = sum_k w_k * T[src_k] over the receiver's CSR segment; w_k from the edge-weight MLP (COMPUTE) or given
template <typename T, bool COMPUTE>
__global__ void __launch_bounds__(WEC_THREADS) wec_fwd_kernel(WecArgs a) {
  const int64_t n = (int64_t)blockIdx.x * (WEC_THREADS / 32) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int lane = threadIdx.x & 31;
  const T* Q = reinterpret_cast<const T*>(a.Q);
  T* w = reinterpret_cast<T*>(a.w);
  const int b = a.rowptr[n], e = a.rowptr[n + 1];
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
  for (int k = b; k < e; ++k) {
    const int64_t s = a.src[k];
    const int pk = a.perm[k];
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (colok) t = load4(Q + s * a.ldq + a.toff + c);
    float wgt;
    if (COMPUTE) {
      float2 As = load2(Q + s * a.ldq + 2 * lane);
      const float len = edge_len(a.pos, a.pos_dim, s, n);
      const float h0 = fmaxf(As.x + Bn.x + wl.x * len, 0.f);
      const float h1 = fmaxf(As.y + Bn.y + wl.y * len, 0.f);
      const float sc = warp_sum(h0 * w2v.x + h1 * w2v.y) + b2v;
      wgt = round_to<T>(1.f / (1.f + expf(-sc)));
      if (lane == 0) store1(w + pk, wgt);
    } else {
      wgt = load1(w + pk);
    }
    acc.x = fmaf(wgt, t.x, acc.x); acc.y = fmaf(wgt, t.y, acc.y);
    acc.z = fmaf(wgt, t.z, acc.z); acc.w = fmaf(wgt, t.w, acc.w);
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
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const T* Q = reinterpret_cast<const T*>(a.Q);
  const T* G = reinterpret_cast<const T*>(a.g_out);
  const T* gwx = reinterpret_cast<const T*>(a.g_w_ext);
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
    const int b = a.rowptr[n], e = a.rowptr[n + 1];
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (colok) g = load4(G + n * a.out_dim + c);
    if (a.mean) {
      const float cnt = (float)(e - b > 1 ? e - b : 1);
      g.x /= cnt; g.y /= cnt; g.z /= cnt; g.w /= cnt;
    }
    float2 Bn = make_float2(0.f, 0.f), dB = Bn;
    if (COMPUTE) Bn = load2(Q + n * a.ldq + WEC_HID + 2 * lane);
    for (int k = b; k < e; ++k) {
      const int64_t s = a.src[k];
      const int pk = a.perm[k];
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (colok) t = load4(Q + s * a.ldq + a.toff + c);
      float dw = warp_sum(t.x * g.x + t.y * g.y + t.z * g.z + t.w * g.w);
      if (gwx) dw += load1(gwx + pk);
      if (COMPUTE) {
        float2 As = load2(Q + s * a.ldq + 2 * lane);
        const float len = edge_len(a.pos, a.pos_dim, s, n);
        const float h0 = fmaxf(As.x + Bn.x + wl.x * len, 0.f);
        const float h1 = fmaxf(As.y + Bn.y + wl.y * len, 0.f);
        const float sc = warp_sum(h0 * w2v.x + h1 * w2v.y) + b2v;
        const float wgt = 1.f / (1.f + expf(-sc));
        const float ds = dw * wgt * (1.f - wgt);
        if (lane == 0) a.ds[k] = ds;
        const float dz0 = h0 > 0.f ? ds * w2v.x : 0.f;
        const float dz1 = h1 > 0.f ? ds * w2v.y : 0.f;
        dB.x += dz0; dB.y += dz1;
        dwl.x = fmaf(dz0, len, dwl.x); dwl.y = fmaf(dz1, len, dwl.y);
        dw2.x = fmaf(ds, h0, dw2.x); dw2.y = fmaf(ds, h1, dw2.y);
        db2 += ds;
      } else {
        if (lane == 0) store1(reinterpret_cast<T*>(a.g_w) + pk, dw);
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
  const int64_t s = (int64_t)blockIdx.x * (WEC_THREADS / 32) + (threadIdx.x >> 5);
  if (s >= a.N) return;
  const int lane = threadIdx.x & 31;
  const T* Q = reinterpret_cast<const T*>(a.Q);
  const T* G = reinterpret_cast<const T*>(a.g_out);
  const T* w = reinterpret_cast<const T*>(a.w);
  T* dQ = reinterpret_cast<T*>(a.dQ);
  const int c = lane * 4;
  const bool colok = c < a.out_dim;
  float2 As = make_float2(0.f, 0.f), wl = As, w2v = As, dA = As;
  if (COMPUTE) {
    As = load2(Q + s * a.ldq + 2 * lane);
    wl = *reinterpret_cast<const float2*>(a.w1_len + 2 * lane);
    w2v = *reinterpret_cast<const float2*>(a.w2 + 2 * lane);
  }
  float4 dT = make_float4(0.f, 0.f, 0.f, 0.f);
  const int b = a.sptr[s], e = a.sptr[s + 1];
  for (int j = b; j < e; ++j) {
    const int k = a.sperm[j];
    const int64_t n = a.dst[k];
    const float wgt = load1(w + a.perm[k]);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (colok) g = load4(G + n * a.out_dim + c);
    if (a.mean) {
      const int deg = a.rowptr[n + 1] - a.rowptr[n];
      const float cnt = (float)(deg > 1 ? deg : 1);
      g.x /= cnt; g.y /= cnt; g.z /= cnt; g.w /= cnt;
    }
    dT.x = fmaf(wgt, g.x, dT.x); dT.y = fmaf(wgt, g.y, dT.y);
    dT.z = fmaf(wgt, g.z, dT.z); dT.w = fmaf(wgt, g.w, dT.w);
    if (COMPUTE) {
      float2 Bn = load2(Q + n * a.ldq + WEC_HID + 2 * lane);
      const float len = edge_len(a.pos, a.pos_dim, s, n);
      const float ds = a.ds[k];
      if (As.x + Bn.x + wl.x * len > 0.f) dA.x += ds * w2v.x;
      if (As.y + Bn.y + wl.y * len > 0.f) dA.y += ds * w2v.y;
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
  AERO_CHECK_ARG(sptr && sperm && dst && dist && status && workspace, "aero_bfs_levels: null pointer");
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
  if (d->compute_w) wec_fwd_kernel<T, true><<<grid, WEC_THREADS, 0, st>>>(a);
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
