// sort_plan.cu -- integer / indexing kernels: stable radix sort, prefix scan, graph plan
// (receiver-CSR + sender-CSR), content hash, and the bistride pooling index kernels.
// Everything here is bit-exact integer work; no atomics decide an output position.
#include "common.cuh"

namespace aero {

// =============================================================================================
// Stable LSD radix sort, 8-bit digits.
//   pass = histogram (per block, per digit) -> exclusive scan (digit-major) -> stable scatter.
// Stability inside a block comes from ranking with __match_any_sync in lane order, warps in
// order, 256-element sub-tiles in order.
// =============================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_SUBTILES = 16;
constexpr int RS_CHUNK = RS_THREADS * RS_SUBTILES;  // keys per block

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                             int shift, int nblocks, int32_t* __restrict__ hist) {
  __shared__ int32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_CHUNK;
#pragma unroll 4
  for (int s = 0; s < RS_SUBTILES; ++s) {
    int64_t i = base + (int64_t)s * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFF], 1);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// single-block exclusive scan over m int32 entries (in place)
__global__ void __launch_bounds__(1024) excl_scan_single_kernel(int32_t* __restrict__ data, int64_t m) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry_s;
  const int t = threadIdx.x;
  int64_t per = (m + 1023) / 1024;
  int64_t lo = (int64_t)t * per, hi = lo + per < m ? lo + per : m;
  int32_t s = 0;
  for (int64_t i = lo; i < hi; ++i) s += data[i];
  // block exclusive scan of s
  int32_t v = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t u = __shfl_up_sync(0xffffffffu, v, o);
    if ((t & 31) >= o) v += u;
  }
  if ((t & 31) == 31) warp_sums[t >> 5] = v;
  __syncthreads();
  if (t < 32) {
    int32_t w = warp_sums[t];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t u = __shfl_up_sync(0xffffffffu, w, o);
      if (t >= o) w += u;
    }
    warp_sums[t] = w;
  }
  __syncthreads();
  int32_t excl = v - s + ((t >> 5) > 0 ? warp_sums[(t >> 5) - 1] : 0);
  (void)carry_s;
  int32_t run = excl;
  for (int64_t i = lo; i < hi; ++i) {
    int32_t d = data[i];
    data[i] = run;
    run += d;
  }
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                const int32_t* __restrict__ vals_in,
                                                                uint64_t* __restrict__ keys_out,
                                                                int32_t* __restrict__ vals_out, int64_t n,
                                                                int shift, int nblocks,
                                                                const int32_t* __restrict__ hist_scanned) {
  __shared__ int32_t base[256];
  __shared__ int32_t warp_cnt[RS_THREADS / 32][256];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  base[t] = hist_scanned[(size_t)t * nblocks + blockIdx.x];
  int64_t cbase = (int64_t)blockIdx.x * RS_CHUNK;
  for (int s = 0; s < RS_SUBTILES; ++s) {
    int64_t tile0 = cbase + (int64_t)s * RS_THREADS;
    if (tile0 >= n) break;  // uniform across the block
#pragma unroll
    for (int q = 0; q < RS_THREADS / 32; ++q) warp_cnt[q][t] = 0;
    __syncthreads();
    int64_t i = tile0 + t;
    bool valid = i < n;
    uint64_t key = valid ? keys_in[i] : 0ull;
    int32_t val = valid ? vals_in[i] : 0;
    uint32_t d = valid ? (uint32_t)((key >> shift) & 0xFF) : 0xFFFFFFFFu;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) warp_cnt[w][d] = __popc(peers);
    __syncthreads();
    {
      int32_t off = base[t];
#pragma unroll
      for (int q = 0; q < RS_THREADS / 32; ++q) {
        int32_t c = warp_cnt[q][t];
        warp_cnt[q][t] = off;
        off += c;
      }
      base[t] = off;
    }
    __syncthreads();
    if (valid) {
      int32_t pos = warp_cnt[w][d] + rank;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
  }
}

static size_t sort_ws_bytes(int64_t n) {
  int64_t nb = cdiv(n > 0 ? n : 1, RS_CHUNK);
  return align_up((size_t)nb * 256 * sizeof(int32_t), 256);
}

static int sort_pairs(uint64_t* ka, int32_t* va, uint64_t* kb, int32_t* vb, int64_t n, int key_bits,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < sort_ws_bytes(n)) {
    set_error("sort_pairs: workspace %zu < %zu", ws_bytes, sort_ws_bytes(n));
    return AERO_EWORKSPACE;
  }
  if (n <= 0) return AERO_OK;
  int passes = (key_bits + 7) / 8;
  int nb = (int)cdiv(n, RS_CHUNK);
  int32_t* hist = reinterpret_cast<int32_t*>(ws);
  uint64_t *kin = ka, *kout = kb;
  int32_t *vin = va, *vout = vb;
  for (int p = 0; p < passes; ++p) {
    rs_hist_kernel<<<nb, RS_THREADS, 0, st>>>(kin, n, p * 8, nb, hist);
    AERO_LAUNCH_CHECK();
    excl_scan_single_kernel<<<1, 1024, 0, st>>>(hist, (int64_t)nb * 256);
    AERO_LAUNCH_CHECK();
    rs_scatter_kernel<<<nb, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, p * 8, nb, hist);
    AERO_LAUNCH_CHECK();
    uint64_t* tk = kin; kin = kout; kout = tk;
    int32_t* tv = vin; vin = vout; vout = tv;
  }
  // result currently in (kin, vin)
  if (kin != kb) {
    AERO_CUDA(cudaMemcpyAsync(kb, kin, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    AERO_CUDA(cudaMemcpyAsync(vb, vin, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  return AERO_OK;
}

static int bit_length(uint64_t v) {
  int b = 0;
  while (v) { ++b; v >>= 1; }
  return b;
}

// =============================================================================================
// Inclusive prefix sum of int32 (three kernels, any n)
// =============================================================================================
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 4;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ int32_t block_incl_scan(int32_t v, int32_t* warp_sums /*[8]*/) {
  const int t = threadIdx.x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t u = __shfl_up_sync(0xffffffffu, v, o);
    if ((t & 31) >= o) v += u;
  }
  if ((t & 31) == 31) warp_sums[t >> 5] = v;
  __syncthreads();
  int32_t add = 0;
  for (int q = 0; q < (t >> 5); ++q) add += warp_sums[q];
  __syncthreads();
  return v + add;
}

__global__ void __launch_bounds__(SC_THREADS) scan_block_sums_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                     int32_t* __restrict__ block_sums) {
  __shared__ int32_t ws[SC_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
  int32_t s = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j)
    if (base + j < n) s += in[base + j];
  int32_t incl = block_incl_scan(s, ws);
  if (threadIdx.x == SC_THREADS - 1) block_sums[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                const int32_t* __restrict__ block_offs,
                                                                int32_t* __restrict__ out) {
  __shared__ int32_t ws[SC_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
  int32_t v[SC_ITEMS];
  int32_t s = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    s += v[j];
  }
  int32_t incl = block_incl_scan(s, ws);
  int32_t run = incl - s + block_offs[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) {
    run += v[j];
    if (base + j < n) out[base + j] = run;
  }
}

size_t scan_ws_bytes(int64_t n) { return align_up((size_t)cdiv(n > 0 ? n : 1, SC_TILE) * sizeof(int32_t), 256); }

// out may alias in
int inclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, cudaStream_t st) {
  if (n <= 0) return AERO_OK;
  int nb = (int)cdiv(n, SC_TILE);
  int32_t* bs = reinterpret_cast<int32_t*>(ws);
  scan_block_sums_kernel<<<nb, SC_THREADS, 0, st>>>(in, n, bs);
  AERO_LAUNCH_CHECK();
  excl_scan_single_kernel<<<1, 1024, 0, st>>>(bs, nb);
  AERO_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, SC_THREADS, 0, st>>>(in, n, bs, out);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

// =============================================================================================
// Graph plan
// =============================================================================================
__global__ void plan_keys_kernel(const int64_t* __restrict__ idx, int64_t E, int64_t N,
                                 uint64_t* __restrict__ keys, int32_t* __restrict__ vals,
                                 int32_t* __restrict__ status) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  int64_t v = idx[k];
  if (v < 0 || v >= N) {
    atomicAdd(status, 1);
    v = 0;
  }
  keys[k] = (uint64_t)v;
  vals[k] = (int32_t)k;
}

// after the receiver sort: perm = vals, dst = keys, src = sender[perm]
__global__ void plan_finish_recv_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                        const int64_t* __restrict__ sender, int64_t E, int64_t N,
                                        int32_t* __restrict__ perm, int32_t* __restrict__ src,
                                        int32_t* __restrict__ dst, uint64_t* __restrict__ keys2,
                                        int32_t* __restrict__ vals2, int32_t* __restrict__ status) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  int32_t p = vals[k];
  int64_t s = sender[p];
  if (s < 0 || s >= N) {
    atomicAdd(status, 1);
    s = 0;
  }
  perm[k] = p;
  dst[k] = (int32_t)keys[k];
  src[k] = (int32_t)s;
  keys2[k] = (uint64_t)s;
  vals2[k] = (int32_t)k;
}

// offsets from sorted keys: ptr[n] = first position whose key >= n, ptr[n_seg] = n_items
__global__ void ptr_from_sorted_kernel(const uint64_t* __restrict__ keys, int64_t n_items, int64_t n_seg,
                                       int32_t* __restrict__ ptr) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_items) return;
  int64_t prev = (k == 0) ? -1 : (int64_t)keys[k - 1];
  int64_t cur = (k == n_items) ? n_seg : (int64_t)keys[k];
  for (int64_t n = prev + 1; n <= cur && n <= n_seg; ++n) ptr[n] = (int32_t)k;
}

__global__ void copy_i32_kernel(const int32_t* __restrict__ a, int32_t* __restrict__ b, int64_t n) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) b[k] = a[k];
}

// =============================================================================================
// content hash: sum over 8-byte words of splitmix64(word ^ (index * phi)) (order independent)
// =============================================================================================
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void hash_kernel(const uint64_t* __restrict__ w, int64_t nwords, unsigned long long* __restrict__ out) {
  uint64_t acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x)
    acc += splitmix64(w[i] ^ ((uint64_t)i * 0x9E3779B97F4A7C15ull));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)acc);
}

// =============================================================================================
// Bistride pooling plan (bsms_mgn.py:231-262)
// =============================================================================================
__global__ void pool_heads_kernel(const int64_t* __restrict__ batch, int64_t N, int32_t* __restrict__ head) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) head[i] = (i == 0 || batch[i] != batch[i - 1]) ? 1 : 0;
}
// gord[i] (inclusive scan of head) - 1 = ordinal of the node's graph; record graph starts
__global__ void pool_gstart_kernel(const int32_t* __restrict__ head, const int32_t* __restrict__ gscan, int64_t N,
                                   int32_t* __restrict__ gstart) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && head[i]) gstart[gscan[i] - 1] = (int32_t)i;
}
__device__ __forceinline__ uint64_t orderable_f64(double x) {
  if (x == 0.0) x = 0.0;  // -0.0 -> +0.0 (they compare equal in torch.argsort)
  uint64_t b = (uint64_t)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void pool_xkeys_kernel(const double* __restrict__ posx, int64_t N, uint64_t* __restrict__ keys,
                                  int32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  keys[i] = posx ? orderable_f64(posx[i]) : 0ull;
  vals[i] = (int32_t)i;
}
__global__ void pool_gkeys_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ gscan, int64_t N,
                                  uint64_t* __restrict__ keys) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < N) keys[p] = (uint64_t)(gscan[order[p]] - 1);
}
// flag[p] = 1 when the rank of position p inside its graph is a multiple of stride
__global__ void pool_flags_kernel(const uint64_t* __restrict__ gkeys_sorted, const int32_t* __restrict__ gstart,
                                  int64_t N, int64_t stride, int32_t* __restrict__ flag) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  int64_t r = p - gstart[gkeys_sorted[p]];
  flag[p] = (r % stride == 0) ? 1 : 0;
}
__global__ void pool_assign_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ flag,
                                   const int32_t* __restrict__ cscan, const int64_t* __restrict__ batch, int64_t N,
                                   int64_t* __restrict__ f2c, int64_t* __restrict__ coarse_batch,
                                   int64_t* __restrict__ counts) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  int32_t node = order[p];
  int64_t c = (int64_t)cscan[p] - 1;
  f2c[node] = c;
  if (flag[p]) coarse_batch[c] = batch[node];
  if (p == N - 1) counts[0] = (int64_t)cscan[p];
}

// =============================================================================================
// Edge coarsening (bsms_mgn.py:274-288)
// =============================================================================================
__global__ void coarse_keys_kernel(const int64_t* __restrict__ ei, int64_t E, const int64_t* __restrict__ f2c,
                                   int64_t nc1, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  int64_t r = f2c[ei[k]], c = f2c[ei[E + k]];
  keys[k] = (uint64_t)(r * nc1 + c);
  vals[k] = (int32_t)k;
}
__global__ void uniq_heads_kernel(const uint64_t* __restrict__ sk, int64_t n, int32_t* __restrict__ head) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) head[k] = (k == 0 || sk[k] != sk[k - 1]) ? 1 : 0;
}
__global__ void coarse_finish_kernel(const uint64_t* __restrict__ sk, const int32_t* __restrict__ sv,
                                     const int32_t* __restrict__ head, const int32_t* __restrict__ uscan, int64_t E,
                                     int64_t nc1, int64_t* __restrict__ cei, int64_t* __restrict__ inverse,
                                     int32_t* __restrict__ gptr, int32_t* __restrict__ glist,
                                     int64_t* __restrict__ counts) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  int64_t u = (int64_t)uscan[k] - 1;
  inverse[sv[k]] = u;
  glist[k] = sv[k];
  if (head[k]) {
    uint64_t key = sk[k];
    cei[u] = (int64_t)(key / (uint64_t)nc1);
    cei[E + u] = (int64_t)(key % (uint64_t)nc1);
    gptr[u] = (int32_t)k;
  }
  if (k == E - 1) {
    counts[0] = u + 1;
    gptr[u + 1] = (int32_t)E;
  }
}

__global__ void group_keys_kernel(const int64_t* __restrict__ g, int64_t n, uint64_t* __restrict__ keys,
                                  int32_t* __restrict__ vals, int32_t* __restrict__ g32) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (uint64_t)g[i];
  vals[i] = (int32_t)i;
  g32[i] = (int32_t)g[i];
}

}  // namespace aero

using namespace aero;

static inline dim3 grid1d(int64_t n, int threads = 256) { return dim3((unsigned)cdiv(n > 0 ? n : 1, threads)); }

extern "C" size_t aero_sort_pairs_workspace_bytes(int64_t n) { return sort_ws_bytes(n); }

extern "C" int aero_sort_pairs_u64(uint64_t* keys_in, int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out,
                                   int64_t n, int key_bits, void* workspace, size_t workspace_bytes, void* stream) {
  AERO_CHECK_ARG(n >= 0 && key_bits >= 0 && key_bits <= 64, "aero_sort_pairs_u64: bad n/key_bits");
  AERO_CHECK_ARG(n == 0 || (keys_in && vals_in && keys_out && vals_out && workspace), "aero_sort_pairs_u64: null pointer");
  return sort_pairs(keys_in, vals_in, keys_out, vals_out, n, key_bits, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t aero_graph_plan_workspace_bytes(int64_t E, int64_t N) {
  (void)N;
  size_t e = (size_t)(E > 0 ? E : 1);
  return 2 * align_up(e * 8, 256) + 2 * align_up(e * 4, 256) + sort_ws_bytes(E) + 256;
}

extern "C" int aero_graph_plan_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr, int32_t* perm,
                                     int32_t* src, int32_t* dst, int32_t* sptr, int32_t* sperm, int32_t* status,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(E >= 0 && N >= 0 && E < 2147483647LL && N < 2147483647LL, "aero_graph_plan_build: E=%lld N=%lld out of range", (long long)E, (long long)N);
  AERO_CHECK_ARG(rowptr && sptr && status && workspace, "aero_graph_plan_build: null pointer");
  AERO_CHECK_ARG(E == 0 || (edge_index && perm && src && dst && sperm), "aero_graph_plan_build: null pointer");
  if (workspace_bytes < aero_graph_plan_workspace_bytes(E, N)) {
    set_error("aero_graph_plan_build: workspace %zu < %zu", workspace_bytes, aero_graph_plan_workspace_bytes(E, N));
    return AERO_EWORKSPACE;
  }
  AERO_CUDA(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
  Carver cv(workspace);
  size_t e = (size_t)(E > 0 ? E : 1);
  uint64_t* ka = cv.take<uint64_t>(e);
  uint64_t* kb = cv.take<uint64_t>(e);
  int32_t* va = cv.take<int32_t>(e);
  int32_t* vb = cv.take<int32_t>(e);
  void* sws = cv.base + cv.off;
  size_t sws_bytes = sort_ws_bytes(E);
  int bits = bit_length(N > 0 ? (uint64_t)(N - 1) : 0);
  if (E > 0) {
    plan_keys_kernel<<<grid1d(E), 256, 0, st>>>(edge_index + E, E, N, ka, va, status);
    AERO_LAUNCH_CHECK();
    int rc = sort_pairs(ka, va, kb, vb, E, bits, sws, sws_bytes, st);
    if (rc) return rc;
    plan_finish_recv_kernel<<<grid1d(E), 256, 0, st>>>(kb, vb, edge_index, E, N, perm, src, dst, ka, va, status);
    AERO_LAUNCH_CHECK();
  }
  ptr_from_sorted_kernel<<<grid1d(E + 1), 256, 0, st>>>(kb, E, N, rowptr);
  AERO_LAUNCH_CHECK();
  if (E > 0) {
    int rc = sort_pairs(ka, va, kb, vb, E, bits, sws, sws_bytes, st);
    if (rc) return rc;
    copy_i32_kernel<<<grid1d(E), 256, 0, st>>>(vb, sperm, E);
    AERO_LAUNCH_CHECK();
  }
  ptr_from_sorted_kernel<<<grid1d(E + 1), 256, 0, st>>>(kb, E, N, sptr);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" int aero_hash_u64(const void* data, int64_t nbytes, uint64_t* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(out && nbytes >= 0 && (nbytes % 8) == 0, "aero_hash_u64: nbytes must be a multiple of 8");
  AERO_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t), st));
  if (nbytes == 0) return AERO_OK;
  AERO_CHECK_ARG(data != nullptr, "aero_hash_u64: null data");
  int64_t nw = nbytes / 8;
  int blocks = (int)(cdiv(nw, 256) < 1184 ? cdiv(nw, 256) : 1184);
  hash_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint64_t*>(data), nw,
                                      reinterpret_cast<unsigned long long*>(out));
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" size_t aero_stride_pool_workspace_bytes(int64_t N) {
  size_t n = (size_t)(N > 0 ? N : 1);
  return 2 * align_up(n * 8, 256) + 6 * align_up(n * 4, 256) + sort_ws_bytes(N) + scan_ws_bytes(N) + 256;
}

extern "C" int aero_stride_pool_plan(const int64_t* batch, const double* posx, int64_t N, int64_t stride,
                                     int64_t* fine_to_coarse, int64_t* coarse_batch, int64_t* counts, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(N >= 0 && N < 2147483647LL && stride >= 1, "aero_stride_pool_plan: bad N/stride");
  AERO_CHECK_ARG(counts && workspace, "aero_stride_pool_plan: null pointer");
  AERO_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int64_t), st));
  if (N == 0) return AERO_OK;
  AERO_CHECK_ARG(batch && fine_to_coarse && coarse_batch, "aero_stride_pool_plan: null pointer");
  if (workspace_bytes < aero_stride_pool_workspace_bytes(N)) {
    set_error("aero_stride_pool_plan: workspace too small");
    return AERO_EWORKSPACE;
  }
  Carver cv(workspace);
  uint64_t* ka = cv.take<uint64_t>(N);
  uint64_t* kb = cv.take<uint64_t>(N);
  int32_t* va = cv.take<int32_t>(N);
  int32_t* vb = cv.take<int32_t>(N);
  int32_t* head = cv.take<int32_t>(N);
  int32_t* gscan = cv.take<int32_t>(N);
  int32_t* gstart = cv.take<int32_t>(N);
  int32_t* flag = cv.take<int32_t>(N);
  void* sws = cv.take<char>(sort_ws_bytes(N));
  void* cws = cv.take<char>(scan_ws_bytes(N));
  int rc;
  pool_heads_kernel<<<grid1d(N), 256, 0, st>>>(batch, N, head);
  AERO_LAUNCH_CHECK();
  if ((rc = inclusive_scan_i32(head, gscan, N, cws, st))) return rc;
  pool_gstart_kernel<<<grid1d(N), 256, 0, st>>>(head, gscan, N, gstart);
  AERO_LAUNCH_CHECK();
  // order nodes by x (stable), then by graph ordinal (stable) == per-graph argsort of pos[:,0]
  pool_xkeys_kernel<<<grid1d(N), 256, 0, st>>>(posx, N, ka, va);
  AERO_LAUNCH_CHECK();
  int32_t* order = va;
  if (posx) {
    if ((rc = sort_pairs(ka, va, kb, vb, N, 64, sws, sort_ws_bytes(N), st))) return rc;
    order = vb;
  }
  // second key: graph ordinal of the node at each position
  uint64_t* gk_in = (order == vb) ? ka : kb;
  uint64_t* gk_out = (order == vb) ? kb : ka;
  int32_t* ord_out = (order == vb) ? va : vb;
  pool_gkeys_kernel<<<grid1d(N), 256, 0, st>>>(order, gscan, N, gk_in);
  AERO_LAUNCH_CHECK();
  if ((rc = sort_pairs(gk_in, order, gk_out, ord_out, N, bit_length((uint64_t)N), sws, sort_ws_bytes(N), st))) return rc;
  pool_flags_kernel<<<grid1d(N), 256, 0, st>>>(gk_out, gstart, N, stride, flag);
  AERO_LAUNCH_CHECK();
  int32_t* cscan = head;  // head no longer needed
  if ((rc = inclusive_scan_i32(flag, cscan, N, cws, st))) return rc;
  pool_assign_kernel<<<grid1d(N), 256, 0, st>>>(ord_out, flag, cscan, batch, N, fine_to_coarse, coarse_batch, counts);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" size_t aero_coarsen_edges_workspace_bytes(int64_t E) {
  size_t e = (size_t)(E > 0 ? E : 1);
  return 2 * align_up(e * 8, 256) + 4 * align_up(e * 4, 256) + sort_ws_bytes(E) + scan_ws_bytes(E) + 256;
}

extern "C" int aero_coarsen_edges(const int64_t* edge_index, int64_t E, const int64_t* fine_to_coarse, int64_t Nc,
                                  int64_t* coarse_edge_index, int64_t* inverse, int32_t* gptr, int32_t* glist,
                                  int64_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(E >= 0 && E < 2147483647LL && Nc >= 0, "aero_coarsen_edges: bad E/Nc");
  AERO_CHECK_ARG(counts && workspace && gptr, "aero_coarsen_edges: null pointer");
  AERO_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int64_t), st));
  AERO_CUDA(cudaMemsetAsync(gptr, 0, sizeof(int32_t), st));
  if (E == 0) return AERO_OK;
  AERO_CHECK_ARG(edge_index && fine_to_coarse && coarse_edge_index && inverse && glist, "aero_coarsen_edges: null pointer");
  if (workspace_bytes < aero_coarsen_edges_workspace_bytes(E)) {
    set_error("aero_coarsen_edges: workspace too small");
    return AERO_EWORKSPACE;
  }
  int64_t nc1 = Nc > 1 ? Nc : 1;
  // keys < nc1*nc1
  unsigned __int128 mk = (unsigned __int128)nc1 * (unsigned __int128)nc1 - 1;
  AERO_CHECK_ARG((mk >> 63) == 0, "aero_coarsen_edges: Nc^2 overflows int64 (same limit as the reference key)");
  int bits = bit_length((uint64_t)mk);
  Carver cv(workspace);
  uint64_t* ka = cv.take<uint64_t>(E);
  uint64_t* kb = cv.take<uint64_t>(E);
  int32_t* va = cv.take<int32_t>(E);
  int32_t* vb = cv.take<int32_t>(E);
  int32_t* head = cv.take<int32_t>(E);
  int32_t* uscan = cv.take<int32_t>(E);
  void* sws = cv.take<char>(sort_ws_bytes(E));
  void* cws = cv.take<char>(scan_ws_bytes(E));
  int rc;
  coarse_keys_kernel<<<grid1d(E), 256, 0, st>>>(edge_index, E, fine_to_coarse, nc1, ka, va);
  AERO_LAUNCH_CHECK();
  if ((rc = sort_pairs(ka, va, kb, vb, E, bits, sws, sort_ws_bytes(E), st))) return rc;
  uniq_heads_kernel<<<grid1d(E), 256, 0, st>>>(kb, E, head);
  AERO_LAUNCH_CHECK();
  if ((rc = inclusive_scan_i32(head, uscan, E, cws, st))) return rc;
  coarse_finish_kernel<<<grid1d(E), 256, 0, st>>>(kb, vb, head, uscan, E, nc1, coarse_edge_index, inverse, gptr, glist,
                                                  counts);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" size_t aero_group_lists_workspace_bytes(int64_t n) {
  size_t e = (size_t)(n > 0 ? n : 1);
  return 2 * align_up(e * 8, 256) + 2 * align_up(e * 4, 256) + sort_ws_bytes(n) + 256;
}

extern "C" int aero_group_lists(const int64_t* group_of, int64_t n, int64_t n_groups, int32_t* gptr, int32_t* glist,
                                int32_t* group32, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(n >= 0 && n < 2147483647LL && n_groups >= 0 && n_groups < 2147483647LL, "aero_group_lists: bad sizes");
  AERO_CHECK_ARG(gptr && workspace, "aero_group_lists: null pointer");
  if (workspace_bytes < aero_group_lists_workspace_bytes(n)) {
    set_error("aero_group_lists: workspace too small");
    return AERO_EWORKSPACE;
  }
  Carver cv(workspace);
  size_t e = (size_t)(n > 0 ? n : 1);
  uint64_t* ka = cv.take<uint64_t>(e);
  uint64_t* kb = cv.take<uint64_t>(e);
  int32_t* va = cv.take<int32_t>(e);
  int32_t* vb = cv.take<int32_t>(e);
  void* sws = cv.take<char>(sort_ws_bytes(n));
  if (n > 0) {
    AERO_CHECK_ARG(group_of && glist && group32, "aero_group_lists: null pointer");
    group_keys_kernel<<<grid1d(n), 256, 0, st>>>(group_of, n, ka, va, group32);
    AERO_LAUNCH_CHECK();
    int rc = sort_pairs(ka, va, kb, vb, n, bit_length(n_groups > 0 ? (uint64_t)(n_groups - 1) : 0), sws, sort_ws_bytes(n), st);
    if (rc) return rc;
    copy_i32_kernel<<<grid1d(n), 256, 0, st>>>(vb, glist, n);
    AERO_LAUNCH_CHECK();
  }
  ptr_from_sorted_kernel<<<grid1d(n + 1), 256, 0, st>>>(kb, n, n_groups, gptr);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}
