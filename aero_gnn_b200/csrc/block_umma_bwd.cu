// block_umma_bwd.cu -- backward of the fused MGN block on tcgen05 tensor cores.
//
// One persistent CTA (128 threads) per SM owns one 128-row tile at a time and runs, per tile,
//   forward recompute : (L+2) GEMMs   h_m = act(h_{m-1} W_m^T + ..),  y = h_L W_out^T + b
//   LayerNorm backward in registers (thread = row): dL/dy, plus the column sums for d(gamma), d(beta)
//   for m = L+1 .. 1  : dW_m += G_m^T H_{m-1}   (both operands MN-major views of the same row tiles,
//                                               fp32 accumulators stay in TMEM for the whole kernel)
//                       G_{m-1} = (G_m W_m) * act'(H_{m-1})   (W_m read as an MN-major operand), written in place
//                       over H_{m-1}
//   m = 0             : g_main = G_0 W_main (+ residual gradient)
// i.e. 3(L+1)+1 GEMMs of 128x128x128 per tile, none of whose operands ever leaves shared memory / TMEM.
//
// Shared memory holds max(L+2,3) activation tiles and TWO weight slots; weight matrix m lives in slot (m & 1) and
// is streamed from the L2-resident bf16 image with cp.async.bulk one GEMM ahead (4 x 32 KB per tile for L = 2).
// TMEM: columns [0,128) working accumulator, [128 m, 128 m + 128) the dW_m accumulator (512 columns for L = 2).
// Bias / LayerNorm-parameter gradients are column sums taken by thread = column over the bf16 tiles while the
// MMAs run.  Per-CTA partial gradients are written once at the end and reduced in CTA order (deterministic).
#include "umma_block.cuh"

namespace aero {

constexpr int BWD_THREADS = 128;

__device__ __forceinline__ int h_tile(int m) { return m == 0 ? 1 : (m == 1 ? 0 : m); }   // tile holding H_m

// column sum of a bf16 row tile: thread = column c, rows in order
__device__ __forceinline__ float tile_col_sum(const uint8_t* tile, int c) {
  const uint8_t* colp = tile + (c >> 6) * PANEL_BYTES + (c & 7) * 2;
  const int cc = (c >> 3) & 7;
  float s = 0.f;
#pragma unroll 8
  for (int t = 0; t < 128; ++t) {
    uint16_t h = *reinterpret_cast<const uint16_t*>(colp + t * 128 + ((cc ^ (t & 7)) << 4));
    s += __uint_as_float((uint32_t)h << 16);
  }
  return s;
}

__global__ void __launch_bounds__(BWD_THREADS, 1) umma_block_bwd_kernel(UmmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int L = a.L;
  const int NT = (L + 2) > 3 ? (L + 2) : 3;
  uint8_t* Wslot = smem;                                    // 2 tiles
  uint8_t* X = Wslot + 2 * TILE_BYTES;                      // NT tiles
  float* vec = reinterpret_cast<float*>(X + (size_t)NT * TILE_BYTES);
  int* sidx0 = reinterpret_cast<int*>(vec + (L + 3) * 128);
  int* sidx1 = sidx0 + 128;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sidx1 + 128);   // [0] mma, [1..2] weight slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3);

  const int tid = threadIdx.x, q = tid >> 5;
  const int row = tid;
  {
    const float* vs = reinterpret_cast<const float*>(a.prep + (size_t)(L + 2) * TILE_BYTES);
    for (int i = tid; i < (L + 3) * 128; i += BWD_THREADS) vec[i] = vs[i];
  }
  if (tid == 0) {
    mbar_init(smem_u32(&mbar[0]), 1);
    mbar_init(smem_u32(&mbar[1]), 1);
    mbar_init(smem_u32(&mbar[2]), 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t bar_mma = smem_u32(&mbar[0]);
  const uint32_t x_s = smem_u32(X);
  const uint32_t w_s = smem_u32(Wslot);
  uint8_t* G = X + (size_t)(NT - 1) * TILE_BYTES;
  const uint32_t g_s = x_s + (uint32_t)(NT - 1) * TILE_BYTES;
  const int act = a.act;

  // ---- weight streaming state (thread 0 only) ----
  int slot_mat[2] = {-1, -1};
  bool slot_pending[2] = {false, false};
  uint32_t slot_phase[2] = {0, 0};
  auto prefetch = [&](int m) {   // thread 0
    int s = m & 1;
    if (slot_mat[s] == m) return;
    uint32_t bar = smem_u32(&mbar[1 + s]);
    mbar_expect_tx(bar, TILE_BYTES);
    bulk_g2s(w_s + (uint32_t)s * TILE_BYTES, a.prep + (size_t)m * TILE_BYTES, TILE_BYTES, bar);
    slot_mat[s] = m;
    slot_pending[s] = true;
  };
  auto acquire = [&](int m) -> uint32_t {   // thread 0: weight m resident -> its smem address
    int s = m & 1;
    if (slot_mat[s] != m) prefetch(m);
    if (slot_pending[s]) {
      mbar_wait(smem_u32(&mbar[1 + s]), slot_phase[s]);
      slot_phase[s] ^= 1;
      slot_pending[s] = false;
    }
    return w_s + (uint32_t)s * TILE_BYTES;
  };
  if (tid == 0) {
    prefetch(0);
    if (L + 1 >= 1) prefetch(1);
  }

  uint32_t phase = 0;
  bool first_tile = true;
  float db[UMMA_MAX_L_BWD + 1];
#pragma unroll
  for (int i = 0; i <= UMMA_MAX_L_BWD; ++i) db[i] = 0.f;
  float dgam = 0.f, dbet = 0.f;

  const int64_t tiles = (a.rows + 127) / 128;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    const int nrows = (int)((a.rows - row0) < 128 ? (a.rows - row0) : 128);
    const bool valid = row < nrows;
    __syncthreads();   // previous tile fully consumed
    // ---- stage main rows, gather indices, and the incoming gradient tile (g_out + g_agg[receiver]) ----
    if (a.main_f32) stage_rows<true>(X, a.main, a.main_scale, row0, nrows, tid);
    else stage_rows<false>(X, a.main, nullptr, row0, nrows, tid);
    {
      int64_t r = row0 + tid;
      sidx0[tid] = valid ? (a.idx0 ? a.idx0[r] : (int)r) : 0;
      sidx1[tid] = valid ? (a.idx1 ? a.idx1[r] : -1) : -1;
    }
    {
      const int chunk = tid & 15;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        int r = (tid >> 4) + i * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < nrows) {
          v = *reinterpret_cast<const uint4*>(a.g_out + (row0 + r) * 128 + chunk * 8);
          if (a.g_agg) {
            int n = a.idx1[row0 + r];
            const float* gp = a.g_agg + (size_t)n * 128 + chunk * 8;
            float4 p0 = *reinterpret_cast<const float4*>(gp);
            float4 p1 = *reinterpret_cast<const float4*>(gp + 4);
            v.x = pack_bf16(bf16_lo(v.x) + p0.x, bf16_hi(v.x) + p0.y);
            v.y = pack_bf16(bf16_lo(v.y) + p0.z, bf16_hi(v.y) + p0.w);
            v.z = pack_bf16(bf16_lo(v.z) + p1.x, bf16_hi(v.z) + p1.y);
            v.w = pack_bf16(bf16_lo(v.w) + p1.z, bf16_hi(v.w) + p1.w);
          }
        }
        *reinterpret_cast<uint4*>(G + tile_chunk_off(r, chunk)) = v;
      }
    }
    fence_async_smem();
    __syncthreads();

    // ---- forward recompute ----
    for (int m = 0; m <= L + 1; ++m) {
      if (tid == 0) {
        uint32_t wa = acquire(m);
        fence_after_sync();
        uint32_t a_addr = (m == 0) ? x_s : x_s + (uint32_t)h_tile(m - 1) * TILE_BYTES;
        issue_gemm(tmem_base, a_addr, false, wa, false, false);
        mma_commit(bar_mma);
        if (m + 1 <= L + 1) prefetch(m + 1);   // other slot: last read by GEMM m-1, already complete
      }
      if (m == 0) dbet += tile_col_sum(G, tid);   // d(beta) column sum of the incoming gradient, under the MMA
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      fence_after_sync();
      if (m <= L) {
        uint8_t* Ht = X + (size_t)h_tile(m) * TILE_BYTES;
        const __nv_bfloat16* p0 = nullptr;
        const __nv_bfloat16* p1 = nullptr;
        if (m == 0 && valid) {
          p0 = a.P + (int64_t)sidx0[row] * a.ldp + a.poff0;
          if (sidx1[row] >= 0) p1 = a.P + (int64_t)sidx1[row] * a.ldp + a.poff1;
        }
        const float* bias = m > 0 ? vec + (m - 1) * 128 : nullptr;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint4 g0[4], g1[4];
          if (m == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              g0[j] = p0 ? *reinterpret_cast<const uint4*>(p0 + c * 32 + j * 8) : make_uint4(0u, 0u, 0u, 0u);
              g1[j] = p1 ? *reinterpret_cast<const uint4*>(p1 + c * 32 + j * 8) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
          float v[32];
          tmem_ld32(tlane + (uint32_t)(c * 32), v);
          if (m == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              add_bf16x8(v + 8 * j, g0[j]);
              add_bf16x8(v + 8 * j, g1[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bias[c * 32 + j];
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? relu_or_act(v[j], act) : 0.f;
          store_row32(Ht, row, c, v);
        }
        fence_before_sync();
        fence_async_smem();
        __syncthreads();
      }
    }
    // ---- LayerNorm backward (thread = row); G holds g = dL/d(out) and ends up holding dL/dy ----
    {
      float v[128];
      const float* bo = vec + L * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float t[32];
        tmem_ld32(tlane + (uint32_t)(c * 32), t);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[c * 32 + j] = t[j] + bo[c * 32 + j];
      }
      fence_before_sync();
      if (a.use_ln) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 128; ++j) s += v[j];
        const float mean = s * (1.f / 128.f);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          float d = v[j] - mean;
          ss = fmaf(d, d, ss);
        }
        const float rstd = rsqrtf(ss * (1.f / 128.f) + 1e-5f);
        const float* gam = vec + (L + 1) * 128;
        float m1 = 0.f, m2 = 0.f;
        uint32_t gp[64];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint4 g4 = *reinterpret_cast<const uint4*>(G + tile_chunk_off(row, j));
          gp[4 * j + 0] = g4.x; gp[4 * j + 1] = g4.y; gp[4 * j + 2] = g4.z; gp[4 * j + 3] = g4.w;
          float z[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float glo = bf16_lo(gp[4 * j + k]), ghi = bf16_hi(gp[4 * j + k]);
            float hlo = (v[8 * j + 2 * k] - mean) * rstd, hhi = (v[8 * j + 2 * k + 1] - mean) * rstd;
            v[8 * j + 2 * k] = hlo;
            v[8 * j + 2 * k + 1] = hhi;
            float wlo = glo * gam[8 * j + 2 * k], whi = ghi * gam[8 * j + 2 * k + 1];
            m1 += wlo + whi;
            m2 = fmaf(wlo, hlo, fmaf(whi, hhi, m2));
            z[2 * k] = glo * hlo;
            z[2 * k + 1] = ghi * hhi;
          }
          uint4 zq;
          zq.x = pack_bf16(z[0], z[1]); zq.y = pack_bf16(z[2], z[3]);
          zq.z = pack_bf16(z[4], z[5]); zq.w = pack_bf16(z[6], z[7]);
          *reinterpret_cast<uint4*>(G + tile_chunk_off(row, j)) = zq;   // z = g * yhat, for d(gamma)
        }
        m1 *= (1.f / 128.f);
        m2 *= (1.f / 128.f);
        __syncthreads();
        dgam += tile_col_sum(G, tid);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float glo = bf16_lo(gp[4 * j + k]), ghi = bf16_hi(gp[4 * j + k]);
            o[2 * k] = rstd * (glo * gam[8 * j + 2 * k] - m1 - v[8 * j + 2 * k] * m2);
            o[2 * k + 1] = rstd * (ghi * gam[8 * j + 2 * k + 1] - m1 - v[8 * j + 2 * k + 1] * m2);
          }
          uint4 oq;
          oq.x = pack_bf16(o[0], o[1]); oq.y = pack_bf16(o[2], o[3]);
          oq.z = pack_bf16(o[4], o[5]); oq.w = pack_bf16(o[6], o[7]);
          *reinterpret_cast<uint4*>(G + tile_chunk_off(row, j)) = oq;
        }
      }
      // use_ln == 0: dL/dy = g, already in G
      fence_async_smem();
      __syncthreads();
    }
    // ---- backward through the Linear layers ----
    uint8_t* Gc = G;
    uint32_t gc_s = g_s;
    for (int m = L + 1; m >= 0; --m) {
      if (tid == 0) {
        uint32_t wa = acquire(m);
        fence_after_sync();
        if (m >= 1) {
          uint32_t h_addr = x_s + (uint32_t)h_tile(m - 1) * TILE_BYTES;
          issue_gemm(tmem_base + (uint32_t)(128 * m), gc_s, true, h_addr, true, !first_tile);   // dW_m += G^T H
        }
        issue_gemm(tmem_base, gc_s, false, wa, true, false);                                     // G W_m
        mma_commit(bar_mma);
        if (m >= 1) prefetch(m - 1);
      }
      if (m >= 1) db[m - 1] += tile_col_sum(Gc, tid);   // bias gradient of Linear m (index m-1: 0..L-1 hidden, L out)
      if (m == 0) {
        // g_h0 leaves through a coalesced copy of its tile while the last GEMM runs
        const int chunk = tid & 15;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          int r = (tid >> 4) + i * 8;
          if (r < nrows)
            *reinterpret_cast<uint4*>(a.g_h0 + (row0 + r) * 128 + chunk * 8) =
                *reinterpret_cast<const uint4*>(Gc + tile_chunk_off(r, chunk));
        }
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      fence_after_sync();
      if (m >= 1) {
        uint8_t* Ht = X + (size_t)h_tile(m - 1) * TILE_BYTES;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[32];
          tmem_ld32(tlane + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 h4 = *reinterpret_cast<const uint4*>(Ht + tile_chunk_off(row, c * 4 + j));
            uint32_t hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float hlo = bf16_lo(hh[k]), hhi = bf16_hi(hh[k]);
              if (act == AERO_ACT_RELU) {
                v[8 * j + 2 * k] = hlo > 0.f ? v[8 * j + 2 * k] : 0.f;
                v[8 * j + 2 * k + 1] = hhi > 0.f ? v[8 * j + 2 * k + 1] : 0.f;
              } else {
                v[8 * j + 2 * k] *= act_grad_from_out(hlo, act);
                v[8 * j + 2 * k + 1] *= act_grad_from_out(hhi, act);
              }
            }
          }
          store_row32(Ht, row, c, v);   // in place: G_{m-1} over H_{m-1}
        }
        fence_before_sync();
        fence_async_smem();
        __syncthreads();
        Gc = Ht;
        gc_s = x_s + (uint32_t)h_tile(m - 1) * TILE_BYTES;
      } else {
        // g_main = G_0 W_main (* scale) (+ residual gradient)
        const float sc = (a.main_scale && valid) ? a.main_scale[row0 + row] : 1.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[32];
          tmem_ld32(tlane + (uint32_t)(c * 32), v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= sc;
            if (a.has_resid_grad) {
              const uint4* gr = reinterpret_cast<const uint4*>(a.g_out + (row0 + row) * 128 + c * 32);
              const float* ga = a.g_agg ? a.g_agg + (size_t)sidx1[row] * 128 + c * 32 : nullptr;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                add_bf16x8(v + 8 * j, gr[j]);
                if (ga) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) v[8 * j + k] += ga[8 * j + k];
                }
              }
            }
            if (a.main_f32) {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.g_main) + (row0 + row) * 128 + c * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.g_main) + (row0 + row) * 128 + c * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o4;
                o4.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]); o4.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                o4.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); o4.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                op[j] = o4;
              }
            }
          }
        }
        fence_before_sync();
      }
    }
    first_tile = false;
  }

  // ---- flush per-CTA partial gradients: dW_m from TMEM, vectors from registers ----
  __syncthreads();
  fence_after_sync();
  const PackedLayout pl{L};
  float* part = a.w_part + (size_t)blockIdx.x * pl.total();
  for (int m = 1; m <= L + 1; ++m) {
    float* dst = part + (m == L + 1 ? pl.w_out() : pl.w_hidden(m - 1)) + (size_t)row * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float v[32];
      tmem_ld32(tlane + (uint32_t)(128 * m + c * 32), v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(dst + c * 32 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  for (int l = 0; l < L; ++l) part[pl.b_hidden(l) + tid] = db[l];
  part[pl.b_out() + tid] = db[L];
  part[pl.gamma() + tid] = dgam;
  part[pl.beta() + tid] = dbet;
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tmem_base);
}

// ---- host side ------------------------------------------------------------------------------------
static size_t bwd_smem(int L) {
  int nt = (L + 2) > 3 ? (L + 2) : 3;
  return 1024 + (size_t)(2 + nt) * TILE_BYTES + (size_t)(L + 3) * 512 + 1024 + 3 * 8 + 16;
}
static int bwd_grid_umma(int64_t rows) {
  int64_t tiles = cdiv(rows > 0 ? rows : 1, 128);
  return (int)(tiles < sm_count() ? tiles : sm_count());
}

size_t umma_bwd_workspace_bytes(const aero_block_desc* d) {
  PackedLayout pl{d->L};
  return align_up((size_t)bwd_grid_umma(d->rows) * pl.total() * sizeof(float), 256);
}

int umma_block_bwd(const aero_block_desc* d, cudaStream_t st) {
  if (d->dtype != AERO_BF16 || d->L > UMMA_MAX_L_BWD) {
    set_error("umma_block_bwd: needs bf16 rows and L <= %d", UMMA_MAX_L_BWD);
    return AERO_EUNSUPPORTED;
  }
  UmmaArgs a = make_uargs(d);
  PackedLayout pl{d->L};
  const size_t off = pl.w_hidden(0);
  if (d->rows == 0) {
    AERO_CUDA(cudaMemsetAsync(d->g_w + off, 0, (pl.total() - off) * sizeof(float), st));
    return AERO_OK;
  }
  a.w_part = reinterpret_cast<float*>(d->workspace);
  static bool attr_set = false;
  if (!attr_set) {
    AERO_CUDA(cudaFuncSetAttribute(umma_block_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)bwd_smem(UMMA_MAX_L_BWD)));
    attr_set = true;
  }
  int grid = bwd_grid_umma(d->rows);
  umma_block_bwd_kernel<<<grid, BWD_THREADS, bwd_smem(d->L), st>>>(a);
  AERO_LAUNCH_CHECK();
  return launch_reduce_partials(a.w_part + off, grid, pl.total(), d->g_w + off, pl.total() - off, st);
}

}  // namespace aero

extern "C" int aero_has_umma_bwd(void) { return 1; }
