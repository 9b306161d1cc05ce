// block_umma_bwd.cu -- backward of the fused MGN block on tcgen05 tensor cores.
//
// One persistent CTA (512 threads, 16 warps) per SM owns one 128-row tile at a time and runs, per tile,
//   forward recompute : (L+2) GEMMs   h_m = act(h_{m-1} W_m^T + ..),  y = h_L W_out^T + b
//                       (L+1 when the forward kept h_0: aero_block_desc.h0 -- no gather, no first GEMM/epilogue)
//   LayerNorm backward (thread = (row, 32-column chunk), row statistics exchanged through shared memory): dL/dy
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
// Epilogues: warp w reads TMEM lanes 32*(w%4).., columns 32*(w/4)..; one 32-column chunk per thread.
// Bias / LayerNorm-parameter gradients are column sums over the bf16 tiles (thread = (column, row quarter)),
// taken while the MMAs run.  Per-CTA partial gradients are written once at the end and reduced in CTA order.
// The MMAs of a phase are issued by the elected lane of warp 0 inside a warp-uniform branch (umma.cuh: elect_one),
// which keeps the descriptors in uniform registers; every staging phase issues all its global loads before its
// first shared-memory store.
#include <stdlib.h>
#include "umma_block.cuh"

namespace aero {

constexpr int BWD_THREADS = 512;

// Optional phase timing (build with -DAERO_PHASE_TIMING): one observer thread of CTA 0 accumulates clock64 deltas per
// phase into g_phase[]; read back with aero_debug_phase_read().  Compiled out of the production library.
#ifdef AERO_PHASE_TIMING
__device__ long long g_phase[32];
#ifndef AERO_PHASE_TID
#define AERO_PHASE_TID 64   // observer thread; 0 = the MMA-issuing lane (then phases 15..18 time acquire / issue)
#endif
#define PHASE_INIT() long long _pt = clock64(); const bool _obs = (blockIdx.x == 0 && threadIdx.x == AERO_PHASE_TID)
#define PHASE(k) do { if (_obs) { long long _n = clock64(); g_phase[k] += _n - _pt; _pt = _n; } } while (0)
#else
#define PHASE_INIT() do { } while (0)
#define PHASE(k) do { } while (0)
#endif

__device__ __forceinline__ int h_tile(int m) { return m == 0 ? 1 : (m == 1 ? 0 : m); }   // tile holding H_m

template <bool RELU>
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
  float* red = reinterpret_cast<float*>(sidx1 + 128);          // [4 chunks][128 rows] float4 (LayerNorm row sums)
  uint64_t* mbar = reinterpret_cast<uint64_t*>(red + 2048);    // [0] mma, [1..2] weight slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int q = wid & 3;         // TMEM lane quarter
  const int ch = wid >> 2;       // 32-column chunk owned in epilogues
  const int row = q * 32 + lane;
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
  const int act = RELU ? AERO_ACT_RELU : a.act;   // compile-time ReLU: the transcendental paths are not even in the binary

  // ---- weight streaming + MMA issue belong to warp 0 (a warp-uniform branch; one elected lane touches the hardware) ----
  const bool w0 = __shfl_sync(0xffffffffu, wid, 0) == 0;
  int slot_mat[2] = {-1, -1};
  bool slot_pending[2] = {false, false};
  uint32_t slot_phase[2] = {0, 0};
  auto prefetch = [&](int m) {
    int s = m & 1;
    if (slot_mat[s] == m) return;
    uint32_t bar = smem_u32(&mbar[1 + s]);
    if (slot_pending[s]) {   // a transfer nobody waited for: consume its phase before the barrier is re-armed
      mbar_wait(bar, slot_phase[s]);
      slot_phase[s] ^= 1;
      slot_pending[s] = false;
    }
    if (elect_one()) {
      mbar_expect_tx(bar, TILE_BYTES);
      bulk_g2s(w_s + (uint32_t)s * TILE_BYTES, a.prep + (size_t)m * TILE_BYTES, TILE_BYTES, bar);
    }
    __syncwarp();
    slot_mat[s] = m;
    slot_pending[s] = true;
  };
  auto acquire = [&](int m) -> uint32_t {
    int s = m & 1;
    if (slot_mat[s] != m) prefetch(m);
    if (slot_pending[s]) {
      mbar_wait(smem_u32(&mbar[1 + s]), slot_phase[s]);
      slot_phase[s] ^= 1;
      slot_pending[s] = false;
    }
    return w_s + (uint32_t)s * TILE_BYTES;
  };
  if (w0) {   // the first two matrices the recompute needs (layer 0 is skipped when h_0 was kept)
    const int m0 = a.h0 ? 1 : 0;
    prefetch(m0);
    if (m0 + 1 <= L + 1) prefetch(m0 + 1);
  }

  // incoming gradient of a tile, coalesced: tile[r] = bf16( g_out[row0+r] (+ g_agg[receiver of row]) )
  auto stage_gtot = [&](uint8_t* tile, int64_t row0, int nrows) {
    // receiver ids come from the staged index list; every global load is issued before the first shared-memory store
    const int chunk = tid & 15;
    uint4 gv[4];
    float4 ga[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (tid >> 4) + i * 32;
      gv[i] = make_uint4(0u, 0u, 0u, 0u);
      ga[i][0] = ga[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) {
        gv[i] = *reinterpret_cast<const uint4*>(a.g_out + (row0 + r) * 128 + chunk * 8);
        const int n = a.g_agg ? sidx1[r] : -1;
        if (n >= 0) {
          const float* gp = a.g_agg + (size_t)n * 128 + chunk * 8;
          ga[i][0] = *reinterpret_cast<const float4*>(gp);
          ga[i][1] = *reinterpret_cast<const float4*>(gp + 4);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (tid >> 4) + i * 32;
      uint4 v = gv[i];
      v.x = pack_bf16(bf16_lo(v.x) + ga[i][0].x, bf16_hi(v.x) + ga[i][0].y);
      v.y = pack_bf16(bf16_lo(v.y) + ga[i][0].z, bf16_hi(v.y) + ga[i][0].w);
      v.z = pack_bf16(bf16_lo(v.z) + ga[i][1].x, bf16_hi(v.z) + ga[i][1].y);
      v.w = pack_bf16(bf16_lo(v.w) + ga[i][1].z, bf16_hi(v.w) + ga[i][1].w);
      *reinterpret_cast<uint4*>(tile + tile_chunk_off(r, chunk)) = v;
    }
  };

  uint32_t phase = 0;
  bool first_tile = true;
  float db[UMMA_MAX_L_BWD + 1];
#pragma unroll
  for (int i = 0; i <= UMMA_MAX_L_BWD; ++i) db[i] = 0.f;
  float dgam = 0.f, dbet = 0.f, db0 = 0.f;

  const int64_t tiles = (a.rows + 127) / 128;
  PHASE_INIT();
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    const int nrows = (int)((a.rows - row0) < 128 ? (a.rows - row0) : 128);
    const bool valid = row < nrows;
    __syncthreads();   // previous tile fully consumed
    PHASE(0);
    // ---- stage main rows, gather indices, and the incoming gradient tile (g_out + g_agg[receiver]) ----
    if (tid < 128) {
      int64_t r = row0 + tid;
      bool ok = tid < nrows;
      sidx0[tid] = ok ? (a.idx0 ? a.idx0[r] : (int)r) : 0;
      sidx1[tid] = ok ? (a.idx1 ? a.idx1[r] : -1) : -1;
    }
    // rows staged for the recompute: `main` into tile 0, or the kept h_0 rows straight into the tile of H_0
    const bool have_h0 = a.h0 != nullptr;
    const __nv_bfloat16* rows_bf16 = have_h0 ? a.h0 : reinterpret_cast<const __nv_bfloat16*>(a.main);
    uint8_t* Rt = have_h0 ? X + (size_t)h_tile(0) * TILE_BYTES : X;
    if (a.main_f32 && !have_h0) {
      stage_rows<true, BWD_THREADS>(X, a.main, a.main_scale, row0, nrows, tid);
      if (a.g_agg) __syncthreads();   // stage_gtot reads the receiver ids staged just above
      stage_gtot(G, row0, nrows);
    } else {
      // edge block: issue every global load of the tile (main rows, incoming gradient rows, receiver ids, then the
      // receiver gradient rows) before the first shared-memory store, so their latencies overlap
      const int chunk = tid & 15;
      uint4 mv[4], gv[4];
      int dn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (tid >> 4) + i * 32;
        mv[i] = gv[i] = make_uint4(0u, 0u, 0u, 0u);
        dn[i] = -1;
        if (r < nrows) {
          mv[i] = *reinterpret_cast<const uint4*>(rows_bf16 + (row0 + r) * 128 + chunk * 8);
          gv[i] = *reinterpret_cast<const uint4*>(a.g_out + (row0 + r) * 128 + chunk * 8);
          if (a.g_agg) dn[i] = a.idx1[row0 + r];
        }
      }
      float4 ga[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ga[i][0] = ga[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dn[i] >= 0) {
          const float* gp = a.g_agg + (size_t)dn[i] * 128 + chunk * 8;
          ga[i][0] = *reinterpret_cast<const float4*>(gp);
          ga[i][1] = *reinterpret_cast<const float4*>(gp + 4);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (tid >> 4) + i * 32;
        *reinterpret_cast<uint4*>(Rt + tile_chunk_off(r, chunk)) = mv[i];
        uint4 v = gv[i];
        if (dn[i] >= 0) {
          v.x = pack_bf16(bf16_lo(v.x) + ga[i][0].x, bf16_hi(v.x) + ga[i][0].y);
          v.y = pack_bf16(bf16_lo(v.y) + ga[i][0].z, bf16_hi(v.y) + ga[i][0].w);
          v.z = pack_bf16(bf16_lo(v.z) + ga[i][1].x, bf16_hi(v.z) + ga[i][1].y);
          v.w = pack_bf16(bf16_lo(v.w) + ga[i][1].z, bf16_hi(v.w) + ga[i][1].w);
        }
        *reinterpret_cast<uint4*>(G + tile_chunk_off(r, chunk)) = v;
      }
    }
    fence_async_smem();
    __syncthreads();
    // pull the next tile's rows into L2 while this one computes (HBM latency off the critical path)
    {
      const int64_t nrow0 = (tile + gridDim.x) * 128;
      if (nrow0 < a.rows) {
        const int64_t r = nrow0 + (tid >> 2);
        if (r < a.rows) {
          const int part4 = tid & 3;
          if (part4 < 2) {
            if (a.main_f32 && !have_h0) {
              prefetch_l2(reinterpret_cast<const float*>(a.main) + r * 128 + part4 * 64);
              prefetch_l2(reinterpret_cast<const float*>(a.main) + r * 128 + part4 * 64 + 32);
            } else {
              prefetch_l2(rows_bf16 + r * 128 + part4 * 64);
            }
          } else {
            prefetch_l2(a.g_out + r * 128 + (part4 - 2) * 64);
          }
        }
      }
    }

    PHASE(1);   // staging
    // ---- forward recompute (from layer 1 when h_0 was kept from the forward) ----
    const int m_first = have_h0 ? 1 : 0;
    for (int m = m_first; m <= L + 1; ++m) {
      if (w0) {
        PHASE(19);  // (lane 0) arrival at the issue point
        uint32_t wa = acquire(m);
        PHASE(15);  // (lane 0) recompute: wait for the weight slot
        fence_after_sync();
        uint32_t a_addr = (m == 0) ? x_s : x_s + (uint32_t)h_tile(m - 1) * TILE_BYTES;
        if (elect_one()) {
          issue_gemm(tmem_base, a_addr, false, wa, false, false);
          mma_commit(bar_mma);
        }
        __syncwarp();
        if (m + 1 <= L + 1) prefetch(m + 1);   // other slot: last read by GEMM m-1, already complete
        PHASE(16);  // (lane 0) recompute: MMA issue + next weight prefetch
      }
      if (m == 0) {
        // under the first MMA: the coalesced gather P_s[src] + P_d[dst] into the (still free) tile that will hold
        // H_0, and d(beta) = column sums of the incoming gradient
        stage_gather_sum<BWD_THREADS>(X + (size_t)h_tile(0) * TILE_BYTES, a.P, a.ldp, a.poff0, a.poff1, sidx0, sidx1, nrows, tid);
      }
      if (m == m_first) dbet += tile_col_sums_512(G, wid, lane);
      PHASE(2);   // fwd: issue + column sums
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      fence_after_sync();
      PHASE(3);   // fwd: wait for MMA
      if (m == 0) {
        // h0 in place over the gathered pre-projection rows
        uint8_t* Ht = X + (size_t)h_tile(0) * TILE_BYTES;
        __syncthreads();
        first_epilogue_chunk(tlane, ch, act, Ht, row);
        fence_before_sync();
        fence_async_smem();
        __syncthreads();
      } else if (m <= L) {
        uint8_t* Ht = X + (size_t)h_tile(m) * TILE_BYTES;
        hidden_epilogue_chunk(tlane, ch, nullptr, nullptr, vec + (m - 1) * 128, act, Ht, row);
        fence_before_sync();
        fence_async_smem();
        __syncthreads();
      }
    }
    PHASE(4);   // fwd epilogues (+sync)
    // ---- LayerNorm backward; G holds g = dL/d(out) and ends up holding dL/dy ----
    {
      float v[32];
      tmem_ld32(tlane + (uint32_t)(ch * 32), v);
      fence_before_sync();
      const float* bo = vec + L * 128 + ch * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b4 = *reinterpret_cast<const float4*>(bo + 4 * j);
        v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
      }
      if (a.use_ln) {
        const float* gam = vec + (L + 1) * 128 + ch * 32;
        // this thread's 32 gradient values (4 x 16 B of the G tile), kept packed
        uint32_t gp[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 g4 = *reinterpret_cast<const uint4*>(G + tile_chunk_off(row, ch * 4 + j));
          gp[4 * j] = g4.x; gp[4 * j + 1] = g4.y; gp[4 * j + 2] = g4.z; gp[4 * j + 3] = g4.w;
        }
        // one pass, four row sums: S1 = sum y, S2 = sum y^2, A = sum g*gamma, B = sum g*gamma*y
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f, aa = 0.f, ab = 0.f, ba = 0.f, bb = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float ylo = v[2 * j], yhi = v[2 * j + 1];
          const float wlo = bf16_lo(gp[j]) * gam[2 * j], whi = bf16_hi(gp[j]) * gam[2 * j + 1];
          s1a += ylo; s1b += yhi;
          s2a = fmaf(ylo, ylo, s2a); s2b = fmaf(yhi, yhi, s2b);
          aa += wlo; ab += whi;
          ba = fmaf(wlo, ylo, ba); bb = fmaf(whi, yhi, bb);
        }
        reinterpret_cast<float4*>(red)[ch * 128 + row] = make_float4(s1a + s1b, s2a + s2b, aa + ab, ba + bb);
        __syncthreads();
        float S1 = 0.f, S2 = 0.f, A = 0.f, B = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 t4 = reinterpret_cast<const float4*>(red)[k * 128 + row];
          S1 += t4.x; S2 += t4.y; A += t4.z; B += t4.w;
        }
        const float mean = S1 * (1.f / 128.f);
        const float rstd = rsqrtf(fmaxf(S2 * (1.f / 128.f) - mean * mean, 0.f) + 1e-5f);
        const float m1 = A * (1.f / 128.f);
        const float m2 = rstd * (B - mean * A) * (1.f / 128.f);   // mean of g*gamma*yhat
        // dL/dy back into the G tile (same thread, same bytes); z = g*yhat stays in registers for d(gamma)
        float z[32];
        uint32_t op[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float glo = bf16_lo(gp[j]), ghi = bf16_hi(gp[j]);
          const float hlo = (v[2 * j] - mean) * rstd, hhi = (v[2 * j + 1] - mean) * rstd;
          z[2 * j] = glo * hlo;
          z[2 * j + 1] = ghi * hhi;
          op[j] = pack_bf16(rstd * (glo * gam[2 * j] - m1 - hlo * m2), rstd * (ghi * gam[2 * j + 1] - m1 - hhi * m2));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(G + tile_chunk_off(row, ch * 4 + j)) = make_uint4(op[4 * j], op[4 * j + 1], op[4 * j + 2], op[4 * j + 3]);
        // d(gamma): column sums of z over the warp's 32 rows by a shuffle transpose-reduce (lane l ends with the
        // sum of column 32*ch + l); fixed order -> deterministic.  No shared memory, no barrier.
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) {
          const bool upper = (lane & sft) != 0;
#pragma unroll
          for (int i = 0; i < sft; ++i) {
            const float keep = upper ? z[i + sft] : z[i];
            const float send = upper ? z[i] : z[i + sft];
            z[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
          }
        }
        dgam += z[0];
      }
      // use_ln == 0: dL/dy = g, already in G
      fence_async_smem();
      __syncthreads();
    }
    PHASE(5);   // LayerNorm backward
    // ---- backward through the Linear layers ----
    uint8_t* Gc = G;
    uint32_t gc_s = g_s;
    for (int m = L + 1; m >= 0; --m) {
      if (w0) {
        PHASE(19);
        uint32_t wa = acquire(m);
        PHASE(17);  // (lane 0) backward: wait for the weight slot
        fence_after_sync();
        if (elect_one()) {
          if (m >= 1) {
            uint32_t h_addr = x_s + (uint32_t)h_tile(m - 1) * TILE_BYTES;
            issue_gemm(tmem_base + (uint32_t)(128 * m), gc_s, true, h_addr, true, !first_tile);   // dW_m += G^T H
          }
          issue_gemm(tmem_base, gc_s, false, wa, true, false);                                     // G W_m
          mma_commit(bar_mma);
        }
        __syncwarp();
        if (m >= 1) prefetch(m - 1);
        PHASE(18);  // (lane 0) backward: MMA issue + next weight prefetch
      }
      PHASE(14);  // bwd: (thread 0: MMA issue)
      if (m >= 1) db[m - 1] += tile_col_sums_512(Gc, wid, lane);   // bias gradient of Linear m
      if (m == 0) {
        db0 += tile_col_sums_512(Gc, wid, lane);                   // gradient of the first Linear's bias
        unstage_rows<BWD_THREADS>(Gc, a.g_h0, row0, nrows, tid);   // g_h0 leaves while the last GEMM runs
        if (a.has_resid_grad) stage_gtot(G, row0, nrows);          // residual gradient, coalesced, into the free G tile
      }
      PHASE(6);   // bwd: issue + column sums + g_h0 store
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      fence_after_sync();
      PHASE(7);   // bwd: wait for MMA
      if (m >= 1) {
        uint8_t* Ht = X + (size_t)h_tile(m - 1) * TILE_BYTES;
        float v[32];
        tmem_ld32(tlane + (uint32_t)(ch * 32), v);
        PHASE(10);  // bwd: tcgen05.ld
        mask_by_act_grad(v, Ht, row, ch, act);
        store_row32(Ht, row, ch, v);   // in place: G_{m-1} over H_{m-1}
        PHASE(11);  // bwd: mask + pack + store
        fence_before_sync();
        fence_async_smem();
        PHASE(12);  // bwd: fences
        __syncthreads();
        PHASE(13);  // bwd: barrier
        Gc = Ht;
        gc_s = x_s + (uint32_t)h_tile(m - 1) * TILE_BYTES;
      } else {
        // g_main = G_0 W_main (* scale) (+ residual gradient)
        if (a.has_resid_grad) __syncthreads();   // staged residual gradient visible
        float v[32];
        tmem_ld32(tlane + (uint32_t)(ch * 32), v);
        fence_before_sync();
        if (valid) {
          if (a.main_scale) {
            const float sc = a.main_scale[row0 + row];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= sc;
          }
          if (a.has_resid_grad) add_tile_chunk(v, G, row, ch);
          if (a.main_f32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.g_main) + (row0 + row) * 128 + ch * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        if (!a.main_f32) {
          // bf16 gradient rows leave through tile 0 (free since m = 1) so the global store is coalesced
          store_row32(X, row, ch, v);
          __syncthreads();
          unstage_rows<BWD_THREADS>(X, reinterpret_cast<__nv_bfloat16*>(a.g_main), row0, nrows, tid);
        }
      }
    }
    PHASE(8);   // bwd epilogues (+sync)
    first_tile = false;
  }

  // ---- flush per-CTA partial gradients: dW_m from TMEM, vectors from registers (row quarters summed in order) ----
  __syncthreads();
  fence_after_sync();
  const PackedLayout pl{L};
  float* part_out = a.w_part + (size_t)blockIdx.x * pl.total();
  for (int m = 1; m <= L + 1; ++m) {
    float* dst = part_out + (m == L + 1 ? pl.w_out() : pl.w_hidden(m - 1)) + (size_t)row * 128 + ch * 32;
    float v[32];
    tmem_ld32(tlane + (uint32_t)(128 * m + ch * 32), v);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  fence_before_sync();
  // vectors: lanes with (lane & 3) == 0 own one column each (tile_col_sums_512)
  if ((lane & 3) == 0) {
    const int c = col_of_lane_512(wid, lane);
    for (int l = 0; l < L; ++l) part_out[pl.b_hidden(l) + c] = db[l];
    part_out[pl.b_out() + c] = db[L];
    part_out[pl.beta() + c] = dbet;
    part_out[pl.bias0() + c] = db0;
  }
  __syncthreads();
  red[q * 128 + ch * 32 + lane] = dgam;   // partial over the rows of lane quarter q
  __syncthreads();
  if (tid < 128) part_out[pl.gamma() + tid] = (red[tid] + red[128 + tid]) + (red[256 + tid] + red[384 + tid]);
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tmem_base);
}

// ---- host side ------------------------------------------------------------------------------------
static size_t bwd_smem(int L) {
  int nt = (L + 2) > 3 ? (L + 2) : 3;
  return 1024 + (size_t)(2 + nt) * TILE_BYTES + (size_t)(L + 3) * 512 + 1024 + 8192 + 3 * 8 + 16;
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
  int grid = bwd_grid_umma(d->rows);
  // TMA-fed variant (block_umma_bwd2.cu) whenever the forward kept h_0; AERO_BWD_V1=1 keeps the first-generation kernel
  const char* v1_env = getenv("AERO_BWD_V1");
  const bool force_v1 = v1_env && v1_env[0] == '1';
  if (!force_v1 && umma_bwd2_applicable(d)) {
    int rc = umma_block_bwd2(d, a, grid, st);
    if (rc) return rc;
    return launch_reduce_partials(a.w_part + off, grid, pl.total(), d->g_w + off, pl.total() - off, st);
  }
  // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute
  static bool attr_set[64] = {false};
  int dev = 0;
  AERO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    AERO_CUDA(cudaFuncSetAttribute(umma_block_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)bwd_smem(UMMA_MAX_L_BWD)));
    AERO_CUDA(cudaFuncSetAttribute(umma_block_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)bwd_smem(UMMA_MAX_L_BWD)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  if (d->act == AERO_ACT_RELU) umma_block_bwd_kernel<true><<<grid, BWD_THREADS, bwd_smem(d->L), st>>>(a);
  else umma_block_bwd_kernel<false><<<grid, BWD_THREADS, bwd_smem(d->L), st>>>(a);
  AERO_LAUNCH_CHECK();
  return launch_reduce_partials(a.w_part + off, grid, pl.total(), d->g_w + off, pl.total() - off, st);
}

}  // namespace aero

extern "C" int aero_has_umma_bwd(void) { return 1; }

#ifdef AERO_PHASE_TIMING
extern "C" int aero_debug_phase_read(long long* out_host, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out_host, aero::g_phase, sizeof(long long) * 32);
  if (reset) {
    long long z[32] = {0};
    cudaMemcpyToSymbol(aero::g_phase, z, sizeof(z));
  }
  return 0;
}
#endif
