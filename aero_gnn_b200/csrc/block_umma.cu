// block_umma.cu -- fused MGN block, forward, on the 5th-generation tensor cores (tcgen05.mma, bf16 operands, fp32
// accumulators in TMEM).  bf16 storage path of aero_block_fwd.
//
// One persistent CTA per SM, 512 threads = two independent 256-thread tile groups.  A group owns one 128-row tile:
//   * all (L+2) weight matrices of the block stay in shared memory for the whole kernel as bf16 SWIZZLE_128B row
//     tiles (umma.cuh), loaded once per CTA;
//   * the group stages its rows into its activation tile; the elected lane of the group's first warp (warp-uniform
//     branch, descriptors in uniform registers) issues the 8 tcgen05.mma (K = 16 each)
//     of a 128x128x128 GEMM into the group's 128 TMEM columns and commits to an mbarrier;
//   * epilogue: thread = (row, 64-column half) -- warp w reads TMEM lanes 32*(w%4).., columns 64*(w/4).. with
//     tcgen05.ld, adds the gathered pre-projections (layer 0) or the bias, applies the activation, packs to bf16
//     and writes the next GEMM's A operand into the same activation tile (the MMA that read it has completed);
//   * last epilogue: bias + LayerNorm (fp32 row statistics, the two column halves exchange partial sums through
//     shared memory) + residual, bf16 pack into the tile; the tile is then stored to HBM with coalesced 16-byte
//     chunks and reduced per receiver (one warp per CSR segment, fp32, fixed order) into agg;
//   * while one group is in an epilogue, the other group's MMAs run: 16 resident warps per SM keep both the
//     tensor pipe and the LSU/ALU pipes busy.  Epilogue loops are chunked (32 columns) and not unrolled across
//     chunks so the instruction footprint stays inside the instruction cache.
#include "umma_block.cuh"
#include "tma.cuh"

namespace aero {

constexpr int FWD_GROUPS = 2;
constexpr int FWD_GT = 256;                       // threads per tile group
constexpr int FWD_THREADS = FWD_GROUPS * FWD_GT;

// ---- weight images ------------------------------------------------------------------------------
__global__ void umma_prepare_kernel(const float* __restrict__ w, int L, uint8_t* __restrict__ prep) {
  const PackedLayout pl{L};
  const int nm = L + 2;
  int total = nm * 128 * 16;  // (matrix, row, chunk)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int m = i / 2048, o = (i % 2048) / 16, c = i % 16;
    const float* src = w + (size_t)m * 16384 + (size_t)o * 128 + c * 8;
    uint4 v;
    v.x = pack_bf16(src[0], src[1]);
    v.y = pack_bf16(src[2], src[3]);
    v.z = pack_bf16(src[4], src[5]);
    v.w = pack_bf16(src[6], src[7]);
    *reinterpret_cast<uint4*>(prep + (size_t)m * TILE_BYTES + tile_chunk_off(o, c)) = v;
  }
  float* vec = reinterpret_cast<float*>(prep + (size_t)nm * TILE_BYTES);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (L + 3) * 128; i += gridDim.x * blockDim.x)
    vec[i] = w[pl.b_hidden(0) + i];
}

// =============================================================================================
// forward
// =============================================================================================
// tm_main / tm_resid / tm_out / tm_h0: SWIZZLE_128B tensor maps of the bf16 row matrices a.main, a.resid, a.out, a.h0
// (tma.cuh); a map whose matrix is absent (fp32 main, no residual, no h0) is a copy of tm_out and never used.
template <bool RELU>
__global__ void __launch_bounds__(FWD_THREADS, 1)
umma_block_fwd_kernel(UmmaArgs a, const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_resid,
                      const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_h0,
                      const __grid_constant__ CUtensorMap tm_mlat, const __grid_constant__ CUtensorMap tm_h1,
                      const __grid_constant__ CUtensorMap tm_h2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int L = a.L;
  uint8_t* Wimg = smem;
  uint8_t* Abuf = Wimg + (size_t)(L + 2) * TILE_BYTES;
  float* vec = reinterpret_cast<float*>(Abuf + (size_t)FWD_GROUPS * TILE_BYTES);
  int* sidx = reinterpret_cast<int*>(vec + (L + 3) * 128);                     // [group][2][128]
  float2* red = reinterpret_cast<float2*>(sidx + FWD_GROUPS * 256);            // [group][2][128] LN partials
  uint32_t* segmask = reinterpret_cast<uint32_t*>(red + FWD_GROUPS * 256);     // [group][8]: 4 masks + 2 flags
  int* segs = reinterpret_cast<int*>(segmask + FWD_GROUPS * 8);                // [group][132]: heads + count
  uint64_t* mbar = reinterpret_cast<uint64_t*>(segs + FWD_GROUPS * 132);       // [group][3]: mma, main rows, residual rows
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3 * FWD_GROUPS);

  const int tid = threadIdx.x, grp = tid / FWD_GT, gt = tid % FWD_GT, lane = tid & 31;
  const int gw = gt >> 5;            // warp inside the group, 0..7
  const int q = gw & 3;              // TMEM lane quarter (== warp id % 4)
  const int hf = gw >> 2;            // column half
  const int row = q * 32 + lane;
  const bool gw0 = __shfl_sync(0xffffffffu, gw, 0) == 0;
  // weights + vectors, once per CTA
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.prep);
    uint4* dst = reinterpret_cast<uint4*>(Wimg);
    const int n16 = (L + 2) * (TILE_BYTES / 16);
    for (int i = tid; i < n16; i += FWD_THREADS) dst[i] = src[i];
    const float* vs = reinterpret_cast<const float*>(a.prep + (size_t)(L + 2) * TILE_BYTES);
    for (int i = tid; i < (L + 3) * 128; i += FWD_THREADS) vec[i] = vs[i];
  }
  if (tid == 0) {
    for (int w = 0; w < 3 * FWD_GROUPS; ++w) mbar_init(smem_u32(&mbar[w]), 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<FWD_GROUPS * 128>(tmem_slot);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tacc = tmem_base + (uint32_t)grp * 128u;
  const uint32_t tlane = tacc + ((uint32_t)(q * 32) << 16);
  uint8_t* A = Abuf + (size_t)grp * TILE_BYTES;
  const uint32_t a_s = smem_u32(A);
  const uint32_t w_s = smem_u32(Wimg);
  const uint32_t bar_s = smem_u32(&mbar[3 * grp]);
  const uint32_t bar_in = smem_u32(&mbar[3 * grp + 1]), bar_res = smem_u32(&mbar[3 * grp + 2]);
  uint32_t ph_in = 0, ph_res = 0;
  const bool tma_main = !a.main_f32;
  int* sidx0 = sidx + grp * 256;
  int* sidx1 = sidx0 + 128;
  float2* gred = red + grp * 256;
  uint32_t* gmask = segmask + grp * 8;
  int* gseg = segs + grp * 132;
  const int bar_id = 1 + grp;
  uint32_t phase = 0;
  const int act = RELU ? AERO_ACT_RELU : a.act;

  const int64_t tiles = (a.rows + 127) / 128;
  // gather indices (and the two receiver ids across the tile edges) of a tile are fetched while the PREVIOUS tile of
  // the group computes, so no tile starts by waiting for an index load
  int nsrc = -1, ndst = -1, nprev = -2, nnext = -2;
  auto fetch_indices = [&](int64_t t) {
    nsrc = ndst = -1;
    nprev = nnext = -2;
    const int64_t r = t * 128 + gt;
    if (t < tiles && gt < 128 && r < a.rows) {
      nsrc = a.idx0 ? __ldg(a.idx0 + r) : (int)r;
      ndst = a.idx1 ? __ldg(a.idx1 + r) : -1;
      if (a.agg) {
        if (lane == 0 && r > 0) nprev = __ldg(a.idx1 + r - 1);
        if ((gt == 127 || r == a.rows - 1) && r + 1 < a.rows) nnext = __ldg(a.idx1 + r + 1);
      }
    }
  };
  fetch_indices((int64_t)blockIdx.x * FWD_GROUPS + grp);
  for (int64_t tile = (int64_t)blockIdx.x * FWD_GROUPS + grp; tile < tiles; tile += (int64_t)gridDim.x * FWD_GROUPS) {
    const int64_t row0 = tile * 128;
    const int nrows = (int)((a.rows - row0) < 128 ? (a.rows - row0) : 128);
    if (gw0) {   // the previous tile's output store has finished reading the activation tile
      if (elect_one()) tma::store_wait_read();
      __syncwarp();
    }
    named_sync(bar_id, FWD_GT);   // previous tile of this group fully consumed
    if (tma_main) {
      // bf16 rows: fetched by the TMA engine straight into the UMMA tile format while the index work below runs;
      // the group's next tile is pulled into L2 at the same time
      if (gw0) {
        if (elect_one()) {
          mbar_expect_tx(bar_in, TILE_BYTES);
          tma::load_tile(a_s, &tm_main, (int)row0, bar_in);
          const int64_t nr0 = (tile + (int64_t)gridDim.x * FWD_GROUPS) * 128;
          if (nr0 < a.rows) tma::prefetch_tile_l2(&tm_main, (int)nr0);
        }
        __syncwarp();
      }
    } else {
      stage_rows<true, FWD_GT>(A, a.main, a.main_scale, row0, nrows, gt);
    }
    if (gt < 128) {
      int64_t r = row0 + gt;
      bool ok = gt < nrows;
      int i0 = ok ? nsrc : 0;
      int i1 = ok ? ndst : -1;
      sidx0[gt] = i0;
      sidx1[gt] = i1;
      if (a.agg) {
        // receiver-segment heads of this tile (rows are in CSR order) + completeness of the two boundary segments
        // neighbours across the tile boundary are read straight from the (receiver-sorted) index list: no load
        // depends on another one
        int prev = __shfl_up_sync(0xffffffffu, i1, 1);
        if (lane == 0) prev = (!ok || r == 0) ? -2 : nprev;
        const int next = (gt == nrows - 1) ? nnext : -2;
        uint32_t m = __ballot_sync(0xffffffffu, ok && (gt == 0 || i1 != prev));
        if (lane == 0) gmask[gw] = m;
        if (gt == 0) gmask[4] = (prev == i1) ? 1u : 0u;              // head segment started in an earlier tile
        if (gt == nrows - 1) gmask[5] = (next == i1) ? 1u : 0u;      // tail segment continues in the next tile
      }
    }
    fence_async_smem();
    named_sync(bar_id, FWD_GT);
    if (a.agg && gt < 128) {
      // compact list of segment heads (row numbers) from the ballot masks
      const uint32_t m0 = gmask[0], m1 = gmask[1], m2 = gmask[2], m3 = gmask[3];
      const uint32_t mine = gw == 0 ? m0 : (gw == 1 ? m1 : (gw == 2 ? m2 : m3));
      if ((mine >> lane) & 1u) {
        int pos = __popc(mine & ((1u << lane) - 1u));
        if (gw > 0) pos += __popc(m0);
        if (gw > 1) pos += __popc(m1);
        if (gw > 2) pos += __popc(m2);
        gseg[pos] = gt;
      }
      if (gt == 0) gseg[128] = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
    }
    // this tile's gathered pre-projection rows -> L1, so the layer-0 epilogue (one GEMM later) does not wait on L2
    if (gt < 128 && gt < nrows) {
      const __nv_bfloat16* ps = a.P + (int64_t)sidx0[gt] * a.ldp + a.poff0;
      prefetch_l1(ps);
      prefetch_l1(ps + 64);
      const int d1 = sidx1[gt];
      if (d1 >= 0 && (gt == 0 || sidx1[gt - 1] != d1)) {
        const __nv_bfloat16* pp = a.P + (int64_t)d1 * a.ldp + a.poff1;
        prefetch_l1(pp);
        prefetch_l1(pp + 64);
      }
    }
    // pull this group's next tile into L2 while the current one computes
    if (a.main_f32) {
      const int64_t r = (tile + (int64_t)gridDim.x * FWD_GROUPS) * 128 + (gt >> 1);
      if (r < a.rows) {
        const float* p = reinterpret_cast<const float*>(a.main) + r * 128 + (gt & 1) * 64;
        prefetch_l2(p);
        prefetch_l2(p + 32);
      }
    }
    // gather indices of the NEXT tile (used at its start; their pre-projected rows are also prefetched into L2 after
    // this tile's first epilogue, so the layer-0 gathers of the next tile do not pay DRAM latency)
    fetch_indices(tile + (int64_t)gridDim.x * FWD_GROUPS);

    for (int layer = 0; layer <= L + 1; ++layer) {
      if (gw0) {   // first warp of the group, warp-uniform branch; one elected lane issues
        if (layer == 0 && tma_main) {
          mbar_wait(bar_in, ph_in);
          ph_in ^= 1;
        }
        fence_after_sync();
        if (elect_one()) {
          issue_gemm(tacc, a_s, false, w_s + (uint32_t)layer * TILE_BYTES, false, false);
          mma_commit(bar_s);
          if (layer == 0 && a.main_lat) {   // the staged (scaled, rounded) fp32 rows leave as a latent-dtype copy
            tma::store_tile(&tm_mlat, a_s, (int)row0);
            tma::store_commit();
          }
        }
        __syncwarp();
      }
      // keep h_0 (and, under the keep-all policy, H_1 / H_2) for the backward: a TMA store reads the activation tile
      // while the GEMM that consumes it does
      const bool keep_this = layer >= 1 && (layer == 1 ? a.h0 != nullptr : (layer <= 3 && a.hh[layer - 2] != nullptr));
      if (keep_this && gw0) {
        if (elect_one()) {
          tma::store_tile(layer == 1 ? &tm_h0 : (layer == 2 ? &tm_h1 : &tm_h2), a_s, (int)row0);
          tma::store_commit();
        }
        __syncwarp();
      }
      mbar_wait(bar_s, phase);
      phase ^= 1;
      fence_after_sync();
      if (keep_this) {
        // the next epilogue overwrites the tile: behind the store's reads (long finished; the barrier is the hand-off)
        if (gw0) {
          if (elect_one()) tma::store_wait_read();
          __syncwarp();
        }
        named_sync(bar_id, FWD_GT);
      }

      if (layer == 0 && a.main_lat) {   // the copy-out above has read the tile before the gather below overwrites it
        if (gw0) {
          if (elect_one()) tma::store_wait_read();
          __syncwarp();
        }
        named_sync(bar_id, FWD_GT);
      }
      if (layer == 0) {
        // the MMA that read A has completed: A now receives the coalesced gather P_s[src] + P_d[dst], then each
        // thread turns its (row, chunk) into h0 in place
        stage_gather_sum<FWD_GT>(A, a.P, a.ldp, a.poff0, a.poff1, sidx0, sidx1, nrows, gt);
        named_sync(bar_id, FWD_GT);
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) first_epilogue_chunk(tlane, hf * 2 + cc, act, A, row);
        fence_before_sync();
        fence_async_smem();
        named_sync(bar_id, FWD_GT);
      } else if (layer <= L) {
        const float* bias = vec + (layer - 1) * 128;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) hidden_epilogue_chunk(tlane, hf * 2 + cc, nullptr, nullptr, bias, act, A, row);
        fence_before_sync();
        fence_async_smem();
        named_sync(bar_id, FWD_GT);
      } else {
        // ---- output epilogue: bias, LayerNorm, residual ----
        const float* bo = vec + L * 128;
        float mean = 0.f, rstd = 1.f;
        // GEMM_out has consumed A: the residual rows are fetched into it by the TMA engine under the statistics pass
        if (a.resid && gw0) {
          if (elect_one()) {
            mbar_expect_tx(bar_res, TILE_BYTES);
            tma::load_tile(a_s, &tm_resid, (int)row0, bar_res);
          }
          __syncwarp();
        }
        if (a.use_ln) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int c = hf * 2 + cc;
            float v[32];
            tmem_ld32(tlane + (uint32_t)(c * 32), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 b4 = *reinterpret_cast<const float4*>(bo + c * 32 + 4 * j);
              float y0 = v[4 * j] + b4.x, y1 = v[4 * j + 1] + b4.y, y2 = v[4 * j + 2] + b4.z, y3 = v[4 * j + 3] + b4.w;
              s0 += y0; s1 += y1; s2 += y2; s3 += y3;
              t0 = fmaf(y0, y0, t0); t1 = fmaf(y1, y1, t1); t2 = fmaf(y2, y2, t2); t3 = fmaf(y3, y3, t3);
            }
          }
          gred[hf * 128 + row] = make_float2((s0 + s1) + (s2 + s3), (t0 + t1) + (t2 + t3));
          named_sync(bar_id, FWD_GT);
          float2 pa = gred[row], pb = gred[128 + row];
          mean = (pa.x + pb.x) * (1.f / 128.f);
          float var = fmaxf((pa.y + pb.y) * (1.f / 128.f) - mean * mean, 0.f);
          rstd = rsqrtf(var + 1e-5f);
        }
        const float* gam = vec + (L + 1) * 128;
        const float* bet = vec + (L + 2) * 128;
        if (a.resid) {
          mbar_wait(bar_res, ph_res);
          ph_res ^= 1;
        }
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = hf * 2 + cc;
          float v[32];
          tmem_ld32(tlane + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b4 = *reinterpret_cast<const float4*>(bo + c * 32 + 4 * j);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
          if (a.use_ln) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 g4 = *reinterpret_cast<const float4*>(gam + c * 32 + 4 * j);
              float4 e4 = *reinterpret_cast<const float4*>(bet + c * 32 + 4 * j);
              v[4 * j] = fmaf((v[4 * j] - mean) * rstd, g4.x, e4.x);
              v[4 * j + 1] = fmaf((v[4 * j + 1] - mean) * rstd, g4.y, e4.y);
              v[4 * j + 2] = fmaf((v[4 * j + 2] - mean) * rstd, g4.z, e4.z);
              v[4 * j + 3] = fmaf((v[4 * j + 3] - mean) * rstd, g4.w, e4.w);
            }
          }
          if (a.resid) add_tile_chunk(v, A, row, c);      // rows beyond nrows were staged as zeros
          store_row32(A, row, c, v);
        }
        fence_before_sync();
        fence_async_smem();
        named_sync(bar_id, FWD_GT);
        // the output tile leaves through a TMA store (rows beyond the matrix are clipped)
        if (gw0) {
          if (elect_one()) {
            tma::store_tile(&tm_out, a_s, (int)row0);
            tma::store_commit();
          }
          __syncwarp();
        }
        if (a.agg) {
          // receiver sums over the bf16-rounded rows: one warp per CSR segment (segments k = gw, gw+8, ..),
          // lane = 4 columns, rows in order
          const uint8_t* base = A + (lane >> 4) * PANEL_BYTES + (lane & 1) * 8;
          const int ch = (lane >> 1) & 7;
          const bool head_open = gmask[4] != 0, tail_open = gmask[5] != 0;
          const int nseg = gseg[128];
          for (int k = gw; k < nseg; k += 8) {
            const int rs = gseg[k];
            const int re = (k + 1 < nseg) ? gseg[k + 1] : nrows;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int t = rs; t < re; ++t) {
              uint2 u = *reinterpret_cast<const uint2*>(base + t * 128 + ((ch ^ (t & 7)) << 4));
              acc.x += bf16_lo(u.x); acc.y += bf16_hi(u.x); acc.z += bf16_lo(u.y); acc.w += bf16_hi(u.y);
            }
            const bool open = (rs == 0 && head_open) || (re == nrows && tail_open);
            float* dst = open ? a.agg_part + ((size_t)tile * 2 + (rs == 0 ? 0 : 1)) * 128
                              : a.agg + (size_t)sidx1[rs] * 128;
            *reinterpret_cast<float4*>(dst + lane * 4) = acc;
          }
        }
      }
    }
  }
  if (gw0) {
    if (elect_one()) tma::store_wait_all();
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<FWD_GROUPS * 128>(tmem_base);
}

// =============================================================================================
// primitive self-test: one 128x128x128 GEMM in each operand orientation
// =============================================================================================
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const __nv_bfloat16* __restrict__ Ag,
                                                               const __nv_bfloat16* __restrict__ Bg, float* __restrict__ C,
                                                               int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* At = smem;
  uint8_t* Bt = smem + TILE_BYTES;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(Bt + TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x;
  stage_rows<false, 128>(At, Ag, nullptr, 0, 128, tid);
  stage_rows<false, 128>(Bt, Bg, nullptr, 0, 128, tid);
  if (tid == 0) {
    mbar_init(smem_u32(mbar), 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<128>(tmem_slot);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tacc = *tmem_slot;
  if (tid == 0) {
    issue_gemm(tacc, smem_u32(At), mode == 2, smem_u32(Bt), mode >= 1, false);
    mma_commit(smem_u32(mbar));
  }
  mbar_wait(smem_u32(mbar), 0);
  fence_after_sync();
  const uint32_t tlane = tacc + ((uint32_t)((tid >> 5) * 32) << 16);
  for (int c = 0; c < 4; ++c) {
    float v[32];
    tmem_ld32(tlane + (uint32_t)(c * 32), v);
    for (int j = 0; j < 32; ++j) C[(size_t)tid * 128 + c * 32 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<128>(tacc);
}

// ---- host side ------------------------------------------------------------------------------------
size_t umma_prepared_bytes(int L) { return (size_t)(L + 2) * TILE_BYTES + (size_t)(L + 3) * 128 * sizeof(float); }

int umma_prepare(const float* w, int L, void* prepared, cudaStream_t st) {
  if (L > UMMA_MAX_L) {
    set_error("umma_prepare: L=%d > %d (use the CUDA-core path)", L, UMMA_MAX_L);
    return AERO_EUNSUPPORTED;
  }
  umma_prepare_kernel<<<32, 256, 0, st>>>(w, L, reinterpret_cast<uint8_t*>(prepared));
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

static size_t fwd_smem(int L) {
  return 1024 + (size_t)(L + 2 + FWD_GROUPS) * TILE_BYTES + (size_t)(L + 3) * 512 +
         (size_t)FWD_GROUPS * (1024 + 2048 + 32 + 528 + 24) + 16;
}

size_t umma_block_workspace_bytes(const aero_block_desc* d, int backward) {
  if (!backward) {
    int64_t tiles = cdiv(d->rows > 0 ? d->rows : 1, 128);
    return d->agg ? align_up((size_t)tiles * 2 * 128 * sizeof(float), 256) : 256;
  }
  return umma_bwd_workspace_bytes(d);
}

int umma_block_fwd(const aero_block_desc* d, cudaStream_t st) {
  if (d->dtype != AERO_BF16 || d->L > UMMA_MAX_L) {
    set_error("umma_block_fwd: needs bf16 rows and L <= %d", UMMA_MAX_L);
    return AERO_EUNSUPPORTED;
  }
  UmmaArgs a = make_uargs(d);
  if (d->agg) {
    a.agg_part = reinterpret_cast<float*>(d->workspace);
    // no memset of the aggregate: every receiver with rows is written by its tile (or by the fix-up when its run
    // straddles tiles), and the fix-up launch zeroes the receivers without rows
    if (d->rows == 0 && !(d->flags & AERO_BLOCK_AGG_NO_CLEAR))
      AERO_CUDA(cudaMemsetAsync(d->agg, 0, (size_t)d->n_nodes * 128 * sizeof(float), st));
  }
  if (d->rows == 0) return AERO_OK;
  // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute
  static bool attr_set[64] = {false};
  int dev = 0;
  AERO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    AERO_CUDA(cudaFuncSetAttribute(umma_block_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)fwd_smem(UMMA_MAX_L)));
    AERO_CUDA(cudaFuncSetAttribute(umma_block_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)fwd_smem(UMMA_MAX_L)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  int64_t tiles = cdiv(d->rows, 128);
  int64_t want = cdiv(tiles, FWD_GROUPS);
  int grid = (int)(want < sm_count() ? want : sm_count());
  CUtensorMap tm_main, tm_resid, tm_out, tm_h0, tm_mlat, tm_h1, tm_h2;
  if ((d->h_hidden[0] || d->h_hidden[1]) && !(d->L == 2 && d->h0 && d->h_hidden[0] && d->h_hidden[1])) {
    set_error("umma_block_fwd: h_hidden needs L == 2, h0 and both of H_1, H_2");
    return AERO_EINVAL;
  }
  if (tma::make_rows_map(d->h_hidden[0] ? d->h_hidden[0] : d->out, d->rows, &tm_h1) ||
      tma::make_rows_map(d->h_hidden[1] ? d->h_hidden[1] : d->out, d->rows, &tm_h2)) {
    set_error("umma_block_fwd: cuTensorMapEncodeTiled failed (h_hidden)");
    return AERO_ECUDA;
  }
  if (d->main_lat && !d->main_f32) {
    set_error("umma_block_fwd: main_lat is the latent-dtype copy of fp32 main rows (main_f32 = 1 only)");
    return AERO_EINVAL;
  }
  if (tma::make_rows_map(d->out, d->rows, &tm_out) || tma::make_rows_map(d->main_lat ? d->main_lat : d->out, d->rows, &tm_mlat) ||
      tma::make_rows_map(d->main_f32 ? d->out : d->main, d->rows, &tm_main) ||
      tma::make_rows_map(d->resid ? d->resid : d->out, d->rows, &tm_resid) ||
      tma::make_rows_map(d->h0 ? d->h0 : d->out, d->rows, &tm_h0)) {
    set_error("umma_block_fwd: cuTensorMapEncodeTiled failed (row matrices must be 16-byte aligned)");
    return AERO_ECUDA;
  }
  if (d->act == AERO_ACT_RELU)
    umma_block_fwd_kernel<true><<<grid, FWD_THREADS, fwd_smem(d->L), st>>>(a, tm_main, tm_resid, tm_out, tm_h0, tm_mlat, tm_h1, tm_h2);
  else
    umma_block_fwd_kernel<false><<<grid, FWD_THREADS, fwd_smem(d->L), st>>>(a, tm_main, tm_resid, tm_out, tm_h0, tm_mlat, tm_h1, tm_h2);
  AERO_LAUNCH_CHECK();
  if (d->agg)
    return launch_agg_fixup128(a.agg_part, d->rowptr, d->idx1, d->agg, d->rows, d->n_nodes,
                               !(d->flags & AERO_BLOCK_AGG_NO_CLEAR), st);
  return AERO_OK;
}

}  // namespace aero

using namespace aero;

extern "C" int aero_has_umma(void) { return 1; }

extern "C" int aero_umma_selftest(const void* a_bf16, const void* b_bf16, float* c, int mode, void* stream) {
  AERO_CHECK_ARG(a_bf16 && b_bf16 && c && mode >= 0 && mode <= 2, "aero_umma_selftest: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = 1024 + 2 * umma::TILE_BYTES + 64;
  AERO_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(a_bf16),
                                             reinterpret_cast<const __nv_bfloat16*>(b_bf16), c, mode);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}
