// block_umma_bwd2.cu -- backward of the fused MGN block on tcgen05 tensor cores, TMA-fed variant.
//
// Used when the forward kept the first hidden activation h_0 (aero_block_desc.h0) and 1 <= L <= 2 -- the processor
// configuration (models/mgnLayer.py:177-213 of the reference run backwards).  Same mathematics as block_umma_bwd.cu
// (recompute of layers 1..L+1 from h_0, LayerNorm backward, dW_m += G_m^T H_{m-1} resident in TMEM,
// G_{m-1} = (G_m W_m) * act'(H_{m-1}) in place), different data movement:
//   * the two row tiles a 128-row tile reads (h_0 rows, incoming gradient rows) are fetched by the TMA engine
//     (cp.async.bulk.tensor, SWIZZLE_128B tensor maps) straight into the UMMA tile format, for the NEXT tile while
//     the current one computes, into shared-memory tiles the current tile has already finished with -- the four
//     activation tiles rotate roles from tile to tile; no thread stages a row;
//   * the two row tiles it writes (g_h0 and g_main) leave through TMA stores;
//   * the receiver-gradient rows g_agg[dst] (fp32) are added where the gradient is consumed (LayerNorm backward
//     and the residual add of the last epilogue), each thread reading its own 128-byte piece;
//   * bias / LayerNorm-beta gradients (column sums of bf16 row tiles) are taken by the legacy tensor path:
//     ldmatrix.trans fragments of the tile times an all-ones A operand (mma.sync m16n8k16), 4 + 8 instructions per
//     warp and tile instead of ~90; d(beta)'s receiver part is sum_n deg(n) g_agg[n], one pass at the end;
//   * ReLU masks are applied to the packed bf16x2 words (add + prmt sign replication + and).
#include "umma_block.cuh"
#include "tma.cuh"

namespace aero {

constexpr int B2_THREADS = 512;

// column sums of a bf16 row tile, warp w owns the 8 columns of chunk w: s0 / s1 += sums of columns 8w + 2(lane%4) + {0,1}
// (every lane holds the totals of its lane%4 pair; lanes 0..3 are the designated owners).  Exact: bf16 x 1.0 products
// accumulated in fp32 in row order 0..127.
__device__ __forceinline__ void colsum_mma(const uint8_t* tile, int warp, int lane, float& s0, float& s1) {
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
  const uint32_t ones = 0x3F803F80u;
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    const int r = kb * 32 + lane;
    const uint32_t addr = smem_u32(tile + tile_chunk_off(r, warp));
    uint32_t b0, b1, b2, b3;
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                 : "r"(addr)
                 : "memory");
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, {%0, %1, %2, %3};\n"
        : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
        : "r"(ones), "r"(b0), "r"(b1));
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, {%0, %1, %2, %3};\n"
        : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
        : "r"(ones), "r"(b2), "r"(b3));
  }
  s0 += c0;
  s1 += c1;
}

// packed ReLU mask: keep each bf16 half of v where the matching half of the (non-negative) activation word h is non-zero
__device__ __forceinline__ uint32_t relu_mask_bf16x2(uint32_t v, uint32_t h) {
  const uint32_t t = h + 0x7FFF7FFFu;   // bit 15 / 31 set  <=>  the half is non-zero (h halves are in [0, 0x7FFF])
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBB99;\n" : "=r"(m) : "r"(t));   // replicate those sign bits over their halves
  return v & m;
}

// G_{m-1} = acc * act'(H_{m-1}) for one (row, 32-column chunk), written in place over H_{m-1}
// (the stores wait for `bar`: the weight-gradient GEMM still reads H_{m-1} while the values are computed)
template <bool RELU>
__device__ __forceinline__ void mask_epilogue_chunk(uint32_t taddr, uint8_t* Ht, int row, int c, int act, uint32_t bar,
                                                    uint32_t parity) {
  float v[32];
  tmem_ld32(taddr, v);
  if (RELU) {
    uint4 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 h4 = *reinterpret_cast<const uint4*>(Ht + tile_chunk_off(row, c * 4 + j));
      o[j].x = relu_mask_bf16x2(pack_bf16(v[8 * j + 0], v[8 * j + 1]), h4.x);
      o[j].y = relu_mask_bf16x2(pack_bf16(v[8 * j + 2], v[8 * j + 3]), h4.y);
      o[j].z = relu_mask_bf16x2(pack_bf16(v[8 * j + 4], v[8 * j + 5]), h4.z);
      o[j].w = relu_mask_bf16x2(pack_bf16(v[8 * j + 6], v[8 * j + 7]), h4.w);
    }
    mbar_wait(bar, parity);
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(Ht + tile_chunk_off(row, c * 4 + j)) = o[j];
  } else {
    mask_by_act_grad(v, Ht, row, c, act);
    mbar_wait(bar, parity);
    store_row32(Ht, row, c, v);
  }
}

// KEPT: the forward also kept H_1 .. H_L (aero_block_desc.h_hidden, L == 2): they are fetched by the TMA engine like
// h_0 and the recompute of the hidden layers disappears -- two GEMM phases and two epilogues per tile less, for
// +256 B/row/layer kept from the forward and read here.
template <bool RELU, bool KEPT, bool F32T>   // F32T: fp32 g_main rows leave through TMA (node block, !KEPT)
__global__ void __launch_bounds__(B2_THREADS, 1)
umma_block_bwd2_kernel(UmmaArgs a, const __grid_constant__ CUtensorMap tm_h0, const __grid_constant__ CUtensorMap tm_gout,
                       const __grid_constant__ CUtensorMap tm_gh0, const __grid_constant__ CUtensorMap tm_gmain,
                       const __grid_constant__ CUtensorMap tm_h1, const __grid_constant__ CUtensorMap tm_h2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int L = a.L;
  uint8_t* Wslot = smem;                                    // 2 tiles
  uint8_t* X = Wslot + 2 * TILE_BYTES;                      // 4 activation tiles, roles rotate
  float* vec = reinterpret_cast<float*>(X + (size_t)4 * TILE_BYTES);
  float* red = vec + (UMMA_MAX_L_BWD + 3) * 128;             // [4 chunks][128 rows] float4 (LayerNorm row sums)
  uint64_t* mbar = reinterpret_cast<uint64_t*>(red + 2048);  // [0] mma, [1..2] weight slots, [3] h0, [4] g_out, [5] reload, [6] dW, [7] h1, [8] h2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 9);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int q = wid & 3;         // TMEM lane quarter
  const int ch = wid >> 2;       // 32-column chunk owned in epilogues
  const int row = q * 32 + lane;
  {
    const float* vs = reinterpret_cast<const float*>(a.prep + (size_t)(L + 2) * TILE_BYTES);
    for (int i = tid; i < (L + 3) * 128; i += B2_THREADS) vec[i] = vs[i];
  }
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(smem_u32(&mbar[i]), 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32);
  const uint32_t bar_mma = smem_u32(&mbar[0]);
  const uint32_t bar_h0 = smem_u32(&mbar[3]), bar_g = smem_u32(&mbar[4]), bar_r = smem_u32(&mbar[5]), bar_dw = smem_u32(&mbar[6]);
  const uint32_t bar_h1 = smem_u32(&mbar[7]), bar_h2 = smem_u32(&mbar[8]);
  const uint32_t x_s = smem_u32(X);
  const uint32_t w_s = smem_u32(Wslot);
  const int act = RELU ? AERO_ACT_RELU : a.act;
  const bool resid = a.has_resid_grad != 0;
  const bool out_tma = !a.main_f32;

  // ---- weight streaming, TMA and MMA issue belong to warp 0 (warp-uniform branch; one elected lane touches the hardware) ----
  const bool w0 = __shfl_sync(0xffffffffu, wid, 0) == 0;
  // slot s = m & 1 holds matrix m.  State of both slots in one register (no local-memory arrays): per slot 3 bits
  // (resident matrix + 1), 1 bit (transfer in flight), 1 bit (mbarrier phase); slot s at bit 8*s.
  uint32_t wst = 0;
  auto prefetch = [&](int m) {
    const int s = m & 1, sh = 8 * s;
    if ((int)((wst >> sh) & 7u) == m + 1) return;
    const uint32_t bar = smem_u32(&mbar[1 + s]);
    if ((wst >> (sh + 3)) & 1u) {   // a transfer nobody waited for: consume its phase before the barrier is re-armed
      mbar_wait(bar, (wst >> (sh + 4)) & 1u);
      wst ^= 1u << (sh + 4);
    }
    if (elect_one()) {
      mbar_expect_tx(bar, TILE_BYTES);
      bulk_g2s(w_s + (uint32_t)s * TILE_BYTES, a.prep + (size_t)m * TILE_BYTES, TILE_BYTES, bar);
    }
    __syncwarp();
    wst = (wst & ~(15u << sh)) | ((uint32_t)(m + 1) << sh) | (8u << sh);
  };
  auto acquire = [&](int m) -> uint32_t {
    const int s = m & 1, sh = 8 * s;
    if ((int)((wst >> sh) & 7u) != m + 1) prefetch(m);
    if ((wst >> (sh + 3)) & 1u) {
      mbar_wait(smem_u32(&mbar[1 + s]), (wst >> (sh + 4)) & 1u);
      wst ^= 1u << (sh + 4);
      wst &= ~(8u << sh);
    }
    return w_s + (uint32_t)s * TILE_BYTES;
  };

  // tile roles: pH[m] holds H_m (later G_m), pG the incoming gradient (later dL/dy, later the reloaded gradient / output)
  int pH0 = 0, pH1 = 1, pH2 = 2, pG = 3;   // L == 1: pH2 is a spare tile
  auto tile_of_h = [&](int m) { return m == 0 ? pH0 : (m == 1 ? pH1 : pH2); };

  const int64_t tiles = (a.rows + 127) / 128;
  if (w0) {
    if ((int64_t)blockIdx.x < tiles) {
      if (elect_one()) {
        tma::prefetch_map(&tm_h0);
        tma::prefetch_map(&tm_gout);
        tma::prefetch_map(&tm_gh0);
        if (out_tma || F32T) tma::prefetch_map(&tm_gmain);
        const int r0 = (int)((int64_t)blockIdx.x * 128);
        mbar_expect_tx(bar_h0, TILE_BYTES);
        tma::load_tile(x_s + (uint32_t)pH0 * TILE_BYTES, &tm_h0, r0, bar_h0);
        mbar_expect_tx(bar_g, TILE_BYTES);
        tma::load_tile(x_s + (uint32_t)pG * TILE_BYTES, &tm_gout, r0, bar_g);
        if (KEPT) {
          mbar_expect_tx(bar_h2, TILE_BYTES);
          tma::load_tile(x_s + (uint32_t)pH2 * TILE_BYTES, &tm_h2, r0, bar_h2);
          mbar_expect_tx(bar_h1, TILE_BYTES);
          tma::load_tile(x_s + (uint32_t)pH1 * TILE_BYTES, &tm_h1, r0, bar_h1);
        }
      }
      __syncwarp();
    }
    if (KEPT) {
      prefetch(L + 1);   // W_out: output GEMM and the first backward phase; W_2 follows in the other slot
      prefetch(L);
    } else {
      prefetch(1);
      prefetch(2);   // L >= 1: matrix 2 exists (W_2 or W_out)
    }
  }
  bool h_posted = true;   // KEPT: this tile's H_1 / H_0 loads are already in flight (first tile: posted above)
  uint32_t ph_h1 = 0, ph_h2 = 0;

  uint32_t ph_mma = 0, ph_h0 = 0, ph_g = 0, ph_r = 0, ph_dw = 0;
  bool first_tile = true;
  float db[UMMA_MAX_L_BWD + 1][2];
#pragma unroll
  for (int i = 0; i <= UMMA_MAX_L_BWD; ++i) db[i][0] = db[i][1] = 0.f;
  float dbet0 = 0.f, dbet1 = 0.f, db00 = 0.f, db01 = 0.f, dgam = 0.f;

  // receiver id of this thread's row, fetched one tile ahead (its latency never meets a dependent load)
  int dst_next = -1;
  if (a.g_agg && (int64_t)blockIdx.x * 128 + row < a.rows) dst_next = __ldg(a.idx1 + (int64_t)blockIdx.x * 128 + row);

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    const int nrows = (int)((a.rows - row0) < 128 ? (a.rows - row0) : 128);
    const bool valid = row < nrows;
    const bool has_next = tile + gridDim.x < tiles;
    const int next_r0 = (int)((tile + gridDim.x) * 128);
    // this thread's piece of the receiver-gradient row (fp32): added to the incoming gradient and to the residual
    const float* gag = nullptr;
    if (dst_next >= 0) {
      gag = a.g_agg + (size_t)dst_next * 128 + ch * 32;
      prefetch_l1(gag);   // 128 bytes = one line; read three GEMMs later (LayerNorm backward) and again by the last epilogue
    }
    dst_next = -1;
    if (a.g_agg && has_next && (int64_t)next_r0 + row < a.rows) dst_next = __ldg(a.idx1 + (int64_t)next_r0 + row);

    // Tiles released by backward phase `done` (its MMAs have completed and every warp has passed the phase's closing
    // barrier, so neither the tensor core nor a column-sum reader still touches them) receive what the TMA engine is
    // asked for next.  Called by the elected lane of warp 0 at the start of the following phase.
    auto release = [&](int done) {
      if (done == L + 1 && resid) {      // the incoming gradient rows again, for the residual add of the last epilogue
        mbar_expect_tx(bar_r, TILE_BYTES);
        tma::load_tile(x_s + (uint32_t)pG * TILE_BYTES, &tm_gout, (int)row0, bar_r);
      }
      if (KEPT) {
        // the next tile starts with its incoming gradient and H_2 resident: they go to the tiles phases 2 and 1 free
        if (done == L && has_next) {
          mbar_expect_tx(bar_g, TILE_BYTES);
          tma::load_tile(x_s + (uint32_t)pH2 * TILE_BYTES, &tm_gout, next_r0, bar_g);
        }
        if (done == 1 && has_next) {
          mbar_expect_tx(bar_h2, TILE_BYTES);
          tma::load_tile(x_s + (uint32_t)pH1 * TILE_BYTES, &tm_h2, next_r0, bar_h2);
        }
        return;
      }
      if (done == L && has_next) {       // next tile's h_0 rows: H_L's tile (L == 2) or the spare (L == 1)
        mbar_expect_tx(bar_h0, TILE_BYTES);
        tma::load_tile(x_s + (uint32_t)pH2 * TILE_BYTES, &tm_h0, next_r0, bar_h0);
      }
      if (done == 1 && has_next) {       // next tile's incoming gradient rows: H_1's tile
        mbar_expect_tx(bar_g, TILE_BYTES);
        tma::load_tile(x_s + (uint32_t)pH1 * TILE_BYTES, &tm_gout, next_r0, bar_g);
      }
    };

    if (KEPT) {
      // this tile's H_1 and H_0 rows go to the two tiles the previous tile's output stores are leaving (needed two and
      // three phases from now); its gradient rows and H_2 were fetched while the previous tile computed
      if (w0) {
        if (elect_one() && !h_posted) {
          tma::store_wait_read();
          mbar_expect_tx(bar_h1, TILE_BYTES);
          tma::load_tile(x_s + (uint32_t)pH1 * TILE_BYTES, &tm_h1, (int)row0, bar_h1);
          mbar_expect_tx(bar_h0, TILE_BYTES);
          tma::load_tile(x_s + (uint32_t)pH0 * TILE_BYTES, &tm_h0, (int)row0, bar_h0);
        }
        __syncwarp();
      }
      h_posted = false;
      mbar_wait(bar_h2, ph_h2);
      ph_h2 ^= 1;
      mbar_wait(bar_g, ph_g);
      ph_g ^= 1;
    } else {
      mbar_wait(bar_h0, ph_h0);
      ph_h0 ^= 1;
    }
    if (w0 && has_next) {   // pull the next tile's rows into L2 now; the loads above then hit L2
      if (elect_one()) {
        tma::prefetch_tile_l2(&tm_h0, next_r0);
        tma::prefetch_tile_l2(&tm_gout, next_r0);
        if (KEPT) {
          tma::prefetch_tile_l2(&tm_h1, next_r0);
          tma::prefetch_tile_l2(&tm_h2, next_r0);
        }
      }
      __syncwarp();
    }

    // ---- forward recompute of layers 1..L (hidden) ----
    for (int m = 1; m <= (KEPT ? 0 : L); ++m) {
      if (w0) {
        uint32_t wa = acquire(m);
        fence_after_sync();
        if (elect_one()) {
          if (m == 1) tma::store_wait_read();   // the previous tile's output tiles have left (they become H_1, H_2)
          issue_gemm(tmem_base, x_s + (uint32_t)tile_of_h(m - 1) * TILE_BYTES, false, wa, false, false);
          mma_commit(bar_mma);
        }
        __syncwarp();
        prefetch(m + 1);   // other slot: last read by GEMM m-1, already complete
      }
      if (m == 1) {
        // under the first MMA: d(beta) part 1 = column sums of the incoming gradient rows
        mbar_wait(bar_g, ph_g);
        ph_g ^= 1;
        colsum_mma(X + (size_t)pG * TILE_BYTES, wid, lane, dbet0, dbet1);
      }
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      fence_after_sync();
      hidden_epilogue_chunk(tlane - (uint32_t)(ch * 32), ch, nullptr, nullptr, vec + (m - 1) * 128, act,
                            X + (size_t)tile_of_h(m) * TILE_BYTES, row);
      fence_before_sync();
      fence_async_smem();
      __syncthreads();
    }
    // ---- output layer + LayerNorm backward; the G tile holds g_out and ends up holding dL/dy ----
    {
      uint8_t* G = X + (size_t)pG * TILE_BYTES;
      if (w0) {
        uint32_t wa = acquire(L + 1);
        fence_after_sync();
        if (elect_one()) {
          issue_gemm(tmem_base, x_s + (uint32_t)tile_of_h(L) * TILE_BYTES, false, wa, false, false);
          mma_commit(bar_mma);
        }
        __syncwarp();
        if (KEPT) prefetch(L);   // W_2 into the other slot (held W_main, consumed by the previous tile's last GEMM)
      }
      if (KEPT) colsum_mma(G, wid, lane, dbet0, dbet1);   // d(beta) part 1 = column sums of the incoming gradient rows
      // under the MMA: this thread's 32 incoming gradient values (4 x 16 B of the G tile) + receiver gradient piece
      uint32_t gp[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 g4 = *reinterpret_cast<const uint4*>(G + tile_chunk_off(row, ch * 4 + j));
        gp[4 * j] = g4.x; gp[4 * j + 1] = g4.y; gp[4 * j + 2] = g4.z; gp[4 * j + 3] = g4.w;
      }
      if (gag) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(gag) + j);
          gp[2 * j] = pack_bf16(bf16_lo(gp[2 * j]) + t4.x, bf16_hi(gp[2 * j]) + t4.y);
          gp[2 * j + 1] = pack_bf16(bf16_lo(gp[2 * j + 1]) + t4.z, bf16_hi(gp[2 * j + 1]) + t4.w);
        }
      }
      // ... and, still under the MMA, the part of the LayerNorm backward that does not need y: A = sum g*gamma
      const float* gam = vec + (L + 1) * 128 + ch * 32;
      float aa = 0.f, ab = 0.f;
      if (a.use_ln) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          aa = fmaf(bf16_lo(gp[j]), gam[2 * j], aa);
          ab = fmaf(bf16_hi(gp[j]), gam[2 * j + 1], ab);
        }
      }
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      fence_after_sync();
      float v[32];
      tmem_ld32(tlane, v);
      fence_before_sync();
      const float* bo = vec + L * 128 + ch * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b4 = *reinterpret_cast<const float4*>(bo + 4 * j);
        v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
      }
      if (a.use_ln) {
        // one pass, three more row sums: S1 = sum y, S2 = sum y^2, B = sum g*gamma*y
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f, ba = 0.f, bb = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float ylo = v[2 * j], yhi = v[2 * j + 1];
          s1a += ylo; s1b += yhi;
          s2a = fmaf(ylo, ylo, s2a); s2b = fmaf(yhi, yhi, s2b);
          ba = fmaf(bf16_lo(gp[j]) * gam[2 * j], ylo, ba);
          bb = fmaf(bf16_hi(gp[j]) * gam[2 * j + 1], yhi, bb);
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
        // dL/dy = rstd * (g*gamma - mean(g*gamma) - yhat * mean(g*gamma*yhat)),  yhat = y*rstd - mean*rstd
        const float c0 = -mean * rstd;
        const float c1 = -rstd * A * (1.f / 128.f);
        const float c2 = -rstd * rstd * (B - mean * A) * (1.f / 128.f);
        float z[32];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint32_t op[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = 4 * jj + k;
            const float glo = bf16_lo(gp[j]), ghi = bf16_hi(gp[j]);
            const float hlo = fmaf(v[2 * j], rstd, c0), hhi = fmaf(v[2 * j + 1], rstd, c0);
            z[2 * j] = glo * hlo;           // g * yhat: d(gamma) terms
            z[2 * j + 1] = ghi * hhi;
            op[k] = pack_bf16(fmaf(hlo, c2, fmaf(glo * gam[2 * j], rstd, c1)),
                              fmaf(hhi, c2, fmaf(ghi * gam[2 * j + 1], rstd, c1)));
          }
          *reinterpret_cast<uint4*>(G + tile_chunk_off(row, ch * 4 + jj)) = make_uint4(op[0], op[1], op[2], op[3]);
        }
        // d(gamma): column sums of z over the warp's 32 rows by a shuffle transpose-reduce (lane l ends with the
        // sum of column 32*ch + l); fixed order -> deterministic.  (Deferring it under the next phase's MMAs was
        // measured slower: 32 more live registers across the phase boundary.)
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
      } else if (gag) {
        // no LayerNorm: dL/dy = g (with the receiver part added)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(G + tile_chunk_off(row, ch * 4 + j)) = make_uint4(gp[4 * j], gp[4 * j + 1], gp[4 * j + 2], gp[4 * j + 3]);
      }
      fence_async_smem();
      __syncthreads();
    }
    // ---- backward through the Linear layers L+1 .. 1 ----
    for (int m = L + 1; m >= 1; --m) {
      const int tg = (m == L + 1) ? pG : tile_of_h(m);       // tile holding G_m
      const int th = tile_of_h(m - 1);                        // tile holding H_{m-1}, receives G_{m-1}
      if (KEPT && m <= L) {   // H_{m-1} of this tile has landed (H_L was awaited at the top)
        if (m == 2) {
          mbar_wait(bar_h1, ph_h1);
          ph_h1 ^= 1;
        } else {
          mbar_wait(bar_h0, ph_h0);
          ph_h0 ^= 1;
        }
      }
      if (w0) {
        uint32_t wa = acquire(m);
        fence_after_sync();
        if (elect_one()) {
          release(m + 1);
          const uint32_t g_addr = x_s + (uint32_t)tg * TILE_BYTES;
          // the data gradient first, with its own completion barrier: the epilogue starts on it while the weight
          // gradient runs, and only its in-place stores over H_{m-1} (the dW operand) wait for the second barrier
          issue_gemm(tmem_base, g_addr, false, wa, true, false);                                                           // G W_m
          mma_commit(bar_mma);
          issue_gemm(tmem_base + (uint32_t)(128 * m), g_addr, true, x_s + (uint32_t)th * TILE_BYTES, true, !first_tile);   // dW_m += G^T H
          mma_commit(bar_dw);
        }
        __syncwarp();
        prefetch(m - 1);
      }
      // under the MMAs: bias gradient of Linear m = column sums of G_m
      {
        float t0 = 0.f, t1 = 0.f;
        colsum_mma(X + (size_t)tg * TILE_BYTES, wid, lane, t0, t1);
#pragma unroll
        for (int i = 0; i <= UMMA_MAX_L_BWD; ++i) {   // constant indices: db stays in registers
          db[i][0] += (i == m - 1) ? t0 : 0.f;
          db[i][1] += (i == m - 1) ? t1 : 0.f;
        }
      }
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      fence_after_sync();
      mask_epilogue_chunk<RELU>(tlane, X + (size_t)th * TILE_BYTES, row, ch, act, bar_dw, ph_dw);
      ph_dw ^= 1;
      fence_before_sync();
      fence_async_smem();
      __syncthreads();
    }
    // ---- m = 0: g_main = G_0 W_main (* scale) (+ residual gradient); G_0 leaves as g_h0 ----
    {
      uint8_t* G0 = X + (size_t)pH0 * TILE_BYTES;
      if (w0) {
        uint32_t wa = acquire(0);
        fence_after_sync();
        if (elect_one()) {
          release(1);
          issue_gemm(tmem_base, x_s + (uint32_t)pH0 * TILE_BYTES, false, wa, true, false);
          mma_commit(bar_mma);
          tma::store_tile(&tm_gh0, x_s + (uint32_t)pH0 * TILE_BYTES, (int)row0);
          tma::store_commit();
        }
        __syncwarp();
      }
      colsum_mma(G0, wid, lane, db00, db01);   // gradient of the first Linear's bias
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      fence_after_sync();
      if (w0 && has_next) {
        if (KEPT) prefetch(L + 1);   // slot 1 held W_1 (consumed by phase 1): the next tile's W_out
        else if (!F32T) prefetch(2);     // slot 0 held W_main, just consumed: the next tile's second matrix
        // (fp32 g_main: slot 0 first stages two of the four fp32 output panels; W_2 follows at the next tile's first GEMM)
      }
      uint8_t* O = X + (size_t)pG * TILE_BYTES;
      if (resid) {
        mbar_wait(bar_r, ph_r);
        ph_r ^= 1;
      }
      float v[32];
      tmem_ld32(tlane, v);
      fence_before_sync();
      if (a.main_scale && valid) {
        const float sc = a.main_scale[row0 + row];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sc;
      }
      if (resid) {
        add_tile_chunk(v, O, row, ch);
        if (gag) {   // L2 hits: the same rows were read by this tile's LayerNorm backward
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(gag) + j);
            v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
          }
        }
      }
      if (F32T) {
        // fp32 rows (the node block's g_agg) leave through TMA as well: 128 x 128 fp32 = four 32-column panels; chunks
        // 0, 1 stage theirs in the O tile, chunks 2, 3 in weight slot 0 (W_main, consumed by the GEMM above).  Row-per-
        // lane 16-byte global stores would be 32 half-filled sectors per warp instruction (4096 requests per tile).
        uint8_t* S = (ch < 2 ? O : Wslot) + (size_t)(ch & 1) * PANEL_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(S + row * 128 + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        wst &= ~7u;                  // slot 0 no longer holds W_main
        fence_async_smem();
        __syncthreads();
        if (w0) {
          if (elect_one()) {
            const uint32_t o_s = x_s + (uint32_t)pG * TILE_BYTES;
            tma::store_panel_f32(&tm_gmain, o_s, 0, (int)row0);
            tma::store_panel_f32(&tm_gmain, o_s + PANEL_BYTES, 32, (int)row0);
            tma::store_panel_f32(&tm_gmain, w_s, 64, (int)row0);
            tma::store_panel_f32(&tm_gmain, w_s + PANEL_BYTES, 96, (int)row0);
            tma::store_commit();   // the next tile's first GEMM waits for these reads before either region is reused
          }
          __syncwarp();
        }
      } else if (a.main_f32) {
        if (valid) {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.g_main) + (row0 + row) * 128 + ch * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncthreads();   // every reader of this tile's shared memory is done before the roles rotate
      } else {
        store_row32(O, row, ch, v);   // in place over the reloaded gradient rows (same thread, same bytes)
        fence_async_smem();
        __syncthreads();
        if (w0) {
          if (elect_one()) {
            tma::store_tile(&tm_gmain, x_s + (uint32_t)pG * TILE_BYTES, (int)row0);
            tma::store_commit();
          }
          __syncwarp();
        }
      }
    }
    // ---- rotate the tile roles: the next tile's inputs are (being) loaded into the tiles freed above ----
    {
      const int oH0 = pH0, oH1 = pH1, oH2 = pH2, oG = pG;
      if (KEPT) {
        pG = oH2;    // next incoming gradient rows (posted after phase 2)
        pH2 = oH1;   // next H_2 rows (posted after phase 1)
        pH1 = oG;    // next H_1 rows: posted at the top of the next tile, once the g_main store has read the tile
        pH0 = oH0;   // next h_0 rows: likewise after the g_h0 store
        first_tile = false;
        continue;
      }
      pH0 = oH2;   // next h_0 rows
      pG = oH1;    // next incoming gradient rows
      pH1 = oG;    // written by the next tile's first epilogue, after the g_main store has read it
      pH2 = oH0;   // L == 2: H_2 of the next tile; L == 1: the spare (after the g_h0 store has read it)
    }
    first_tile = false;
  }

  // ---- flush per-CTA partial gradients: dW_m from TMEM, vectors from registers ----
  if (w0) {
    if (elect_one()) tma::store_wait_all();
    __syncwarp();
  }
  __syncthreads();
  fence_after_sync();
  const PackedLayout pl{L};
  float* part_out = a.w_part + (size_t)blockIdx.x * pl.total();
  const bool worked = (int64_t)blockIdx.x < tiles;
  for (int m = 1; m <= L + 1; ++m) {
    float* dst = part_out + (m == L + 1 ? pl.w_out() : pl.w_hidden(m - 1)) + (size_t)row * 128 + ch * 32;
    float v[32];
    if (worked) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(128 * m + ch * 32), v);
    else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  fence_before_sync();
  // d(beta) part 2: sum over receivers of (in-degree) x g_agg row, this CTA's slice of the nodes, thread = (column, quarter)
  {
    float acc = 0.f;
    if (a.g_agg && a.rowptr) {
      const int64_t per = (a.n_nodes + gridDim.x - 1) / gridDim.x;
      const int64_t n0 = (int64_t)blockIdx.x * per;
      const int64_t n1 = (n0 + per) < a.n_nodes ? (n0 + per) : a.n_nodes;
      const int col = tid & 127;
      int64_t n = n0 + (tid >> 7);
      for (; n + 28 < n1; n += 32) {   // eight rows per trip, all loads issued before the first use
        float g[8], c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          g[u] = __ldg(a.g_agg + (size_t)(n + 4 * u) * 128 + col);
          c[u] = (float)(__ldg(a.rowptr + n + 4 * u + 1) - __ldg(a.rowptr + n + 4 * u));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = fmaf(c[u], g[u], acc);
      }
      for (; n < n1; n += 4) {
        const float cnt = (float)(a.rowptr[n + 1] - a.rowptr[n]);
        acc = fmaf(cnt, __ldg(a.g_agg + (size_t)n * 128 + col), acc);
      }
    }
    red[tid] = acc;
  }
  __syncthreads();
  // vectors: lanes 0..3 of warp w own columns 8w + 2*lane + {0,1} (colsum_mma)
  if (lane < 4) {
    const int c = 8 * wid + 2 * lane;
#pragma unroll
    for (int l = 0; l <= UMMA_MAX_L_BWD; ++l) {   // db[l] = bias gradient of Linear l+1; Linear L+1 is the output layer
      if (l < L) {
        part_out[pl.b_hidden(l) + c] = db[l][0];
        part_out[pl.b_hidden(l) + c + 1] = db[l][1];
      } else if (l == L) {
        part_out[pl.b_out() + c] = db[l][0];
        part_out[pl.b_out() + c + 1] = db[l][1];
      }
    }
    part_out[pl.beta() + c] = dbet0 + ((red[c] + red[128 + c]) + (red[256 + c] + red[384 + c]));
    part_out[pl.beta() + c + 1] = dbet1 + ((red[c + 1] + red[128 + c + 1]) + (red[256 + c + 1] + red[384 + c + 1]));
    part_out[pl.bias0() + c] = db00;
    part_out[pl.bias0() + c + 1] = db01;
  }
  __syncthreads();
  red[q * 128 + ch * 32 + lane] = dgam;   // partial over the rows of lane quarter q
  __syncthreads();
  if (tid < 128) part_out[pl.gamma() + tid] = (red[tid] + red[128 + tid]) + (red[256 + tid] + red[384 + tid]);
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tmem_base);
}

// ---- host side ------------------------------------------------------------------------------------
static size_t bwd2_smem() {
  return 1024 + (size_t)6 * TILE_BYTES + (size_t)(UMMA_MAX_L_BWD + 3) * 512 + 8192 + 9 * 8 + 16;
}

bool umma_bwd2_applicable(const aero_block_desc* d) {
  if (d->h0 == nullptr || d->L < 1 || d->L > 2 || d->rows <= 0) return false;
  if (d->g_agg && !d->rowptr) return false;   // d(beta) needs the receiver degrees
  const uintptr_t al = (uintptr_t)d->h0 | (uintptr_t)d->g_out | (uintptr_t)d->g_h0 | (uintptr_t)d->g_main;
  return (al & 15) == 0;
}

int umma_block_bwd2(const aero_block_desc* d, UmmaArgs a, int grid, cudaStream_t st) {
  CUtensorMap tm_h0, tm_gout, tm_gh0, tm_gmain, tm_h1, tm_h2;
  const bool kept = d->L == 2 && d->h_hidden[0] && d->h_hidden[1];
  if (tma::make_rows_map(kept ? d->h_hidden[0] : d->h0, d->rows, &tm_h1) ||
      tma::make_rows_map(kept ? d->h_hidden[1] : d->h0, d->rows, &tm_h2) ||
      tma::make_rows_map(d->h0, d->rows, &tm_h0) || tma::make_rows_map(d->g_out, d->rows, &tm_gout) ||
      tma::make_rows_map(d->g_h0, d->rows, &tm_gh0) ||
      (d->main_f32 ? tma::make_rows_map_f32(d->g_main, d->rows, &tm_gmain) : tma::make_rows_map(d->g_main, d->rows, &tm_gmain))) {
    set_error("umma_block_bwd2: cuTensorMapEncodeTiled failed");
    return AERO_ECUDA;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  AERO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
#define AERO_B2_ATTR(R, K, F) \
  AERO_CUDA(cudaFuncSetAttribute(umma_block_bwd2_kernel<R, K, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd2_smem()))
    AERO_B2_ATTR(true, false, false); AERO_B2_ATTR(false, false, false);
    AERO_B2_ATTR(true, true, false);  AERO_B2_ATTR(false, true, false);
    AERO_B2_ATTR(true, false, true);  AERO_B2_ATTR(false, false, true);
#undef AERO_B2_ATTR
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const bool relu = d->act == AERO_ACT_RELU;
  const bool f32t = d->main_f32 && !kept;
#define AERO_B2_LAUNCH(R, K, F) \
  umma_block_bwd2_kernel<R, K, F><<<grid, B2_THREADS, bwd2_smem(), st>>>(a, tm_h0, tm_gout, tm_gh0, tm_gmain, tm_h1, tm_h2)
  if (kept) {
    if (relu) AERO_B2_LAUNCH(true, true, false);
    else AERO_B2_LAUNCH(false, true, false);
  } else if (f32t) {
    if (relu) AERO_B2_LAUNCH(true, false, true);
    else AERO_B2_LAUNCH(false, false, true);
  } else {
    if (relu) AERO_B2_LAUNCH(true, false, false);
    else AERO_B2_LAUNCH(false, false, false);
  }
#undef AERO_B2_LAUNCH
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

}  // namespace aero
