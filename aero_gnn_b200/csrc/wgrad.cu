// wgrad.cu -- weight-gradient reductions over rows on tcgen05, warp-specialised and fed by TMA.
//
//   dW[j] (128 x 128, fp32) = A[:, 128 j : 128 j + 128]^T  B        j = 0 .. a-1   (a = 1 or 2)
//
// A [rows, 128 a] and B [rows, 128] are bf16 row matrices in HBM (row strides lda / 128).  This is the first Linear's
// weight gradient of a block (mgnLayer.py:97-103 run backwards: dW_e = g_h0^T e over the E edge rows; dW_na = g_h0n^T
// agg), and the projection weight gradients [g_P_s | g_P_d]^T x, g_h0n^T x -- plain reductions over the rows with no
// epilogue per tile, i.e. HBM-bound streaming.  One persistent CTA per SM:
//   warp 0     TMA producer: per 128-row tile the A panel(s) and the B tile land in a ring of shared-memory stages
//              (SWIZZLE_128B tensor maps = the UMMA tile format), signalled on the stage's `full` mbarrier;
//   warp 1     MMA issuer: 8 tcgen05.mma (both operands MN-major views of the row tiles, K = the 128 rows) per A panel
//              into the panel's 128 TMEM columns, accumulating over every tile of the CTA; tcgen05.commit releases the
//              stage (`empty` mbarrier);
//   warps 2..9 (SEG only) receiver sums of A panel 0 taken from the same shared-memory tile: the rows are in
//              receiver-CSR order, so out_seg[n] = sum of the A rows of receiver n (the gradient of the receiver-side
//              pre-projection P_d) comes from the bytes the weight gradient needs anyway -- one pass over g_h0 instead of
//              two.  Runs that straddle tiles leave fp32 partial rows for gpd_fixup_kernel.
// At the end the accumulators are written as per-CTA partials and summed in CTA order (deterministic).
#include "umma_block.cuh"
#include "tma.cuh"

namespace aero {

constexpr int WG_THREADS = 320;   // 10 warps: producer, MMA issuer, 8 consumers
constexpr int WG_CONS = 8;
constexpr uint32_t WG_IDS_BOX = 136;      // receiver ids of rows [row0 - 4, row0 + 132): the tile and its two neighbours
constexpr uint32_t WG_IDS_BYTES = 1024;  // keeps the stages 1024-byte aligned

struct WgradArgs {
  int64_t rows, n_nodes, seg_ld;
  int a;                         // A panels of 128 columns (1 or 2)
  int stages;
  const int32_t* dst;            // SEG: receiver of each row
  __nv_bfloat16* seg_out;        // SEG: [n_nodes, seg_ld >= 128] bf16 receiver sums of A panel 0
  float* seg_part;               // SEG: [tiles][2][128] fp32 partial rows of runs that straddle tiles
  float* part;                   // [grid][a][128][128] fp32 per-CTA partial results
};

template <bool SEG>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(WgradArgs g, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
             const __grid_constant__ CUtensorMap tm_ids) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int a = g.a, S = g.stages;
  const uint32_t tile_bytes = (uint32_t)(a + 1) * TILE_BYTES;                 // A panel(s) + B
  const uint32_t stage_bytes = tile_bytes + (SEG ? WG_IDS_BYTES : 0u);        // + the tile's receiver ids
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);   // full[S], empty[S]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2 * S);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&mbar[s]), 1);
      mbar_init(smem_u32(&mbar[S + s]), 1 + (SEG ? WG_CONS : 0));
    }
    fence_mbar_init();
  }
  if (wid == 0) tmem_alloc<256>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t s0 = smem_u32(smem);
  const int64_t tiles = (g.rows + 127) / 128;
  const int64_t my_tiles = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (wid == 0) {
    // ---- TMA producer ----
    if (elect_one()) {
      tma::prefetch_map(&tm_a);
      tma::prefetch_map(&tm_b);
      if (SEG) tma::prefetch_map(&tm_ids);
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int s = (int)(i % S);
        const uint32_t use = (uint32_t)(i / S);
        if (use > 0) mbar_wait(smem_u32(&mbar[S + s]), (use - 1) & 1);   // the stage's previous contents are consumed
        const int row0 = (int)((blockIdx.x + i * gridDim.x) * 128);
        const uint32_t base = s0 + (uint32_t)s * stage_bytes, full = smem_u32(&mbar[s]);
        mbar_expect_tx(full, tile_bytes + (SEG ? WG_IDS_BOX * 4u : 0u));
        for (int p = 0; p < 2 * a; ++p) tma::load_panel(base + (uint32_t)p * PANEL_BYTES, &tm_a, p, row0, full);
        tma::load_tile(base + (uint32_t)a * TILE_BYTES, &tm_b, row0, full);
        if (SEG) tma::load_ids(base + tile_bytes, &tm_ids, row0 - 4, full);   // out-of-range ids arrive as zeros
      }
    }
    __syncwarp();
  } else if (wid == 1) {
    // ---- MMA issuer ----
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int s = (int)(i % S);
      mbar_wait(smem_u32(&mbar[s]), (uint32_t)(i / S) & 1);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t base = s0 + (uint32_t)s * stage_bytes;
        for (int j = 0; j < a; ++j)
          issue_gemm(tmem_base + (uint32_t)(128 * j), base + (uint32_t)j * TILE_BYTES, true, base + (uint32_t)a * TILE_BYTES,
                     true, i > 0);
        mma_commit(smem_u32(&mbar[S + s]));
      }
      __syncwarp();
    }
  } else if (SEG) {
    // ---- receiver sums of A panel 0 (consumer warp c owns the runs whose head row lies in [16c, 16c + 16)) ----
    const int c = wid - 2;
    // the tile's receiver ids (and those of the rows just outside it) arrive in the stage with the tile, by TMA:
    // a consumer never waits on HBM latency under a saturated bus
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int s = (int)(i % S);
      const int64_t tile = blockIdx.x + i * gridDim.x, row0 = tile * 128;
      const int nrows = (int)((g.rows - row0) < 128 ? (g.rows - row0) : 128);
      mbar_wait(smem_u32(&mbar[s]), (uint32_t)(i / S) & 1);
      const int* ids = reinterpret_cast<const int*>(smem + (size_t)s * stage_bytes + tile_bytes);   // ids[4 + r] = dst[row0 + r]
      int dq[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int r = q4 * 32 + lane;
        dq[q4] = r < nrows ? ids[4 + r] : -1;
      }
      const int before = (row0 > 0) ? ids[3] : -2;
      const int after = (row0 + nrows < g.rows) ? ids[4 + nrows] : -2;
      uint32_t hm[4];
      int carry = -3;   // receiver of the previous row quarter's last row
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        int prev = __shfl_up_sync(0xffffffffu, dq[q4], 1);
        if (lane == 0) prev = carry;
        const int r = q4 * 32 + lane;
        hm[q4] = __ballot_sync(0xffffffffu, r < nrows && (r == 0 || dq[q4] != prev));
        carry = __shfl_sync(0xffffffffu, dq[q4], 31);
      }
      const int first_dst = __shfl_sync(0xffffffffu, dq[0], 0);
      const int last_q = (nrows - 1) >> 5, last_l = (nrows - 1) & 31;
      const int last_dst = __shfl_sync(0xffffffffu, last_q == 0 ? dq[0] : (last_q == 1 ? dq[1] : (last_q == 2 ? dq[2] : dq[3])), last_l);
      const bool head_open = before == first_dst, tail_open = after == last_dst;
      auto next_head = [&](int r) -> int {
        int w = r >> 5;
        uint32_t m = (r & 31) == 31 ? 0u : ((w == 0 ? hm[0] : (w == 1 ? hm[1] : (w == 2 ? hm[2] : hm[3]))) & (0xffffffffu << ((r & 31) + 1)));
        while (m == 0u && ++w < 4) m = (w == 1 ? hm[1] : (w == 2 ? hm[2] : hm[3]));
        return m ? w * 32 + __ffs(m) - 1 : nrows;
      };
      const uint32_t word = (c >> 1) == 0 ? hm[0] : ((c >> 1) == 1 ? hm[1] : ((c >> 1) == 2 ? hm[2] : hm[3]));
      const uint32_t mine = (word >> ((c & 1) * 16)) & 0xffffu;
      const uint8_t* base = smem + (size_t)s * stage_bytes + (lane >> 4) * PANEL_BYTES + (lane & 1) * 8;
      const int chk = (lane >> 1) & 7;
      for (uint32_t bits = mine; bits; bits &= bits - 1) {
        const int rs = 16 * c + __ffs(bits) - 1, re = next_head(rs);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = rs; t < re; t += 4) {   // four rows in flight; rows past the run add zeros (same sums, row order)
          uint2 u[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            u[k] = (t + k < re) ? *reinterpret_cast<const uint2*>(base + (t + k) * 128 + ((chk ^ ((t + k) & 7)) << 4))
                                : make_uint2(0u, 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc.x += bf16_lo(u[k].x); acc.y += bf16_hi(u[k].x); acc.z += bf16_lo(u[k].y); acc.w += bf16_hi(u[k].y);
          }
        }
        const int q4 = rs >> 5;
        const int n = __shfl_sync(0xffffffffu, q4 == 0 ? dq[0] : (q4 == 1 ? dq[1] : (q4 == 2 ? dq[2] : dq[3])), rs & 31);
        if ((rs == 0 && head_open) || (re == nrows && tail_open)) {
          *reinterpret_cast<float4*>(g.seg_part + ((size_t)tile * 2 + (rs == 0 ? 0 : 1)) * 128 + lane * 4) = acc;
        } else {
          uint2 o;
          o.x = pack_bf16(acc.x, acc.y);
          o.y = pack_bf16(acc.z, acc.w);
          *reinterpret_cast<uint2*>(g.seg_out + (size_t)n * g.seg_ld + lane * 4) = o;
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&mbar[S + s])) : "memory");
    }
  }
  // ---- accumulators -> per-CTA partials (warps 2..5 read the four TMEM lane quarters) ----
  if (wid == 1 && my_tiles > 0) {   // every MMA has completed when the last stage's `empty` barrier flips
    const int64_t i = my_tiles - 1;
    mbar_wait(smem_u32(&mbar[S + (int)(i % S)]), (uint32_t)(i / S) & 1);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (wid >= 2 && wid < 6) {
    const int q = wid - 2;   // TMEM lane quarter must equal warp id % 4: warps 2..5 -> quarters 2,3,0,1
    const int lq = wid & 3;
    (void)q;
    const int row = lq * 32 + lane;
    float* out = g.part + (size_t)blockIdx.x * a * 16384;
    for (int j = 0; j < a; ++j) {
      for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        if (my_tiles > 0) tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(128 * j + 32 * cb), v);
        else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0.f;
        }
        float4* dstp = reinterpret_cast<float4*>(out + (size_t)j * 16384 + (size_t)row * 128 + cb * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) dstp[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (wid == 0) tmem_dealloc<256>(tmem_base);
}

static int wg_grid(int64_t rows) {
  int64_t tiles = cdiv(rows > 0 ? rows : 1, 128);
  return (int)(tiles < sm_count() ? tiles : sm_count());
}
static int wg_stages(int a) { return a == 1 ? 3 : 2; }
static size_t wg_smem(int a);
static size_t wg_smem_max() { return wg_smem(1) > wg_smem(2) ? wg_smem(1) : wg_smem(2); }
static size_t wg_smem(int a) { return 1024 + (size_t)wg_stages(a) * ((a + 1) * TILE_BYTES + WG_IDS_BYTES) + 2 * 4 * 8 + 64; }

}  // namespace aero

using namespace aero;

extern "C" size_t aero_wgrad_workspace_bytes(int64_t rows, int a_panels, int with_seg) {
  size_t b = align_up((size_t)wg_grid(rows) * a_panels * 16384 * sizeof(float), 256);
  if (with_seg) b += align_up((size_t)cdiv(rows > 0 ? rows : 1, 128) * 2 * 128 * sizeof(float), 256);
  return b;
}

extern "C" int aero_wgrad(const void* A, int64_t lda, int a_panels, const void* B, int64_t rows, float* dW,
                          const int32_t* dst, const int32_t* rowptr, int64_t n_nodes, void* seg_out, int64_t seg_ld,
                          void* workspace, size_t workspace_bytes, void* stream) {
  g_launch_count = 0;
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG((a_panels == 1 || a_panels == 2) && rows >= 0 && dW && lda >= 128 * a_panels && (lda % 8) == 0,
                 "aero_wgrad: bad arguments");
  const bool seg = seg_out != nullptr;
  AERO_CHECK_ARG(!seg || ((dst || rows == 0) && rowptr && n_nodes >= 0 && seg_ld >= 128 && (seg_ld % 4) == 0), "aero_wgrad: receiver sums need dst, rowptr, n_nodes");
  AERO_CHECK_ARG(workspace && workspace_bytes >= aero_wgrad_workspace_bytes(rows, a_panels, seg), "aero_wgrad: workspace");
  if (rows == 0) {
    AERO_CUDA(cudaMemsetAsync(dW, 0, (size_t)a_panels * 16384 * sizeof(float), st));
    if (seg) return launch_gpd_fixup(nullptr, rowptr, dst, reinterpret_cast<__nv_bfloat16*>(seg_out), seg_ld, 0, n_nodes, st);
    return AERO_OK;
  }
  AERO_CHECK_ARG(A && B, "aero_wgrad: null operand");
  CUtensorMap tm_a, tm_b, tm_ids;
  if (tma::make_rows_map_ld(A, rows, 128 * a_panels, lda, &tm_a) || tma::make_rows_map(B, rows, &tm_b) ||
      (seg ? tma::make_ids_map(dst, rows, WG_IDS_BOX, &tm_ids) : (tm_ids = tm_b, 0))) {
    set_error("aero_wgrad: cuTensorMapEncodeTiled failed (operands must be 16-byte aligned bf16 row matrices)");
    return AERO_ECUDA;
  }
  const int grid = wg_grid(rows);
  WgradArgs g;
  g.rows = rows; g.n_nodes = n_nodes; g.seg_ld = seg_ld; g.a = a_panels; g.stages = wg_stages(a_panels);
  g.dst = dst; g.seg_out = reinterpret_cast<__nv_bfloat16*>(seg_out);
  g.part = reinterpret_cast<float*>(workspace);
  g.seg_part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) +
                                        align_up((size_t)grid * a_panels * 16384 * sizeof(float), 256));
  static bool attr_set[64] = {false};
  int dev = 0;
  AERO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    AERO_CUDA(cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem_max()));
    AERO_CUDA(cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem_max()));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  if (seg) wgrad_kernel<true><<<grid, WG_THREADS, wg_smem(a_panels), st>>>(g, tm_a, tm_b, tm_ids);
  else wgrad_kernel<false><<<grid, WG_THREADS, wg_smem(a_panels), st>>>(g, tm_a, tm_b, tm_ids);
  AERO_LAUNCH_CHECK();
  if (seg) {
    int rc = launch_gpd_fixup(g.seg_part, rowptr, dst, g.seg_out, seg_ld, rows, n_nodes, st);
    if (rc) return rc;
  }
  return launch_reduce_partials(g.part, grid, (size_t)a_panels * 16384, dW, (size_t)a_panels * 16384, st);
}
