// rowgemm.cu -- the two row GEMMs of a processor step that are not inside a fused block, on tcgen05, warp-specialised
// and fed by TMA:
//
//   projection    P[rows, 128 nb]  = x[rows,128] W^T + b              W = [W_s; W_d; W_nx] (nb = 3 blocks of 128 outputs)
//                 -- the sum-trick pre-projection of the node latents (mgnLayer.py:97-103: the first Linear of the edge
//                 block applied per node instead of per edge; processor.py header)
//   back-proj.    g_x[rows,128]    = sum_ka A_ka[rows,128] W[128 ka : 128 ka + 128, :]  + G_x
//                 -- A_0 | A_1 | A_2 = g_P_s | g_P_d | g_h0n: the gradient of the node latents through that projection,
//                 one K = 384 contraction with the incoming gradient added in the epilogue (one rounding to bf16)
//
// Both are HBM-bound streaming over the rows with the 3 weight tiles resident in shared memory.  One persistent CTA
// per SM:
//   warp 0      TMA producer: the weight tiles once, then one 128-row A tile per ring slot (3 slots)
//   warp 1      MMA issuer: per output block `na` tcgen05.mma chains into one of 4 TMEM accumulators; tcgen05.commit
//               frees the ring slot (after its last use) and hands the accumulator to the epilogue
//   warps 2..9  epilogue: TMEM -> registers (+ bias / + addend rows) -> bf16 -> SWIZZLE_128B staging tile -> one TMA store
//               per output block (out-of-range rows are clipped by the tensor map)
// Exactly one of na / nb may exceed 1 (na * nb <= 3 weight tiles).
#include "umma_block.cuh"
#include "tma.cuh"

namespace aero {

constexpr int RG_THREADS = 320;
constexpr int RG_SLOTS = 3;
constexpr int RG_ACC = 4;
constexpr int RG_EPI = 256;   // epilogue threads (warps 2..9)

struct RowGemmArgs {
  int64_t rows, add_ld;
  int na, nb, w_mn;
  const __nv_bfloat16* bias;   // [128 nb] or null
  const __nv_bfloat16* add;    // [rows, 128 nb] with row stride add_ld, or null
};

__device__ __forceinline__ void mbar_arrive(uint32_t saddr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(saddr) : "memory");
}

__global__ void __launch_bounds__(RG_THREADS, 1)
row_gemm_kernel(RowGemmArgs g, const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_w,
                const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* Wt = smem;                                        // 3 weight tiles
  uint8_t* ring = Wt + 3 * TILE_BYTES;                       // RG_SLOTS A tiles
  uint8_t* stage = ring + RG_SLOTS * TILE_BYTES;             // output staging tile
  uint64_t* mbar = reinterpret_cast<uint64_t*>(stage + TILE_BYTES);
  // [0..2] slot full, [3..5] slot empty, [6..9] accumulator full, [10..13] accumulator empty, [14] weights
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 15);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int na = g.na, nb = g.nb;
  if (tid == 0) {
    for (int s = 0; s < 2 * RG_SLOTS; ++s) mbar_init(smem_u32(&mbar[s]), 1);
    for (int b = 0; b < RG_ACC; ++b) {
      mbar_init(smem_u32(&mbar[6 + b]), 1);
      mbar_init(smem_u32(&mbar[10 + b]), RG_EPI / 32);
    }
    mbar_init(smem_u32(&mbar[14]), 1);
    fence_mbar_init();
  }
  if (wid == 0) tmem_alloc<RG_ACC * 128>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t w_s = smem_u32(Wt), ring_s = smem_u32(ring);
  const int64_t tiles = (g.rows + 127) / 128;
  const int64_t my_tiles = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (wid == 0) {
    // ---- TMA producer ----
    if (lane == 0) {
      tma::prefetch_map(&tm_a0);
      tma::prefetch_map(&tm_w);
      tma::prefetch_map(&tm_out);
      const int nw = na * nb;
      mbar_expect_tx(smem_u32(&mbar[14]), (uint32_t)nw * TILE_BYTES);
      for (int i = 0; i < nw; ++i) tma::load_tile(w_s + (uint32_t)i * TILE_BYTES, &tm_w, 128 * i, smem_u32(&mbar[14]));
      uint32_t cnt = 0;
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int row0 = (int)((blockIdx.x + i * gridDim.x) * 128);
        for (int ka = 0; ka < na; ++ka, ++cnt) {
          const uint32_t s = cnt % RG_SLOTS, use = cnt / RG_SLOTS;
          if (use > 0) mbar_wait(smem_u32(&mbar[3 + s]), (use - 1) & 1);
          const uint32_t full = smem_u32(&mbar[s]);
          mbar_expect_tx(full, TILE_BYTES);
          tma::load_tile(ring_s + s * TILE_BYTES, ka == 0 ? &tm_a0 : (ka == 1 ? &tm_a1 : &tm_a2), row0, full);
        }
      }
    }
    __syncwarp();
  } else if (wid == 1) {
    // ---- MMA issuer (lane 0 issues every MMA and every commit, so each commit covers the MMAs before it) ----
    mbar_wait(smem_u32(&mbar[14]), 0);
    uint32_t cnt = 0, item = 0;
    for (int64_t i = 0; i < my_tiles; ++i, cnt += (uint32_t)na) {
      for (int nbi = 0; nbi < nb; ++nbi, ++item) {
        const uint32_t b = item % RG_ACC, useb = item / RG_ACC;
        if (useb > 0) mbar_wait(smem_u32(&mbar[10 + b]), (useb - 1) & 1);   // the epilogue has drained this accumulator
        for (int ka = 0; ka < na; ++ka) {
          const uint32_t c = cnt + (uint32_t)ka, s = c % RG_SLOTS;
          if (nbi == 0) mbar_wait(smem_u32(&mbar[s]), (c / RG_SLOTS) & 1);
          fence_after_sync();
          if (lane == 0) {
            issue_gemm(tmem_base + 128u * b, ring_s + s * TILE_BYTES, false, w_s + (uint32_t)(ka * nb + nbi) * TILE_BYTES,
                       g.w_mn != 0, ka > 0);
            if (nbi == nb - 1) mma_commit(smem_u32(&mbar[3 + s]));   // last reader of the slot
          }
          __syncwarp();
        }
        if (lane == 0) mma_commit(smem_u32(&mbar[6 + b]));
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: warps 2..9; TMEM lane quarter = warp id % 4, column half = (warp - 2) / 4 ----
    const int et = tid - 64;
    const int lq = wid & 3, half = (wid - 2) >> 2;
    const int row = lq * 32 + lane;
    uint32_t item = 0;
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int64_t row0 = (int64_t)(blockIdx.x + i * gridDim.x) * 128;
      for (int nbi = 0; nbi < nb; ++nbi, ++item) {
        const uint32_t b = item % RG_ACC;
        mbar_wait(smem_u32(&mbar[6 + b]), (item / RG_ACC) & 1);
        fence_after_sync();
        float v0[32], v1[32];
        const uint32_t tl = tmem_base + ((uint32_t)(lq * 32) << 16) + 128u * b + (uint32_t)(half * 64);
        tmem_ld32(tl, v0);
        tmem_ld32(tl + 32u, v1);
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&mbar[10 + b]));
        if (g.bias) {
          const uint4* bp = reinterpret_cast<const uint4*>(g.bias + nbi * 128 + half * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            add_bf16x8(v0 + 8 * j, __ldg(bp + j));
            add_bf16x8(v1 + 8 * j, __ldg(bp + 4 + j));
          }
        }
        if (g.add && row0 + row < g.rows) {
          const uint4* ap = reinterpret_cast<const uint4*>(g.add + (size_t)(row0 + row) * g.add_ld + nbi * 128 + half * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            add_bf16x8(v0 + 8 * j, __ldg(ap + j));
            add_bf16x8(v1 + 8 * j, __ldg(ap + 4 + j));
          }
        }
        if (et == 0) tma::store_wait_read();        // the previous block's store has left the staging tile
        named_sync(1, RG_EPI);
        store_row32(stage, row, half * 2, v0);
        store_row32(stage, row, half * 2 + 1, v1);
        fence_async_smem();
        named_sync(1, RG_EPI);
        if (et == 0) {
          tma::store_tile_at(&tm_out, smem_u32(stage), nbi * 128, (int)row0);
          tma::store_commit();
        }
      }
    }
    if (et == 0) tma::store_wait_all();
  }
  fence_before_sync();
  __syncthreads();
  if (wid == 0) tmem_dealloc<RG_ACC * 128>(tmem_base);
}

static size_t rg_smem() { return 1024 + (size_t)(3 + RG_SLOTS + 1) * TILE_BYTES + 15 * 8 + 16; }

}  // namespace aero

using namespace aero;

extern "C" int aero_row_gemm(const void* const* a_blocks, const int64_t* a_ld, int na, const void* W, int w_mn, int nb,
                             const void* bias, const void* add, int64_t add_ld, void* out, int64_t out_ld, int64_t rows,
                             void* stream) {
  g_launch_count = 0;
  cudaStream_t st = (cudaStream_t)stream;
  AERO_CHECK_ARG(na >= 1 && nb >= 1 && na * nb <= 3 && (na == 1 || nb == 1) && rows >= 0, "aero_row_gemm: na x nb must be 1 x {1,2,3} or {1,2,3} x 1");
  if (rows == 0) return AERO_OK;
  AERO_CHECK_ARG(a_blocks && a_ld && W && out && out_ld >= 128 * nb && (out_ld % 8) == 0, "aero_row_gemm: bad arguments");
  AERO_CHECK_ARG(!add || (add_ld >= 128 * nb && (add_ld % 8) == 0 && ((uintptr_t)add & 15) == 0), "aero_row_gemm: addend rows must be 16-byte aligned");
  AERO_CHECK_ARG(!bias || ((uintptr_t)bias & 15) == 0, "aero_row_gemm: bias must be 16-byte aligned");
  CUtensorMap tm_a[3], tm_w, tm_out;
  for (int k = 0; k < 3; ++k) {
    const int kk = k < na ? k : 0;
    if (!a_blocks[kk] || tma::make_rows_map_ld(a_blocks[kk], rows, 128, a_ld[kk], &tm_a[k])) {
      set_error("aero_row_gemm: cuTensorMapEncodeTiled failed for A block %d (16-byte aligned bf16 rows, ld %% 8 == 0)", kk);
      return AERO_ECUDA;
    }
  }
  if (tma::make_rows_map(W, 128 * (int64_t)(na * nb), &tm_w) || tma::make_rows_map_ld(out, rows, 128 * nb, out_ld, &tm_out)) {
    set_error("aero_row_gemm: cuTensorMapEncodeTiled failed for W / out");
    return AERO_ECUDA;
  }
  RowGemmArgs g;
  g.rows = rows; g.add_ld = add_ld; g.na = na; g.nb = nb; g.w_mn = w_mn;
  g.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  g.add = reinterpret_cast<const __nv_bfloat16*>(add);
  static bool attr_set[64] = {false};
  int dev = 0;
  AERO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    AERO_CUDA(cudaFuncSetAttribute(row_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rg_smem()));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t tiles = cdiv(rows, 128);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  row_gemm_kernel<<<grid, RG_THREADS, rg_smem(), st>>>(g, tm_a[0], tm_a[1], tm_a[2], tm_w, tm_out);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}
