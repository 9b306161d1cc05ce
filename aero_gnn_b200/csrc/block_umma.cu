// block_umma.cu -- fused MGN block on the 5th-generation tensor cores (tcgen05.mma, bf16 operands, fp32
// accumulators in TMEM).  bf16 storage path of aero_block_fwd / aero_block_bwd.
//
// Forward kernel (one persistent CTA per SM, two independent 128-thread warpgroups per CTA):
//   * all (L+2) weight matrices of the block live in shared memory for the whole kernel as bf16 SWIZZLE_128B
//     row tiles (umma.cuh), loaded once per CTA;
//   * a warpgroup owns a 128-row tile: it stages the rows into its activation tile, one elected thread issues the
//     8 tcgen05.mma (K = 16 each) of a 128x128x128 GEMM into the warpgroup's 128 TMEM columns and commits to an
//     mbarrier; the 128 threads then read the accumulator with tcgen05.ld (thread = row), apply the epilogue
//     (gathered pre-projections / bias, activation, bf16 pack) and write the next GEMM's A operand back into the
//     same activation tile;
//   * the last epilogue keeps the whole row in registers: bias, LayerNorm (fp32 statistics), residual, bf16 pack;
//     the output tile goes through shared memory so the global store is coalesced and the receiver sums
//     (segmented, CSR order, fp32) read it column-wise;
//   * while one warpgroup is in an epilogue the other one's MMAs keep the tensor pipe busy.
#include "umma_block.cuh"

namespace aero {

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
template <int NWG>
__global__ void __launch_bounds__(NWG * 128, 1) umma_block_fwd_kernel(UmmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int L = a.L;
  uint8_t* Wimg = smem;
  uint8_t* Abuf = Wimg + (size_t)(L + 2) * TILE_BYTES;
  float* vec = reinterpret_cast<float*>(Abuf + (size_t)NWG * TILE_BYTES);
  int* sidx = reinterpret_cast<int*>(vec + (L + 3) * 128);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sidx + NWG * 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + NWG);

  const int tid = threadIdx.x, wg = tid >> 7, wt = tid & 127, lane = tid & 31, q = (tid >> 5) & 3;
  // weights + vectors, once per CTA
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.prep);
    uint4* dst = reinterpret_cast<uint4*>(Wimg);
    const int n16 = (L + 2) * (TILE_BYTES / 16);
    for (int i = tid; i < n16; i += NWG * 128) dst[i] = src[i];
    const float* vs = reinterpret_cast<const float*>(a.prep + (size_t)(L + 2) * TILE_BYTES);
    for (int i = tid; i < (L + 3) * 128; i += NWG * 128) vec[i] = vs[i];
  }
  if (tid == 0) {
    for (int w = 0; w < NWG; ++w) mbar_init(smem_u32(&mbar[w]), 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<NWG * 128>(tmem_slot);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tacc = tmem_base + (uint32_t)wg * 128u;
  const uint32_t tlane = tacc + ((uint32_t)(q * 32) << 16);   // this warp's lane quarter
  uint8_t* A = Abuf + (size_t)wg * TILE_BYTES;
  const uint32_t a_s = smem_u32(A);
  const uint32_t w_s = smem_u32(Wimg);
  const uint32_t bar_s = smem_u32(&mbar[wg]);
  int* sidx0 = sidx + wg * 256;
  int* sidx1 = sidx0 + 128;
  uint32_t phase = 0;
  const int act = a.act;
  const int row = wt;   // epilogue: thread = tile row = TMEM lane

  const int64_t tiles = (a.rows + 127) / 128;
  for (int64_t tile = (int64_t)blockIdx.x * NWG + wg; tile < tiles; tile += (int64_t)gridDim.x * NWG) {
    const int64_t row0 = tile * 128;
    const int nrows = (int)((a.rows - row0) < 128 ? (a.rows - row0) : 128);
    wg_sync(1 + wg);   // previous tile of this warpgroup fully consumed
    if (a.main_f32) stage_rows<true>(A, a.main, a.main_scale, row0, nrows, wt);
    else stage_rows<false>(A, a.main, nullptr, row0, nrows, wt);
    {
      int64_t r = row0 + wt;
      bool ok = wt < nrows;
      sidx0[wt] = ok ? (a.idx0 ? a.idx0[r] : (int)r) : 0;
      sidx1[wt] = ok ? (a.idx1 ? a.idx1[r] : -1) : -1;
    }
    fence_async_smem();
    wg_sync(1 + wg);

    for (int layer = 0; layer <= L + 1; ++layer) {
      if (wt == 0) {
        fence_after_sync();
        issue_gemm(tacc, a_s, false, w_s + (uint32_t)layer * TILE_BYTES, false, false);
        mma_commit(bar_s);
      }
      mbar_wait(bar_s, phase);
      phase ^= 1;
      fence_after_sync();

      if (layer <= L) {
        // hidden epilogue: (+ gathered pre-projections | + bias), activation, bf16, back into the A tile
        const bool valid = row < nrows;
        const __nv_bfloat16* p0 = nullptr;
        const __nv_bfloat16* p1 = nullptr;
        if (layer == 0 && valid) {
          p0 = a.P + (int64_t)sidx0[row] * a.ldp + a.poff0;
          if (sidx1[row] >= 0) p1 = a.P + (int64_t)sidx1[row] * a.ldp + a.poff1;
        }
        const float* bias = layer > 0 ? vec + (layer - 1) * 128 : nullptr;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint4 g0[4], g1[4];
          if (layer == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              g0[j] = p0 ? *reinterpret_cast<const uint4*>(p0 + c * 32 + j * 8) : make_uint4(0u, 0u, 0u, 0u);
              g1[j] = p1 ? *reinterpret_cast<const uint4*>(p1 + c * 32 + j * 8) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
          float v[32];
          tmem_ld32(tlane + (uint32_t)(c * 32), v);
          if (layer == 0) {
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
          store_row32(A, row, c, v);
        }
        fence_before_sync();
        fence_async_smem();
        wg_sync(1 + wg);
      } else {
        // output epilogue: bias, LayerNorm, residual; whole row in registers
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
          const float* bet = vec + (L + 2) * 128;
#pragma unroll
          for (int j = 0; j < 128; ++j) v[j] = fmaf((v[j] - mean) * rstd, gam[j], bet[j]);
        }
        const bool valid = row < nrows;
        if (a.resid && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.resid + (row0 + row) * 128);
#pragma unroll
          for (int j = 0; j < 16; ++j) add_bf16x8(v + 8 * j, rp[j]);
        }
        if (!valid) {
#pragma unroll
          for (int j = 0; j < 128; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) store_row32(A, row, c, v + 32 * c);
        wg_sync(1 + wg);
        // coalesced store of the output tile
        {
          const int chunk = wt & 15;
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            int r = (wt >> 4) + i * 8;
            if (r < nrows)
              *reinterpret_cast<uint4*>(a.out + (row0 + r) * 128 + chunk * 8) =
                  *reinterpret_cast<const uint4*>(A + tile_chunk_off(r, chunk));
          }
        }
        if (a.agg) {
          // receiver sums over the bf16-rounded rows: thread = column, CSR order
          const int c = wt;
          const uint8_t* colp = A + (c >> 6) * PANEL_BYTES + (c & 7) * 2;
          const int cc = (c >> 3) & 7;
          const int64_t tile_end = row0 + nrows;
          int r = 0;
          while (r < nrows) {
            int n = sidx1[r];
            int b = a.rowptr[n], e = a.rowptr[n + 1];
            int re = (int)(((int64_t)e < tile_end ? (int64_t)e : tile_end) - row0);
            float s = 0.f;
            for (int t = r; t < re; ++t) {
              uint16_t h = *reinterpret_cast<const uint16_t*>(colp + t * 128 + ((cc ^ (t & 7)) << 4));
              s += __uint_as_float((uint32_t)h << 16);
            }
            bool complete = (b >= row0) && (e <= tile_end);
            if (complete) a.agg[(size_t)n * 128 + c] = s;
            else a.agg_part[((size_t)tile * 2 + (r == 0 ? 0 : 1)) * 128 + c] = s;
            r = re;
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<NWG * 128>(tmem_base);
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
  stage_rows<false>(At, Ag, nullptr, 0, 128, tid);
  stage_rows<false>(Bt, Bg, nullptr, 0, 128, tid);
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
constexpr int FWD_NWG = 2;

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

static size_t fwd_smem(int L, int nwg) {
  return 1024 + (size_t)(L + 2 + nwg) * TILE_BYTES + (size_t)(L + 3) * 512 + (size_t)nwg * 1024 + (size_t)nwg * 8 + 16;
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
    AERO_CUDA(cudaMemsetAsync(d->agg, 0, (size_t)d->n_nodes * 128 * sizeof(float), st));
  }
  if (d->rows == 0) return AERO_OK;
  static bool attr_set = false;
  size_t smem = fwd_smem(d->L, FWD_NWG);
  if (!attr_set) {
    AERO_CUDA(cudaFuncSetAttribute(umma_block_fwd_kernel<FWD_NWG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)fwd_smem(UMMA_MAX_L, FWD_NWG)));
    attr_set = true;
  }
  int64_t tiles = cdiv(d->rows, 128);
  int grid = (int)(cdiv(tiles, FWD_NWG) < sm_count() ? cdiv(tiles, FWD_NWG) : sm_count());
  umma_block_fwd_kernel<FWD_NWG><<<grid, FWD_NWG * 128, smem, st>>>(a);
  AERO_LAUNCH_CHECK();
  if (d->agg) return launch_agg_fixup(a.agg_part, d->rowptr, d->agg, d->rows, d->n_nodes, 128, d->idx1, st);
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
