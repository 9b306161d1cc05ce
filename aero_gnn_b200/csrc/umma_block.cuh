// umma_block.cuh -- argument block and device helpers shared by the tcgen05 forward / backward block kernels.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace aero {
using namespace umma;

struct UmmaArgs {
  int L, act, use_ln, main_f32, has_resid_grad;
  int64_t rows, n_nodes, ldp, poff0, poff1;
  const void* main;
  const float* main_scale;
  const __nv_bfloat16* resid;
  const __nv_bfloat16* P;
  const int32_t* idx0;
  const int32_t* idx1;
  const int32_t* rowptr;
  const uint8_t* prep;
  __nv_bfloat16* out;
  float* agg;
  float* agg_part;
  const __nv_bfloat16* g_out;
  const float* g_agg;
  void* g_main;
  __nv_bfloat16* g_h0;
  float* w_part;
};

// ---- helpers --------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  uint32_t s = smem_u32(p);
  return p + (((s + 1023u) & ~1023u) - s);
}

// stage rows [row0, row0+nrows) of a [rows,128] matrix into a row tile (zero padded), by one warpgroup
template <bool F32>
__device__ __forceinline__ void stage_rows(uint8_t* tile, const void* src, const float* scale, int64_t row0, int nrows,
                                           int wt) {
  const int chunk = wt & 15;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    int r = (wt >> 4) + i * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < nrows) {
      if (F32) {
        const float* p = reinterpret_cast<const float*>(src) + (row0 + r) * 128 + chunk * 8;
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        float s = scale ? scale[row0 + r] : 1.f;
        v.x = pack_bf16(a.x * s, a.y * s);
        v.y = pack_bf16(a.z * s, a.w * s);
        v.z = pack_bf16(b.x * s, b.y * s);
        v.w = pack_bf16(b.z * s, b.w * s);
      } else {
        v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + (row0 + r) * 128 + chunk * 8);
      }
    }
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(r, chunk)) = v;
  }
}

// one 128x128x128 GEMM: D(tmem) = A(tile) * B(tile), issued by a single thread
__device__ __forceinline__ void issue_gemm(uint32_t tacc, uint32_t a_saddr, bool a_mn, uint32_t b_saddr, bool b_mn,
                                           bool accumulate_first) {
  const uint32_t idesc = make_idesc(a_mn, b_mn);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    uint64_t ad = a_mn ? desc_mnmajor(a_saddr, kk) : desc_kmajor(a_saddr, kk);
    uint64_t bd = b_mn ? desc_mnmajor(b_saddr, kk) : desc_kmajor(b_saddr, kk);
    mma_bf16(tacc, ad, bd, idesc, accumulate_first || kk > 0);
  }
}

__device__ __forceinline__ void add_bf16x8(float* v, uint4 q) {
  v[0] += bf16_lo(q.x); v[1] += bf16_hi(q.x); v[2] += bf16_lo(q.y); v[3] += bf16_hi(q.y);
  v[4] += bf16_lo(q.z); v[5] += bf16_hi(q.z); v[6] += bf16_lo(q.w); v[7] += bf16_hi(q.w);
}

// write 32 activations of `row` (columns 32*c32 ..) as bf16 into a row tile
__device__ __forceinline__ void store_row32(uint8_t* tile, int row, int c32, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
    q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
    q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
    q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(row, c32 * 4 + j)) = q;
  }
}

__device__ __forceinline__ float relu_or_act(float v, int act) { return act == AERO_ACT_RELU ? fmaxf(v, 0.f) : act_fwd(v, act); }

constexpr int UMMA_MAX_L = 2;
constexpr int UMMA_MAX_L_BWD = 2;
size_t umma_bwd_workspace_bytes(const aero_block_desc* d);

static inline UmmaArgs make_uargs(const aero_block_desc* d) {
  UmmaArgs a;
  a.L = d->L; a.act = d->act; a.use_ln = d->use_ln; a.main_f32 = d->main_f32; a.has_resid_grad = d->has_resid_grad;
  a.rows = d->rows; a.n_nodes = d->n_nodes; a.ldp = d->ldp; a.poff0 = d->poff0; a.poff1 = d->poff1;
  a.main = d->main; a.main_scale = d->main_scale;
  a.resid = reinterpret_cast<const __nv_bfloat16*>(d->resid);
  a.P = reinterpret_cast<const __nv_bfloat16*>(d->P);
  a.idx0 = d->idx0; a.idx1 = d->idx1; a.rowptr = d->rowptr;
  a.prep = reinterpret_cast<const uint8_t*>(d->prepared);
  a.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  a.agg = d->agg; a.agg_part = nullptr;
  a.g_out = reinterpret_cast<const __nv_bfloat16*>(d->g_out);
  a.g_agg = d->g_agg; a.g_main = d->g_main;
  a.g_h0 = reinterpret_cast<__nv_bfloat16*>(d->g_h0);
  a.w_part = nullptr;
  return a;
}

}  // namespace aero
