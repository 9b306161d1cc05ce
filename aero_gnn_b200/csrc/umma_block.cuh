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
  __nv_bfloat16* h0;   // fwd: optional output; bwd: optional input replacing the layer-0 recompute
  __nv_bfloat16* main_lat;   // fwd, main_f32: optional bf16 copy of the staged (scaled) main rows
  __nv_bfloat16* hh[2];      // fwd: optional outputs H_1, H_2 (kept for a backward without recompute)
};

constexpr int UMMA_MAX_L = 2;
constexpr int UMMA_MAX_L_BWD = 2;
size_t umma_bwd_workspace_bytes(const aero_block_desc* d);
// TMA-fed backward (block_umma_bwd2.cu): kept h_0, 1 <= L <= 2
bool umma_bwd2_applicable(const aero_block_desc* d);

// ---- helpers --------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  uint32_t s = smem_u32(p);
  return p + (((s + 1023u) & ~1023u) - s);
}

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// stage rows [row0, row0+nrows) of a [rows,128] matrix into a row tile (zero padded) with NT cooperating
// threads (t = 0..NT-1): a warp covers two whole rows per pass, so global reads are fully coalesced
template <bool F32, int NT>
__device__ __forceinline__ void stage_rows(uint8_t* tile, const void* src, const float* scale, int64_t row0, int nrows,
                                           int t) {
  const int chunk = t & 15;
  constexpr int RPP = NT / 16;   // rows per pass
  // four rows per batch, every global load of a batch in flight before its first shared-memory store
#pragma unroll 1
  for (int b0 = 0; b0 < 128 / RPP; b0 += 4) {
    if (F32) {
      float4 a[4], b[4];
      float s[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (t >> 4) + (b0 + i) * RPP;
        a[i] = b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        s[i] = 1.f;
        if (r < nrows) {
          const float* p = reinterpret_cast<const float*>(src) + (row0 + r) * 128 + chunk * 8;
          a[i] = *reinterpret_cast<const float4*>(p);
          b[i] = *reinterpret_cast<const float4*>(p + 4);
          if (scale) s[i] = scale[row0 + r];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (t >> 4) + (b0 + i) * RPP;
        uint4 v;
        v.x = pack_bf16(a[i].x * s[i], a[i].y * s[i]);
        v.y = pack_bf16(a[i].z * s[i], a[i].w * s[i]);
        v.z = pack_bf16(b[i].x * s[i], b[i].y * s[i]);
        v.w = pack_bf16(b[i].z * s[i], b[i].w * s[i]);
        *reinterpret_cast<uint4*>(tile + tile_chunk_off(r, chunk)) = v;
      }
    } else {
      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (t >> 4) + (b0 + i) * RPP;
        v[i] = make_uint4(0u, 0u, 0u, 0u);
        if (r < nrows)
          v[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + (row0 + r) * 128 + chunk * 8);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (t >> 4) + (b0 + i) * RPP;
        *reinterpret_cast<uint4*>(tile + tile_chunk_off(r, chunk)) = v[i];
      }
    }
  }
}

// copy the valid rows of a row tile to a [rows,128] bf16 matrix (coalesced), NT cooperating threads
template <int NT>
__device__ __forceinline__ void unstage_rows(const uint8_t* tile, __nv_bfloat16* dst, int64_t row0, int nrows, int t) {
  const int chunk = t & 15;
  constexpr int RPP = NT / 16;
#pragma unroll 4
  for (int i = 0; i < 128 / RPP; ++i) {
    int r = (t >> 4) + i * RPP;
    if (r < nrows)
      *reinterpret_cast<uint4*>(dst + (row0 + r) * 128 + chunk * 8) = *reinterpret_cast<const uint4*>(tile + tile_chunk_off(r, chunk));
  }
}

// One batch (rows (t>>4) + (b..b+3) * NT/16) of the gather below, split into its load half and its store half so a
// caller can keep the loads in flight across a wait.
template <int NT>
__device__ __forceinline__ void gather_load4(const __nv_bfloat16* P, int64_t ldp, int64_t off0, int64_t off1,
                                             const int* sidx0, const int* sidx1, int nrows, int t, int b,
                                             uint4 (&v)[4], uint4 (&q)[4]) {
  const int chunk = t & 15;
  constexpr int RPP = NT / 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (t >> 4) + (b + i) * RPP;
    v[i] = q[i] = make_uint4(0u, 0u, 0u, 0u);
    if (r < nrows) {
      v[i] = *reinterpret_cast<const uint4*>(P + (int64_t)sidx0[r] * ldp + off0 + chunk * 8);
      const int i1 = sidx1[r];
      if (i1 >= 0) q[i] = *reinterpret_cast<const uint4*>(P + (int64_t)i1 * ldp + off1 + chunk * 8);
    }
  }
}
template <int NT>
__device__ __forceinline__ void gather_store4(uint8_t* tile, const int* sidx1, int nrows, int t, int b,
                                              const uint4 (&v)[4], const uint4 (&q)[4]) {
  const int chunk = t & 15;
  constexpr int RPP = NT / 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (t >> 4) + (b + i) * RPP;
    const bool two = r < nrows && sidx1[r] >= 0;   // a single operand passes through bit-exactly
    uint4 o = v[i];
    if (two) {
      o.x = add_bf16x2(v[i].x, q[i].x);
      o.y = add_bf16x2(v[i].y, q[i].y);
      o.z = add_bf16x2(v[i].z, q[i].z);
      o.w = add_bf16x2(v[i].w, q[i].w);
    }
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(r, chunk)) = o;
  }
}

// Coalesced gather of the pre-projected rows: tile[r] = bf16( P[idx0[r]] (+ P[idx1[r]]) ) for the rows of one tile.
// A half-warp reads one whole 256-byte row, so every global request touches 2-4 cache lines instead of the 32 a
// thread-per-row access would; the epilogue then reads its (row, chunk) from shared memory.
template <int NT>
__device__ __forceinline__ void stage_gather_sum(uint8_t* tile, const __nv_bfloat16* P, int64_t ldp, int64_t off0,
                                                 int64_t off1, const int* sidx0, const int* sidx1, int nrows, int t) {
  // four rows per batch: all eight global loads of a batch are in flight before its first shared-memory store
#pragma unroll 1
  for (int b = 0; b < 128 / (NT / 16); b += 4) {
    uint4 v[4], q[4];
    gather_load4<NT>(P, ldp, off0, off1, sidx0, sidx1, nrows, t, b, v, q);
    gather_store4<NT>(tile, sidx1, nrows, t, b, v, q);
  }
}

// v[0..31] += the 32 bf16 values stored at (row, chunk c) of a row tile
__device__ __forceinline__ void add_tile_chunk(float* v, const uint8_t* tile, int row, int c);

// one 128x128x128 GEMM: D(tmem) = A(tile) * B(tile), issued by a single thread
__device__ __forceinline__ void issue_gemm(uint32_t tacc, uint32_t a_saddr, bool a_mn, uint32_t b_saddr, bool b_mn,
                                           bool accumulate_first) {
  const uint32_t idesc = make_idesc(a_mn, b_mn);
  // descriptors of K-step kk differ from those of step 0 only in the 14-bit start-address field (16-byte units):
  // K-major +2 per step inside a panel, +1024 for the second panel; MN-major +128 per step
  const uint64_t ad0 = a_mn ? desc_mnmajor(a_saddr, 0) : desc_kmajor(a_saddr, 0);
  const uint64_t bd0 = b_mn ? desc_mnmajor(b_saddr, 0) : desc_kmajor(b_saddr, 0);
  const uint32_t ahi = (uint32_t)(ad0 >> 32), bhi = (uint32_t)(bd0 >> 32);
  const uint32_t alo = (uint32_t)ad0, blo = (uint32_t)bd0;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint32_t ka = a_mn ? (uint32_t)kk * 128u : (uint32_t)((kk >> 2) * 1024 + (kk & 3) * 2);
    const uint32_t kb = b_mn ? (uint32_t)kk * 128u : (uint32_t)((kk >> 2) * 1024 + (kk & 3) * 2);
    const uint64_t ad = ((uint64_t)ahi << 32) | (uint64_t)(alo + ka);
    const uint64_t bd = ((uint64_t)bhi << 32) | (uint64_t)(blo + kb);
    mma_bf16(tacc, ad, bd, idesc, accumulate_first || kk > 0);
  }
}

__device__ __forceinline__ void add_bf16x8(float* v, uint4 q) {
  v[0] += bf16_lo(q.x); v[1] += bf16_hi(q.x); v[2] += bf16_lo(q.y); v[3] += bf16_hi(q.y);
  v[4] += bf16_lo(q.z); v[5] += bf16_hi(q.z); v[6] += bf16_lo(q.w); v[7] += bf16_hi(q.w);
}

__device__ __forceinline__ void add_tile_chunk(float* v, const uint8_t* tile, int row, int c) {
#pragma unroll
  for (int j = 0; j < 4; ++j) add_bf16x8(v + 8 * j, *reinterpret_cast<const uint4*>(tile + tile_chunk_off(row, c * 4 + j)));
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

// same, with ReLU fused into the bf16 conversion
__device__ __forceinline__ void store_row32_relu(uint8_t* tile, int row, int c32, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_bf16_relu(v[8 * j + 0], v[8 * j + 1]);
    q.y = pack_bf16_relu(v[8 * j + 2], v[8 * j + 3]);
    q.z = pack_bf16_relu(v[8 * j + 4], v[8 * j + 5]);
    q.w = pack_bf16_relu(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(row, c32 * 4 + j)) = q;
  }
}

__device__ __forceinline__ float relu_or_act(float v, int act) { return act == AERO_ACT_RELU ? fmaxf(v, 0.f) : act_fwd(v, act); }

// hidden-layer epilogue for one (row, 32-column chunk): accumulator + (gathered pre-projections | bias),
// activation, bf16, into the destination row tile.  p0/p1: gathered rows (layer 0) or nullptr; bias: smem.
__device__ __forceinline__ void hidden_epilogue_chunk(uint32_t tacc_lane, int c, const __nv_bfloat16* p0,
                                                      const __nv_bfloat16* p1, const float* bias, int act,
                                                      uint8_t* dst_tile, int row) {
  uint4 g0[4], g1[4];
  if (p0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) g0[j] = *reinterpret_cast<const uint4*>(p0 + c * 32 + j * 8);
  }
  if (p1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) g1[j] = *reinterpret_cast<const uint4*>(p1 + c * 32 + j * 8);
  }
  float v[32];
  tmem_ld32(tacc_lane + (uint32_t)(c * 32), v);
  if (p0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) add_bf16x8(v + 8 * j, g0[j]);
  }
  if (p1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) add_bf16x8(v + 8 * j, g1[j]);
  }
  if (bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b4 = *reinterpret_cast<const float4*>(bias + c * 32 + 4 * j);
      v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
    }
  }
  if (act == AERO_ACT_RELU) {   // uniform branch: the transcendental activations stay out of the hot path
    store_row32_relu(dst_tile, row, c, v);
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], act);
  store_row32(dst_tile, row, c, v);
}

// first-layer epilogue for one (row, chunk): accumulator + staged gathered sum (read from `tile`), activation, bf16,
// written back IN PLACE over the staged values (same thread, same bytes)
__device__ __forceinline__ void first_epilogue_chunk(uint32_t tacc_lane, int c, int act, uint8_t* tile, int row) {
  float v[32];
  tmem_ld32(tacc_lane + (uint32_t)(c * 32), v);
  add_tile_chunk(v, tile, row, c);
  if (act == AERO_ACT_RELU) {
    store_row32_relu(tile, row, c, v);
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], act);
  store_row32(tile, row, c, v);
}

// v[j] *= act'(h[j]) for the 32 activations h stored at (row, chunk c) of a bf16 row tile
__device__ __forceinline__ void mask_by_act_grad(float* v, const uint8_t* h_tile_ptr, int row, int c, int act) {
  uint32_t hh[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 h4 = *reinterpret_cast<const uint4*>(h_tile_ptr + tile_chunk_off(row, c * 4 + j));
    hh[4 * j] = h4.x; hh[4 * j + 1] = h4.y; hh[4 * j + 2] = h4.z; hh[4 * j + 3] = h4.w;
  }
  if (act == AERO_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // bf16 sign/zero tests on the packed halves: h > 0  <=>  (bits & 0x7fff) != 0 and sign clear
      v[2 * j] = (int)(hh[j] << 16) > 0 ? v[2 * j] : 0.f;
      v[2 * j + 1] = (int)(hh[j] & 0xffff0000u) > 0 ? v[2 * j + 1] : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[2 * j] *= act_grad_from_out(bf16_lo(hh[j]), act);
      v[2 * j + 1] *= act_grad_from_out(bf16_hi(hh[j]), act);
    }
  }
}

// Column sums of a bf16 row tile by a 512-thread CTA: warp w owns the 8 columns of chunk w, lane l sums rows
// l, l+32, l+64, l+96, then a recursive-halving shuffle reduction leaves the total of column 8*w + (l>>2 & 7)...
// precisely: after the reduction lane l holds the total of column 8*w + colsel(l) where colsel(l) = ((l>>4)&1)*4 +
// ((l>>3)&1)*2 + ((l>>2)&1); lanes with (l & 3) == 0 are the designated owners.  Fixed order -> deterministic.
__device__ __forceinline__ float tile_col_sums_512(const uint8_t* tile, int warp, int lane) {
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = lane + 32 * i;
    uint4 q = *reinterpret_cast<const uint4*>(tile + tile_chunk_off(r, warp));
    s[0] += bf16_lo(q.x); s[1] += bf16_hi(q.x); s[2] += bf16_lo(q.y); s[3] += bf16_hi(q.y);
    s[4] += bf16_lo(q.z); s[5] += bf16_hi(q.z); s[6] += bf16_lo(q.w); s[7] += bf16_hi(q.w);
  }
  // halve the vector, double the lane span: xor 16 keeps 4 values, xor 8 keeps 2, xor 4 keeps 1
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float t[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float mine = b4 ? s[4 + j] : s[j], give = b4 ? s[j] : s[4 + j];
    t[j] = mine + __shfl_xor_sync(0xffffffffu, give, 16);
  }
  float u[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float mine = b3 ? t[2 + j] : t[j], give = b3 ? t[j] : t[2 + j];
    u[j] = mine + __shfl_xor_sync(0xffffffffu, give, 8);
  }
  float mine = b2 ? u[1] : u[0], give = b2 ? u[0] : u[1];
  float w = mine + __shfl_xor_sync(0xffffffffu, give, 4);
  w += __shfl_xor_sync(0xffffffffu, w, 2);
  w += __shfl_xor_sync(0xffffffffu, w, 1);
  return w;
}
__device__ __forceinline__ int col_of_lane_512(int warp, int lane) {
  return 8 * warp + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

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
  a.h0 = reinterpret_cast<__nv_bfloat16*>(d->h0);
  a.main_lat = reinterpret_cast<__nv_bfloat16*>(d->main_lat);
  a.hh[0] = reinterpret_cast<__nv_bfloat16*>(d->h_hidden[0]);
  a.hh[1] = reinterpret_cast<__nv_bfloat16*>(d->h_hidden[1]);
  a.agg = d->agg; a.agg_part = nullptr;
  a.g_out = reinterpret_cast<const __nv_bfloat16*>(d->g_out);
  a.g_agg = d->g_agg; a.g_main = d->g_main;
  a.g_h0 = reinterpret_cast<__nv_bfloat16*>(d->g_h0);
  a.w_part = nullptr;
  return a;
}

int umma_block_bwd2(const aero_block_desc* d, UmmaArgs a, int grid, cudaStream_t st);

}  // namespace aero
