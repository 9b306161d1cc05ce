// umma.cuh -- tcgen05 / TMEM / mbarrier primitives for sm_100a, written as inline PTX.
//
// Shared-memory tile format used by every operand in this library ("row tile"):
//   a [128 rows][128 features] bf16 tile is stored as two 16 KB panels (features 0..63, 64..127);
//   inside a panel a row is 128 B, rows are consecutive, and the eight 16-byte chunks of a row are
//   XOR-swizzled with (row & 7) -- exactly the canonical SWIZZLE_128B layout of the UMMA descriptors.
//   The same physical tile serves as
//     * a K-major operand   (rows = M or N index, features = K):  SBO = 1024, K-step 16 = +32 B, panel switch at K=64
//     * an MN-major operand (features = M or N index, rows = K):  LBO = 16384 (panel stride), SBO = 1024, K-step 16 rows = +2048 B
//   so activations, weights and gradients are staged once and read in whichever orientation a GEMM needs
//   (forward, data-gradient, weight-gradient).  Tiles must be 1024-byte aligned.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aero {
namespace umma {

constexpr uint32_t TILE_BYTES = 32768;
constexpr uint32_t PANEL_BYTES = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of the 16-byte chunk holding features [8*chunk, 8*chunk+8) of `row` inside a row tile
__device__ __forceinline__ uint32_t tile_chunk_off(int row, int chunk /*0..15*/) {
  return (uint32_t)((chunk >> 3) * PANEL_BYTES + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

// ---- descriptors -------------------------------------------------------------------------------
// SWIZZLE_128B shared-memory matrix descriptor (version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// K-major operand, K-step kk (16 elements each) of a 128-wide tile
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, int kk) {
  return make_desc(tile_saddr + (uint32_t)(kk >> 2) * PANEL_BYTES + (uint32_t)(kk & 3) * 32, 16, 1024);
}
// MN-major operand, K-step kk = 16 rows
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, int kk) {
  return make_desc(tile_saddr + (uint32_t)kk * 2048, PANEL_BYTES, 1024);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = 128
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                  // D format fp32
         | (1u << 7) | (1u << 10)   // A, B format bf16
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((128u >> 3) << 17)      // N
         | ((128u >> 4) << 24);     // M
}

// ---- tcgen05 -------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// make completion of all prior tcgen05.mma of this thread arrive on an mbarrier
__device__ __forceinline__ void mma_commit(uint32_t mbar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar_saddr) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t saddr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t saddr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(saddr), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait with back-off: a protocol bug traps (fails the launch) instead of hanging the GPU, and waiting
// warps do not steal issue slots from the warps doing epilogue work
__device__ __forceinline__ void mbar_wait(uint32_t saddr, uint32_t parity) {
  if (mbar_try_wait(saddr, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(saddr, parity)) {   // try_wait suspends the warp in hardware until the phase flips or a time limit
#ifdef AERO_WAIT_NANOSLEEP
    __nanosleep(AERO_WAIT_NANOSLEEP);
#endif
    if (++spins > (1u << 26)) asm volatile("trap;\n");
  }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t saddr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(saddr), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t mbar_saddr) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_saddr),
               "l"(src), "r"(bytes), "r"(mbar_saddr)
               : "memory");
}

// one lane of a fully converged warp; the compiler keeps code under this predicate on the uniform datapath, so the
// tcgen05.mma operands are built in uniform registers (a `threadIdx.x == 0` branch makes it fall back to a
// vector-register + R2UR "waterfall" of ~20 dependent instructions per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}

// named barrier for one 128-thread warpgroup
__device__ __forceinline__ void wg_sync(int id) { asm volatile("bar.sync %0, 128;\n" ::"r"(id) : "memory"); }

// ---- packing -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// max(x, 0) fused into the conversion (one instruction for two activations)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// packed sum of two bf16 pairs, each lane rounded once (same result as an fp32 add followed by cvt.rn)
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.bf16x2 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

}  // namespace umma
}  // namespace aero
