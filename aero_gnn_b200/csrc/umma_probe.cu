// umma_probe.cu -- hardware probes for the tcgen05 data paths this library does not use yet (diagnostics only;
// nothing on the product path calls them).  They answer, on the actual part, the questions the next kernel design
// depends on:
//   mode 0: where does an M = 64 (cta_group::1) accumulator live in TMEM?  TMEM is zeroed by a full M = 128 MMA with
//           a zero A operand, then one M = 64, N = 128, K = 128 GEMM is issued at lane offset 0 and all 128 lanes x
//           128 columns are dumped.
//   mode 1: the same with the D address at lane offset 16 (can two M = 64 accumulators share one column range?)
//   mode 2: A operand from tensor memory (".ts" form): the bf16 A tile is written with tcgen05.st as 64 packed
//           32-bit columns per lane, the GEMM reads A from TMEM and B from shared memory; dump = A * B^T.
#include "umma_block.cuh"

namespace aero {

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// A operand in tensor memory
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __nv_bfloat16* __restrict__ Ag,
                                                            const __nv_bfloat16* __restrict__ Bg, float* __restrict__ C,
                                                            int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* At = smem;
  uint8_t* Bt = smem + TILE_BYTES;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(Bt + TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, wid = tid >> 5;
  // zero A tile, real B tile
  for (int i = tid; i < (int)(TILE_BYTES / 16); i += 128) reinterpret_cast<uint4*>(At)[i] = make_uint4(0u, 0u, 0u, 0u);
  stage_rows<false, 128>(Bt, Bg, nullptr, 0, 128, tid);
  if (tid == 0) {
    mbar_init(smem_u32(mbar), 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<256>(tmem_slot);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tacc = *tmem_slot;
  const uint32_t tlane = tacc + ((uint32_t)(wid * 32) << 16);
  uint32_t phase = 0;
  // 1) D[128 x 128] = 0 * B^T : every lane of the accumulator columns holds 0
  if (tid == 0) {
    issue_gemm(tacc, smem_u32(At), false, smem_u32(Bt), false, false);
    mma_commit(smem_u32(mbar));
  }
  mbar_wait(smem_u32(mbar), phase);
  phase ^= 1;
  fence_after_sync();
  __syncthreads();
  if (mode <= 1) {
    stage_rows<false, 128>(At, Ag, nullptr, 0, 64, tid);    // rows 0..63 of A, the rest stays zero
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t d = tacc + ((uint32_t)(mode == 1 ? 16 : 0) << 16);
      const uint32_t idesc = idesc_mn(64, 128);
      for (int kk = 0; kk < 8; ++kk)
        mma_bf16(d, desc_kmajor(smem_u32(At), kk), desc_kmajor(smem_u32(Bt), kk), idesc, kk > 0);
      mma_commit(smem_u32(mbar));
    }
  } else {
    // A into TMEM columns [128, 192): lane = row, 64 packed words = 128 bf16 of the row
    uint32_t w[32];
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(Ag + (size_t)tid * 128);
    for (int half = 0; half < 2; ++half) {
      for (int j = 0; j < 32; ++j) w[j] = arow[half * 32 + j];
      tmem_st32(tlane + 128u + (uint32_t)(half * 32), w);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc = idesc_mn(128, 128);
      for (int kk = 0; kk < 8; ++kk)   // K-step of 16 bf16 = 8 packed columns of A
        mma_bf16_ts(tacc, tacc + 128u + (uint32_t)(kk * 8), desc_kmajor(smem_u32(Bt), kk), idesc, kk > 0);
      mma_commit(smem_u32(mbar));
    }
  }
  mbar_wait(smem_u32(mbar), phase);
  fence_after_sync();
  for (int c = 0; c < 4; ++c) {
    float v[32];
    tmem_ld32(tlane + (uint32_t)(c * 32), v);
    for (int j = 0; j < 32; ++j) C[(size_t)tid * 128 + c * 32 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<256>(tacc);
}

// MMA throughput per operand orientation: `reps` back-to-back 128x128x128 GEMMs (8 MMAs each) between two clock
// reads taken by the issuing thread around issue .. commit-wait; out[0] = cycles, out[1] = MMAs issued.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(const __nv_bfloat16* __restrict__ Ag,
                                                           const __nv_bfloat16* __restrict__ Bg, long long* __restrict__ out,
                                                           int a_mn, int b_mn, int reps) {
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
  if (tid < 32) {
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) issue_gemm(tacc, smem_u32(At), a_mn != 0, smem_u32(Bt), b_mn != 0, r > 0);
      mma_commit(smem_u32(mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(mbar), 0);
    long long t1 = clock64();
    if (tid == 0) {
      out[0] = t1 - t0;
      out[1] = 8LL * reps;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<128>(tacc);
}

}  // namespace aero

using namespace aero;

extern "C" int aero_umma_rate_probe(const void* a_bf16, const void* b_bf16, long long* out2, int a_mn, int b_mn, int reps,
                                    void* stream) {
  AERO_CHECK_ARG(a_bf16 && b_bf16 && out2 && reps > 0, "aero_umma_rate_probe: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = 1024 + 2 * umma::TILE_BYTES + 64;
  AERO_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_rate_kernel<<<1, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(a_bf16),
                                         reinterpret_cast<const __nv_bfloat16*>(b_bf16), out2, a_mn, b_mn, reps);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" int aero_umma_probe(const void* a_bf16, const void* b_bf16, float* c, int mode, void* stream) {
  AERO_CHECK_ARG(a_bf16 && b_bf16 && c && mode >= 0 && mode <= 2, "aero_umma_probe: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = 1024 + 2 * umma::TILE_BYTES + 64;
  AERO_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(a_bf16),
                                          reinterpret_cast<const __nv_bfloat16*>(b_bf16), c, mode);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}
