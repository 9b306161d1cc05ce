// block_umma.cu -- fused MGN block on tcgen05 tensor cores (bf16 operands, fp32 TMEM accumulators).
// Placeholder until the tcgen05 kernels land: reports the path as unavailable (no silent fallback).
#include "common.cuh"

namespace aero {

size_t umma_prepared_bytes(int L) { (void)L; return 256; }
int umma_prepare(const float*, int, void*, cudaStream_t) {
  set_error("UMMA path not built into this library");
  return AERO_EUNSUPPORTED;
}
size_t umma_block_workspace_bytes(const aero_block_desc*, int) { return 256; }
int umma_block_fwd(const aero_block_desc*, cudaStream_t) {
  set_error("UMMA path not built into this library");
  return AERO_EUNSUPPORTED;
}
int umma_block_bwd(const aero_block_desc*, cudaStream_t) {
  set_error("UMMA path not built into this library");
  return AERO_EUNSUPPORTED;
}

}  // namespace aero

extern "C" int aero_has_umma(void) { return 0; }
