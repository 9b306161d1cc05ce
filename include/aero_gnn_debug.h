/*
 * aero_gnn_debug.h -- hardware probes of libaero_probe.so (built from csrc/umma_probe.cu).
 *
 * Diagnostics used while designing the kernels: they are NOT part of the product boundary
 * (include/aero_gnn.h / libaero_sm100.so) and nothing on the message-passing path calls them.
 */
#ifndef AERO_GNN_DEBUG_H_
#define AERO_GNN_DEBUG_H_

#ifdef __cplusplus
extern "C" {
#endif

/* tcgen05 hardware probes (diagnostics for kernel design; not on the product path).
 * c[128][128] receives all 128 TMEM lanes x 128 accumulator columns after
 *   mode 0: TMEM zeroed, then one M = 64 GEMM (rows 0..63 of a) at lane offset 0
 *   mode 1: the same with the accumulator address at lane offset 16
 *   mode 2: c = a * b^T with the A operand read from tensor memory (written there with tcgen05.st) */
int aero_umma_probe(const void* a_bf16, const void* b_bf16, float* c, int mode, void* stream);
/* cycles for `reps` back-to-back 128x128x128 GEMMs with the given operand orientations (0 = K-major, 1 = MN-major):
 * out2[0] = SM cycles from first issue to completion, out2[1] = number of tcgen05.mma issued (device int64[2]) */
int aero_umma_rate_probe(const void* a_bf16, const void* b_bf16, long long* out2, int a_mn, int b_mn, int reps,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AERO_GNN_DEBUG_H_ */
