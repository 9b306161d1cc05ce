// abi.cu -- extern "C" entry points of libaero_sm100.so that dispatch to a kernel family.
#include <stdarg.h>
#include "common.cuh"

namespace aero {

static thread_local char g_err[512] = "";
thread_local int g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static int validate_block(const aero_block_desc* d, int backward) {
  AERO_CHECK_ARG(d != nullptr, "aero_block: null descriptor");
  if (d->dtype != AERO_F32 && d->dtype != AERO_BF16) {
    set_error("aero_block: unsupported dtype %d (fp32 and bf16 only)", d->dtype);
    return AERO_EUNSUPPORTED;
  }
  if (d->act < AERO_ACT_RELU || d->act > AERO_ACT_LEAKY_RELU) {
    set_error("aero_block: unsupported activation %d", d->act);
    return AERO_EUNSUPPORTED;
  }
  AERO_CHECK_ARG(d->L >= 0 && d->rows >= 0 && d->n_nodes >= 0, "aero_block: negative size");
  AERO_CHECK_ARG(d->rows < 2147483647LL && d->n_nodes < 2147483647LL, "aero_block: sizes exceed int32");
  AERO_CHECK_ARG(d->prepared != nullptr, "aero_block: prepared weights missing");
  if (d->rows > 0) {
    AERO_CHECK_ARG(d->main && (d->P || (backward && d->h0)), "aero_block: null main/P");
    if (d->h0 && d->path != AERO_PATH_UMMA) {
      set_error("aero_block: the h0 buffer is supported by AERO_PATH_UMMA only");
      return AERO_EUNSUPPORTED;
    }
    AERO_CHECK_ARG((d->ldp % 8) == 0 && (d->poff0 % 8) == 0 && (d->poff1 % 8) == 0, "aero_block: P strides must be multiples of 8");
    if (!backward) AERO_CHECK_ARG(d->out, "aero_block_fwd: null out");  /* resid == NULL: no residual */
    if (!backward && d->agg) AERO_CHECK_ARG(d->rowptr && d->idx1, "aero_block_fwd: agg needs rowptr and idx1 (= receiver)");
    if (backward) AERO_CHECK_ARG(d->g_out && d->g_main && d->g_h0 && d->g_w, "aero_block_bwd: null gradient buffer");
    if (backward && d->g_agg) AERO_CHECK_ARG(d->idx1, "aero_block_bwd: g_agg needs idx1 (= receiver)");
  }
  size_t need = aero_block_workspace_bytes(d, backward);
  if (need > 0 && (d->workspace == nullptr || d->workspace_bytes < need)) {
    set_error("aero_block: workspace %zu < %zu", d->workspace_bytes, need);
    return AERO_EWORKSPACE;
  }
  return AERO_OK;
}

}  // namespace aero

using namespace aero;

extern "C" const char* aero_last_error(void) { return g_err; }
extern "C" int aero_version(void) { return 100; }
extern "C" int aero_last_launch_count(void) { return g_launch_count; }

extern "C" size_t aero_block_prepared_bytes(int L, int path) {
  return path == AERO_PATH_UMMA ? umma_prepared_bytes(L) : simt_prepared_bytes(L);
}

extern "C" int aero_block_prepare(const float* w, int L, int path, void* prepared, void* stream) {
  AERO_CHECK_ARG(w && prepared && L >= 0, "aero_block_prepare: bad arguments");
  g_launch_count = 0;
  if (path == AERO_PATH_UMMA) return umma_prepare(w, L, prepared, (cudaStream_t)stream);
  if (path == AERO_PATH_SIMT) return simt_prepare(w, L, prepared, (cudaStream_t)stream);
  set_error("aero_block_prepare: unknown path %d", path);
  return AERO_EINVAL;
}

extern "C" size_t aero_block_workspace_bytes(const aero_block_desc* d, int backward) {
  if (!d) return 0;
  return d->path == AERO_PATH_UMMA ? umma_block_workspace_bytes(d, backward) : simt_block_workspace_bytes(d, backward);
}

extern "C" int aero_block_fwd(const aero_block_desc* d, void* stream) {
  g_launch_count = 0;
  int rc = validate_block(d, 0);
  if (rc) return rc;
  if (d->path == AERO_PATH_UMMA) return umma_block_fwd(d, (cudaStream_t)stream);
  return simt_block_fwd(d, (cudaStream_t)stream);
}

extern "C" int aero_block_bwd(const aero_block_desc* d, void* stream) {
  g_launch_count = 0;
  int rc = validate_block(d, 1);
  if (rc) return rc;
  if (d->path == AERO_PATH_UMMA) return umma_block_bwd(d, (cudaStream_t)stream);
  return simt_block_bwd(d, (cudaStream_t)stream);
}
