// train_tail.cu -- the tail of a training step on the device: MSE loss + its gradient, multi-tensor Adam.
//
// Reference: utils.py:191-195 (loss = MSELoss(pred, y); loss.backward(); optimizer.step(); optimizer.zero_grad();
// loss.item() every batch) and train.py:207-211 (torch.optim.Adam(lr, weight_decay)).  After the processor is fused a
// small-mesh step is bound by launches and host syncs, so here
//   * the loss and dL/dpred come out of ONE pass over the predictions (+ one tiny fixed-order reduction), the loss
//     stays on the device (no per-batch .item());
//   * every parameter tensor of the model is updated by ONE launch driven by a device-resident segment table
//     (torch.optim.Adam semantics: L2 weight decay added to the gradient, bias-corrected moments), bf16 parameters
//     optionally backed by fp32 master copies.
#include <math.h>
#include "common.cuh"

namespace aero {

constexpr int MSE_BLOCKS = 256;

template <typename T>
__global__ void __launch_bounds__(256) mse_partial_kernel(const T* __restrict__ pred, const float* __restrict__ tgt,
                                                          T* __restrict__ grad, float* __restrict__ part, int64_t rows,
                                                          int cols, int64_t ld_pred, int64_t ld_grad, float gscale) {
  const int64_t n = rows * cols;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float d = load1(pred + r * ld_pred + c) - tgt[i];
    acc = fmaf(d, d, acc);
    store1(grad + r * ld_grad + c, gscale * d);
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {   // fixed tree: deterministic
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) mse_final_kernel(const float* __restrict__ part, int n_parts, float lscale,
                                                        float* __restrict__ loss) {
  __shared__ float red[256];
  red[threadIdx.x] = threadIdx.x < n_parts ? part[threadIdx.x] : 0.f;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] * lscale;
}

template <typename P, typename G>
__device__ __forceinline__ void adam_elems(const aero_adam_seg& s, float lr_c, float b1, float b2, float eps, float wd,
                                           float inv_sqrt_c2) {
  P* p = reinterpret_cast<P*>(s.param);
  const G* g = reinterpret_cast<const G*>(s.grad);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float w = s.master ? s.master[i] : load1(p + i);
    float gi = load1(g + i);
    gi = fmaf(wd, w, gi);                                   // torch.optim.Adam: L2 penalty folded into the gradient
    const float m = fmaf(b1, s.m[i], (1.f - b1) * gi);
    const float v = fmaf(b2, s.v[i], (1.f - b2) * gi * gi);
    s.m[i] = m;
    s.v[i] = v;
    const float wn = w - lr_c * m / (sqrtf(v) * inv_sqrt_c2 + eps);
    if (s.master) s.master[i] = wn;
    store1(p + i, wn);
  }
}

// Per-tensor step counts (torch.optim.Adam keeps one per parameter: a parameter without a gradient skips the step
// and its bias corrections lag).  steps_in is read by every block of a segment, steps_out written by one thread, the
// caller swaps the two arrays between launches -- no race, no second launch.
__global__ void __launch_bounds__(256) adam_kernel(const aero_adam_seg* __restrict__ segs, const int32_t* __restrict__ steps_in,
                                                   int32_t* __restrict__ steps_out, float lr, float b1, float b2, float eps,
                                                   float wd) {
  const aero_adam_seg s = segs[blockIdx.y];
  const int t_prev = steps_in[blockIdx.y];
  const bool live = s.grad != nullptr && s.n > 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) steps_out[blockIdx.y] = t_prev + (live ? 1 : 0);
  if (!live) return;
  // bias corrections in double, as torch computes them in Python floats
  const double c1 = 1.0 - pow((double)b1, (double)(t_prev + 1)), c2 = 1.0 - pow((double)b2, (double)(t_prev + 1));
  const float lr_c = (float)((double)lr / c1), inv_sqrt_c2 = (float)(1.0 / sqrt(c2));
  if (s.p_dtype == AERO_F32) {
    if (s.g_dtype == AERO_F32) adam_elems<float, float>(s, lr_c, b1, b2, eps, wd, inv_sqrt_c2);
    else adam_elems<float, __nv_bfloat16>(s, lr_c, b1, b2, eps, wd, inv_sqrt_c2);
  } else {
    if (s.g_dtype == AERO_F32) adam_elems<__nv_bfloat16, float>(s, lr_c, b1, b2, eps, wd, inv_sqrt_c2);
    else adam_elems<__nv_bfloat16, __nv_bfloat16>(s, lr_c, b1, b2, eps, wd, inv_sqrt_c2);
  }
}

}  // namespace aero

using namespace aero;

extern "C" size_t aero_mse_workspace_bytes(void) { return MSE_BLOCKS * sizeof(float); }

extern "C" int aero_mse_loss_grad(const void* pred, const float* target, void* grad, float* loss, int64_t rows, int64_t cols,
                                  int64_t ld_pred, int64_t ld_grad, int dtype, float loss_scale, float grad_scale,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  g_launch_count = 0;
  AERO_CHECK_ARG(rows >= 0 && cols > 0 && cols < (1 << 20) && ld_pred >= cols && ld_grad >= cols, "aero_mse_loss_grad: bad sizes");
  AERO_CHECK_ARG(loss && workspace && workspace_bytes >= aero_mse_workspace_bytes(), "aero_mse_loss_grad: workspace");
  AERO_CHECK_ARG(rows == 0 || (pred && target && grad), "aero_mse_loss_grad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(workspace);
  const int64_t n = rows * cols;
  int blocks = (int)(cdiv(n > 0 ? n : 1, 256 * 8) < MSE_BLOCKS ? cdiv(n > 0 ? n : 1, 256 * 8) : MSE_BLOCKS);
  if (dtype == AERO_F32)
    mse_partial_kernel<float><<<blocks, 256, 0, st>>>((const float*)pred, target, (float*)grad, part, rows, (int)cols, ld_pred,
                                                      ld_grad, grad_scale);
  else if (dtype == AERO_BF16)
    mse_partial_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)pred, target, (__nv_bfloat16*)grad, part,
                                                              rows, (int)cols, ld_pred, ld_grad, grad_scale);
  else {
    set_error("aero_mse_loss_grad: unsupported dtype %d", dtype);
    return AERO_EUNSUPPORTED;
  }
  AERO_LAUNCH_CHECK();
  mse_final_kernel<<<1, 256, 0, st>>>(part, blocks, loss_scale, loss);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}

extern "C" int aero_adam_step(const aero_adam_seg* segs_device, int n_segs, int64_t max_elems, const int32_t* steps_in,
                              int32_t* steps_out, float lr, float beta1, float beta2, float eps, float weight_decay,
                              void* stream) {
  g_launch_count = 0;
  AERO_CHECK_ARG(n_segs >= 0 && n_segs <= 65535 && (n_segs == 0 || (segs_device && steps_in && steps_out)) &&
                     steps_in != steps_out && max_elems >= 0,
                 "aero_adam_step: bad arguments");
  if (n_segs == 0) return AERO_OK;
  int64_t bx = cdiv(max_elems > 0 ? max_elems : 1, 256 * 4);
  if (bx > 32) bx = 32;
  dim3 grid((unsigned)bx, (unsigned)n_segs);
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(segs_device, steps_in, steps_out, lr, beta1, beta2, eps, weight_decay);
  AERO_LAUNCH_CHECK();
  return AERO_OK;
}
