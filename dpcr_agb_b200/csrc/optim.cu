// optim.cu -- one-launch AdaBelief over a flat fp32 parameter buffer.
//
// Restates R:core/optimizer/adabelief.py:130-199 (decoupled weight decay, rectified update) with the
// glue R:models/base_model.py:241-246 fused in front: GradScaler unscale, clip_grad_value_, inf skip.
//
// hyper_host (float[16], host memory):
//   [0] lr  [1] beta1  [2] beta2  [3] eps  [4] weight_decay
//   [5] step_size  -- the rectified step size of adabelief.py:175-187 for THIS step (host computes it
//                     from the step count); <= 0 means "no parameter update" (:196-199)
//   [6] use_denom  -- 1 if num_sma >= 5 (adaptive update, :193-195), 0 for the SGD-like branch
//   [7] inv_scale  -- 1 / GradScaler scale    [8] clip -- clip_grad_value_ bound, <= 0 disables
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) grad_check_kernel(const float* __restrict__ g, int64_t n, float inv_scale,
                                                         float* __restrict__ found_inf) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bad |= !isfinite(g[i] * inv_scale);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.0f;
}

struct AbHyper {
  float lr, b1, b2, eps, wd, step_size, use_denom, inv_scale, clip;
};

__global__ void __launch_bounds__(256) adabelief_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ s, int64_t n,
                                                        AbHyper h, const float* __restrict__ hyper_dev,
                                                        const float* __restrict__ found_inf) {
  if (found_inf && *found_inf != 0.f) return;  // GradScaler.step skips the optimiser on inf/nan
  if (hyper_dev)                               // hyper-parameters live on the device (captured-graph replays)
    h = AbHyper{hyper_dev[0], hyper_dev[1], hyper_dev[2], hyper_dev[3], hyper_dev[4],
                hyper_dev[5], hyper_dev[6], hyper_dev[7], hyper_dev[8]};
  const float decay = 1.0f - h.lr * h.wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * h.inv_scale;
    if (h.clip > 0.f) gi = fminf(fmaxf(gi, -h.clip), h.clip);
    float pi = p[i] * decay;                                  // adabelief.py:131-135
    float mi = m[i] * h.b1 + gi * (1.0f - h.b1);             // :147
    const float r = gi - mi;                                  // :148
    float si = s[i] * h.b2 + r * r * (1.0f - h.b2) + h.eps;  // :149, :159 (eps is added in place)
    if (h.use_denom != 0.f)
      pi -= h.step_size * h.lr * mi / (sqrtf(si) + h.eps);    // :193-195
    else if (h.step_size > 0.f)
      pi -= h.step_size * h.lr * mi;                          // :196-197
    p[i] = pi;
    m[i] = mi;
    s[i] = si;
  }
}

}  // namespace

extern "C" int32_t b2s_grad_check(const float* grad, int64_t numel, float inv_scale, float* found_inf_dev,
                                  b2s_stream_t stream) {
  B2S_CHECK_ARG(numel >= 0 && found_inf_dev, "bad arguments");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(found_inf_dev, 0, sizeof(float), st));
  if (numel == 0) return B2S_OK;
  B2S_CHECK_ARG(grad, "null pointer");
  grad_check_kernel<<<grid_for(numel, 256), 256, 0, st>>>(grad, numel, inv_scale, found_inf_dev);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_adabelief_step(float* param, const float* grad, float* exp_avg, float* exp_avg_var,
                                      int64_t numel, const float* hyper_host, const float* hyper_dev,
                                      const float* found_inf_dev, b2s_stream_t stream) {
  B2S_CHECK_ARG(numel >= 0 && ((hyper_host != nullptr) != (hyper_dev != nullptr)),
                "exactly one of hyper_host / hyper_dev must be given");
  if (numel == 0) return B2S_OK;
  B2S_CHECK_ARG(param && grad && exp_avg && exp_avg_var, "null pointer");
  AbHyper h{};
  if (hyper_host)
    h = AbHyper{hyper_host[0], hyper_host[1], hyper_host[2], hyper_host[3], hyper_host[4],
                hyper_host[5], hyper_host[6], hyper_host[7], hyper_host[8]};
  adabelief_kernel<<<grid_for(numel, 256), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_var, numel, h,
                                                                       hyper_dev, found_inf_dev);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}
