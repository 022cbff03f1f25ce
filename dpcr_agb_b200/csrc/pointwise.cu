// pointwise.cu -- the bandwidth-bound stages: max pooling (a9), per-plot reductions and broadcasts
// (a5, a10, a11), batch-norm statistics / apply (a13), GELU (a14), column sums (bias gradient).
// Every kernel streams fp32 [N, C] rows with coalesced 128-byte warp accesses and is bounded by HBM.
//
// Reference call sites (R: = /root/reference/torch-points3d/torch_points3d/modules/MinkowskiEngine/):
//   a9 R:SENet.py:53   a10/a11 R:senet_block.py:43-50, R:common.py:44-48, R:SENet.py:63,117
//   a13 R:SENet.py:35,51,98 + R:resnet_block.py:51-55   a14 R:common.py:41
#include "common.cuh"
#include <math.h>

namespace {

constexpr int PW_THREADS = 256;

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// ---------------------------------------------------------------- max pooling ---------------
__global__ void __launch_bounds__(PW_THREADS) maxpool_fwd_kernel(const float* __restrict__ x,
                                                                 const int* __restrict__ nbr, int64_t n_out, int c,
                                                                 int k3, float* __restrict__ y,
                                                                 int* __restrict__ arg) {
  const int64_t total = n_out * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = e / c;
    const int ch = (int)(e - o * c);
    float best = -INFINITY;
    int bi = -1;
    for (int k = 0; k < k3; ++k) {
      const int i = __ldg(&nbr[(int64_t)k * n_out + o]);
      if (i < 0) continue;
      const float v = __ldg(&x[(int64_t)i * c + ch]);
      if (bi < 0 || v > best || (v == best && i < bi)) {
        best = v;
        bi = i;
      }
    }
    y[e] = bi >= 0 ? best : 0.f;
    arg[e] = bi;
  }
}

__global__ void __launch_bounds__(PW_THREADS) maxpool_bwd_kernel(const float* __restrict__ gy,
                                                                 const int* __restrict__ arg, int64_t n_out, int c,
                                                                 float* __restrict__ gx) {
  const int64_t total = n_out * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int a = arg[e];
    if (a >= 0) atomicAdd(&gx[(int64_t)a * c + (int)(e % c)], gy[e]);
  }
}

// ---------------------------------------------------------------- per-plot segments ---------
__global__ void __launch_bounds__(PW_THREADS) batch_counts_kernel(const int* __restrict__ rb, int stride, int64_t n,
                                                                  int nb, int* __restrict__ counts) {
  constexpr int CHUNK = 64;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t r0 = t * CHUNK, r1 = min(r0 + CHUNK, n);
  int cur = -1, run = 0;
  for (int64_t r = r0; r < r1; ++r) {
    int b = rb[r * stride];
    if (b != cur) {
      if (run && (unsigned)cur < (unsigned)nb) atomicAdd(&counts[cur], run);
      cur = b;
      run = 0;
    }
    ++run;
  }
  if (run && (unsigned)cur < (unsigned)nb) atomicAdd(&counts[cur], run);
}

// Column-wise reduction skeleton: block = 32 channel lanes x 8 row lanes; blockIdx.y tiles channels by 32,
// blockIdx.x tiles rows by ROWS_PER_CTA.  A warp reads 32 consecutive floats of one row (128 B).
constexpr int ROWS_PER_CTA = 256;

// y[b, ch] += sum over rows of x (optionally times x2), segmented by batch id; runs of equal batch id are
// accumulated in registers and flushed with one atomicAdd.
template <bool MUL>
__global__ void __launch_bounds__(PW_THREADS) segment_sum_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ x2,
                                                                 const int* __restrict__ rb, int stride, int64_t n,
                                                                 int c, int nb, float* __restrict__ y) {
  const int ch = blockIdx.y * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_CTA;
  const int64_t r1 = min(r0 + ROWS_PER_CTA, n);
  if (ch >= c) return;
  int cur = -1;
  float acc = 0.f;
  for (int64_t r = r0 + ty; r < r1; r += 8) {
    const int b = __ldg(&rb[r * stride]);
    if (b != cur) {
      if (cur >= 0 && cur < nb) atomicAdd(&y[(int64_t)cur * c + ch], acc);
      cur = b;
      acc = 0.f;
    }
    float v = x[r * c + ch];
    if (MUL) v *= x2[r * c + ch];
    acc += v;
  }
  if (cur >= 0 && cur < nb) atomicAdd(&y[(int64_t)cur * c + ch], acc);
}

__global__ void __launch_bounds__(PW_THREADS) scale_rows_kernel(float* __restrict__ y, const float* __restrict__ scale,
                                                                int nb, int c) {
  const int total = nb * c;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) y[e] *= scale[e / c];
}

__global__ void __launch_bounds__(PW_THREADS) segment_bcast_kernel(const float* __restrict__ y,
                                                                   const int* __restrict__ rb, int stride, int64_t n,
                                                                   int c, const float* __restrict__ scale,
                                                                   float* __restrict__ out) {
  const int64_t total = n * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / c;
    const int ch = (int)(e - r * c);
    const int b = __ldg(&rb[r * stride]);
    float v = __ldg(&y[(int64_t)b * c + ch]);
    if (scale) v *= __ldg(&scale[b]);
    out[e] = v;
  }
}

__global__ void __launch_bounds__(PW_THREADS) bcast_mul_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ y,
                                                               const int* __restrict__ rb, int stride, int64_t n,
                                                               int c, int y_c, float* __restrict__ out) {
  const int64_t total = n * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / c;
    const int ch = (int)(e - r * c);
    const int b = __ldg(&rb[r * stride]);
    out[e] = x[e] * __ldg(&y[(int64_t)b * y_c + (y_c == 1 ? 0 : ch)]);
  }
}

// ---------------------------------------------------------------- column reductions ---------
// MODE 0: sum x            -> ws[ch]
// MODE 1: sum x, sum x^2   -> ws[ch], ws[c+ch]           (batch-norm statistics)
// MODE 2: sum g', sum g'*xhat  with g' = g * act'(pre)   (batch-norm backward)
template <int MODE, typename ACC>
__global__ void __launch_bounds__(PW_THREADS) colreduce_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ g,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ invstd,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int64_t n, int c,
                                                               int act, ACC* __restrict__ ws) {
  const int ch = blockIdx.y * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_CTA;
  const int64_t r1 = min(r0 + ROWS_PER_CTA, n);
  __shared__ float sm[2][8][33];
  float a0 = 0.f, a1 = 0.f;
  if (ch < c) {
    float mu = 0.f, is = 1.f, ga = 1.f, be = 0.f;
    if (MODE == 2) {
      mu = mean[ch];
      is = invstd[ch];
      ga = gamma ? gamma[ch] : 1.f;
      be = beta ? beta[ch] : 0.f;
    }
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float v = x[r * c + ch];
      if (MODE == 0) {
        a0 += v;
      } else if (MODE == 1) {
        a0 += v;
        a1 += v * v;
      } else {
        const float xh = (v - mu) * is;
        float gg = g[r * c + ch];
        if (act == 1) gg *= gelu_grad_f(xh * ga + be);
        a0 += gg;
        a1 += gg * xh;
      }
    }
  }
  sm[0][ty][threadIdx.x & 31] = a0;
  sm[1][ty][threadIdx.x & 31] = a1;
  __syncthreads();
  if (ty == 0 && ch < c) {
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      a0 += sm[0][j][threadIdx.x];
      a1 += sm[1][j][threadIdx.x];
    }
    atomicAdd(&ws[ch], (ACC)a0);
    if (MODE != 0) atomicAdd(&ws[c + ch], (ACC)a1);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ ws, int64_t n, int c, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean, float* __restrict__ invstd) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double m = ws[ch] / (double)n;
  double var = ws[c + ch] / (double)n - m * m;
  if (var < 0.0) var = 0.0;
  mean[ch] = (float)m;
  invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
  if (running_var) {
    const double unb = n > 1 ? var * (double)n / (double)(n - 1) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
  }
}

__global__ void double_to_float_kernel(const double* __restrict__ in, float* __restrict__ out, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = (float)in[i];
}

__global__ void __launch_bounds__(PW_THREADS) bn_apply_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ invstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int64_t n, int c,
                                                              int act, float* __restrict__ y) {
  const int64_t total = n * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c);
    float v = (x[e] - __ldg(&mean[ch])) * __ldg(&invstd[ch]);
    v = v * (gamma ? __ldg(&gamma[ch]) : 1.f) + (beta ? __ldg(&beta[ch]) : 0.f);
    y[e] = act == 1 ? gelu_f(v) : v;
  }
}

__global__ void __launch_bounds__(PW_THREADS) bn_bwd_apply_kernel(
    const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ sums, int64_t n, int c, int act, int training, float* __restrict__ gx) {
  const int64_t total = n * c;
  const float inv_n = 1.f / (float)n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c);
    const float is = __ldg(&invstd[ch]);
    const float ga = gamma ? __ldg(&gamma[ch]) : 1.f;
    const float xh = (x[e] - __ldg(&mean[ch])) * is;
    float g = gy[e];
    if (act == 1) g *= gelu_grad_f(xh * ga + (beta ? __ldg(&beta[ch]) : 0.f));
    if (training) g = g - __ldg(&sums[ch]) * inv_n - xh * __ldg(&sums[c + ch]) * inv_n;
    gx[e] = g * ga * is;
  }
}

__global__ void __launch_bounds__(PW_THREADS) gelu_fwd_kernel(const float* __restrict__ x, int64_t numel,
                                                              float* __restrict__ y) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = gelu_f(x[e]);
}

__global__ void __launch_bounds__(PW_THREADS) gelu_bwd_kernel(const float* __restrict__ gy,
                                                              const float* __restrict__ x, int64_t numel,
                                                              float* __restrict__ gx) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel; e += (int64_t)gridDim.x * blockDim.x)
    gx[e] = gy[e] * gelu_grad_f(x[e]);
}

// s = a + b (kept for the backward), y = gelu(s): the residual join of every block (senet_block.py:93-94)
__global__ void __launch_bounds__(PW_THREADS) add_gelu_fwd_kernel(const float4* __restrict__ a,
                                                                  const float4* __restrict__ b, int64_t n4,
                                                                  float4* __restrict__ s, float4* __restrict__ y) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    const float4 av = a[e], bv = b[e];
    const float4 sv = make_float4(av.x + bv.x, av.y + bv.y, av.z + bv.z, av.w + bv.w);
    s[e] = sv;
    y[e] = make_float4(gelu_f(sv.x), gelu_f(sv.y), gelu_f(sv.z), gelu_f(sv.w));
  }
}

dim3 colgrid(int64_t n, int c) { return dim3((unsigned)ceil_div64(n, ROWS_PER_CTA), (unsigned)((c + 31) / 32)); }

}  // namespace

// ================================================================= C ABI ======================
extern "C" int32_t b2s_maxpool_fwd(const float* x, const int32_t* nbr, int64_t n_out, int32_t c, int32_t k3, float* y,
                                   int32_t* arg, b2s_stream_t stream) {
  B2S_CHECK_ARG(n_out >= 0 && c > 0 && k3 > 0, "bad sizes");
  if (n_out == 0) return B2S_OK;
  B2S_CHECK_ARG(x && nbr && y && arg, "null pointer");
  maxpool_fwd_kernel<<<grid_for(n_out * c, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(x, nbr, n_out, c, k3, y,
                                                                                           arg);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_maxpool_bwd(const float* gy, const int32_t* arg, int64_t n_in, int64_t n_out, int32_t c,
                                   float* gx, b2s_stream_t stream) {
  B2S_CHECK_ARG(n_in >= 0 && n_out >= 0 && c > 0, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (n_in > 0) {
    B2S_CHECK_ARG(gx, "null pointer");
    B2S_CUDA(cudaMemsetAsync(gx, 0, n_in * c * sizeof(float), st));
  }
  if (n_out == 0) return B2S_OK;
  B2S_CHECK_ARG(gy && arg, "null pointer");
  maxpool_bwd_kernel<<<grid_for(n_out * c, PW_THREADS), PW_THREADS, 0, st>>>(gy, arg, n_out, c, gx);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_batch_counts(const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                    int32_t num_batches, int32_t* counts, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && num_batches > 0 && counts && row_batch_stride > 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(counts, 0, num_batches * sizeof(int), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(row_batch, "null pointer");
  batch_counts_kernel<<<(unsigned)ceil_div64(ceil_div64(n, 64), PW_THREADS), PW_THREADS, 0, st>>>(
      row_batch, row_batch_stride, n, num_batches, counts);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_segment_sum(const float* x, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                   int32_t c, int32_t num_batches, const float* scale, float* y,
                                   b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && num_batches > 0 && y && row_batch_stride > 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(y, 0, (size_t)num_batches * c * sizeof(float), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && row_batch, "null pointer");
  segment_sum_kernel<false><<<colgrid(n, c), PW_THREADS, 0, st>>>(x, nullptr, row_batch, row_batch_stride, n, c,
                                                                  num_batches, y);
  if (scale) scale_rows_kernel<<<grid_for((int64_t)num_batches * c, PW_THREADS), PW_THREADS, 0, st>>>(y, scale, num_batches, c);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_segment_bcast(const float* y, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                     int32_t c, const float* scale, float* x_out, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && row_batch_stride > 0, "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(y && row_batch && x_out, "null pointer");
  segment_bcast_kernel<<<grid_for(n * c, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(y, row_batch,
                                                                                         row_batch_stride, n, c,
                                                                                         scale, x_out);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bcast_mul_fwd(const float* x, const float* y, const int32_t* row_batch,
                                     int32_t row_batch_stride, int64_t n, int32_t c, int32_t y_c, float* out,
                                     b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && (y_c == c || y_c == 1) && row_batch_stride > 0, "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && y && row_batch && out, "null pointer");
  bcast_mul_kernel<<<grid_for(n * c, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(x, y, row_batch,
                                                                                     row_batch_stride, n, c, y_c,
                                                                                     out);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bcast_mul_bwd(const float* g, const float* x, const float* y, const int32_t* row_batch,
                                     int32_t row_batch_stride, int64_t n, int32_t c, int32_t num_batches, float* gx,
                                     float* gy, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && num_batches > 0 && row_batch_stride > 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (gy) B2S_CUDA(cudaMemsetAsync(gy, 0, (size_t)num_batches * c * sizeof(float), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(g && row_batch, "null pointer");
  if (gx) {
    B2S_CHECK_ARG(y, "null pointer");
    bcast_mul_kernel<<<grid_for(n * c, PW_THREADS), PW_THREADS, 0, st>>>(g, y, row_batch, row_batch_stride, n, c, c,
                                                                         gx);
  }
  if (gy) {
    B2S_CHECK_ARG(x, "null pointer");
    segment_sum_kernel<true><<<colgrid(n, c), PW_THREADS, 0, st>>>(g, x, row_batch, row_batch_stride, n, c,
                                                                   num_batches, gy);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_colsum(const float* x, int64_t n, int32_t c, float* out, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && out, "bad arguments");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(out, 0, c * sizeof(float), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x, "null pointer");
  colreduce_kernel<0, float><<<colgrid(n, c), PW_THREADS, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, nullptr, n,
                                                                   c, 0, out);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bn_stats(const float* x, int64_t n, int32_t c, float eps, float momentum, float* running_mean,
                                float* running_var, double* stats_ws, float* mean, float* invstd,
                                b2s_stream_t stream) {
  B2S_CHECK_ARG(n > 0 && c > 0, "n > 0 and c > 0");
  B2S_CHECK_ARG(x && stats_ws && mean && invstd, "null pointer");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(stats_ws, 0, 2 * (size_t)c * sizeof(double), st));
  colreduce_kernel<1, double><<<colgrid(n, c), PW_THREADS, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                    n, c, 0, stats_ws);
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(stats_ws, n, c, eps, momentum, running_mean, running_var, mean,
                                                      invstd);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bn_apply(const float* x, const float* mean, const float* invstd, const float* gamma,
                                const float* beta, int64_t n, int32_t c, int32_t act, float* y,
                                b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && (act == 0 || act == 1), "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && mean && invstd && y, "null pointer");
  bn_apply_kernel<<<grid_for(n * c, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(x, mean, invstd, gamma, beta, n,
                                                                                    c, act, y);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bn_bwd_reduce(const float* gy, const float* x, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, int64_t n, int32_t c, int32_t act,
                                     double* stats_ws, float* sums, b2s_stream_t stream) {
  B2S_CHECK_ARG(n > 0 && c > 0 && (act == 0 || act == 1), "bad arguments");
  B2S_CHECK_ARG(gy && x && mean && invstd && stats_ws && sums, "null pointer");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(stats_ws, 0, 2 * (size_t)c * sizeof(double), st));
  colreduce_kernel<2, double><<<colgrid(n, c), PW_THREADS, 0, st>>>(x, gy, mean, invstd, gamma, beta, n, c, act,
                                                                    stats_ws);
  double_to_float_kernel<<<(2 * c + 127) / 128, 128, 0, st>>>(stats_ws, sums, 2 * c);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bn_bwd_apply(const float* gy, const float* x, const float* mean, const float* invstd,
                                    const float* gamma, const float* beta, const float* sums, int64_t n, int32_t c,
                                    int32_t act, int32_t training, float* gx, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && (act == 0 || act == 1), "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(gy && x && mean && invstd && gx && (sums || !training), "null pointer");
  bn_bwd_apply_kernel<<<grid_for(n * c, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(
      gy, x, mean, invstd, gamma, beta, sums, n, c, act, training, gx);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_gelu_fwd(const float* x, int64_t numel, float* y, b2s_stream_t stream) {
  B2S_CHECK_ARG(numel >= 0, "numel >= 0");
  if (numel == 0) return B2S_OK;
  B2S_CHECK_ARG(x && y, "null pointer");
  gelu_fwd_kernel<<<grid_for(numel, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(x, numel, y);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_add_gelu_fwd(const float* a, const float* b, int64_t numel, float* sum, float* y,
                                    b2s_stream_t stream) {
  B2S_CHECK_ARG(numel >= 0 && numel % 4 == 0, "numel must be a non-negative multiple of 4");
  if (numel == 0) return B2S_OK;
  B2S_CHECK_ARG(a && b && sum && y, "null pointer");
  add_gelu_fwd_kernel<<<grid_for(numel / 4, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), numel / 4,
      reinterpret_cast<float4*>(sum), reinterpret_cast<float4*>(y));
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_gelu_bwd(const float* gy, const float* x, int64_t numel, float* gx, b2s_stream_t stream) {
  B2S_CHECK_ARG(numel >= 0, "numel >= 0");
  if (numel == 0) return B2S_OK;
  B2S_CHECK_ARG(gy && x && gx, "null pointer");
  gelu_bwd_kernel<<<grid_for(numel, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(gy, x, numel, gx);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}
