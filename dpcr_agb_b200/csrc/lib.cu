// lib.cu -- library plumbing (errors, version, device check) and the convolution dispatch of the C ABI.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

static thread_local char g_last_error[512] = "";

void b2s_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

extern "C" const char* b2s_last_error(void) { return g_last_error; }
extern "C" int32_t b2s_version(void) { return 100; /* 0.1.0 */ }

extern "C" int32_t b2s_device_check(void) {
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  B2S_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B2S_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    b2s_set_error("b2s_device_check: device %d is sm_%d%d; libb200sparse is built for sm_100a (B200) only", dev, major,
                  minor);
    return B2S_ECUDA;
  }
  return B2S_OK;
}

// launch-tuning knobs (wgrad_tc.cu / conv_tc.cu read them at every launch)
int g_b2s_wg_nbp = -1, g_b2s_wg_lag = -1, g_b2s_wg_occ2 = -1, g_b2s_tc_rot = -1;
int g_b2s_tc_ca = -1, g_b2s_tc_occ1 = -1, g_b2s_wg_ca = -1;
int g_b2s_wg_wv = -1, g_b2s_tc_m256 = -1, g_b2s_tc_ta = -1;
int g_b2s_cr_v4 = -1, g_b2s_cr_cap = -1, g_b2s_cr_unroll = -1;
int g_b2s_precise = -1;

// Operand mode of the tensor-core convolutions: 1 (default) = split-bf16 pairs, 16-17 significant bits per operand,
// three kind::f16 products per term; 0 = TF32 round-to-nearest operands, one kind::tf32 product.  Environment
// B2S_PRECISE read once; b2s_set_tuning("precise", v) overrides.  Unlike the other knobs this one changes results.
int b2s_precise() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("B2S_PRECISE");
    env = e ? (atoi(e) != 0 ? 1 : 0) : 1;
  }
  return g_b2s_precise >= 0 ? (g_b2s_precise != 0 ? 1 : 0) : env;
}

extern "C" int32_t b2s_set_tuning(const char* key, int32_t value) {
  B2S_CHECK_ARG(key != nullptr, "key is NULL");
  if (!strcmp(key, "wg_nbp")) g_b2s_wg_nbp = value;
  else if (!strcmp(key, "wg_lag")) g_b2s_wg_lag = value;
  else if (!strcmp(key, "wg_occ2")) g_b2s_wg_occ2 = value;
  else if (!strcmp(key, "tc_rot")) g_b2s_tc_rot = value;
  else if (!strcmp(key, "tc_ca")) g_b2s_tc_ca = value;
  else if (!strcmp(key, "tc_occ1")) g_b2s_tc_occ1 = value;
  else if (!strcmp(key, "wg_ca")) g_b2s_wg_ca = value;
  else if (!strcmp(key, "wg_wv")) g_b2s_wg_wv = value;
  else if (!strcmp(key, "tc_m256")) g_b2s_tc_m256 = value;
  else if (!strcmp(key, "tc_ta")) g_b2s_tc_ta = value;
  else if (!strcmp(key, "cr_v4")) g_b2s_cr_v4 = value;
  else if (!strcmp(key, "cr_cap")) g_b2s_cr_cap = value;
  else if (!strcmp(key, "cr_unroll")) g_b2s_cr_unroll = value;
  else if (!strcmp(key, "precise")) g_b2s_precise = value;
  else {
    b2s_set_error("b2s_set_tuning: unknown key '%s'", key);
    return B2S_EINVAL;
  }
  return B2S_OK;
}

// implemented in conv_simt.cu / conv_tc.cu
int b2s_conv_gather_gemm_simt(const float* x, const float* w, const float* bias, const int32_t* nbr, int64_t n_out,
                              const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, int32_t w_layout,
                              float* y, cudaStream_t st);
int b2s_conv_wgrad_simt(const float* x, const float* gy, const int32_t* nbr, int64_t n_out, const int32_t* n_out_dev,
                        int32_t c_in, int32_t c_out, int32_t k3, float* gw, cudaStream_t st);
bool b2s_conv_tc_supported(int32_t c_in, int32_t c_out, int32_t k3, int64_t n_out);
int64_t b2s_conv_tc_workspace_bytes(int32_t c_in, int32_t c_out, int32_t k3, int64_t n_in);
int b2s_conv_gather_gemm_tc(const float* x, const float* w, const float* bias, const int32_t* nbr, int64_t n_in,
                            int64_t n_out, const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, int32_t w_layout, float* y,
                            void* workspace, int64_t workspace_bytes, cudaStream_t st, float* col_stats,
                            int* stats_rows);
// pointwise.cu: per-128-row partial column sums / sums of squares of x into col_stats, and the header of that buffer
void b2s_launch_col_partials(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* col_stats, cudaStream_t st);
bool b2s_wgrad_tc_supported(int32_t c_in, int32_t c_out, int32_t k3, int64_t n_out, bool has_map);
int64_t b2s_wgrad_tc_workspace_bytes(int32_t c_in, int64_t n_in);
int b2s_conv_wgrad_tc(const float* x, const float* gy, const int32_t* nbr, int64_t n_in, int64_t n_out,
                      const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, float* gw, void* workspace, cudaStream_t st);

void b2s_launch_round_tf32(const float* in, int64_t n, const int32_t* n_dev, int32_t c, float* out, cudaStream_t st);

static inline int64_t al256(int64_t b) { return (b + 255) & ~(int64_t)255; }
// bytes of the internally rounded operand copies (0 when the caller pre-rounds; c_in <= 4 is rounded while padding)
static int64_t round_copy_fwd(int64_t n_in, int32_t c_in, int32_t pre) { return (pre || c_in <= 4) ? 0 : al256(n_in * c_in * 4); }
static int64_t round_copy_wg_x(int64_t n_in, int32_t c_in, int32_t pre) { return (pre || c_in <= 4) ? 0 : al256(n_in * c_in * 4); }
static int64_t round_copy_wg_gy(int64_t n_out, int32_t c_out, int32_t pre) { return pre ? 0 : al256(n_out * c_out * 4); }

extern "C" int64_t b2s_conv_workspace_bytes(int64_t n_in, int64_t n_out, int32_t c_in, int32_t c_out, int32_t k3,
                                            int32_t prerounded) {
  if (c_in <= 0 || c_out <= 0 || k3 <= 0 || n_in < 0 || n_out < 0) return -1;
  int64_t fwd = b2s_conv_tc_supported(c_in, c_out, k3, n_out)
                    ? al256(b2s_conv_tc_workspace_bytes(c_in, c_out, k3, n_in)) + round_copy_fwd(n_in, c_in, prerounded)
                    : 0;
  int64_t wg = b2s_wgrad_tc_supported(c_in, c_out, k3, n_out, true)
                   ? al256(b2s_wgrad_tc_workspace_bytes(c_in, n_in)) + round_copy_wg_x(n_in, c_in, prerounded) +
                         round_copy_wg_gy(n_out, c_out, prerounded)
                   : 0;
  return fwd > wg ? fwd : wg;
}

extern "C" int32_t b2s_conv_gather_gemm(const float* x, const float* w, const float* bias, const int32_t* nbr,
                                        int64_t n_in, int64_t n_out, const int32_t* n_out_dev, int32_t c_in,
                                        int32_t c_out, int32_t k3, int32_t w_layout, float* y, void* workspace,
                                        int64_t workspace_bytes, int32_t impl, float* col_stats,
                                        b2s_stream_t stream) {
  B2S_CHECK_ARG(n_in >= 0 && n_out >= 0 && c_in > 0 && c_out > 0 && k3 > 0, "bad sizes");
  B2S_CHECK_ARG(w_layout >= 0 && w_layout <= 31 && (w_layout & 8) == 0, "w_layout: bits 0-2 and 4");
  B2S_CHECK_ARG(!(w_layout & 16) || ((w_layout & 4) && c_in > 4),
                "a prebuilt weight image (bit 4) needs pre-rounded operands (bit 2) and c_in > 4");
  B2S_CHECK_ARG(impl >= 0 && impl <= 2, "impl must be 0, 1 or 2");
  B2S_CHECK_ARG(nbr || (k3 == 1 && n_in == n_out), "nbr may be null only for the identity map (k3 == 1)");
  if (n_out == 0) return B2S_OK;
  B2S_CHECK_ARG(x && w && y, "null pointer");
  cudaStream_t st = as_stream(stream);
  const bool tc_ok = b2s_conv_tc_supported(c_in, c_out, k3, n_out);
  if (impl == 2 && !tc_ok) {
    b2s_set_error("b2s_conv_gather_gemm: tcgen05 kernel does not cover c_in=%d c_out=%d k3=%d", c_in, c_out, k3);
    return B2S_EINVAL;
  }
  if (impl == 2 || (impl == 0 && tc_ok)) {
    const int pre = (w_layout & 4) ? 1 : 0;
    const int64_t tc_bytes = al256(b2s_conv_tc_workspace_bytes(c_in, c_out, k3, n_in));
    B2S_CHECK_ARG(workspace && workspace_bytes >= tc_bytes + round_copy_fwd(n_in, c_in, pre),
                  "workspace too small (see b2s_conv_workspace_bytes)");
    B2S_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                  "workspace must be 256-byte aligned, x and y 16-byte aligned");
    const float* xin = x;
    if (!pre && c_in > 4) {   // tcgen05 truncates fp32 operands: round the gathered operand to nearest first
      float* xr = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + tc_bytes);
      b2s_launch_round_tf32(x, n_in, nullptr, c_in, xr, st);
      xin = xr;
    }
    int stats_rows = 0;
    if ((w_layout & 16) && !tc_ok) {
      b2s_set_error("b2s_conv_gather_gemm: prebuilt weight image for a shape the tcgen05 kernel does not cover");
      return B2S_EINVAL;
    }
    if (b2s_conv_gather_gemm_tc(xin, w, bias, nbr, n_in, n_out, n_out_dev, c_in, c_out, k3, w_layout & 19, y, workspace,
                                workspace_bytes, st, col_stats, &stats_rows))
      return B2S_ECUDA;
    if (col_stats && stats_rows == 0) b2s_launch_col_partials(y, n_out, n_out_dev, c_out, col_stats, st);
  } else {
    if ((w_layout & 4) && b2s_precise()) {
      b2s_set_error("b2s_conv_gather_gemm: operand-form (split-bf16) input on the SIMT path (c_in=%d c_out=%d impl=%d)",
                    c_in, c_out, impl);
      return B2S_EINVAL;
    }
    b2s_conv_gather_gemm_simt(x, w, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, w_layout & 3, y, st);
    if (col_stats) b2s_launch_col_partials(y, n_out, n_out_dev, c_out, col_stats, st);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int b2s_conv_dgrad_perm_tc(const float* x, const float* w, const int32_t* nbr, int64_t n_fine, const int32_t* n_fine_dev,
                           int32_t c_in, int32_t c_out, const int32_t* ksize, int32_t w_layout, const int32_t* perm,
                           const int32_t* bounds, float* y, void* workspace, cudaStream_t st);
int64_t b2s_conv_tc_image_bytes(int32_t c_in, int32_t c_out, int32_t k3);
int64_t b2s_conv_tc_prebuilt_image_bytes(int32_t c_in, int32_t c_out, int32_t k3);
int b2s_conv_weight_image_tc(const float* w, int32_t c_in, int32_t c_out, int32_t k3, int32_t w_layout, float* img,
                             cudaStream_t st);

extern "C" int64_t b2s_conv_weight_image_bytes(int32_t c_in, int32_t c_out, int32_t k3) {
  if (c_in <= 4 || c_in % 32 != 0 || c_out % 64 != 0 || k3 <= 0) return -1;   // shapes of the tcgen05 kernels only
  return al256(b2s_conv_tc_prebuilt_image_bytes(c_in, c_out, k3));
}

extern "C" int32_t b2s_conv_weight_image(const float* w, int32_t c_in, int32_t c_out, int32_t k3, int32_t w_layout,
                                         void* img, int64_t img_bytes, b2s_stream_t stream) {
  B2S_CHECK_ARG(w && img && w_layout >= 0 && w_layout <= 3, "bad arguments");
  const int64_t need = b2s_conv_weight_image_bytes(c_in, c_out, k3);
  B2S_CHECK_ARG(need > 0 && img_bytes >= need && (reinterpret_cast<uintptr_t>(img) & 255) == 0,
                "shape not covered, or image buffer too small / misaligned");
  if (b2s_conv_weight_image_tc(w, c_in, c_out, k3, w_layout, reinterpret_cast<float*>(img), as_stream(stream)))
    return B2S_ECUDA;
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int64_t b2s_conv_dgrad_strided_workspace_bytes(int32_t c_gy, int32_t c_x, int32_t k3) {
  if (c_gy <= 0 || c_x <= 0 || k3 <= 0) return -1;
  return al256(b2s_conv_tc_image_bytes(c_gy, c_x, k3));
}

extern "C" int32_t b2s_conv_dgrad_strided(const float* gy, const float* w, const int32_t* inv_nbr, const int32_t* perm,
                                          const int32_t* bounds, int64_t n_fine, const int32_t* n_fine_dev,
                                          int32_t c_gy, int32_t c_x, const int32_t* kernel_size_host, float* gx,
                                          void* workspace, int64_t workspace_bytes, int32_t flags,
                                          b2s_stream_t stream) {
  B2S_CHECK_ARG(n_fine >= 0 && c_gy > 0 && c_x > 0 && kernel_size_host, "bad sizes");
  const int k3 = kernel_size_host[0] * kernel_size_host[1] * kernel_size_host[2];
  B2S_CHECK_ARG(k3 >= 1 && k3 <= 27, "kernel volume must be 1..27");
  B2S_CHECK_ARG(c_gy % 32 == 0 && c_x % 64 == 0, "tcgen05 path: c_gy % 32 == 0 and c_x % 64 == 0");
  if (n_fine == 0) return B2S_OK;
  B2S_CHECK_ARG(gy && w && inv_nbr && perm && bounds && gx, "null pointer");
  B2S_CHECK_ARG(workspace && workspace_bytes >= b2s_conv_dgrad_strided_workspace_bytes(c_gy, c_x, k3) &&
                    (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                "workspace too small or misaligned");
  if (b2s_conv_dgrad_perm_tc(gy, w, inv_nbr, n_fine, n_fine_dev, c_gy, c_x, kernel_size_host, 1 | ((flags & 1) ? 16 : 0),
                             perm, bounds, gx, workspace, as_stream(stream)))
    return B2S_ECUDA;
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_conv_wgrad(const float* x, const float* gy, const int32_t* nbr, int64_t n_in, int64_t n_out,
                                  const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, float* gw,
                                  void* workspace, int64_t workspace_bytes, int32_t impl, int32_t flags,
                                  b2s_stream_t stream) {
  B2S_CHECK_ARG(n_in >= 0 && n_out >= 0 && c_in > 0 && c_out > 0 && k3 > 0 && gw, "bad sizes");
  B2S_CHECK_ARG(impl >= 0 && impl <= 2, "impl must be 0, 1 or 2");
  B2S_CHECK_ARG(nbr || (k3 == 1 && n_in == n_out), "nbr may be null only for the identity map (k3 == 1)");
  cudaStream_t st = as_stream(stream);
  if (n_out == 0) {
    B2S_CUDA(cudaMemsetAsync(gw, 0, (size_t)k3 * c_in * c_out * sizeof(float), st));
    return B2S_OK;
  }
  B2S_CHECK_ARG(x && gy, "null pointer");
  const bool tc_ok = b2s_wgrad_tc_supported(c_in, c_out, k3, n_out, nbr != nullptr);
  if (impl == 2 && !tc_ok) {
    b2s_set_error("b2s_conv_wgrad: tcgen05 kernel does not cover c_in=%d c_out=%d k3=%d", c_in, c_out, k3);
    return B2S_EINVAL;
  }
  if (impl == 2 || (impl == 0 && tc_ok)) {
    const int pre = flags & 1;
    const int64_t own = al256(b2s_wgrad_tc_workspace_bytes(c_in, n_in));
    const int64_t need = own + round_copy_wg_x(n_in, c_in, pre) + round_copy_wg_gy(n_out, c_out, pre);
    B2S_CHECK_ARG(need == 0 || (workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0),
                  "workspace too small or not 256-byte aligned (see b2s_conv_workspace_bytes)");
    const float *xin = x, *gin = gy;
    if (!pre) {
      char* wsp = reinterpret_cast<char*>(workspace) + own;
      if (c_in > 4) {
        b2s_launch_round_tf32(x, n_in, nullptr, c_in, reinterpret_cast<float*>(wsp), st);
        xin = reinterpret_cast<const float*>(wsp);
        wsp += round_copy_wg_x(n_in, c_in, pre);
      }
      b2s_launch_round_tf32(gy, n_out, n_out_dev, c_out, reinterpret_cast<float*>(wsp), st);
      gin = reinterpret_cast<const float*>(wsp);
    }
    if (b2s_conv_wgrad_tc(xin, gin, nbr, n_in, n_out, n_out_dev, c_in, c_out, k3, gw, workspace, st)) return B2S_ECUDA;
  } else {
    if ((flags & 1) && b2s_precise()) {
      b2s_set_error("b2s_conv_wgrad: operand-form (split-bf16) inputs on the SIMT path (c_in=%d c_out=%d impl=%d)", c_in,
                    c_out, impl);
      return B2S_EINVAL;
    }
    b2s_conv_wgrad_simt(x, gy, nbr, n_out, n_out_dev, c_in, c_out, k3, gw, st);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}
