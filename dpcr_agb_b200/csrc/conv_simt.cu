// conv_simt.cu -- fp32 SIMT implementation of the output-stationary gather-GEMM (forward / dgrad) and
// of wgrad.  This is the bit-for-bit fp32 yardstick for the tcgen05 kernels in conv_tc.cu (same neighbour
// table, same accumulation order per out row) and the path taken for shapes the tensor-core kernel does
// not cover (c_in or c_out not a multiple of 8, tiny maps).
//
// Reference call sites: R:modules/MinkowskiEngine/SENet.py:49-52,94-97; resnet_block.py:48-54,95-107.
#include "common.cuh"

namespace {

constexpr int BM = 64;   // out rows per CTA
constexpr int BN = 64;   // out channels per CTA
constexpr int THREADS = 256;

// y[o, n0:n0+64] = bias + sum_k sum_ci x[nbr[k,o], ci] * B_k[ci, co]
template <int BK>
__global__ void __launch_bounds__(THREADS) gather_gemm_simt_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ bias,
                                                                   const int* __restrict__ nbr, int64_t n_out,
                                                                   const int* __restrict__ n_out_dev, int c_in,
                                                                   int c_out, int k3, int w_layout,
                                                                   float* __restrict__ y) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  if ((int64_t)blockIdx.x * BM >= n_out) return;
  __shared__ int idx_s[BM];
  __shared__ float a_s[BK][BM + 4];
  __shared__ float b_s[BK][BN + 4];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < k3; ++k) {
    int my = -1;
    if (t < BM) {
      const int64_t o = m0 + t;
      my = o < n_out ? (nbr ? nbr[(int64_t)k * pitch + o] : (int)o) : -1;
      idx_s[t] = my;
    }
    if (!__syncthreads_or(my >= 0)) continue;  // no row of this tile has a neighbour at offset k
    for (int c0 = 0; c0 < c_in; c0 += BK) {
      // A tile: BM rows x BK input channels, stored transposed
      for (int e = t; e < BM * BK; e += THREADS) {
        const int m = e / BK, kk = e % BK;
        const int i = idx_s[m];
        const int ci = c0 + kk;
        a_s[kk][m] = (i >= 0 && ci < c_in) ? __ldg(&x[(int64_t)i * c_in + ci]) : 0.f;
      }
      // B tile: BK input channels x BN output channels
      for (int e = t; e < BK * BN; e += THREADS) {
        int kk, nn;
        if (!(w_layout & 1)) {
          kk = e / BN;
          nn = e % BN;
        } else {
          nn = e / BK;
          kk = e % BK;
        }
        const int ci = c0 + kk, co = n0 + nn;
        float v = 0.f;
        const int kw = (w_layout & 2) ? k3 - 1 - k : k;  // bit 1: kernel index reversed (symmetric maps)
        if (ci < c_in && co < c_out)
          v = !(w_layout & 1) ? __ldg(&w[((int64_t)kw * c_in + ci) * c_out + co])
                              : __ldg(&w[((int64_t)kw * c_out + co) * c_in + ci]);
        b_s[kk][nn] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&a_s[kk][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&b_s[kk][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t o = m0 + ty * 4 + i;
    if (o >= n_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co < c_out) y[o * c_out + co] = acc[i][j] + (bias ? __ldg(&bias[co]) : 0.f);
    }
  }
}

// gw[k, ci0:ci0+64, co0:co0+64] += sum over a slice of out rows of x[nbr[k,o], ci] * gy[o, co]
constexpr int WG_ROWS = 16;
__global__ void __launch_bounds__(THREADS) wgrad_simt_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ gy,
                                                             const int* __restrict__ nbr, int64_t n_out,
                                                             const int* __restrict__ n_out_dev, int c_in, int c_out,
                                                             int co_tiles, int64_t rows_per_split,
                                                             float* __restrict__ gw) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  __shared__ int idx_s[WG_ROWS];
  __shared__ float a_s[WG_ROWS][BM + 4];  // [row][ci]
  __shared__ float b_s[WG_ROWS][BN + 4];  // [row][co]
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int k = blockIdx.x;
  const int ci0 = (blockIdx.y / co_tiles) * BM;
  const int co0 = (blockIdx.y % co_tiles) * BN;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(r_begin + rows_per_split, n_out);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += WG_ROWS) {
    int my = -1;
    if (t < WG_ROWS) {
      const int64_t o = r0 + t;
      my = o < r_end ? (nbr ? nbr[(int64_t)k * pitch + o] : (int)o) : -1;
      idx_s[t] = my;
    }
    if (!__syncthreads_or(my >= 0)) continue;
    for (int e = t; e < WG_ROWS * BM; e += THREADS) {
      const int r = e / BM, c = e % BM;
      const int i = idx_s[r];
      const int ci = ci0 + c;
      a_s[r][c] = (i >= 0 && ci < c_in) ? __ldg(&x[(int64_t)i * c_in + ci]) : 0.f;
    }
    for (int e = t; e < WG_ROWS * BN; e += THREADS) {
      const int r = e / BN, c = e % BN;
      const int co = co0 + c;
      b_s[r][c] = (idx_s[r] >= 0 && co < c_out) ? __ldg(&gy[(r0 + r) * c_out + co]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WG_ROWS; ++r) {
      const float4 a4 = *reinterpret_cast<const float4*>(&a_s[r][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&b_s[r][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= c_in) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < c_out && acc[i][j] != 0.f) atomicAdd(&gw[((int64_t)k * c_in + ci) * c_out + co], acc[i][j]);
    }
  }
}

// wgrad for c_in <= 4 (the k7 stem, c_in = 3): one warp walks 32 out rows at a time, lanes own output
// channels; rows without a neighbour at this offset (89 % for the stem) are skipped warp-uniformly.
constexpr int SC_WARPS = 8;
__global__ void __launch_bounds__(SC_WARPS * 32) wgrad_smallcin_kernel(const float* __restrict__ x,
                                                                      const float* __restrict__ gy,
                                                                      const int* __restrict__ nbr, int64_t n_out,
                                                                      const int* __restrict__ n_out_dev, int c_in,
                                                                      int c_out, int64_t rows_per_split,
                                                                      float* __restrict__ gw) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  __shared__ float red[SC_WARPS][4][64];
  const int k = blockIdx.x;
  const int co0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(r_begin + rows_per_split, n_out);
  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
  const int coA = co0 + lane, coB = co0 + 32 + lane;
  for (int64_t r0 = r_begin + warp * 32; r0 < r_end; r0 += SC_WARPS * 32) {
    const int64_t o_l = r0 + lane;
    const int mine = o_l < r_end ? nbr[(int64_t)k * pitch + o_l] : -1;
    unsigned live = __ballot_sync(0xffffffffu, mine >= 0);
    while (live) {
      const int src = __ffs(live) - 1;
      live &= live - 1;
      const int i = __shfl_sync(0xffffffffu, mine, src);
      const int64_t o = r0 + src;
      const float ga = coA < c_out ? __ldg(&gy[o * c_out + coA]) : 0.f;
      const float gb = coB < c_out ? __ldg(&gy[o * c_out + coB]) : 0.f;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (ci < c_in) {
          const float xv = __ldg(&x[(int64_t)i * c_in + ci]);
          acc[ci][0] = fmaf(xv, ga, acc[ci][0]);
          acc[ci][1] = fmaf(xv, gb, acc[ci][1]);
        }
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < 4; ++ci) {
    red[warp][ci][lane] = acc[ci][0];
    red[warp][ci][lane + 32] = acc[ci][1];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 4 * 64; e += SC_WARPS * 32) {
    const int ci = e / 64, c = e % 64;
    if (ci >= c_in || co0 + c >= c_out) continue;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < SC_WARPS; ++wv) s += red[wv][ci][c];
    if (s != 0.f) atomicAdd(&gw[((int64_t)k * c_in + ci) * c_out + co0 + c], s);
  }
}

}  // namespace

int b2s_conv_gather_gemm_simt(const float* x, const float* w, const float* bias, const int32_t* nbr, int64_t n_out,
                              const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, int32_t w_layout,
                              float* y, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(n_out, BM), (unsigned)((c_out + BN - 1) / BN));
  if (c_in <= 4)
    gather_gemm_simt_kernel<4><<<grid, THREADS, 0, st>>>(x, w, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, w_layout, y);
  else
    gather_gemm_simt_kernel<16><<<grid, THREADS, 0, st>>>(x, w, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, w_layout, y);
  return 0;
}

int b2s_conv_wgrad_simt(const float* x, const float* gy, const int32_t* nbr, int64_t n_out, const int32_t* n_out_dev,
                        int32_t c_in, int32_t c_out, int32_t k3, float* gw, cudaStream_t st) {
  cudaMemsetAsync(gw, 0, (size_t)k3 * c_in * c_out * sizeof(float), st);
  if (c_in <= 4 && nbr) {
    const int co_tiles = (c_out + 63) / 64;
    int64_t splits = (4LL * B2S_NUM_SMS) / ((int64_t)k3 * co_tiles) + 1;
    const int64_t max_splits = ceil_div64(n_out, SC_WARPS * 32);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t rows = ceil_div64(n_out, splits);
    rows = ceil_div64(rows, SC_WARPS * 32) * (SC_WARPS * 32);
    splits = ceil_div64(n_out, rows);
    dim3 grid((unsigned)k3, (unsigned)co_tiles, (unsigned)splits);
    wgrad_smallcin_kernel<<<grid, SC_WARPS * 32, 0, st>>>(x, gy, nbr, n_out, n_out_dev, c_in, c_out, rows, gw);
    return 0;
  }
  const int ci_tiles = (c_in + BM - 1) / BM, co_tiles = (c_out + BN - 1) / BN;
  int64_t base = (int64_t)k3 * ci_tiles * co_tiles;
  int64_t splits = (4LL * B2S_NUM_SMS + base - 1) / base;
  const int64_t max_splits = ceil_div64(n_out, WG_ROWS * 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rows = ceil_div64(n_out, splits);
  rows = ceil_div64(rows, WG_ROWS) * WG_ROWS;
  splits = ceil_div64(n_out, rows);
  dim3 grid((unsigned)k3, (unsigned)(ci_tiles * co_tiles), (unsigned)splits);
  wgrad_simt_kernel<<<grid, THREADS, 0, st>>>(x, gy, nbr, n_out, n_out_dev, c_in, c_out, co_tiles, rows, gw);
  return 0;
}
