// pointwise.cu -- the bandwidth-bound stages: max pooling (a9), per-plot reductions and broadcasts
// (a5, a10, a11), batch-norm statistics / apply (a13), GELU (a14), column sums (bias gradient).
// Every kernel streams fp32 [N, C] rows: consecutive threads own consecutive 16-byte channel vectors of a row
// (then the next row), so a warp touches 512 contiguous bytes; per-channel parameters live in registers.
// All of them are bounded by HBM.  Row counts may live on the device (n_dev, see b200sparse.h "row counts").
//
// Reference call sites (R: = /root/reference/torch-points3d/torch_points3d/modules/MinkowskiEngine/):
//   a9 R:SENet.py:53   a10/a11 R:senet_block.py:43-50, R:common.py:44-48, R:SENet.py:63,117
//   a13 R:SENet.py:35,51,98 + R:resnet_block.py:51-55   a14 R:common.py:41
#include "common.cuh"
#include "tc_ptx.cuh"
#include <math.h>

extern int g_b2s_cr_v4, g_b2s_cr_cap, g_b2s_cr_unroll;   // lib.cu (b2s_set_tuning)

namespace {

constexpr int PW_THREADS = 256;

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// ---------------------------------------------------------------- vector helpers ------------
template <int VEC>
struct V {
  float v[VEC];
};
template <int VEC>
__device__ __forceinline__ V<VEC> ldv(const float* p) {
  V<VEC> r;
  if (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x, r.v[1 % VEC] = t.y, r.v[2 % VEC] = t.z, r.v[3 % VEC] = t.w;
  } else {
    r.v[0] = *p;
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ V<VEC> ldgv(const float* p) {
  V<VEC> r;
  if (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x, r.v[1 % VEC] = t.y, r.v[2 % VEC] = t.z, r.v[3 % VEC] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const V<VEC>& r) {
  if (VEC == 4)
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1 % VEC], r.v[2 % VEC], r.v[3 % VEC]);
  else
    *p = r.v[0];
}
template <int VEC>
__device__ __forceinline__ V<VEC> splat(float s) {
  V<VEC> r;
#pragma unroll
  for (int j = 0; j < VEC; ++j) r.v[j] = s;
  return r;
}
// TF32 round-to-nearest copy of a vector: the operand form of the tensor-core convolutions.  Producers whose output
// feeds a convolution write it alongside (or instead of) the plain result, which takes the separate b2s_round_tf32
// pass -- one more read and write of the tensor, one more launch -- off the step.
template <int VEC>
__device__ __forceinline__ V<VEC> tf32v(const V<VEC>& a) {
  V<VEC> r;
#pragma unroll
  for (int j = 0; j < VEC; ++j) r.v[j] = __uint_as_float(tc::rna_tf32(__float_as_uint(a.v[j])));
  return r;
}
// Operand-form store of VEC consecutive channels (ch ..) of row r of an [n, c] matrix: the layout the tensor-core
// convolutions gather.  mode 0: TF32 round-to-nearest values in place.  mode 1 (split-bf16, see tc_ptx.cuh): every
// 32-channel block of a row keeps its 128 bytes and holds the 32 bf16 high parts in the first 64 bytes and the 32 bf16
// residuals in the last 64 -- a 16-byte chunk is then either eight high parts or eight residuals, so that a gathered
// row lands as [h | l] halves of a K-major tile and as separate H / L atoms of an MN-major one.
template <int VEC>
__device__ __forceinline__ void st_operand(float* base, int64_t r, int c, int ch, const V<VEC>& a, int mode) {
  if (mode == 0) {
    stv<VEC>(base + r * c + ch, tf32v<VEC>(a));
    return;
  }
  char* blk = reinterpret_cast<char*>(base + r * c + (ch & ~31)) + (ch & 31) * 2;
  uint32_t h[VEC], l[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) tc::split_bf16(a.v[j], h[j], l[j]);
  if (VEC == 4) {
    *reinterpret_cast<uint2*>(blk) = make_uint2(h[0] | (h[1 % VEC] << 16), h[2 % VEC] | (h[3 % VEC] << 16));
    *reinterpret_cast<uint2*>(blk + 64) = make_uint2(l[0] | (l[1 % VEC] << 16), l[2 % VEC] | (l[3 % VEC] << 16));
  } else {
    *reinterpret_cast<unsigned short*>(blk) = (unsigned short)h[0];
    *reinterpret_cast<unsigned short*>(blk + 64) = (unsigned short)l[0];
  }
}
// per-channel parameter vector (nullable pointer -> constant)
template <int VEC>
__device__ __forceinline__ V<VEC> ldparam(const float* p, int ch, float dflt) {
  return p ? ldgv<VEC>(p + ch) : splat<VEC>(dflt);
}

// Thread -> (row lane, channel vector) mapping shared by the row-streaming kernels.
//   cv = channel vectors per row, tpr = threads per row, rpb = rows per block pass
struct RowMap {
  int cv, tpr, rpb, tr, tc;
  bool active;
};
template <int VEC>
__device__ __forceinline__ RowMap row_map(int c) {
  RowMap m;
  m.cv = c / VEC;
  m.tpr = m.cv < (int)blockDim.x ? m.cv : (int)blockDim.x;
  m.rpb = (int)blockDim.x / m.tpr;
  m.tr = (int)threadIdx.x / m.tpr;
  m.tc = (int)threadIdx.x % m.tpr;
  m.active = m.tr < m.rpb;
  return m;
}
#define B2S_ROW_LOOP(m, n, r)                                                                       \
  for (int64_t r = (int64_t)blockIdx.x * (m).rpb + (m).tr; r < (n); r += (int64_t)gridDim.x * (m).rpb)

// rows per pass of the unrolled row-streaming kernels (loads of all rows of a pass issued first); B2S_PW_UNROLL=1
// selects the one-row-per-pass form.  Measured (tools/bn_bench.py, graph replays): +5-10 % on the 256 k- and 422 k-row
// layers, -10-25 % below ~4 M elements (too few rows per thread), hence the size threshold.
constexpr int64_t PW_UNROLL_MIN = (int64_t)1 << 22;
static int pw_unroll() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2S_PW_UNROLL");
    v = e ? atoi(e) : 4;
  }
  return v;
}

static inline int rows_grid(int64_t n, int c, int vec) {
  const int cv = c / vec;
  const int tpr = cv < PW_THREADS ? cv : PW_THREADS;
  const int rpb = PW_THREADS / tpr;
  return grid_for(ceil_div64(n, rpb) * PW_THREADS, PW_THREADS);
}
static inline int vec_of(int c, const void* a, const void* b = nullptr, const void* d = nullptr, const void* e = nullptr) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return (c % 4 == 0 && al(a) && al(b) && al(d) && al(e)) ? 4 : 1;
}

// ---------------------------------------------------------------- max pooling ---------------
template <int VEC>
__global__ void __launch_bounds__(PW_THREADS) maxpool_fwd_kernel(const float* __restrict__ x,
                                                                 const int* __restrict__ nbr, int64_t n_out,
                                                                 const int* __restrict__ n_dev, int c, int k3,
                                                                 float* __restrict__ y, int* __restrict__ arg,
                                                                 float* __restrict__ y_tf32, int opm, int batched) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_dev);
  const RowMap m = row_map<VEC>(c);
  if (!m.active) return;
  for (int v0 = m.tc; v0 < m.cv; v0 += m.tpr) {
    const int ch = v0 * VEC;
    B2S_ROW_LOOP(m, n_out, o) {
      V<VEC> best = splat<VEC>(-INFINITY);
      int bi[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) bi[j] = -1;
      // nine kernel offsets at a time: their table entries, then the rows of the existing neighbours, are all in
      // flight before the first comparison (one offset after the other made every row a dependent pair of loads:
      // 0.19 ms for the k3 s2 pool of a 32-plot batch, 3x its HBM time)
      if (!batched) {
        for (int k = 0; k < k3; ++k) {
          const int i = __ldg(&nbr[(int64_t)k * pitch + o]);
          if (i < 0) continue;
          const V<VEC> v = ldgv<VEC>(x + (int64_t)i * c + ch);
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            if (bi[j] < 0 || v.v[j] > best.v[j] || (v.v[j] == best.v[j] && i < bi[j])) {
              best.v[j] = v.v[j];
              bi[j] = i;
            }
          }
        }
      } else
      for (int k0 = 0; k0 < k3; k0 += 9) {
        int idx[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) idx[q] = k0 + q < k3 ? __ldg(&nbr[(int64_t)(k0 + q) * pitch + o]) : -1;
        V<VEC> v[9];
#pragma unroll
        for (int q = 0; q < 9; ++q)
          if (idx[q] >= 0) v[q] = ldgv<VEC>(x + (int64_t)idx[q] * c + ch);
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          const int i = idx[q];
          if (i < 0) continue;
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            if (bi[j] < 0 || v[q].v[j] > best.v[j] || (v[q].v[j] == best.v[j] && i < bi[j])) {
              best.v[j] = v[q].v[j];
              bi[j] = i;
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if (bi[j] < 0) best.v[j] = 0.f;
      if (VEC == 4) *reinterpret_cast<int4*>(arg + o * c + ch) = make_int4(bi[0], bi[1], bi[2], bi[3]);
      else
#pragma unroll
        for (int j = 0; j < VEC; ++j) arg[o * c + ch + j] = bi[j];
      stv<VEC>(y + o * c + ch, best);
      if (y_tf32) st_operand<VEC>(y_tf32, o, c, ch, best, opm);
    }
  }
}

// Max pooling with a WARP per 32 consecutive out rows (c = 4 * TPR, TPR threads per row, k3 <= 27).  The kernel above
// reads its 27 table entries per row group as 27 separate sectors with 8 useful bytes each and then walks the offsets
// one dependent (entry -> row) load pair at a time.  Here the warp first stages the entries of its 32 rows coalesced
// (one 128-byte line per offset, all offsets in flight) in shared memory together with a presence mask per row; a row
// then visits only the offsets that exist, four row loads in flight at a time.  Same comparisons in the same
// (ascending offset) order: identical values and arg-max rows.
constexpr int MP_WARPS = 8;
template <int TPR>
__global__ void __launch_bounds__(MP_WARPS * 32) maxpool_fwd_warp_kernel(const float* __restrict__ x,
                                                                         const int* __restrict__ nbr, int64_t n_out,
                                                                         const int* __restrict__ n_dev, int c, int k3,
                                                                         float* __restrict__ y, int* __restrict__ arg,
                                                                         float* __restrict__ y_tf32, int opm) {
  __shared__ int s_idx[MP_WARPS][27][32];
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_dev);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int RPP = 32 / TPR;                 // rows per pass
  const int sub = lane / TPR, ch = (lane % TPR) * 4;
  for (int64_t o0 = ((int64_t)blockIdx.x * MP_WARPS + warp) * 32; o0 < n_out; o0 += (int64_t)gridDim.x * MP_WARPS * 32) {
    const int64_t ol = o0 + lane;
    unsigned mymask = 0;
#pragma unroll 9
    for (int k = 0; k < k3; ++k) {
      const int i = ol < n_out ? __ldg(nbr + (int64_t)k * pitch + ol) : -1;
      s_idx[warp][k][lane] = i;
      mymask |= (i >= 0 ? 1u : 0u) << k;
    }
    __syncwarp();
#pragma unroll 1
    for (int p = 0; p < 32 / RPP; ++p) {
      const int r = p * RPP + sub;
      const int64_t o = o0 + r;
      unsigned mask = __shfl_sync(0xffffffffu, mymask, r);
      V<4> best = splat<4>(-INFINITY);
      int bi[4] = {-1, -1, -1, -1};
      while (mask) {
        int ii[4];
        V<4> vv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ii[q] = -1;
          if (mask) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            ii[q] = s_idx[warp][k][r];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (ii[q] >= 0) vv[q] = ldgv<4>(x + (int64_t)ii[q] * c + ch);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (ii[q] < 0) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (bi[j] < 0 || vv[q].v[j] > best.v[j] || (vv[q].v[j] == best.v[j] && ii[q] < bi[j])) {
              best.v[j] = vv[q].v[j];
              bi[j] = ii[q];
            }
          }
        }
      }
      if (o < n_out) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (bi[j] < 0) best.v[j] = 0.f;
        *reinterpret_cast<int4*>(arg + o * c + ch) = make_int4(bi[0], bi[1], bi[2], bi[3]);
        stv<4>(y + o * c + ch, best);
        if (y_tf32) st_operand<4>(y_tf32, o, c, ch, best, opm);
      }
    }
    __syncwarp();                               // the next 32 rows overwrite the staged entries
  }
}

// ---------------------------------------------------------------- local sum / average pooling --
// y[o,:] = post[o] * sum_k pre[i] * x[i,:], i = nbr[k,o] >= 0 (pre / post nullable).  Forward of MinkowskiSumPooling /
// MinkowskiAvgPooling (post = 1 / #neighbours) on the forward table, and their backward on the TRANSPOSED table with
// pre = the same per-out-row factor (gx[i] = sum over the out rows o that pooled i of gy[o] / count[o]).
template <int VEC>
__global__ void __launch_bounds__(PW_THREADS) sumpool_kernel(const float* __restrict__ x, const int* __restrict__ nbr,
                                                             const float* __restrict__ pre,
                                                             const float* __restrict__ post, int64_t n_out,
                                                             const int* __restrict__ n_dev, int c, int k3,
                                                             float* __restrict__ y) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_dev);
  const RowMap m = row_map<VEC>(c);
  if (!m.active) return;
  for (int v0 = m.tc; v0 < m.cv; v0 += m.tpr) {
    const int ch = v0 * VEC;
    B2S_ROW_LOOP(m, n_out, o) {
      V<VEC> acc = splat<VEC>(0.f);
      for (int k = 0; k < k3; ++k) {
        const int i = __ldg(&nbr[(int64_t)k * pitch + o]);
        if (i < 0) continue;
        const V<VEC> v = ldgv<VEC>(x + (int64_t)i * c + ch);
        const float f = pre ? __ldg(&pre[i]) : 1.f;
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc.v[j] = fmaf(v.v[j], f, acc.v[j]);
      }
      if (post) {
        const float f = __ldg(&post[o]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc.v[j] *= f;
      }
      stv<VEC>(y + o * c + ch, acc);
    }
  }
}

// inv[o] = 1 / max(1, number of k with nbr[k,o] >= 0)
__global__ void __launch_bounds__(PW_THREADS) nbr_inv_counts_kernel(const int* __restrict__ nbr, int k3, int64_t n,
                                                                    const int* __restrict__ n_dev,
                                                                    float* __restrict__ inv) {
  const int64_t pitch = n;
  n = b2s_rows(n, n_dev);
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    int cnt = 0;
    for (int k = 0; k < k3; ++k) cnt += __ldg(&nbr[(int64_t)k * pitch + o]) >= 0;
    inv[o] = 1.f / (float)(cnt > 0 ? cnt : 1);
  }
}

__global__ void __launch_bounds__(PW_THREADS) maxpool_bwd_kernel(const float* __restrict__ gy,
                                                                 const int* __restrict__ arg, int64_t n_out,
                                                                 const int* __restrict__ n_dev, int c,
                                                                 float* __restrict__ gx) {
  n_out = b2s_rows(n_out, n_dev);
  const int64_t total = n_out * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int a = arg[e];
    if (a >= 0) atomicAdd(&gx[(int64_t)a * c + (int)(e % c)], gy[e]);
  }
}

// ---------------------------------------------------------------- per-plot segments ---------
__global__ void __launch_bounds__(PW_THREADS) batch_counts_kernel(const int* __restrict__ rb, int stride, int64_t n,
                                                                  const int* __restrict__ n_dev, int nb,
                                                                  int* __restrict__ counts) {
  n = b2s_rows(n, n_dev);
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
                                                                 const int* __restrict__ n_dev, int c, int nb,
                                                                 float* __restrict__ y) {
  n = b2s_rows(n, n_dev);
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

// Per-plot maximum (MinkowskiGlobalMaxPooling, R:networks.py:39, R:PointNet.py:28): y[b, c] = max over the rows of plot b,
// arg[b, c] = the lowest row holding it (-1 and y = 0 for a plot without rows).  Rows are batch-sorted, so a CTA =
// (plot, 64-channel slab) finds its row range by bisection of the batch column and reduces it without atomics:
// 16 row lanes x 16 channel quads, then a shared-memory fold over the row lanes.  Deterministic.
__global__ void __launch_bounds__(256) segment_max_kernel(const float* __restrict__ x, const int* __restrict__ rb,
                                                          int stride, int64_t n, const int* __restrict__ n_dev, int c,
                                                          float* __restrict__ y, int* __restrict__ arg) {
  n = b2s_rows(n, n_dev);
  const int b = blockIdx.x, c0 = blockIdx.y * 64;
  auto lower = [&](int key) {       // first row with batch id >= key
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(&rb[mid * stride]) < key) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  const int64_t r0 = lower(b), r1 = lower(b + 1);
  const int q = threadIdx.x & 15, lane = threadIdx.x >> 4;       // channel quad, row lane
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int bi[4] = {-1, -1, -1, -1};
  const int ch = c0 + q * 4;
  if (ch < c) {
    for (int64_t r = r0 + lane; r < r1; r += 16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ch + j < c) {
          const float v = __ldg(&x[r * c + ch + j]);
          if (bi[j] < 0 || v > best[j]) {          // rows ascend per lane: strict > keeps the lowest row on ties
            best[j] = v;
            bi[j] = (int)r;
          }
        }
      }
    }
  }
  __shared__ float sv[16][64];
  __shared__ int si[16][64];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sv[lane][q * 4 + j] = best[j];
    si[lane][q * 4 + j] = bi[j];
  }
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < c) {
    float v = -INFINITY;
    int i = -1;
    for (int l = 0; l < 16; ++l) {
      const int il = si[l][threadIdx.x];
      const float vl = sv[l][threadIdx.x];
      if (il >= 0 && (i < 0 || vl > v || (vl == v && il < i))) {
        v = vl;
        i = il;
      }
    }
    y[(int64_t)b * c + c0 + threadIdx.x] = i >= 0 ? v : 0.f;
    arg[(int64_t)b * c + c0 + threadIdx.x] = i;
  }
}

// gx[arg[b, c], c] = gy[b, c] into a zeroed gx
__global__ void __launch_bounds__(256) segment_max_bwd_kernel(const float* __restrict__ gy, const int* __restrict__ arg,
                                                              int64_t total, int c, float* __restrict__ gx) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = arg[e];
    if (i >= 0) gx[(int64_t)i * c + (e % c)] = gy[e];
  }
}

// out[r, :] = y[batch(r), :] * scale[batch(r)]
template <int VEC>
__global__ void __launch_bounds__(PW_THREADS) segment_bcast_kernel(const float* __restrict__ y,
                                                                   const int* __restrict__ rb, int stride, int64_t n,
                                                                   const int* __restrict__ n_dev, int c,
                                                                   const float* __restrict__ scale,
                                                                   float* __restrict__ out) {
  n = b2s_rows(n, n_dev);
  const RowMap m = row_map<VEC>(c);
  if (!m.active) return;
  for (int v0 = m.tc; v0 < m.cv; v0 += m.tpr) {
    const int ch = v0 * VEC;
    B2S_ROW_LOOP(m, n, r) {
      const int b = __ldg(&rb[r * stride]);
      V<VEC> v = ldgv<VEC>(y + (int64_t)b * c + ch);
      if (scale) {
        const float s = __ldg(&scale[b]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v.v[j] *= s;
      }
      stv<VEC>(out + r * c + ch, v);
    }
  }
}

// out[r, :] = x[r, :] * y[batch(r), :]      (y_c == 1: per-plot scalar)
template <int VEC>
__global__ void __launch_bounds__(PW_THREADS) bcast_mul_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ y,
                                                               const int* __restrict__ rb, int stride, int64_t n,
                                                               const int* __restrict__ n_dev, int c, int y_c,
                                                               float* __restrict__ out) {
  n = b2s_rows(n, n_dev);
  const RowMap m = row_map<VEC>(c);
  if (!m.active) return;
  for (int v0 = m.tc; v0 < m.cv; v0 += m.tpr) {
    const int ch = v0 * VEC;
    B2S_ROW_LOOP(m, n, r) {
      const int b = __ldg(&rb[r * stride]);
      V<VEC> v = ldv<VEC>(x + r * c + ch);
      if (y_c == 1) {
        const float s = __ldg(&y[b]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v.v[j] *= s;
      } else {
        const V<VEC> g = ldgv<VEC>(y + (int64_t)b * c + ch);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v.v[j] *= g.v[j];
      }
      stv<VEC>(out + r * c + ch, v);
    }
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
                                                               const float* __restrict__ beta, int64_t n,
                                                               const int* __restrict__ n_dev, int c, int act,
                                                               ACC* __restrict__ ws) {
  n = b2s_rows(n, n_dev);
  const int ch = blockIdx.y * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_CTA;
  if (r0 >= n) return;
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

// Vectorised form of the same reductions (c % 64 == 0, 16-byte aligned rows).  A 512-thread CTA owns a slab of 64
// channels (blockIdx.y) and a strided share of the rows (blockIdx.x): 16 threads x float4 per row, 32 row lanes, U
// independent 16-byte loads in flight per operand and thread (64 KB per SM).  Partial sums are combined across the
// row lanes by a shared-memory tree and leave the CTA as one atomic per channel.  Same-address fp64 atomics
// serialise at ~20 ns each in L2 (measured: tools/pw_bench.py), so the launch keeps slabs x row shares at about one
// CTA per SM: the chain on any accumulator is at most 148 deep (the scalar kernel issues one per 256 rows).
constexpr int CR_THREADS = 512;
constexpr int CR_SLAB = 64;                         // channels per CTA
constexpr int CR_TPR = CR_SLAB / 4;                 // threads per row
constexpr int CR_RPB = CR_THREADS / CR_TPR;         // row lanes
// What the LAST CTA of a column reduction does with the finished sums (it is elected by a ticket counter stored behind
// the 2c accumulators; the caller zeroes accumulators and ticket with one memset): nothing, the batch-norm statistics
// (mean / invstd / running buffers) or the fp64 -> fp32 copy of the backward sums.  Saves one dependent launch per call.
struct CrFinal {
  int kind;              // 0 none, 1 batch-norm statistics, 2 copy to float
  float eps, momentum;
  float* running_mean;
  float* running_var;
  float* out0;           // kind 1: mean [c]; kind 2: sums [2c]
  float* out1;           // kind 1: invstd [c]
};

template <int MODE, typename ACC, int U>
__global__ void __launch_bounds__(CR_THREADS, 1) colreduce_v4_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ g,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ invstd,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, int64_t n,
                                                                     const int* __restrict__ n_dev, int c, int act,
                                                                     ACC* __restrict__ ws, CrFinal fin) {
  n = b2s_rows(n, n_dev);
  __shared__ float4 sm[2][CR_THREADS];
  const int tc = threadIdx.x % CR_TPR, tr = threadIdx.x / CR_TPR;
  const int ch = blockIdx.y * CR_SLAB + tc * 4;
  const int64_t step = (int64_t)gridDim.x * CR_RPB;
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
  V<4> mu = splat<4>(0.f), is = splat<4>(1.f), ga = splat<4>(1.f), be = splat<4>(0.f);
  if (MODE == 2) {
    mu = ldgv<4>(mean + ch), is = ldgv<4>(invstd + ch);
    ga = ldparam<4>(gamma, ch, 1.f), be = ldparam<4>(beta, ch, 0.f);
  }
  auto accumulate = [&](const V<4>& xv, const V<4>& gv) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = xv.v[j];
      if (MODE == 0) {
        a0[j] += v;
      } else if (MODE == 1) {
        a0[j] += v;
        a1[j] += v * v;
      } else {
        const float xh = (v - mu.v[j]) * is.v[j];
        float gg = gv.v[j];
        if (act == 1) gg *= gelu_grad_f(xh * ga.v[j] + be.v[j]);
        a0[j] += gg;
        a1[j] += gg * xh;
      }
    }
  };
  int64_t r = (int64_t)blockIdx.x * CR_RPB + tr;
  for (; r + (U - 1) * step < n; r += U * step) {
    V<4> xv[U], gv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) xv[u] = ldv<4>(x + (r + u * step) * c + ch);
    if (MODE == 2) {
#pragma unroll
      for (int u = 0; u < U; ++u) gv[u] = ldv<4>(g + (r + u * step) * c + ch);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) accumulate(xv[u], gv[u]);
  }
  for (; r < n; r += step) {   // < U rows left
    const V<4> xv = ldv<4>(x + r * c + ch);
    V<4> gv = xv;
    if (MODE == 2) gv = ldv<4>(g + r * c + ch);
    accumulate(xv, gv);
  }
  sm[0][threadIdx.x] = make_float4(a0[0], a0[1], a0[2], a0[3]);
  if (MODE != 0) sm[1][threadIdx.x] = make_float4(a1[0], a1[1], a1[2], a1[3]);
  __syncthreads();
#pragma unroll
  for (int sft = CR_RPB >> 1; sft > 0; sft >>= 1) {
    if (tr < sft) {
      const int o = threadIdx.x + sft * CR_TPR;
      float4 p = sm[0][threadIdx.x], q = sm[0][o];
      sm[0][threadIdx.x] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
      if (MODE != 0) {
        p = sm[1][threadIdx.x], q = sm[1][o];
        sm[1][threadIdx.x] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < CR_SLAB) {       // one atomic per thread: channel blockIdx.y * 64 + threadIdx.x
    const float* s0 = reinterpret_cast<const float*>(&sm[0][0]);
    const float* s1 = reinterpret_cast<const float*>(&sm[1][0]);
    const int cc = blockIdx.y * CR_SLAB + threadIdx.x;
    atomicAdd(&ws[cc], (ACC)s0[threadIdx.x]);
    if (MODE != 0) atomicAdd(&ws[c + cc], (ACC)s1[threadIdx.x]);
  }
  if (MODE != 0 && fin.kind != 0) {
    __shared__ int is_last;
    __threadfence();                                   // this CTA's atomics are visible before its ticket
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned* ticket = reinterpret_cast<unsigned*>(ws + 2 * c);
      is_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      const double dn = n > 0 ? (double)n : 1.0;
      if (fin.kind == 1) {
        for (int ch2 = threadIdx.x; ch2 < c; ch2 += blockDim.x) {
          const double m = (double)__ldcg(&ws[ch2]) / dn;
          double var = (double)__ldcg(&ws[c + ch2]) / dn - m * m;
          if (var < 0.0) var = 0.0;
          fin.out0[ch2] = (float)m;
          fin.out1[ch2] = (float)(1.0 / sqrt(var + (double)fin.eps));
          if (fin.running_mean)
            fin.running_mean[ch2] = (1.f - fin.momentum) * fin.running_mean[ch2] + fin.momentum * (float)m;
          if (fin.running_var) {
            const double unb = n > 1 ? var * dn / (dn - 1.0) : var;
            fin.running_var[ch2] = (1.f - fin.momentum) * fin.running_var[ch2] + fin.momentum * (float)unb;
          }
        }
      } else {
        for (int e = threadIdx.x; e < 2 * c; e += blockDim.x) fin.out0[e] = (float)__ldcg(&ws[e]);
      }
    }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ ws, int64_t n, const int* __restrict__ n_dev, int c,
                                   float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean,
                                   float* __restrict__ invstd) {
  n = b2s_rows(n, n_dev);
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double dn = n > 0 ? (double)n : 1.0;
  const double m = ws[ch] / dn;
  double var = ws[c + ch] / dn - m * m;
  if (var < 0.0) var = 0.0;
  mean[ch] = (float)m;
  invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
  if (running_var) {
    const double unb = n > 1 ? var * dn / (dn - 1.0) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
  }
}

__global__ void double_to_float_kernel(const double* __restrict__ in, float* __restrict__ out, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = (float)in[i];
}

template <int VEC, int PW_U>
__global__ void __launch_bounds__(PW_THREADS) bn_apply_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ invstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int64_t n,
                                                              const int* __restrict__ n_dev, int c, int act,
                                                              float* __restrict__ y, float* __restrict__ y_tf32,
                                                              int opm) {
  n = b2s_rows(n, n_dev);
  const RowMap m = row_map<VEC>(c);
  if (!m.active) return;
  for (int v0 = m.tc; v0 < m.cv; v0 += m.tpr) {
    const int ch = v0 * VEC;
    const V<VEC> mu = ldgv<VEC>(mean + ch), is = ldgv<VEC>(invstd + ch);
    const V<VEC> ga = ldparam<VEC>(gamma, ch, 1.f), be = ldparam<VEC>(beta, ch, 0.f);
    // PW_U rows per pass with all of their loads issued first: one 16-byte load per thread and pass kept 16 KB in
    // flight per SM (1 024 resident threads), a third of what HBM needs; the kernel ran at half of the copy bandwidth
    const int64_t stride = (int64_t)gridDim.x * m.rpb;
    for (int64_t r0 = (int64_t)blockIdx.x * m.rpb + m.tr; r0 < n; r0 += stride * PW_U) {
      V<VEC> vu[PW_U];
#pragma unroll
      for (int u = 0; u < PW_U; ++u)
        if (r0 + u * stride < n) vu[u] = ldv<VEC>(x + (r0 + u * stride) * c + ch);
#pragma unroll
      for (int u = 0; u < PW_U; ++u) {
        const int64_t r = r0 + u * stride;
        if (r >= n) break;
        V<VEC> v = vu[u];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float t = (v.v[j] - mu.v[j]) * is.v[j];
          t = t * ga.v[j] + be.v[j];
          v.v[j] = act == 1 ? gelu_f(t) : t;
        }
        if (y) stv<VEC>(y + r * c + ch, v);
        if (y_tf32) st_operand<VEC>(y_tf32, r, c, ch, v, opm);
      }
    }
  }
}

// COLSUM: also accumulate the column sums of gx (the bias gradient of the convolution in front of the batch norm, which
// otherwise re-reads gx in its own b2s_colsum launch): per-thread partial sums over the thread's rows, a shared-memory
// tree over the row lanes of the block, one partial ROW per block (row blockIdx.x of colsum; the consumer adds the
// rows -- ~1 000 same-address atomics per channel serialised in L2 and cost more than the pass they replaced).  Needs
// every thread of the block active and the same number of channel passes for all of them (the launcher checks).
template <int VEC, bool COLSUM, int PW_U>
__global__ void __launch_bounds__(PW_THREADS) bn_bwd_apply_kernel(
    const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ sums, int64_t n, const int* __restrict__ n_dev, int c, int act, int training,
    float* __restrict__ gx, float* __restrict__ gx_tf32, int opm, float* __restrict__ colsum) {
  __shared__ float red[COLSUM ? PW_THREADS * VEC : 1];
  n = b2s_rows(n, n_dev);
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  const RowMap m = row_map<VEC>(c);
  if (!m.active) return;
  for (int v0 = m.tc; v0 < m.cv; v0 += m.tpr) {
    const int ch = v0 * VEC;
    const V<VEC> mu = ldgv<VEC>(mean + ch), is = ldgv<VEC>(invstd + ch);
    const V<VEC> ga = ldparam<VEC>(gamma, ch, 1.f), be = ldparam<VEC>(beta, ch, 0.f);
    V<VEC> s0 = splat<VEC>(0.f), s1 = splat<VEC>(0.f);
    V<VEC> cs = splat<VEC>(0.f);
    if (training) {
      s0 = ldgv<VEC>(sums + ch);
      s1 = ldgv<VEC>(sums + c + ch);
    }
    const int64_t stride = (int64_t)gridDim.x * m.rpb;           // PW_U rows per pass, loads first (see bn_apply_kernel)
    for (int64_t r0 = (int64_t)blockIdx.x * m.rpb + m.tr; r0 < n; r0 += stride * PW_U) {
      V<VEC> xu[PW_U], gu[PW_U];
#pragma unroll
      for (int u = 0; u < PW_U; ++u) {
        if (r0 + u * stride < n) {
          xu[u] = ldv<VEC>(x + (r0 + u * stride) * c + ch);
          gu[u] = ldv<VEC>(gy + (r0 + u * stride) * c + ch);
        }
      }
#pragma unroll
      for (int u = 0; u < PW_U; ++u) {
        const int64_t r = r0 + u * stride;
        if (r >= n) break;
        const V<VEC> xv = xu[u];
        V<VEC> g = gu[u];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float xh = (xv.v[j] - mu.v[j]) * is.v[j];
          float t = g.v[j];
          if (act == 1) t *= gelu_grad_f(xh * ga.v[j] + be.v[j]);
          if (training) t = t - s0.v[j] * inv_n - xh * s1.v[j] * inv_n;
          g.v[j] = t * ga.v[j] * is.v[j];
          if (COLSUM) cs.v[j] += g.v[j];
        }
        stv<VEC>(gx + r * c + ch, g);
        if (gx_tf32) st_operand<VEC>(gx_tf32, r, c, ch, g, opm);
      }
    }
    if (COLSUM) {
      __syncthreads();                     // the previous channel pass has read its sums
#pragma unroll
      for (int j = 0; j < VEC; ++j) red[threadIdx.x * VEC + j] = cs.v[j];
      __syncthreads();
      for (int sft = m.rpb >> 1; sft > 0; sft >>= 1) {   // rpb is a power of two here
        if (m.tr < sft) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) red[threadIdx.x * VEC + j] += red[(threadIdx.x + sft * m.tpr) * VEC + j];
        }
        __syncthreads();
      }
      if (m.tr == 0) {                     // this block's partial row (plain stores; the consumer adds the rows)
#pragma unroll
        for (int j = 0; j < VEC; ++j) colsum[(int64_t)blockIdx.x * c + ch + j] = red[threadIdx.x * VEC + j];
      }
    }
  }
}

// flat elementwise kernels (numel = rows * c)
template <int VEC, int OP>  // OP 0: y = gelu(x)   1: gx = gy * gelu'(x)   2: s = a + b, y = gelu(s)   3: y = tf32(x)
__global__ void __launch_bounds__(PW_THREADS) flat_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          int64_t n, const int* __restrict__ n_dev, int c,
                                                          float* __restrict__ o0, float* __restrict__ o1,
                                                          float* __restrict__ o2, int opm) {
  n = b2s_rows(n, n_dev);
  const int64_t total = n * c / VEC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const V<VEC> av = ldv<VEC>(a + e * VEC);
    V<VEC> r0, r1;
    if (OP == 0) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) r0.v[j] = gelu_f(av.v[j]);
      stv<VEC>(o0 + e * VEC, r0);
    } else if (OP == 3) {
      st_operand<VEC>(o0, (e * VEC) / c, c, (int)((e * VEC) % c), av, opm);
    } else if (OP == 1) {
      const V<VEC> bv = ldv<VEC>(b + e * VEC);
#pragma unroll
      for (int j = 0; j < VEC; ++j) r0.v[j] = av.v[j] * gelu_grad_f(bv.v[j]);
      stv<VEC>(o0 + e * VEC, r0);
    } else {
      const V<VEC> bv = ldv<VEC>(b + e * VEC);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        r0.v[j] = av.v[j] + bv.v[j];
        r1.v[j] = gelu_f(r0.v[j]);
      }
      stv<VEC>(o0 + e * VEC, r0);
      if (o1) stv<VEC>(o1 + e * VEC, r1);
      if (o2) st_operand<VEC>(o2, (e * VEC) / c, c, (int)((e * VEC) % c), r1, opm);
    }
  }
}

// ---------------------------------------------------------------- fused SE block tail ---------
// The tail of every SE residual block (R:senet_block.py:33-50,83-94):
//     y = act( drop_path( u * sigmoid(fc2(act(fc1(mean_plot(u))))) ) + residual )
// as three kernels forward (per-plot mean = segment_sum; the excitation MLP of ALL plots in one launch; gate x
// drop-path scale x residual add x GELU in one pass over the rows) and three backward -- instead of ~14 launches
// forward and ~20 backward with 21 passes over the [N, C] tensors.  Same arithmetic in the same order per element
// (u * (gate * keep) + res); the MLP sums are fp32 FMAs over <= 2048 terms.

// The excitation MLP, forward: (1) one warp per (plot, hidden unit): h_pre = W1 p + b1 -- every plot streams W1 once;
// (2) CTA = (plot, chunk of SE_CHUNK gate channels): h = gelu(h_pre) in shared memory, a warp per gate channel:
// gate = sigmoid(W2 h + b2), gate_eff = gate * keep[plot].  (A single kernel that recomputed the hidden layer per
// channel chunk read W1 sixteen times per plot at 2048 channels: 68 us per call in MSENet50.)
constexpr int SE_CHUNK = 128;
__global__ void __launch_bounds__(256) se_hidden_kernel(const float* __restrict__ pooled, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, int c, int h,
                                                        float* __restrict__ h_pre) {
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.y * 8 + warp;
  if (j >= h) return;
  const float* p = pooled + (int64_t)b * c;
  const float* w = w1 + (int64_t)j * c;
  float acc = 0.f;
  for (int i = lane; i < c; i += 32) acc = fmaf(__ldg(&w[i]), __ldg(&p[i]), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) h_pre[(int64_t)b * h + j] = acc + (b1 ? b1[j] : 0.f);
}

__global__ void __launch_bounds__(256) se_gate_fwd_kernel(const float* __restrict__ h_pre, const float* __restrict__ w2,
                                                          const float* __restrict__ b2, const float* __restrict__ keep,
                                                          int c, int h, float* __restrict__ gate,
                                                          float* __restrict__ gate_eff) {
  extern __shared__ float se_sm[];
  float* hh = se_sm;         // [h]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int j = tid; j < h; j += blockDim.x) hh[j] = gelu_f(h_pre[(int64_t)b * h + j]);
  __syncthreads();
  const float kp = keep ? keep[b] : 1.f;
  const int i1 = min(c, (int)(blockIdx.y + 1) * SE_CHUNK);
  for (int i = blockIdx.y * SE_CHUNK + warp; i < i1; i += (int)(blockDim.x >> 5)) {
    float acc = 0.f;
    for (int j = lane; j < h; j += 32) acc = fmaf(__ldg(&w2[(int64_t)i * h + j]), hh[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float z = acc + (b2 ? b2[i] : 0.f);
      const float g = 1.f / (1.f + expf(-z));
      gate[(int64_t)b * c + i] = g;
      gate_eff[(int64_t)b * c + i] = g * kp;
    }
  }
}

// Backward: (1) one CTA per plot: gz2 = gradient at the sigmoid input (stored, and kept in shared memory),
// gh[j] = sum_i gz2[i] W2[i, j] with thread = hidden unit (W2 rows are read coalesced; 256 / h thread groups split the
// channel range and are combined in shared memory), gh_pre = gh * gelu'(h_pre); also h = gelu(h_pre) for the
// parameter-gradient kernel (gh_pre: [2, B, h] scratch, second half = h).  (2) thread per channel:
// g_pooled[i] = inv_count * sum_j gh_pre[j] W1[j, i] (coalesced over i).  Every plot streams W2 and W1 once.
__global__ void __launch_bounds__(256) se_gate_bwd_kernel(const float* __restrict__ g_gate_eff,
                                                          const float* __restrict__ keep, const float* __restrict__ gate,
                                                          const float* __restrict__ h_pre, const float* __restrict__ w2,
                                                          int nb, int c, int h, float* __restrict__ gz2,
                                                          float* __restrict__ gh_pre) {
  extern __shared__ float se_sm[];
  float* z = se_sm;          // [c]
  float* part = se_sm + c;   // [256] partial sums of the thread groups
  const int b = blockIdx.x, tid = threadIdx.x;
  const float kp = keep ? keep[b] : 1.f;
  for (int i = tid; i < c; i += blockDim.x) {
    const float g = gate[(int64_t)b * c + i];
    const float v = g_gate_eff[(int64_t)b * c + i] * kp * g * (1.f - g);
    z[i] = v;
    gz2[(int64_t)b * c + i] = v;
  }
  __syncthreads();
  // hidden units are handled in rounds of `hw` (<= 256) units; within a round 256 / hw groups split the channels
  const int hw = h < 256 ? h : 256;
  const int groups = 256 / hw;
  const int jl = tid % hw, grp = tid / hw;
  for (int j0 = 0; j0 < h; j0 += hw) {
    const int j = j0 + jl;
    float acc = 0.f;
    if (grp < groups && j < h)
      for (int i = grp; i < c; i += groups) acc = fmaf(z[i], __ldg(&w2[(int64_t)i * h + j]), acc);
    part[tid] = acc;
    __syncthreads();
    if (grp == 0 && j < h) {
      for (int q = 1; q < groups; ++q) acc += part[q * hw + jl];
      const float pre = h_pre[(int64_t)b * h + j];
      gh_pre[(int64_t)b * h + j] = acc * gelu_grad_f(pre);
      gh_pre[((int64_t)nb + b) * h + j] = gelu_f(pre);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) se_pooled_grad_kernel(const float* __restrict__ gh_pre,
                                                             const float* __restrict__ w1,
                                                             const float* __restrict__ inv_count, int c, int h,
                                                             float* __restrict__ g_pooled) {
  extern __shared__ float se_sm[];
  float* hp = se_sm;         // [h]
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < h; j += blockDim.x) hp[j] = gh_pre[(int64_t)b * h + j];
  __syncthreads();
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float acc = 0.f;
  for (int j = 0; j < h; ++j) acc = fmaf(hp[j], __ldg(&w1[(int64_t)j * c + i]), acc);
  g_pooled[(int64_t)b * c + i] = acc * (inv_count ? inv_count[b] : 1.f);
}

// parameter gradients of the two linears: sums over the plots (h_act = gelu(h_pre), stored by se_gate_bwd_kernel)
__global__ void __launch_bounds__(256) se_param_grad_kernel(const float* __restrict__ gz2,
                                                            const float* __restrict__ gh_pre,
                                                            const float* __restrict__ h_act,
                                                            const float* __restrict__ pooled, int nb, int c, int h,
                                                            float* __restrict__ gw1, float* __restrict__ gb1,
                                                            float* __restrict__ gw2, float* __restrict__ gb2) {
  const int hc = h * c;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 2 * hc + h + c; e += gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (e < hc) {                         // gw1[j, i] = sum_b gh_pre[b, j] * pooled[b, i]
      const int j = e / c, i = e % c;
      for (int b = 0; b < nb; ++b) acc = fmaf(gh_pre[b * h + j], pooled[(int64_t)b * c + i], acc);
      gw1[e] = acc;
    } else if (e < 2 * hc) {              // gw2[i, j] = sum_b gz2[b, i] * h[b, j]
      const int f = e - hc, i = f / h, j = f % h;
      for (int b = 0; b < nb; ++b) acc = fmaf(gz2[(int64_t)b * c + i], h_act[b * h + j], acc);
      gw2[f] = acc;
    } else if (e < 2 * hc + h) {
      const int j = e - 2 * hc;
      for (int b = 0; b < nb; ++b) acc += gh_pre[b * h + j];
      if (gb1) gb1[j] = acc;
    } else {
      const int i = e - 2 * hc - h;
      for (int b = 0; b < nb; ++b) acc += gz2[(int64_t)b * c + i];
      if (gb2) gb2[i] = acc;
    }
  }
}

// s = u * gate_eff[plot] + res ; y = gelu(s)   (y and / or its TF32 twin)
template <int VEC>
__global__ void __launch_bounds__(PW_THREADS) gated_add_gelu_fwd_kernel(
    const float* __restrict__ u, const float* __restrict__ gate_eff, const float* __restrict__ res,
    const int* __restrict__ rb, int stride, int64_t n, const int* __restrict__ n_dev, int c, float* __restrict__ s,
    float* __restrict__ y, float* __restrict__ y_tf32, int opm) {
  n = b2s_rows(n, n_dev);
  const RowMap m = row_map<VEC>(c);
  if (!m.active) return;
  for (int v0 = m.tc; v0 < m.cv; v0 += m.tpr) {
    const int ch = v0 * VEC;
    B2S_ROW_LOOP(m, n, r) {
      const int b = __ldg(&rb[r * stride]);
      const V<VEC> g = ldgv<VEC>(gate_eff + (int64_t)b * c + ch);
      const V<VEC> uv = ldv<VEC>(u + r * c + ch), rv = ldv<VEC>(res + r * c + ch);
      V<VEC> sv, yv;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        sv.v[j] = uv.v[j] * g.v[j] + rv.v[j];
        yv.v[j] = gelu_f(sv.v[j]);
      }
      stv<VEC>(s + r * c + ch, sv);
      if (y) stv<VEC>(y + r * c + ch, yv);
      if (y_tf32) st_operand<VEC>(y_tf32, r, c, ch, yv, opm);
    }
  }
}

// g_s = gy * gelu'(s) (gradient of the residual), g_u = g_s * gate_eff[plot], g_gate[plot] += g_s * u.
// CTA = 64-channel slab (blockIdx.y) x CONTIGUOUS row chunk (blockIdx.x), so that a thread's rows (chunk rows
// tr, tr + 32, ...) stay within one plot almost always: the per-plot product sum lives in registers and is flushed
// with one fp32 atomic per channel when the plot changes and at the end.
__global__ void __launch_bounds__(CR_THREADS, 1) gated_add_gelu_bwd_kernel(
    const float* __restrict__ gy, const float* __restrict__ s, const float* __restrict__ u,
    const float* __restrict__ gate_eff, const int* __restrict__ rb, int stride, int64_t n,
    const int* __restrict__ n_dev, int c, int nb, int64_t rows_per_cta, float* __restrict__ g_s,
    float* __restrict__ g_u, float* __restrict__ g_gate) {
  n = b2s_rows(n, n_dev);
  const int tc = threadIdx.x % CR_TPR, tr = threadIdx.x / CR_TPR;
  const int ch = blockIdx.y * CR_SLAB + tc * 4;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(r0 + rows_per_cta, n);
  int cur = -1;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  V<4> g = splat<4>(0.f);
  auto flush = [&]() {
    if (cur >= 0 && cur < nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&g_gate[(int64_t)cur * c + ch + j], acc[j]);
    }
  };
  constexpr int U = 2;
  for (int64_t r = r0 + tr; r < r1; r += U * CR_RPB) {
    V<4> gv[U], sv[U], uv[U];
    int bb[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int64_t rr = r + q * CR_RPB;
      const bool ok = rr < r1;
      bb[q] = ok ? __ldg(&rb[rr * stride]) : -1;
      gv[q] = ok ? ldv<4>(gy + rr * c + ch) : splat<4>(0.f);
      sv[q] = ok ? ldv<4>(s + rr * c + ch) : splat<4>(0.f);
      uv[q] = ok ? ldv<4>(u + rr * c + ch) : splat<4>(0.f);
    }
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int64_t rr = r + q * CR_RPB;
      if (rr >= r1) break;
      if (bb[q] != cur) {
        flush();
        cur = bb[q];
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        g = (cur >= 0 && cur < nb) ? ldgv<4>(gate_eff + (int64_t)cur * c + ch) : splat<4>(0.f);
      }
      V<4> gs, gu;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gs.v[j] = gv[q].v[j] * gelu_grad_f(sv[q].v[j]);
        gu.v[j] = gs.v[j] * g.v[j];
        acc[j] = fmaf(gs.v[j], uv[q].v[j], acc[j]);
      }
      stv<4>(g_s + rr * c + ch, gs);
      stv<4>(g_u + rr * c + ch, gu);
    }
  }
  flush();
}

// float4 form of segment_sum (c % 64 == 0): CTA = 64-channel slab x CONTIGUOUS row chunk, a thread's rows (chunk rows
// tr, tr + 32, ...) almost always belong to one plot, whose partial sum stays in registers and leaves as one fp32
// atomic per channel (times the per-plot scale, e.g. 1 / rows for the mean) when the plot changes and at the end.
__global__ void __launch_bounds__(CR_THREADS, 1) segment_sum_v4_kernel(const float* __restrict__ x,
                                                                       const int* __restrict__ rb, int stride,
                                                                       int64_t n, const int* __restrict__ n_dev, int c,
                                                                       int nb, int64_t rows_per_cta,
                                                                       const float* __restrict__ scale,
                                                                       float* __restrict__ y) {
  n = b2s_rows(n, n_dev);
  const int tc = threadIdx.x % CR_TPR, tr = threadIdx.x / CR_TPR;
  const int ch = blockIdx.y * CR_SLAB + tc * 4;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(r0 + rows_per_cta, n);
  int cur = -1;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  auto flush = [&]() {
    if (cur >= 0 && cur < nb) {
      const float sc = scale ? __ldg(&scale[cur]) : 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&y[(int64_t)cur * c + ch + j], acc[j] * sc);
    }
  };
  constexpr int U = 4;
  for (int64_t r = r0 + tr; r < r1; r += U * CR_RPB) {
    V<4> xv[U];
    int bb[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int64_t rr = r + q * CR_RPB;
      const bool ok = rr < r1;
      bb[q] = ok ? __ldg(&rb[rr * stride]) : -1;
      xv[q] = ok ? ldv<4>(x + rr * c + ch) : splat<4>(0.f);
    }
#pragma unroll
    for (int q = 0; q < U; ++q) {
      if (r + q * CR_RPB >= r1) break;
      if (bb[q] != cur) {
        flush();
        cur = bb[q];
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += xv[q].v[j];
    }
  }
  flush();
}

// x[r, :] += y[plot(r), :]
template <int VEC>
__global__ void __launch_bounds__(PW_THREADS) bcast_add_kernel(float* __restrict__ x, const float* __restrict__ y,
                                                               const int* __restrict__ rb, int stride, int64_t n,
                                                               const int* __restrict__ n_dev, int c) {
  n = b2s_rows(n, n_dev);
  const RowMap m = row_map<VEC>(c);
  if (!m.active) return;
  for (int v0 = m.tc; v0 < m.cv; v0 += m.tpr) {
    const int ch = v0 * VEC;
    B2S_ROW_LOOP(m, n, r) {
      const int b = __ldg(&rb[r * stride]);
      const V<VEC> a = ldgv<VEC>(y + (int64_t)b * c + ch);
      V<VEC> v = ldv<VEC>(x + r * c + ch);
#pragma unroll
      for (int j = 0; j < VEC; ++j) v.v[j] += a.v[j];
      stv<VEC>(x + r * c + ch, v);
    }
  }
}

dim3 colgrid(int64_t n, int c) { return dim3((unsigned)ceil_div64(n, ROWS_PER_CTA), (unsigned)((c + 31) / 32)); }

// column reduction launch: the float4 kernel whenever the rows are 16-byte aligned vectors, else the scalar one
template <int MODE, typename ACC>
bool launch_colreduce(const float* x, const float* g, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, int64_t n, const int* n_dev, int c, int act, ACC* ws, cudaStream_t st,
                      CrFinal fin = CrFinal{}) {
  const int v4 = g_b2s_cr_v4 >= 0 ? g_b2s_cr_v4 : 1;
  if (v4 && c % CR_SLAB == 0 && vec_of(c, x, g, mean, invstd) == 4 && vec_of(c, gamma, beta) == 4) {
    // slabs x row shares ~ `cap` (default one) CTA per SM, >= 4 rows per thread
    const int slabs = c / CR_SLAB;
    const int64_t ctas = (int64_t)B2S_NUM_SMS * (g_b2s_cr_cap > 0 ? g_b2s_cr_cap : 1);
    int64_t shares = (ctas + slabs - 1) / slabs;
    const int64_t max_shares = ceil_div64(n, (int64_t)CR_RPB * 4);
    if (shares > max_shares) shares = max_shares;
    if (shares < 1) shares = 1;
    constexpr int U = MODE == 2 ? 4 : 8;
    colreduce_v4_kernel<MODE, ACC, U><<<dim3((unsigned)shares, (unsigned)slabs), CR_THREADS, 0, st>>>(
        x, g, mean, invstd, gamma, beta, n, n_dev, c, act, ws, fin);
    return fin.kind != 0;
  }
  colreduce_kernel<MODE, ACC><<<colgrid(n, c), PW_THREADS, 0, st>>>(x, g, mean, invstd, gamma, beta, n, n_dev, c, act,
                                                                    ws);
  return false;
}

}  // namespace

// operand form written by the `*_tf32` outputs: split-bf16 blocks need whole 32-channel blocks (what the tensor-core
// convolutions require of their operands anyway); other widths keep TF32 values
static inline int operand_mode(int c) { return (b2s_precise() && c % 32 == 0) ? 1 : 0; }

// ================================================================= C ABI ======================
extern "C" int32_t b2s_maxpool_fwd(const float* x, const int32_t* nbr, int64_t n_out, const int32_t* n_out_dev,
                                   int32_t c, int32_t k3, float* y, int32_t* arg, float* y_tf32,
                                   b2s_stream_t stream) {
  B2S_CHECK_ARG(n_out >= 0 && c > 0 && k3 > 0, "bad sizes");
  if (n_out == 0) return B2S_OK;
  B2S_CHECK_ARG(x && nbr && y && arg, "null pointer");
  cudaStream_t st = as_stream(stream);
  static int batched = -1;
  if (batched < 0) {
    const char* e = getenv("B2S_POOL_BATCH");
    batched = e ? atoi(e) : 2;                 // 0: one offset after the other, 1: nine at a time, 2: warp per 32 rows
  }
  if (batched >= 2 && vec_of(c, x, y, y_tf32) == 4 && k3 <= 27 && (c == 32 || c == 64 || c == 128) &&
      (reinterpret_cast<uintptr_t>(arg) & 15) == 0) {
    const int64_t blocks64 = (n_out + MP_WARPS * 32 - 1) / (MP_WARPS * 32);
    const int blocks = (int)(blocks64 < (int64_t)B2S_NUM_SMS * 8 ? blocks64 : (int64_t)B2S_NUM_SMS * 8);
    if (c == 32)
      maxpool_fwd_warp_kernel<8><<<blocks, MP_WARPS * 32, 0, st>>>(x, nbr, n_out, n_out_dev, c, k3, y, arg, y_tf32,
                                                                   operand_mode(c));
    else if (c == 64)
      maxpool_fwd_warp_kernel<16><<<blocks, MP_WARPS * 32, 0, st>>>(x, nbr, n_out, n_out_dev, c, k3, y, arg, y_tf32,
                                                                    operand_mode(c));
    else
      maxpool_fwd_warp_kernel<32><<<blocks, MP_WARPS * 32, 0, st>>>(x, nbr, n_out, n_out_dev, c, k3, y, arg, y_tf32,
                                                                    operand_mode(c));
    B2S_LAUNCH_CHECK();
    return B2S_OK;
  }
  if (vec_of(c, x, y, y_tf32) == 4)
    maxpool_fwd_kernel<4><<<rows_grid(n_out, c, 4), PW_THREADS, 0, st>>>(x, nbr, n_out, n_out_dev, c, k3, y, arg,
                                                                         y_tf32, operand_mode(c), batched);
  else
    maxpool_fwd_kernel<1><<<rows_grid(n_out, c, 1), PW_THREADS, 0, st>>>(x, nbr, n_out, n_out_dev, c, k3, y, arg,
                                                                         y_tf32, operand_mode(c), batched);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_sumpool(const float* x, const int32_t* nbr, const float* pre_scale, const float* post_scale,
                               int64_t n_out, const int32_t* n_out_dev, int32_t c, int32_t k3, float* y,
                               b2s_stream_t stream) {
  B2S_CHECK_ARG(n_out >= 0 && c > 0 && k3 > 0, "bad sizes");
  if (n_out == 0) return B2S_OK;
  B2S_CHECK_ARG(x && nbr && y, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (vec_of(c, x, y) == 4)
    sumpool_kernel<4><<<rows_grid(n_out, c, 4), PW_THREADS, 0, st>>>(x, nbr, pre_scale, post_scale, n_out, n_out_dev, c,
                                                                     k3, y);
  else
    sumpool_kernel<1><<<rows_grid(n_out, c, 1), PW_THREADS, 0, st>>>(x, nbr, pre_scale, post_scale, n_out, n_out_dev, c,
                                                                     k3, y);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_nbr_inv_counts(const int32_t* nbr, int32_t k3, int64_t n, const int32_t* n_dev, float* inv,
                                      b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && k3 > 0, "bad sizes");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(nbr && inv, "null pointer");
  nbr_inv_counts_kernel<<<grid_for(n, PW_THREADS), PW_THREADS, 0, as_stream(stream)>>>(nbr, k3, n, n_dev, inv);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_maxpool_bwd(const float* gy, const int32_t* arg, int64_t n_in, int64_t n_out,
                                   const int32_t* n_out_dev, int32_t c, float* gx, b2s_stream_t stream) {
  B2S_CHECK_ARG(n_in >= 0 && n_out >= 0 && c > 0, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (n_in > 0) {
    B2S_CHECK_ARG(gx, "null pointer");
    B2S_CUDA(cudaMemsetAsync(gx, 0, n_in * c * sizeof(float), st));
  }
  if (n_out == 0) return B2S_OK;
  B2S_CHECK_ARG(gy && arg, "null pointer");
  maxpool_bwd_kernel<<<grid_for(n_out * c, PW_THREADS), PW_THREADS, 0, st>>>(gy, arg, n_out, n_out_dev, c, gx);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_batch_counts(const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                    const int32_t* n_dev, int32_t num_batches, int32_t* counts,
                                    b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && num_batches > 0 && counts && row_batch_stride > 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(counts, 0, num_batches * sizeof(int), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(row_batch, "null pointer");
  batch_counts_kernel<<<(unsigned)ceil_div64(ceil_div64(n, 64), PW_THREADS), PW_THREADS, 0, st>>>(
      row_batch, row_batch_stride, n, n_dev, num_batches, counts);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_segment_sum(const float* x, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                   const int32_t* n_dev, int32_t c, int32_t num_batches, const float* scale,
                                   float* y, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && num_batches > 0 && y && row_batch_stride > 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(y, 0, (size_t)num_batches * c * sizeof(float), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && row_batch, "null pointer");
  if (c % CR_SLAB == 0 && vec_of(c, x, y) == 4) {
    const int slabs = c / CR_SLAB;
    int64_t chunks = (2LL * B2S_NUM_SMS + slabs - 1) / slabs;
    const int64_t max_chunks = ceil_div64(n, (int64_t)CR_RPB * 4);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t rows = ceil_div64(n, chunks);
    rows = ceil_div64(rows, CR_RPB) * CR_RPB;
    chunks = ceil_div64(n, rows);
    segment_sum_v4_kernel<<<dim3((unsigned)chunks, (unsigned)slabs), CR_THREADS, 0, st>>>(
        x, row_batch, row_batch_stride, n, n_dev, c, num_batches, rows, scale, y);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
  }
  segment_sum_kernel<false><<<colgrid(n, c), PW_THREADS, 0, st>>>(x, nullptr, row_batch, row_batch_stride, n, n_dev,
                                                                  c, num_batches, y);
  if (scale) scale_rows_kernel<<<grid_for((int64_t)num_batches * c, PW_THREADS), PW_THREADS, 0, st>>>(y, scale, num_batches, c);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_segment_max(const float* x, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                   const int32_t* n_dev, int32_t c, int32_t num_batches, float* y, int32_t* arg,
                                   b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && num_batches >= 0 && row_batch_stride > 0, "bad arguments");
  if (num_batches == 0) return B2S_OK;
  B2S_CHECK_ARG((x || n == 0) && (row_batch || n == 0) && y && arg, "null pointer");
  segment_max_kernel<<<dim3((unsigned)num_batches, (unsigned)((c + 63) / 64)), 256, 0, as_stream(stream)>>>(
      x, row_batch, row_batch_stride, n, n_dev, c, y, arg);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_segment_max_bwd(const float* gy, const int32_t* arg, int64_t n, int32_t c, int32_t num_batches,
                                       float* gx, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && num_batches >= 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (n > 0) B2S_CUDA(cudaMemsetAsync(gx, 0, (size_t)n * c * sizeof(float), st));
  if (n == 0 || num_batches == 0) return B2S_OK;
  B2S_CHECK_ARG(gy && arg && gx, "null pointer");
  const int64_t total = (int64_t)num_batches * c;
  segment_max_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(gy, arg, total, c, gx);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_segment_bcast(const float* y, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                     const int32_t* n_dev, int32_t c, const float* scale, float* x_out,
                                     b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && row_batch_stride > 0, "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(y && row_batch && x_out, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (vec_of(c, y, x_out) == 4)
    segment_bcast_kernel<4><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(y, row_batch, row_batch_stride, n, n_dev, c,
                                                                       scale, x_out);
  else
    segment_bcast_kernel<1><<<rows_grid(n, c, 1), PW_THREADS, 0, st>>>(y, row_batch, row_batch_stride, n, n_dev, c,
                                                                       scale, x_out);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

static void launch_bcast_mul(const float* x, const float* y, const int32_t* rb, int32_t stride, int64_t n,
                             const int32_t* n_dev, int32_t c, int32_t y_c, float* out, cudaStream_t st) {
  if (vec_of(c, x, out, y_c == 1 ? nullptr : y) == 4)
    bcast_mul_kernel<4><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(x, y, rb, stride, n, n_dev, c, y_c, out);
  else
    bcast_mul_kernel<1><<<rows_grid(n, c, 1), PW_THREADS, 0, st>>>(x, y, rb, stride, n, n_dev, c, y_c, out);
}

extern "C" int32_t b2s_bcast_mul_fwd(const float* x, const float* y, const int32_t* row_batch,
                                     int32_t row_batch_stride, int64_t n, const int32_t* n_dev, int32_t c,
                                     int32_t y_c, float* out, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && (y_c == c || y_c == 1) && row_batch_stride > 0, "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && y && row_batch && out, "null pointer");
  launch_bcast_mul(x, y, row_batch, row_batch_stride, n, n_dev, c, y_c, out, as_stream(stream));
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bcast_mul_bwd(const float* g, const float* x, const float* y, const int32_t* row_batch,
                                     int32_t row_batch_stride, int64_t n, const int32_t* n_dev, int32_t c,
                                     int32_t num_batches, float* gx, float* gy, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && num_batches > 0 && row_batch_stride > 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (gy) B2S_CUDA(cudaMemsetAsync(gy, 0, (size_t)num_batches * c * sizeof(float), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(g && row_batch, "null pointer");
  if (gx) {
    B2S_CHECK_ARG(y, "null pointer");
    launch_bcast_mul(g, y, row_batch, row_batch_stride, n, n_dev, c, c, gx, st);
  }
  if (gy) {
    B2S_CHECK_ARG(x, "null pointer");
    segment_sum_kernel<true><<<colgrid(n, c), PW_THREADS, 0, st>>>(g, x, row_batch, row_batch_stride, n, n_dev, c,
                                                                   num_batches, gy);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_colsum(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* out,
                              b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && out, "bad arguments");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(out, 0, c * sizeof(float), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x, "null pointer");
  launch_colreduce<0, float>(x, nullptr, nullptr, nullptr, nullptr, nullptr, n, n_dev, c, 0, out, st);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bn_stats(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float eps,
                                float momentum, float* running_mean, float* running_var, double* stats_ws,
                                float* mean, float* invstd, b2s_stream_t stream) {
  B2S_CHECK_ARG(n > 0 && c > 0, "n > 0 and c > 0");
  B2S_CHECK_ARG(x && stats_ws && mean && invstd, "null pointer");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(stats_ws, 0, (2 * (size_t)c + 1) * sizeof(double), st));
  CrFinal fin{1, eps, momentum, running_mean, running_var, mean, invstd};
  if (!launch_colreduce<1, double>(x, nullptr, nullptr, nullptr, nullptr, nullptr, n, n_dev, c, 0, stats_ws, st, fin))
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(stats_ws, n, n_dev, c, eps, momentum, running_mean,
                                                        running_var, mean, invstd);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// ---- batch-norm statistics from a convolution's partial rows ----------------------------------------------------
// col_stats (float): P_cap = ceil(n / 128) partial rows of [sums (c) | sums of squares (c)], then one header float =
// the number of out rows a partial row covers (128 or 256: the row tile of the kernel that wrote them).  The
// convolution epilogues write one partial row per row tile with plain stores -- no atomics: thousands of same-address
// fp64 atomics serialise at ~20 ns each in L2, which cost more than the statistics pass they replaced.
// Fallback producer (split-K launches, SIMT path): 128 rows per block, thread = channel.
__global__ void __launch_bounds__(256) col_partials_kernel(const float* __restrict__ x, int64_t n,
                                                           const int* __restrict__ n_dev, int c,
                                                           float* __restrict__ col_stats) {
  const int64_t pitch = n;
  n = b2s_rows(n, n_dev);
  if (blockIdx.x == 0 && threadIdx.x == 0) col_stats[((pitch + 127) / 128) * 2 * c] = 128.f;
  const int64_t r0 = (int64_t)blockIdx.x * 128, r1 = min(r0 + 128, n);
  if (r0 >= n) return;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      const float v = x[r * c + ch];
      s1 += v;
      s2 += v * v;
    }
    col_stats[(int64_t)blockIdx.x * 2 * c + ch] = s1;
    col_stats[(int64_t)blockIdx.x * 2 * c + c + ch] = s2;
  }
}

void b2s_launch_col_partials(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* col_stats,
                             cudaStream_t st) {
  if (n <= 0) return;
  col_partials_kernel<<<(unsigned)ceil_div64(n, 128), 256, 0, st>>>(x, n, n_dev, c, col_stats);
}

// block = 32 channels x 32 partial-row lanes: fp64 sums over the live partial rows, then the finalisation of b2s_bn_stats
__global__ void __launch_bounds__(1024) bn_finalize_partials_kernel(const float* __restrict__ col_stats, int64_t n,
                                                                    const int* __restrict__ n_dev, int c, float eps,
                                                                    float momentum, float* __restrict__ running_mean,
                                                                    float* __restrict__ running_var,
                                                                    float* __restrict__ mean,
                                                                    float* __restrict__ invstd) {
  __shared__ double sh[2][32][33];
  const int64_t pitch = n;
  n = b2s_rows(n, n_dev);
  const int lane_c = threadIdx.x & 31, lane_p = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane_c;
  int rpt = (int)col_stats[((pitch + 127) / 128) * 2 * c];
  if (rpt <= 0) rpt = 128;
  const int64_t P = (n + rpt - 1) / rpt;
  double s1 = 0.0, s2 = 0.0;
  if (ch < c) {
#pragma unroll 4
    for (int64_t p = lane_p; p < P; p += 32) {
      s1 += (double)__ldg(&col_stats[p * 2 * c + ch]);
      s2 += (double)__ldg(&col_stats[p * 2 * c + c + ch]);
    }
  }
  sh[0][lane_p][lane_c] = s1;
  sh[1][lane_p][lane_c] = s2;
  __syncthreads();
  if (lane_p == 0 && ch < c) {
#pragma unroll
    for (int j = 1; j < 32; ++j) {
      s1 += sh[0][j][lane_c];
      s2 += sh[1][j][lane_c];
    }
    const double dn = n > 0 ? (double)n : 1.0;
    const double m = s1 / dn;
    double var = s2 / dn - m * m;
    if (var < 0.0) var = 0.0;
    mean[ch] = (float)m;
    invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
    if (running_var) {
      const double unb = n > 1 ? var * dn / (dn - 1.0) : var;
      running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
    }
  }
}

extern "C" int64_t b2s_conv_col_stats_elems(int64_t n_out, int32_t c_out) {
  if (n_out < 0 || c_out <= 0) return -1;
  return ceil_div64(n_out > 0 ? n_out : 1, 128) * 2 * c_out + 4;
}

extern "C" int32_t b2s_bn_finalize(const float* col_stats, int64_t n, const int32_t* n_dev, int32_t c, float eps,
                                   float momentum, float* running_mean, float* running_var, float* mean,
                                   float* invstd, b2s_stream_t stream) {
  B2S_CHECK_ARG(n > 0 && c > 0 && col_stats && mean && invstd, "bad arguments");
  bn_finalize_partials_kernel<<<(c + 31) / 32, 1024, 0, as_stream(stream)>>>(col_stats, n, n_dev, c, eps, momentum,
                                                                            running_mean, running_var, mean, invstd);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bn_apply(const float* x, const float* mean, const float* invstd, const float* gamma,
                                const float* beta, int64_t n, const int32_t* n_dev, int32_t c, int32_t act, float* y,
                                float* y_tf32, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && (act == 0 || act == 1), "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && mean && invstd && (y || y_tf32), "null pointer");
  cudaStream_t st = as_stream(stream);
  if (vec_of(c, x, y, mean, invstd) == 4 && vec_of(c, gamma, beta, y_tf32) == 4) {
    if (pw_unroll() > 1 && n * c >= PW_UNROLL_MIN)
      bn_apply_kernel<4, 4><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(x, mean, invstd, gamma, beta, n, n_dev, c, act, y,
                                                                       y_tf32, operand_mode(c));
    else
      bn_apply_kernel<4, 1><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(x, mean, invstd, gamma, beta, n, n_dev, c, act, y,
                                                                       y_tf32, operand_mode(c));
  } else {
    bn_apply_kernel<1, 1><<<rows_grid(n, c, 1), PW_THREADS, 0, st>>>(x, mean, invstd, gamma, beta, n, n_dev, c, act, y,
                                                                     y_tf32, operand_mode(c));
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bn_bwd_reduce(const float* gy, const float* x, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, int64_t n, const int32_t* n_dev,
                                     int32_t c, int32_t act, double* stats_ws, float* sums, b2s_stream_t stream) {
  B2S_CHECK_ARG(n > 0 && c > 0 && (act == 0 || act == 1), "bad arguments");
  B2S_CHECK_ARG(gy && x && mean && invstd && stats_ws && sums, "null pointer");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(stats_ws, 0, (2 * (size_t)c + 1) * sizeof(double), st));
  CrFinal fin{2, 0.f, 0.f, nullptr, nullptr, sums, nullptr};
  if (!launch_colreduce<2, double>(x, gy, mean, invstd, gamma, beta, n, n_dev, c, act, stats_ws, st, fin))
    double_to_float_kernel<<<(2 * c + 127) / 128, 128, 0, st>>>(stats_ws, sums, 2 * c);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// out[c] = sum over `rows` rows of x[rows, c]: the partial rows of b2s_bn_bwd_apply's gx_colsum.  Block = 32 channels x
// 32 row lanes (a thousand rows x 64 channels is a ~3 us job; torch's generic reduction took 15 us for it).
__global__ void __launch_bounds__(1024) sum_rows_kernel(const float* __restrict__ x, int64_t rows, int c,
                                                        float* __restrict__ out) {
  __shared__ float sh[32][33];
  const int lane_c = threadIdx.x & 31, lane_r = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane_c;
  float s = 0.f;
  if (ch < c) {
#pragma unroll 4
    for (int64_t r = lane_r; r < rows; r += 32) s += __ldg(&x[r * c + ch]);
  }
  sh[lane_r][lane_c] = s;
  __syncthreads();
  if (lane_r == 0 && ch < c) {
#pragma unroll
    for (int j = 1; j < 32; ++j) s += sh[j][lane_c];
    out[ch] = s;
  }
}

extern "C" int32_t b2s_sum_rows(const float* x, int64_t rows, int32_t c, float* out, b2s_stream_t stream) {
  B2S_CHECK_ARG(rows >= 0 && c > 0 && out && (x || rows == 0), "bad arguments");
  sum_rows_kernel<<<(c + 31) / 32, 1024, 0, as_stream(stream)>>>(x, rows, c, out);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// rows of the gx_colsum output of b2s_bn_bwd_apply: one partial row per block of its launch
extern "C" int64_t b2s_bn_bwd_colsum_rows(int64_t n, int32_t c) {
  if (n < 0 || c <= 0) return -1;
  return c % 4 == 0 ? rows_grid(n > 0 ? n : 1, c, 4) : 1;
}

extern "C" int32_t b2s_bn_bwd_apply(const float* gy, const float* x, const float* mean, const float* invstd,
                                    const float* gamma, const float* beta, const float* sums, int64_t n,
                                    const int32_t* n_dev, int32_t c, int32_t act, int32_t training, float* gx,
                                    float* gx_tf32, float* gx_colsum, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && (act == 0 || act == 1), "bad arguments");
  cudaStream_t st = as_stream(stream);
  const int64_t cs_rows = gx_colsum ? b2s_bn_bwd_colsum_rows(n, c) : 0;
  if (n == 0) {
    if (gx_colsum) B2S_CUDA(cudaMemsetAsync(gx_colsum, 0, (size_t)cs_rows * c * sizeof(float), st));
    return B2S_OK;
  }
  B2S_CHECK_ARG(gy && x && mean && invstd && gx && (sums || !training), "null pointer");
  bool colsum_done = false;
  if (vec_of(c, gy, x, gx, mean) == 4 && vec_of(c, invstd, gamma, beta, sums) == 4 && vec_of(c, gx_tf32) == 4) {
    const int cv = c / 4, tpr = cv < PW_THREADS ? cv : PW_THREADS;
    const int rpb = PW_THREADS / tpr;
    // the fused column sums need all threads active, equal channel passes and a power-of-two tree over the row lanes
    if (gx_colsum && PW_THREADS % tpr == 0 && cv % tpr == 0 && (rpb & (rpb - 1)) == 0 &&
        (reinterpret_cast<uintptr_t>(gx_colsum) & 3) == 0) {
      if (pw_unroll() > 1 && n * c >= PW_UNROLL_MIN)
        bn_bwd_apply_kernel<4, true, 4><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(
            gy, x, mean, invstd, gamma, beta, sums, n, n_dev, c, act, training, gx, gx_tf32, operand_mode(c), gx_colsum);
      else
        bn_bwd_apply_kernel<4, true, 1><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(
            gy, x, mean, invstd, gamma, beta, sums, n, n_dev, c, act, training, gx, gx_tf32, operand_mode(c), gx_colsum);
      colsum_done = true;
    } else {
      bn_bwd_apply_kernel<4, false, 1><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(
          gy, x, mean, invstd, gamma, beta, sums, n, n_dev, c, act, training, gx, gx_tf32, operand_mode(c), nullptr);
    }
  } else {
    bn_bwd_apply_kernel<1, false, 1><<<rows_grid(n, c, 1), PW_THREADS, 0, st>>>(
        gy, x, mean, invstd, gamma, beta, sums, n, n_dev, c, act, training, gx, gx_tf32, operand_mode(c), nullptr);
  }
  if (gx_colsum && !colsum_done) {     // shapes the fused form does not cover: the separate column reduction into row 0
    B2S_CUDA(cudaMemsetAsync(gx_colsum, 0, (size_t)cs_rows * c * sizeof(float), st));
    launch_colreduce<0, float>(gx, nullptr, nullptr, nullptr, nullptr, nullptr, n, n_dev, c, 0, gx_colsum, st);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

template <int OP>
static int32_t launch_flat(const float* a, const float* b, int64_t n, const int32_t* n_dev, int32_t c, float* o0,
                           float* o1, cudaStream_t st, float* o2 = nullptr) {
  if (vec_of(c, a, b, o0, o1) == 4 && vec_of(c, o2) == 4)
    flat_kernel<4, OP><<<grid_for(n * c / 4, PW_THREADS), PW_THREADS, 0, st>>>(a, b, n, n_dev, c, o0, o1, o2,
                                                                               operand_mode(c));
  else
    flat_kernel<1, OP><<<grid_for(n * c, PW_THREADS), PW_THREADS, 0, st>>>(a, b, n, n_dev, c, o0, o1, o2,
                                                                               operand_mode(c));
  return 0;
}

// internal: rounding pass used by the convolution entry points when the caller did not pre-round
void b2s_launch_round_tf32(const float* in, int64_t n, const int32_t* n_dev, int32_t c, float* out, cudaStream_t st) {
  launch_flat<3>(in, nullptr, n, n_dev, c, out, nullptr, st);
}

extern "C" int32_t b2s_round_tf32(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* y,
                                  b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0, "n >= 0 and c > 0");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && y, "null pointer");
  launch_flat<3>(x, nullptr, n, n_dev, c, y, nullptr, as_stream(stream));
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_gelu_fwd(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* y,
                                b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0, "n >= 0 and c > 0");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && y, "null pointer");
  launch_flat<0>(x, nullptr, n, n_dev, c, y, nullptr, as_stream(stream));
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_add_gelu_fwd(const float* a, const float* b, int64_t n, const int32_t* n_dev, int32_t c,
                                    float* sum, float* y, float* y_tf32, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0, "n >= 0 and c > 0");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(a && b && sum && (y || y_tf32), "null pointer");
  launch_flat<2>(a, b, n, n_dev, c, sum, y, as_stream(stream), y_tf32);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_gelu_bwd(const float* gy, const float* x, int64_t n, const int32_t* n_dev, int32_t c,
                                float* gx, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0, "n >= 0 and c > 0");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(gy && x && gx, "null pointer");
  launch_flat<1>(gy, x, n, n_dev, c, gx, nullptr, as_stream(stream));
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// ---------------------------------------------------------------- fused SE block tail (C ABI) ---
extern "C" int32_t b2s_se_gate_fwd(const float* pooled, const float* w1, const float* b1, const float* w2,
                                   const float* b2, const float* keep, int32_t num_batches, int32_t c, int32_t h,
                                   float* h_pre, float* gate, float* gate_eff, b2s_stream_t stream) {
  B2S_CHECK_ARG(num_batches > 0 && c > 0 && h > 0 && (c + h) * sizeof(float) <= 48 * 1024, "bad sizes");
  B2S_CHECK_ARG(pooled && w1 && w2 && h_pre && gate && gate_eff, "null pointer");
  cudaStream_t st = as_stream(stream);
  se_hidden_kernel<<<dim3(num_batches, (h + 7) / 8), 256, 0, st>>>(pooled, w1, b1, c, h, h_pre);
  se_gate_fwd_kernel<<<dim3(num_batches, (c + SE_CHUNK - 1) / SE_CHUNK), 256, h * sizeof(float), st>>>(
      h_pre, w2, b2, keep, c, h, gate, gate_eff);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_se_gate_bwd(const float* g_gate_eff, const float* keep, const float* gate, const float* h_pre,
                                   const float* pooled, const float* w1, const float* w2, const float* inv_count,
                                   int32_t num_batches, int32_t c, int32_t h, float* gz2, float* gh_pre,
                                   float* g_pooled, float* gw1, float* gb1, float* gw2, float* gb2,
                                   b2s_stream_t stream) {
  B2S_CHECK_ARG(num_batches > 0 && c > 0 && h > 0 && (c + 256) * sizeof(float) <= 48 * 1024 &&
                    h * sizeof(float) <= 48 * 1024, "bad sizes");
  B2S_CHECK_ARG(g_gate_eff && gate && h_pre && pooled && w1 && w2 && gz2 && gh_pre && g_pooled && gw1 && gw2,
                "null pointer");
  cudaStream_t st = as_stream(stream);
  se_gate_bwd_kernel<<<num_batches, 256, (c + 256) * sizeof(float), st>>>(g_gate_eff, keep, gate, h_pre, w2,
                                                                          num_batches, c, h, gz2, gh_pre);
  se_pooled_grad_kernel<<<dim3(num_batches, (c + 255) / 256), 256, h * sizeof(float), st>>>(gh_pre, w1, inv_count, c, h,
                                                                                          g_pooled);
  se_param_grad_kernel<<<grid_for(2 * (int64_t)h * c + h + c, 256), 256, 0, st>>>(
      gz2, gh_pre, gh_pre + (int64_t)num_batches * h, pooled, num_batches, c, h, gw1, gb1, gw2, gb2);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_gated_add_gelu_fwd(const float* u, const float* gate_eff, const float* res,
                                          const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                          const int32_t* n_dev, int32_t c, float* sum, float* y, float* y_tf32,
                                          b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && row_batch_stride > 0, "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(u && gate_eff && res && row_batch && sum && (y || y_tf32), "null pointer");
  cudaStream_t st = as_stream(stream);
  if (vec_of(c, u, res, sum, gate_eff) == 4 && vec_of(c, y, y_tf32) == 4)
    gated_add_gelu_fwd_kernel<4><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(u, gate_eff, res, row_batch,
                                                                            row_batch_stride, n, n_dev, c, sum, y, y_tf32,
                                                                            operand_mode(c));
  else
    gated_add_gelu_fwd_kernel<1><<<rows_grid(n, c, 1), PW_THREADS, 0, st>>>(u, gate_eff, res, row_batch,
                                                                            row_batch_stride, n, n_dev, c, sum, y, y_tf32,
                                                                            operand_mode(c));
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_gated_add_gelu_bwd(const float* gy, const float* sum, const float* u, const float* gate_eff,
                                          const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                          const int32_t* n_dev, int32_t c, int32_t num_batches, float* g_res,
                                          float* g_u, float* g_gate_eff, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && c % CR_SLAB == 0 && num_batches > 0 && row_batch_stride > 0,
                "c must be a multiple of 64");
  B2S_CHECK_ARG(g_gate_eff, "null pointer");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(g_gate_eff, 0, (size_t)num_batches * c * sizeof(float), st));
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(gy && sum && u && gate_eff && row_batch && g_res && g_u, "null pointer");
  B2S_CHECK_ARG(vec_of(c, gy, sum, u, gate_eff) == 4 && vec_of(c, g_res, g_u, g_gate_eff) == 4, "16-byte aligned rows");
  const int slabs = c / CR_SLAB;
  int64_t chunks = (2LL * B2S_NUM_SMS + slabs - 1) / slabs;       // ~two CTAs per SM over all slabs
  const int64_t max_chunks = ceil_div64(n, (int64_t)CR_RPB * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t rows = ceil_div64(n, chunks);
  rows = ceil_div64(rows, CR_RPB) * CR_RPB;
  chunks = ceil_div64(n, rows);
  gated_add_gelu_bwd_kernel<<<dim3((unsigned)chunks, (unsigned)slabs), CR_THREADS, 0, st>>>(
      gy, sum, u, gate_eff, row_batch, row_batch_stride, n, n_dev, c, num_batches, rows, g_res, g_u, g_gate_eff);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_bcast_add_(float* x, const float* y, const int32_t* row_batch, int32_t row_batch_stride,
                                  int64_t n, const int32_t* n_dev, int32_t c, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && c > 0 && row_batch_stride > 0, "bad arguments");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(x && y && row_batch, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (vec_of(c, x, y) == 4)
    bcast_add_kernel<4><<<rows_grid(n, c, 4), PW_THREADS, 0, st>>>(x, y, row_batch, row_batch_stride, n, n_dev, c);
  else
    bcast_add_kernel<1><<<rows_grid(n, c, 1), PW_THREADS, 0, st>>>(x, y, row_batch, row_batch_stride, n, n_dev, c);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}
