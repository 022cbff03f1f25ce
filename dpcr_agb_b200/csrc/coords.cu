// coords.cu -- integer stages of the hot path: voxel quantisation (a1), coordinate hash and strided
// maps (a2, a3), kernel maps (a4).  All HBM / L2-latency bound; no tensor cores here by design.
//
// Reference call sites (R: = /root/reference/torch-points3d/torch_points3d/):
//   a1  R:core/data_transform/grid_transform.py:112-128
//   a2  R:models/instance/minkowski.py:74
//   a3/a4  implicit in R:modules/MinkowskiEngine/SENet.py:53,94-97 and resnet_block.py:48-54
#include "common.cuh"
#include <limits.h>

// =============================================================================================
// exclusive scan of int32 (three small kernels; n is at most a few 10^7 here)
// =============================================================================================
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across a block of SCAN_THREADS; returns exclusive prefix,
// writes the block total to *total
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    int si = warp_incl_scan(s);
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = si - s;
    if (lane == SCAN_THREADS / 32 - 1) block_total = si;
  }
  __syncthreads();
  int res = incl - v + warp_sums[wid];
  *total = block_total;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int* __restrict__ in, int64_t n,
                                                                   const int* __restrict__ n_dev,
                                                                   int* __restrict__ block_sums) {
  n = b2s_rows(n, n_dev);
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    int64_t i = base + j;
    if (i < n) s += in[i];
  }
  int total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(int* __restrict__ block_sums, int nb,
                                                                       int* __restrict__ total_out) {
  int carry = 0;
  for (int base = 0; base < nb; base += SCAN_THREADS) {
    int i = base + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int total;
    int ex = block_excl_scan(v, &total);
    if (i < nb) block_sums[i] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                                  int64_t n, const int* __restrict__ n_dev,
                                                                  const int* __restrict__ block_sums) {
  n = b2s_rows(n, n_dev);
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    int64_t i = base + j;
    v[j] = i < n ? in[i] : 0;
    s += v[j];
  }
  int total;
  int ex = block_excl_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    int64_t i = base + j;
    if (i < n) out[i] = ex;
    ex += v[j];
  }
}

// in-place capable exclusive scan; total (device int) may be null
int launch_exclusive_scan(const int* in, int* out, int64_t n, const int* n_dev, int* total_dev, void* ws,
                          cudaStream_t st) {
  if (n <= 0) {
    if (total_dev) cudaMemsetAsync(total_dev, 0, sizeof(int), st);
    return 0;
  }
  int nb = (int)ceil_div64(n, SCAN_TILE);
  int* block_sums = reinterpret_cast<int*>(ws);
  scan_reduce_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, n, n_dev, block_sums);
  scan_block_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(block_sums, nb, total_dev);
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, out, n, n_dev, block_sums);
  return 0;
}

}  // namespace

extern "C" int64_t b2s_scan_workspace_bytes(int64_t n) {
  return (ceil_div64(n > 0 ? n : 1, SCAN_TILE) + 1) * (int64_t)sizeof(int);
}

// =============================================================================================
// (a1) voxel quantisation
// =============================================================================================
namespace {

__global__ void init_bounds_kernel(int* bounds) {
  if (threadIdx.x < 3) bounds[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) bounds[threadIdx.x] = INT_MIN;
}

// q = rint(pos / size) in fp32 (IEEE divide, round-half-even) -- grid_transform.py:116
__global__ void __launch_bounds__(256) quantize_points_kernel(const float* __restrict__ pos, int64_t n,
                                                              const int* __restrict__ n_dev, float size,
                                                              int* __restrict__ q, int* __restrict__ bounds) {
  n = b2s_rows(n, n_dev);
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  const int64_t total = n * 3;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    float v = rintf(__fdiv_rn(pos[e], size));
    int iv = (int)v;
    q[e] = iv;
    int d = (int)(e % 3);
    // branch-free select keeps lo/hi in registers
    lo[0] = d == 0 ? min(lo[0], iv) : lo[0];
    lo[1] = d == 1 ? min(lo[1], iv) : lo[1];
    lo[2] = d == 2 ? min(lo[2], iv) : lo[2];
    hi[0] = d == 0 ? max(hi[0], iv) : hi[0];
    hi[1] = d == 1 ? max(hi[1], iv) : hi[1];
    hi[2] = d == 2 ? max(hi[2], iv) : hi[2];
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (lo[d] != INT_MAX) atomicMin(&bounds[d], lo[d]);
      if (hi[d] != INT_MIN) atomicMax(&bounds[3 + d], hi[d]);
    }
  }
}

struct QBox {
  int lo[3];
  int dim[3];  // nx, ny, nz
};

__device__ __forceinline__ int64_t cell_of(const int* __restrict__ q, const int* __restrict__ plot, int64_t p,
                                           const QBox& box) {
  int x = q[3 * p] - box.lo[0], y = q[3 * p + 1] - box.lo[1], z = q[3 * p + 2] - box.lo[2];
  if ((unsigned)x >= (unsigned)box.dim[0] || (unsigned)y >= (unsigned)box.dim[1] || (unsigned)z >= (unsigned)box.dim[2])
    return -1;
  return (((int64_t)plot[p] * box.dim[2] + z) * box.dim[1] + y) * box.dim[0] + x;
}

__global__ void __launch_bounds__(256) quantize_mark_kernel(const int* __restrict__ q, const int* __restrict__ plot,
                                                            int64_t n, const int* __restrict__ n_dev, int num_plots,
                                                            QBox box, unsigned* __restrict__ bitmap,
                                                            int* __restrict__ oob) {
  n = b2s_rows(n, n_dev);
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = ((unsigned)plot[p] < (unsigned)num_plots) ? cell_of(q, plot, p, box) : -1;
    if (cell < 0) {
      *oob = 1;
      continue;
    }
    atomicOr(&bitmap[cell >> 5], 1u << (cell & 31));
  }
}

__global__ void __launch_bounds__(256) popc_kernel(const unsigned* __restrict__ bitmap, int64_t words,
                                                   int* __restrict__ pc) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x)
    pc[w] = __popc(bitmap[w]);
}

__global__ void quantize_total_kernel(const int* __restrict__ oob, int* __restrict__ num_voxels) {
  if (*oob) *num_voxels = -1;
}

__global__ void __launch_bounds__(256) quantize_rep_kernel(const int* __restrict__ q, const int* __restrict__ plot,
                                                           const int* __restrict__ order, int64_t n,
                                                           const int* __restrict__ n_dev, QBox box,
                                                           const unsigned* __restrict__ bitmap,
                                                           const int* __restrict__ prefix, int64_t rep_cap,
                                                           int* __restrict__ rep) {
  n = b2s_rows(n, n_dev);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = order ? order[j] : j;
    int64_t cell = cell_of(q, plot, p, box);
    if (cell < 0) continue;
    int64_t w = cell >> 5;
    unsigned bit = (unsigned)(cell & 31);
    int r = prefix[w] + __popc(bitmap[w] & ((1u << bit) - 1u));
    if (r < rep_cap) atomicMax(&rep[r], (int)j);  // last position in the shuffled order wins
  }
}

__global__ void __launch_bounds__(256) quantize_emit_kernel(const int* __restrict__ q, const int* __restrict__ plot,
                                                            const int* __restrict__ order, int64_t m,
                                                            const int* __restrict__ m_dev,
                                                            int* __restrict__ out_coords, int* __restrict__ rep_src) {
  m = b2s_rows(m, m_dev);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
    int j = rep_src[r];
    int p = order ? order[j] : j;
    int4 c = make_int4(plot[p], q[3 * (int64_t)p], q[3 * (int64_t)p + 1], q[3 * (int64_t)p + 2]);
    reinterpret_cast<int4*>(out_coords)[r] = c;
    rep_src[r] = p;
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ in, const int* __restrict__ idx,
                                                          int64_t m, const int* __restrict__ m_dev, int c,
                                                          float* __restrict__ out) {
  m = b2s_rows(m, m_dev);
  const int64_t total = m * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e / c;
    int ch = (int)(e - r * c);
    out[e] = in[(int64_t)idx[r] * c + ch];
  }
}

struct QWs {
  unsigned* bitmap;
  int* prefix;
  int* oob;
  void* scan_ws;
  int64_t words;
  int64_t bytes;
};

bool quantize_layout(int64_t num_plots, const int32_t* dims, void* ws, QWs* out) {
  if (num_plots <= 0 || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return false;
  int64_t cells = num_plots * (int64_t)dims[0] * dims[1] * dims[2];
  if (cells <= 0 || cells > ((int64_t)1 << 36)) return false;
  int64_t words = ceil_div64(cells, 32);
  auto align = [](int64_t b) { return (b + 255) & ~(int64_t)255; };
  char* p = reinterpret_cast<char*>(ws);
  int64_t off = 0;
  out->bitmap = reinterpret_cast<unsigned*>(p + off);
  off += align(words * 4);
  out->prefix = reinterpret_cast<int*>(p + off);
  off += align(words * 4);
  out->oob = reinterpret_cast<int*>(p + off);
  off += 256;
  out->scan_ws = p + off;
  off += align(b2s_scan_workspace_bytes(words));
  out->words = words;
  out->bytes = off;
  return true;
}

}  // namespace

extern "C" int32_t b2s_quantize_points(const float* pos, int64_t n, const int32_t* n_dev, float size,
                                       int32_t* qcoords, int32_t* bounds, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && size > 0.f, "n >= 0 and size > 0");
  B2S_CHECK_ARG(bounds && (n == 0 || (pos && qcoords)), "null pointer");
  cudaStream_t st = as_stream(stream);
  init_bounds_kernel<<<1, 32, 0, st>>>(bounds);
  if (n > 0) quantize_points_kernel<<<grid_for(n * 3, 256), 256, 0, st>>>(pos, n, n_dev, size, qcoords, bounds);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int64_t b2s_quantize_workspace_bytes(int64_t num_plots, const int32_t* dims_host) {
  QWs l;
  if (!dims_host || !quantize_layout(num_plots, dims_host, nullptr, &l)) return -1;
  return l.bytes;
}

extern "C" int32_t b2s_quantize_count(const int32_t* qcoords, const int32_t* plot_of_point, int64_t n,
                                      const int32_t* n_dev, int32_t num_plots, const int32_t* lo_host, const int32_t* dims_host,
                                      void* workspace, int64_t workspace_bytes, int32_t* num_voxels_dev,
                                      b2s_stream_t stream) {
  B2S_CHECK_ARG(lo_host && dims_host && workspace && num_voxels_dev, "null pointer");
  QWs l;
  if (!quantize_layout(num_plots, dims_host, workspace, &l)) {
    b2s_set_error("b2s_quantize_count: voxel box %d x %d x %d x %d plots is empty or above 2^36 cells", dims_host[0],
                  dims_host[1], dims_host[2], num_plots);
    return B2S_EOVERFLOW;
  }
  B2S_CHECK_ARG(workspace_bytes >= l.bytes, "workspace too small");
  cudaStream_t st = as_stream(stream);
  QBox box{{lo_host[0], lo_host[1], lo_host[2]}, {dims_host[0], dims_host[1], dims_host[2]}};
  B2S_CUDA(cudaMemsetAsync(l.bitmap, 0, l.words * 4, st));
  B2S_CUDA(cudaMemsetAsync(l.oob, 0, 4, st));
  if (n > 0)
    quantize_mark_kernel<<<grid_for(n, 256), 256, 0, st>>>(qcoords, plot_of_point, n, n_dev, num_plots, box, l.bitmap,
                                                           l.oob);
  popc_kernel<<<grid_for(l.words, 256), 256, 0, st>>>(l.bitmap, l.words, l.prefix);
  launch_exclusive_scan(l.prefix, l.prefix, l.words, nullptr, num_voxels_dev, l.scan_ws, st);
  quantize_total_kernel<<<1, 1, 0, st>>>(l.oob, num_voxels_dev);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_quantize_fill(const int32_t* qcoords, const int32_t* plot_of_point, const int32_t* order,
                                     int64_t n, const int32_t* n_dev, int32_t num_plots, const int32_t* lo_host,
                                     const int32_t* dims_host, void* workspace, int64_t num_voxels,
                                     const int32_t* num_voxels_dev, int32_t* out_coords, int32_t* out_src,
                                     b2s_stream_t stream) {
  B2S_CHECK_ARG(lo_host && dims_host && workspace, "null pointer");
  B2S_CHECK_ARG(num_voxels >= 0 && num_voxels <= n, "num_voxels out of range");
  if (num_voxels == 0) return B2S_OK;
  B2S_CHECK_ARG(out_coords && out_src, "null output");
  QWs l;
  B2S_CHECK_ARG(quantize_layout(num_plots, dims_host, workspace, &l), "bad voxel box");
  cudaStream_t st = as_stream(stream);
  QBox box{{lo_host[0], lo_host[1], lo_host[2]}, {dims_host[0], dims_host[1], dims_host[2]}};
  B2S_CUDA(cudaMemsetAsync(out_src, 0xFF, num_voxels * 4, st));  // -1
  quantize_rep_kernel<<<grid_for(n, 256), 256, 0, st>>>(qcoords, plot_of_point, order, n, n_dev, box, l.bitmap,
                                                        l.prefix, num_voxels, out_src);
  quantize_emit_kernel<<<grid_for(num_voxels, 256), 256, 0, st>>>(qcoords, plot_of_point, order, num_voxels,
                                                                   num_voxels_dev, out_coords, out_src);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_gather_rows(const float* in, const int32_t* idx, int64_t m, const int32_t* m_dev, int32_t c,
                                   float* out, b2s_stream_t stream) {
  B2S_CHECK_ARG(m >= 0 && c > 0, "m >= 0 and c > 0");
  if (m == 0) return B2S_OK;
  B2S_CHECK_ARG(in && idx && out, "null pointer");
  gather_rows_kernel<<<grid_for(m * c, 256), 256, 0, as_stream(stream)>>>(in, idx, m, m_dev, c, out);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// =============================================================================================
// (a2, a3) coordinate hash + strided maps
// =============================================================================================
namespace {

__global__ void __launch_bounds__(256) coordmap_insert_kernel(const int4* __restrict__ coords, int64_t n,
                                                              const int* __restrict__ n_dev, int tsx, int tsy,
                                                              int tsz, B2sEntry* __restrict__ table, uint64_t mask,
                                                              int* __restrict__ slot, int* __restrict__ info) {
  n = b2s_rows(n, n_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = coords[i];
    if (c.x < 0 || c.x >= 65535 || abs(c.y) >= B2S_COORD_LIMIT || abs(c.z) >= B2S_COORD_LIMIT ||
        abs(c.w) >= B2S_COORD_LIMIT) {
      info[1] = 1;
      slot[i] = -1;
      continue;
    }
    if (c.x > info[2]) atomicMax(&info[2], c.x);  // largest batch id (rows are batch-sorted: few atomics)
    const uint64_t key = b2s_pack_key(c.x, b2s_floor_to(c.y, tsx), b2s_floor_to(c.z, tsy), b2s_floor_to(c.w, tsz));
    uint64_t s = b2s_hash64(key) & mask;
    for (;;) {
      unsigned long long old = atomicCAS(&table[s].key, (unsigned long long)B2S_KEY_EMPTY, (unsigned long long)key);
      if (old == B2S_KEY_EMPTY || old == key) {
        atomicMin(reinterpret_cast<unsigned*>(&table[s].val), (unsigned)i);  // first occurrence wins
        slot[i] = (int)s;
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

__global__ void __launch_bounds__(256) coordmap_flag_kernel(const B2sEntry* __restrict__ table,
                                                            const int* __restrict__ slot, int64_t n,
                                                            const int* __restrict__ n_dev, int* __restrict__ flag) {
  n = b2s_rows(n, n_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = slot[i];
    flag[i] = (s >= 0 && table[s].val == (int)i) ? 1 : 0;
  }
}

// firsts: emit floored coordinate at their rank and rewrite the table value to that rank.
// A non-first row i' can never pass the (val == i') test: val is either first(i') < i' or rank <= first.
__global__ void __launch_bounds__(256) coordmap_emit_kernel(const int4* __restrict__ coords, int64_t n,
                                                            const int* __restrict__ n_dev, int tsx, int tsy, int tsz,
                                                            B2sEntry* __restrict__ table,
                                                            const int* __restrict__ slot,
                                                            const int* __restrict__ rank, int64_t out_cap,
                                                            int4* __restrict__ out) {
  n = b2s_rows(n, n_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = slot[i];
    if (s < 0) continue;
    if (table[s].val == (int)i) {
      int r = rank[i];
      int4 c = coords[i];
      if (r < out_cap) {
        out[r] = make_int4(c.x, b2s_floor_to(c.y, tsx), b2s_floor_to(c.z, tsy), b2s_floor_to(c.w, tsz));
        table[s].val = r;
      } else {
        table[s].val = -1;  // beyond the caller's capacity: the key resolves to "no row" (the caller checks info[0])
      }
    }
  }
}

__global__ void __launch_bounds__(256) coordmap_in2out_kernel(const B2sEntry* __restrict__ table,
                                                              const int* __restrict__ slot, int64_t n,
                                                              const int* __restrict__ n_dev,
                                                              int* __restrict__ in2out) {
  n = b2s_rows(n, n_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = slot[i];
    in2out[i] = s >= 0 ? table[s].val : -1;
  }
}

bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" int64_t b2s_hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

extern "C" int32_t b2s_coordmap_insert(const int32_t* coords, int64_t n, const int32_t* n_dev,
                                       const int32_t* ts_host, void* table, int64_t capacity, int32_t* slot, int32_t* rank, int32_t* info_dev,
                                       void* scan_workspace, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && n < INT_MAX, "0 <= n < 2^31");
  B2S_CHECK_ARG(is_pow2(capacity) && capacity >= 2 * n, "capacity must be a power of two >= 2n");
  B2S_CHECK_ARG(table && info_dev && ts_host && scan_workspace, "null pointer");
  B2S_CHECK_ARG(ts_host[0] > 0 && ts_host[1] > 0 && ts_host[2] > 0, "tensor stride must be positive");
  B2S_CHECK_ARG((reinterpret_cast<uintptr_t>(coords) & 15) == 0 && (reinterpret_cast<uintptr_t>(table) & 15) == 0,
                "coords and table must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(table, 0xFF, capacity * sizeof(B2sEntry), st));
  B2S_CUDA(cudaMemsetAsync(info_dev, 0, 4 * sizeof(int), st));
  B2S_CUDA(cudaMemsetAsync(info_dev + 2, 0xFF, sizeof(int), st));  // max batch id starts at -1
  if (n > 0) {
    B2S_CHECK_ARG(coords && slot && rank, "null pointer");
    coordmap_insert_kernel<<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n, n_dev,
                                                             ts_host[0], ts_host[1], ts_host[2],
                                                             reinterpret_cast<B2sEntry*>(table),
                                                             (uint64_t)capacity - 1, slot, info_dev);
    coordmap_flag_kernel<<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const B2sEntry*>(table), slot, n, n_dev,
                                                           rank);
    launch_exclusive_scan(rank, rank, n, n_dev, info_dev, scan_workspace, st);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_coordmap_fill(const int32_t* coords, int64_t n, const int32_t* n_dev, const int32_t* ts_host,
                                     void* table, int64_t capacity, const int32_t* slot, const int32_t* rank,
                                     int32_t* out_coords, int64_t out_capacity, int32_t* in2out,
                                     b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && is_pow2(capacity), "bad sizes");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(coords && table && slot && rank && out_coords && ts_host, "null pointer");
  B2S_CHECK_ARG((reinterpret_cast<uintptr_t>(out_coords) & 15) == 0, "out_coords must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  coordmap_emit_kernel<<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n, n_dev, ts_host[0],
                                                         ts_host[1], ts_host[2], reinterpret_cast<B2sEntry*>(table),
                                                         slot, rank, out_capacity,
                                                         reinterpret_cast<int4*>(out_coords));
  if (in2out)
    coordmap_in2out_kernel<<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const B2sEntry*>(table), slot, n, n_dev,
                                                             in2out);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// =============================================================================================
// (a4) kernel maps
// =============================================================================================
namespace {

struct KmParams {
  int K[3];
  int step[3];
  int sign;
  int k3;
  int group;  // offsets handled per blockIdx.y
};

// One thread per query row, blockIdx.y selects a group of kernel offsets.  Writes of nbr[k, q] are
// coalesced across the warp for every k; the 16-byte coordinate is loaded once per group.
__global__ void __launch_bounds__(256) kernel_map_kernel(const int4* __restrict__ query, int64_t n,
                                                         const int* __restrict__ n_dev,
                                                         const B2sEntry* __restrict__ table, uint64_t mask,
                                                         KmParams p, int* __restrict__ nbr) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b2s_rows(n, n_dev)) return;
  const int4 c = query[q];
  const int k0 = blockIdx.y * p.group;
  const int k1 = min(k0 + p.group, p.k3);
  const int hx = (p.K[0] & 1) ? p.K[0] / 2 : 0, hy = (p.K[1] & 1) ? p.K[1] / 2 : 0, hz = (p.K[2] & 1) ? p.K[2] / 2 : 0;
  int ix = k0 % p.K[0], iy = (k0 / p.K[0]) % p.K[1], iz = k0 / (p.K[0] * p.K[1]);
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    const int x = c.y + p.sign * (ix - hx) * p.step[0];
    const int y = c.z + p.sign * (iy - hy) * p.step[1];
    const int z = c.w + p.sign * (iz - hz) * p.step[2];
    nbr[(int64_t)k * n + q] = b2s_table_find(table, mask, b2s_pack_key(c.x, x, y, z));
    if (++ix == p.K[0]) {
      ix = 0;
      if (++iy == p.K[1]) {
        iy = 0;
        ++iz;
      }
    }
  }
}

// Kernel map through the quantiser's occupancy index instead of the hash: the rows of a quantised map are sorted by
// cell (plot, z, y, x), so  row(cell) = prefix[cell >> 5] + popc(bitmap[cell >> 5] & below(cell & 31)).  The index of
// a full batch is ~5 MB (L2 / L1 resident) and the x-neighbours of x-consecutive rows share words, which turns the
// 343 random 16-byte hash probes per row of the k7 stem map into L1 hits.  Same output as kernel_map_kernel, bit for bit.
__global__ void __launch_bounds__(256) kernel_map_dense_kernel(const int4* __restrict__ query, int64_t n,
                                                               const int* __restrict__ n_dev,
                                                               const unsigned* __restrict__ bitmap,
                                                               const int* __restrict__ prefix, QBox box,
                                                               int num_plots, KmParams p, int* __restrict__ nbr) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b2s_rows(n, n_dev)) return;
  const int4 c = query[q];
  const int k0 = blockIdx.y * p.group;
  const int k1 = min(k0 + p.group, p.k3);
  const int hx = (p.K[0] & 1) ? p.K[0] / 2 : 0, hy = (p.K[1] & 1) ? p.K[1] / 2 : 0, hz = (p.K[2] & 1) ? p.K[2] / 2 : 0;
  int ix = k0 % p.K[0], iy = (k0 / p.K[0]) % p.K[1], iz = k0 / (p.K[0] * p.K[1]);
  const bool plot_ok = (unsigned)c.x < (unsigned)num_plots;
  if (p.step[0] == 1) {
    // x-lines: the K[0] cells of one (iy, iz) line are consecutive bits of the occupancy bitmap (at most two words), so
    // the words and their ranks are loaded once per line -- 2-4 loads instead of 2 K[0] -- and every offset of the line
    // is a popcount.  With sign = -1 the line is walked downwards; the bits are the same.
    int wi0 = 0;
    int64_t cellbase = 0;
    unsigned w0 = 0, w1 = 0;
    int p0 = 0, p1 = 0;
    bool line_ok = false;
    for (int k = k0; k < k1; ++k) {
      if (k == k0 || ix == 0) {
        const int y = c.z + p.sign * (iy - hy) * p.step[1] - box.lo[1];
        const int z = c.w + p.sign * (iz - hz) * p.step[2] - box.lo[2];
        line_ok = plot_ok && (unsigned)y < (unsigned)box.dim[1] && (unsigned)z < (unsigned)box.dim[2];
        if (line_ok) {
          const int xlo = c.y - box.lo[0] - (p.sign > 0 ? hx : (p.K[0] - 1 - hx));   // smallest x of the line
          const int xa = max(xlo, 0), xb = min(xlo + p.K[0] - 1, box.dim[0] - 1);
          cellbase = (((int64_t)c.x * box.dim[2] + z) * box.dim[1] + y) * box.dim[0];
          line_ok = xa <= xb;
          if (line_ok) {
            wi0 = (int)((cellbase + xa) >> 5);
            w0 = __ldg(&bitmap[wi0]);
            p0 = __ldg(&prefix[wi0]);
            if ((int)((cellbase + xb) >> 5) != wi0) {
              w1 = __ldg(&bitmap[wi0 + 1]);
              p1 = __ldg(&prefix[wi0 + 1]);
            }
          }
        }
      }
      int row = -1;
      const int x = c.y + p.sign * (ix - hx) - box.lo[0];
      if (line_ok && (unsigned)x < (unsigned)box.dim[0]) {
        const int64_t cell = cellbase + x;
        const bool first = (int)(cell >> 5) == wi0;
        const unsigned w = first ? w0 : w1;
        const unsigned bit = (unsigned)(cell & 31);
        if ((w >> bit) & 1u) row = (first ? p0 : p1) + __popc(w & ((1u << bit) - 1u));
      }
      nbr[(int64_t)k * n + q] = row;
      if (++ix == p.K[0]) {
        ix = 0;
        if (++iy == p.K[1]) {
          iy = 0;
          ++iz;
        }
      }
    }
    return;
  }
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    const int x = c.y + p.sign * (ix - hx) * p.step[0] - box.lo[0];
    const int y = c.z + p.sign * (iy - hy) * p.step[1] - box.lo[1];
    const int z = c.w + p.sign * (iz - hz) * p.step[2] - box.lo[2];
    int row = -1;
    if (plot_ok && (unsigned)x < (unsigned)box.dim[0] && (unsigned)y < (unsigned)box.dim[1] &&
        (unsigned)z < (unsigned)box.dim[2]) {
      const int64_t cell = (((int64_t)c.x * box.dim[2] + z) * box.dim[1] + y) * box.dim[0] + x;
      const unsigned w = __ldg(&bitmap[cell >> 5]);
      const unsigned bit = (unsigned)(cell & 31);
      if ((w >> bit) & 1u) row = __ldg(&prefix[cell >> 5]) + __popc(w & ((1u << bit) - 1u));
    }
    nbr[(int64_t)k * n + q] = row;
    if (++ix == p.K[0]) {
      ix = 0;
      if (++iy == p.K[1]) {
        iy = 0;
        ++iz;
      }
    }
  }
}

// x-line form of a stride-1 kernel map over a quantised (cell-sorted) map: ONE 32-bit word per (query row, line of
// K[0] x-consecutive kernel offsets):  word = (base << 8) | mask,  mask bit ix set <=> the cell of offset ix of the
// line is occupied, base = row of the lowest occupied cell of the window.  Rows are sorted by cell (plot, z, y, x), so
// the occupied cells of a line are CONSECUTIVE rows base, base + 1, ...: row(ix) = base + popc(mask & below(ix)).
// 49 words per row instead of the 343 table entries of the k7 stem (83 MB instead of 580 MB at batch 32), built from
// at most two bitmap words and one prefix entry per line.
__global__ void __launch_bounds__(256) kernel_map_lines_kernel(const int4* __restrict__ query, int64_t n,
                                                               const int* __restrict__ n_dev,
                                                               const unsigned* __restrict__ bitmap,
                                                               const int* __restrict__ prefix, QBox box,
                                                               int num_plots, KmParams p, int nlines,
                                                               unsigned* __restrict__ lines) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b2s_rows(n, n_dev)) return;
  const int4 c = query[q];
  const int l0 = blockIdx.y * p.group;
  const int l1 = min(l0 + p.group, nlines);
  const int hx = (p.K[0] & 1) ? p.K[0] / 2 : 0, hy = (p.K[1] & 1) ? p.K[1] / 2 : 0, hz = (p.K[2] & 1) ? p.K[2] / 2 : 0;
  const bool plot_ok = (unsigned)c.x < (unsigned)num_plots;
  const int xlo = c.y - box.lo[0] - hx;                                    // cell x of offset ix = 0
  const int xa = max(xlo, 0), xb = min(xlo + p.K[0] - 1, box.dim[0] - 1);
  const bool row_ok = plot_ok && xa <= xb;
  const unsigned span = (unsigned)(xb - xa), wmask = (2u << span) - 1u, shl = (unsigned)(xa - xlo);
  // cell of (xa, y, z) = plot_base + (z * dim1 + y) * dim0 + xa, walked with additions: iy fastest
  const int64_t plot_base = (int64_t)c.x * box.dim[2] * box.dim[1] * box.dim[0] + xa;
  int iy = l0 % p.K[1], iz = l0 / p.K[1];
  int y = c.z + (iy - hy) * p.step[1] - box.lo[1];
  int z = c.w + (iz - hz) * p.step[2] - box.lo[2];
  const int y_first = c.z - hy * p.step[1] - box.lo[1];
  unsigned* out = lines + (int64_t)l0 * n + q;
  for (int l = l0; l < l1; ++l, out += n) {
    unsigned word = 0;
    if (row_ok && (unsigned)y < (unsigned)box.dim[1] && (unsigned)z < (unsigned)box.dim[2]) {
      const int64_t c0 = plot_base + ((int64_t)z * box.dim[1] + y) * box.dim[0];
      const int64_t wi = c0 >> 5;
      const unsigned b0 = (unsigned)c0 & 31u;
      const unsigned w0 = __ldg(&bitmap[wi]);
      unsigned bits = w0 >> b0;
      if (b0 + span > 31u) bits |= __ldg(&bitmap[wi + 1]) << (32u - b0);   // the window crosses into the next word
      bits &= wmask;
      if (bits) word = ((unsigned)(__ldg(&prefix[wi]) + __popc(w0 & ((1u << b0) - 1u))) << 8) | (bits << shl);
    }
    *out = word;
    y += p.step[1];
    if (++iy == p.K[1]) {
      iy = 0;
      y = y_first;
      z += p.step[2];
    }
  }
}

__global__ void __launch_bounds__(256) pair_count_kernel(const int* __restrict__ nbr, int64_t n,
                                                         int* __restrict__ counts) {
  const int k = blockIdx.y;
  int c = 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    c += nbr[(int64_t)k * n + q] >= 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counts[k], c);
}

// one CTA per kernel offset walks the row of the table in order, so pairs come out sorted by out row
__global__ void __launch_bounds__(SCAN_THREADS) pair_fill_kernel(const int* __restrict__ nbr, int64_t n,
                                                                 const int64_t* __restrict__ offsets,
                                                                 int* __restrict__ in_idx, int* __restrict__ out_idx) {
  const int k = blockIdx.x;
  int64_t base_out = offsets[k];
  for (int64_t q0 = 0; q0 < n; q0 += SCAN_THREADS) {
    int64_t q = q0 + threadIdx.x;
    int v = q < n ? nbr[(int64_t)k * n + q] : -1;
    int total;
    int ex = block_excl_scan(v >= 0 ? 1 : 0, &total);
    if (v >= 0) {
      in_idx[base_out + ex] = v;
      out_idx[base_out + ex] = (int)q;
    }
    base_out += total;
  }
}

}  // namespace

extern "C" int32_t b2s_kernel_map(const int32_t* query_coords, int64_t n_query, const int32_t* n_query_dev,
                                  const void* table, int64_t capacity,
                                  const int32_t* kernel_size_host, const int32_t* step_host, int32_t sign,
                                  int32_t* nbr, b2s_stream_t stream) {
  B2S_CHECK_ARG(n_query >= 0 && is_pow2(capacity), "bad sizes");
  B2S_CHECK_ARG(kernel_size_host && step_host && table, "null pointer");
  B2S_CHECK_ARG(sign == 1 || sign == -1, "sign must be +1 or -1");
  KmParams p;
  for (int d = 0; d < 3; ++d) {
    B2S_CHECK_ARG(kernel_size_host[d] >= 1 && kernel_size_host[d] <= 15 && step_host[d] >= 1, "kernel size 1..15");
    p.K[d] = kernel_size_host[d];
    p.step[d] = step_host[d];
    B2S_CHECK_ARG((int64_t)p.K[d] * p.step[d] < (B2S_COORD_BIAS - B2S_COORD_LIMIT), "kernel reach too large");
  }
  p.sign = sign;
  p.k3 = p.K[0] * p.K[1] * p.K[2];
  if (n_query == 0) return B2S_OK;
  B2S_CHECK_ARG(query_coords && nbr, "null pointer");
  // enough CTAs for >= 2 waves of 148 SMs x 8 even when the map is small
  int64_t row_blocks = ceil_div64(n_query, 256);
  int groups = 1;
  while (row_blocks * groups < 2 * B2S_NUM_SMS * 8 && groups < p.k3) ++groups;
  p.group = (p.k3 + groups - 1) / groups;
  groups = (p.k3 + p.group - 1) / p.group;
  dim3 grid((unsigned)row_blocks, (unsigned)groups);
  kernel_map_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const int4*>(query_coords), n_query,
                                                         n_query_dev, reinterpret_cast<const B2sEntry*>(table),
                                                         (uint64_t)capacity - 1, p, nbr);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_kernel_map_dense(const int32_t* query_coords, int64_t n_query, const int32_t* n_query_dev,
                                        const void* quantize_workspace, int32_t num_plots, const int32_t* lo_host,
                                        const int32_t* dims_host, const int32_t* kernel_size_host,
                                        const int32_t* step_host, int32_t sign, int32_t* nbr, b2s_stream_t stream) {
  B2S_CHECK_ARG(n_query >= 0 && kernel_size_host && step_host && quantize_workspace && lo_host && dims_host,
                "bad arguments");
  B2S_CHECK_ARG(sign == 1 || sign == -1, "sign must be +1 or -1");
  QWs l;
  B2S_CHECK_ARG(quantize_layout(num_plots, dims_host, const_cast<void*>(quantize_workspace), &l), "bad voxel box");
  KmParams p;
  for (int d = 0; d < 3; ++d) {
    B2S_CHECK_ARG(kernel_size_host[d] >= 1 && kernel_size_host[d] <= 15 && step_host[d] >= 1, "kernel size 1..15");
    p.K[d] = kernel_size_host[d];
    p.step[d] = step_host[d];
  }
  p.sign = sign;
  p.k3 = p.K[0] * p.K[1] * p.K[2];
  if (n_query == 0) return B2S_OK;
  B2S_CHECK_ARG(query_coords && nbr, "null pointer");
  int64_t row_blocks = ceil_div64(n_query, 256);
  int groups = 1;
  while (row_blocks * groups < 2 * B2S_NUM_SMS * 8 && groups < p.k3) ++groups;
  p.group = (p.k3 + groups - 1) / groups;
  groups = (p.k3 + p.group - 1) / p.group;
  QBox box{{lo_host[0], lo_host[1], lo_host[2]}, {dims_host[0], dims_host[1], dims_host[2]}};
  dim3 grid((unsigned)row_blocks, (unsigned)groups);
  kernel_map_dense_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const int4*>(query_coords), n_query,
                                                               n_query_dev, l.bitmap, l.prefix, box, num_plots, p,
                                                               nbr);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_kernel_map_lines(const int32_t* query_coords, int64_t n_query, const int32_t* n_query_dev,
                                        const void* quantize_workspace, int32_t num_plots, const int32_t* lo_host,
                                        const int32_t* dims_host, const int32_t* kernel_size_host,
                                        const int32_t* step_host, uint32_t* lines, b2s_stream_t stream) {
  B2S_CHECK_ARG(n_query >= 0 && kernel_size_host && step_host && quantize_workspace && lo_host && dims_host,
                "bad arguments");
  QWs l;
  B2S_CHECK_ARG(quantize_layout(num_plots, dims_host, const_cast<void*>(quantize_workspace), &l), "bad voxel box");
  KmParams p;
  for (int d = 0; d < 3; ++d) {
    B2S_CHECK_ARG(kernel_size_host[d] >= 1 && kernel_size_host[d] <= 15 && step_host[d] >= 1, "kernel size 1..15");
    p.K[d] = kernel_size_host[d];
    p.step[d] = step_host[d];
  }
  B2S_CHECK_ARG(p.K[0] <= 8 && p.step[0] == 1, "x-lines need kernel_size[0] <= 8 and an x step of 1");
  if (n_query >= ((int64_t)1 << 24)) {
    b2s_set_error("b2s_kernel_map_lines: %lld rows do not fit the 24-bit row field of a line word", (long long)n_query);
    return B2S_EOVERFLOW;
  }
  p.sign = 1;
  p.k3 = p.K[0] * p.K[1] * p.K[2];
  if (n_query == 0) return B2S_OK;
  B2S_CHECK_ARG(query_coords && lines, "null pointer");
  const int nlines = p.K[1] * p.K[2];
  int64_t row_blocks = ceil_div64(n_query, 256);
  int groups = 1;
  while (row_blocks * groups < 2 * B2S_NUM_SMS * 8 && groups < nlines) ++groups;
  p.group = (nlines + groups - 1) / groups;
  groups = (nlines + p.group - 1) / p.group;
  QBox box{{lo_host[0], lo_host[1], lo_host[2]}, {dims_host[0], dims_host[1], dims_host[2]}};
  dim3 grid((unsigned)row_blocks, (unsigned)groups);
  kernel_map_lines_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const int4*>(query_coords), n_query,
                                                               n_query_dev, l.bitmap, l.prefix, box, num_plots, p,
                                                               nlines, lines);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// =============================================================================================
// parity plan: fine rows of a stride-2 map sorted by their position inside the 2x2x2 coarse cell
// =============================================================================================
namespace {

__device__ __forceinline__ int parity_class(const int4 c, int tx, int ty, int tz) {
  // bit d set <=> the coordinate is off the coarse lattice (not a multiple of the coarse tensor stride)
  return ((c.y - b2s_floor_to(c.y, tx)) != 0 ? 1 : 0) | ((c.z - b2s_floor_to(c.z, ty)) != 0 ? 2 : 0) |
         ((c.w - b2s_floor_to(c.w, tz)) != 0 ? 4 : 0);
}

__global__ void __launch_bounds__(256) parity_count_kernel(const int4* __restrict__ coords, int64_t n,
                                                           const int* __restrict__ n_dev, int tx, int ty, int tz,
                                                           int* __restrict__ counts) {
  n = b2s_rows(n, n_dev);
  __shared__ int h[8];
  if (threadIdx.x < 8) h[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&h[parity_class(coords[i], tx, ty, tz)], 1);
  __syncthreads();
  if (threadIdx.x < 8 && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

// bounds[c] = first tile of class c (tiles of 128 rows), bounds[8] = number of tiles; cursor[c] = first slot of class c
__global__ void parity_bounds_kernel(const int* __restrict__ counts, int* __restrict__ bounds, int* __restrict__ cursor) {
  if (threadIdx.x == 0) {
    int tile = 0;
    for (int c = 0; c < 8; ++c) {
      bounds[c] = tile;
      cursor[c] = tile * 128;
      tile += (counts[c] + 127) / 128;
    }
    bounds[8] = tile;
  }
}

__global__ void __launch_bounds__(256) parity_scatter_kernel(const int4* __restrict__ coords, int64_t n,
                                                             const int* __restrict__ n_dev, int tx, int ty, int tz,
                                                             int* __restrict__ cursor, int* __restrict__ perm) {
  n = b2s_rows(n, n_dev);
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    const int cls = i < n ? parity_class(coords[i], tx, ty, tz) : -1;
    // warp-aggregated slot claim: one atomic per (warp, class)
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const unsigned m = __ballot_sync(0xffffffffu, cls == c);
      if (m == 0) continue;
      const int leader = __ffs(m) - 1;
      int basev = 0;
      if ((int)lane == leader) basev = atomicAdd(&cursor[c], __popc(m));
      basev = __shfl_sync(0xffffffffu, basev, leader);
      if (cls == c) perm[basev + __popc(m & ((1u << lane) - 1u))] = (int)i;
    }
  }
}

}  // namespace

extern "C" int64_t b2s_parity_plan_rows(int64_t n) { return (ceil_div64(n > 0 ? n : 1, 128) + 8) * 128; }

extern "C" int32_t b2s_parity_plan(const int32_t* coords, int64_t n, const int32_t* n_dev, const int32_t* ts_coarse_host,
                                   int32_t* perm, int32_t* bounds, int32_t* scratch16, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && ts_coarse_host && perm && bounds && scratch16, "bad arguments");
  B2S_CHECK_ARG(ts_coarse_host[0] > 0 && ts_coarse_host[1] > 0 && ts_coarse_host[2] > 0, "tensor stride must be positive");
  cudaStream_t st = as_stream(stream);
  int* counts = scratch16;        // [8]
  int* cursor = scratch16 + 8;    // [8]
  B2S_CUDA(cudaMemsetAsync(scratch16, 0, 16 * sizeof(int), st));
  B2S_CUDA(cudaMemsetAsync(perm, 0xFF, b2s_parity_plan_rows(n) * sizeof(int), st));   // -1 = padding
  if (n > 0) {
    B2S_CHECK_ARG(coords, "null pointer");
    parity_count_kernel<<<grid_for(n, 256, 4), 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n, n_dev,
                                                             ts_coarse_host[0], ts_coarse_host[1], ts_coarse_host[2],
                                                             counts);
  }
  parity_bounds_kernel<<<1, 32, 0, st>>>(counts, bounds, cursor);
  if (n > 0)
    parity_scatter_kernel<<<grid_for(n, 256, 4), 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n, n_dev,
                                                               ts_coarse_host[0], ts_coarse_host[1], ts_coarse_host[2],
                                                               cursor, perm);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_kernel_map_pair_counts(const int32_t* nbr, int32_t k3, int64_t n_query, int32_t* counts,
                                              b2s_stream_t stream) {
  B2S_CHECK_ARG(k3 > 0 && n_query >= 0 && counts, "bad arguments");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(counts, 0, k3 * sizeof(int), st));
  if (n_query == 0) return B2S_OK;
  B2S_CHECK_ARG(nbr, "null pointer");
  dim3 grid((unsigned)grid_for(n_query, 256, 2), (unsigned)k3);
  pair_count_kernel<<<grid, 256, 0, st>>>(nbr, n_query, counts);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_kernel_map_pairs_fill(const int32_t* nbr, int32_t k3, int64_t n_query, const int64_t* offsets,
                                             int32_t* in_idx, int32_t* out_idx, b2s_stream_t stream) {
  B2S_CHECK_ARG(k3 > 0 && n_query >= 0, "bad arguments");
  if (n_query == 0) return B2S_OK;
  B2S_CHECK_ARG(nbr && offsets && in_idx && out_idx, "null pointer");
  pair_fill_kernel<<<k3, SCAN_THREADS, 0, as_stream(stream)>>>(nbr, n_query, offsets, in_idx, out_idx);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// =============================================================================================
// (f2) GPU input pipeline -- the arithmetic transforms the reference runs per sample on the CPU before the quantiser
// (R:conf/data/instance/NFI/transforms/sparse-xy.yaml:105-152 test_transform; the same ops sit inside train_transform):
//   ScalePos(op="div")           R:core/data_transform/transforms.py:590-598   pos = pos / (sx, sy, sz)
//   MoveCenterPosPerSample       :722-739                                      pos += (cx, cy, cz)
//   StartZFromZero               :766-769                                      z -= min z of the plot
//   Polygon2dExtend              :1489-1496  keep the points whose (x, y) lies inside the polygon (matplotlib
//                                            Path.contains_points == the crossings test restated below, in double)
//   MaxPoints                    :1772-1791  choice = randperm(n)[:num], rows in choice order
//   AddOnes / XYZFeature(z) / AddXYDistanceToCenter / AddFeatsByKeys   R:core/data_transform/features.py:307-383
//                                            x = [1, z, ||(x, y) - centre + 1e-6||_2]  (torch PairwiseDistance)
//   RandomCoordsFlip(ignored z) / ShiftVoxels   :1046-1054 / R:core/data_transform/sparse_transforms.py:49-55
//                                            on the quantised coordinates, per plot
// Points of a plot are contiguous (collated batch); every kernel takes the device point count.
// =============================================================================================
namespace {

// order-preserving map float -> uint32 (for atomicMin on floats of either sign)
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

constexpr int PP_CHUNK = 64;   // consecutive points per thread: one atomic per plot change instead of one per point

__global__ void __launch_bounds__(256) plot_minz_kernel(const float* __restrict__ pos, const int* __restrict__ plot,
                                                        int64_t n, const int* __restrict__ n_dev, float sz, float cz,
                                                        unsigned* __restrict__ minz) {
  n = b2s_rows(n, n_dev);
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PP_CHUNK;
  int cur = -1;
  unsigned best = 0xFFFFFFFFu;
  for (int64_t i = i0; i < n && i < i0 + PP_CHUNK; ++i) {
    const int p = __ldg(&plot[i]);
    if (p != cur) {
      if (cur >= 0) atomicMin(&minz[cur], best);
      cur = p;
      best = 0xFFFFFFFFu;
    }
    const unsigned o = f2ord(__fadd_rn(__fdiv_rn(__ldg(&pos[i * 3 + 2]), sz), cz));
    best = o < best ? o : best;
  }
  if (cur >= 0) atomicMin(&minz[cur], best);
}

struct Polygon {
  int nv;
  double vx[16], vy[16];
};

// matplotlib's point_in_path (src/_path.h, the crossings test of the comp.graphics.algorithms FAQ) for one closed
// polygon, in double like Path.contains_points
__device__ __forceinline__ bool inside_polygon(const Polygon& pg, double tx, double ty) {
  bool inside = false;
  double x0 = pg.vx[pg.nv - 1], y0 = pg.vy[pg.nv - 1];
  bool yflag0 = y0 >= ty;
  for (int j = 0; j < pg.nv; ++j) {
    const double x1 = pg.vx[j], y1 = pg.vy[j];
    const bool yflag1 = y1 >= ty;
    if (yflag0 != yflag1) {
      if (((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == yflag1) inside = !inside;
    }
    yflag0 = yflag1;
    x0 = x1;
    y0 = y1;
  }
  return inside;
}

__global__ void __launch_bounds__(256) plot_transform_kernel(const float* __restrict__ pos, const int* __restrict__ plot,
                                                             int64_t n, const int* __restrict__ n_dev, float sx,
                                                             float sy, float sz, float cx, float cy, float cz,
                                                             const unsigned* __restrict__ minz, const Polygon pg,
                                                             float* __restrict__ out_pos, int* __restrict__ keep) {
  n = b2s_rows(n, n_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __fadd_rn(__fdiv_rn(__ldg(&pos[i * 3]), sx), cx);
    const float y = __fadd_rn(__fdiv_rn(__ldg(&pos[i * 3 + 1]), sy), cy);
    const float z = __fsub_rn(__fadd_rn(__fdiv_rn(__ldg(&pos[i * 3 + 2]), sz), cz), ord2f(__ldg(&minz[__ldg(&plot[i])])));
    out_pos[i * 3] = x;
    out_pos[i * 3 + 1] = y;
    out_pos[i * 3 + 2] = z;
    keep[i] = (pg.nv < 3 || inside_polygon(pg, (double)x, (double)y)) ? 1 : 0;
  }
}

// stable compaction: rows with keep != 0 move to their exclusive-scan position
__global__ void __launch_bounds__(256) compact_points_kernel(const float* __restrict__ pos, const int* __restrict__ plot,
                                                             const int* __restrict__ keep, const int* __restrict__ idx,
                                                             int64_t n, const int* __restrict__ n_dev,
                                                             float* __restrict__ out_pos, int* __restrict__ out_plot) {
  n = b2s_rows(n, n_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!keep[i]) continue;
    const int64_t o = idx[i];
    out_pos[o * 3] = pos[i * 3];
    out_pos[o * 3 + 1] = pos[i * 3 + 1];
    out_pos[o * 3 + 2] = pos[i * 3 + 2];
    out_plot[o] = plot[i];
  }
}

// feats = [1, z, ||(x, y) - (cx, cy) + 1e-6||]: torch.nn.PairwiseDistance(eps=1e-6) evaluates
// sqrt(fma(dy, dy, dx*dx)) with dx, dy = (x - cx) + eps in fp32 on the CPU (checked bit for bit, tests/test_oracle.py)
__global__ void __launch_bounds__(256) point_features_kernel(const float* __restrict__ pos, int64_t n,
                                                             const int* __restrict__ n_dev, float cx, float cy,
                                                             float* __restrict__ feats) {
  n = b2s_rows(n, n_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float dx = __fadd_rn(__fsub_rn(pos[i * 3], cx), 1e-6f), dy = __fadd_rn(__fsub_rn(pos[i * 3 + 1], cy), 1e-6f);
    feats[i * 3] = 1.0f;
    feats[i * 3 + 1] = pos[i * 3 + 2];
    feats[i * 3 + 2] = __fsqrt_rn(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
  }
}

// MaxPoints: point i of plot p with rank r = rank[i] (its position in the plot's random permutation) goes to row
// offsets[p] + r if r < num; offsets = exclusive sum of min(count, num)
__global__ void capped_offsets_kernel(const int* __restrict__ counts, int num_plots, int num, int* __restrict__ offsets) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int s = 0;
    for (int p = 0; p < num_plots; ++p) {
      offsets[p] = s;
      s += counts[p] < num ? counts[p] : num;
    }
    offsets[num_plots] = s;
  }
}
__global__ void __launch_bounds__(256) select_by_rank_kernel(const float* __restrict__ pos, const int* __restrict__ plot,
                                                             const int* __restrict__ rank, int64_t n,
                                                             const int* __restrict__ n_dev, int num,
                                                             const int* __restrict__ offsets,
                                                             float* __restrict__ out_pos, int* __restrict__ out_plot) {
  n = b2s_rows(n, n_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = rank[i];
    if (r < 0 || r >= num) continue;
    const int p = plot[i];
    const int64_t o = (int64_t)offsets[p] + r;
    out_pos[o * 3] = pos[i * 3];
    out_pos[o * 3 + 1] = pos[i * 3 + 1];
    out_pos[o * 3 + 2] = pos[i * 3 + 2];
    out_plot[o] = p;
  }
}

// per-plot maximum of the x and y voxel coordinates (RandomCoordsFlip flips about the maximum of the sample)
__global__ void __launch_bounds__(256) coords_max_kernel(const int* __restrict__ coords, int64_t m,
                                                         const int* __restrict__ m_dev, int* __restrict__ maxc) {
  m = b2s_rows(m, m_dev);
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PP_CHUNK;
  int cur = -1, bx = INT_MIN, by = INT_MIN;
  for (int64_t i = i0; i < m && i < i0 + PP_CHUNK; ++i) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
    if (c.x != cur) {
      if (cur >= 0) {
        atomicMax(&maxc[cur * 2], bx);
        atomicMax(&maxc[cur * 2 + 1], by);
      }
      cur = c.x;
      bx = by = INT_MIN;
    }
    bx = max(bx, c.y);
    by = max(by, c.z);
  }
  if (cur >= 0) {
    atomicMax(&maxc[cur * 2], bx);
    atomicMax(&maxc[cur * 2 + 1], by);
  }
}
// aug int32 [B, 5] = (flip_x, flip_y, shift_x, shift_y, shift_z) per plot
__global__ void __launch_bounds__(256) coords_augment_kernel(int* __restrict__ coords, int64_t m,
                                                             const int* __restrict__ m_dev, const int* __restrict__ aug,
                                                             const int* __restrict__ maxc) {
  m = b2s_rows(m, m_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    int4 c = reinterpret_cast<int4*>(coords)[i];
    const int* a = aug + c.x * 5;
    if (a[0]) c.y = maxc[c.x * 2] - c.y;
    if (a[1]) c.z = maxc[c.x * 2 + 1] - c.z;
    c.y += a[2];
    c.z += a[3];
    c.w += a[4];
    reinterpret_cast<int4*>(coords)[i] = c;
  }
}

}  // namespace

extern "C" int32_t b2s_plot_transform(const float* pos, const int32_t* plot_of_point, int64_t n, const int32_t* n_dev,
                                      int32_t num_plots, const float* scale_host, const float* center_host,
                                      const double* polygon_host, int32_t num_vertices, uint32_t* minz_scratch,
                                      float* out_pos, int32_t* keep, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && num_plots > 0 && scale_host && center_host && num_vertices >= 0 && num_vertices <= 16,
                "bad arguments (at most 16 polygon vertices)");
  B2S_CHECK_ARG(num_vertices == 0 || polygon_host, "polygon is NULL");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(pos && plot_of_point && minz_scratch && out_pos && keep, "null pointer");
  cudaStream_t st = as_stream(stream);
  Polygon pg{};
  pg.nv = num_vertices;
  for (int j = 0; j < num_vertices; ++j) {
    pg.vx[j] = polygon_host[2 * j];
    pg.vy[j] = polygon_host[2 * j + 1];
  }
  B2S_CUDA(cudaMemsetAsync(minz_scratch, 0xFF, (size_t)num_plots * sizeof(uint32_t), st));
  plot_minz_kernel<<<grid_for(ceil_div64(n, PP_CHUNK), 256), 256, 0, st>>>(pos, plot_of_point, n, n_dev, scale_host[2],
                                                                          center_host[2], minz_scratch);
  plot_transform_kernel<<<grid_for(n, 256), 256, 0, st>>>(pos, plot_of_point, n, n_dev, scale_host[0], scale_host[1],
                                                          scale_host[2], center_host[0], center_host[1], center_host[2],
                                                          minz_scratch, pg, out_pos, keep);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_compact_points(const float* pos, const int32_t* plot_of_point, const int32_t* keep, int64_t n,
                                      const int32_t* n_dev, int32_t* index_scratch, void* scan_workspace,
                                      float* out_pos, int32_t* out_plot, int32_t* num_kept_dev, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && num_kept_dev, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    B2S_CUDA(cudaMemsetAsync(num_kept_dev, 0, sizeof(int32_t), st));
    return B2S_OK;
  }
  B2S_CHECK_ARG(pos && plot_of_point && keep && index_scratch && scan_workspace && out_pos && out_plot, "null pointer");
  launch_exclusive_scan(keep, index_scratch, n, n_dev, num_kept_dev, scan_workspace, st);
  compact_points_kernel<<<grid_for(n, 256), 256, 0, st>>>(pos, plot_of_point, keep, index_scratch, n, n_dev, out_pos,
                                                          out_plot);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_point_features(const float* pos, int64_t n, const int32_t* n_dev, float center_x, float center_y,
                                      float* feats, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0, "n >= 0");
  if (n == 0) return B2S_OK;
  B2S_CHECK_ARG(pos && feats, "null pointer");
  point_features_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(pos, n, n_dev, center_x, center_y, feats);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_select_by_rank(const float* pos, const int32_t* plot_of_point, const int32_t* rank, int64_t n,
                                      const int32_t* n_dev, int32_t num_plots, int32_t num, const int32_t* counts,
                                      int32_t* offsets, float* out_pos, int32_t* out_plot, b2s_stream_t stream) {
  B2S_CHECK_ARG(n >= 0 && num_plots > 0 && num > 0 && counts && offsets, "bad arguments");
  cudaStream_t st = as_stream(stream);
  capped_offsets_kernel<<<1, 32, 0, st>>>(counts, num_plots, num, offsets);
  if (n > 0) {
    B2S_CHECK_ARG(pos && plot_of_point && rank && out_pos && out_plot, "null pointer");
    select_by_rank_kernel<<<grid_for(n, 256), 256, 0, st>>>(pos, plot_of_point, rank, n, n_dev, num, offsets, out_pos,
                                                            out_plot);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_coords_augment(int32_t* coords, int64_t m, const int32_t* m_dev, int32_t num_plots,
                                      const int32_t* aug_dev, int32_t* max_scratch, b2s_stream_t stream) {
  B2S_CHECK_ARG(m >= 0 && num_plots > 0, "bad arguments");
  if (m == 0) return B2S_OK;
  B2S_CHECK_ARG(coords && aug_dev && max_scratch, "null pointer");
  cudaStream_t st = as_stream(stream);
  B2S_CUDA(cudaMemsetAsync(max_scratch, 0x80, (size_t)num_plots * 2 * sizeof(int32_t), st));   // 0x80808080 < any coord
  coords_max_kernel<<<grid_for(ceil_div64(m, PP_CHUNK), 256), 256, 0, st>>>(coords, m, m_dev, max_scratch);
  coords_augment_kernel<<<grid_for(m, 256), 256, 0, st>>>(coords, m, m_dev, aug_dev, max_scratch);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}
