// wgrad_tc.cu -- tcgen05 (kind::tf32) weight gradient of the sparse convolution.
//
//   gw[k, ci, co] = sum over out rows o of  x[nbr[k, o], ci] * gy[o, co]
//
// As a GEMM per kernel offset k: D[M = ci, N = co] with the reduction (UMMA "K") running over OUT ROWS.
// Both operands are gathered feature rows, i.e. M/N-contiguous ("MN-major") in shared memory, which for
// 32-bit elements requires the SWIZZLE_128B_BASE32B layout (see tc_ptx.cuh):
//   A stage [128 ci x 32 rows]: block j = ci/32 at j*4096, inside it row r at r*128 (atoms of 4 rows),
//   B stage [BN  co x 32 rows]: same shape with j = co/32,
//   the four 32-byte units of every 128-byte row XOR-ed with (r & 3).  One tcgen05.mma (M=128, N=BN, K=8)
//   consumes 8 rows (two atoms); descriptors carry LBO = 4096 (next 32 channels), SBO = 512 (next 4 rows).
// Rows without a neighbour at offset k are zero-filled by LDGSTS (no global read for either operand).
// The out rows are split across CTAs (grid.y); partial tiles are combined with vector fp32 reductions.
//
// Reference call site: autograd of MinkowskiConvolution (R:models/base_model.py:262 -> loss.backward()).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace {

using namespace tc;

constexpr int WG_BM = 128;        // input channels per CTA == UMMA M
constexpr int WG_ROWS = 32;       // out rows per pipeline stage (4 MMAs of K = 8 rows)
constexpr int WG_A_STAGE = WG_BM * WG_ROWS * 4;   // 16 KB
constexpr int WG_PRODUCERS = 128;
constexpr int WG_THREADS = 160;
constexpr int WG_LAG = 2;
constexpr uint32_t ATOM_BYTES = 1024;             // 8 rows x 128 B consumed per MMA (two 4-row swizzle atoms)
constexpr uint32_t SBO_BYTES = 512;               // next 4-row atom along the reduction
constexpr uint32_t LBO_BYTES = WG_ROWS * 128;     // next 32-channel block

template <int BN, int STAGES>
struct WgSmem {
  static constexpr int B_STAGE = BN * WG_ROWS * 4;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * WG_A_STAGE;
  static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 2) * 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// byte offset of 16-byte chunk `chunk` (of a feature row piece) for stage-local row `row`
__device__ __forceinline__ uint32_t mn_offset(int row, int chunk) {
  const int j = chunk >> 3, c = chunk & 7;
  const int unit = (c >> 1) ^ (row & 3);          // 32-byte unit swizzle (Swizzle<2,5,2>)
  return (uint32_t)j * LBO_BYTES + (uint32_t)row * 128u + (uint32_t)(((unit << 1) | (c & 1)) << 4);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1)
    wgrad_tc_kernel(const float* __restrict__ x, const float* __restrict__ gy, const int* __restrict__ nbr,
                    int64_t n_out, const int* __restrict__ n_out_dev, int c_in, int c_out, int ci_tiles, int co_tiles,
                    int64_t rows_per_split, int use_atomic, float* __restrict__ gw) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  using L = WgSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base + L::A_OFF, b_base = base + L::B_OFF, bar_base = base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int t = blockIdx.x;
  const int cot = t % co_tiles;
  t /= co_tiles;
  const int cit = t % ci_tiles;
  const int k = t / ci_tiles;
  const int ci0 = cit * WG_BM, co0 = cot * BN;
  const int ci_valid = min(WG_BM, c_in - ci0);
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(r_begin + rows_per_split, n_out);
  const int T = (int)((r_end - r_begin + WG_ROWS - 1) / WG_ROWS);
  if (T <= 0) return;  // uniform across the CTA

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), WG_PRODUCERS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp < 4) {
    const int a_chunks = ci_valid >> 2;  // 16-byte chunks of the A row piece (multiple of 8)
    auto publish = [&](int it_done) {    // round this thread's own chunks of stage it_done to tf32, then hand over
      const uint32_t a_st = a_base + (it_done % STAGES) * WG_A_STAGE, b_st = b_base + (it_done % STAGES) * L::B_STAGE;
#pragma unroll
      for (int p = 0; p < WG_ROWS / 4; ++p) {
        const int row = p * 4 + warp;
        if (lane < a_chunks) round_chunk_tf32(a_st + mn_offset(row, lane));
#pragma unroll
        for (int q = 0; q < (BN + 127) / 128; ++q) {
          const int chunk = q * 32 + lane;
          if (chunk < BN / 4) round_chunk_tf32(b_st + mn_offset(row, chunk));
        }
      }
      fence_proxy_async();
      mbar_arrive(full_bar(it_done % STAGES));
    };
    for (int it = 0; it < T; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t a_stage = a_base + s * WG_A_STAGE, b_stage = b_base + s * L::B_STAGE;
      const int64_t r0 = r_begin + (int64_t)it * WG_ROWS;
#pragma unroll
      for (int p = 0; p < WG_ROWS / 4; ++p) {
        const int row = p * 4 + warp;
        const int64_t o = r0 + row;
        int i = -1;
        if (o < r_end) i = nbr ? __ldg(&nbr[(int64_t)k * pitch + o]) : (int)o;
        const uint32_t nbytes = i >= 0 ? 16u : 0u;
        const int64_t xi = i >= 0 ? i : 0, oo = i >= 0 ? o : 0;
        if (lane < a_chunks) cp_async16(a_stage + mn_offset(row, lane), x + xi * c_in + ci0 + lane * 4, nbytes);
#pragma unroll
        for (int q = 0; q < (BN + 127) / 128; ++q) {
          const int chunk = q * 32 + lane;
          if (chunk < BN / 4) cp_async16(b_stage + mn_offset(row, chunk), gy + oo * c_out + co0 + chunk * 4, nbytes);
        }
      }
      cp_async_commit();
      if (it >= WG_LAG) {
        cp_async_wait<WG_LAG>();
        publish(it - WG_LAG);
      }
    }
    if (T >= 2) {
      cp_async_wait<1>();
      publish(T - 2);
    }
    cp_async_wait<0>();
    publish(T - 1);

    // ---- epilogue: thread = one input channel (TMEM lane), 32 output channels per tcgen05.ld
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int ci = warp * 32 + lane;
    const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
    float* dst_row = gw + ((int64_t)k * c_in + ci0 + ci) * c_out + co0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
      if (ci < ci_valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (use_atomic)
            red_add_v4(dst_row + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                       __uint_as_float(v[j + 3]));
          else
            *reinterpret_cast<float4*>(dst_row + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                            __uint_as_float(v[j + 3]));
        }
      }
    }
    tc_fence_before();
  } else {
    constexpr uint32_t IDESC = idesc_tf32(WG_BM, BN, 1, 1);  // both operands MN-major
    for (int it = 0; it < T; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int g = 0; g < WG_ROWS / 8; ++g) {
          const uint64_t a_desc = smem_desc_sw128_base32(a_base + s * WG_A_STAGE + g * ATOM_BYTES, LBO_BYTES, SBO_BYTES);
          const uint64_t b_desc = smem_desc_sw128_base32(b_base + s * L::B_STAGE + g * ATOM_BYTES, LBO_BYTES, SBO_BYTES);
          mma_tf32(tmem_d, a_desc, b_desc, IDESC, (it | g) ? 1u : 0u);
        }
        mma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (lane == 0) mma_commit(accum_bar);
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<BN>(tmem_d);
  }
}

// ---------------------------------------------------------------------------------------------
// c_in <= 4 (the k7 stem, c_in = 3):  D[M = co, N = 64 kernel offsets x 4 padded input channels] over out rows.
//   A stage = gy rows (M/N-major, 128 output channels), B stage = for every row the 64 neighbours'
//   4-float feature vectors (one 16-byte LDGSTS per (row, offset), zero-filled where nbr = -1).
// grid.x = offset groups of 64 x output-channel tiles of 128, grid.y = row splits.
// ---------------------------------------------------------------------------------------------
constexpr int SM_BN = 256;
constexpr int SM_B_STAGE = SM_BN * WG_ROWS * 4;   // 32 KB
template <int STAGES>
struct WgSmallSmem {
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * WG_A_STAGE;
  static constexpr int BAR_OFF = B_OFF + STAGES * SM_B_STAGE;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 2) * 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1)
    wgrad_small_tc_kernel(const float4* __restrict__ x4, const float* __restrict__ gy, const int* __restrict__ nbr,
                          int64_t n_out, const int* __restrict__ n_out_dev, int c_in, int c_out, int k3, int co_tiles,
                          int64_t rows_per_split, float* __restrict__ gw) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  using L = WgSmallSmem<STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base + L::A_OFF, b_base = base + L::B_OFF, bar_base = base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cot = blockIdx.x % co_tiles;
  const int k0 = (blockIdx.x / co_tiles) * 64;
  const int co0 = cot * WG_BM;
  const int co_valid = min(WG_BM, c_out - co0);
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(r_begin + rows_per_split, n_out);
  const int T = (int)((r_end - r_begin + WG_ROWS - 1) / WG_ROWS);
  if (T <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), WG_PRODUCERS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<SM_BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp < 4) {
    const int a_chunks = co_valid >> 2;
    auto publish = [&](int it_done) {    // gy chunks are rounded here; x4 was rounded by the padding kernel
      const uint32_t a_st = a_base + (it_done % STAGES) * WG_A_STAGE;
#pragma unroll
      for (int p = 0; p < WG_ROWS / 4; ++p)
        if (lane < a_chunks) round_chunk_tf32(a_st + mn_offset(p * 4 + warp, lane));
      fence_proxy_async();
      mbar_arrive(full_bar(it_done % STAGES));
    };
    const int kA = k0 + lane, kB = k0 + 32 + lane;       // the two kernel offsets this lane gathers
    for (int it = 0; it < T; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t a_stage = a_base + s * WG_A_STAGE, b_stage = b_base + s * SM_B_STAGE;
      const int64_t r0 = r_begin + (int64_t)it * WG_ROWS;
#pragma unroll
      for (int p = 0; p < WG_ROWS / 4; ++p) {
        const int row = p * 4 + warp;
        const int64_t o = r0 + row;
        const bool live = o < r_end;
        const int64_t oo = live ? o : 0;
        if (lane < a_chunks) cp_async16(a_stage + mn_offset(row, lane), gy + oo * c_out + co0 + lane * 4, live ? 16u : 0u);
        int ia = -1, ib = -1;
        if (live && kA < k3) ia = __ldg(&nbr[(int64_t)kA * pitch + o]);
        if (live && kB < k3) ib = __ldg(&nbr[(int64_t)kB * pitch + o]);
        cp_async16(b_stage + mn_offset(row, lane), x4 + (ia >= 0 ? ia : 0), ia >= 0 ? 16u : 0u);
        cp_async16(b_stage + mn_offset(row, lane + 32), x4 + (ib >= 0 ? ib : 0), ib >= 0 ? 16u : 0u);
      }
      cp_async_commit();
      if (it >= WG_LAG) {
        cp_async_wait<WG_LAG>();
        publish(it - WG_LAG);
      }
    }
    if (T >= 2) {
      cp_async_wait<1>();
      publish(T - 2);
    }
    cp_async_wait<0>();
    publish(T - 1);

    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int co = warp * 32 + lane;
    const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < SM_BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
      if (co < co_valid) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int col = c0 + jj;
          const int k = k0 + (col >> 2), ci = col & 3;
          if (ci < c_in && k < k3) atomicAdd(&gw[((int64_t)k * c_in + ci) * c_out + co0 + co], __uint_as_float(v[jj]));
        }
      }
    }
    tc_fence_before();
  } else {
    constexpr uint32_t IDESC = idesc_tf32(WG_BM, SM_BN, 1, 1);
    for (int it = 0; it < T; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int g = 0; g < WG_ROWS / 8; ++g) {
          const uint64_t a_desc = smem_desc_sw128_base32(a_base + s * WG_A_STAGE + g * ATOM_BYTES, LBO_BYTES, SBO_BYTES);
          const uint64_t b_desc = smem_desc_sw128_base32(b_base + s * SM_B_STAGE + g * ATOM_BYTES, LBO_BYTES, SBO_BYTES);
          mma_tf32(tmem_d, a_desc, b_desc, IDESC, (it | g) ? 1u : 0u);
        }
        mma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (lane == 0) mma_commit(accum_bar);
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<SM_BN>(tmem_d);
  }
}

__global__ void __launch_bounds__(256) wg_pad_rows4_kernel(const float* __restrict__ x, int64_t n, int c,
                                                           float4* __restrict__ x4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < c; ++j) v[j] = __uint_as_float(rna_tf32(__float_as_uint(x[i * c + j])));
    x4[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

int launch_wgrad_small(const float* x, const float* gy, const int* nbr, int64_t n_in, int64_t n_out, const int* n_out_dev,
                       int c_in, int c_out, int k3, float* gw, void* workspace, cudaStream_t st) {
  constexpr int STAGES = 4;
  using L = WgSmallSmem<STAGES>;
  auto kern = wgrad_small_tc_kernel<STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) != cudaSuccess) {
      b2s_set_error("wgrad_small_tc: cannot opt in to %d bytes of shared memory", L::DYN_BYTES);
      return -1;
    }
    attr_set = true;
  }
  float4* x4 = reinterpret_cast<float4*>(workspace);
  wg_pad_rows4_kernel<<<grid_for(n_in, 256), 256, 0, st>>>(x, n_in, c_in, x4);
  const int co_tiles = (c_out + WG_BM - 1) / WG_BM, groups = (k3 + 63) / 64;
  const int64_t base = (int64_t)groups * co_tiles;
  int64_t splits = (3LL * B2S_NUM_SMS + base - 1) / base;
  const int64_t max_splits = ceil_div64(n_out, 8 * WG_ROWS);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rows = ceil_div64(n_out, splits);
  rows = ceil_div64(rows, WG_ROWS) * WG_ROWS;
  splits = ceil_div64(n_out, rows);
  cudaMemsetAsync(gw, 0, (size_t)k3 * c_in * c_out * sizeof(float), st);
  dim3 grid((unsigned)base, (unsigned)splits);
  kern<<<grid, WG_THREADS, L::DYN_BYTES, st>>>(x4, gy, nbr, n_out, n_out_dev, c_in, c_out, k3, co_tiles, rows, gw);
  return 0;
}

bool wgrad_tc_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2S_DISABLE_TC");
    const char* w = getenv("B2S_DISABLE_TC_WGRAD");
    v = ((e && e[0] == '1') || (w && w[0] == '1')) ? 1 : 0;
  }
  return v == 1;
}

template <int BN, int STAGES>
int launch_wgrad(const float* x, const float* gy, const int* nbr, int64_t n_out, const int* n_out_dev, int c_in, int c_out,
                 int k3, float* gw, cudaStream_t st) {
  using L = WgSmem<BN, STAGES>;
  auto kern = wgrad_tc_kernel<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) != cudaSuccess) {
      b2s_set_error("wgrad_tc: cannot opt in to %d bytes of shared memory", L::DYN_BYTES);
      return -1;
    }
    attr_set = true;
  }
  const int ci_tiles = (c_in + WG_BM - 1) / WG_BM, co_tiles = c_out / BN;
  const int64_t base = (int64_t)k3 * ci_tiles * co_tiles;
  // aim at ~3 waves of 148 SMs; every split gets at least 8 stages of rows
  int64_t splits = (3LL * B2S_NUM_SMS + base - 1) / base;
  const int64_t max_splits = ceil_div64(n_out, 8 * WG_ROWS);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rows = ceil_div64(n_out, splits);
  rows = ceil_div64(rows, WG_ROWS) * WG_ROWS;
  splits = ceil_div64(n_out, rows);
  const int use_atomic = splits > 1 ? 1 : 0;
  if (use_atomic) cudaMemsetAsync(gw, 0, (size_t)k3 * c_in * c_out * sizeof(float), st);
  dim3 grid((unsigned)base, (unsigned)splits);
  kern<<<grid, WG_THREADS, L::DYN_BYTES, st>>>(x, gy, nbr, n_out, n_out_dev, c_in, c_out, ci_tiles, co_tiles, rows,
                                               use_atomic, gw);
  return 0;
}

}  // namespace

bool b2s_wgrad_tc_supported(int32_t c_in, int32_t c_out, int32_t k3, int64_t n_out, bool has_map) {
  (void)k3;
  if (wgrad_tc_disabled() || n_out <= 0) return false;
  if (c_in <= 4) return has_map && c_out % 32 == 0;
  return c_in % 32 == 0 && c_out % 64 == 0;
}

int64_t b2s_wgrad_tc_workspace_bytes(int32_t c_in, int64_t n_in) { return c_in <= 4 ? ((n_in * 16 + 255) & ~(int64_t)255) : 0; }

int b2s_conv_wgrad_tc(const float* x, const float* gy, const int32_t* nbr, int64_t n_in, int64_t n_out,
                      const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, float* gw, void* workspace,
                      cudaStream_t st) {
  if (c_in <= 4) return launch_wgrad_small(x, gy, nbr, n_in, n_out, n_out_dev, c_in, c_out, k3, gw, workspace, st);
  const int bn = c_out % 256 == 0 ? 256 : (c_out % 128 == 0 ? 128 : 64);
  if (bn == 256) return launch_wgrad<256, 4>(x, gy, nbr, n_out, n_out_dev, c_in, c_out, k3, gw, st);
  if (bn == 128) return launch_wgrad<128, 3>(x, gy, nbr, n_out, n_out_dev, c_in, c_out, k3, gw, st);
  return launch_wgrad<64, 4>(x, gy, nbr, n_out, n_out_dev, c_in, c_out, k3, gw, st);
}
