// stem_lines.cu -- convolution of a few input channels (c_in <= 4: the k7 stem) through the X-LINE form of its kernel
// map (b2s_kernel_map_lines, coords.cu), split-bf16 operand mode, tcgen05 kind::f16 with fp32 accumulation in TMEM.
//
// Why a second pair of kernels for one layer: the stem is 1.6 % of the step's FLOPs but was 23 % of its time.  Its 343
// offsets are 89 % empty, and the general output-stationary kernels (conv_tc.cu SMALL mode, wgrad_tc.cu
// wgrad_small_tc_kernel) pay one 16-byte LDGSTS per (row, offset) -- zero-fills included -- at the ~27 B/clk/SM the
// LSU sustains on scattered rows, on top of a 580 MB [343, N] table written once and read twice per step.  Here
//   * the map is one word per (row, LINE of 7 x-consecutive offsets): (base << 8) | mask.  Rows are sorted by
//     (plot, z, y, x), so the existing neighbours of a line are the consecutive rows base, base+1, ...;
//   * a pipeline stage is one line: K = 8 slots (7 offsets + a zero slot) x 8 bf16 [h0..h3 | l0..l3] = one 128-byte
//     swizzle row per out row.  A producer thread owns one out row: it loads the line word, issues ONLY the existing
//     neighbours' 16-byte operand rows (ld.global.nc, through L1: x-adjacent rows share them) a few stages ahead into
//     registers and writes the 8 slots with st.shared.v4 (128 B/clk/SM) -- no zero-fill traffic to L2 at all;
//   * forward: M = 256 rows per CTA (two accumulator tiles share every 24 KB weight stage [H|H], [M|M], [L|0]; one
//     N = 192 MMA per K step and tile; the epilogue adds the three column groups);
//   * wgrad: D[M = [gh; gl] (64 + 64 output channels), N = up to 8 lines x 64] over out rows, both operands MN-major,
//     two accumulators (N <= 256 each) per CTA so that the gy rows are staged once per 7-8 lines.
// Both are then bound by the tensor pipe (the three-term weight split is 3x the MMA work of a TF32 product), not by
// the gather.  Results are those of the table-driven kernels: same operand splits, same products, fp32 accumulation.
//
// Reference call site: R:modules/MinkowskiEngine/SENet.py:49-52 (ME.MinkowskiConvolution(3, 64, kernel_size=7)).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

void b2s_launch_col_partials(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* col_stats, cudaStream_t st);

namespace {

using namespace tc;

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// operand rows of the input: [h0 h1 h2 h3 | l0 l1 l2 l3] bf16 per row (channels >= c are zero)
__global__ void __launch_bounds__(256) lines_pad_rows_kernel(const float* __restrict__ x, int64_t n, int c,
                                                             uint4* __restrict__ x4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
    for (int j = 0; j < c; ++j) split_bf16(x[i * c + j], h[j], l[j]);
    x4[i] = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), l[0] | (l[1] << 16), l[2] | (l[3] << 16));
  }
}

// weight image: img[(line * 3 + j) * c_out + n] = 128-byte row of 64 bf16, element kq = slot * 8 + s8 (ci = s8 & 3),
// 16-byte chunks XOR-swizzled by (n & 7).  W = H + M + L (24 significant bits): j = 0 [H | H], j = 1 [M | M],
// j = 2 [L | 0], against operand rows [h | l]: h*H + l*H + h*M + l*M + h*L.
__global__ void __launch_bounds__(256) lines_prep_weights_kernel(const float* __restrict__ w, int c_in, int c_out,
                                                                 int kx, int nlines, uint32_t* __restrict__ img) {
  const int64_t total = (int64_t)nlines * 3 * c_out * 32;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int pos = (int)(e & 31);
    const int64_t rn = e >> 5;
    const int n = (int)(rn % c_out);
    const int lj = (int)(rn / c_out);
    const int line = lj / 3, j = lj % 3;
    const int slot = (pos >> 2) ^ (n & 7);       // logical chunk (= slot) stored at this physical position
    uint32_t out[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int s8 = (pos & 3) * 2 + q, ci = s8 & 3;
      float v = 0.f;
      if (slot < kx && ci < c_in) v = w[((int64_t)(line * kx + slot) * c_in + ci) * c_out + n];
      uint32_t h, m, l;
      split_bf16x3(v, h, m, l);
      out[q] = j == 0 ? h : (j == 1 ? m : (s8 < 4 ? l : 0u));
    }
    img[e] = out[0] | (out[1] << 16);
  }
}

// the operand rows of one (out row, line): slot s holds the row of offset s of the line, or zeros
__device__ __forceinline__ void load_line(const uint4* __restrict__ x4, uint32_t word, uint4 (&dst)[8]) {
  const uint32_t mask = word & 0xffu;
  const uint4* src = x4 + (word >> 8);
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if ((mask >> s) & 1u) v = __ldg(src + __popc(mask & ((1u << s) - 1u)));
    dst[s] = v;
  }
}

// =============================================================================================
// forward
// =============================================================================================
constexpr int LF_TILE = 128 * 128;           // one 128-row A tile of a stage
constexpr int LF_A_STAGE = 2 * LF_TILE;      // M = 256
constexpr int LF_BN = 64;
constexpr int LF_B_IMG = LF_BN * 128;
constexpr int LF_B_STAGE = 3 * LF_B_IMG;     // 24 KB
constexpr int LF_THREADS = 288;              // 8 producer / epilogue warps + the MMA warp

template <int STAGES>
struct LfSmem {
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * LF_A_STAGE;
  static constexpr int BAR_OFF = B_OFF + STAGES * LF_B_STAGE;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 2) * 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <int STAGES, int DEPTH>
__global__ void __launch_bounds__(LF_THREADS, 1)
    conv_lines_fwd_kernel(const uint4* __restrict__ x4, const uint32_t* __restrict__ wimg,
                          const float* __restrict__ bias, const uint32_t* __restrict__ lines, int64_t n_out,
                          const int* __restrict__ n_out_dev, int c_out, int nlines, float* __restrict__ y) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  const int64_t m0 = (int64_t)blockIdx.x * 256;
  if (m0 >= n_out) return;                       // uniform across the CTA
  const int n0 = blockIdx.y * LF_BN;
  using L = LfSmem<STAGES>;
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
  const int T = nlines;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 256 + 1);           // 256 producer rows + 1 arrive.expect_tx for the weight images
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);     // tile 0: columns 0..191, tile 1: 256..447
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp < 8) {
    // ===================== producers: thread = out row =====================
    const int64_t o = m0 + tid;
    const bool live = o < n_out;
    const uint32_t* lt = lines + o;
    const uint32_t row_off = (uint32_t)(tid >> 7) * LF_TILE + (uint32_t)(tid & 127) * 128u;
    const uint32_t r7 = (uint32_t)tid & 7u;
    auto ldw = [&](int l) -> uint32_t { return (live && l < T) ? __ldg(lt + (int64_t)l * pitch) : 0u; };
    uint4 d[DEPTH][8];
    uint32_t wq[DEPTH];
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) wq[j] = ldw(j);
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) {
      load_line(x4, wq[j], d[j]);
      wq[j] = ldw(j + DEPTH);
    }
    int s = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int it0 = 0; it0 < T; it0 += DEPTH) {
#pragma unroll
      for (int j = 0; j < DEPTH; ++j) {
        const int it = it0 + j;
        if (it < T) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (tid == 0) {
            mbar_arrive_expect_tx(full_bar(s), LF_B_STAGE);
#pragma unroll
            for (int i = 0; i < 3; ++i)
              bulk_g2s(b_base + s * LF_B_STAGE + i * LF_B_IMG,
                       wimg + (((int64_t)it * 3 + i) * c_out + n0) * 32, LF_B_IMG, full_bar(s));
          }
          const uint32_t dst = a_base + (uint32_t)s * LF_A_STAGE + row_off;
#pragma unroll
          for (int q = 0; q < 8; ++q) sts128(dst + (((uint32_t)q ^ r7) << 4), d[j][q]);
          fence_proxy_async();
          mbar_arrive(full_bar(s));
          load_line(x4, wq[j], d[j]);            // stage it + DEPTH
          wq[j] = ldw(it + 2 * DEPTH);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
    // ===================== epilogue: warp = (tile, lane quadrant) =====================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q4 = warp & 3, tile = warp >> 2;
    const int64_t orow = m0 + tile * 128 + q4 * 32 + lane;
    const uint32_t t_lane = tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(tile * 256);
#pragma unroll 1
    for (int c0 = 0; c0 < LF_BN; c0 += 32) {
      uint32_t v[32], u[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 1; i < 3; ++i) {              // the column groups of the [M | M] and [L | 0] images
        tmem_ld32(t_lane + (uint32_t)(i * LF_BN + c0), u);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
      }
      if (orow < n_out) {
        float* dst = y + orow * c_out + n0 + c0;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float4 r;
          r.x = __uint_as_float(v[e]) + (bias ? __ldg(&bias[n0 + c0 + e]) : 0.f);
          r.y = __uint_as_float(v[e + 1]) + (bias ? __ldg(&bias[n0 + c0 + e + 1]) : 0.f);
          r.z = __uint_as_float(v[e + 2]) + (bias ? __ldg(&bias[n0 + c0 + e + 2]) : 0.f);
          r.w = __uint_as_float(v[e + 3]) + (bias ? __ldg(&bias[n0 + c0 + e + 3]) : 0.f);
          *reinterpret_cast<float4*>(dst + e) = r;
        }
      }
    }
    tc_fence_before();
  } else {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC = idesc_bf16(128, 3 * LF_BN, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t b_desc = smem_desc_sw128(b_base + s * LF_B_STAGE, 16, 1024);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint64_t a_desc = smem_desc_sw128(a_base + s * LF_A_STAGE + t * LF_TILE, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)         // K = 16 bf16 = two slots per MMA
            mma_bf16(tmem_d + (uint32_t)(t * 256), a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), IDESC,
                     (it | kk) ? 1u : 0u);
        }
        mma_commit(empty_bar(s));
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (elect_one()) mma_commit(accum_bar);
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_d);
  }
}

// ---------------------------------------------------------------------------------------------
// forward, operand tile in TENSOR MEMORY.  The kernel above is bound by shared-memory bandwidth, not by the tensor
// pipe: per stage the MMAs read 2 x 4 x (4 KB of A + 6 KB of B) = 80 KB of operands from shared memory, the producers
// write 32 KB of A and the bulk copy 24 KB of B -- 136 KB at 128 B/clk = 1 090 clocks against 768 clocks of MMA.
// Here a producer thread (= out row = TMEM lane) writes its 128-byte operand row straight into tensor memory
// (tcgen05.st 32x32b.x32, 256 B/clk) and the MMA takes A from TMEM, so shared memory carries the weight stages only
// (48 KB of reads + 24 KB of writes per stage).  TMEM: accumulators 2 x 192 columns, operand ring 2 stages x 2 tiles x
// 32 columns = 512.  Weights: a ring of SB stages filled by their own warp (the bulk copies need ~1 300 clocks, more
// than one stage of MMA).
// ---------------------------------------------------------------------------------------------
constexpr int LT_THREADS = 320;              // 8 producer / epilogue warps, warp 8 = MMA issuer, warp 9 = weight loader
constexpr int LT_SA = 2;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint4 (&d)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(d[0].x), "r"(d[0].y), "r"(d[0].z), "r"(d[0].w), "r"(d[1].x), "r"(d[1].y), "r"(d[1].z), "r"(d[1].w),
      "r"(d[2].x), "r"(d[2].y), "r"(d[2].z), "r"(d[2].w), "r"(d[3].x), "r"(d[3].y), "r"(d[3].z), "r"(d[3].w),
      "r"(d[4].x), "r"(d[4].y), "r"(d[4].z), "r"(d[4].w), "r"(d[5].x), "r"(d[5].y), "r"(d[5].z), "r"(d[5].w),
      "r"(d[6].x), "r"(d[6].y), "r"(d[6].z), "r"(d[6].w), "r"(d[7].x), "r"(d[7].y), "r"(d[7].z), "r"(d[7].w)
      : "memory");
}
template <int SB>
struct LtSmem {
  static constexpr int B_OFF = 0;
  static constexpr int BAR_OFF = SB * LF_B_STAGE;
  static constexpr int NBAR = 2 * LT_SA + 2 * SB + 2;
  static constexpr int TOTAL = BAR_OFF + NBAR * 8;
  static constexpr int DYN_BYTES = (TOTAL + 1024) > 120 * 1024 ? (TOTAL + 1024) : 120 * 1024;   // one CTA per SM (TMEM)
};

template <int SB, int DEPTH>
__global__ void __launch_bounds__(LT_THREADS, 1)
    conv_lines_fwd_tmem_kernel(const uint4* __restrict__ x4, const uint32_t* __restrict__ wimg,
                               const float* __restrict__ bias, const uint32_t* __restrict__ lines, int64_t n_out,
                               const int* __restrict__ n_out_dev, int c_out, int nlines, float* __restrict__ y,
                               float* __restrict__ col_stats) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  const int64_t m0 = (int64_t)blockIdx.x * 256;
  if (m0 >= n_out) return;                       // uniform across the CTA
  const int n0 = blockIdx.y * LF_BN;
  using L = LtSmem<SB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t b_base = base + L::B_OFF, bar_base = base + L::BAR_OFF;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (LT_SA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * LT_SA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * LT_SA + SB + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * LT_SA + 2 * SB);
  const uint32_t tmem_slot = bar_base + 8u * (2 * LT_SA + 2 * SB + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * LT_SA + 2 * SB + 1));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = nlines;

  if (tid == 0) {
    for (int s = 0; s < LT_SA; ++s) {
      mbar_init(a_full(s), 256);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);     // D tile 0: columns 0..191, D tile 1: 192..383, operands: 384..511
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;
  constexpr uint32_t A_COL = 384;

  if (warp < 8) {
    // ===================== producers: thread = out row = TMEM lane =====================
    const int64_t o = m0 + tid;
    const bool live = o < n_out;
    const uint32_t* lt = lines + o;
    const int tile = warp >> 2;
    const uint32_t t_row = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    auto ldw = [&](int l) -> uint32_t { return (live && l < T) ? __ldg(lt + (int64_t)l * pitch) : 0u; };
    uint4 d[DEPTH][8];
    uint32_t wq[DEPTH];
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) wq[j] = ldw(j);
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) {
      load_line(x4, wq[j], d[j]);
      wq[j] = ldw(j + DEPTH);
    }
#pragma unroll 1
    for (int it0 = 0; it0 < T; it0 += DEPTH) {
#pragma unroll
      for (int j = 0; j < DEPTH; ++j) {
        const int it = it0 + j;
        if (it < T) {
          const int sa = it & 1;
          mbar_wait(a_empty(sa), (((uint32_t)it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          tmem_st32(t_row + A_COL + (uint32_t)((sa * 2 + tile) * 32), d[j]);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(a_full(sa));
          load_line(x4, wq[j], d[j]);            // stage it + DEPTH
          wq[j] = ldw(it + 2 * DEPTH);
        }
      }
    }
    // ===================== epilogue: warp = (tile, lane quadrant) =====================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int64_t orow = m0 + tid;
    const uint32_t t_lane = t_row + (uint32_t)(tile * 192);
#pragma unroll 1
    for (int c0 = 0; c0 < LF_BN; c0 += 32) {
      uint32_t v[32], u[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 1; i < 3; ++i) {
        tmem_ld32(t_lane + (uint32_t)(i * LF_BN + c0), u);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
      }
      if (col_stats) {     // batch-norm statistics of the output, accumulated while the tile is in registers
        float r[32];
#pragma unroll
        for (int e = 0; e < 32; ++e)
          r[e] = orow < n_out ? __uint_as_float(v[e]) + (bias ? __ldg(&bias[n0 + c0 + e]) : 0.f) : 0.f;
        epilogue_col_stats(r, lane, reinterpret_cast<float2*>(smem + L::B_OFF) + warp * LF_BN + c0);
      }
      if (orow < n_out) {
        float* dst = y + orow * c_out + n0 + c0;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float4 r;
          r.x = __uint_as_float(v[e]) + (bias ? __ldg(&bias[n0 + c0 + e]) : 0.f);
          r.y = __uint_as_float(v[e + 1]) + (bias ? __ldg(&bias[n0 + c0 + e + 1]) : 0.f);
          r.z = __uint_as_float(v[e + 2]) + (bias ? __ldg(&bias[n0 + c0 + e + 2]) : 0.f);
          r.w = __uint_as_float(v[e + 3]) + (bias ? __ldg(&bias[n0 + c0 + e + 3]) : 0.f);
          *reinterpret_cast<float4*>(dst + e) = r;
        }
      }
    }
    if (col_stats) {       // the eight epilogue warps (two row tiles) combine their per-column sums
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2* red = reinterpret_cast<const float2*>(smem + L::B_OFF);
      if (tid < LF_BN) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) {
          s1 += red[wq * LF_BN + tid].x;
          s2 += red[wq * LF_BN + tid].y;
        }
        float* part = col_stats + (int64_t)blockIdx.x * 2 * c_out;     // this 256-row tile's partial row
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) col_stats[((pitch + 127) / 128) * 2 * c_out] = 256.f;
        part[n0 + tid] = s1;
        part[c_out + n0 + tid] = s2;
      }
    }
    tc_fence_before();
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC = idesc_bf16(128, 3 * LF_BN, 0, 0);
    int sb = 0;
    uint32_t phb = 0;
    for (int it = 0; it < T; ++it) {
      const int sa = it & 1;
      mbar_wait(b_full(sb), phb);
      mbar_wait(a_full(sa), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t b_desc = smem_desc_sw128(b_base + sb * LF_B_STAGE, 16, 1024);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t a_t = tmem_d + A_COL + (uint32_t)((sa * 2 + t) * 32);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)         // K = 16 bf16 = 8 TMEM columns = two slots per MMA
            mma_bf16_ta(tmem_d + (uint32_t)(t * 192), a_t + (uint32_t)(kk * 8), b_desc + (uint64_t)(kk * 2), IDESC,
                        (it | kk) ? 1u : 0u);
        }
        mma_commit(a_empty(sa));
        mma_commit(b_empty(sb));
      }
      __syncwarp();
      if (++sb == SB) {
        sb = 0;
        phb ^= 1u;
      }
    }
    if (elect_one()) mma_commit(accum_bar);
    __syncwarp();
  } else if (elect_one()) {
    // ===================== weight loader (warp 9, one lane) =====================
    int sb = 0;
    uint32_t phb = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(b_empty(sb), phb ^ 1u);
      mbar_arrive_expect_tx(b_full(sb), LF_B_STAGE);
#pragma unroll
      for (int i = 0; i < 3; ++i)
        bulk_g2s(b_base + sb * LF_B_STAGE + i * LF_B_IMG, wimg + (((int64_t)it * 3 + i) * c_out + n0) * 32, LF_B_IMG,
                 b_full(sb));
      if (++sb == SB) {
        sb = 0;
        phb ^= 1u;
      }
    }
  }
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_d);
  }
}

// =============================================================================================
// weight gradient
// =============================================================================================
constexpr int LW_ROWS = 32;                  // out rows per stage: two MMAs of K = 16 rows per accumulator
constexpr int LW_A_STAGE = 8192;             // gy rows: [gh atoms | gl atoms], 64 channels x 32 rows x 2 B each
constexpr int LW_LINE = 4096;                // one line of a stage: 32 rows x 128 B
constexpr int LW_B_STAGE = 8 * LW_LINE;      // up to 8 lines per CTA
constexpr int LW_STAGE = LW_A_STAGE + LW_B_STAGE;
constexpr int LW_THREADS = 288;

template <int STAGES>
struct LwSmem {
  static constexpr int BAR_OFF = STAGES * LW_STAGE;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 2) * 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <int STAGES, int DEPTH>
__global__ void __launch_bounds__(LW_THREADS, 1)
    wgrad_lines_kernel(const uint4* __restrict__ x4, const float* __restrict__ gy, const uint32_t* __restrict__ lines,
                       int64_t n_out, const int* __restrict__ n_out_dev, int c_in, int c_out, int nlines, int kx,
                       int lpg, int co_tiles, int64_t rows_per_split, float* __restrict__ gw) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  using L = LwSmem<STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * STAGES + 1));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cot = blockIdx.x % co_tiles;
  const int line0 = (blockIdx.x / co_tiles) * lpg;
  const int nl = min(lpg, nlines - line0);       // lines of this CTA (<= 8)
  const int n1 = min(nl, 4), n2 = nl - n1;       // accumulator 1: lines 0..3, accumulator 2: lines 4..
  const int co0 = cot * 64;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(r_begin + rows_per_split, n_out);
  const int T = (int)((r_end - r_begin + LW_ROWS - 1) / LW_ROWS);
  if (T <= 0 || nl <= 0) return;                 // uniform across the CTA

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 256);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp < 8) {
    // ===================== producers: warp = line of the group, lane = row of the stage =====================
    const bool my_line = warp < nl;
    const uint32_t* lt = lines + (int64_t)(line0 + (my_line ? warp : 0)) * pitch;
    // B: MN-major SWIZZLE_128B atoms of 64 N-elements (8 slots x 8 bf16) x 8 rows, 1 KB per 8 rows, 4 KB per line
    const uint32_t b_off = LW_A_STAGE + (uint32_t)warp * LW_LINE + (uint32_t)(lane >> 3) * 1024u + (uint32_t)(lane & 7) * 128u;
    const uint32_t l7 = (uint32_t)lane & 7u;
    // A: the two 16-byte chunks (rows ra, ra + 16) of the stage's gy rows this thread moves.  A gy row of 64 channels
    // in operand form is [h 0..31 | l 0..31 | h 32..63 | l 32..63]; h chunks go to the first 4 KB (M rows 0..63), l
    // chunks to the second (M rows 64..127), each as atoms of 64 channels x 8 rows
    // (chunk order inside the 16 lanes of a row: lanes 0-7 take the eight h chunks of both 32-channel blocks, lanes
    // 8-15 the l chunks, so that a quarter-warp -- the unit a 16-byte shared store is processed in -- writes eight
    // different chunk columns; [block | h/l | chunk] order put h and l of one chunk, 4 KB apart, on the same banks)
    const int l16 = tid & 15, ra = tid >> 4;
    const int c16 = ((l16 & 4) << 1) | ((l16 & 8) >> 1) | (l16 & 3);
    const uint32_t a_chunk = (uint32_t)((((c16 >> 3) & 1) << 2) | (c16 & 3));
    auto a_off = [&](int row) -> uint32_t {
      return (uint32_t)((c16 >> 2) & 1) * 4096u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
             ((a_chunk ^ (uint32_t)(row & 7)) << 4);
    };
    const uint32_t a_off0 = a_off(ra), a_off1 = a_off(ra + 16);
    const uint4* gy4 = reinterpret_cast<const uint4*>(gy + co0) + c16;
    const int64_t gy_pitch4 = c_out / 4;
    auto ldw = [&](int it) -> uint32_t {
      const int64_t o = r_begin + (int64_t)it * LW_ROWS + lane;
      return (my_line && it < T && o < r_end) ? __ldg(lt + o) : 0u;
    };
    auto ldg = [&](int it, uint4 (&g)[2]) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t o = r_begin + (int64_t)it * LW_ROWS + ra + 16 * i;
        g[i] = (it < T && o < r_end) ? __ldg(gy4 + o * gy_pitch4) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    uint4 d[DEPTH][8], g[DEPTH][2];
    uint32_t wq[DEPTH];
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) wq[j] = ldw(j);
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) {
      load_line(x4, wq[j], d[j]);
      ldg(j, g[j]);
      wq[j] = ldw(j + DEPTH);
    }
    int s = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int it0 = 0; it0 < T; it0 += DEPTH) {
#pragma unroll
      for (int j = 0; j < DEPTH; ++j) {
        const int it = it0 + j;
        if (it < T) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t st = base + (uint32_t)s * LW_STAGE;
#pragma unroll
          for (int q = 0; q < 8; ++q) sts128(st + b_off + (((uint32_t)q ^ l7) << 4), d[j][q]);
          sts128(st + a_off0, g[j][0]);
          sts128(st + a_off1, g[j][1]);
          fence_proxy_async();
          mbar_arrive(full_bar(s));
          load_line(x4, wq[j], d[j]);            // stage it + DEPTH
          ldg(it + DEPTH, g[j]);
          wq[j] = ldw(it + 2 * DEPTH);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
    // ===================== epilogue: TMEM lane = [gh co | gl co], column = (line, slot, [h ci | l ci]) ==========
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q4 = warp & 3, acc = warp >> 2;      // warps 0-3 read accumulator 1, warps 4-7 accumulator 2
    const int co = (q4 * 32 + lane) & 63;
    const int ncols = (acc ? n2 : n1) * 64;
    const uint32_t t_lane = tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
      const int line = line0 + acc * 4 + (c0 >> 6);
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const int slot = ((c0 & 63) >> 3) + g8;
        if (slot < kx) {
          float* dst = gw + ((int64_t)(line * kx + slot) * c_in) * c_out + co0 + co;
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
            if (ci < c_in)
              atomicAdd(dst + (int64_t)ci * c_out, __uint_as_float(v[g8 * 8 + ci]) + __uint_as_float(v[g8 * 8 + 4 + ci]));
        }
      }
    }
    tc_fence_before();
  } else {
    // ===================== MMA issuer =====================
    const uint32_t idesc1 = idesc_bf16(128, n1 * 64, 1, 1);   // both operands MN-major
    const uint32_t idesc2 = idesc_bf16(128, n2 > 0 ? n2 * 64 : 64, 1, 1);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_stage = base + (uint32_t)s * LW_STAGE, b_stage = a_stage + LW_A_STAGE;
#pragma unroll
        for (int k16 = 0; k16 < LW_ROWS / 16; ++k16) {
          const uint64_t a_desc = smem_desc_sw128(a_stage + k16 * 2048, 4096, 1024);
          mma_bf16(tmem_d, a_desc, smem_desc_sw128(b_stage + k16 * 2048, LW_LINE, 1024), idesc1, (it | k16) ? 1u : 0u);
          if (n2 > 0)
            mma_bf16(tmem_d + 256u, a_desc, smem_desc_sw128(b_stage + 4 * LW_LINE + k16 * 2048, LW_LINE, 1024), idesc2,
                     (it | k16) ? 1u : 0u);
        }
        mma_commit(empty_bar(s));
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (elect_one()) mma_commit(accum_bar);
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_d);
  }
}

bool lines_tc_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2S_DISABLE_TC");
    const char* l = getenv("B2S_DISABLE_LINES");
    v = ((e && e[0] == '1') || (l && l[0] == '1')) ? 1 : 0;
  }
  return v == 1;
}

int lines_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

inline int64_t al256(int64_t b) { return (b + 255) & ~(int64_t)255; }
inline int64_t image_bytes(int c_out, int nlines) { return al256((int64_t)nlines * 3 * c_out * 128); }

}  // namespace

extern "C" int32_t b2s_conv_lines_supported(int32_t c_in, int32_t c_out, const int32_t* ks) {
  if (!ks || lines_tc_disabled() || !b2s_precise()) return 0;
  if (c_in < 1 || c_in > 4 || c_out < 64 || c_out % 64 != 0) return 0;
  if (ks[0] < 1 || ks[0] > 8 || ks[1] < 1 || ks[2] < 1 || ks[1] * ks[2] > 4096) return 0;
  return 1;
}

extern "C" int64_t b2s_conv_lines_workspace_bytes(int64_t n_in, int32_t c_in, int32_t c_out, const int32_t* ks) {
  if (!ks || n_in < 0 || c_in < 1 || c_in > 4 || c_out < 64 || c_out % 64 != 0) return -1;
  return image_bytes(c_out, ks[1] * ks[2]) + al256(n_in * 16);
}

extern "C" int32_t b2s_conv_lines_fwd(const float* x, const float* w, const float* bias, const uint32_t* lines,
                                      int64_t n_in, int64_t n_out, const int32_t* n_out_dev, int32_t c_in,
                                      int32_t c_out, const int32_t* ks, float* y, void* workspace,
                                      int64_t workspace_bytes, float* col_stats, b2s_stream_t stream) {
  B2S_CHECK_ARG(n_in >= 0 && n_out >= 0 && ks, "bad sizes");
  if (!b2s_conv_lines_supported(c_in, c_out, ks)) {
    b2s_set_error("b2s_conv_lines_fwd: shape c_in=%d c_out=%d or operand mode not covered (see b2s_conv_lines_supported)",
                  c_in, c_out);
    return B2S_EINVAL;
  }
  if (n_out == 0) return B2S_OK;
  B2S_CHECK_ARG(x && w && lines && y && n_in > 0 && n_in < ((int64_t)1 << 24), "null pointer or row count");
  const int nlines = ks[1] * ks[2], kx = ks[0];
  B2S_CHECK_ARG(workspace && workspace_bytes >= b2s_conv_lines_workspace_bytes(n_in, c_in, c_out, ks) &&
                    (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                "workspace too small or misaligned (see b2s_conv_lines_workspace_bytes)");
  cudaStream_t st = as_stream(stream);
  uint32_t* img = reinterpret_cast<uint32_t*>(workspace);
  uint4* x4 = reinterpret_cast<uint4*>(reinterpret_cast<char*>(workspace) + image_bytes(c_out, nlines));
  lines_prep_weights_kernel<<<grid_for((int64_t)nlines * 3 * c_out * 32, 256), 256, 0, st>>>(w, c_in, c_out, kx, nlines,
                                                                                            img);
  lines_pad_rows_kernel<<<grid_for(n_in, 256), 256, 0, st>>>(x, n_in, c_in, x4);
  dim3 grid((unsigned)ceil_div64(n_out, 256), (unsigned)(c_out / LF_BN));
  static const int variant = lines_env("B2S_LINES_FWD", 0);
  bool stats_fused = false;
#define LF_LAUNCH(S, D)                                                                                         \
  do {                                                                                                          \
    auto kern = conv_lines_fwd_kernel<S, D>;                                                                    \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LfSmem<S>::DYN_BYTES));  \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    kern<<<grid, LF_THREADS, LfSmem<S>::DYN_BYTES, st>>>(x4, img, bias, lines, n_out, n_out_dev, c_out, nlines, y); \
  } while (0)
#define LT_LAUNCH(SB, D)                                                                                        \
  do {                                                                                                          \
    auto kern = conv_lines_fwd_tmem_kernel<SB, D>;                                                              \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LtSmem<SB>::DYN_BYTES)); \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    kern<<<grid, LT_THREADS, LtSmem<SB>::DYN_BYTES, st>>>(x4, img, bias, lines, n_out, n_out_dev, c_out, nlines, y, \
                                                          col_stats);                                           \
    stats_fused = true;                                                                                         \
  } while (0)
  if (variant == 1) LF_LAUNCH(3, 2);
  else if (variant == 2) LF_LAUNCH(4, 3);
  else if (variant == 3) LF_LAUNCH(3, 3);
  else if (variant == 4) LT_LAUNCH(4, 3);
  else if (variant == 5) LT_LAUNCH(6, 2);
  else LT_LAUNCH(6, 3);
#undef LF_LAUNCH
#undef LT_LAUNCH
  if (col_stats && !stats_fused) b2s_launch_col_partials(y, n_out, n_out_dev, c_out, col_stats, st);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

extern "C" int32_t b2s_conv_lines_wgrad(const float* x, const float* gy, const uint32_t* lines, int64_t n_in,
                                        int64_t n_out, const int32_t* n_out_dev, int32_t c_in, int32_t c_out,
                                        const int32_t* ks, float* gw, void* workspace, int64_t workspace_bytes,
                                        b2s_stream_t stream) {
  B2S_CHECK_ARG(n_in >= 0 && n_out >= 0 && ks && gw, "bad sizes");
  if (!b2s_conv_lines_supported(c_in, c_out, ks)) {
    b2s_set_error("b2s_conv_lines_wgrad: shape c_in=%d c_out=%d or operand mode not covered", c_in, c_out);
    return B2S_EINVAL;
  }
  cudaStream_t st = as_stream(stream);
  const int nlines = ks[1] * ks[2], kx = ks[0];
  B2S_CUDA(cudaMemsetAsync(gw, 0, (size_t)kx * nlines * c_in * c_out * sizeof(float), st));
  if (n_out == 0) return B2S_OK;
  B2S_CHECK_ARG(x && gy && lines && n_in > 0 && n_in < ((int64_t)1 << 24), "null pointer or row count");
  B2S_CHECK_ARG(workspace && workspace_bytes >= b2s_conv_lines_workspace_bytes(n_in, c_in, c_out, ks) &&
                    (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0,
                "workspace too small or misaligned (see b2s_conv_lines_workspace_bytes)");
  uint4* x4 = reinterpret_cast<uint4*>(reinterpret_cast<char*>(workspace) + image_bytes(c_out, nlines));
  lines_pad_rows_kernel<<<grid_for(n_in, 256), 256, 0, st>>>(x, n_in, c_in, x4);
  const int groups = (nlines + 7) / 8, lpg = (nlines + groups - 1) / groups, co_tiles = c_out / 64;
  const int64_t base = (int64_t)((nlines + lpg - 1) / lpg) * co_tiles;
  static const int waves = lines_env("B2S_LINES_WG_WAVES", 2);
  int64_t splits = ((int64_t)waves * B2S_NUM_SMS) / base;      // whole waves of one CTA per SM
  const int64_t max_splits = ceil_div64(n_out, 8 * LW_ROWS);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rows = ceil_div64(n_out, splits);
  rows = ceil_div64(rows, LW_ROWS) * LW_ROWS;
  splits = ceil_div64(n_out, rows);
  dim3 grid((unsigned)base, (unsigned)splits);
  static const int variant = lines_env("B2S_LINES_WG", 0);
#define LW_LAUNCH(S, D)                                                                                         \
  do {                                                                                                          \
    auto kern = wgrad_lines_kernel<S, D>;                                                                       \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LwSmem<S>::DYN_BYTES));  \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    kern<<<grid, LW_THREADS, LwSmem<S>::DYN_BYTES, st>>>(x4, gy, lines, n_out, n_out_dev, c_in, c_out, nlines, kx, lpg, \
                                                         co_tiles, rows, gw);                                   \
  } while (0)
  if (variant == 1) LW_LAUNCH(4, 2);
  else if (variant == 2) LW_LAUNCH(5, 3);
  else LW_LAUNCH(4, 3);
#undef LW_LAUNCH
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}
