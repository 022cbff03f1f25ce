// wgrad_tc.cu -- tcgen05 (kind::tf32) weight gradient of the sparse convolution.
//
//   gw[k, ci, co] = sum over out rows o of  x[nbr[k, o], ci] * gy[o, co]
//
// As a GEMM: D[M = (kernel offset, ci), N = co] with the reduction (UMMA "K") running over OUT ROWS.  Both operands
// are feature rows, i.e. M/N-contiguous ("MN-major") in shared memory, which for 32-bit elements requires the
// SWIZZLE_128B_BASE32B layout (see tc_ptx.cuh).
//
// One CTA owns a GROUP of KG kernel offsets x one slice of ci (CB blocks of 32 channels) x one tile of BN output
// channels x one range of out rows.  Per pipeline stage of R = 16 out rows it stages
//   B: the gy rows ONCE                                 [BN/32 blocks][16 rows][128 B]
//   A: for every offset of the group the gathered x rows [KG*CB blocks][16 rows][128 B]  (zero-filled where nbr = -1)
// and issues, per M tile of 128 (= 4 consecutive A blocks: 4/CB offsets x CB*32 channels), 2 x tcgen05.mma
// (M=128, N=BN, K=8 rows) into that tile's own TMEM accumulator (columns t*BN ...).  Sharing the gy rows between
// the offsets of a group is what this layout buys: the 64->64 layers read gy 2x instead of 27x and fill all 128
// lanes of the tensor core with two offsets x 64 channels.
// Producers: 8 warps, 8 lanes per (block, row) item = one 128-byte LDGSTS run; the neighbour indices of the NEXT
// stage are prefetched into registers while the current stage's copies are in flight.
// Operands must already be TF32-representable (b2s_round_tf32): tcgen05 kind::tf32 truncates.
// Row ranges are split across CTAs (grid.y); partial tiles are combined with vector fp32 reductions.
//
// Reference call site: autograd of MinkowskiConvolution (R:models/base_model.py:262 -> loss.backward()).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

extern int g_b2s_wg_nbp, g_b2s_wg_lag, g_b2s_wg_occ2, g_b2s_wg_ca, g_b2s_wg_wv;   // lib.cu (b2s_set_tuning)

namespace {

using namespace tc;

constexpr int WG_BM = 128;        // UMMA M
constexpr int WG_ROWS = 32;       // (small-c_in kernel) out rows per pipeline stage
constexpr int WG_A_STAGE = WG_BM * WG_ROWS * 4;   // 16 KB
constexpr int WG_PRODUCERS = 128;
constexpr int WG_THREADS = 160;
constexpr int WG_LAG = 2;
constexpr uint32_t ATOM_BYTES = 1024;             // 8 rows x 128 B consumed per MMA (two 4-row swizzle atoms)
constexpr uint32_t SBO_BYTES = 512;               // next 4-row atom along the reduction
constexpr uint32_t LBO_BYTES = WG_ROWS * 128;     // next 32-channel block

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

// ---------------------------------------------------------------------------------------------
// general kernel: offset groups sharing the gy rows
// ---------------------------------------------------------------------------------------------
constexpr int G_R = 16;                         // out rows per stage (2 MMAs of K = 8 rows per M tile)
constexpr int G_BLOCK = G_R * 128;              // bytes of one 32-channel block of a stage
constexpr int G_PRODUCERS = 256;                // warps 0..7
constexpr int G_THREADS = 288;                  // + warp 8 = MMA issuer
constexpr int G_MAX_STAGES = 8;
constexpr int G_MAX_LAG = 6;                    // producers may run up to this many stages ahead of their hand-over

// cp.async.wait_group with a run-time count (uniform across the CTA)
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    default: cp_async_wait<6>(); break;
  }
}

__device__ __forceinline__ uint32_t g_offset(int row, int c8) {  // chunk c8 (0..7) of a block-local row
  const int unit = (c8 >> 1) ^ (row & 3);
  return (uint32_t)row * 128u + (uint32_t)(((unit << 1) | (c8 & 1)) << 4);
}

struct G2Params {
  int c_in, c_out, k3;
  int KG, CB;          // offsets per group, 32-channel blocks per offset inside this CTA's ci slice
  int MT;              // M tiles of 128 = ceil(KG*CB / 4)
  int ci_tiles, co_tiles, groups;
  int stages;
  int lag;             // stages a producer keeps in flight behind the one it is issuing (< stages)
  int use_atomic;
  int l1;              // 1: gather the x rows through L1 (offsets of a group re-read the same rows within a stage)
  int precise;         // 1: split-bf16 operands (see below), 0: TF32 operands
  int64_t rows_per_split;
};

// MAXR: A items per thread per stage (>= NBP / 2; 16 covers 32 blocks x 16 rows / 32 items per round);
// MINB: CTAs per SM the register budget is sized for (2 needs TCOLS <= 256 and <= 113 KB of shared memory)
template <int BN, int TCOLS, int MAXR, int MINB>
__global__ void __launch_bounds__(G_THREADS, MINB)
    wgrad_group_kernel(const float* __restrict__ x, const float* __restrict__ gy, const int* __restrict__ nbr,
                       int64_t n_out, const int* __restrict__ n_out_dev, G2Params p, float* __restrict__ gw) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int NBP = p.MT * 4;                                   // A blocks per stage incl. zero padding
  const uint32_t a_stage_bytes = (uint32_t)NBP * G_BLOCK, b_stage_bytes = (uint32_t)(BN / 32) * G_BLOCK;
  const uint32_t stage_bytes = a_stage_bytes + b_stage_bytes;
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (G_MAX_STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * G_MAX_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * G_MAX_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (size_t)p.stages * stage_bytes + 8 * (2 * G_MAX_STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int t = blockIdx.x;
  const int cot = t % p.co_tiles;
  t /= p.co_tiles;
  const int cit = t % p.ci_tiles;
  const int grp = t / p.ci_tiles;
  const int k0 = grp * p.KG;
  const int ci0 = cit * p.CB * 32, co0 = cot * BN;
  const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_split;
  const int64_t r_end = min(r_begin + p.rows_per_split, n_out);
  const int T = (int)((r_end - r_begin + G_R - 1) / G_R);
  if (T <= 0) return;  // uniform across the CTA

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), G_PRODUCERS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp < 8) {
    // ===================== producers =====================
    // item = (block b, row r) = one 128-byte run; 8 lanes per item, 32 items per round of the 256 producers.
    // With 16 rows per block, item j*32 + slot is block 2j + h, row r with h = slot >> 4, r = slot & 15: the row and
    // the swizzled destination are per-thread constants and everything else is linear in the round j.
    const int l8 = tid & 7;
    const int slot = tid >> 3;
    const int h = slot >> 4, r = slot & (G_R - 1);
    const int a_rounds = NBP >> 1;                  // NBP blocks, two per round (NBP is a multiple of 4)
    constexpr int B_ROUNDS = BN / 64;
    const int cb_mask = p.CB - 1, cb_shift = 31 - __clz(p.CB);   // CB is a power of two
    // Split-bf16 mode: a gathered 128-byte run is [h of 32 channels | l of 32 channels] (chunks 0-3 / 4-7).  The H and L
    // halves go to separate regions of the stage (first / second half of the A part, likewise of the B part), each laid
    // out as MN-major SWIZZLE_128B atoms of 64 channels x 8 rows: two 32-channel blocks side by side per 128-byte atom
    // row, 16-byte chunks XOR-ed with (row & 7), the next 8 rows 1 KB further, the next 64 channels 2 KB further.
    const int prc = p.precise;
    const uint32_t in_atom = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
                             (uint32_t)((((h << 2) | (l8 & 3)) ^ (r & 7)) << 4);
    const uint32_t dst0 = prc ? (uint32_t)(l8 >> 2) * (a_stage_bytes >> 1) + in_atom
                              : (uint32_t)h * G_BLOCK + g_offset(r, l8);
    const uint32_t dstb0 = prc ? a_stage_bytes + (uint32_t)(l8 >> 2) * (b_stage_bytes >> 1) + in_atom
                               : a_stage_bytes + dst0;
    const uint32_t rstride = prc ? (uint32_t)G_BLOCK : (uint32_t)(2 * G_BLOCK);   // per round of two blocks
    unsigned valid = 0;                              // bit j: round j addresses a real (offset, channel block)
#pragma unroll
    for (int j = 0; j < MAXR; ++j) {
      const int bb = 2 * j + h, kk = bb >> cb_shift;
      if (j < a_rounds && kk < p.KG && k0 + kk < p.k3 && ci0 + (bb & cb_mask) * 32 < p.c_in) valid |= 1u << j;
    }
    const int* nbr_r = nbr ? nbr + r : nullptr;
    const float* x_l = x + ci0 + l8 * 4;
    const float* gy_l = gy + co0 + h * 32 + l8 * 4;
    int idx[MAXR];
    const bool l1 = p.l1 != 0;

    int idx_n[MAXR], idx_nn[MAXR];       // neighbour rows of the next two stages (two-deep index prefetch)
    auto load_idx = [&](int it, int (&dst)[MAXR]) {   // neighbour rows of stage `it` for this thread's A items
      const int64_t r0 = r_begin + (int64_t)it * G_R;
      const bool live = r0 + r < r_end;
#pragma unroll
      for (int j = 0; j < MAXR; ++j) {
        int v = -1;
        if (((valid >> j) & 1u) && live) {
          const int k = k0 + ((2 * j + h) >> cb_shift);
          v = nbr_r ? __ldg(nbr_r + (int64_t)k * pitch + r0) : (int)(r0 + r);
        }
        dst[j] = v;
      }
    };
    auto publish = [&](int it_done) {
      fence_proxy_async();
      mbar_arrive(full_bar(it_done % p.stages));
    };

    auto stage = [&](int it, const int (&cur)[MAXR], int (&fill)[MAXR]) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t a_dst = base + (uint32_t)s * stage_bytes + dst0;
      const uint32_t b_dst = base + (uint32_t)s * stage_bytes + dstb0;
      const int64_t o = r_begin + (int64_t)it * G_R + r;
#pragma unroll
      for (int j = 0; j < MAXR; ++j) {
        if (j < a_rounds) {
          const int i = cur[j];
          const float* src = x_l + (int64_t)(i >= 0 ? i : 0) * p.c_in + ((2 * j + h) & cb_mask) * 32;
          cp_async16_sel(a_dst + (uint32_t)j * rstride, src, i >= 0 ? 16u : 0u, l1);
        }
      }
      {
        const bool live = o < r_end;
        const float* src = gy_l + (live ? o : 0) * p.c_out;
#pragma unroll
        for (int j = 0; j < B_ROUNDS; ++j)
          cp_async16(b_dst + (uint32_t)j * rstride, src + j * 64, live ? 16u : 0u);
      }
      cp_async_commit();
      load_idx(it + 2, fill);             // overlaps with two stages of copies (the table streams from HBM)
      if (it >= p.lag) {
        cp_async_wait_dyn(p.lag);
        publish(it - p.lag);
      }
    };
    // three index sets used round-robin (loop unrolled by three): no register moves out of a set whose loads are in
    // flight -- such a move waits for the loads and cuts the run-ahead to one stage
    load_idx(0, idx);
    load_idx(1, idx_n);
    for (int it = 0; it < T; it += 3) {
      stage(it, idx, idx_nn);
      if (it + 1 < T) stage(it + 1, idx_n, idx);
      if (it + 2 < T) stage(it + 2, idx_nn, idx_n);
    }
    for (int r = T < p.lag ? T : p.lag; r > 0; --r) {   // drain: stage T - r is complete once <= r - 1 groups are pending
      cp_async_wait_dyn(r - 1);
      publish(T - r);
    }

    // ===================== epilogue: TMEM lane = (block 4t + q, channel lane), 32 columns per tcgen05.ld
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                        // TMEM lane quadrant this warp may read
    const uint32_t t_lane = tmem_d + ((uint32_t)(q * 32) << 16);
    for (int mt = (warp >> 2); mt < p.MT; mt += 2) {   // warps 0-3 take even M tiles, warps 4-7 odd ones
      const int b = mt * 4 + q;
      const int kk = b / p.CB, cb = b % p.CB;
      const int k = k0 + kk;
      const int ci = ci0 + cb * 32 + lane;
      const bool ok = kk < p.KG && k < p.k3 && ci < p.c_in;
      float* dst_row = gw + ((int64_t)k * p.c_in + ci) * p.c_out + co0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(mt * BN + c0), v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (p.use_atomic)
              red_add_v4(dst_row + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                         __uint_as_float(v[j + 3]));
            else
              *reinterpret_cast<float4*>(dst_row + c0 + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC = idesc_tf32(WG_BM, BN, 1, 1);  // both operands MN-major
    constexpr uint32_t IDESC16 = idesc_bf16(WG_BM, BN, 1, 1);
    for (int it = 0; it < T; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_stage = base + (uint32_t)s * stage_bytes, b_stage = a_stage + a_stage_bytes;
        for (int mt = 0; mt < p.MT; ++mt) {
          if (p.precise) {
            // one kind::f16 MMA (K = 16 rows) per term: xh*gh + xl*gh + xh*gl into the tile's accumulator
            const uint32_t ah = a_stage + (uint32_t)mt * (2 * G_BLOCK), al = ah + (a_stage_bytes >> 1);
            const uint32_t bh = b_stage, bl = b_stage + (b_stage_bytes >> 1);
            const uint32_t d = tmem_d + (uint32_t)(mt * BN);
            mma_bf16(d, smem_desc_sw128(ah, G_BLOCK, 1024), smem_desc_sw128(bh, G_BLOCK, 1024), IDESC16, it ? 1u : 0u);
            mma_bf16(d, smem_desc_sw128(al, G_BLOCK, 1024), smem_desc_sw128(bh, G_BLOCK, 1024), IDESC16, 1u);
            mma_bf16(d, smem_desc_sw128(ah, G_BLOCK, 1024), smem_desc_sw128(bl, G_BLOCK, 1024), IDESC16, 1u);
            continue;
          }
#pragma unroll
          for (int g = 0; g < G_R / 8; ++g) {
            const uint64_t a_desc =
                smem_desc_sw128_base32(a_stage + (uint32_t)mt * 4 * G_BLOCK + g * ATOM_BYTES, G_BLOCK, SBO_BYTES);
            const uint64_t b_desc = smem_desc_sw128_base32(b_stage + g * ATOM_BYTES, G_BLOCK, SBO_BYTES);
            mma_tf32(tmem_d + (uint32_t)(mt * BN), a_desc, b_desc, IDESC, (it | g) ? 1u : 0u);
          }
        }
        mma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(accum_bar);
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_d);
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
                          int64_t rows_per_split, int l1_on, int precise, float* __restrict__ gw) {
  // Split-bf16 mode: the operand row of a neighbour is 8 bf16 [h0..h3 | l0..l3], so an offset takes 8 columns of N and
  // a CTA owns 32 offsets (N = 256); gy rows are [h | l] per 32 channels and go to separate H / L atoms
  // (MN-major SWIZZLE_128B, see wgrad_group_kernel).  D = (gh + gl)^T [xh | xl]; the epilogue adds the h and l columns.
  const bool l1 = l1_on != 0;
  const int KPG = precise ? 32 : 64;             // kernel offsets per CTA
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
  const int k0 = (blockIdx.x / co_tiles) * KPG;
  const int co0 = cot * WG_BM;
  const int co_valid = min(WG_BM, c_out - co0);
  const bool stack_hl = precise && co_valid <= 64;   // split-bf16: [gh ; gl] stacked along M (see the producers)
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
    auto publish = [&](int it_done) {
      fence_proxy_async();
      mbar_arrive(full_bar(it_done % STAGES));
    };
    // B gather mapping: lane = stage row, this warp owns the 16 kernel offsets k0 + 16 warp + p.  A warp then reads 32
    // consecutive entries of one neighbour-table row (128 B) and -- rows being sorted along x -- mostly consecutive
    // 16-byte feature rows, instead of 32 scattered sectors per instruction with lanes walking the offsets of one row.
    constexpr int NK = 16;
    const int nk = precise ? 8 : NK;               // offsets of this warp
    const int kw = k0 + warp * nk;
    // neighbour rows are fetched two stages ahead of their gather (the table streams from HBM)
    auto load_nbr = [&](int it, int (&da)[NK]) {
      const int64_t o = r_begin + (int64_t)it * WG_ROWS + lane;
      const bool live = o < r_end;
#pragma unroll
      for (int p = 0; p < NK; ++p)
        da[p] = (live && p < nk && kw + p < k3) ? __ldg(&nbr[(int64_t)(kw + p) * pitch + o]) : -1;
    };
    auto stage = [&](int it, const int (&cur)[NK], int (&fill)[NK]) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t a_stage = a_base + s * WG_A_STAGE, b_stage = b_base + s * SM_B_STAGE;
      const int64_t r0 = r_begin + (int64_t)it * WG_ROWS;
      if (lane < a_chunks) {
#pragma unroll
        for (int p = 0; p < WG_ROWS / 4; ++p) {
          const int row = p * 4 + warp;
          const int64_t o = r0 + row;
          const bool live = o < r_end;
          // split-bf16: chunk `lane` of the row = block lane / 8, chunks 0-3 h / 4-7 l; atom = 64 channels x 8 rows
          // (<= 64 output channels: the l parts take the second 64-lane atom of M instead of a second region, so
          // that ONE M = 128 MMA covers gh and gl)
          const uint32_t l_off = stack_hl ? 4096u : 8192u;
          const uint32_t dst =
              precise ? (uint32_t)((lane >> 2) & 1) * l_off + (uint32_t)(lane >> 4) * 4096u + (uint32_t)(row >> 3) * 1024u +
                            (uint32_t)(row & 7) * 128u + (uint32_t)((((((lane >> 3) & 1) << 2) | (lane & 3)) ^ (row & 7)) << 4)
                      : mn_offset(row, lane);
          cp_async16(a_stage + dst, gy + (live ? o : 0) * c_out + co0 + lane * 4, live ? 16u : 0u);
        }
      }
#pragma unroll
      for (int p = 0; p < NK; ++p) {
        if (p >= nk) break;
        const int v = cur[p];
        const uint32_t dst = precise ? (uint32_t)warp * 4096u + (uint32_t)(lane >> 3) * 1024u + (uint32_t)(lane & 7) * 128u +
                                           (uint32_t)((p ^ (lane & 7)) << 4)
                                     : mn_offset(lane, warp * NK + p);
        cp_async16_sel(b_stage + dst, x4 + (v >= 0 ? v : 0), v >= 0 ? 16u : 0u, l1);
      }
      cp_async_commit();
      load_nbr(it + 2, fill);
      if (it >= WG_LAG) {
        cp_async_wait<WG_LAG>();
        publish(it - WG_LAG);
      }
    };
    // three index sets used round-robin (loop unrolled by three): a set is loaded two stages before its gathers and
    // never copied in between -- a register move out of a set whose loads are in flight waits for them, which held
    // this kernel to one stage per memory latency
    int ia[NK], ia_n[NK], ia_nn[NK];
    load_nbr(0, ia);
    load_nbr(1, ia_n);
    for (int it = 0; it < T; it += 3) {
      stage(it, ia, ia_nn);
      if (it + 1 < T) stage(it + 1, ia_n, ia);
      if (it + 2 < T) stage(it + 2, ia_nn, ia_n);
    }
    if (T >= 2) {
      cp_async_wait<1>();
      publish(T - 2);
    }
    cp_async_wait<0>();
    publish(T - 1);

    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int co = stack_hl ? ((warp * 32 + lane) & 63) : warp * 32 + lane;   // stacked: lanes 64.. hold the gl terms
    const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < SM_BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
      if (co < co_valid && precise) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int k = k0 + ((c0 + g8 * 8) >> 3);
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
            if (ci < c_in && k < k3)
              atomicAdd(&gw[((int64_t)k * c_in + ci) * c_out + co0 + co],
                        __uint_as_float(v[g8 * 8 + ci]) + __uint_as_float(v[g8 * 8 + 4 + ci]));
        }
      } else if (co < co_valid) {
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
    constexpr uint32_t IDESC16 = idesc_bf16(WG_BM, SM_BN, 1, 1);
    for (int it = 0; it < T; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const bool issuer = elect_one();
      if (issuer && precise) {
#pragma unroll
        for (int g = 0; g < WG_ROWS / 16; ++g) {      // K = 16 rows per MMA; A terms gh, gl against [xh | xl]
          const uint64_t b_desc = smem_desc_sw128(b_base + s * SM_B_STAGE + g * 2048, 4096, 1024);
          mma_bf16(tmem_d, smem_desc_sw128(a_base + s * WG_A_STAGE + g * 2048, 4096, 1024), b_desc, IDESC16,
                   (it | g) ? 1u : 0u);
          if (!stack_hl)
            mma_bf16(tmem_d, smem_desc_sw128(a_base + s * WG_A_STAGE + 8192 + g * 2048, 4096, 1024), b_desc, IDESC16, 1u);
        }
        mma_commit(empty_bar(s));
      } else if (issuer) {
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
    if (elect_one()) mma_commit(accum_bar);
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<SM_BN>(tmem_d);
  }
}

int wg_ca_knob();

__global__ void __launch_bounds__(256) wg_pad_rows4_kernel(const float* __restrict__ x, int64_t n, int c, int precise,
                                                           float4* __restrict__ x4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (precise) {     // [h0 h1 h2 h3 | l0 l1 l2 l3] bf16
      uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
      for (int j = 0; j < c; ++j) split_bf16(x[i * c + j], h[j], l[j]);
      x4[i] = make_float4(__uint_as_float(h[0] | (h[1] << 16)), __uint_as_float(h[2] | (h[3] << 16)),
                          __uint_as_float(l[0] | (l[1] << 16)), __uint_as_float(l[2] | (l[3] << 16)));
      continue;
    }
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
  const int precise = b2s_precise();
  wg_pad_rows4_kernel<<<grid_for(n_in, 256), 256, 0, st>>>(x, n_in, c_in, precise, x4);
  const int kpg = precise ? 32 : 64;
  const int co_tiles = (c_out + WG_BM - 1) / WG_BM, groups = (k3 + kpg - 1) / kpg;
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
  kern<<<grid, WG_THREADS, L::DYN_BYTES, st>>>(x4, gy, nbr, n_out, n_out_dev, c_in, c_out, k3, co_tiles, rows,
                                               wg_ca_knob(), precise, gw);
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

int wg_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Tuning knobs: nbp_cap = cap on the A blocks a stage holds (smaller stages -> deeper ring), lag = stages a producer
// keeps in flight (0 = derive from the ring depth), occ2 = 1 sizes the kernel for two CTAs per SM where the
// accumulators fit 256 TMEM columns.  Environment (B2S_WG_NBP / _LAG / _OCC2) read once; b2s_set_tuning overrides.
struct WgTuning {
  int nbp_cap, lag, occ2, ca;
};
WgTuning wg_tuning() {
  // defaults measured per operand mode (tools/sweep.py, profiles/r02_sweep_bf16x2.json): with split-bf16 operands a
  // stage is three MMAs per M tile instead of two, and ONE CTA per SM with a hand-over lag of one stage runs the 64->64
  // layers 20 % faster (0.48 -> 0.385 ms) than two CTAs per SM with a lag of two, which is the better choice for TF32
  static const WgTuning env = {wg_env("B2S_WG_NBP", 16), wg_env("B2S_WG_LAG", -1), wg_env("B2S_WG_OCC2", -1),
                               wg_env("B2S_WG_CA", 0)};
  WgTuning t = env;
  if (t.lag < 0) t.lag = b2s_precise() ? 1 : 2;
  if (t.occ2 < 0) t.occ2 = b2s_precise() ? 0 : 1;
  if (g_b2s_wg_nbp >= 0) t.nbp_cap = g_b2s_wg_nbp;
  if (g_b2s_wg_lag >= 0) t.lag = g_b2s_wg_lag;
  if (g_b2s_wg_occ2 >= 0) t.occ2 = g_b2s_wg_occ2;
  if (g_b2s_wg_ca >= 0) t.ca = g_b2s_wg_ca;
  return t;
}

int wg_ca_knob() { return wg_tuning().ca; }

template <int BN, int TCOLS, int MAXR, int MINB>
int launch_group(const float* x, const float* gy, const int* nbr, int64_t n_out, const int* n_out_dev, G2Params p,
                 float* gw, cudaStream_t st) {
  auto kern = wgrad_group_kernel<BN, TCOLS, MAXR, MINB>;
  const int stage_bytes = (p.MT * 4 + BN / 32) * G_BLOCK;
  const int bar_bytes = (2 * G_MAX_STAGES + 2) * 8 + 1024;
  const int budget = MINB == 2 ? 112 * 1024 : 226 * 1024;   // per-CTA shared memory incl. barriers and alignment slack
  int stages = (budget - bar_bytes) / stage_bytes;
  if (stages > G_MAX_STAGES) stages = G_MAX_STAGES;
  if (stages < 2) {
    b2s_set_error("wgrad_tc: stage of %d bytes does not fit twice in shared memory", stage_bytes);
    return -1;
  }
  p.stages = stages;
  p.l1 = wg_tuning().ca;
  p.precise = b2s_precise();
  // default: leave one stage being consumed and one being filled beyond the in-flight ones when the ring allows it
  int lag = wg_tuning().lag > 0 ? wg_tuning().lag : (stages >= 4 ? stages - 2 : stages - 1);
  if (lag > stages - 1) lag = stages - 1;
  if (lag < 1) lag = 1;
  if (lag > G_MAX_LAG) lag = G_MAX_LAG;
  p.lag = lag;
  int dyn = stages * stage_bytes + bar_bytes;
  // one CTA per SM unless sized for two: a CTA may hold all 512 TMEM columns, a second one would wait for them
  if (MINB == 1 && dyn < 116 * 1024) dyn = 116 * 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      b2s_set_error("wgrad_tc: cannot opt in to 227 KB of shared memory");
      return -1;
    }
    attr_set = true;
  }
  const int64_t base = (int64_t)p.groups * p.ci_tiles * p.co_tiles;
  // `wv` half-waves of the resident CTAs (default 4 = two waves); every split gets at least 16 stages of rows.  More
  // splits balance the SMs, fewer splits mean fewer partial tiles to combine with atomics (each split adds one copy
  // of the whole gradient tile to the reduction traffic)
  const int wv = g_b2s_wg_wv > 0 ? g_b2s_wg_wv : wg_env("B2S_WG_WV", 4);
  int64_t splits = ((int64_t)wv * MINB * B2S_NUM_SMS / 2 + base - 1) / base;
  const int64_t max_splits = ceil_div64(n_out, 16 * G_R);
  if (splits > max_splits) splits = max_splits;
  // reduction traffic = splits x (whole gradient tile), gathered traffic ~ n_out x k3 x c_in: where the unsplit grid
  // already has a few dozen CTAs, keep the former below about half of the latter (the 512-channel maps with ~2k rows
  // ran 1.3x faster with 1-2 splits than with 3-6: tools/sweep.py, wg_wv)
  const int64_t traffic_cap = n_out / (2 * (int64_t)p.c_out);
  if (base >= 48 && splits > traffic_cap) splits = traffic_cap;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rows = ceil_div64(n_out, splits);
  rows = ceil_div64(rows, G_R) * G_R;
  splits = ceil_div64(n_out, rows);
  p.rows_per_split = rows;
  p.use_atomic = splits > 1 ? 1 : 0;
  if (p.use_atomic) cudaMemsetAsync(gw, 0, (size_t)p.k3 * p.c_in * p.c_out * sizeof(float), st);
  dim3 grid((unsigned)base, (unsigned)splits);
  kern<<<grid, G_THREADS, dyn, st>>>(x, gy, nbr, n_out, n_out_dev, p, gw);
  return 0;
}

// picks the kernel instance for a given accumulator width: MAXR 8 when a stage holds <= 16 A blocks, two CTAs per SM
// when asked for and the accumulators fit half of TMEM
template <int BN, int TCOLS>
int launch_group_sel(const float* x, const float* gy, const int* nbr, int64_t n_out, const int* n_out_dev,
                     const G2Params& p, float* gw, cudaStream_t st) {
  if (p.MT * 4 <= 16) {
    if constexpr (TCOLS <= 256) {
      if (wg_tuning().occ2) return launch_group<BN, TCOLS, 8, 2>(x, gy, nbr, n_out, n_out_dev, p, gw, st);
    }
    return launch_group<BN, TCOLS, 8, 1>(x, gy, nbr, n_out, n_out_dev, p, gw, st);
  }
  return launch_group<BN, TCOLS, 16, 1>(x, gy, nbr, n_out, n_out_dev, p, gw, st);
}

}  // namespace

bool b2s_wgrad_tc_supported(int32_t c_in, int32_t c_out, int32_t k3, int64_t n_out, bool has_map) {
  (void)k3;
  if (wgrad_tc_disabled() || n_out <= 0) return false;
  if (c_in <= 4) return has_map && c_out % 32 == 0;
  const int blocks = c_in / 32;   // the group kernel slices ci into power-of-two runs of 32-channel blocks (<= 8)
  return c_in % 32 == 0 && c_out % 64 == 0 && (blocks >= 8 || (blocks & (blocks - 1)) == 0);
}

int64_t b2s_wgrad_tc_workspace_bytes(int32_t c_in, int64_t n_in) { return c_in <= 4 ? ((n_in * 16 + 255) & ~(int64_t)255) : 0; }

int b2s_conv_wgrad_tc(const float* x, const float* gy, const int32_t* nbr, int64_t n_in, int64_t n_out,
                      const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, float* gw, void* workspace,
                      cudaStream_t st) {
  if (c_in <= 4) return launch_wgrad_small(x, gy, nbr, n_in, n_out, n_out_dev, c_in, c_out, k3, gw, workspace, st);
  const int bn = c_out % 256 == 0 ? 256 : (c_out % 128 == 0 ? 128 : 64);
  G2Params p{};
  p.c_in = c_in, p.c_out = c_out, p.k3 = k3;
  const int blocks_in = c_in / 32;
  p.CB = blocks_in < 8 ? blocks_in : 8;                  // ci slice of at most 256 channels per CTA
  p.ci_tiles = (blocks_in + p.CB - 1) / p.CB;
  p.co_tiles = c_out / bn;
  int mt_max = 512 / bn;                                 // TMEM columns
  const int mt_cap = wg_tuning().nbp_cap / 4;            // stage-size cap: NBP = 4 MT A blocks of 2 KB
  if (mt_cap >= 1 && mt_max > mt_cap) mt_max = mt_cap;
  int kg = (mt_max * 4) / p.CB;
  if (kg < 1) kg = 1;
  if (kg > k3) kg = k3;
  p.groups = (k3 + kg - 1) / kg;
  kg = (k3 + p.groups - 1) / p.groups;                   // even out the groups (27 -> 14 + 13, 7+7+7+6 ...)
  p.KG = kg;
  p.MT = (kg * p.CB + 3) / 4;
  const int cols = p.MT * bn;
  if (bn == 256) return cols <= 256 ? launch_group_sel<256, 256>(x, gy, nbr, n_out, n_out_dev, p, gw, st)
                                    : launch_group_sel<256, 512>(x, gy, nbr, n_out, n_out_dev, p, gw, st);
  if (bn == 128) return cols <= 128 ? launch_group_sel<128, 128>(x, gy, nbr, n_out, n_out_dev, p, gw, st)
                       : cols <= 256 ? launch_group_sel<128, 256>(x, gy, nbr, n_out, n_out_dev, p, gw, st)
                                     : launch_group_sel<128, 512>(x, gy, nbr, n_out, n_out_dev, p, gw, st);
  return cols <= 64 ? launch_group_sel<64, 64>(x, gy, nbr, n_out, n_out_dev, p, gw, st)
       : cols <= 128 ? launch_group_sel<64, 128>(x, gy, nbr, n_out, n_out_dev, p, gw, st)
       : cols <= 256 ? launch_group_sel<64, 256>(x, gy, nbr, n_out, n_out_dev, p, gw, st)
                     : launch_group_sel<64, 512>(x, gy, nbr, n_out, n_out_dev, p, gw, st);
}
