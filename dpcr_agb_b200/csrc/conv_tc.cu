// conv_tc.cu -- tcgen05 (kind::tf32, fp32 accumulate in TMEM) kernels for the sparse convolution.
//
// Forward / dgrad: OUTPUT-STATIONARY implicit GEMM over the neighbour table.
//   CTA tile = 128 out rows x BN out channels, accumulator = 128 TMEM lanes x BN fp32 columns.
//   K loop   = (kernel offset k) x (32-channel chunk of c_in): per step
//       A [128 x 32] = rows x[nbr[k, o], chunk]   gathered by 128 producer threads with 16-byte LDGSTS
//                      (zero-fill when nbr = -1) straight into the 128B-swizzled K-major UMMA layout,
//       B [BN  x 32] = a pre-swizzled image of W[k] (built once per call by prep_weights_kernel), fetched
//                      with ONE 1-D bulk copy (UBLKCP) whose bytes complete on the stage's mbarrier,
//       4 x tcgen05.mma (M=128, N=BN, K=8) issued by one elected thread; tcgen05.commit frees the stage.
//   Epilogue: tcgen05.ld 32x32b -> + bias -> fp32 rows of y.  No atomics, deterministic.
//   SMALL mode (c_in <= 4, the k7 stem with c_in = 3): x is padded to 4 floats per row so that one
//   16-byte chunk is one (row, offset) gather; a K step covers 8 kernel offsets.
//
// Reference call sites: R:modules/MinkowskiEngine/SENet.py:49-52,94-97; resnet_block.py:48-54,95-107.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>

extern int g_b2s_tc_rot, g_b2s_tc_ca, g_b2s_tc_occ1, g_b2s_tc_m256, g_b2s_tc_ta;   // lib.cu (b2s_set_tuning)

namespace {

using namespace tc;

constexpr int BM = 128;            // out rows per CTA == UMMA M
constexpr int BK = 32;             // fp32 per K step == one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int PRODUCERS = 128;     // warps 0..3 gather A and run the epilogue; warp 4 issues MMA
constexpr int TC_THREADS = 160;
constexpr int SMALL_NB = 3;        // weight images per stage of the small-c_in kernel in split-bf16 mode

// ---------------------------------------------------------------------------------------------
// weight image: img[it][n][32] (128 bytes per n, 16-byte chunks XOR-swizzled by n & 7)
//   general: it = k * (c_in/32) + cc, element kk = ci - 32*cc
//   small  : it = k / 8, element kk = (k % 8) * 4 + ci      (ci < c_in <= 4, rest zero)
// B_k[ci, co] = w[(k*c_in + ci)*c_out + co] (layout bit0 = 0) or w[(k*c_out + co)*c_in + ci] (bit0 = 1);
// layout bit1 reverses the kernel index (k -> k3-1-k)
// ---------------------------------------------------------------------------------------------
// Split-bf16 mode (`precise`, see tc_ptx.cuh): a 128-byte image row holds 64 bf16.
//   general: [H(32 channels) | L(32 channels)] of the same (offset, channel chunk) -- same bytes as the TF32 image;
//   small  : per kernel offset 8 bf16 matching the operand row [h0..h3 | l0..l3], in SMALL_NB = 3 images
//            (it*3 + j): j = 0 [H | H], j = 1 [M | M], j = 2 [L | 0] with W = H + M + L (24 bits: the k7 stem's
//            weights need more than the 16-17 bits of a pair, and the l*M cross term, measured end to end: without
//            them the stem kernel's own gradient is off by 4e-3 at BASELINE plot size).
// Image rows for the kernel that keeps the A operand in tensor memory (gather_gemm_ta_kernel, precise == 2): position
// p of the 64 bf16 of a row ([H | L], 16 per MMA k-step) holds channel 8 ((p % 16) / 4) + 4 ((p / 16) % 2) + p % 4 of the
// 32-channel chunk -- the order in which that kernel's row fragments land in tensor memory.
__device__ __forceinline__ int ta_channel(int p) { return 8 * ((p & 15) >> 2) + 4 * ((p >> 4) & 1) + (p & 3); }

__device__ __forceinline__ float weight_at(const float* __restrict__ w, int c_in, int c_out, int k3, int w_layout, int k,
                                           int ci, int n) {
  if (k >= k3 || ci >= c_in) return 0.f;
  const int kw = (w_layout & 2) ? k3 - 1 - k : k;  // bit 1: kernel index reversed (symmetric maps)
  return !(w_layout & 1) ? w[((int64_t)kw * c_in + ci) * c_out + n] : w[((int64_t)kw * c_out + n) * c_in + ci];
}

__global__ void __launch_bounds__(256) prep_weights_kernel(const float* __restrict__ w, int c_in, int c_out, int k3,
                                                           int w_layout, int small, int T, int precise,
                                                           float* __restrict__ img) {
  const int nb = (precise && small) ? SMALL_NB : 1;
  const int64_t total = (int64_t)T * nb * c_out * BK;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int pos = (int)(e % BK);             // physical position inside the 128-byte row
    const int64_t rn = e / BK;
    const int n = (int)(rn % c_out);
    const int itj = (int)(rn / c_out);
    const int chunk = (pos >> 2) ^ (n & 7);     // logical 16-byte chunk stored at this physical slot
    if (precise) {
      const int it = itj / nb, j = itj % nb;
      const int k16 = chunk * 8 + (pos & 3) * 2;  // the two bf16 of this 4-byte slot: k16, k16 + 1
      uint32_t out[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int kq = k16 + q;
        uint32_t h, m, l;
        if (small) {
          const int s8 = kq & 7;
          split_bf16x3(weight_at(w, c_in, c_out, k3, w_layout, it * 8 + (kq >> 3), s8 & 3, n), h, m, l);
          out[q] = j == 0 ? h : (j == 1 ? m : (s8 < 4 ? l : 0u));
        } else {
          const int kc = c_in / BK;
          const int ch = precise == 2 ? ta_channel(kq) : (kq & 31);
          split_bf16(weight_at(w, c_in, c_out, k3, w_layout, it / kc, (it % kc) * BK + ch, n), h, l);
          out[q] = kq < 32 ? h : l;
        }
      }
      img[e] = __uint_as_float(out[0] | (out[1] << 16));
      continue;
    }
    const int it = itj;
    const int kk = chunk * 4 + (pos & 3);
    int k, ci;
    if (small) {
      k = it * 8 + (kk >> 2);
      ci = kk & 3;
    } else {
      const int kc = c_in / BK;
      k = it / kc;
      ci = (it % kc) * BK + kk;
    }
    float v = 0.f;
    if (k < k3 && ci < c_in) {
      const int kw = (w_layout & 2) ? k3 - 1 - k : k;  // bit 1: kernel index reversed (symmetric maps)
      v = !(w_layout & 1) ? w[((int64_t)kw * c_in + ci) * c_out + n] : w[((int64_t)kw * c_out + n) * c_in + ci];
    }
    // round-to-nearest to TF32 here: whatever the tensor core does with the low 13 mantissa bits of B is then a no-op
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    img[e] = __uint_as_float(r);
  }
}

// Same image for the forward layout (w[k][ci][co], co contiguous) through a shared-memory transpose: the kernel above
// reads that layout with ci fastest across threads -- one 32-byte sector per element.  Block = (iteration, 32 output
// channels): rows of 32 co are read coalesced, 128-byte image rows are written coalesced.
__global__ void __launch_bounds__(256) prep_weights_t_kernel(const float* __restrict__ w, int c_in, int c_out, int k3,
                                                             int w_layout, int precise, float* __restrict__ img) {
  __shared__ float t[32][33];
  const int it = blockIdx.x, n0 = blockIdx.y * 32, kc = c_in / BK;
  const int k = it / kc, ci0 = (it % kc) * BK;
  const int kw = (w_layout & 2) ? k3 - 1 - k : k;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) t[r][tx] = w[((int64_t)kw * c_in + ci0 + r) * c_out + n0 + tx];
  __syncthreads();
#pragma unroll
  for (int y = ty; y < 32; y += 8) {
    const int n = n0 + y;
    const int chunk = (tx >> 2) ^ (n & 7);
    uint32_t r;
    if (precise) {
      const int k16 = chunk * 8 + (tx & 3) * 2;   // [H(32) | L(32)]: the two bf16 of this slot
      uint32_t h0, l0, h1, l1;
      const int ch = precise == 2 ? ta_channel(k16) : (k16 & 31);   // k16 is even: its partner is channel ch + 1
      split_bf16(t[ch][y], h0, l0);
      split_bf16(t[ch + 1][y], h1, l1);
      r = k16 < 32 ? (h0 | (h1 << 16)) : (l0 | (l1 << 16));
    } else {
      const int kk = chunk * 4 + (tx & 3);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(t[kk][y]));
    }
    img[((int64_t)it * c_out + n) * BK + tx] = __uint_as_float(r);
  }
}

__global__ void __launch_bounds__(256) pad_rows4_kernel(const float* __restrict__ x, int64_t n, int c, int precise,
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

template <int BN, int STAGES, int NB = 1>
struct SmemLayout {
  static constexpr int B_IMG_BYTES = BN * 128;
  static constexpr int B_STAGE_BYTES = NB * BN * 128;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_STAGE_BYTES;
  static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE_BYTES;  // full[S], empty[S], accum, tmem slot
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 2) * 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;                   // slack for manual 1024-byte alignment
};

// PERM mode (dgrad of stride-2 convolutions).  In the transposed map of a stride-2 conv a fine row only has partners
// at the kernel offsets whose parity matches the row's position inside its 2x2x2 cell: 27/8 = 3.4 offsets on average
// instead of 27, so the dense loop over all offsets gathers ~89 % zero rows.  b2s_parity_plan sorts the fine rows by
// parity class (tile-aligned); a tile then walks only the offsets of its class (klist) and writes its rows through
// the permutation.
struct PermArgs {
  const int* perm;      // [tiles * 128] output row of every tile row, -1 = padding
  const int* bounds;    // [9] first tile of every parity class; bounds[8] = number of tiles
  unsigned char nk[8];  // offsets per class
  unsigned char klist[8][27];
};

// LAG: a producer hands over stage (it - LAG) after issuing the copies of stage it (LAG + 1 stages of copies in flight)
// flags: bit 0 = offset rotation, bit 1 = L1-allocating gathers, bit 2 = split-bf16 operands (NB > 1 implies it)
// PW: producer warps (4, or 8 in SMALL mode: threads 128.. take the second half of a stage's kernel offsets -- the
// split-bf16 stem keeps three weight images per stage and therefore one CTA per SM; eight gathering warps restore the
// number of copies in flight that two 4-warp CTAs had).  Warp PW issues the MMAs; warps 0-3 run the epilogue.
template <int BN, int STAGES, bool SMALL, int LAG, bool PERM = false, int NB = 1, int PW = 4>
__global__ void __launch_bounds__((PW + 1) * 32, 1)
    gather_gemm_tc_kernel(const float* __restrict__ x, const float* __restrict__ wimg, const float* __restrict__ bias,
                          const int* __restrict__ nbr, int64_t n_out, const int* __restrict__ n_out_dev, int c_in,
                          int c_out, int k3, int T_total, int it_per_split, float* __restrict__ y,
                          const PermArgs pa, int flags, float* __restrict__ col_stats) {
  const int rot_on = flags & 1;
  const bool l1 = (flags & 2) != 0;                // gather through L1 (cp.async.ca) instead of L2 only (.cg)
  const bool precise = NB > 1 || (flags & 4) != 0;
  const int64_t pitch = n_out;                     // row pitch of the neighbour table (the caller's capacity)
  n_out = b2s_rows(n_out, n_out_dev);
  int cls = 0;
  if (PERM) {
    const int tile = blockIdx.x;
    if (tile >= __ldg(&pa.bounds[8])) return;      // beyond the last class tile (uniform across the CTA)
    while (cls < 7 && tile >= __ldg(&pa.bounds[cls + 1])) ++cls;
    T_total = (int)pa.nk[cls] * (c_in / BK);
    it_per_split = T_total;
  } else if ((int64_t)blockIdx.x * BM >= n_out) {
    return;                                        // whole tile beyond the live rows (uniform across the CTA)
  }
  using L = SmemLayout<BN, STAGES, NB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base + L::A_OFF, b_base = base + L::B_OFF;
  const uint32_t bar_base = base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // split-K: blockIdx.z owns iterations [it0, it0 + T) of the (offset x channel-chunk) loop; partial tiles are
  // combined with fp32 vector reductions into a zero-initialised y (small maps: too few row tiles to fill 148 SMs)
  const int it0 = blockIdx.z * it_per_split;
  const int T = min(it_per_split, T_total - it0);
  if (T <= 0) {
    if (PERM) {   // a class without any offset (K = 1: rows off the coarse lattice): its rows are zero
      for (int e = tid; e < BM * (BN / 4); e += (PW + 1) * 32) {
        const int o = __ldg(&pa.perm[m0 + e / (BN / 4)]);
        if (o >= 0) reinterpret_cast<float4*>(y + (int64_t)o * c_out + n0)[e % (BN / 4)] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }
  const bool split = gridDim.z > 1;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), PW * 32 + 1);    // gather arrivals + 1 arrive.expect_tx for the B bulk copy
      mbar_init(empty_bar(s), 1);             // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  static_assert(PW == 4 || (PW == 8 && SMALL), "eight producer warps: SMALL mode only");
  // WIDE: the NB weight images of a stage are consecutive rows of shared memory, so ONE MMA of N = NB * BN reads the
  // A stage once for all of them (N = 64 MMAs re-read the 16 KB A stage per image: the split-bf16 stem was bound by
  // shared-memory operand reads); the accumulator holds NB column groups that the epilogue adds.
  constexpr bool WIDE = NB > 1 && NB * BN <= 256;
  constexpr int TCOLS = WIDE ? 256 : BN;
  if (warp == PW) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp < PW) {
    // ===================== producers: gather A, fetch B =====================
    constexpr int NPT = SMALL ? 32 / PW : 8;        // SMALL: kernel offsets of a stage handled by this thread
    const int srow = tid & 127;                      // SMALL: tile row of this thread
    const int pbase = SMALL ? (tid >> 7) * NPT : 0;  // SMALL: first of its offsets
    const int chunk = tid & 7;   // 16-byte chunk inside the 128-byte row
    const int rsub = tid >> 3;   // rows rsub + 16 p
    const int kc = SMALL ? 1 : c_in / BK;
    // Neighbour indices are fetched TWO index groups ahead of their use (a group = one kernel offset; in SMALL mode
    // one stage = 8 offsets), so the gather never waits on the dependent nbr -> row-address load chain.
    int idx[8], idx1[8], idx2[8];
    const int* nbr_t = nbr ? nbr + m0 + rsub : nullptr;          // this thread's rows: + 16 p
    int orow[8];                                                  // PERM: output (= table query) row of tile row p
    if (PERM) {
#pragma unroll
      for (int p = 0; p < 8; ++p) orow[p] = __ldg(&pa.perm[m0 + rsub + 16 * p]);
    }
    // ROT (tuning knob "tc_rot"): every row tile starts its walk over the kernel offsets at a different offset, so
    // that the CTAs running side by side do not all fetch the same weight tile from the same L2 slices at once
    const int rot = (!PERM && !SMALL && rot_on && gridDim.z == 1) ? (int)((blockIdx.x * 11u) % (unsigned)k3) : 0;
    auto kof = [&](int g) -> int {                                // PERM: g-th offset of this tile's parity class
      if (PERM) return g < (int)pa.nk[cls] ? (int)pa.klist[cls][g] : k3;
      const int k = g + rot;
      return (k >= k3 && g < k3) ? k - k3 : k;
    };
    auto load_group = [&](int g, int (&dst)[8]) {                 // g: offset index (general) / stage index (SMALL)
      if (SMALL) {
        // thread = tile row, p = the 8 kernel offsets of the stage: a warp reads 32 consecutive entries of one table
        // row (128 B) and, rows being sorted along x, gathers mostly consecutive feature rows -- instead of 32
        // scattered 16-byte sectors per instruction when the lanes of a warp walk the offsets of one row
        const int64_t o = m0 + srow;
#pragma unroll
        for (int p = 0; p < NPT; ++p) {
          const int k = g * 8 + pbase + p;
          dst[p] = (k < k3 && o < n_out) ? (nbr ? __ldg(nbr + (int64_t)k * pitch + o) : (int)o) : -1;
        }
        return;
      }
      const int k = kof(g);
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int v = -1;
        if (PERM) {
          if (k < k3 && orow[p] >= 0) v = __ldg(nbr + (int64_t)k * pitch + orow[p]);
        } else {
          const int64_t o = m0 + rsub + 16 * p;
          if (k < k3 && o < n_out) v = nbr_t ? __ldg(nbr_t + (int64_t)k * pitch + 16 * p) : (int)o;
        }
        dst[p] = v;
      }
    };
    const int g0 = SMALL ? it0 : it0 / kc;                        // it0 is a multiple of kc

    auto publish = [&](int s_done) {  // all of this thread's LDGSTS for that stage have landed
      fence_proxy_async();
      mbar_arrive(full_bar(s_done));
    };

    int s = 0, cc = 0, g = g0;
    uint32_t ph = 0;
    int s_pub = 0;                    // stage to hand over next (LAG iterations behind)
    auto issue_stage = [&](int it, const int (&cur)[8]) {   // wait for the slot, start the B copy and the A gathers
      mbar_wait(empty_bar(s), ph ^ 1u);
      // index of this iteration's weight tile in the image: (offset, channel chunk), offset-major
      const int git = (PERM || !SMALL) ? kof(g) * kc + cc : it0 + it;
      if (tid == 0) {
        mbar_arrive_expect_tx(full_bar(s), L::B_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < NB; ++j)
          bulk_g2s(b_base + s * L::B_STAGE_BYTES + j * L::B_IMG_BYTES,
                   wimg + (((int64_t)git * NB + j) * c_out + n0) * BK, L::B_IMG_BYTES, full_bar(s));
      }
      const uint32_t a_stage = a_base + s * A_STAGE_BYTES;
#pragma unroll
      for (int p = 0; p < NPT; ++p) {
        const int row = SMALL ? srow : rsub + 16 * p;
        const int i = cur[p];
        const float* src = SMALL ? x + (int64_t)(i >= 0 ? i : 0) * 4
                                 : x + (int64_t)(i >= 0 ? i : 0) * c_in + cc * BK + chunk * 4;
        cp_async16_sel(a_stage + sw128_offset(row, SMALL ? pbase + p : chunk), src, i >= 0 ? 16u : 0u, l1);
      }
      cp_async_commit();
    };
    auto finish_stage = [&](int it) {                        // advance the ring, hand over stage it - LAG
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
      if (it >= LAG) {
        cp_async_wait<LAG>();
        publish(s_pub);
        if (++s_pub == STAGES) s_pub = 0;
      }
    };

    if (SMALL) {
      // one index group per stage: three register sets used round-robin (loop unrolled by three), so that a set is
      // loaded two stages before its gathers WITHOUT being copied in between -- a register move out of a set whose
      // loads are in flight waits for them, which held the k7 stem to one stage per memory latency
      int r0[8], r1[8], r2[8];
      load_group(g0, r0);
      load_group(g0 + 1, r1);
      for (int it = 0; it < T; it += 3) {
        issue_stage(it, r0);
        load_group(g0 + it + 2, r2);
        finish_stage(it);
        if (it + 1 < T) {
          issue_stage(it + 1, r1);
          load_group(g0 + it + 3, r0);
          finish_stage(it + 1);
        }
        if (it + 2 < T) {
          issue_stage(it + 2, r2);
          load_group(g0 + it + 4, r1);
          finish_stage(it + 2);
        }
      }
    } else {
      // same move-free scheme per index group (= kernel offset, kc stages each): the set of group g is refilled with
      // group g + 3 as soon as the group's last stage has been issued
      load_group(g0, idx);
      load_group(g0 + 1, idx1);
      load_group(g0 + 2, idx2);
      int it = 0;
      auto run_group = [&](const int (&cur)[8]) {
        for (cc = 0; cc < kc && it < T; ++cc, ++it) {
          issue_stage(it, cur);
          finish_stage(it);
        }
        ++g;
      };
      while (it < T) {
        run_group(idx);
        load_group(g + 2, idx);
        if (it < T) {
          run_group(idx1);
          load_group(g + 2, idx1);
        }
        if (it < T) {
          run_group(idx2);
          load_group(g + 2, idx2);
        }
      }
    }
    // drain the last min(LAG, T) stages
    cp_async_wait<0>();
    for (int r = T < LAG ? T : LAG; r > 0; --r) {
      publish(s_pub);
      if (++s_pub == STAGES) s_pub = 0;
    }

    // ===================== epilogue: TMEM -> registers -> y =====================
    if (warp < 4) {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int64_t o = PERM ? (int64_t)__ldg(&pa.perm[m0 + warp * 32 + lane]) : m0 + warp * 32 + lane;
    const bool o_ok = PERM ? o >= 0 : o < n_out;
    const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
      if (WIDE) {            // add the column groups of the other weight images
#pragma unroll
        for (int j = 1; j < NB; ++j) {
          uint32_t u[32];
          tmem_ld32(t_lane + (uint32_t)(j * BN + c0), u);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(u[q]));
        }
      }
      if (col_stats) {     // batch-norm statistics of the output (host: only when the tile is final, i.e. no split-K)
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = o_ok ? __uint_as_float(v[j]) + (bias ? __ldg(&bias[n0 + c0 + j]) : 0.f) : 0.f;
        epilogue_col_stats(r, lane, reinterpret_cast<float2*>(smem + L::A_OFF) + warp * BN + c0);
      }
      if (o_ok) {
        float* dst = y + o * c_out + n0 + c0;
        const bool add_bias = bias && it0 == 0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 r;
          r.x = __uint_as_float(v[j]) + (add_bias ? __ldg(&bias[n0 + c0 + j]) : 0.f);
          r.y = __uint_as_float(v[j + 1]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 1]) : 0.f);
          r.z = __uint_as_float(v[j + 2]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 2]) : 0.f);
          r.w = __uint_as_float(v[j + 3]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 3]) : 0.f);
          if (split)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(r.x), "f"(r.y), "f"(r.z),
                         "f"(r.w)
                         : "memory");
          else
            *reinterpret_cast<float4*>(dst + j) = r;
        }
      }
    }
    if (col_stats) {       // the four epilogue warps combine their per-column sums into this row tile's partial row
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float2* red = reinterpret_cast<const float2*>(smem + L::A_OFF);
      float* part = col_stats + (int64_t)blockIdx.x * 2 * c_out;
      if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)    // header behind the partial rows: rows per partial row
        col_stats[((pitch + 127) / 128) * 2 * c_out] = (float)BM;
      for (int c = tid; c < BN; c += 128) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int wq = 0; wq < 4; ++wq) {
          s1 += red[wq * BN + c].x;
          s2 += red[wq * BN + c].y;
        }
        part[n0 + c] = s1;
        part[c_out + n0 + c] = s2;
      }
    }
    tc_fence_before();
    }
  } else {
    // ===================== MMA issuer: warp PW stays converged, lane 0 issues =====================
    constexpr uint32_t IDESC = idesc_tf32(BM, BN, 0, 0), IDESC16 = idesc_bf16(BM, BN, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_desc = smem_desc_sw128(a_base + s * A_STAGE_BYTES, 16, 1024);
        const uint64_t b_desc = smem_desc_sw128(b_base + s * L::B_STAGE_BYTES, 16, 1024);
        if (!precise) {
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk)  // advance 32 bytes (8 tf32) inside the swizzle row per MMA
            mma_tf32(tmem_d, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), IDESC, (it | kk) ? 1u : 0u);
        } else if (NB == 1) {
          // rows are [h | l] x [H | L], 32 bytes (16 bf16) per MMA: h*H (steps 0,1 x 0,1), l*H (2,3 x 0,1), h*L (0,1 x 2,3)
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const int ak = q < 4 ? q : q - 4, bk = q < 2 ? q : q - 2;
            mma_bf16(tmem_d, a_desc + (uint64_t)(ak * 2), b_desc + (uint64_t)(bk * 2), IDESC16, (it | q) ? 1u : 0u);
          }
        } else if (WIDE) {
          constexpr uint32_t IDESCW = idesc_bf16(BM, WIDE ? NB * BN : BN, 0, 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_bf16(tmem_d, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), IDESCW, (it | kk) ? 1u : 0u);
        } else {
          // small c_in: operand rows [h4 | l4] per offset against the images [H | H], [M | M], [L | 0]
#pragma unroll
          for (int j = 0; j < NB; ++j)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              mma_bf16(tmem_d, a_desc + (uint64_t)(kk * 2),
                       b_desc + (uint64_t)((j * L::B_IMG_BYTES) >> 4) + (uint64_t)(kk * 2), IDESC16,
                       (it | j | kk) ? 1u : 0u);
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
  if (warp == PW) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_d);
  }
}


// ---------------------------------------------------------------------------------------------
// M = 256 variant of the general kernel (c_in % 32 == 0, no split-K, no permutation): TWO 128-row accumulator tiles
// per CTA share every weight stage, so the weight image -- half of the L2->SM traffic of the 64- and 128-channel
// layers, whose conv kernels run at 60-80 % of the L2 fabric ceiling -- is fetched once per 256 output rows instead
// of once per 128.  8 producer warps (tile half = warp / 4), warp 8 issues 2 x 4 MMAs per stage; one CTA per SM with
// the same bytes in flight as two CTAs of the M = 128 kernel.
// ---------------------------------------------------------------------------------------------
constexpr int TC2_THREADS = 288;
template <int BN, int STAGES>
struct SmemLayout2 {
  static constexpr int A_STAGE = 2 * A_STAGE_BYTES;
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_STAGE;
  static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 2) * 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <int BN, int STAGES, int LAG>
__global__ void __launch_bounds__(TC2_THREADS, 1)
    gather_gemm_tc2_kernel(const float* __restrict__ x, const float* __restrict__ wimg, const float* __restrict__ bias,
                           const int* __restrict__ nbr, int64_t n_out, const int* __restrict__ n_out_dev, int c_in,
                           int c_out, int k3, int T_total, int it_per_split, float* __restrict__ y, int precise,
                           float* __restrict__ col_stats) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  if ((int64_t)blockIdx.x * (2 * BM) >= n_out) return;
  // split-K as in the M = 128 kernel: blockIdx.z owns iterations [it0, it0 + T), partial tiles are combined with
  // fp32 vector reductions into a zero-initialised y
  const int it0 = blockIdx.z * it_per_split;       // a multiple of c_in / 32
  const int T = min(it_per_split, T_total - it0);
  if (T <= 0) return;
  const bool split = gridDim.z > 1;
  using L = SmemLayout2<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base + L::A_OFF, b_base = base + L::B_OFF;
  const uint32_t bar_base = base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * (2 * BM);
  const int n0 = blockIdx.y * BN;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2 * PRODUCERS + 1);   // 256 gather arrivals + 1 arrive.expect_tx for the B bulk copy
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp < 8) {
    const int half = tid >> 7, t = tid & 127;
    const int chunk = t & 7, rsub = t >> 3;
    const int kc = c_in / BK;
    const int64_t mh = m0 + half * BM;
    const int* nbr_t = nbr ? nbr + mh + rsub : nullptr;
    int idx[8], idx1[8], idx2[8];
    auto load_group = [&](int g, int (&dst)[8]) {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int64_t o = mh + rsub + 16 * p;
        int v = -1;
        if (g < k3 && o < n_out) v = nbr_t ? __ldg(nbr_t + (int64_t)g * pitch + 16 * p) : (int)o;
        dst[p] = v;
      }
    };
    auto publish = [&](int s_done) {
      fence_proxy_async();
      mbar_arrive(full_bar(s_done));
    };
    const int g0 = it0 / kc;
    int s = 0, cc = 0, g = g0, s_pub = 0;
    uint32_t ph = 0;
    auto issue_stage = [&](const int (&cur)[8]) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (tid == 0) {
        mbar_arrive_expect_tx(full_bar(s), L::B_STAGE_BYTES);
        bulk_g2s(b_base + s * L::B_STAGE_BYTES, wimg + ((int64_t)(g * kc + cc) * c_out + n0) * BK, L::B_STAGE_BYTES,
                 full_bar(s));
      }
      const uint32_t a_stage = a_base + s * L::A_STAGE + half * A_STAGE_BYTES;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int i = cur[p];
        cp_async16(a_stage + sw128_offset(rsub + 16 * p, chunk),
                   x + (int64_t)(i >= 0 ? i : 0) * c_in + cc * BK + chunk * 4, i >= 0 ? 16u : 0u);
      }
      cp_async_commit();
    };
    auto finish_stage = [&](int it) {
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
      if (it >= LAG) {
        cp_async_wait<LAG>();
        publish(s_pub);
        if (++s_pub == STAGES) s_pub = 0;
      }
    };
    load_group(g0, idx);
    load_group(g0 + 1, idx1);
    load_group(g0 + 2, idx2);
    int it = 0;
    auto run_group = [&](const int (&cur)[8]) {
      for (cc = 0; cc < kc && it < T; ++cc, ++it) {
        issue_stage(cur);
        finish_stage(it);
      }
      ++g;
    };
    while (it < T) {
      run_group(idx);
      load_group(g + 2, idx);
      if (it < T) {
        run_group(idx1);
        load_group(g + 2, idx1);
      }
      if (it < T) {
        run_group(idx2);
        load_group(g + 2, idx2);
      }
    }
    cp_async_wait<0>();
    for (int r = T < LAG ? T : LAG; r > 0; --r) {
      publish(s_pub);
      if (++s_pub == STAGES) s_pub = 0;
    }

    // epilogue: warps 0-3 read tile 0, warps 4-7 tile 1; a warp may touch TMEM lanes 32 (warp % 4) ...
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int64_t o = mh + q * 32 + lane;
    const uint32_t t_lane = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * BN);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
      if (col_stats) {
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = o < n_out ? __uint_as_float(v[j]) + (bias ? __ldg(&bias[n0 + c0 + j]) : 0.f) : 0.f;
        epilogue_col_stats(r, lane, reinterpret_cast<float2*>(smem + L::A_OFF) + warp * BN + c0);
      }
      if (o < n_out) {
        float* dst = y + o * c_out + n0 + c0;
        const bool add_bias = bias && it0 == 0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 r;
          r.x = __uint_as_float(v[j]) + (add_bias ? __ldg(&bias[n0 + c0 + j]) : 0.f);
          r.y = __uint_as_float(v[j + 1]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 1]) : 0.f);
          r.z = __uint_as_float(v[j + 2]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 2]) : 0.f);
          r.w = __uint_as_float(v[j + 3]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 3]) : 0.f);
          if (split)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(r.x), "f"(r.y), "f"(r.z),
                         "f"(r.w)
                         : "memory");
          else
            *reinterpret_cast<float4*>(dst + j) = r;
        }
      }
    }
    if (col_stats) {       // all eight epilogue warps (two row tiles) combine their per-column sums
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2* red = reinterpret_cast<const float2*>(smem + L::A_OFF);
      float* part = col_stats + (int64_t)blockIdx.x * 2 * c_out;
      if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) col_stats[((pitch + 127) / 128) * 2 * c_out] = (float)(2 * BM);
      for (int c = tid; c < BN; c += 256) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) {
          s1 += red[wq * BN + c].x;
          s2 += red[wq * BN + c].y;
        }
        part[n0 + c] = s1;
        part[c_out + n0 + c] = s2;
      }
    }
    tc_fence_before();
  } else {
    constexpr uint32_t IDESC = idesc_tf32(BM, BN, 0, 0), IDESC16 = idesc_bf16(BM, BN, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t b_desc = smem_desc_sw128(b_base + s * L::B_STAGE_BYTES, 16, 1024);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t a_desc = smem_desc_sw128(a_base + s * L::A_STAGE + mt * A_STAGE_BYTES, 16, 1024);
          if (!precise) {
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk)
              mma_tf32(tmem_d + (uint32_t)(mt * BN), a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), IDESC,
                       (it | kk) ? 1u : 0u);
          } else {   // split-bf16: [h | l] x [H | L] -> h*H, l*H, h*L (see gather_gemm_tc_kernel)
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              const int ak = q < 4 ? q : q - 4, bk = q < 2 ? q : q - 2;
              mma_bf16(tmem_d + (uint32_t)(mt * BN), a_desc + (uint64_t)(ak * 2), b_desc + (uint64_t)(bk * 2), IDESC16,
                       (it | q) ? 1u : 0u);
            }
          }
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
    tmem_dealloc<2 * BN>(tmem_d);
  }
}


// ---------------------------------------------------------------------------------------------
// TMA variant of the general (c_in % 32 == 0) kernel: the A rows are gathered by the TMA unit
// (cp.async.bulk.tensor ... tile::gather4, 128B swizzle) instead of 16-byte LDGSTS -- the LSU path saturates at
// ~27 B/clk/SM (4 tag look-ups + 4-5 shared-memory wavefronts per 512-byte warp instruction, zero-fills included),
// which capped the tensor pipe at ~20-37 %.  One producer warp: lane j owns rows 4j..4j+3 of the tile and issues one
// gather4 per stage; rows without a neighbour use an out-of-range row index and are zero-filled by the hardware.
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
    gather_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmx, const float* __restrict__ wimg,
                           const float* __restrict__ bias, const int* __restrict__ nbr, int64_t n_out,
                           const int* __restrict__ n_out_dev, int oob_row, int c_in, int c_out, int k3, int T_total,
                           int it_per_split, float* __restrict__ y) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  if ((int64_t)blockIdx.x * BM >= n_out) return;
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base + L::A_OFF, b_base = base + L::B_OFF;
  const uint32_t bar_base = base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int it0 = blockIdx.z * it_per_split;
  const int T = min(it_per_split, T_total - it0);
  if (T <= 0) return;
  const bool split = gridDim.z > 1;

  if (tid == 0) {
    prefetch_tmap(&tmx);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // one arrive.expect_tx; the TMA and bulk copies complete the bytes
      mbar_init(empty_bar(s), 1);   // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== producer warp: lane j gathers rows 4j .. 4j+3 =====================
    const int kc = c_in / BK;
    const int64_t r0 = m0 + 4 * lane;
    const int* nbr_t = nbr ? nbr + r0 : nullptr;
    auto load_group = [&](int k, int (&dst)[4]) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        int v = -1;
        if (k < k3 && r0 + p < n_out) v = nbr_t ? __ldg(nbr_t + (int64_t)k * pitch + p) : (int)(r0 + p);
        dst[p] = v >= 0 ? v : oob_row;
      }
    };
    int idx[4], idx1[4], idx2[4];
    int g = it0 / kc;
    load_group(g, idx);
    load_group(g + 1, idx1);
    load_group(g + 2, idx2);
    int s = 0, cc = 0;
    uint32_t ph = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (lane == 0) {
        mbar_arrive_expect_tx(full_bar(s), A_STAGE_BYTES + L::B_STAGE_BYTES);
        bulk_g2s(b_base + s * L::B_STAGE_BYTES, wimg + ((int64_t)(it0 + it) * c_out + n0) * BK, L::B_STAGE_BYTES,
                 full_bar(s));
      }
      __syncwarp();
      tma_gather4(a_base + s * A_STAGE_BYTES + lane * 512, &tmx, cc * BK, idx[0], idx[1], idx[2], idx[3], full_bar(s));
      if (++cc == kc) {
        cc = 0;
        ++g;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          idx[p] = idx1[p];
          idx1[p] = idx2[p];
        }
        load_group(g + 2, idx2);
      }
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
  if (warp < 4) {
    // ===================== epilogue: TMEM -> registers -> y =====================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int64_t o = m0 + warp * 32 + lane;
    const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
      if (o < n_out) {
        float* dst = y + o * c_out + n0 + c0;
        const bool add_bias = bias && it0 == 0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 r;
          r.x = __uint_as_float(v[j]) + (add_bias ? __ldg(&bias[n0 + c0 + j]) : 0.f);
          r.y = __uint_as_float(v[j + 1]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 1]) : 0.f);
          r.z = __uint_as_float(v[j + 2]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 2]) : 0.f);
          r.w = __uint_as_float(v[j + 3]) + (add_bias ? __ldg(&bias[n0 + c0 + j + 3]) : 0.f);
          if (split)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(r.x), "f"(r.y), "f"(r.z),
                         "f"(r.w)
                         : "memory");
          else
            *reinterpret_cast<float4*>(dst + j) = r;
        }
      }
    }
    tc_fence_before();
  } else {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC = idesc_tf32(BM, BN, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_desc = smem_desc_sw128(a_base + s * A_STAGE_BYTES, 16, 1024);
        const uint64_t b_desc = smem_desc_sw128(b_base + s * L::B_STAGE_BYTES, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk)
          mma_tf32(tmem_d, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), IDESC, (it | kk) ? 1u : 0u);
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
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<BN>(tmem_d);
  }
}

// 2-D tensor map over a row-major fp32 matrix [rows, cols] with box {BK columns, 1 row} and 128-byte swizzle
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool make_row_map(CUtensorMap* m, const float* base, int64_t rows, int cols) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, 1};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES>
int launch_tma(const float* x, int64_t n_in, const float* wimg, const float* bias, const int* nbr, int64_t n_out,
               const int* n_out_dev, int c_in, int c_out, int k3, int T, float* y, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gather_gemm_tma_kernel<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) != cudaSuccess) {
      b2s_set_error("conv_tc: cannot opt in to %d bytes of shared memory", L::DYN_BYTES);
      return -1;
    }
    attr_set = true;
  }
  CUtensorMap tmx;
  if (!make_row_map(&tmx, x, n_in, c_in)) {
    b2s_set_error("conv_tc: cuTensorMapEncodeTiled failed for a [%lld, %d] fp32 matrix", (long long)n_in, c_in);
    return -1;
  }
  const int64_t ctas = ceil_div64(n_out, BM) * (c_out / BN);
  const int kc = c_in / BK;
  int splits = 1;
  if (ctas < B2S_NUM_SMS) {
    splits = (int)((B2S_NUM_SMS + ctas - 1) / ctas);
    const int max_splits = T / 16 > 0 ? T / 16 : 1;
    if (splits > max_splits) splits = max_splits;
  }
  int per = (T + splits - 1) / splits;
  per = ((per + kc - 1) / kc) * kc;
  splits = (T + per - 1) / per;
  if (splits > 1) cudaMemsetAsync(y, 0, (size_t)n_out * c_out * sizeof(float), st);
  dim3 grid((unsigned)ceil_div64(n_out, BM), (unsigned)(c_out / BN), (unsigned)splits);
  kern<<<grid, TC_THREADS, L::DYN_BYTES, st>>>(tmx, wimg, bias, nbr, n_out, n_out_dev, (int)n_in, c_in, c_out, k3, T, per,
                                              y);
  return 0;
}


// ---------------------------------------------------------------------------------------------
// A operand in TENSOR MEMORY (split-bf16 mode, c_in a power-of-two multiple of 32, no split-K, no permutation).
// The shared-memory kernels above are bound by the L1 / shared-memory data pipe, not by the tensor cores: per
// (128-row tile, kernel offset) of a 64 -> 64 layer the LDGSTS gather writes 32 KB (in 64-byte beats), the six MMAs
// per stage read 6 KB of operands each (48 clocks at 128 B / clk against 32 clocks of tensor time: 72 KB per tile and
// offset) and the tensor pipe idles at 31 %.  Here the
// gathered rows never touch shared memory: four lanes load one row chunk (2 x LDG.128 each, a quad covers 64
// contiguous bytes per instruction) straight into the register fragment of tcgen05.st.16x256b, the MMAs take A from
// tensor memory (tcgen05.mma [d], [a], b-desc) and shared memory carries the weights only.
//   CTA = TILES x 128 out rows x BN channels, 4 TILES + 2 warps:
//   warps 0 .. 4 TILES - 1  producers + epilogue: warp w owns lanes 32 (w % 4) .. + 31 of tile w / 4 -- the quarter of
//              tensor memory a warp may address; per stage a thread loads 4 rows x 32 bytes and issues two 16-lane stores
//   next warp  MMA issuer (elect.sync lane): per stage and tile 6 x (M 128, N BN, K 16): h*H (k-steps 0,1),
//              l*H (2,3 x 0,1), h*L (0,1 x 2,3)
//   last warp  weight loader: one bulk copy per stage
// Tensor memory: accumulators in columns [0, TILES * BN), A ring (TA_SA stages x TILES tiles x 32 columns) behind them.
// Measured (tools/ta_bench.py, profiles/r02_ta_kernel_*.txt): L2 -> L1 traffic 1.84 -> 0.80 GB and 4 % less time than
// the shared-memory kernel on the 64 -> 64 layers; what holds it is the producers (row-load latency + instructions per
// gathered byte), not a saturated unit.
// The k-step positions of a row fragment are a fixed permutation of the chunk's channels (ta_channel); the weight
// image is built with the same permutation, so the products pair up unchanged.
// ---------------------------------------------------------------------------------------------
// TILES: 128-row accumulator tiles per CTA.  2 = one CTA per SM whose tiles share every weight stage; 1 = two CTAs per
// SM (6 warps, 168 registers), the prologue / epilogue of one overlapping the main loop of the other.
template <int BN, int SB, int TA_SA>
struct SmemTA {
  static constexpr int B_STAGE = BN * 128;
  static constexpr int B_OFF = 0;
  static constexpr int BAR_OFF = SB * B_STAGE;
  static constexpr int NBAR = 2 * TA_SA + 2 * SB + 2;
  static constexpr int TOTAL = BAR_OFF + NBAR * 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <int BN, int SB, int DEPTH, int TA_SA, int TILES>
__global__ void __launch_bounds__(TILES * 128 + 64, 3 - TILES)
    gather_gemm_ta_kernel(const float* __restrict__ x, const float* __restrict__ wimg, const float* __restrict__ bias,
                          const int* __restrict__ nbr, int64_t n_out, const int* __restrict__ n_out_dev, int c_in,
                          int c_out, int k3, int kc_shift, float* __restrict__ y) {
  const int64_t pitch = n_out;
  n_out = b2s_rows(n_out, n_out_dev);
  const int64_t m0 = (int64_t)blockIdx.x * (TILES * 128);
  if (m0 >= n_out) return;                       // uniform across the CTA
  const int n0 = blockIdx.y * BN;
  using L = SmemTA<BN, SB, TA_SA>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t b_base = base + L::B_OFF, bar_base = base + L::BAR_OFF;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (TA_SA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * TA_SA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * TA_SA + SB + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * TA_SA + 2 * SB);
  const uint32_t tmem_slot = bar_base + 8u * (2 * TA_SA + 2 * SB + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * TA_SA + 2 * SB + 1));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = k3 << kc_shift;

  if (tid == 0) {
    for (int s = 0; s < TA_SA; ++s) {
      mbar_init(a_full(s), TILES * 128);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  constexpr int TNEED = TILES * (BN + TA_SA * 32);
  constexpr int TCOLS = TNEED <= 128 ? 128 : (TNEED <= 256 ? 256 : 512);
  constexpr uint32_t A_COL = TILES * BN;
  constexpr int MMA_WARP = TILES * 4;
  if (warp == MMA_WARP) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp < MMA_WARP) {
    // ===================== producers =====================
    const int tile = warp >> 2, quarter = warp & 3, qd = lane & 3;
    const int64_t r0 = m0 + tile * 128 + quarter * 32 + (lane >> 2);      // this thread's rows: r0 + {0, 8, 16, 24}
    const uint32_t t_quarter = tmem_d + ((uint32_t)(quarter * 32) << 16);
    const int kcm = (1 << kc_shift) - 1;
    auto ldidx = [&](int it, int (&ix)[4]) {
      const int k = it >> kc_shift;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t o = r0 + 8 * q;
        ix[q] = (it < T && o < n_out) ? __ldg(nbr + (int64_t)k * pitch + o) : -1;
      }
    };
    auto ldrows = [&](int it, const int (&ix)[4], uint4 (&v)[8]) {
      const int cc = it & kcm;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ix[q] >= 0) {          // bytes [16 qd, + 16) of the h half and of the l half of the 128-byte row chunk
          const uint4* p = reinterpret_cast<const uint4*>(x + (int64_t)ix[q] * c_in + cc * BK) + qd;
          v[2 * q] = __ldg(p);
          v[2 * q + 1] = __ldg(p + 4);
        } else {
          v[2 * q] = make_uint4(0u, 0u, 0u, 0u);
          v[2 * q + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    };
    auto store = [&](uint32_t taddr, const uint4 (&v)[8]) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {   // lanes 16 g .. 16 g + 15 of the quarter: rows r0 + 16 g (a) and r0 + 16 g + 8 (b)
        const uint4 &a0 = v[4 * g], &a1 = v[4 * g + 1], &b0 = v[4 * g + 2], &b1 = v[4 * g + 3];
        const uint32_t r[16] = {a0.x, a0.y, b0.x, b0.y, a0.z, a0.w, b0.z, b0.w,
                                a1.x, a1.y, b1.x, b1.y, a1.z, a1.w, b1.z, b1.w};
        tmem_st_16x256b_x4(taddr + ((uint32_t)(g * 16) << 16), r);
      }
    };
    uint4 d[DEPTH][8];
    int wq[DEPTH][4];
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) ldidx(j, wq[j]);
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) {
      ldrows(j, wq[j], d[j]);
      ldidx(j + DEPTH, wq[j]);
    }
#pragma unroll 1
    for (int it0 = 0; it0 < T; it0 += DEPTH) {
#pragma unroll
      for (int j = 0; j < DEPTH; ++j) {
        const int it = it0 + j;
        if (it < T) {
          const int sa = it % TA_SA;
          mbar_wait(a_empty(sa), (((uint32_t)(it / TA_SA)) & 1u) ^ 1u);
          tc_fence_after();
          store(t_quarter + A_COL + (uint32_t)((sa * TILES + tile) * 32), d[j]);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(a_full(sa));
          ldrows(it + DEPTH, wq[j], d[j]);       // stage it + DEPTH (indices past the last stage are -1: no loads)
          ldidx(it + 2 * DEPTH, wq[j]);
        }
      }
    }
    // ===================== epilogue: warp = (tile, lane quarter) =====================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int64_t orow = m0 + tile * 128 + quarter * 32 + lane;
    const uint32_t t_lane = t_quarter + (uint32_t)(tile * BN);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)c0, v);
      tmem_ld_wait();
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
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC = idesc_bf16(128, BN, 0, 0);
    int sb = 0;
    uint32_t phb = 0;
    for (int it = 0; it < T; ++it) {
      const int sa = it % TA_SA;
      mbar_wait(b_full(sb), phb);
      mbar_wait(a_full(sa), ((uint32_t)(it / TA_SA)) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t b_desc = smem_desc_sw128(b_base + sb * L::B_STAGE, 16, 1024);
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          const uint32_t a_t = tmem_d + A_COL + (uint32_t)((sa * TILES + t) * 32);
#pragma unroll
          for (int q = 0; q < 6; ++q) {          // K = 16 bf16 = 8 columns of tensor memory / 32 bytes of the image row
            const int ak = q < 4 ? q : q - 4, bk = q < 2 ? q : q - 2;
            mma_bf16_ta(tmem_d + (uint32_t)(t * BN), a_t + (uint32_t)(ak * 8), b_desc + (uint64_t)(bk * 2), IDESC,
                        (it | q) ? 1u : 0u);
          }
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
      mbar_arrive_expect_tx(b_full(sb), L::B_STAGE);
      bulk_g2s(b_base + sb * L::B_STAGE, wimg + ((int64_t)it * c_out + n0) * BK, L::B_STAGE, b_full(sb));
      if (++sb == SB) {
        sb = 0;
        phb ^= 1u;
      }
    }
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_d);
  }
}

int tc_knob(int global, const char* env_name, int dflt) {
  if (global >= 0) return global;
  const char* e = getenv(env_name);
  return e ? atoi(e) : dflt;
}
int tc_ca() { return tc_knob(g_b2s_tc_ca, "B2S_TC_CA", 0); }
int tc_occ1() { return tc_knob(g_b2s_tc_occ1, "B2S_TC_OCC1", 0); }

template <int BN, int STAGES>
int launch_perm(const float* x, const float* wimg, const int* nbr, int64_t n_out, const int* n_out_dev, int c_in,
                int c_out, int k3, float* y, const PermArgs& pa, int64_t tiles_cap, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gather_gemm_tc_kernel<BN, STAGES, false, 2, true>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) != cudaSuccess) {
      b2s_set_error("conv_tc: cannot opt in to %d bytes of shared memory", L::DYN_BYTES);
      return -1;
    }
    attr_set = true;
  }
  dim3 grid((unsigned)tiles_cap, (unsigned)(c_out / BN), 1);
  kern<<<grid, TC_THREADS, L::DYN_BYTES, st>>>(x, wimg, nullptr, nbr, n_out, n_out_dev, c_in, c_out, k3, 0, 0, y, pa,
                                               (tc_ca() ? 2 : 0) | (b2s_precise() ? 4 : 0), nullptr);
  return 0;
}

bool tc_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2S_DISABLE_TC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int tc_rot() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("B2S_TC_ROT");
    env = e ? atoi(e) : 0;
  }
  return g_b2s_tc_rot >= 0 ? g_b2s_tc_rot : env;
}

template <int BN, int STAGES, int LAG>
int launch_tc2(const float* x, const float* wimg, const float* bias, const int* nbr, int64_t n_out, const int* n_out_dev,
               int c_in, int c_out, int k3, int T, float* y, cudaStream_t st, float* col_stats = nullptr,
               int* stats_rows = nullptr) {
  using L = SmemLayout2<BN, STAGES>;
  static_assert(LAG < STAGES, "producers run LAG stages ahead of their hand-over");
  auto kern = gather_gemm_tc2_kernel<BN, STAGES, LAG>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) != cudaSuccess) {
      b2s_set_error("conv_tc: cannot opt in to %d bytes of shared memory", L::DYN_BYTES);
      return -1;
    }
    attr_set = true;
  }
  const int64_t ctas = ceil_div64(n_out, 2 * BM) * (c_out / BN);
  const int kc = c_in / BK;
  int splits = 1;
  if (ctas < B2S_NUM_SMS) {   // too few output tiles for 148 SMs: split the K loop, but stay within ONE wave (one CTA
    splits = (int)(B2S_NUM_SMS / ctas);             // per SM: a 149th CTA would double the kernel's time)
    const int max_splits = T / 16 > 0 ? T / 16 : 1;                 // >= 16 stages per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int per = (T + splits - 1) / splits;
  per = ((per + kc - 1) / kc) * kc;
  splits = (T + per - 1) / per;
  if (splits > 1) cudaMemsetAsync(y, 0, (size_t)n_out * c_out * sizeof(float), st);
  dim3 grid((unsigned)ceil_div64(n_out, 2 * BM), (unsigned)(c_out / BN), (unsigned)splits);
  const bool fuse = col_stats != nullptr && splits == 1;      // split-K tiles are partial sums: no statistics there
  if (stats_rows) *stats_rows = fuse ? 2 * BM : 0;            // rows per partial row of col_stats (0: not produced)
  kern<<<grid, TC2_THREADS, L::DYN_BYTES, st>>>(x, wimg, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, per, y,
                                                b2s_precise(), fuse ? col_stats : nullptr);
  return 0;
}

int tc_m256() { return tc_knob(g_b2s_tc_m256, "B2S_TC_M256", 1); }
// "tc_ta": 0 = off, 1 = 64-wide output tiles take the A operand from tensor memory (default), 2 = 128-wide ones too,
// 3 = as 2 and also maps with fewer CTAs than SMs (tests)
int tc_ta() { return tc_knob(g_b2s_tc_ta, "B2S_TC_TA", 1); }

// shapes gather_gemm_ta_kernel covers: split-bf16 operands, c_in = 32 * 2^s, at least one 256-row CTA per SM (no
// split-K in that kernel), a neighbour table, no statistics epilogue
bool ta_applies(int c_in, int c_out, int64_t n_out, bool has_nbr, bool stats) {
  const int mode = tc_ta();
  if (!mode || !b2s_precise() || !has_nbr || stats || c_in < BK || c_in % BK != 0) return false;
  const int kc = c_in / BK;
  if (kc & (kc - 1)) return false;
  const int bn = c_out % 256 == 0 ? 256 : (c_out % 128 == 0 ? 128 : 64);
  if (!(bn == 64 || (bn == 128 && mode >= 2))) return false;
  return mode >= 3 || ceil_div64(n_out, 2 * BM) * (c_out / bn) >= B2S_NUM_SMS;
}

template <int BN, int SB, int DEPTH, int SA, int TILES>
int launch_ta(const float* x, const float* wimg, const float* bias, const int* nbr, int64_t n_out, const int* n_out_dev,
              int c_in, int c_out, int k3, float* y, cudaStream_t st) {
  using L = SmemTA<BN, SB, SA>;
  auto kern = gather_gemm_ta_kernel<BN, SB, DEPTH, SA, TILES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) != cudaSuccess) {
      b2s_set_error("conv_tc: cannot opt in to %d bytes of shared memory", L::DYN_BYTES);
      return -1;
    }
    attr_set = true;
  }
  int sh = 0;
  while ((BK << sh) < c_in) ++sh;
  dim3 grid((unsigned)ceil_div64(n_out, TILES * BM), (unsigned)(c_out / BN), 1);
  kern<<<grid, TILES * 128 + 64, L::DYN_BYTES, st>>>(x, wimg, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, sh, y);
  return 0;
}

template <int BN, int STAGES, bool SMALL, int LAG = 2, int NB = 1, int PW = 4>
int launch_tc(const float* x, const float* wimg, const float* bias, const int* nbr, int64_t n_out, const int* n_out_dev,
              int c_in, int c_out, int k3, int T, float* y, cudaStream_t st, float* col_stats = nullptr,
              int* stats_rows = nullptr) {
  using L = SmemLayout<BN, STAGES, NB>;
  static_assert(LAG < STAGES, "producers run LAG stages ahead of their hand-over");
  auto kern = gather_gemm_tc_kernel<BN, STAGES, SMALL, LAG, false, NB, PW>;
  static bool attr_set = false;
  if (!attr_set) {
    constexpr int OPT_IN = L::DYN_BYTES > 116 * 1024 ? L::DYN_BYTES : 116 * 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, OPT_IN) != cudaSuccess) {
      b2s_set_error("conv_tc: cannot opt in to %d bytes of shared memory", OPT_IN);
      return -1;
    }
    attr_set = true;
  }
  // tc_occ1: ask for more than half of the SM's shared memory so that ONE CTA runs per SM and the rest of the
  // unified array stays L1 (the .ca gathers live there)
  const int dyn = (tc_occ1() && L::DYN_BYTES < 116 * 1024) ? 116 * 1024 : L::DYN_BYTES;
  const int64_t ctas = ceil_div64(n_out, BM) * (c_out / BN);
  const int kc = SMALL ? 1 : c_in / BK;
  int splits = 1;
  if (ctas < B2S_NUM_SMS) {                      // too few output tiles for 148 SMs: split the K loop
    splits = (int)((B2S_NUM_SMS + ctas - 1) / ctas);
    const int max_splits = T / 16 > 0 ? T / 16 : 1;                 // >= 16 stages per split
    if (splits > max_splits) splits = max_splits;
  }
  int per = (T + splits - 1) / splits;
  per = ((per + kc - 1) / kc) * kc;
  splits = (T + per - 1) / per;
  if (splits > 1) cudaMemsetAsync(y, 0, (size_t)n_out * c_out * sizeof(float), st);
  dim3 grid((unsigned)ceil_div64(n_out, BM), (unsigned)(c_out / BN), (unsigned)splits);
  const bool fuse = col_stats != nullptr && splits == 1;      // split-K tiles are partial sums: no statistics there
  if (stats_rows) *stats_rows = fuse ? BM : 0;
  kern<<<grid, (PW + 1) * 32, dyn, st>>>(x, wimg, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, per, y, PermArgs{},
                                      (tc_rot() ? 1 : 0) | (tc_ca() ? 2 : 0) | (b2s_precise() ? 4 : 0),
                                      fuse ? col_stats : nullptr);
  return 0;
}

}  // namespace

bool b2s_conv_tc_supported(int32_t c_in, int32_t c_out, int32_t k3, int64_t n_out) {
  (void)k3;
  if (tc_disabled() || n_out <= 0) return false;
  if (c_out % 64 != 0) return false;
  return c_in <= 4 || c_in % BK == 0;
}

static inline int64_t align256(int64_t b) { return (b + 255) & ~(int64_t)255; }

static int iterations(int c_in, int k3) { return c_in <= 4 ? (k3 + 7) / 8 : k3 * (c_in / BK); }

int64_t b2s_conv_tc_image_bytes(int32_t c_in, int32_t c_out, int32_t k3) {   // small c_in: room for SMALL_NB images
  return align256((int64_t)iterations(c_in, k3) * (c_in <= 4 ? SMALL_NB : 1) * c_out * 128);
}

// A prebuilt image (b2s_conv_weight_image) does not know the row count of the call that will use it, hence not which
// kernel: for the shapes gather_gemm_ta_kernel may take it holds both forms back to back.
static bool ta_shape(int c_in, int c_out) {
  const int kc = c_in / BK;
  return c_in >= BK && c_in % BK == 0 && !(kc & (kc - 1)) && c_out % 64 == 0 && c_out % 256 != 0;
}
int64_t b2s_conv_tc_prebuilt_image_bytes(int32_t c_in, int32_t c_out, int32_t k3) {
  return b2s_conv_tc_image_bytes(c_in, c_out, k3) * (ta_shape(c_in, c_out) ? 2 : 1);
}

int64_t b2s_conv_tc_workspace_bytes(int32_t c_in, int32_t c_out, int32_t k3, int64_t n_in) {
  return b2s_conv_tc_image_bytes(c_in, c_out, k3) + (c_in <= 4 ? align256(n_in * 16) : 0);
}

static void launch_prep_weights(const float* w, int c_in, int c_out, int k3, int w_layout, bool small, int T, float* img,
                                cudaStream_t st, bool ta = false) {
  const int precise = b2s_precise() ? (ta ? 2 : 1) : 0;     // 2: rows in the channel order of gather_gemm_ta_kernel
  if (!small && !(w_layout & 1) && c_in % BK == 0 && c_out % 32 == 0)
    prep_weights_t_kernel<<<dim3((unsigned)T, (unsigned)(c_out / 32)), 256, 0, st>>>(w, c_in, c_out, k3, w_layout,
                                                                                     precise, img);
  else
    prep_weights_kernel<<<grid_for((int64_t)T * (small && precise ? SMALL_NB : 1) * c_out * BK, 256), 256, 0, st>>>(
        w, c_in, c_out, k3, w_layout, small ? 1 : 0, T, precise, img);
}

// the weight image of one convolution (what b2s_conv_gather_gemm builds at the head of its workspace), built ahead
int b2s_conv_weight_image_tc(const float* w, int32_t c_in, int32_t c_out, int32_t k3, int32_t w_layout, float* img,
                             cudaStream_t st) {
  launch_prep_weights(w, c_in, c_out, k3, w_layout & 3, false, iterations(c_in, k3), img, st);
  if (ta_shape(c_in, c_out) && b2s_precise())
    launch_prep_weights(w, c_in, c_out, k3, w_layout & 3, false, iterations(c_in, k3),
                        img + b2s_conv_tc_image_bytes(c_in, c_out, k3) / sizeof(float), st, true);
  return 0;
}

int b2s_conv_gather_gemm_tc(const float* x, const float* w, const float* bias, const int32_t* nbr, int64_t n_in,
                            int64_t n_out, const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, int32_t w_layout, float* y,
                            void* workspace, int64_t workspace_bytes, cudaStream_t st, float* col_stats,
                            int* stats_rows) {
  if (stats_rows) *stats_rows = 0;
  const bool small = c_in <= 4;
  const int T = iterations(c_in, k3);
  float* img = reinterpret_cast<float*>(workspace);
  bool ta = !small && ta_applies(c_in, c_out, n_out, nbr != nullptr, col_stats != nullptr);
  if ((w_layout & 16) && workspace_bytes < b2s_conv_tc_prebuilt_image_bytes(c_in, c_out, k3)) ta = false;   // one form only
  if (!(w_layout & 16)) launch_prep_weights(w, c_in, c_out, k3, w_layout & 3, small, T, img, st, ta);   // bit 4: image prebuilt
  else if (ta) img += b2s_conv_tc_image_bytes(c_in, c_out, k3) / sizeof(float);   // ... in both forms: the second one
  const float* xin = x;
  if (small) {
    float4* x4 = reinterpret_cast<float4*>(reinterpret_cast<char*>(workspace) + b2s_conv_tc_image_bytes(c_in, c_out, k3));
    pad_rows4_kernel<<<grid_for(n_in, 256), 256, 0, st>>>(x, n_in, c_in, b2s_precise(), x4);
    xin = reinterpret_cast<const float*>(x4);
  }
  const int bn = c_out % 256 == 0 ? 256 : (c_out % 128 == 0 ? 128 : 64);
  if (ta) {
    static int tiles = -1;
    if (tiles < 0) {
      const char* e = getenv("B2S_TA_TILES");
      tiles = e ? atoi(e) : 1;
    }
    if (bn == 128) return launch_ta<128, 5, 3, 4, 2>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, y, st);
    if (tiles == 2) return launch_ta<64, 8, 3, 4, 2>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, y, st);
    if (tiles == 12) return launch_ta<64, 6, 3, 2, 1>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, y, st);
    return launch_ta<64, 6, 3, 4, 1>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, y, st);
  }
  if (small && b2s_precise()) {   // three weight images per stage: one CTA per SM, deeper ring
    if (bn == 256) return launch_tc<256, 2, true, 1, SMALL_NB, 8>(xin, img, bias, nbr, n_out, n_out_dev, 4, c_out, k3, T, y, st);
    if (bn == 128) return launch_tc<128, 3, true, 2, SMALL_NB, 8>(xin, img, bias, nbr, n_out, n_out_dev, 4, c_out, k3, T, y, st);
    return launch_tc<64, 5, true, 3, SMALL_NB, 8>(xin, img, bias, nbr, n_out, n_out_dev, 4, c_out, k3, T, y, st);
  }
  if (small) {
    if (bn == 256) return launch_tc<256, 4, true>(xin, img, bias, nbr, n_out, n_out_dev, 4, c_out, k3, T, y, st);
    if (bn == 128) return launch_tc<128, 3, true>(xin, img, bias, nbr, n_out, n_out_dev, 4, c_out, k3, T, y, st);
    return launch_tc<64, 4, true>(xin, img, bias, nbr, n_out, n_out_dev, 4, c_out, k3, T, y, st);
  }
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("B2S_TC_VARIANT");
    variant = e ? atoi(e) : 0;
  }
  if (variant >= 10 && n_in > 0 && !b2s_precise()) {   // TMA row gather (TF32 operands only) (variant 10: default stages; 11: deeper)
    if (bn == 256) return launch_tma<256, 4>(xin, n_in, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
    if (bn == 128) return variant == 11 ? launch_tma<128, 6>(xin, n_in, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st)
                                        : launch_tma<128, 3>(xin, n_in, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
    return variant == 11 ? launch_tma<64, 8>(xin, n_in, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st)
                         : launch_tma<64, 4>(xin, n_in, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
  }
  // M = 256 tiles: measured 1.5x faster for 128-wide output tiles (0.193 -> 0.127 ms at 69.5 k rows x 128 -> 128),
  // neutral for 64-wide ones (bound by the LSU gather, not by weight traffic) and 1.5x SLOWER for 256-wide ones
  // (three 64 KB stages, split-K reductions) -- knob "tc_m256": 0 off, 1 = 128-wide tiles only (default), 2 = wherever
  // it can run (tests), 3 = 64- and 128-wide tiles
  const int m256 = tc_m256();
  if (m256 && n_out > BM && (bn == 128 || m256 == 2 || (m256 == 3 && bn == 64))) {
    if (bn == 256) return launch_tc2<256, 3, 2>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st, col_stats, stats_rows);
    if (bn == 128) return launch_tc2<128, 4, 2>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st, col_stats, stats_rows);
    return launch_tc2<64, 5, 3>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st, col_stats, stats_rows);
  }
  if (bn == 256) return launch_tc<256, 4, false>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st, col_stats, stats_rows);
  if (bn == 128) {
    if (variant == 1) return launch_tc<128, 3, false, 2>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
    if (variant == 2) return launch_tc<128, 6, false, 4>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
    if (variant == 3) return launch_tc<128, 6, false, 5>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
    return launch_tc<128, 3, false>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st, col_stats, stats_rows);
  }
  if (variant == 1) return launch_tc<64, 3, false, 2>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
  if (variant == 2) return launch_tc<64, 4, false, 3>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
  if (variant == 3) return launch_tc<64, 8, false, 6>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st);
  return launch_tc<64, 4, false>(xin, img, bias, nbr, n_out, n_out_dev, c_in, c_out, k3, T, y, st, col_stats, stats_rows);
}


// dgrad of a stride-2 convolution through the parity plan (see PermArgs).  x = rounded grad_out [n_coarse, c_in],
// nbr = transposed table [k3, n_fine], y = grad_in [n_fine, c_out]; w_layout as in b2s_conv_gather_gemm.
int b2s_conv_dgrad_perm_tc(const float* x, const float* w, const int32_t* nbr, int64_t n_fine, const int32_t* n_fine_dev,
                           int32_t c_in, int32_t c_out, const int32_t* ksize, int32_t w_layout, const int32_t* perm,
                           const int32_t* bounds, float* y, void* workspace, cudaStream_t st) {
  const int k3 = ksize[0] * ksize[1] * ksize[2];
  const int T = iterations(c_in, k3);
  float* img = reinterpret_cast<float*>(workspace);
  if (!(w_layout & 16)) launch_prep_weights(w, c_in, c_out, k3, w_layout & 3, false, T, img, st);
  PermArgs pa{};
  pa.perm = perm;
  pa.bounds = bounds;
  // class bit d set <=> the fine coordinate is OFF the coarse lattice in dimension d; the partner offsets are those
  // whose centred index has the same parity as the position inside the cell
  for (int c = 0; c < 8; ++c) {
    int n = 0;
    for (int iz = 0; iz < ksize[2]; ++iz)
      for (int iy = 0; iy < ksize[1]; ++iy)
        for (int ix = 0; ix < ksize[0]; ++ix) {
          const int i3[3] = {ix, iy, iz};
          bool ok = true;
          for (int d = 0; d < 3; ++d) {
            const int centred = (ksize[d] & 1) ? i3[d] - ksize[d] / 2 : i3[d];
            if ((abs(centred) & 1) != ((c >> d) & 1)) ok = false;
          }
          if (ok) pa.klist[c][n++] = (unsigned char)(ix + ksize[0] * (iy + ksize[1] * iz));
        }
    pa.nk[c] = (unsigned char)n;
  }
  const int64_t tiles_cap = ceil_div64(n_fine, BM) + 8;
  const int bn = c_out % 256 == 0 ? 256 : (c_out % 128 == 0 ? 128 : 64);
  if (bn == 256) return launch_perm<256, 4>(x, img, nbr, n_fine, n_fine_dev, c_in, c_out, k3, y, pa, tiles_cap, st);
  if (bn == 128) return launch_perm<128, 3>(x, img, nbr, n_fine, n_fine_dev, c_in, c_out, k3, y, pa, tiles_cap, st);
  return launch_perm<64, 4>(x, img, nbr, n_fine, n_fine_dev, c_in, c_out, k3, y, pa, tiles_cap, st);
}

// wgrad on tensor cores: see wgrad_tc.cu
