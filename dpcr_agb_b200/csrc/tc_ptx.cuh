// tc_ptx.cuh -- hand-written PTX wrappers for the Blackwell (sm_100a) tensor-core path:
// mbarrier, cp.async (LDGSTS) with zero-fill, 1-D bulk copy (UBLKCP), tcgen05 alloc/mma/commit/ld, and the
// shared-memory / instruction descriptors for kind::tf32.  Bit layouts follow the PTX ISA "tcgen05"
// matrix-descriptor and instruction-descriptor tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ mbarrier -----------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------ async copies -------------
// 16-byte LDGSTS; src_bytes == 0 writes 16 zero bytes (used for rows without a neighbour)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// same copy, allocating the line in L1 as well: gathers of neighbouring kernel offsets touch the same feature rows
// again within a few pipeline stages, and an L1 hit takes that re-read off the L2 fabric
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_sel(uint32_t dst, const void* src, uint32_t src_bytes, bool l1) {
  if (l1) cp_async16_ca(dst, src, src_bytes);
  else cp_async16(dst, src, src_bytes);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes (cp.async / st.shared) visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// round-to-nearest (ties away) fp32 -> tf32 bit pattern.  tcgen05 kind::tf32 TRUNCATES the low 13 mantissa bits of
// its fp32 operands (measured: tools/umma_probe.cu, "rounding probe"), which biases every product towards zero;
// operands are therefore rounded explicitly before the tensor core sees them.
__device__ __forceinline__ uint32_t rna_tf32(uint32_t v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(__uint_as_float(v)));
  return r;
}
// in-place rna rounding of one 16-byte chunk of shared memory (a chunk this thread's own LDGSTS has filled)
__device__ __forceinline__ void round_chunk_tf32(uint32_t saddr) {
  uint32_t a, b, c, d;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(saddr) : "memory");
  a = rna_tf32(a); b = rna_tf32(b); c = rna_tf32(c); d = rna_tf32(d);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// TMA row gather: 4 rows (row indices r0..r3, any order, out-of-range rows are zero-filled) x one box of columns
// starting at `col` of a 2-D tensor map whose box is {box_cols, 1}, written as 4 consecutive rows at `dst` with the
// map's shared-memory swizzle applied; bytes complete on the mbarrier (SASS: UTMALDG ... gather4).
__device__ __forceinline__ void tma_gather4(uint32_t dst, const void* tmap, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// One lane of a CONVERGED warp (elect.sync).  Single-thread tcgen05 instructions (mma, commit) are uniform-datapath
// instructions: under a data-dependent branch such as `lane == 0` the compiler wraps every one of them in a
// divergence loop (ELECT / BRA.U.ANY, about 140 clocks per MMA measured with tools/mma_rate_probe.cu, which caps a
// CTA at one N = 64 MMA per 140 clocks whatever the operands do); under the predicate of elect.sync it issues them
// back to back.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ tcgen05 ------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, kind::f16 (here: bf16 inputs, fp32 accumulate; K = 16 per instruction) -- the split-bf16 operand mode
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all tcgen05 ops issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc], bf16 inputs, fp32 accumulate: the A operand (128 lanes = rows, 8 columns = 16
// bf16 of K per instruction) is read from tensor memory instead of shared memory
__device__ __forceinline__ void mma_bf16_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 lanes x 128 bytes (32 columns) per warp instruction, fragment layout of the 16x256b shape repeated four times
// along the columns: thread t holds, for j = 0..3, columns 8 j + 2 (t % 4) + {0, 1} of lane t / 4 in r[4 j + {0, 1}]
// and of lane t / 4 + 8 in r[4 j + {2, 3}].  The lane field of taddr is the first of the 16 lanes (a multiple of 16
// inside the warp's own 32-lane quarter).
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ------------------------------------------------------------------ descriptors --------------
// Shared-memory matrix descriptor (64-bit):
//   [0,14)  start address >> 4          [16,30) leading-dimension byte offset >> 4
//   [32,46) stride-dimension byte offset >> 4     [46,48) version = 1 on sm_100
//   [49,52) base offset (0: tiles are 1024-byte aligned)      [61,64) swizzle: 2 = 128-byte
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MN-major operands of 32-bit types (tf32) must use the "128-byte swizzle with 32-byte base" layout (type 1,
// Swizzle<2,5,2>): atoms of 4 K-rows x 128 bytes, the four 32-byte units of a row XOR-ed with (row & 3).
// LBO = byte stride between atoms along M/N (next 32 elements), SBO = stride between atoms along K (next 4
// rows).  Plain SWIZZLE_128B (type 2) with MN-major tf32 silently yields zeros (measured, tools/umma_probe.cu).
__device__ __forceinline__ uint64_t smem_desc_sw128_base32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

// Instruction descriptor (32-bit) for kind::tf32, fp32 accumulate:
//   [4,6) D format: 1 = F32     [7,10) A format: 2 = TF32     [10,13) B format: 2 = TF32
//   [15] A major (0 = K, 1 = MN)   [16] B major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// kind::f16 with bf16 inputs and fp32 accumulate: A / B format 1 = BF16, otherwise the same fields
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------ split-bf16 operands ------
// "Precise" operand mode (b2s_set_tuning("precise", 1), the default): an fp32 value v travels as the bf16 pair
// (h, l) with h = bf16(v), l = bf16(v - h), i.e. 16-17 significant bits in the same four bytes, and a product is
// evaluated as h*H + l*H + h*L on the kind::f16 tensor-core path (bf16 products are exact in the fp32 accumulator).
// Per operand the error is <= 2^-17 |v| against 2^-11 |v| of a TF32 operand, which is what it takes to hold 1e-3 on
// every gradient of a training step END TO END (tests/test_gpu_model.py::test_full_size_training_parity_fp32).
// h and l are returned in the low 16 bits.
__device__ __forceinline__ uint32_t bf16_rn_bits(float v) {
  uint32_t r;
  asm("{\n\t.reg .b16 t;\n\tcvt.rn.bf16.f32 t, %1;\n\tmov.b32 %0, {t, 0};\n\t}" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void split_bf16(float v, uint32_t& h, uint32_t& l) {
  h = bf16_rn_bits(v);
  l = bf16_rn_bits(v - __uint_as_float(h << 16));
}
// three-term split (24 significant bits) for the k7 stem's weights
__device__ __forceinline__ void split_bf16x3(float v, uint32_t& h, uint32_t& m, uint32_t& l) {
  split_bf16(v, h, m);
  l = bf16_rn_bits((v - __uint_as_float(h << 16)) - __uint_as_float(m << 16));
}

// Column sums of a 32 x 32 register tile held one row per lane: on return lane j holds the total of v[j] over the 32
// lanes (butterfly transpose-reduce: 31 shuffles instead of 32 x 5).  v is clobbered.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// Batch-norm statistics in a convolution epilogue: a warp holds 32 rows x 32 columns of final output values (lane =
// row, invalid rows zeroed); per column the sum and the sum of squares over the warp's rows go to `red[warp][col]`.
__device__ __forceinline__ void epilogue_col_stats(const float (&r)[32], int lane, float2* red_row) {
  float f[32], q[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    f[j] = r[j];
    q[j] = r[j] * r[j];
  }
  const float s1 = warp_transpose_sum(f, lane), s2 = warp_transpose_sum(q, lane);
  red_row[lane] = make_float2(s1, s2);
}

// byte offset of 16-byte chunk `chunk` (0..7) of 128-byte row `row` inside a 128B-swizzled tile
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

}  // namespace tc
