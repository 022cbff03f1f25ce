// umma_probe.cu -- single-CTA probe of tcgen05 kind::tf32 with MN-major operands: tries descriptor / layout
// variants and reports which reproduces D[m][n] = sum_k A[k][m] * B[k][n].  Debug tool, not part of the library.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../dpcr_agb_b200/csrc/tc_ptx.cuh"
using namespace tc;

constexpr int M = 128, N = 64, KR = 32;  // KR reduction rows per stage (4 MMAs of K=8)

struct Variant {
  int layout;        // 0: [j][g] atoms (MN atom outer), 1: [g][j] atoms (K atom outer)
  uint32_t lbo, sbo; // descriptor fields (bytes)
  uint32_t adv;      // byte advance of the start address per MMA (per K atom of 8 rows)
  int a_major, b_major;
  int swz;           // 1: 128B swizzle, 0: none(interleave) -- only 1 used here
};

__device__ __forceinline__ uint32_t off_mn(int layout, int row, int chunk, int natoms_mn) {
  const int j = chunk >> 3, c = chunk & 7, g = row >> 3, rr = row & 7;
  if (layout == 2) {  // 128B swizzle with 32-byte base (Swizzle<2,5,2>): atoms of 4 rows x 128 B, MN atom outer
    const int u = c >> 1, h = c & 1;
    return (uint32_t)j * (KR * 128u) + (uint32_t)row * 128u + (uint32_t)((((u ^ (row & 3)) << 1) | h) << 4);
  }
  uint32_t atom = layout == 0 ? (uint32_t)(j * (KR / 8) + g) : (uint32_t)(g * natoms_mn + j);
  return atom * 1024u + (uint32_t)rr * 128u + (uint32_t)((c ^ rr) << 4);
}
__device__ __forceinline__ uint64_t desc_any(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)type << 61;
  return d;
}

__global__ void __launch_bounds__(160, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ D, Variant v) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base, b_base = base + M * KR * 4, bar = b_base + N * KR * 4, slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + M * KR * 4 + N * KR * 4 + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc<N>(slot);
  // fill smem with plain stores (generic proxy), then proxy fence
  for (int e = tid; e < KR * (M / 4); e += 160) {
    const int row = e / (M / 4), chunk = e % (M / 4);
    float4 val = *reinterpret_cast<const float4*>(A + row * M + chunk * 4);
    *reinterpret_cast<float4*>(smem + off_mn(v.layout, row, chunk, M / 32)) = val;
  }
  for (int e = tid; e < KR * (N / 4); e += 160) {
    const int row = e / (N / 4), chunk = e % (N / 4);
    float4 val = *reinterpret_cast<const float4*>(B + row * N + chunk * 4);
    *reinterpret_cast<float4*>(smem + M * KR * 4 + off_mn(v.layout, row, chunk, N / 32)) = val;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *slot_ptr;
  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(M, N, v.a_major, v.b_major);
      for (int g = 0; g < KR / 8; ++g) {
        const uint64_t ad = desc_any(a_base + g * v.adv, v.lbo, v.sbo, v.swz);
        const uint64_t bd = desc_any(b_base + g * v.adv, v.lbo, v.sbo, v.swz);
        mma_tf32(tmem_d, ad, bd, idesc, g ? 1u : 0u);
      }
      mma_commit(bar);
    }
    __syncwarp();
  } else {
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc<N>(tmem_d); }
}

int main() {
  std::vector<float> A(KR * M), B(KR * N), ref(M * N), out(M * N);
  for (int k = 0; k < KR; ++k) for (int m = 0; m < M; ++m) A[k * M + m] = (float)(((k * 7 + m * 3) % 11) - 5);
  for (int k = 0; k < KR; ++k) for (int n = 0; n < N; ++n) B[k * N + n] = (float)(((k * 5 + n * 13) % 7) - 3);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    float s = 0; for (int k = 0; k < KR; ++k) s += A[k * M + m] * B[k * N + n]; ref[m * N + n] = s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = M * KR * 4 + N * KR * 4 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const uint32_t g4 = (KR / 8) * 1024;   // 4096: stride between MN atoms in layout 0
  Variant vs[] = {
      {0, g4, 1024, 1024, 1, 1, 2},      // V0: SWIZZLE_128B (type 2), as first tried in wgrad_tc.cu
      {2, g4, 512, 1024, 1, 1, 1},       // V1: SWIZZLE_128B_BASE32B (type 1): LBO = next MN atom, SBO = next 4 rows
      {2, 512, g4, 1024, 1, 1, 1},       // V2: same, LBO/SBO swapped
      {2, g4, 1024, 1024, 1, 1, 1},      // V3: type 1 with SBO = 8 rows
      {0, g4, 1024, 1024, 1, 1, 1},      // V4: type 1 descriptor over the 16B-swizzled bytes (expected wrong)
  };
  for (int vi = 0; vi < (int)(sizeof(vs) / sizeof(vs[0])); ++vi) {
    cudaMemset(dD, 0xFF, out.size() * 4);
    probe_kernel<<<1, 160, smem>>>(dA, dB, dD, vs[vi]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", vi, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0, zeros = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs((double)out[i] - ref[i]); if (d > maxerr) maxerr = d; bad += d > 1e-3; zeros += out[i] == 0.f; }
    printf("variant %d (layout %d lbo %u sbo %u adv %u major %d%d): maxerr %.3f bad %d/%d zeros %d  out[0..3]=%g %g %g %g ref=%g %g %g %g  out[row33]=%g ref=%g\n",
           vi, vs[vi].layout, vs[vi].lbo, vs[vi].sbo, vs[vi].adv, vs[vi].a_major, vs[vi].b_major, maxerr, bad, M * N, zeros,
           out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3], out[33 * N + 5], ref[33 * N + 5]);
  }
  // ---- does kind::tf32 truncate or round the fp32 operand?  A[0][0] = 1 + 2^-11 + 2^-12 (just above half a tf32 ulp)
  std::fill(A.begin(), A.end(), 0.f); std::fill(B.begin(), B.end(), 0.f);
  A[0] = 1.0f + 1.0f / 2048.0f + 1.0f / 4096.0f;  B[0] = 1.0f;
  A[1] = 1.0f + 1.0f / 4096.0f;                    // below half an ulp: both modes give 1
  A[2] = -(1.0f + 1.0f / 2048.0f + 1.0f / 4096.0f);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  probe_kernel<<<1, 160, smem>>>(dA, dB, dD, vs[1]);
  cudaDeviceSynchronize();
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  printf("rounding probe: D[0][0]=%.10f (trunc -> 1.0, round-nearest -> %.10f)  D[1][0]=%.10f  D[2][0]=%.10f\n",
         out[0], 1.0 + 1.0 / 1024.0, out[1 * N], out[2 * N]);
  return 0;
}
