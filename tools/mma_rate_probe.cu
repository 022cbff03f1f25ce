// mma_rate_probe.cu -- issue-rate probe for tcgen05.mma (kind::f16) and tcgen05.commit on one SM: one thread issues
// ROUNDS x { K MMAs ; C commits } (compile-time unrolled, constant operands) and the CTA measures clock64 from the first
// issue to the completion of a final commit.  Operand values are irrelevant (whatever shared / tensor memory holds).
// Debug tool, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/mma_rate_probe tools/mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../dpcr_agb_b200/csrc/tc_ptx.cuh"
using namespace tc;

constexpr int ROUNDS = 2000;

template <int M, int N, int K, int COMMITS, int A_TMEM, int ELECT>
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]), barf = smem_u32(&bars[2]);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    mbar_init(barf, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0x3f803f80u;
  fence_proxy_async();
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
  if (ELECT ? (warp == 0 && elect_one()) : (tid == 0)) {
    constexpr uint32_t idesc = idesc_bf16(M, N, 0, 0);
    const uint64_t a_desc = smem_desc_sw128(base, 16, 1024);
    const uint64_t b_desc = smem_desc_sw128(base + 32 * 1024, 16, 1024);
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (A_TMEM) mma_bf16_ta(tmem, tmem + 256 + (uint32_t)((j & 3) * 8), b_desc + (uint64_t)((j & 3) * 2), idesc, 1u);
        else mma_bf16(tmem, a_desc + (uint64_t)((j & 3) * 2), b_desc + (uint64_t)((j & 3) * 2), idesc, 1u);
      }
#pragma unroll
      for (int j = 0; j < COMMITS; ++j) mma_commit(j ? bar1 : bar0);
    }
    const long long t1 = clock64();
    mma_commit(barf);
    mbar_wait(barf, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int M, int N, int K, int COMMITS, int A_TMEM, int ELECT>
void run(long long* out) {
  auto kern = probe<M, N, K, COMMITS, A_TMEM, ELECT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  kern<<<1, 128, 100 * 1024>>>(out);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
    exit(1);
  }
  const double ideal = (double)K * N / 2.0 * (M == 64 ? 1.0 : 1.0);   // M128 N K16 at 8192 dense bf16 flop / clk
  printf("M=%3d N=%3d  %2d MMAs + %d commits per round, A from %s, issued by %s: %8.1f clk / round = %6.1f clk / MMA "
         "(issue loop %7.1f / round), M=128 tensor time %6.1f\n",
         M, N, K, COMMITS, A_TMEM ? "tmem" : "smem", ELECT ? "elect.sync lane" : "tid == 0    ", (double)out[1] / ROUNDS,
         K ? (double)out[1] / ROUNDS / K : 0.0, (double)out[0] / ROUNDS, ideal);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 2 * sizeof(long long));
  run<128, 64, 6, 0, 0, 0>(out);
  run<128, 64, 6, 1, 0, 0>(out);
  run<128, 64, 6, 0, 0, 1>(out);
  run<128, 64, 6, 1, 0, 1>(out);
  run<128, 64, 12, 2, 1, 0>(out);
  run<128, 64, 12, 2, 1, 1>(out);
  run<128, 64, 24, 0, 0, 0>(out);
  run<128, 64, 24, 0, 1, 0>(out);
  run<128, 128, 8, 0, 0, 0>(out);
  run<128, 128, 8, 1, 0, 0>(out);
  run<128, 128, 8, 0, 1, 0>(out);
  run<128, 192, 8, 0, 1, 0>(out);
  run<128, 192, 8, 2, 1, 0>(out);
  run<128, 256, 8, 0, 0, 0>(out);
  run<128, 256, 8, 1, 0, 0>(out);
  run<128, 256, 8, 0, 1, 0>(out);
  run<64, 256, 8, 0, 0, 0>(out);
  run<64, 128, 8, 0, 0, 0>(out);
  run<64, 64, 8, 0, 0, 0>(out);
  run<128, 16, 8, 0, 0, 0>(out);
  run<128, 32, 8, 0, 0, 0>(out);
  run<128, 64, 0, 1, 0, 0>(out);
  run<128, 64, 0, 2, 0, 0>(out);
  run<128, 64, 1, 1, 0, 0>(out);
  return 0;
}
