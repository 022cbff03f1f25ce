// common.cuh -- shared helpers for libb200sparse (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/b200sparse.h"

#ifndef B2S_NUM_SMS
#define B2S_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
#endif

// ------------------------------------------------------------------ error plumbing ----------
void b2s_set_error(const char* fmt, ...);

#define B2S_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      b2s_set_error("%s: invalid argument: %s", __func__, msg);    \
      return B2S_EINVAL;                                           \
    }                                                              \
  } while (0)

#define B2S_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      b2s_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(_e));           \
      return B2S_ECUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define B2S_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      b2s_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(_e));       \
      return B2S_ECUDA;                                                                      \
    }                                                                                        \
  } while (0)

// operand mode of the tensor-core convolutions (lib.cu): 1 = split-bf16 pairs (default), 0 = TF32
int b2s_precise();

static inline cudaStream_t as_stream(b2s_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride kernel: enough CTAs to cover the work, capped at a few waves of 148 SMs
static inline int grid_for(int64_t work_items, int block, int ctas_per_sm = 8) {
  int64_t need = ceil_div64(work_items, block);
  int64_t cap = (int64_t)B2S_NUM_SMS * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ------------------------------------------------------------------ device-side row counts --
// Every entry point that takes a row count n also takes `const int32_t* n_dev`.  NULL: n is exact.  Otherwise n is
// the CAPACITY (array pitch, launch bound) and the kernel reads the actual count from device memory, clamped to
// [0, n].  Launch shapes then do not depend on the data, so a whole training step can be captured in a CUDA graph.
__device__ __forceinline__ int64_t b2s_rows(int64_t n, const int* __restrict__ n_dev) {
  if (n_dev) {
    const int64_t v = (int64_t)__ldg(n_dev);
    n = v < 0 ? 0 : (v < n ? v : n);
  }
  return n;
}

// ------------------------------------------------------------------ coordinate keys ---------
// 64-bit key: batch | z | y | x, 16 bits each, spatial fields biased by 2^15 so that kernel
// offsets may reach below zero.  Key order == lexicographic (batch, z, y, x).
#define B2S_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define B2S_COORD_BIAS 32768
#define B2S_COORD_LIMIT 32000  // |coordinate| must stay below this so that +-K*step never wraps

__host__ __device__ __forceinline__ uint64_t b2s_pack_key(int b, int x, int y, int z) {
  return ((uint64_t)(uint16_t)b << 48) | ((uint64_t)(uint16_t)(z + B2S_COORD_BIAS) << 32) |
         ((uint64_t)(uint16_t)(y + B2S_COORD_BIAS) << 16) | (uint64_t)(uint16_t)(x + B2S_COORD_BIAS);
}

__host__ __device__ __forceinline__ uint64_t b2s_hash64(uint64_t k) {  // murmur3 fmix64
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// floor division toward -inf by a positive divisor, then back to a multiple of it
__host__ __device__ __forceinline__ int b2s_floor_to(int c, int ts) {
  if (ts == 1) return c;
  int q = c / ts;
  if ((c % ts != 0) && (c < 0)) --q;
  return q * ts;
}

// 16-byte hash entry
struct __align__(16) B2sEntry {
  unsigned long long key;
  int val;
  int pad;
};

// device-side lookup: returns row or -1
__device__ __forceinline__ int b2s_table_find(const B2sEntry* __restrict__ table, uint64_t mask, uint64_t key) {
  uint64_t s = b2s_hash64(key) & mask;
#pragma unroll 1
  for (;;) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(table + s));
    const uint64_t k = ((uint64_t)e.y << 32) | e.x;
    if (k == key) return (int)e.z;
    if (k == B2S_KEY_EMPTY) return -1;
    s = (s + 1) & mask;
  }
}
