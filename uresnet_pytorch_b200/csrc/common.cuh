// Shared helpers of libb200scn (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "b200scn.h"

namespace b200scn {

void set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      b200scn::set_error(__VA_ARGS__); \
      return B200SCN_ERR_ARG;          \
    }                                  \
  } while (0)

#define B200_CHECK_LAUNCH(name)                                                  \
  do {                                                                           \
    cudaError_t e__ = cudaPeekAtLastError();                                     \
    if (e__ != cudaSuccess) {                                                    \
      b200scn::set_error("%s: CUDA error: %s", name, cudaGetErrorString(e__));   \
      return B200SCN_ERR_CUDA;                                                   \
    }                                                                            \
  } while (0)

#define B200_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      b200scn::set_error("%s: CUDA error: %s", #call, cudaGetErrorString(e__));     \
      return B200SCN_ERR_CUDA;                                                      \
    }                                                                               \
  } while (0)

int sm_count();
void count_launches(int n);

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// live row count: min(*n_dev, n_cap) when a device count is given
__device__ __forceinline__ int live_rows(const int32_t* n_dev, int64_t n_cap) {
  int n = (int)n_cap;
  if (n_dev != nullptr) {
    int v = __ldg(n_dev);
    n = v < n ? (v < 0 ? 0 : v) : n;
  }
  return n;
}

constexpr uint64_t kEmptyKey = ~0ull;

__device__ __forceinline__ uint64_t mix64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// probe an open-addressing table (linear probing); returns the value or -1
__device__ __forceinline__ int hash_find(const uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tvals,
                                         uint64_t mask, uint64_t key) {
  uint64_t s = mix64(key) & mask;
  while (true) {
    uint64_t k = __ldg(tkeys + s);
    if (k == key) return __ldg(tvals + s);
    if (k == kEmptyKey) return -1;
    s = (s + 1) & mask;
  }
}

}  // namespace b200scn
