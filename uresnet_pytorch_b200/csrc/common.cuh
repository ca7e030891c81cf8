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

// Programmatic dependent launch: the grid may be scheduled while its stream predecessor drains (the launch
// latency of the ~400 short kernels of a step overlaps the previous kernel's tail); every kernel launched
// this way calls pdl_wait() before it touches global memory, and none triggers early, so a completed grid
// still implies that everything before it in the stream completed.  B200SCN_PDL=0 turns it off.
bool pdl_enabled();

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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

// Implicit rule "table" of a dense grid: maps (output voxel row, kernel tap) to the input voxel
// row of a channels-last [B, (D,) H, W, C] tensor, or -1.  mode 1: strided correlation
// (in = out * stride + tap - pad), mode 2: its transpose (in = (out + pad - tap) / stride).
struct DenseGeom {
  int mode, dims;
  int in_sz[3], out_sz[3];  // (z, y, x); z = 1 for 2D
  int ks, stride, pad;
};

__device__ __forceinline__ int dense_index(const DenseGeom& gm, int o, int k) {
  int oc[3], kc[3];
  oc[2] = o % gm.out_sz[2]; o /= gm.out_sz[2];
  oc[1] = o % gm.out_sz[1]; o /= gm.out_sz[1];
  oc[0] = o % gm.out_sz[0]; o /= gm.out_sz[0];
  kc[2] = k % gm.ks; k /= gm.ks;
  kc[1] = k % gm.ks; k /= gm.ks;
  kc[0] = k;  // 0 for 2D
  int lin = o;  // batch index
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    int ic;
    if (gm.dims == 2 && d == 0) {
      ic = 0;
    } else if (gm.mode == 1) {
      ic = oc[d] * gm.stride + kc[d] - gm.pad;
      if (ic < 0 || ic >= gm.in_sz[d]) return -1;
    } else {
      const int tnum = oc[d] + gm.pad - kc[d];
      if (tnum < 0 || tnum % gm.stride != 0) return -1;
      ic = tnum / gm.stride;
      if (ic >= gm.in_sz[d]) return -1;
    }
    lin = lin * gm.in_sz[d] + ic;
  }
  return lin;
}

}  // namespace b200scn
