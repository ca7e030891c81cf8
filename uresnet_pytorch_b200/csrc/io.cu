// Input / output layer feature movement (scn.InputLayer mode-3 sums and scn.OutputLayer,
// /root/reference/uresnet/models/uresnet_sparse.py:20,24) and batch collation.  Pure HBM-bound row copies:
// one thread per (row, column), consecutive threads on
// consecutive columns.
#include "common.cuh"

namespace b200scn {
namespace {

__global__ void k_gather_rows(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ idx,
                              int64_t n_cap, const int32_t* __restrict__ n_dev, int C, float* __restrict__ dst,
                              int64_t ldd) {
  const int64_t n = live_rows(n_dev, n_cap);
  const int64_t total = n * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / C;
    const int c = (int)(e - r * C);
    dst[r * ldd + c] = __ldg(src + (int64_t)__ldg(idx + r) * lds + c);
  }
}

__global__ void k_segsum_rows(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ first_idx,
                              const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg, int64_t n_cap,
                              const int32_t* __restrict__ n_dev, int C, float* __restrict__ dst, int64_t ldd) {
  const int64_t n = live_rows(n_dev, n_cap);
  const int64_t total = n * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / C;
    const int c = (int)(e - r * C);
    float v = __ldg(src + (int64_t)__ldg(first_idx + r) * lds + c);
    const int b = __ldg(seg_off + r), en = __ldg(seg_off + r + 1);
    for (int j = b; j < en; ++j) v += __ldg(src + (int64_t)__ldg(seg + j) * lds + c);
    dst[r * ldd + c] = v;
  }
}

// Collation (/root/reference/uresnet/iotools/iotools_sparse.py:131-141 pads every event's
// voxels with its batch id on the host and concatenates in float64): events arrive back to back as int32
// coordinates [N, D] plus event offsets; the kernel writes the [N, D + 1] matrix the input layer reads, the sample
// id found by a binary search over the (few) offsets.
__global__ void k_collate(const int32_t* __restrict__ coords, int64_t ldc, const int32_t* __restrict__ offsets, int nb,
                          int D, int first_id, int32_t* __restrict__ out, int64_t ldo, int32_t* __restrict__ n_out) {
  const int n = __ldg(offsets + nb);
  if (n_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *n_out = n;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = nb;  // last event whose offset is <= r
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(offsets + mid) <= r) lo = mid; else hi = mid;
    }
    for (int d = 0; d < D; ++d) out[r * ldo + d] = __ldg(coords + r * ldc + d);
    out[r * ldo + D] = first_id + lo;
  }
}

inline int ew_grid(int64_t elems) {
  int64_t g = cdiv(elems, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

extern "C" int b200scn_gather_rows(const float* src, int64_t lds, const int32_t* idx, int64_t n_cap,
                                   const int32_t* n_dev, int C, float* dst, int64_t ldd, void* stream) {
  B200_CHECK_ARG(C >= 1, "gather_rows: C must be >= 1");
  if (n_cap == 0) return 0;
  k_gather_rows<<<ew_grid(n_cap * C), 256, 0, (cudaStream_t)stream>>>(src, lds, idx, n_cap, n_dev, C, dst, ldd);
  count_launches(1);
  B200_CHECK_LAUNCH("gather_rows");
  return 0;
}

extern "C" int b200scn_segsum_rows(const float* src, int64_t lds, const int32_t* first_idx, const int32_t* seg_off,
                                   const int32_t* seg, int64_t n_cap, const int32_t* n_dev, int C, float* dst,
                                   int64_t ldd, void* stream) {
  B200_CHECK_ARG(C >= 1, "segsum_rows: C must be >= 1");
  if (n_cap == 0) return 0;
  k_segsum_rows<<<ew_grid(n_cap * C), 256, 0, (cudaStream_t)stream>>>(src, lds, first_idx, seg_off, seg, n_cap, n_dev,
                                                                      C, dst, ldd);
  count_launches(1);
  B200_CHECK_LAUNCH("segsum_rows");
  return 0;
}

extern "C" int b200scn_collate(const int32_t* coords, int64_t ldc, const int32_t* offsets_dev, int nb, int D,
                               int first_id, int64_t n_cap, int32_t* out, int64_t ldo, int32_t* n_out_dev,
                               void* stream) {
  B200_CHECK_ARG(nb >= 1 && D >= 1 && D <= 3 && ldo >= D + 1 && ldc >= D, "collate: bad shape");
  k_collate<<<ew_grid(n_cap < 1 ? 1 : n_cap), 256, 0, (cudaStream_t)stream>>>(coords, ldc, offsets_dev, nb, D, first_id,
                                                                            out, ldo, n_out_dev);
  count_launches(1);
  B200_CHECK_LAUNCH("collate");
  return 0;
}
