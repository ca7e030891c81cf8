// GPU hash-table rulebook builder (north-star job 1).
//
// Replaces scn's single-threaded CPU Metadata (InputLayer rules, SubmanifoldConvolution
// rules, Convolution/Deconvolution rules -- the scn calls made at
// /root/reference/uresnet/models/uresnet_sparse.py:20-22).  Design:
//   * one open-addressing table (64-bit packed key -> int32 row) per U-Net level, linear
//     probing, load factor <= 0.5, built with atomicCAS on the key and atomicMin on the
//     input index so that the *first occurrence* of every key is known without any ordering
//     assumption between threads;
//   * row ids = exclusive scan over the first-occurrence flags in input order, which
//     reproduces the sequential "hand out ids in order of discovery" semantics exactly and
//     deterministically;
//   * every kernel takes the live row count from device memory (n_dev), so the whole
//     pyramid is built without a host round-trip and can sit inside a CUDA graph;
//   * tables come out offset-major (table[k][row]) so the conv kernels read them coalesced.
// All of this is HBM/L2-bound integer work: coalesced streaming of keys, random 8-byte
// probes into the table.
#include "common.cuh"

namespace b200scn {
namespace {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanBlock * kScanItems;  // 1024

__device__ __forceinline__ int64_t load_coord(const void* p, int dtype, int64_t idx) {
  switch (dtype) {
    case B200SCN_I32: return (int64_t) reinterpret_cast<const int32_t*>(p)[idx];
    case B200SCN_I64: return reinterpret_cast<const int64_t*>(p)[idx];
    case B200SCN_F32: return (int64_t) reinterpret_cast<const float*>(p)[idx];
    default: return (int64_t) reinterpret_cast<const double*>(p)[idx];
  }
}

__device__ __forceinline__ int hash_insert_min(uint64_t* tkeys, int32_t* tvals, uint64_t mask, uint64_t key,
                                               int i) {
  uint64_t s = mix64(key) & mask;
  while (true) {
    unsigned long long prev =
        atomicCAS(reinterpret_cast<unsigned long long*>(tkeys + s), (unsigned long long)kEmptyKey,
                  (unsigned long long)key);
    if (prev == kEmptyKey || prev == key) {
      atomicMin(reinterpret_cast<unsigned int*>(tvals + s), (unsigned int)i);
      return (int)s;
    }
    s = (s + 1) & mask;
  }
}

// ---- insertion kernels ------------------------------------------------------------------
__global__ void k_insert_coords(const void* __restrict__ coords, int dtype, int64_t ldc, int has_sample, int D,
                                int64_t n_cap, const int32_t* __restrict__ n_dev, uint64_t* tkeys, int32_t* tvals,
                                uint64_t mask, uint64_t* __restrict__ keys, int32_t* __restrict__ slot_of,
                                int32_t* err_dev) {
  const int n = live_rows(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t key = 0;
    bool bad = false;
    for (int d = 0; d < D; ++d) {
      int64_t c = load_coord(coords, dtype, (int64_t)i * ldc + d);
      if (c < 0 || c >= 65535) { bad = true; c = 0; }
      key = (key << 16) | (uint64_t)c;
    }
    int64_t b = has_sample ? load_coord(coords, dtype, (int64_t)i * ldc + D) : 0;
    if (b < 0 || b >= 65535) { bad = true; b = 0; }
    key |= (uint64_t)b << 48;
    if (bad) atomicExch(err_dev, 1);
    keys[i] = key;
    slot_of[i] = hash_insert_min(tkeys, tvals, mask, key, i);
  }
}

__global__ void k_insert_coarse(const uint64_t* __restrict__ ukeys, int D, int s, int64_t n_cap,
                                const int32_t* __restrict__ n_dev, uint64_t* tkeys, int32_t* tvals, uint64_t mask,
                                uint64_t* __restrict__ ckeys, int32_t* __restrict__ slot_of,
                                uint8_t* __restrict__ kofs) {
  const int n = live_rows(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t key = ukeys[i];
    uint64_t ck = key & (0xFFFFull << 48);
    int f = 0;
    for (int d = 0; d < D; ++d) {
      const int sh = 16 * (D - 1 - d);
      const unsigned c = (unsigned)((key >> sh) & 0xFFFF);
      ck |= (uint64_t)(c / s) << sh;
      f = f * s + (int)(c % s);
    }
    ckeys[i] = ck;
    kofs[i] = (uint8_t)f;
    slot_of[i] = hash_insert_min(tkeys, tvals, mask, ck, i);
  }
}

__global__ void k_flag_first(const int32_t* __restrict__ tvals, const int32_t* __restrict__ slot_of, int64_t n_cap,
                             const int32_t* __restrict__ n_dev, int32_t* __restrict__ flags) {
  const int n = live_rows(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    flags[i] = (tvals[slot_of[i]] == i) ? 1 : 0;
}

// ---- int32 exclusive scan, three small kernels ---------------------------------------------
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  // blockDim.x == kScanBlock; returns exclusive prefix of v over the block
  __shared__ int warp_sums[kScanBlock / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < kScanBlock / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < kScanBlock / 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < kScanBlock / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  const int base = w > 0 ? warp_sums[w - 1] : 0;
  *total = warp_sums[kScanBlock / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(kScanBlock) k_scan_blocksums(const int32_t* __restrict__ in, int64_t n_cap,
                                                               const int32_t* __restrict__ n_dev,
                                                               int32_t* __restrict__ bsums) {
  const int n = live_rows(n_dev, n_cap);
  const int base = blockIdx.x * kScanTile;
  if (base >= n) return;
  int v = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int i = base + threadIdx.x * kScanItems + j;
    if (i < n) v += in[i];
  }
  int total;
  block_exclusive_scan(v, &total);
  if (threadIdx.x == 0) bsums[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums in place; total -> *total_out and end_out[n]
__global__ void __launch_bounds__(kScanBlock) k_scan_sums(int32_t* bsums, int64_t n_cap,
                                                          const int32_t* __restrict__ n_dev, int32_t* total_out,
                                                          int32_t* end_out) {
  const int n = live_rows(n_dev, n_cap);
  const int nb = (n + kScanTile - 1) / kScanTile;
  int carry = 0;
  for (int base = 0; base < nb; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? bsums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total);
    if (i < nb) bsums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    if (total_out) *total_out = carry;
    if (end_out) end_out[n] = carry;
  }
}

__global__ void __launch_bounds__(kScanBlock) k_scan_apply(const int32_t* __restrict__ in, int64_t n_cap,
                                                           const int32_t* __restrict__ n_dev,
                                                           const int32_t* __restrict__ bsums,
                                                           int32_t* __restrict__ out) {
  const int n = live_rows(n_dev, n_cap);
  const int base = blockIdx.x * kScanTile;
  if (base >= n) return;
  int item[kScanItems];
  int v = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int i = base + threadIdx.x * kScanItems + j;
    item[j] = i < n ? in[i] : 0;
    v += item[j];
  }
  int total;
  int ex = block_exclusive_scan(v, &total) + bsums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int i = base + threadIdx.x * kScanItems + j;
    if (i < n) out[i] = ex;
    ex += item[j];
  }
}

int scan_i32(const int32_t* in, int32_t* out, int64_t n_cap, const int32_t* n_dev, int32_t* total_out,
             int32_t* end_out, int32_t* bsums, cudaStream_t st) {
  const int nb = (int)cdiv(n_cap, kScanTile);
  k_scan_blocksums<<<nb, kScanBlock, 0, st>>>(in, n_cap, n_dev, bsums);
  k_scan_sums<<<1, kScanBlock, 0, st>>>(bsums, n_cap, n_dev, total_out, end_out);
  k_scan_apply<<<nb, kScanBlock, 0, st>>>(in, n_cap, n_dev, bsums, out);
  count_launches(3);
  B200_CHECK_LAUNCH("scan_i32");
  return 0;
}

// ---- row assignment -----------------------------------------------------------------------
__global__ void k_assign_rows(const int32_t* __restrict__ flags, const int32_t* __restrict__ ranks,
                              const int32_t* __restrict__ slot_of, const uint64_t* __restrict__ keys, int64_t n_cap,
                              const int32_t* __restrict__ n_dev, int32_t* tvals, uint64_t* __restrict__ ukeys,
                              int32_t* __restrict__ first_idx) {
  const int n = live_rows(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (flags[i]) {
      const int r = ranks[i];
      tvals[slot_of[i]] = r;
      ukeys[r] = keys[i];
      if (first_idx) first_idx[r] = i;
    }
  }
}

__global__ void k_row_of(const int32_t* __restrict__ tvals, const int32_t* __restrict__ slot_of, int64_t n_cap,
                         const int32_t* __restrict__ n_dev, int32_t* __restrict__ row_of) {
  const int n = live_rows(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    row_of[i] = tvals[slot_of[i]];
}

__global__ void k_parent_child(const int32_t* __restrict__ tvals, const int32_t* __restrict__ slot_of,
                               const uint8_t* __restrict__ kofs, int64_t n_cap, const int32_t* __restrict__ n_dev,
                               int32_t* __restrict__ parent, int32_t* __restrict__ child, int64_t ld_child) {
  const int n = live_rows(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int p = tvals[slot_of[i]];
    parent[i] = p;
    child[(int64_t)kofs[i] * ld_child + p] = i;
  }
}

// ---- duplicate CSR (mode-3 sums in input order) ---------------------------------------------
__global__ void k_dup_count(const int32_t* __restrict__ flags, const int32_t* __restrict__ row_of, int64_t n_cap,
                            const int32_t* __restrict__ n_dev, int32_t* cnt) {
  const int n = live_rows(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (!flags[i]) atomicAdd(cnt + row_of[i], 1);
}

__global__ void k_dup_fill(const int32_t* __restrict__ flags, const int32_t* __restrict__ row_of, int64_t n_cap,
                           const int32_t* __restrict__ n_dev, const int32_t* __restrict__ seg_off, int32_t* cur,
                           int32_t* __restrict__ seg) {
  const int n = live_rows(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (!flags[i]) {
      const int r = row_of[i];
      seg[seg_off[r] + atomicAdd(cur + r, 1)] = i;
    }
}

__global__ void k_dup_sort(const int32_t* __restrict__ seg_off, int64_t n_cap, const int32_t* __restrict__ n_act_dev,
                           int32_t* seg) {
  const int n = live_rows(n_act_dev, n_cap);
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int b = seg_off[r], e = seg_off[r + 1];
    for (int i = b + 1; i < e; ++i) {  // insertion sort; segments are tiny
      const int v = seg[i];
      int j = i - 1;
      while (j >= b && seg[j] > v) {
        seg[j + 1] = seg[j];
        --j;
      }
      seg[j + 1] = v;
    }
  }
}

// ---- submanifold neighbour table ------------------------------------------------------------
// grid.y = K/2 + 1: one probe per (row, offset k < K/2), mirrored entry written for free
__global__ void k_subm_table(const uint64_t* __restrict__ ukeys, int64_t n_cap, const int32_t* __restrict__ n_dev,
                             int D, int ks, int K, const uint64_t* __restrict__ tkeys,
                             const int32_t* __restrict__ tvals, uint64_t mask, int32_t* __restrict__ nbr,
                             int64_t ld) {
  const int n = live_rows(n_dev, n_cap);
  const int k = blockIdx.y;
  const int half = K / 2;
  int delta[3] = {0, 0, 0};
  {
    int rem = k;
    for (int d = D - 1; d >= 0; --d) {
      delta[d] = rem % ks - ks / 2;
      rem /= ks;
    }
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (k == half) {
      nbr[(int64_t)half * ld + i] = i;
      continue;
    }
    const uint64_t key = ukeys[i];
    uint64_t nk = key & (0xFFFFull << 48);
    bool ok = true;
    for (int d = 0; d < D; ++d) {
      const int sh = 16 * (D - 1 - d);
      const int c = (int)((key >> sh) & 0xFFFF) + delta[d];
      ok = ok && c >= 0 && c < 65535;
      nk |= (uint64_t)(unsigned)(c & 0xFFFF) << sh;
    }
    if (!ok) continue;
    const int j = hash_find(tkeys, tvals, mask, nk);
    if (j >= 0) {
      nbr[(int64_t)k * ld + i] = j;          // input row of output i at offset k
      nbr[(int64_t)(K - 1 - k) * ld + j] = i;  // and the mirrored rule
    }
  }
}

struct Scratch {
  int32_t *slot_of, *flags, *ranks, *bsums, *cnt, *cur;
  uint64_t* keys;
};

int64_t scratch_bytes(int64_t n_cap) {
  const int64_t nb = cdiv(n_cap, kScanTile) + 1;
  return 8 * (n_cap + 8) + 4 * (5 * (n_cap + 8) + nb + 8) + 1024;
}

int carve(void* scratch, int64_t bytes, int64_t n_cap, Scratch* s) {
  if (bytes < scratch_bytes(n_cap)) {
    set_error("rulebook scratch too small: %lld < %lld", (long long)bytes, (long long)scratch_bytes(n_cap));
    return B200SCN_ERR_ARG;
  }
  const int64_t nb = cdiv(n_cap, kScanTile) + 1;
  char* p = reinterpret_cast<char*>(scratch);
  s->keys = reinterpret_cast<uint64_t*>(p);
  p += 8 * (n_cap + 8);
  int32_t* q = reinterpret_cast<int32_t*>(p);
  s->slot_of = q; q += n_cap + 8;
  s->flags = q;   q += n_cap + 8;
  s->ranks = q;   q += n_cap + 8;
  s->cnt = q;     q += n_cap + 8;
  s->cur = q;     q += n_cap + 8;
  s->bsums = q;   q += nb + 8;
  return 0;
}

inline int grid_for(int64_t n, int block) {
  int64_t g = cdiv(n, block);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

int scan_i32_public(const int32_t* in, int32_t* out, int64_t n_cap, const int32_t* n_dev, int32_t* total_out,
                    int32_t* end_out, int32_t* bsums, cudaStream_t st) {
  return scan_i32(in, out, n_cap, n_dev, total_out, end_out, bsums, st);
}

}  // namespace b200scn

using namespace b200scn;

extern "C" int64_t b200scn_hash_slots(int64_t n_cap) {
  int64_t s = 1024;
  while (s < 2 * n_cap) s <<= 1;
  return s;
}

extern "C" int64_t b200scn_rulebook_scratch_bytes(int64_t n_cap) { return scratch_bytes(n_cap < 1 ? 1 : n_cap); }

extern "C" int b200scn_input_build(const void* coords, int coord_dtype, int64_t ld_coords, int has_sample_col, int D,
                                   int64_t n_cap, const int32_t* n_dev, uint64_t* tkeys, int32_t* tvals,
                                   int64_t slots, int32_t* in2row, uint64_t* ukeys, int32_t* n_act_dev,
                                   int32_t* first_idx, int32_t* seg_off, int32_t* seg, int32_t* err_dev,
                                   void* scratch, int64_t scratch_bytes_, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(D >= 1 && D <= 3, "input_build: D must be 1..3 (got %d)", D);
  B200_CHECK_ARG(coord_dtype >= 0 && coord_dtype <= 3, "input_build: bad coord dtype %d", coord_dtype);
  B200_CHECK_ARG(slots >= 2 * n_cap && (slots & (slots - 1)) == 0, "input_build: slots must be pow2 >= 2*n_cap");
  B200_CHECK_ARG(n_cap >= 0 && n_cap < (1ll << 30), "input_build: n_cap out of range");
  if (n_cap == 0) {
    B200_CUDA(cudaMemsetAsync(n_act_dev, 0, 4, st));
    B200_CUDA(cudaMemsetAsync(seg_off, 0, 4, st));
    return 0;
  }
  Scratch s;
  if (int rc = carve(scratch, scratch_bytes_, n_cap, &s)) return rc;
  B200_CUDA(cudaMemsetAsync(tkeys, 0xFF, (size_t)slots * 8, st));
  B200_CUDA(cudaMemsetAsync(tvals, 0xFF, (size_t)slots * 4, st));
  B200_CUDA(cudaMemsetAsync(s.cnt, 0, (size_t)(2 * (n_cap + 8)) * 4, st));  // cnt and cur are adjacent
  const int blk = 256, g = grid_for(n_cap, blk);
  k_insert_coords<<<g, blk, 0, st>>>(coords, coord_dtype, ld_coords, has_sample_col, D, n_cap, n_dev, tkeys, tvals,
                                     (uint64_t)slots - 1, s.keys, s.slot_of, err_dev);
  k_flag_first<<<g, blk, 0, st>>>(tvals, s.slot_of, n_cap, n_dev, s.flags);
  if (int rc = scan_i32(s.flags, s.ranks, n_cap, n_dev, n_act_dev, nullptr, s.bsums, st)) return rc;
  k_assign_rows<<<g, blk, 0, st>>>(s.flags, s.ranks, s.slot_of, s.keys, n_cap, n_dev, tvals, ukeys, first_idx);
  k_row_of<<<g, blk, 0, st>>>(tvals, s.slot_of, n_cap, n_dev, in2row);
  // duplicate CSR
  k_dup_count<<<g, blk, 0, st>>>(s.flags, in2row, n_cap, n_dev, s.cnt);
  if (int rc = scan_i32(s.cnt, seg_off, n_cap, n_act_dev, nullptr, seg_off, s.bsums, st)) return rc;
  k_dup_fill<<<g, blk, 0, st>>>(s.flags, in2row, n_cap, n_dev, seg_off, s.cur, seg);
  k_dup_sort<<<g, blk, 0, st>>>(seg_off, n_cap, n_act_dev, seg);
  count_launches(7);
  B200_CHECK_LAUNCH("input_build");
  return 0;
}

extern "C" int b200scn_level_down(const uint64_t* ukeys, int64_t n_cap, const int32_t* n_dev, int D, int sdiv,
                                  uint64_t* tkeys, int32_t* tvals, int64_t slots, int32_t* parent, uint8_t* kofs,
                                  uint64_t* ukeys_next, int32_t* n_next_dev, int32_t* child, int64_t ld_child,
                                  void* scratch, int64_t scratch_bytes_, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(D >= 1 && D <= 3, "level_down: D must be 1..3 (got %d)", D);
  B200_CHECK_ARG(sdiv >= 2 && sdiv <= 6, "level_down: stride must be 2..6 (got %d)", sdiv);
  B200_CHECK_ARG(slots >= 2 * n_cap && (slots & (slots - 1)) == 0, "level_down: slots must be pow2 >= 2*n_cap");
  B200_CHECK_ARG(ld_child >= n_cap, "level_down: ld_child < n_cap");
  if (n_cap == 0) {
    B200_CUDA(cudaMemsetAsync(n_next_dev, 0, 4, st));
    return 0;
  }
  int K = 1;
  for (int d = 0; d < D; ++d) K *= sdiv;
  Scratch s;
  if (int rc = carve(scratch, scratch_bytes_, n_cap, &s)) return rc;
  B200_CUDA(cudaMemsetAsync(tkeys, 0xFF, (size_t)slots * 8, st));
  B200_CUDA(cudaMemsetAsync(tvals, 0xFF, (size_t)slots * 4, st));
  B200_CUDA(cudaMemsetAsync(child, 0xFF, (size_t)K * ld_child * 4, st));
  const int blk = 256, g = grid_for(n_cap, blk);
  k_insert_coarse<<<g, blk, 0, st>>>(ukeys, D, sdiv, n_cap, n_dev, tkeys, tvals, (uint64_t)slots - 1, s.keys,
                                     s.slot_of, kofs);
  k_flag_first<<<g, blk, 0, st>>>(tvals, s.slot_of, n_cap, n_dev, s.flags);
  if (int rc = scan_i32(s.flags, s.ranks, n_cap, n_dev, n_next_dev, nullptr, s.bsums, st)) return rc;
  k_assign_rows<<<g, blk, 0, st>>>(s.flags, s.ranks, s.slot_of, s.keys, n_cap, n_dev, tvals, ukeys_next, nullptr);
  k_parent_child<<<g, blk, 0, st>>>(tvals, s.slot_of, kofs, n_cap, n_dev, parent, child, ld_child);
  count_launches(4);
  B200_CHECK_LAUNCH("level_down");
  return 0;
}

extern "C" int b200scn_subm_table(const uint64_t* ukeys, int64_t n_cap, const int32_t* n_dev, int D, int k,
                                  const uint64_t* tkeys, const int32_t* tvals, int64_t slots, int32_t* nbr,
                                  int64_t ld, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(D >= 1 && D <= 3, "subm_table: D must be 1..3 (got %d)", D);
  B200_CHECK_ARG(k >= 1 && (k & 1) && k <= 7, "subm_table: filter size must be odd and <= 7 (got %d)", k);
  B200_CHECK_ARG(ld >= n_cap, "subm_table: ld < n_cap");
  B200_CHECK_ARG((slots & (slots - 1)) == 0, "subm_table: slots must be a power of two");
  if (n_cap == 0) return 0;
  int K = 1;
  for (int d = 0; d < D; ++d) K *= k;
  B200_CUDA(cudaMemsetAsync(nbr, 0xFF, (size_t)K * ld * 4, st));
  const int blk = 256;
  dim3 grid(grid_for(n_cap, blk), K / 2 + 1);
  k_subm_table<<<grid, blk, 0, st>>>(ukeys, n_cap, n_dev, D, k, K, tkeys, tvals, (uint64_t)slots - 1, nbr, ld);
  count_launches(1);
  B200_CHECK_LAUNCH("subm_table");
  return 0;
}
