// Fused sparse BatchNorm + (leaky) ReLU (north-star job 3): scn.BatchNormReLU
// (/root/reference/uresnet/models/uresnet_sparse.py:23) and the BatchNormLeakyReLU layers
// inside scn.UNet (:22).  Semantics per SURVEY.md 8(a) row a6 / Appendix A.7.
//
// Pure streaming over an [N, C] fp32 matrix.  Statistics are per-thread shifted sums merged with Chan's formula
// in a fixed order, so the result is deterministic and as accurate as a two-pass variance.  Thread layout:
// thread = (column c, row lane rl); consecutive threads read consecutive columns of a row -> coalesced for any C.
// Two regimes:
//   * small inputs (every level of a bs=1 512^3 event): ONE launch per direction -- statistics, a grid-wide
//     barrier on a caller-owned wrapping counter, then the apply phase on the same registers
//     (k_bn_fwd_fused / k_bn_bwd_fused, <= 64 co-resident CTAs);
//   * large inputs: a statistics pass writing per-CTA partials + an apply pass whose CTAs fold them (every CTA in
//     the same order; beyond 64 partials one CTA merges them first).
#include <stdlib.h>

#include "common.cuh"

namespace b200scn {
namespace {

// Rows per partial-sum CTA: ~8192 elements per CTA (32 per thread), so narrow layers (many rows)
// and wide layers (few rows) both produce a few dozen partials -- a short fold in the apply
// kernels and enough CTAs to fill the SMs.  B200SCN_BN_ELEMS overrides the element budget.
static int rows_per_block_host(int C) {
  static int elems = 0;
  if (elems == 0) {
    const char* e = getenv("B200SCN_BN_ELEMS");
    elems = e ? atoi(e) : 8192;
    if (elems < 256) elems = 256;
  }
  int r = elems / C;
  if (r < 16) r = 16;
  if (r > 2048) r = 2048;
  return r;
}

struct Layout {
  int threads, RL, blocks;
};

Layout make_layout(int64_t n_cap, int C) {
  Layout L;
  static int tpb = 0;  // threads per CTA (B200SCN_BN_THREADS, default 1024): short per-thread chains
  if (tpb == 0) {
    const char* e = getenv("B200SCN_BN_THREADS");
    tpb = e ? atoi(e) : 1024;
    if (tpb < 64) tpb = 64;
    if (tpb > 1024) tpb = 1024;
  }
  L.RL = C >= tpb ? 1 : tpb / C;
  L.threads = C * L.RL;
  int64_t b = cdiv(n_cap, rows_per_block_host(C));
  const int64_t cap = (int64_t)sm_count() * 2;  // one co-resident wave of 1024-thread CTAs (see apply_grid)
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  L.blocks = (int)b;
  return L;
}

__device__ __forceinline__ void chan_merge(float& na, float& ma, float& sa, float nb, float mb, float sb) {
  if (nb == 0.f) return;
  const float n = na + nb;
  const float d = mb - ma;
  ma = ma + d * (nb / n);
  sa = sa + sb + d * d * (na * nb / n);
  na = n;
}

// Merge the RL row lanes of every column held in sm[0..T), sm[T..2T), sm[2T..3T) (count, mean, M2)
// with a fixed-shape binary tree (deterministic); the result ends up in lane 0 = sm[c], ...
__device__ __forceinline__ void chan_tree_merge(float* sm, int T, int C, int RL, int c, int rl) {
  int stride = 1;
  while (stride < RL) stride <<= 1;
  for (stride >>= 1; stride >= 1; stride >>= 1) {
    __syncthreads();
    if (rl < stride && rl + stride < RL) {
      const int a = rl * C + c, b = (rl + stride) * C + c;
      float n = sm[a], m = sm[T + a], s = sm[2 * T + a];
      chan_merge(n, m, s, sm[b], sm[T + b], sm[2 * T + b]);
      sm[a] = n;
      sm[T + a] = m;
      sm[2 * T + a] = s;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int live_blocks(int n, int grid_stats, int rows_per_blk) {
  return max(1, min(grid_stats, (n + rows_per_blk - 1) / rows_per_blk));
}

// Pass 1 of the training forward: per-CTA (count, mean, M2) per column.  part layout [blocks][3][C].
// No cross-CTA step here: every CTA of the apply kernel folds the partials itself (same order
// everywhere, so all CTAs agree bit for bit) -- there is no serial "last CTA" phase.
__global__ void k_bn_stats(const float* __restrict__ x, int64_t ldx, int64_t n_cap, const int32_t* __restrict__ n_dev,
                           int C, int RL, int rows_per_blk, float* __restrict__ part) {
  extern __shared__ float sm[];  // [3][threads]
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x, T = blockDim.x;
  const int c = t % C, rl = t / C;
  const int nblk = live_blocks(n, (int)gridDim.x, rows_per_blk);
  if ((int)blockIdx.x >= nblk) return;
  const int rpb = (n + nblk - 1) / nblk;
  const int r0 = blockIdx.x * rpb, r1 = min(n, r0 + rpb);

  float cnt = 0.f, s1 = 0.f, s2 = 0.f, shift = 0.f;
  if (r0 + rl < r1) shift = __ldg(x + (int64_t)(r0 + rl) * ldx + c);
  int r = r0 + rl;
  for (; r + 7 * RL < r1; r += 8 * RL) {  // eight loads in flight, sums in row order as before
    float xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = __ldg(x + (int64_t)(r + u * RL) * ldx + c);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float d = xv[u] - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
      cnt += 1.f;
    }
  }
  for (; r < r1; r += RL) {
    const float d = __ldg(x + (int64_t)r * ldx + c) - shift;
    s1 += d;
    s2 = fmaf(d, d, s2);
    cnt += 1.f;
  }
  float mean = 0.f, m2 = 0.f;
  if (cnt > 0.f) {
    mean = shift + s1 / cnt;
    m2 = fmaxf(s2 - s1 * s1 / cnt, 0.f);
  }
  sm[t] = cnt;
  sm[T + t] = mean;
  sm[2 * T + t] = m2;
  chan_tree_merge(sm, T, C, RL, c, rl);
  if (rl == 0) {
    float* p = part + (int64_t)blockIdx.x * 3 * C;
    p[c] = sm[c];
    p[C + c] = sm[T + c];
    p[2 * C + c] = sm[2 * T + c];
  }
}

// Pass 2: fold the partials (row lane rl takes CTAs rl, rl+RL, ...; lanes merged in order), then
// y = act((x - mean) * invstd * gamma + beta).  CTA 0 also publishes the statistics.
__global__ void k_bn_apply_train(const float* __restrict__ x, int64_t ldx, int64_t n_cap,
                                 const int32_t* __restrict__ n_dev, int C, int RL, int rows_per_blk, int grid_stats,
                                 const float* __restrict__ part, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* running_mean, float* running_var, float eps,
                                 float retention, float leak, float* __restrict__ y, int64_t ldy,
                                 float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  extern __shared__ float sm[];  // [3][threads]
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x, T = blockDim.x;
  const int c = t % C, rl = t / C;
  // grids are sized by the row capacity: CTAs without a live row leave before the fold (CTA 0 publishes)
  if (blockIdx.x != 0 && (int64_t)blockIdx.x * RL >= n) return;
  const int nblk = live_blocks(n, grid_stats, rows_per_blk);
  {
    // fold this lane's share of the partial blocks: four blocks' loads in flight per step (the merge is a
    // serial chain, the L2 round trips need not be), merged in block order
    float N = 0.f, M = 0.f, S = 0.f;
    int b = rl;
    for (; b + 3 * RL < nblk; b += 4 * RL) {
      float pn[4], pm[4], ps[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* p = part + (int64_t)(b + u * RL) * 3 * C;
        pn[u] = __ldcg(p + c); pm[u] = __ldcg(p + C + c); ps[u] = __ldcg(p + 2 * C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) chan_merge(N, M, S, pn[u], pm[u], ps[u]);
    }
    for (; b < nblk; b += RL) {
      const float* p = part + (int64_t)b * 3 * C;
      chan_merge(N, M, S, __ldcg(p + c), __ldcg(p + C + c), __ldcg(p + 2 * C + c));
    }
    sm[t] = N;
    sm[T + t] = M;
    sm[2 * T + t] = S;
  }
  chan_tree_merge(sm, T, C, RL, c, rl);
  const float N = sm[c], M = sm[T + c], S = sm[2 * T + c];
  const float is = rsqrtf((N > 0.f ? S / N : 0.f) + eps);
  if (blockIdx.x == 0 && rl == 0) {
    save_mean[c] = M;
    save_invstd[c] = is;
    if (running_mean != nullptr) {
      running_mean[c] = retention * running_mean[c] + (1.f - retention) * M;
      running_var[c] = retention * running_var[c] + (1.f - retention) * (S / fmaxf(N - 1.f, 1.f));
    }
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float scale = is * g;
  const float shift = b - M * scale;
  // four rows' loads in flight per thread: with one load per trip the store behind it serialises the trips and
  // the kernel runs at a quarter of the memory bandwidth (measured on 300 k-row levels: 1.6 TB/s)
  const int step = gridDim.x * RL;
  int r = blockIdx.x * RL + rl;
  for (; r + 3 * step < n; r += 4 * step) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(x + (int64_t)(r + u * step) * ldx + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float o = fmaf(v[u], scale, shift);
      o = o > 0.f ? o : o * leak;
      y[(int64_t)(r + u * step) * ldy + c] = o;
    }
  }
  for (; r < n; r += step) {
    float v = fmaf(__ldg(x + (int64_t)r * ldx + c), scale, shift);
    v = v > 0.f ? v : v * leak;
    y[(int64_t)r * ldy + c] = v;
  }
}

__global__ void k_bn_apply(const float* __restrict__ x, int64_t ldx, int64_t n_cap, const int32_t* __restrict__ n_dev,
                           int C, int RL, const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ mean, const float* __restrict__ invstd_or_var, float eps,
                           int var_given, float leak, float* __restrict__ y, int64_t ldy) {
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x;
  const int c = t % C, rl = t / C;
  float is = invstd_or_var[c];
  if (var_given) is = rsqrtf(is + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float scale = is * g;
  const float shift = b - mean[c] * scale;
  for (int r = blockIdx.x * RL + rl; r < n; r += gridDim.x * RL) {
    float v = fmaf(__ldg(x + (int64_t)r * ldx + c), scale, shift);
    v = v > 0.f ? v : v * leak;
    y[(int64_t)r * ldy + c] = v;
  }
}

// Backward pass 1: per-CTA sums of dy_masked and dy_masked * xhat.  part layout [blocks][2][C]
__global__ void k_bn_bwd_reduce(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t lddy,
                                int64_t n_cap, const int32_t* __restrict__ n_dev, int C, int RL, int rows_per_blk,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ mean, const float* __restrict__ invstd, float leak,
                                float* __restrict__ part) {
  extern __shared__ float sm[];  // [2][threads]
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x, T = blockDim.x;
  const int c = t % C, rl = t / C;
  const int nblk = live_blocks(n, (int)gridDim.x, rows_per_blk);
  if ((int)blockIdx.x >= nblk) return;
  const int rpb = (n + nblk - 1) / nblk;
  const int r0 = blockIdx.x * rpb, r1 = min(n, r0 + rpb);
  const float mu = mean[c], is = invstd[c];
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  int r = r0 + rl;
  for (; r + 3 * RL < r1; r += 4 * RL) {  // eight loads in flight, sums in row order as before
    float xv[4], dv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      xv[u] = __ldg(x + (int64_t)(r + u * RL) * ldx + c);
      dv[u] = __ldg(dy + (int64_t)(r + u * RL) * lddy + c);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float xh = (xv[u] - mu) * is;
      const float yv = fmaf(xh, g, b);
      const float d = yv > 0.f ? dv[u] : dv[u] * leak;
      s1 += d;
      s2 = fmaf(d, xh, s2);
    }
  }
  for (; r < r1; r += RL) {
    const float xh = (__ldg(x + (int64_t)r * ldx + c) - mu) * is;
    const float yv = fmaf(xh, g, b);
    float d = __ldg(dy + (int64_t)r * lddy + c);
    d = yv > 0.f ? d : d * leak;
    s1 += d;
    s2 = fmaf(d, xh, s2);
  }
  sm[t] = s1;
  sm[T + t] = s2;
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < RL; ++j) {
      s1 += sm[j * C + c];
      s2 += sm[T + j * C + c];
    }
    float* p = part + (int64_t)blockIdx.x * 2 * C;
    p[c] = s1;
    p[C + c] = s2;
  }
}

// Backward pass 2: fold the partials (every CTA, same order), then dx.  CTA 0 writes dgamma / dbeta.
__global__ void k_bn_bwd_apply(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t lddy,
                               int64_t n_cap, const int32_t* __restrict__ n_dev, int C, int RL, int rows_per_blk,
                               int grid_stats, const float* __restrict__ part, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ invstd, float leak, int train, float* __restrict__ dgamma,
                               float* __restrict__ dbeta, float* __restrict__ dx, int64_t lddx) {
  extern __shared__ float sm[];  // [2][threads]
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x, T = blockDim.x;
  const int c = t % C, rl = t / C;
  if (blockIdx.x != 0 && (int64_t)blockIdx.x * RL >= n) return;
  const int nblk = live_blocks(n, grid_stats, rows_per_blk);
  {
    float a1 = 0.f, a2 = 0.f;
    int bk = rl;
    for (; bk + 3 * RL < nblk; bk += 4 * RL) {  // four blocks' loads in flight, summed in block order
      float v1[4], v2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* p = part + (int64_t)(bk + u * RL) * 2 * C;
        v1[u] = __ldcg(p + c); v2[u] = __ldcg(p + C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a1 += v1[u]; a2 += v2[u]; }
    }
    for (; bk < nblk; bk += RL) {
      const float* p = part + (int64_t)bk * 2 * C;
      a1 += __ldcg(p + c);
      a2 += __ldcg(p + C + c);
    }
    sm[t] = a1;
    sm[T + t] = a2;
  }
  __syncthreads();
  float a1 = sm[c], a2 = sm[T + c];
  for (int j = 1; j < RL; ++j) {
    a1 += sm[j * C + c];
    a2 += sm[T + j * C + c];
  }
  if (blockIdx.x == 0 && rl == 0) {
    if (train & 2) {  // accumulate into a persistent gradient buffer (functional.GradSink)
      dbeta[c] += a1;
      dgamma[c] += a2;
    } else {
      dbeta[c] = a1;
      dgamma[c] = a2;
    }
  }
  const float mu = mean[c], is = invstd[c];
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  const float k1 = (train & 1) ? a1 * inv_n : 0.f;
  const float k2 = (train & 1) ? a2 * inv_n : 0.f;
  const float gi = g * is;
  const int step = gridDim.x * RL;  // (four rows' loads in flight per thread, see k_bn_apply_train)
  int r = blockIdx.x * RL + rl;
  for (; r + 3 * step < n; r += 4 * step) {
    float xv[4], dv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      xv[u] = __ldg(x + (int64_t)(r + u * step) * ldx + c);
      dv[u] = __ldg(dy + (int64_t)(r + u * step) * lddy + c);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float xh = (xv[u] - mu) * is;
      const float yv = fmaf(xh, g, b);
      const float d = yv > 0.f ? dv[u] : dv[u] * leak;
      dx[(int64_t)(r + u * step) * lddx + c] = gi * (d - k1 - xh * k2);
    }
  }
  for (; r < n; r += step) {
    const float xh = (__ldg(x + (int64_t)r * ldx + c) - mu) * is;
    const float yv = fmaf(xh, g, b);
    float d = __ldg(dy + (int64_t)r * lddy + c);
    d = yv > 0.f ? d : d * leak;
    dx[(int64_t)r * lddx + c] = gi * (d - k1 - xh * k2);
  }
}

// Large inputs: hundreds of partial blocks.  Folding them inside every apply CTA is O(CTAs x blocks) L2
// traffic (270 MB per BatchNorm layer at 300 k rows x 32 channels -- 20x the feature matrix); one CTA folds
// them once instead and the apply kernels read a single merged block.
__global__ void k_bn_merge_stats(const float* __restrict__ part, int64_t n_cap, const int32_t* __restrict__ n_dev, int C,
                                 int rows_per_blk, int grid_stats, float* __restrict__ merged) {
  // one CTA per 32 channels, 32 row lanes each: a wide layer no longer folds its partial blocks on C / 1024 lanes
  // (5 lanes x 118 dependent merges at 192 channels: 21 us per launch at 768^3)
  extern __shared__ float sm[];  // [3][threads]
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x, T = blockDim.x;
  const int cl = t & 31, rl = t >> 5, RL = T >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool valid = c < C;
  const int nblk = live_blocks(n, grid_stats, rows_per_blk);
  float N = 0.f, M = 0.f, S = 0.f;
  if (valid) {
    int b = rl;
    for (; b + 3 * RL < nblk; b += 4 * RL) {
      float pn[4], pm[4], ps[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* p = part + (int64_t)(b + u * RL) * 3 * C;
        pn[u] = __ldcg(p + c); pm[u] = __ldcg(p + C + c); ps[u] = __ldcg(p + 2 * C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) chan_merge(N, M, S, pn[u], pm[u], ps[u]);
    }
    for (; b < nblk; b += RL) {
      const float* p = part + (int64_t)b * 3 * C;
      chan_merge(N, M, S, __ldcg(p + c), __ldcg(p + C + c), __ldcg(p + 2 * C + c));
    }
  }
  sm[t] = N; sm[T + t] = M; sm[2 * T + t] = S;
  chan_tree_merge(sm, T, 32, RL, cl, rl);
  if (rl == 0 && valid) { merged[c] = sm[cl]; merged[C + c] = sm[T + cl]; merged[2 * C + c] = sm[2 * T + cl]; }
}

__global__ void k_bn_merge_sums(const float* __restrict__ part, int64_t n_cap, const int32_t* __restrict__ n_dev, int C,
                                int rows_per_blk, int grid_stats, float* __restrict__ merged) {
  extern __shared__ float sm[];  // [2][threads]; one CTA per 32 channels, 32 row lanes each
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x, T = blockDim.x;
  const int cl = t & 31, rl = t >> 5, RL = T >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool valid = c < C;
  const int nblk = live_blocks(n, grid_stats, rows_per_blk);
  float a1 = 0.f, a2 = 0.f;
  if (valid) {
    for (int bk = rl; bk < nblk; bk += RL) {
      const float* p = part + (int64_t)bk * 2 * C;
      a1 += __ldcg(p + c);
      a2 += __ldcg(p + C + c);
    }
  }
  sm[t] = a1; sm[T + t] = a2;
  __syncthreads();
  if (rl == 0 && valid) {
    for (int j = 1; j < RL; ++j) { a1 += sm[j * 32 + cl]; a2 += sm[T + j * 32 + cl]; }
    merged[c] = a1; merged[C + c] = a2;
  }
}


// ---- single-launch variants for inputs that fit the SMs' registers (every level of a bs=1 512^3 event) ------
// One co-resident grid (<= one CTA per SM): statistics, a grid-wide barrier, then the apply phase on the SAME
// registers -- the feature matrix is read once and the second launch (5-9 us of pure latency per layer at these
// sizes) disappears.  The barrier state (one wrapping counter) lives in caller memory that is zero before the first
// use and is left zero-count by every launch; all CTAs are resident at once because the grid never exceeds the
// SM count, so the spin cannot starve a CTA that has not started.
constexpr int kFusedV = 8;  // values per thread cached in registers; rows beyond are re-read (L2 hits)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One wrapping counter: atomicInc(cnt, G - 1) returns to zero with the G-th arrival, which is both the release
// signal of this launch and the reset for the next one -- three L2 round trips per CTA (fence, atomic, one
// poll) instead of the eight of a count + generation pair (ncu: the barrier was most of the kernel's 11 us).
__device__ __forceinline__ void grid_barrier(unsigned int* sync, unsigned int G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int old = atomicInc(sync, G - 1u);
    if (old != G - 1u) {
      while (ld_acquire_u32(sync) != 0u) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024, 1)
k_bn_fwd_fused(const float* __restrict__ x, int64_t ldx, int64_t n_cap, const int32_t* __restrict__ n_dev, int C, int RL,
               float* __restrict__ part, unsigned int* sync, const float* __restrict__ gamma,
               const float* __restrict__ beta, float* running_mean, float* running_var, float eps, float retention,
               float leak, float* __restrict__ y, int64_t ldy, float* __restrict__ save_mean,
               float* __restrict__ save_invstd) {
  extern __shared__ float sm[];  // [3][threads]
  pdl_wait();
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x, T = blockDim.x;
  const int c = t % C, rl = t / C;
  const int G = (int)gridDim.x;
  const int rpb = (n + G - 1) / G;
  const int r0 = min(n, (int)blockIdx.x * rpb), r1 = min(n, r0 + rpb);

  float xv[kFusedV];
#pragma unroll
  for (int v = 0; v < kFusedV; ++v) {
    const int r = r0 + rl + v * RL;
    xv[v] = r < r1 ? __ldg(x + (int64_t)r * ldx + c) : 0.f;
  }
  float cnt = 0.f, s1 = 0.f, s2 = 0.f;
  const float shift0 = xv[0];
#pragma unroll
  for (int v = 0; v < kFusedV; ++v) {
    if (r0 + rl + v * RL < r1) {
      const float d = xv[v] - shift0;
      s1 += d;
      s2 = fmaf(d, d, s2);
      cnt += 1.f;
    }
  }
  for (int r = r0 + rl + kFusedV * RL; r < r1; r += RL) {
    const float d = __ldg(x + (int64_t)r * ldx + c) - shift0;
    s1 += d;
    s2 = fmaf(d, d, s2);
    cnt += 1.f;
  }
  float mean = 0.f, m2 = 0.f;
  if (cnt > 0.f) {
    mean = shift0 + s1 / cnt;
    m2 = fmaxf(s2 - s1 * s1 / cnt, 0.f);
  }
  sm[t] = cnt;
  sm[T + t] = mean;
  sm[2 * T + t] = m2;
  chan_tree_merge(sm, T, C, RL, c, rl);
  if (rl == 0) {
    float* p = part + (int64_t)blockIdx.x * 3 * C;
    p[c] = sm[c];
    p[C + c] = sm[T + c];
    p[2 * C + c] = sm[2 * T + c];
  }
  grid_barrier(sync, (unsigned int)G);
  {
    float N = 0.f, M = 0.f, S = 0.f;
    int b = rl;
    for (; b + 3 * RL < G; b += 4 * RL) {
      float pn[4], pm[4], ps[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* p = part + (int64_t)(b + u * RL) * 3 * C;
        pn[u] = __ldcg(p + c); pm[u] = __ldcg(p + C + c); ps[u] = __ldcg(p + 2 * C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) chan_merge(N, M, S, pn[u], pm[u], ps[u]);
    }
    for (; b < G; b += RL) {
      const float* p = part + (int64_t)b * 3 * C;
      chan_merge(N, M, S, __ldcg(p + c), __ldcg(p + C + c), __ldcg(p + 2 * C + c));
    }
    __syncthreads();  // chan_tree_merge above left readers of sm[] behind the barrier only; be explicit
    sm[t] = N;
    sm[T + t] = M;
    sm[2 * T + t] = S;
  }
  chan_tree_merge(sm, T, C, RL, c, rl);
  const float N = sm[c], M = sm[T + c], S = sm[2 * T + c];
  const float is = rsqrtf((N > 0.f ? S / N : 0.f) + eps);
  if (blockIdx.x == 0 && rl == 0) {
    save_mean[c] = M;
    save_invstd[c] = is;
    if (running_mean != nullptr) {
      running_mean[c] = retention * running_mean[c] + (1.f - retention) * M;
      running_var[c] = retention * running_var[c] + (1.f - retention) * (S / fmaxf(N - 1.f, 1.f));
    }
  }
  const float g = gamma ? gamma[c] : 1.f, bb = beta ? beta[c] : 0.f;
  const float scale = is * g;
  const float shift = bb - M * scale;
#pragma unroll
  for (int v = 0; v < kFusedV; ++v) {
    const int r = r0 + rl + v * RL;
    if (r < r1) {
      float o = fmaf(xv[v], scale, shift);
      o = o > 0.f ? o : o * leak;
      y[(int64_t)r * ldy + c] = o;
    }
  }
  for (int r = r0 + rl + kFusedV * RL; r < r1; r += RL) {
    float o = fmaf(__ldg(x + (int64_t)r * ldx + c), scale, shift);
    o = o > 0.f ? o : o * leak;
    y[(int64_t)r * ldy + c] = o;
  }
}

__global__ void __launch_bounds__(1024, 1)
k_bn_bwd_fused(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t lddy, int64_t n_cap,
               const int32_t* __restrict__ n_dev, int C, int RL, float* __restrict__ part, unsigned int* sync,
               const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
               const float* __restrict__ invstd, float leak, int train, float* __restrict__ dgamma,
               float* __restrict__ dbeta, float* __restrict__ dx, int64_t lddx) {
  extern __shared__ float sm[];  // [2][threads]
  pdl_wait();
  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x, T = blockDim.x;
  const int c = t % C, rl = t / C;
  const int G = (int)gridDim.x;
  const int rpb = (n + G - 1) / G;
  const int r0 = min(n, (int)blockIdx.x * rpb), r1 = min(n, r0 + rpb);
  const float mu = mean[c], is = invstd[c];
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;

  float xh[kFusedV], dv[kFusedV];
#pragma unroll
  for (int v = 0; v < kFusedV; ++v) {
    const int r = r0 + rl + v * RL;
    const bool in = r < r1;
    xh[v] = in ? __ldg(x + (int64_t)r * ldx + c) : 0.f;
    dv[v] = in ? __ldg(dy + (int64_t)r * lddy + c) : 0.f;
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int v = 0; v < kFusedV; ++v) {
    xh[v] = (xh[v] - mu) * is;
    const float yv = fmaf(xh[v], g, b);
    dv[v] = yv > 0.f ? dv[v] : dv[v] * leak;
    if (r0 + rl + v * RL < r1) {
      s1 += dv[v];
      s2 = fmaf(dv[v], xh[v], s2);
    }
  }
  for (int r = r0 + rl + kFusedV * RL; r < r1; r += RL) {
    const float h = (__ldg(x + (int64_t)r * ldx + c) - mu) * is;
    const float yv = fmaf(h, g, b);
    float d = __ldg(dy + (int64_t)r * lddy + c);
    d = yv > 0.f ? d : d * leak;
    s1 += d;
    s2 = fmaf(d, h, s2);
  }
  sm[t] = s1;
  sm[T + t] = s2;
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < RL; ++j) {
      s1 += sm[j * C + c];
      s2 += sm[T + j * C + c];
    }
    float* p = part + (int64_t)blockIdx.x * 2 * C;
    p[c] = s1;
    p[C + c] = s2;
  }
  grid_barrier(sync, (unsigned int)G);
  {
    float a1 = 0.f, a2 = 0.f;
    int bk = rl;
    for (; bk + 3 * RL < G; bk += 4 * RL) {
      float v1[4], v2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* p = part + (int64_t)(bk + u * RL) * 2 * C;
        v1[u] = __ldcg(p + c); v2[u] = __ldcg(p + C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a1 += v1[u]; a2 += v2[u]; }
    }
    for (; bk < G; bk += RL) {
      const float* p = part + (int64_t)bk * 2 * C;
      a1 += __ldcg(p + c);
      a2 += __ldcg(p + C + c);
    }
    sm[t] = a1;
    sm[T + t] = a2;
  }
  __syncthreads();
  float a1 = sm[c], a2 = sm[T + c];
  for (int j = 1; j < RL; ++j) {
    a1 += sm[j * C + c];
    a2 += sm[T + j * C + c];
  }
  if (blockIdx.x == 0 && rl == 0) {
    if (train & 2) {  // accumulate into a persistent gradient buffer (functional.GradSink)
      dbeta[c] += a1;
      dgamma[c] += a2;
    } else {
      dbeta[c] = a1;
      dgamma[c] = a2;
    }
  }
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  const float k1 = (train & 1) ? a1 * inv_n : 0.f;
  const float k2 = (train & 1) ? a2 * inv_n : 0.f;
  const float gi = g * is;
#pragma unroll
  for (int v = 0; v < kFusedV; ++v) {
    const int r = r0 + rl + v * RL;
    if (r < r1) dx[(int64_t)r * lddx + c] = gi * (dv[v] - k1 - xh[v] * k2);
  }
  for (int r = r0 + rl + kFusedV * RL; r < r1; r += RL) {
    const float h = (__ldg(x + (int64_t)r * ldx + c) - mu) * is;
    const float yv = fmaf(h, g, b);
    float d = __ldg(dy + (int64_t)r * lddy + c);
    d = yv > 0.f ? d : d * leak;
    dx[(int64_t)r * lddx + c] = gi * (d - k1 - h * k2);
  }
}

// Grid of the single-launch variants, or 0 when the input is too large for them (two-kernel path).
// n_expect: rows a typical event holds at this level (0 = unknown: use the capacity).
int fused_grid(int64_t n_cap, int64_t n_expect, int C, const Layout& L, bool bwd = false) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("B200SCN_BN_FUSED");
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled) return 0;
  int64_t rows = n_expect > 0 && n_expect < n_cap ? n_expect : n_cap;
  const int64_t elems = rows * C;
  const int sms = sm_count();
  // beyond 2x the register-cached capacity the re-reads cost more than the second launch
  if (elems > (int64_t)sms * L.threads * kFusedV * 2 || n_cap * C > (int64_t)sms * L.threads * kFusedV * 8) return 0;
  static int vpt = 0, maxg = 0;  // values per thread the grid is sized for; CTA cap (barrier and fold cost grow with it)
  if (vpt == 0) {
    const char* e = getenv("B200SCN_BN_FUSED_VPT");
    vpt = e ? atoi(e) : 8;
    if (vpt < 1) vpt = 1;
    e = getenv("B200SCN_BN_FUSED_MAXG");
    maxg = e ? atoi(e) : 64;  // measured: 64 CTAs beat 148 (profiles/README.md)
    if (maxg < 1 || maxg > sms) maxg = sms;
  }
  static int maxg_bwd = 0;  // B200SCN_BN_FUSED_MAXG_BWD: CTA cap of the backward kernel alone (A/B runs)
  if (maxg_bwd == 0) {
    const char* e = getenv("B200SCN_BN_FUSED_MAXG_BWD");
    maxg_bwd = e ? atoi(e) : 96;  // twice the data of the forward kernel (x and dy): c2 4.154 ms at 64, 4.134 at 32, 4.041 at 96, 4.048 at 128
    if (maxg_bwd < 1 || maxg_bwd > sms) maxg_bwd = sms;
  }
  int64_t g = cdiv(elems, (int64_t)L.threads * vpt);
  const int cap_g = bwd ? maxg_bwd : maxg;
  if (g > cap_g) g = cap_g;
  if (g < 1) g = 1;
  return (int)g;
}

constexpr int kMergeAbove = 64;  // partial blocks beyond which the merge runs as its own launch

int apply_grid(int64_t n_cap, int RL) {
  int64_t g = cdiv(n_cap, (int64_t)RL * 4);
  // one co-resident wave (two 1024-thread CTAs per SM): every CTA first folds / loads the statistics and merges
  // them through shared memory -- a few microseconds that four waves of CTAs used to pay four times over
  const int64_t cap = (int64_t)sm_count() * 2;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

extern "C" int64_t b200scn_bn_scratch_bytes(int64_t n_cap, int C) {
  Layout L = make_layout(n_cap < 1 ? 1 : n_cap, C < 1 ? 1 : C);
  const int64_t blocks = L.blocks + 1 > sm_count() ? L.blocks + 1 : sm_count();  // single-launch path: one per SM
  return 256 + blocks * 3 * C * 4;
}

extern "C" int b200scn_bn_fwd_train(const float* x, int64_t ldx, int64_t n_cap, const int32_t* n_dev, int C,
                                    const float* gamma, const float* beta, float* running_mean, float* running_var,
                                    float eps, float retention, float leak, float* y, int64_t ldy, float* save_mean,
                                    float* save_invstd, void* scratch, int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(C >= 1 && C <= 1024, "bn: C must be 1..1024 (got %d)", C);
  B200_CHECK_ARG(scratch_bytes >= b200scn_bn_scratch_bytes(n_cap, C), "bn_fwd_train: scratch too small");
  if (n_cap == 0) return 0;
  Layout L = make_layout(n_cap, C);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 256);
  const int rpb = rows_per_block_host(C);
  k_bn_stats<<<L.blocks, L.threads, 3 * L.threads * sizeof(float), st>>>(x, ldx, n_cap, n_dev, C, L.RL, rpb, part);
  const float* part_apply = part;
  int blocks_apply = L.blocks, rpb_apply = rpb;
  if (L.blocks > kMergeAbove) {
    float* merged = part + (int64_t)L.blocks * 3 * C;
    k_bn_merge_stats<<<(C + 31) / 32, 1024, 3 * 1024 * sizeof(float), st>>>(part, n_cap, n_dev, C, rpb, L.blocks, merged);
    part_apply = merged; blocks_apply = 1; rpb_apply = 1 << 30;
    count_launches(1);
  }
  k_bn_apply_train<<<apply_grid(n_cap, L.RL), L.threads, 3 * L.threads * sizeof(float), st>>>(
      x, ldx, n_cap, n_dev, C, L.RL, rpb_apply, blocks_apply, part_apply, gamma, beta, running_mean, running_var, eps,
      retention, leak, y, ldy, save_mean, save_invstd);
  count_launches(2);
  B200_CHECK_LAUNCH("bn_fwd_train");
  return 0;
}

extern "C" int b200scn_bn_fwd_eval(const float* x, int64_t ldx, int64_t n_cap, const int32_t* n_dev, int C,
                                   const float* gamma, const float* beta, const float* running_mean,
                                   const float* running_var, float eps, float leak, float* y, int64_t ldy,
                                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(C >= 1 && C <= 1024, "bn: C must be 1..1024 (got %d)", C);
  if (n_cap == 0) return 0;
  Layout L = make_layout(n_cap, C);
  k_bn_apply<<<apply_grid(n_cap, L.RL), L.threads, 0, st>>>(x, ldx, n_cap, n_dev, C, L.RL, gamma, beta, running_mean,
                                                            running_var, eps, 1, leak, y, ldy);
  count_launches(1);
  B200_CHECK_LAUNCH("bn_fwd_eval");
  return 0;
}

extern "C" int b200scn_bn_bwd(const float* x, int64_t ldx, const float* dy, int64_t lddy, int64_t n_cap,
                              const int32_t* n_dev, int C, const float* gamma, const float* beta, const float* mean,
                              const float* invstd, float leak, int train, float* dx, int64_t lddx, float* dgamma,
                              float* dbeta, void* scratch, int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(C >= 1 && C <= 1024, "bn: C must be 1..1024 (got %d)", C);
  B200_CHECK_ARG(scratch_bytes >= b200scn_bn_scratch_bytes(n_cap, C), "bn_bwd: scratch too small");
  if (n_cap == 0) {
    if (!(train & 2)) {
      B200_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, st));
      B200_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, st));
    }
    return 0;
  }
  Layout L = make_layout(n_cap, C);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 256);
  const int rpb = rows_per_block_host(C);
  k_bn_bwd_reduce<<<L.blocks, L.threads, 2 * L.threads * sizeof(float), st>>>(x, ldx, dy, lddy, n_cap, n_dev, C, L.RL,
                                                                              rpb, gamma, beta, mean, invstd, leak,
                                                                              part);
  const float* part_apply = part;
  int blocks_apply = L.blocks, rpb_apply = rpb;
  if (L.blocks > kMergeAbove) {
    float* merged = part + (int64_t)L.blocks * 3 * C;
    k_bn_merge_sums<<<(C + 31) / 32, 1024, 2 * 1024 * sizeof(float), st>>>(part, n_cap, n_dev, C, rpb, L.blocks, merged);
    part_apply = merged; blocks_apply = 1; rpb_apply = 1 << 30;
    count_launches(1);
  }
  k_bn_bwd_apply<<<apply_grid(n_cap, L.RL), L.threads, 2 * L.threads * sizeof(float), st>>>(
      x, ldx, dy, lddy, n_cap, n_dev, C, L.RL, rpb_apply, blocks_apply, part_apply, gamma, beta, mean, invstd, leak, train,
      dgamma, dbeta, dx, lddx);
  count_launches(2);
  B200_CHECK_LAUNCH("bn_bwd");
  return 0;
}

extern "C" int b200scn_bn_fwd_train_sync(const float* x, int64_t ldx, int64_t n_cap, const int32_t* n_dev, int64_t n_expect,
                                         int C, const float* gamma, const float* beta, float* running_mean,
                                         float* running_var, float eps, float retention, float leak, float* y,
                                         int64_t ldy, float* save_mean, float* save_invstd, uint32_t* sync,
                                         void* scratch, int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(C >= 1 && C <= 1024, "bn: C must be 1..1024 (got %d)", C);
  B200_CHECK_ARG(scratch_bytes >= b200scn_bn_scratch_bytes(n_cap, C), "bn_fwd_train: scratch too small");
  if (n_cap == 0) return 0;
  Layout L = make_layout(n_cap, C);
  const int G = sync != nullptr ? fused_grid(n_cap, n_expect, C, L) : 0;
  if (G == 0)
    return b200scn_bn_fwd_train(x, ldx, n_cap, n_dev, C, gamma, beta, running_mean, running_var, eps, retention, leak,
                                y, ldy, save_mean, save_invstd, scratch, scratch_bytes, stream);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 256);
  launch_pdl(k_bn_fwd_fused, dim3(G), dim3(L.threads), 3 * L.threads * sizeof(float), st, x, ldx, n_cap, n_dev, C, L.RL,
             part, reinterpret_cast<unsigned int*>(sync), gamma, beta, running_mean, running_var, eps, retention, leak, y,
             ldy, save_mean, save_invstd);
  count_launches(1);
  B200_CHECK_LAUNCH("bn_fwd_train_sync");
  return 0;
}

extern "C" int b200scn_bn_bwd_sync(const float* x, int64_t ldx, const float* dy, int64_t lddy, int64_t n_cap,
                                   const int32_t* n_dev, int64_t n_expect, int C, const float* gamma, const float* beta,
                                   const float* mean, const float* invstd, float leak, int train, float* dx,
                                   int64_t lddx, float* dgamma, float* dbeta, uint32_t* sync, void* scratch,
                                   int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(C >= 1 && C <= 1024, "bn: C must be 1..1024 (got %d)", C);
  B200_CHECK_ARG(scratch_bytes >= b200scn_bn_scratch_bytes(n_cap, C), "bn_bwd: scratch too small");
  Layout L = make_layout(n_cap < 1 ? 1 : n_cap, C);
  const int G = (sync != nullptr && n_cap > 0) ? fused_grid(n_cap, n_expect, C, L, true) : 0;
  if (G == 0)
    return b200scn_bn_bwd(x, ldx, dy, lddy, n_cap, n_dev, C, gamma, beta, mean, invstd, leak, train, dx, lddx, dgamma,
                          dbeta, scratch, scratch_bytes, stream);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 256);
  launch_pdl(k_bn_bwd_fused, dim3(G), dim3(L.threads), 2 * L.threads * sizeof(float), st, x, ldx, dy, lddy, n_cap, n_dev,
             C, L.RL, part, reinterpret_cast<unsigned int*>(sync), gamma, beta, mean, invstd, leak, train, dgamma, dbeta,
             dx, lddx);
  count_launches(1);
  B200_CHECK_LAUNCH("bn_bwd_sync");
  return 0;
}
