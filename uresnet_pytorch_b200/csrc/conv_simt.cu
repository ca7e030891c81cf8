// Gather-GEMM sparse convolutions, exact-fp32 (FFMA) kernels: forward/dgrad (table and
// parent forms) and wgrad (north-star job 2).  These are the MATH_FP32 path and the
// any-channel-count path; sparse_tc.cu (tcgen05) and conv_gpl.cu / conv_pairs.cu (mma.sync) hold the tensor-core
// kernels for channel counts that are multiples of 16.
//
// Output-stationary formulation: every CTA owns a tile of output rows, walks the kernel
// offsets k = 0..K-1 in order, gathers the input rows named by table[k][row] into shared
// memory and accumulates in registers -- no atomics, no scatter conflicts, and a fixed
// accumulation order, which is what makes results reproducible
// against the oracle.  scn's scheme for the same layers (per offset: gather rows, GEMM,
// scatter-add) moves every output row once per offset; this one writes it once.
#include "common.cuh"

namespace b200scn {
namespace {

constexpr int kTM = 64;       // output rows per CTA
constexpr int kKC = 32;       // input-channel chunk staged in smem
constexpr int kThreads = 256;

// NC = number of 16-wide output column groups each thread accumulates
template <int NC>
__global__ void __launch_bounds__(kThreads)
k_conv_rows(const float* __restrict__ in, int64_t ldi, const float* __restrict__ W, int64_t w_sk, int64_t w_si,
            int64_t w_so, int kflip, const int32_t* __restrict__ table, int64_t ld,
            const int32_t* __restrict__ parent, const uint8_t* __restrict__ kofs, int K, int64_t n_cap,
            const int32_t* __restrict__ n_dev, int Cin, int Cout, float* __restrict__ out, int64_t ldo,
            int accumulate, const DenseGeom geom, const float* __restrict__ bias) {
  __shared__ float in_s[kTM][kKC + 1];
  __shared__ float w_s[kKC][NC * 16];
  __shared__ int idx_s[kTM];

  const int n = live_rows(n_dev, n_cap);
  const int t = threadIdx.x;
  const int c0 = t & 15, rg = t >> 4;

  for (int row0 = blockIdx.x * kTM; row0 < n; row0 += gridDim.x * kTM) {
    float acc[4][NC];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int m = 0; m < NC; ++m) acc[j][m] = 0.f;

    int my_kofs = -1, my_parent = -1;
    if (parent != nullptr && t < kTM && row0 + t < n) {
      my_kofs = kofs[row0 + t];
      my_parent = parent[row0 + t];
    }

    for (int k = 0; k < K; ++k) {
      int e = -1;
      if (t < kTM && row0 + t < n) {
        if (parent != nullptr) e = (my_kofs == k) ? my_parent : -1;
        else if (table != nullptr) e = table[(int64_t)k * ld + row0 + t];
        else if (geom.mode != 0) e = dense_index(geom, row0 + t, k);
        else e = row0 + t;
      }
      __syncthreads();  // previous offset's readers of idx_s / in_s / w_s are done
      if (t < kTM) idx_s[t] = e;
      if (!__syncthreads_or(e >= 0)) continue;
      const int kk = kflip ? (K - 1 - k) : k;
      bool valid[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) valid[j] = idx_s[rg + 16 * j] >= 0;

      for (int ci0 = 0; ci0 < Cin; ci0 += kKC) {
        if (ci0 > 0) __syncthreads();
        // stage W[kk][ci0 .. ci0+KC) x [0, Cout)
        for (int e2 = t; e2 < kKC * NC * 16; e2 += kThreads) {
          const int cc = e2 / (NC * 16), co = e2 % (NC * 16);
          float v = 0.f;
          if (ci0 + cc < Cin && co < Cout) v = __ldg(W + (int64_t)kk * w_sk + (int64_t)(ci0 + cc) * w_si + (int64_t)co * w_so);
          w_s[cc][co] = v;
        }
        // gather input rows, one warp per row -> 128 B coalesced
#pragma unroll
        for (int it = 0; it < kTM * kKC / kThreads; ++it) {
          const int r = (t + it * kThreads) >> 5, cc = t & 31;
          const int src = idx_s[r];
          float v = 0.f;
          if (src >= 0 && ci0 + cc < Cin) v = __ldg(in + (int64_t)src * ldi + ci0 + cc);
          in_s[r][cc] = v;
        }
        __syncthreads();
        const int kc = min(kKC, Cin - ci0);
        for (int cc = 0; cc < kc; ++cc) {
          float w[NC];
#pragma unroll
          for (int m = 0; m < NC; ++m) w[m] = w_s[cc][c0 + 16 * m];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (valid[j]) {
              const float a = in_s[rg + 16 * j][cc];
#pragma unroll
              for (int m = 0; m < NC; ++m) acc[j][m] = fmaf(a, w[m], acc[j][m]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = row0 + rg + 16 * j;
      if (r < n) {
#pragma unroll
        for (int m = 0; m < NC; ++m) {
          const int c = c0 + 16 * m;
          if (c < Cout) {
            float* p = out + (int64_t)r * ldo + c;
            const float v = acc[j][m] + (bias != nullptr ? __ldg(bias + c) : 0.f);
            *p = accumulate ? (*p + v) : v;
          }
        }
      }
    }
    __syncthreads();
  }
}

// Stem layer (one input feature, uresnet_sparse.py:21): out[o][c] = sum_k in[tbl[k][o]] * W[k][0][c].
// Pure bandwidth work: one thread per (row, output channel), the table entry is a warp broadcast.
__global__ void k_conv_cin1(const float* __restrict__ in, int64_t ldi, const float* __restrict__ W, int64_t w_sk,
                            int64_t w_so, int kflip, const int32_t* __restrict__ table, int64_t ld, int K,
                            int64_t n_cap, const int32_t* __restrict__ n_dev, int Cout, float* __restrict__ out,
                            int64_t ldo) {
  const int64_t n = live_rows(n_dev, n_cap);
  const int64_t total = n * Cout;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / Cout;
    const int c = (int)(e - r * Cout);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const int idx = __ldg(table + (int64_t)k * ld + r);
      if (idx >= 0) acc = fmaf(__ldg(in + (int64_t)idx * ldi), __ldg(W + (int64_t)(kflip ? K - 1 - k : k) * w_sk + c * w_so), acc);
    }
    out[r * ldo + c] = acc;
  }
}

// Its weight gradient (Ca == 1): dW[k][0][c] = sum_r A[ra] * B[rb][c]; per-chunk partials, folded
// in chunk order by k_wgrad_reduce.
__global__ void k_wgrad_ca1(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                            const int32_t* __restrict__ table, int64_t ld, int K, int gather_b, int64_t n_cap,
                            const int32_t* __restrict__ n_dev, int Cb, int rows_per_split, float* __restrict__ part) {
  extern __shared__ float red[];  // [RL][Cb]
  const int n = live_rows(n_dev, n_cap);
  const int k = blockIdx.x, split = blockIdx.y;
  const int r_begin = split * rows_per_split;
  if (r_begin >= n) return;
  const int r_end = min(n, r_begin + rows_per_split);
  const int t = threadIdx.x, c = t % Cb, rl = t / Cb, RL = blockDim.x / Cb;
  float acc = 0.f;
  if (rl < RL) {
    for (int r = r_begin + rl; r < r_end; r += RL) {
      const int e = table ? __ldg(table + (int64_t)k * ld + r) : r;
      if (e >= 0) {
        const int ra = gather_b ? r : e, rb = gather_b ? e : r;
        acc = fmaf(__ldg(A + (int64_t)ra * lda), __ldg(B + (int64_t)rb * ldb + c), acc);
      }
    }
    red[rl * Cb + c] = acc;
  }
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < RL; ++j) acc += red[j * Cb + c];
    part[((int64_t)split * K + k) * Cb + c] = acc;
  }
}

template <int NC>
void launch_rows(int grid, cudaStream_t st, const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si,
                 int64_t w_so, int kflip, const int32_t* table, int64_t ld, const int32_t* parent,
                 const uint8_t* kofs, int K, int64_t n_cap, const int32_t* n_dev, int Cin, int Cout, float* out,
                 int64_t ldo, int accumulate, const DenseGeom& geom, const float* bias) {
  k_conv_rows<NC><<<grid, kThreads, 0, st>>>(in, ldi, W, w_sk, w_si, w_so, kflip, table, ld, parent, kofs, K, n_cap,
                                             n_dev, Cin, Cout, out, ldo, accumulate, geom, bias);
}

int conv_rows_dispatch(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so,
                       int kflip, const int32_t* table, int64_t ld, const int32_t* parent, const uint8_t* kofs,
                       int K, int64_t n_cap, const int32_t* n_dev, int Cin, int Cout, float* out, int64_t ldo,
                       int accumulate, cudaStream_t st, const DenseGeom& geom = DenseGeom{}, const float* bias = nullptr) {
  if (n_cap == 0) return 0;
  if (Cin == 1 && table != nullptr && parent == nullptr && geom.mode == 0 && !accumulate && bias == nullptr) {
    int64_t g = cdiv(n_cap * Cout, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    k_conv_cin1<<<(int)g, 256, 0, st>>>(in, ldi, W, w_sk, w_so, kflip, table, ld, K, n_cap, n_dev, Cout, out, ldo);
    count_launches(1);
    B200_CHECK_LAUNCH("conv_cin1");
    return 0;
  }
  int64_t g64 = cdiv(n_cap, kTM);
  const int64_t gmax = (int64_t)sm_count() * 8;
  const int grid = (int)(g64 < gmax ? g64 : gmax);
  for (int cb = 0; cb < Cout; cb += 128) {  // column blocks of <= 128 outputs
    const int co = Cout - cb < 128 ? Cout - cb : 128;
    const int nc = (co + 15) / 16;
    const float* Wb = W + (int64_t)cb * w_so;
    float* ob = out + cb;
    const float* bb = bias ? bias + cb : nullptr;
#define B200_ROWS(NCV)                                                                                          \
  launch_rows<NCV>(grid, st, in, ldi, Wb, w_sk, w_si, w_so, kflip, table, ld, parent, kofs, K, n_cap, n_dev, Cin, \
                   co, ob, ldo, accumulate, geom, bb)
    switch (nc) {
      case 1: B200_ROWS(1); break;
      case 2: B200_ROWS(2); break;
      case 3: B200_ROWS(3); break;
      case 4: B200_ROWS(4); break;
      case 5: B200_ROWS(5); break;
      case 6: B200_ROWS(6); break;
      default: B200_ROWS(8); break;
    }
#undef B200_ROWS
    count_launches(1);
  }
  B200_CHECK_LAUNCH("conv_rows");
  return 0;
}

// ---- wgrad ------------------------------------------------------------------------------------
constexpr int kWT = 64;   // dW tile is kWT x kWT
constexpr int kWR = 32;   // pairs staged per iteration

__global__ void __launch_bounds__(kThreads)
k_wgrad_partial(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                const int32_t* __restrict__ table, int64_t ld, int K, int gather_b, int64_t n_cap,
                const int32_t* __restrict__ n_dev, int Ca, int Cb, int tiles_a, int tiles_b, int rows_per_split,
                float* __restrict__ part, const DenseGeom geom) {
  __shared__ __align__(16) float a_s[kWR][kWT];
  __shared__ __align__(16) float b_s[kWR][kWT];
  __shared__ int ia_s[kWR], ib_s[kWR];
  __shared__ int wcnt[2];

  const int n = live_rows(n_dev, n_cap);
  const int split = blockIdx.y;
  const int r_begin = split * rows_per_split;
  if (r_begin >= n) return;
  const int r_end = min(n, r_begin + rows_per_split);
  int bx = blockIdx.x;
  const int tb = bx % tiles_b; bx /= tiles_b;
  const int ta = bx % tiles_a; bx /= tiles_a;
  const int k = bx;
  const int ca0 = ta * kWT, cb0 = tb * kWT;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int r0 = r_begin; r0 < r_end; r0 += kWR) {
    // compact the valid (ra, rb) pairs of this row chunk, preserving row order
    int e = -1, r = r0 + t;
    if (t < kWR && r < r_end) e = table ? table[(int64_t)k * ld + r] : (geom.mode != 0 ? dense_index(geom, r, k) : r);
    unsigned m = 0;
    if (t < 32 * (kWR / 32)) m = __ballot_sync(0xffffffffu, e >= 0);
    __syncthreads();  // previous chunk's readers done
    if (t < kWR && (t & 31) == 0) wcnt[t >> 5] = __popc(m);
    __syncthreads();
    int cnt = 0;
#pragma unroll
    for (int w = 0; w < kWR / 32; ++w) cnt += wcnt[w];
    if (t < kWR && e >= 0) {
      int pos = __popc(m & ((1u << (t & 31)) - 1));
      for (int w = 0; w < (t >> 5); ++w) pos += wcnt[w];
      ia_s[pos] = gather_b ? r : e;
      ib_s[pos] = gather_b ? e : r;
    }
    __syncthreads();
    if (cnt == 0) continue;
    // gather: warp per pair, 2 x 32 columns
    for (int it = t; it < kWR * kWT; it += kThreads) {
      const int p = it / kWT, c = it % kWT;
      float va = 0.f, vb = 0.f;
      if (p < cnt) {
        if (ca0 + c < Ca) va = __ldg(A + (int64_t)ia_s[p] * lda + ca0 + c);
        if (cb0 + c < Cb) vb = __ldg(B + (int64_t)ib_s[p] * ldb + cb0 + c);
      }
      a_s[p][c] = va;
      b_s[p][c] = vb;
    }
    __syncthreads();
    for (int p = 0; p < cnt; ++p) {
      const float4 a4 = *reinterpret_cast<const float4*>(&a_s[p][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&b_s[p][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float* dst = part + ((int64_t)split * K + k) * Ca * Cb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ca = ca0 + ty * 4 + i;
    if (ca < Ca) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cb = cb0 + tx * 4 + j;
        if (cb < Cb) dst[(int64_t)ca * Cb + cb] = acc[i][j];
      }
    }
  }
}

__global__ void k_wgrad_reduce(const float* __restrict__ part, int64_t elems, int rows_per_split, int nsplit,
                               int64_t n_cap, const int32_t* __restrict__ n_dev, float* __restrict__ dW,
                               int accumulate) {
  const int n = live_rows(n_dev, n_cap);
  int live = (n + rows_per_split - 1) / rows_per_split;
  if (live > nsplit) live = nsplit;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < elems; e += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int sp = 0; sp < live; ++sp) s += part[(int64_t)sp * elems + e];
    dW[e] = accumulate ? dW[e] + s : s;
  }
}

void wgrad_plan(int64_t n_cap, int K, int Ca, int Cb, int* nsplit, int* rows_per_split) {
  int64_t ns = cdiv(n_cap, Ca == 1 ? 256 : 2048);  // the one-input-channel kernel is pure streaming: small chunks
  if (ns < 1) ns = 1;
  if (ns > 128) ns = 128;
  const int64_t per = (int64_t)K * Ca * Cb * 4;
  int64_t lim = (64ll << 20) / (per > 0 ? per : 1);
  if (lim < 1) lim = 1;
  if (ns > lim) ns = lim;
  int64_t rps = cdiv(n_cap, ns);
  rps = cdiv(rps, kWR) * kWR;
  if (rps < kWR) rps = kWR;
  *nsplit = (int)ns;
  *rows_per_split = (int)rps;
}

}  // namespace

int conv_table_simt(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so, int kflip,
                    const int32_t* table, int64_t ld, int K, int64_t n_cap, const int32_t* n_dev, int Cin, int Cout,
                    float* out, int64_t ldo, int accumulate, cudaStream_t st) {
  return conv_rows_dispatch(in, ldi, W, w_sk, w_si, w_so, kflip, table, ld, nullptr, nullptr, K, n_cap, n_dev, Cin,
                            Cout, out, ldo, accumulate, st);
}

int conv_parent_simt(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so,
                     const int32_t* parent, const uint8_t* kofs, int K, int64_t n_cap, const int32_t* n_dev, int Cin,
                     int Cout, float* out, int64_t ldo, cudaStream_t st) {
  return conv_rows_dispatch(in, ldi, W, w_sk, w_si, w_so, 0, nullptr, 0, parent, kofs, K, n_cap, n_dev, Cin, Cout, out,
                            ldo, 0, st);
}

int conv_dense_simt(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so, int K,
                    int64_t n_out, int Cin, int Cout, float* out, int64_t ldo, const DenseGeom& geom, const float* bias,
                    cudaStream_t st) {
  return conv_rows_dispatch(in, ldi, W, w_sk, w_si, w_so, 0, nullptr, 0, nullptr, nullptr, K, n_out, nullptr, Cin, Cout,
                            out, ldo, 0, st, geom, bias);
}

int64_t wgrad_scratch_bytes_simt(int64_t n_cap, int K, int Ca, int Cb) {
  int ns, rps;
  wgrad_plan(n_cap < 1 ? 1 : n_cap, K, Ca, Cb, &ns, &rps);
  return (int64_t)ns * K * Ca * Cb * 4 + 256;
}

int wgrad_table_simt(const float* A, int64_t lda, const float* B, int64_t ldb, const int32_t* table, int64_t ld, int K,
                     int gather_b, int64_t n_cap, const int32_t* n_dev, int Ca, int Cb, float* dW, int accumulate,
                     void* scratch, int64_t scratch_bytes, cudaStream_t st, const DenseGeom& geom) {
  const int64_t elems = (int64_t)K * Ca * Cb;
  if (n_cap == 0) {
    if (!accumulate) B200_CUDA(cudaMemsetAsync(dW, 0, (size_t)elems * 4, st));
    return 0;
  }
  int ns, rps;
  wgrad_plan(n_cap, K, Ca, Cb, &ns, &rps);
  B200_CHECK_ARG(scratch_bytes >= (int64_t)ns * elems * 4, "wgrad: scratch too small (%lld < %lld)",
                 (long long)scratch_bytes, (long long)((int64_t)ns * elems * 4));
  const int tiles_a = (Ca + kWT - 1) / kWT, tiles_b = (Cb + kWT - 1) / kWT;
  dim3 grid(K * tiles_a * tiles_b, ns);
  if (Ca == 1 && Cb <= 256 && geom.mode == 0) {
    const int RL = 256 / Cb;
    k_wgrad_ca1<<<dim3(K, ns), RL * Cb, RL * Cb * sizeof(float), st>>>(A, lda, B, ldb, table, ld, K, gather_b, n_cap,
                                                                      n_dev, Cb, rps, reinterpret_cast<float*>(scratch));
  } else
  k_wgrad_partial<<<grid, kThreads, 0, st>>>(A, lda, B, ldb, table, ld, K, gather_b, n_cap, n_dev, Ca, Cb, tiles_a,
                                             tiles_b, rps, reinterpret_cast<float*>(scratch), geom);
  int64_t g = cdiv(elems, 256);
  if (g > 1024) g = 1024;
  k_wgrad_reduce<<<(int)g, 256, 0, st>>>(reinterpret_cast<const float*>(scratch), elems, rps, ns, n_cap, n_dev, dW,
                                         accumulate);
  count_launches(2);
  B200_CHECK_LAUNCH("wgrad_table");
  return 0;
}

}  // namespace b200scn
