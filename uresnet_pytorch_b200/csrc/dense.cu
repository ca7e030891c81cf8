// Dense path (north-star job 4): the Conv{2,3}d / ConvTranspose{2,3}d / replicate-pad family of
// /root/reference/uresnet/models/uresnet_dense.py:12-226 as hand-written CUDA over
// channels-last [B, (D,) H, W, C] fp32 tensors.  A dense grid is the degenerate sparse case
// whose rule "table" is arithmetic, so the convolutions reuse the gather-GEMM row kernels of
// conv_simt.cu with an implicit index (common.cuh: dense_index) instead of a table in HBM:
//   conv fwd            out[o] = b + sum_k in[o*s + k - p] W[k]          (mode 1)
//   conv dgrad          din[i] = sum_k dout[(i + p - k)/s] W[k]^T        (mode 2, grids swapped)
//   convT fwd           out[o] = b + sum_k in[(o + p - k)/s] W[k]        (mode 2)
//   convT dgrad         din[i] = sum_k dout[i*s + k - p] W[k]^T          (mode 1, grids swapped)
//   wgrad (both)        dW[k]  = sum_o in[idx(o,k)]^T dout[o]
// The reference pads with F.pad(mode='replicate') before zero-padding-free convolutions
// (uresnet_dense.py:19-26,75-80); pad_replicate_{fwd,bwd} below do the same explicitly.
// This file is the exact-fp32 FFMA path and the entry points; the tensor-core math modes route forward / dgrad
// to the tcgen05 implicit-GEMM kernel (csrc/dense_tc.cu) and wgrad to csrc/wgrad_mma.cu.
#include "common.cuh"

namespace b200scn {

bool wgrad_mma_supported(int64_t, int, int, int, int64_t, int64_t, const void*, const void*);
int64_t wgrad_scratch_bytes_mma(int64_t, int, int, int);
int wgrad_table_mma(const float*, int64_t, const float*, int64_t, const int32_t*, int64_t, int, int, int64_t,
                    const int32_t*, int, int, float*, int, int, void*, int64_t, cudaStream_t, const DenseGeom*);

int conv_dense_simt(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int, int64_t, int, int, float*,
                    int64_t, const DenseGeom&, const float*, cudaStream_t);
int64_t wgrad_scratch_bytes_simt(int64_t, int, int, int);
int wgrad_table_simt(const float*, int64_t, const float*, int64_t, const int32_t*, int64_t, int, int, int64_t,
                     const int32_t*, int, int, float*, int, void*, int64_t, cudaStream_t, const DenseGeom&);

namespace {

struct PadGeom {
  int dims;
  int in_sz[3], out_sz[3], pl[3];
};

__global__ void k_pad_fwd(const float* __restrict__ in, float* __restrict__ out, int64_t n_out, int C, PadGeom pg) {
  const int64_t total = n_out * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t o = e / C;
    const int c = (int)(e - o * C);
    int oc[3];
    oc[2] = (int)(o % pg.out_sz[2]); o /= pg.out_sz[2];
    oc[1] = (int)(o % pg.out_sz[1]); o /= pg.out_sz[1];
    oc[0] = (int)(o % pg.out_sz[0]); o /= pg.out_sz[0];
    int64_t lin = o;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      int ic = oc[d] - pg.pl[d];
      ic = ic < 0 ? 0 : (ic >= pg.in_sz[d] ? pg.in_sz[d] - 1 : ic);
      lin = lin * pg.in_sz[d] + ic;
    }
    out[e] = __ldg(in + lin * C + c);
  }
}

// din[i] = sum of dout over every padded voxel that clamps to i (gather form: no atomics)
__global__ void k_pad_bwd(const float* __restrict__ dout, float* __restrict__ din, int64_t n_in, int C, PadGeom pg) {
  const int64_t total = n_in * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = e / C;
    const int c = (int)(e - i * C);
    int ic[3];
    ic[2] = (int)(i % pg.in_sz[2]); i /= pg.in_sz[2];
    ic[1] = (int)(i % pg.in_sz[1]); i /= pg.in_sz[1];
    ic[0] = (int)(i % pg.in_sz[0]); i /= pg.in_sz[0];
    const int64_t b = i;
    int lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = ic[d] == 0 ? 0 : ic[d] + pg.pl[d];
      hi[d] = ic[d] == pg.in_sz[d] - 1 ? pg.out_sz[d] - 1 : ic[d] + pg.pl[d];
    }
    float s = 0.f;
    for (int z = lo[0]; z <= hi[0]; ++z)
      for (int y = lo[1]; y <= hi[1]; ++y)
        for (int x = lo[2]; x <= hi[2]; ++x)
          s += __ldg(dout + (((b * pg.out_sz[0] + z) * pg.out_sz[1] + y) * (int64_t)pg.out_sz[2] + x) * C + c);
    din[e] = s;
  }
}

// column sums of an [n, C] matrix (bias gradients): per-CTA partials + ordered fold by the last CTA.
// thread = (column c, row lane rl): consecutive threads read consecutive columns of a row (coalesced for any C);
// the row lanes of a column are summed in lane order, the CTA partials in CTA order (deterministic).
constexpr int kColsumThreads = 256;
__global__ void __launch_bounds__(kColsumThreads) k_colsum(const float* __restrict__ x, int64_t ldx, int64_t n, int C,
                                                           float* __restrict__ part, unsigned int* counter,
                                                           float* __restrict__ out) {
  __shared__ bool is_last;
  __shared__ float sm[kColsumThreads];
  const int64_t rpb = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * rpb, r1 = r0 + rpb < n ? r0 + rpb : n;
  const int t = threadIdx.x;
  const int CW = C < kColsumThreads ? C : kColsumThreads;  // columns handled per sweep
  const int RL = kColsumThreads / CW;
  const int cl = t % CW, rl = t / CW;
  for (int c0 = 0; c0 < C; c0 += CW) {
    const int c = c0 + cl;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (rl < RL && c < C) {
      int64_t r = r0 + rl;
      for (; r + 3 * RL < r1; r += 4 * RL) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += __ldg(x + (r + u * RL) * ldx + c);
      }
      for (; r < r1; r += RL) s[0] += __ldg(x + r * ldx + c);
    }
    __syncthreads();
    sm[t] = (s[0] + s[1]) + (s[2] + s[3]);
    __syncthreads();
    if (rl == 0 && c < C) {
      float a = sm[cl];
      for (int q = 1; q < RL; ++q) a += sm[q * CW + cl];
      part[(int64_t)blockIdx.x * C + c] = a;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int b = 0; b < (int)gridDim.x; ++b) a += __ldcg(part + (int64_t)b * C + c);
    out[c] = a;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

int read_geom(const int32_t* g, DenseGeom* out) {
  out->mode = g[0];
  out->dims = g[1];
  for (int d = 0; d < 3; ++d) {
    out->in_sz[d] = g[2 + d];
    out->out_sz[d] = g[5 + d];
  }
  out->ks = g[8];
  out->stride = g[9];
  out->pad = g[10];
  if (out->mode < 1 || out->mode > 2 || (out->dims != 2 && out->dims != 3) || out->ks < 1 || out->ks > 7 ||
      out->stride < 1) {
    set_error("dense: bad geometry (mode=%d dims=%d ks=%d stride=%d)", out->mode, out->dims, out->ks, out->stride);
    return B200SCN_ERR_ARG;
  }
  return 0;
}

inline int taps(const DenseGeom& g) { return g.dims == 3 ? g.ks * g.ks * g.ks : g.ks * g.ks; }
inline int64_t out_rows(const DenseGeom& g, int batch) {
  return (int64_t)batch * g.out_sz[0] * g.out_sz[1] * g.out_sz[2];
}

}  // namespace

int read_geom_public(const int32_t* g, DenseGeom* out) { return read_geom(g, out); }

}  // namespace b200scn

using namespace b200scn;

extern "C" int b200scn_dense_conv(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si,
                                  int64_t w_so, const float* bias, const int32_t* geom_host, int batch, int Cin,
                                  int Cout, float* out, int64_t ldo, void* stream) {
  DenseGeom g;
  if (int rc = read_geom(geom_host, &g)) return rc;
  B200_CHECK_ARG(Cin >= 1 && Cout >= 1 && batch >= 1, "dense_conv: bad shape");
  const int64_t n = out_rows(g, batch);
  B200_CHECK_ARG(n < (1ll << 31), "dense_conv: too many output voxels");
  return conv_dense_simt(in, ldi, W, w_sk, w_si, w_so, taps(g), n, Cin, Cout, out, ldo, g, bias, (cudaStream_t)stream);
}

extern "C" int64_t b200scn_dense_wgrad_scratch_bytes(const int32_t* geom_host, int batch, int Ca, int Cb) {
  DenseGeom g;
  if (read_geom(geom_host, &g)) return -1;
  const int64_t a = wgrad_scratch_bytes_simt(out_rows(g, batch), taps(g), Ca, Cb);
  const int64_t b = wgrad_scratch_bytes_mma(out_rows(g, batch), taps(g), Ca, Cb);
  return a > b ? a : b;
}

/* dW[k][ca][cb] = sum_o A[idx(o,k)][ca] * B[o][cb]  (gather_b == 0)   or   A[o][ca] * B[idx(o,k)][cb]  (gather_b == 1).
 * math_mode != B200SCN_MATH_FP32 and 16-multiple channel counts: the tensor-core kernel of the sparse path
 * (csrc/wgrad_mma.cu) with the rule of (voxel, tap) computed from the geometry instead of read from a table. */
extern "C" int b200scn_dense_wgrad(const float* A, int64_t lda, const float* B, int64_t ldb, const int32_t* geom_host,
                                   int batch, int gather_b, int Ca, int Cb, float* dW, int math_mode, void* scratch,
                                   int64_t scratch_bytes, void* stream) {
  DenseGeom g;
  if (int rc = read_geom(geom_host, &g)) return rc;
  const int64_t n = out_rows(g, batch);
  B200_CHECK_ARG(n < (1ll << 31), "dense_wgrad: too many voxels");
  if (math_mode != B200SCN_MATH_FP32 && taps(g) <= 64 && wgrad_mma_supported(n, taps(g), Ca, Cb, lda, ldb, A, B))
    return wgrad_table_mma(A, lda, B, ldb, nullptr, 0, taps(g), gather_b, n, nullptr, Ca, Cb, dW, 0, math_mode, scratch,
                           scratch_bytes, (cudaStream_t)stream, &g);
  return wgrad_table_simt(A, lda, B, ldb, nullptr, 0, taps(g), gather_b, n, nullptr, Ca, Cb, dW, 0, scratch,
                          scratch_bytes, (cudaStream_t)stream, g);
}

/* sizes are (z, y, x) with z = 1 for 2D; out = in + pl + pr per dimension */
extern "C" int b200scn_pad_replicate(const float* src, float* dst, int batch, int C, const int32_t* in_sz_host,
                                     const int32_t* pl_host, const int32_t* pr_host, int backward, void* stream) {
  PadGeom pg;
  pg.dims = 3;
  int64_t n_in = batch, n_out = batch;
  for (int d = 0; d < 3; ++d) {
    B200_CHECK_ARG(in_sz_host[d] >= 1 && pl_host[d] >= 0 && pr_host[d] >= 0, "pad_replicate: bad geometry");
    pg.in_sz[d] = in_sz_host[d];
    pg.pl[d] = pl_host[d];
    pg.out_sz[d] = in_sz_host[d] + pl_host[d] + pr_host[d];
    n_in *= pg.in_sz[d];
    n_out *= pg.out_sz[d];
  }
  const int64_t total = (backward ? n_in : n_out) * C;
  int64_t grid = cdiv(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (backward)
    k_pad_bwd<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(src, dst, n_in, C, pg);
  else
    k_pad_fwd<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(src, dst, n_out, C, pg);
  count_launches(1);
  B200_CHECK_LAUNCH("pad_replicate");
  return 0;
}

extern "C" int64_t b200scn_colsum_scratch_bytes(int C) { return 256 + (int64_t)sm_count() * 2 * C * 4; }

extern "C" int b200scn_colsum(const float* x, int64_t ldx, int64_t n, int C, float* out, void* scratch,
                              int64_t scratch_bytes, void* stream) {
  B200_CHECK_ARG(C >= 1 && n >= 0, "colsum: bad shape");
  B200_CHECK_ARG(scratch_bytes >= b200scn_colsum_scratch_bytes(C), "colsum: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 256);
  B200_CUDA(cudaMemsetAsync(counter, 0, 4, st));
  int64_t grid = cdiv(n, 256);
  const int64_t cap = (int64_t)sm_count() * 2;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_colsum<<<(int)grid, kColsumThreads, 0, st>>>(x, ldx, n, C, part, counter, out);
  count_launches(1);
  B200_CHECK_LAUNCH("colsum");
  return 0;
}
