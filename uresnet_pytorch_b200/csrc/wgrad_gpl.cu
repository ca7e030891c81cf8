// Weight gradients over the grouped pair lists (csrc/conv_gpl.cu) on the tensor pipe:
//     dW[k][ci][co] = sum over the rule pairs (in -> out) of offset k of A[in][ci] * B[out][co]
// (A = layer input rows, B = output-gradient rows; the list is the one the layer's FORWARD uses, so its
// "out" rows index B and its "in" rows index A for SubmanifoldConvolution, Convolution and Deconvolution
// alike).  The lists already hold the pairs of every (tile, offset) compacted in groups of 16, so there
// is no neighbour-table scan, no compaction and no shared-memory staging here (csrc/wgrad_mma.cu does
// all three per launch): a warp reads a group's record and gathers straight into mma fragments.
//
// One mma.sync m16n8k8 tf32 per (16 input channels, 8 output channels, 8 PAIRS): the pairs are the K
// dimension.  Fragment <-> channel maps (lane = 4 g + t), chosen so that every gather is a vector load of
// consecutive channels of ONE row, 128 contiguous bytes per 8 lanes:
//   A, 32-channel block b: lane loads A[row(pair)][32b + 4g .. +3] for pairs t, t+4 (k-step 0) / t+8, t+12
//       m-tile 2b   : row g -> ci 32b + 4g,     row g+8 -> ci 32b + 4g + 1
//       m-tile 2b+1 : row g -> ci 32b + 4g + 2, row g+8 -> ci 32b + 4g + 3
//      (a trailing 16-channel block uses 8-byte loads: row g -> ci base + 2g, row g+8 -> base + 2g + 1)
//   B, 32-channel block b: lane loads B[row(pair)][32b + 4g .. +3]; n-tile 4b + j: column g -> co 32b + 4g + j
//      (16-channel block: n-tile j in {0,1}: column g -> co base + 2g + j)
// CTA = (offset k [+ extra CTAs for the dense centre offset], part of the tile range, channel slice);
// its 8 warps take tiles round-robin, keep their (ci, co) tiles in registers, are summed in warp order
// through shared memory, and one partial block per CTA is folded by k_wgrad_gpl_fold in fixed order:
// bit-reproducible, no atomics.
#include <stdlib.h>

#include "common.cuh"

namespace b200scn {
namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

struct WgGplParams {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  const int4* g_rec; const int32_t* koff;  // koff[tile][K+1]: first group of every offset inside the tile
  int K, tile_rows, centre, extra;         // centre offset (or -1) gets `extra` additional CTAs per part
  int64_t n_cap; const int32_t* n_dev;
  int Ca, Cb, ca_chunks, cb_chunks;        // channel slices of 16*MT / 8*NT channels (grid.z)
  int parts;
  float* part;                              // [parts][K + extra][Ca][Cb]
  uint32_t ones;
};

__host__ __device__ __forceinline__ int wgg_tiles_per_part(int n_tiles, int parts) { return (n_tiles + parts - 1) / parts; }

// MT m-tiles (16 ci each), NT n-tiles (8 co each) per CTA slice; MT in 1..4, NT in {2, 4}
template <int MT, int NT, bool X3>
__global__ void __launch_bounds__(kWarps * 32) k_wgrad_gpl(const WgGplParams p) {
  extern __shared__ __align__(16) float red[];  // [16 MT][8 NT] cross-warp sum
  pdl_wait();
  const int n = live_rows(p.n_dev, p.n_cap);
  const int R = p.tile_rows, K = p.K;
  const int n_tiles = (n + R - 1) / R;
  const int tpp = wgg_tiles_per_part(n_tiles, p.parts);
  const int t_begin0 = blockIdx.y * tpp, t_end0 = min(n_tiles, t_begin0 + tpp);
  if (t_begin0 >= t_end0 && blockIdx.y != 0) return;  // part 0 always writes (n == 0: zeros)
  // offset and tile sub-range of this CTA: the centre offset's range is cut into 1 + extra pieces
  int k = blockIdx.x, t_begin = t_begin0, t_end = t_end0;
  if (p.extra > 0 && (k == p.centre || k >= K)) {
    const int piece = k >= K ? (k - K + 1) : 0;
    const int len = t_end0 - t_begin0, per = (len + p.extra) / (p.extra + 1);
    t_begin = min(t_end0, t_begin0 + piece * per);
    t_end = min(t_end0, t_begin + per);
    if (k >= K) k = p.centre;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int zc = blockIdx.z / p.cb_chunks, zb = blockIdx.z - zc * p.cb_chunks;
  const int ca0 = zc * (16 * MT), cb0 = zb * (8 * NT);
  const int mt_valid = min(MT, (p.Ca - ca0) >> 4), nt_valid = min(NT, (p.Cb - cb0) >> 3);

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  const float* const a_l = p.A + ca0;
  const float* const b_l = p.B + cb0;
  for (int tile = t_begin + warp; tile < t_end; tile += kWarps) {
    const int gb = __ldg(p.koff + (int64_t)tile * (K + 1) + k), ge = __ldg(p.koff + (int64_t)tile * (K + 1) + k + 1);
    const int base = tile * R;
    for (int grp = gb; grp < ge; ++grp) {
      // entries t and t+4 of the record describe pairs t, t+8 and t+4, t+12
      const int4 e0 = __ldg(p.g_rec + (int64_t)grp * 8 + t), e1 = __ldg(p.g_rec + (int64_t)grp * 8 + t + 4);
#pragma unroll
      for (int s = 0; s < 2; ++s) {  // k-step s: pairs 8s + t (P1) and 8s + t + 4 (P2)
        const int in1 = s ? e0.y : e0.x, in2 = s ? e1.y : e1.x;
        const int o1 = (e0.z >> (8 * s)) & 0xff, o2 = (e1.z >> (8 * s)) & 0xff;
        // padding slots (out == R): any valid rows, their A operand is forced to zero
        const bool v1 = o1 < R, v2 = o2 < R;
        const float* a1 = a_l + (int64_t)in1 * p.lda;
        const float* a2 = a_l + (int64_t)in2 * p.lda;
        const float* b1 = b_l + (int64_t)(base + (v1 ? o1 : 0)) * p.ldb;
        const float* b2 = b_l + (int64_t)(base + (v2 ? o2 : 0)) * p.ldb;
        // B fragments of all n-tiles
        uint32_t bh[NT][2], bl[NT][2];
        if (NT == 4) {
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(b1) + g), w2 = __ldg(reinterpret_cast<const float4*>(b2) + g);
          const float x1[4] = {w1.x, w1.y, w1.z, w1.w}, x2[4] = {w2.x, w2.y, w2.z, w2.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (X3) {
              const float l1 = x1[j] - tf32_hi(x1[j]), l2 = x2[j] - tf32_hi(x2[j]);
              bl[j][0] = __float_as_uint(l1); bl[j][1] = __float_as_uint(l2);
              bh[j][0] = __float_as_uint(x1[j] - l1); bh[j][1] = __float_as_uint(x2[j] - l2);
            } else {
              bh[j][0] = __float_as_uint(x1[j]) & p.ones; bh[j][1] = __float_as_uint(x2[j]) & p.ones;
            }
          }
        } else {
          const float2 w1 = __ldg(reinterpret_cast<const float2*>(b1) + g), w2 = __ldg(reinterpret_cast<const float2*>(b2) + g);
          const float x1[2] = {w1.x, w1.y}, x2[2] = {w2.x, w2.y};
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (X3) {
              const float l1 = x1[j] - tf32_hi(x1[j]), l2 = x2[j] - tf32_hi(x2[j]);
              bl[j][0] = __float_as_uint(l1); bl[j][1] = __float_as_uint(l2);
              bh[j][0] = __float_as_uint(x1[j] - l1); bh[j][1] = __float_as_uint(x2[j] - l2);
            } else {
              bh[j][0] = __float_as_uint(x1[j]) & p.ones; bh[j][1] = __float_as_uint(x2[j]) & p.ones;
            }
          }
        }
        // A fragments, one or two m-tiles per load
#pragma unroll
        for (int mb = 0; mb < (MT + 1) / 2; ++mb) {
          float av[2][4];  // [m-tile inside the block][a0..a3]
          if (2 * mb >= mt_valid) break;  // channel slice narrower than the template (last Ca chunk)
          if (2 * mb + 1 < mt_valid) {  // full 32-channel block
            float4 u1 = __ldg(reinterpret_cast<const float4*>(a1 + 32 * mb) + g);
            float4 u2 = __ldg(reinterpret_cast<const float4*>(a2 + 32 * mb) + g);
            if (!v1) u1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!v2) u2 = make_float4(0.f, 0.f, 0.f, 0.f);
            av[0][0] = u1.x; av[0][1] = u1.y; av[0][2] = u2.x; av[0][3] = u2.y;
            av[1][0] = u1.z; av[1][1] = u1.w; av[1][2] = u2.z; av[1][3] = u2.w;
          } else {  // trailing 16-channel block
            float2 u1 = __ldg(reinterpret_cast<const float2*>(a1 + 32 * mb) + g);
            float2 u2 = __ldg(reinterpret_cast<const float2*>(a2 + 32 * mb) + g);
            if (!v1) u1 = make_float2(0.f, 0.f);
            if (!v2) u2 = make_float2(0.f, 0.f);
            av[0][0] = u1.x; av[0][1] = u1.y; av[0][2] = u2.x; av[0][3] = u2.y;
            av[1][0] = av[1][1] = av[1][2] = av[1][3] = 0.f;
          }
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const int mt = 2 * mb + mi;
            if (mt < MT && mt < mt_valid) {
              uint32_t ah[4], al[4];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                if (X3) {
                  const float lo = av[mi][r] - tf32_hi(av[mi][r]);
                  al[r] = __float_as_uint(lo);
                  ah[r] = __float_as_uint(av[mi][r] - lo);
                } else {
                  ah[r] = __float_as_uint(av[mi][r]) & p.ones;
                }
              }
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                if (X3) {
                  mma_tf32(acc[mt][j], al, bh[j]);
                  mma_tf32(acc[mt][j], ah, bl[j]);
                }
                mma_tf32(acc[mt][j], ah, bh[j]);
              }
            }
          }
        }
      }
    }
  }

  // cross-warp sum in warp order, in the (ci, co) layout of dW
  constexpr int CA = 16 * MT, CB = 8 * NT;
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (mt >= mt_valid) break;  // (an unused m-tile would alias the half-block map of a used one)
        // channel of fragment row g / g+8 of m-tile mt (see the maps above)
        const bool full = (mt | 1) < mt_valid;  // the m-tile belongs to a full 32-channel block
        const int cbase = 32 * (mt >> 1);
        const int ci0 = full ? cbase + 4 * g + 2 * (mt & 1) : cbase + 2 * g;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int ci = ci0 + (r >> 1);
            const int c = 2 * t + (r & 1);  // fragment column
            const int co = NT == 4 ? 4 * c + j : 2 * c + j;
            float* dst = red + ci * CB + co;
            *dst = (w == 0 ? 0.f : *dst) + acc[mt][j][r];
          }
        }
      }
    }
    __syncthreads();
  }
  float* out = p.part + ((int64_t)blockIdx.y * (K + p.extra) + blockIdx.x) * p.Ca * p.Cb;
  for (int e = tid; e < CA * CB; e += blockDim.x) {
    const int ci = e / CB, co = e - ci * CB;
    if ((ci >> 4) < mt_valid && (co >> 3) < nt_valid) out[(int64_t)(ca0 + ci) * p.Cb + cb0 + co] = red[e];
  }
}

__global__ void k_wgrad_gpl_fold(const float* __restrict__ part, int K, int extra, int centre, int64_t cc, int parts,
                                 int tile_rows, int64_t n_cap, const int32_t* __restrict__ n_dev, float* __restrict__ dW,
                                 int accumulate) {
  pdl_wait();
  const int n = live_rows(n_dev, n_cap);
  const int n_tiles = (n + tile_rows - 1) / tile_rows;
  const int tpp = wgg_tiles_per_part(n_tiles, parts);
  const int live = n_tiles > 0 ? (n_tiles + tpp - 1) / tpp : 1;  // part 0 always writes
  const int64_t stride = (int64_t)(K + extra) * cc;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)K * cc; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e / cc);
    const int64_t o = e - (int64_t)k * cc;
    float s = 0.f;
    int y = 0;
    if (k != centre) {
      for (; y + 4 <= live; y += 4) {  // four partial blocks' loads in flight, summed in part order
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcg(part + (y + u) * stride + e);
#pragma unroll
        for (int u = 0; u < 4; ++u) s += v[u];
      }
    }
    for (; y < live; ++y) {
      s += __ldcg(part + y * stride + e);
      if (k == centre)
        for (int x = 0; x < extra; ++x) s += __ldcg(part + y * stride + (int64_t)(K + x) * cc + o);
    }
    dW[e] = accumulate ? dW[e] + s : s;
  }
}

struct WggPlan {
  int MT, NT, ca_chunks, cb_chunks, parts, extra, centre;
};

WggPlan wgg_plan(int K, int Ca, int Cb, int64_t n_cap, int64_t n_expect, int tile_rows, int subm) {
  WggPlan pl;
  pl.MT = Ca >= 64 ? 4 : Ca / 16;
  pl.ca_chunks = (Ca + 16 * pl.MT - 1) / (16 * pl.MT);
  pl.NT = (Cb % 32 == 0) ? 4 : 2;
  pl.cb_chunks = Cb / (8 * pl.NT);
  // the centre offset of a submanifold convolution pairs every row with itself: ~4-7x the pairs of any
  // other offset, so its tile range is shared by 1 + extra CTAs
  pl.centre = (subm && (K & 1)) ? K / 2 : -1;
  pl.extra = pl.centre >= 0 ? 4 : 0;
  static int target = 0;  // B200SCN_WGG_CTAS: CTAs aimed at (default: one per SM -- more CTAs mean more partial blocks to write and fold)
  if (target == 0) {
    const char* e = getenv("B200SCN_WGG_CTAS");
    target = e ? atoi(e) : sm_count();
    if (target < 1) target = 1;
  }
  const int64_t n_live = (n_expect > 0 && n_expect < n_cap) ? n_expect : n_cap;
  const int64_t tiles = cdiv(n_live < 1 ? 1 : n_live, tile_rows);
  int64_t parts = cdiv(target, (int64_t)(K + pl.extra) * pl.ca_chunks * pl.cb_chunks);
  if (parts > tiles) parts = tiles;
  if (parts > 64) parts = 64;
  if (parts < 1) parts = 1;
  pl.parts = (int)parts;
  return pl;
}

template <int MT, int NT, bool X3>
int wgg_launch(const WgGplParams& p, const WggPlan& pl, cudaStream_t st) {
  dim3 grid((unsigned)(p.K + p.extra), (unsigned)pl.parts, (unsigned)(pl.ca_chunks * pl.cb_chunks));
  launch_pdl(k_wgrad_gpl<MT, NT, X3>, grid, dim3(kWarps * 32), (size_t)16 * MT * 8 * NT * 4, st, p);
  return 0;
}

template <int MT, bool X3>
int wgg_launch_n(const WgGplParams& p, const WggPlan& pl, cudaStream_t st) {
  return pl.NT == 4 ? wgg_launch<MT, 4, X3>(p, pl, st) : wgg_launch<MT, 2, X3>(p, pl, st);
}

template <bool X3>
int wgg_launch_m(const WgGplParams& p, const WggPlan& pl, cudaStream_t st) {
  switch (pl.MT) {
    case 1: return wgg_launch_n<1, X3>(p, pl, st);
    case 2: return wgg_launch_n<2, X3>(p, pl, st);
    case 3: return wgg_launch_n<3, X3>(p, pl, st);
    default: return wgg_launch_n<4, X3>(p, pl, st);
  }
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

extern "C" int64_t b200scn_wgrad_gpl_scratch_bytes(int K, int Ca, int Cb) {
  return (int64_t)64 * (K + 4) * Ca * Cb * 4 + 256;
}

extern "C" int b200scn_wgrad_gpl(const float* A, int64_t lda, const float* B, int64_t ldb, int tile_rows,
                                 const void* g_rec, const int32_t* koff, int K, int subm, int64_t n_cap,
                                 const int32_t* n_dev, int64_t n_expect, int Ca, int Cb, float* dW, int accumulate,
                                 int math_mode, void* scratch, int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(Ca % 16 == 0 && Cb % 16 == 0 && Ca >= 16 && Cb >= 16, "wgrad_gpl: channels must be multiples of 16");
  B200_CHECK_ARG(math_mode == B200SCN_MATH_TF32 || math_mode == B200SCN_MATH_TF32X3, "wgrad_gpl: tensor-core math modes only");
  B200_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(B) & 15) == 0, "wgrad_gpl: rows must be 16-byte aligned");
  B200_CHECK_ARG(tile_rows == 32 || tile_rows == 64 || tile_rows == 128, "wgrad_gpl: tile_rows must be 32, 64 or 128");
  B200_CHECK_ARG(K >= 1 && K <= 64, "wgrad_gpl: K must be 1..64");
  B200_CHECK_ARG(scratch_bytes >= b200scn_wgrad_gpl_scratch_bytes(K, Ca, Cb), "wgrad_gpl: scratch too small");
  const int64_t cc = (int64_t)Ca * Cb;
  if (n_cap == 0) {
    if (!accumulate) B200_CUDA(cudaMemsetAsync(dW, 0, (size_t)K * cc * 4, st));
    return 0;
  }
  WggPlan pl = wgg_plan(K, Ca, Cb, n_cap, n_expect, tile_rows, subm);
  WgGplParams p;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.g_rec = reinterpret_cast<const int4*>(g_rec); p.koff = koff; p.K = K;
  p.tile_rows = tile_rows; p.centre = pl.centre; p.extra = pl.extra; p.n_cap = n_cap; p.n_dev = n_dev; p.Ca = Ca; p.Cb = Cb;
  p.ca_chunks = pl.ca_chunks; p.cb_chunks = pl.cb_chunks; p.parts = pl.parts; p.part = reinterpret_cast<float*>(scratch);
  p.ones = 0xffffffffu;
  const int rc = math_mode == B200SCN_MATH_TF32X3 ? wgg_launch_m<true>(p, pl, st) : wgg_launch_m<false>(p, pl, st);
  if (rc) return rc;
  int64_t gsz = cdiv((int64_t)K * cc, 256);
  if (gsz > 1024) gsz = 1024;
  launch_pdl(k_wgrad_gpl_fold, dim3((unsigned)gsz), dim3(256), 0, st, reinterpret_cast<const float*>(scratch), K,
             pl.extra, pl.centre, cc, pl.parts, tile_rows, n_cap, n_dev, dW, accumulate);
  count_launches(2);
  B200_CHECK_LAUNCH("wgrad_gpl");
  return 0;
}
