// Gather-GEMM sparse convolution, "dense pairs" formulation (forward / dgrad of
// SubmanifoldConvolution, Convolution, Deconvolution, NetworkInNetwork) on the tensor pipe.
//
// An output-stationary mma.sync kernel (round 1, removed) paid an MMA for every (offset, 8-site tile)
// that has at least one neighbour -- at the fine U-Net levels only ~2.4 of the 8 columns of
// such an MMA carry data and the bookkeeping around it dominates the issue slots (ncu:
// profiles/README.md).  Here each warp owns R = 32*NCH consecutive output rows and, for every
// kernel offset k in order,
//   1. compacts the valid (input row, local output row) rule pairs of its rows from the
//      offset-major table with ballot + popc (deterministic, row order),
//   2. gathers the input rows of 8 pairs at a time straight into registers (one LDG.128 per
//      lane per 16 channels; a lane's registers are exactly its mma B fragment),
//   3. multiplies the dense 8-pair tile with W_k on mma.sync.m16n8k8 (tf32, optional 3-term
//      split for fp32-equivalent accuracy), and
//   4. adds the result into a per-warp shared-memory accumulator [R][Cout tile] -- plain
//      read-modify-write, because within one offset every output row appears at most once
//       and offsets are processed in order by the same warp.
// The accumulator is written to HBM once, coalesced.  Gathers of offset k+1 are issued before
// the MMAs of offset k (register ping-pong); table entries are prefetched three offsets ahead.
// Weights come either from shared memory (fragment order, pre-split hi/lo; large inputs) or
// straight from L1/L2 (small inputs, where a per-CTA weight fill would dominate).
#include <stdlib.h>

#include "common.cuh"

namespace b200scn {
namespace {

constexpr int kMaxSmem = 220 * 1024;

// strides are 32-bit on the device (checked on the host): address math stays in IMAD.WIDE form
struct PairParams {
  const float* in; int ldi;
  const float* W; int w_sk, w_si, w_so; int kflip;
  const int32_t* table; int ld;
  const int32_t* parent; const uint8_t* kofs;
  int K;
  int64_t n_cap; const int32_t* n_dev;
  float* out; int ldo; int accumulate; int x3;
  int warps;
  int ksplit;  // > 1: the CTA's warps split the kernel offsets of ONE row tile (small inputs)
};

// tf32 "hi" part by truncation (one LOP3; cvt.rna.tf32 expands to ~4 ALU ops on sm_100).  With
// the 3-term split the remainder v - hi is exact, so truncation costs no accuracy; in the
// single-pass mode it is the tensor core's own operand truncation made explicit.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

constexpr int kKB = 9;  // kernel offsets whose rule pairs are listed together (one "batch")

// Q = Cin/16, MT = Cout tile/16, NCH = 32-row chunks per warp, WSMEM = weights in smem
template <int Q, int MT, int NCH, bool WSMEM>
__global__ void __launch_bounds__(NCH == 2 ? 512 : (WSMEM ? 1024 : 288), (NCH == 1 && !WSMEM) ? 2 : 1) k_conv_pairs(const PairParams p) {
  constexpr int R = 32 * NCH, KS = 2 * Q, CT = 16 * MT, AS = CT + 8;
  constexpr int CAP = kKB * R;                // list slots per batch (every offset padded to a multiple of 8)
  constexpr int DEPTH = Q == 1 ? 4 : 2;       // 8-pair groups whose gathers are in flight
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = live_rows(p.n_dev, p.n_cap);
  const int warps = p.warps;
  const int ksplit = p.ksplit;
  if (ksplit > 1 ? ((int64_t)blockIdx.x * R >= n) : ((int64_t)blockIdx.x * warps * R >= n)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = p.K;
  const int cob = blockIdx.y * CT;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // 8-byte weight loads are legal when the strides keep every (ci, co) pair start 8-byte aligned
  const bool w_vec2 = ((reinterpret_cast<uintptr_t>(p.W) & 7) == 0) && (p.w_sk % 2 == 0) && (p.w_si % 2 == 0 || p.w_si == 1) &&
                      (p.w_so % 2 == 0 || p.w_so == 1);

  // ---- shared memory carve-up ----------------------------------------------------------------
  const int w_elems = WSMEM ? K * KS * MT * 128 : 0;
  float* Whi = reinterpret_cast<float*>(smem_raw);
  float* Wlo = Whi + w_elems;  // only filled / read when p.x3
  unsigned char* wbase = smem_raw + (size_t)w_elems * 4 * (p.x3 ? 2 : 1);
  constexpr size_t kAccBytes = (size_t)(R + 1) * AS * 4;  // + 1 dummy row that absorbs the padding slots
  constexpr size_t kListBytes = (size_t)CAP * 4 + (size_t)CAP + (size_t)CAP / 8;
  constexpr size_t kPerWarp = (kAccBytes + kListBytes + 15) & ~(size_t)15;
  // per-warp regions addressed as 32-bit offsets from the shared window (no 64-bit generic pointers)
  const int warp_off = (int)((size_t)w_elems * 4 * (p.x3 ? 2 : 1)) + warp * (int)kPerWarp;  // bytes
  float* const acc = reinterpret_cast<float*>(smem_raw + warp_off);
  int32_t* const in_s = reinterpret_cast<int32_t*>(smem_raw + warp_off + (int)kAccBytes);
  uint8_t* const out_s = smem_raw + warp_off + (int)kAccBytes + CAP * 4;
  uint8_t* const gk_s = smem_raw + warp_off + (int)kAccBytes + CAP * 4 + CAP;

  if (WSMEM) {
    // coalesced read along the contiguous weight dimension, scatter into A-fragment order
    const bool co_fast = (p.w_so == 1);
    const int fast_n = co_fast ? CT : 16 * Q, slow_n = co_fast ? 16 * Q : CT;
    const int total = K * slow_n * fast_n;
    for (int e = tid; e < total; e += blockDim.x) {
      const int f = e % fast_n;
      int rest = e / fast_n;
      const int sl = rest % slow_n;
      const int k = rest / slow_n;
      const int ci = co_fast ? sl : f, col = co_fast ? f : sl;
      const int kk = p.kflip ? (K - 1 - k) : k;
      const float v = __ldg(p.W + (int64_t)kk * p.w_sk + (int64_t)ci * p.w_si + (int64_t)(cob + col) * p.w_so);
      const int c16 = ci & 15, tt = c16 >> 2, rem = c16 & 3;
      const int s = ((ci >> 4) << 1) + (rem >> 1);
      // mma row g of an output tile holds channel 2g, row g+8 channel 2g+1 (adjacent channels per
      // lane: 8-byte weight loads and float2 accumulator updates)
      const int mt = col >> 4, c = col & 15, gg = c >> 1;
      const int r = ((rem & 1) << 1) | (c & 1);
      const int dst = ((((k * KS + s) * MT + mt) * 32) + gg * 4 + tt) * 4 + r;
      const float hi = tf32_hi(v);
      Whi[dst] = hi;
      if (p.x3) Wlo[dst] = v - hi;
    }
    __syncthreads();
  }

  const int64_t n_tiles = ((int64_t)n + R - 1) / R;
  // ksplit == 1: every warp owns its own row tiles.  ksplit > 1: all warps of the CTA work on the
  // same tile, warp w taking offsets w, w + ksplit, ... into a private accumulator; the partial
  // accumulators are summed in fixed warp order at the end (deterministic).
  const int kw = ksplit > 1 ? warp : 0;
  const int kstep = ksplit > 1 ? ksplit : 1;
  const int64_t wt0 = ksplit > 1 ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * warps + warp;
  const int64_t wt_stride = ksplit > 1 ? (int64_t)gridDim.x : (int64_t)gridDim.x * warps;
  const int Kl = kw < K ? (K - kw + kstep - 1) / kstep : 0;  // this warp's offsets: kw + j * kstep

  for (int64_t wt = wt0; wt < n_tiles; wt += wt_stride) {
    const int base = (int)(wt * R);
    __syncwarp();
    for (int e = lane; e < (R + 1) * AS / 4; e += 32)
      reinterpret_cast<float4*>(acc)[e] = make_float4(0.f, 0.f, 0.f, 0.f);

    int pk[NCH], pp[NCH];
    if (p.parent != nullptr) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int row = base + 32 * c + lane;
        pk[c] = -1; pp[c] = -1;
        if (row < n) { pk[c] = p.kofs[row]; pp[c] = __ldg(p.parent + row); }
      }
    }

    int ev[kKB][NCH];
    bool row_ok[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) row_ok[c] = base + 32 * c + lane < n;
    const int32_t* tptr = p.table != nullptr ? p.table + (int64_t)kw * p.ld + base + lane : nullptr;
    const int tstep = kstep * p.ld;  // elements between this warp's consecutive offsets
    auto load_batch = [&](int j0) {
      const int nb = min(kKB, Kl - j0);
#pragma unroll
      for (int jj = 0; jj < kKB; ++jj) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          int e = -1;
          if (jj < nb && row_ok[c]) {
            if (p.parent != nullptr) e = (pk[c] == kw + (j0 + jj) * kstep) ? pp[c] : -1;
            else if (tptr != nullptr) e = __ldg(tptr + 32 * c);
            else e = base + 32 * c + lane;
          }
          ev[jj][c] = e;
        }
        if (tptr != nullptr) tptr += tstep;
      }
    };
    load_batch(0);
    for (int j0 = 0; j0 < Kl; j0 += kKB) {
      // ---- phase 1: list the rule pairs of up to kKB offsets, each offset padded to 8 slots ----
      __syncwarp();  // previous batch's readers of the lists are done
      int total = 0;
      const int nb_cur = min(kKB, Kl - j0);
#pragma unroll
      for (int jj = 0; jj < kKB; ++jj) {
        if (jj < nb_cur) {
          const int g_first = total >> 3;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const uint32_t m = __ballot_sync(0xffffffffu, ev[jj][c] >= 0);
            if (ev[jj][c] >= 0) {
              const int pos = total + __popc(m & lt_mask);
              in_s[pos] = ev[jj][c];
              out_s[pos] = (uint8_t)(32 * c + lane);
            }
            total += __popc(m);
          }
          const int padn = (8 - (total & 7)) & 7;
          if (lane < padn) {
            in_s[total + lane] = -1;
            out_s[total + lane] = (uint8_t)R;  // dummy accumulator row
          }
          total += padn;
          const int g_new = (total >> 3) - g_first;
          if (lane < g_new) gk_s[g_first + lane] = (uint8_t)(kw + (j0 + jj) * kstep);
        }
      }
      __syncwarp();
      const int ng = total >> 3;
      load_batch(j0 + kKB);  // next batch's table entries travel while this batch is multiplied

      // ---- phase 2: uniform loop over dense 8-pair groups ------------------------------------
      float4 G[DEPTH][Q];
      // direct-weight path: all A fragments of the current offset live in registers, fetched as one
      // batch of independent loads when the offset changes (one L2 round trip per offset, not per k-step)
      float wreg[WSMEM ? 1 : KS][WSMEM ? 1 : MT][4];
      int wreg_k = -1;
      auto load_w = [&](int kk) {
        if (WSMEM) return;
        const int kwt = p.kflip ? (K - 1 - kk) : kk;
#pragma unroll
        for (int s = 0; s < (WSMEM ? 1 : KS); ++s) {
#pragma unroll
          for (int mt = 0; mt < (WSMEM ? 1 : MT); ++mt) {
            const float* wp = p.W + (int64_t)kwt * p.w_sk + (int64_t)(16 * (s >> 1) + 4 * t + 2 * (s & 1)) * p.w_si +
                              (int64_t)(cob + 16 * mt + 2 * g) * p.w_so;
            if (p.w_so == 1 && w_vec2) {
              const float2 u0 = __ldg(reinterpret_cast<const float2*>(wp));
              const float2 u1 = __ldg(reinterpret_cast<const float2*>(wp + p.w_si));
              wreg[s][mt][0] = u0.x; wreg[s][mt][1] = u0.y; wreg[s][mt][2] = u1.x; wreg[s][mt][3] = u1.y;
            } else if (p.w_si == 1 && w_vec2) {
              const float2 u0 = __ldg(reinterpret_cast<const float2*>(wp));
              const float2 u1 = __ldg(reinterpret_cast<const float2*>(wp + p.w_so));
              wreg[s][mt][0] = u0.x; wreg[s][mt][2] = u0.y; wreg[s][mt][1] = u1.x; wreg[s][mt][3] = u1.y;
            } else {
#pragma unroll
              for (int r = 0; r < 4; ++r) wreg[s][mt][r] = __ldg(wp + (r >> 1) * p.w_si + (r & 1) * p.w_so);
            }
          }
        }
        wreg_k = kk;
      };
      auto gather = [&](int grp, float4 (&dst)[Q]) {
        const int idx = grp < ng ? in_s[8 * grp + g] : -1;
        const float* src = p.in + (int64_t)(idx >= 0 ? idx : 0) * p.ldi + 4 * t;
#pragma unroll
        for (int q = 0; q < Q; ++q) dst[q] = idx >= 0 ? ldg_f4(src + 16 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) gather(d, G[d]);

      for (int g0 = 0; g0 < ng; g0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
          const int grp = g0 + d;
          if (grp < ng) {
            const int kk = gk_s[grp];
            if (!WSMEM && kk != wreg_k) load_w(kk);
            float dacc[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int r = 0; r < 4; ++r) dacc[mt][r] = 0.f;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
              const float bv[4] = {G[d][q].x, G[d][q].y, G[d][q].z, G[d][q].w};
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int s = 2 * q + h;
                const float bx = bv[2 * h], by = bv[2 * h + 1];
                const float bxh = tf32_hi(bx), byh = tf32_hi(by);
                const uint32_t b0 = __float_as_uint(bxh), b1 = __float_as_uint(byh);
                const uint32_t l0 = __float_as_uint(bx - bxh), l1 = __float_as_uint(by - byh);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  uint32_t ah[4], al[4];
                  if (WSMEM) {
                    const int off = (((kk * KS + s) * MT + mt) * 32 + lane) * 4;
                    const float4 a4 = *reinterpret_cast<const float4*>(Whi + off);
                    ah[0] = __float_as_uint(a4.x); ah[1] = __float_as_uint(a4.y);
                    ah[2] = __float_as_uint(a4.z); ah[3] = __float_as_uint(a4.w);
                    if (p.x3) {
                      const float4 l4 = *reinterpret_cast<const float4*>(Wlo + off);
                      al[0] = __float_as_uint(l4.x); al[1] = __float_as_uint(l4.y);
                      al[2] = __float_as_uint(l4.z); al[3] = __float_as_uint(l4.w);
                    }
                  } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                      const float wv = wreg[WSMEM ? 0 : s][WSMEM ? 0 : mt][r];
                      const float vh = tf32_hi(wv);
                      ah[r] = __float_as_uint(vh);
                      al[r] = __float_as_uint(wv - vh);
                    }
                  }
                  if (p.x3) {
                    mma_tf32(dacc[mt], al, b0, b1);
                    mma_tf32(dacc[mt], ah, l0, l1);
                  }
                  mma_tf32(dacc[mt], ah, b0, b1);
                }
              }
            }
            // this group's buffer is free: fetch the group DEPTH ahead
            gather(grp + DEPTH, G[d]);
            // D^T fragment -> accumulator rows (padding slots land in the dummy row)
            const int o0 = out_s[8 * grp + 2 * t], o1 = out_s[8 * grp + 2 * t + 1];
            float2* a0 = reinterpret_cast<float2*>(acc + o0 * AS + 2 * g);
            float2* a1 = reinterpret_cast<float2*>(acc + o1 * AS + 2 * g);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              float2 u = a0[8 * mt], v = a1[8 * mt];
              u.x += dacc[mt][0]; u.y += dacc[mt][2];
              v.x += dacc[mt][1]; v.y += dacc[mt][3];
              a0[8 * mt] = u;
              a1[8 * mt] = v;
            }
            __syncwarp();
          }
        }
      }
    }

    // ---- epilogue: accumulator(s) -> out, coalesced ------------------------------------------------
    __syncwarp();
    if (ksplit > 1) __syncthreads();
    const bool vec_ok = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const int e_begin = ksplit > 1 ? tid : lane, e_step = ksplit > 1 ? (int)blockDim.x : 32;
    for (int e = e_begin; e < R * (CT / 4); e += e_step) {
      const int row = e / (CT / 4), c4 = e - row * (CT / 4);
      if (base + row < n) {
        float4 v;
        if (ksplit > 1) {
          v = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int w = 0; w < ksplit; ++w) {
            const float4 u = *reinterpret_cast<const float4*>(
                reinterpret_cast<const float*>(wbase + (size_t)w * kPerWarp) + row * AS + 4 * c4);
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
        } else {
          v = *reinterpret_cast<const float4*>(acc + row * AS + 4 * c4);
        }
        float* o = p.out + (int64_t)(base + row) * p.ldo + cob + 4 * c4;
        if (vec_ok) {
          if (p.accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(o);
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
          }
          *reinterpret_cast<float4*>(o) = v;
        } else {
          if (p.accumulate) { v.x += o[0]; v.y += o[1]; v.z += o[2]; v.w += o[3]; }
          o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
      }
    }
    if (ksplit > 1) __syncthreads();  // partial accumulators are re-zeroed by the next tile
  }
}

struct PPlan {
  int Q, MT, NCH, passes, warps, ctas_per_sm, ksplit;
  bool wsmem, ok;
  size_t smem;
};

size_t per_warp_bytes(int MT, int NCH) {
  const int R = 32 * NCH, AS = 16 * MT + 8, CAP = kKB * R;
  size_t b = (size_t)(R + 1) * AS * 4 + (size_t)CAP * 4 + (size_t)CAP + (size_t)CAP / 8;
  return (b + 15) & ~(size_t)15;
}

// Cin here is already a chunk (<= 64, multiple of 16)
PPlan plan_pairs(int K, int Cin, int Cout, int64_t n_cap, int x3) {
  PPlan pl{};
  pl.ok = false;
  if (Cin % 16 || Cout % 16 || Cin > 64 || Cin < 16) return pl;
  pl.Q = Cin / 16;
  const int cot = Cout / 16;
  const bool large = n_cap >= 65536;
  if (large) {
    // widest output tile whose weights (all K offsets, hi [+ lo]) stay resident next to >= 8 warps
    pl.MT = 1;
    for (int mt = 4; mt >= 1; --mt) {
      if (cot % mt != 0) continue;
      const int nch = pl.Q * mt > 2 ? 2 : 1;
      if ((size_t)K * Cin * 16 * mt * 4 * (x3 ? 2 : 1) + 8 * per_warp_bytes(mt, nch) <= (size_t)kMaxSmem) {
        pl.MT = mt;
        break;
      }
    }
  } else {
    pl.MT = (cot % 2 == 0) ? 2 : 1;  // more, smaller tasks when rows are few
  }
  pl.NCH = (large && pl.Q * pl.MT > 2) ? 2 : 1;
  pl.passes = cot / pl.MT;
  const size_t pw = per_warp_bytes(pl.MT, pl.NCH);
  const size_t wbytes = (size_t)K * Cin * 16 * pl.MT * 4 * (x3 ? 2 : 1);
  pl.wsmem = large && (wbytes + 8 * pw <= (size_t)kMaxSmem);
  const size_t fixed = pl.wsmem ? wbytes : 0;
  // small inputs: one row tile per CTA, its warps split the offsets (27 -> 9 warps x 3, 8 -> 4 x 2, ...)
  pl.ksplit = large ? 1 : (K >= 27 ? 9 : (K >= 8 ? 4 : (K >= 4 ? 2 : 1)));
  int warps = large ? (pl.NCH == 1 ? 32 : 16) : (pl.ksplit > 1 ? pl.ksplit : 4);
  while (pl.ksplit == 1 && warps > 2 && fixed + warps * pw > (size_t)kMaxSmem) warps >>= 1;
  if (fixed + warps * pw > (size_t)kMaxSmem) return pl;
  pl.warps = warps;
  pl.smem = fixed + warps * pw;
  int cps = (int)((size_t)(227 * 1024) / (pl.smem + 1024));
  const int by_threads = 2048 / (warps * 32);
  if (cps > by_threads) cps = by_threads;
  if (cps > 8) cps = 8;
  if (cps < 1) cps = 1;
  pl.ctas_per_sm = cps;
  pl.ok = true;
  return pl;
}

template <int Q, int MT, int NCH, bool WSMEM>
int launch_pairs(const PairParams& p, const PPlan& pl, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_conv_pairs<Q, MT, NCH, WSMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kMaxSmem));
    configured = true;
  }
  const int R = 32 * NCH;
  const int64_t tiles = cdiv(p.n_cap, R);
  int64_t gx = pl.ksplit > 1 ? tiles : cdiv(tiles, pl.warps);
  int64_t gmax = ((int64_t)sm_count() * pl.ctas_per_sm) / pl.passes;
  if (!WSMEM) gmax = (int64_t)sm_count() * pl.ctas_per_sm;  // no per-CTA prologue: no reason to hold CTAs back
  if (gmax < 1) gmax = 1;
  if (gx > gmax) gx = gmax;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)pl.passes);
  k_conv_pairs<Q, MT, NCH, WSMEM><<<grid, pl.warps * 32, pl.smem, st>>>(p);
  count_launches(1);
  B200_CHECK_LAUNCH("conv_pairs");
  return 0;
}

template <int Q, int MT>
int launch_qm(const PairParams& p, const PPlan& pl, cudaStream_t st) {
  if (pl.NCH == 2) {
    if (pl.wsmem) return launch_pairs<Q, MT, 2, true>(p, pl, st);
    return launch_pairs<Q, MT, 2, false>(p, pl, st);
  }
  if (pl.wsmem) return launch_pairs<Q, MT, 1, true>(p, pl, st);
  return launch_pairs<Q, MT, 1, false>(p, pl, st);
}

template <int Q>
int launch_q(const PairParams& p, const PPlan& pl, cudaStream_t st) {
  switch (pl.MT) {
    case 1: return launch_qm<Q, 1>(p, pl, st);
    case 2: return launch_qm<Q, 2>(p, pl, st);
    case 3: return launch_qm<Q, 3>(p, pl, st);
    default: return launch_qm<Q, 4>(p, pl, st);
  }
}

}  // namespace

bool conv_pairs_supported(int Cin, int Cout, int K, int64_t ldi, const void* in, int64_t n_cap, int64_t ld,
                          int64_t ldo, int64_t n_in_max) {
  if (Cin % 16 != 0 || Cout % 16 != 0 || K < 1 || K > 64) return false;
  // 32-bit element offsets on the device
  const int64_t lim = (1ll << 31) - 1;
  (void)n_in_max; (void)n_cap;
  if (ldi >= lim || ldo >= lim || (int64_t)K * ld >= lim || (int64_t)K * Cin * Cout >= lim) return false;
  if (ldi % 4 != 0 || (reinterpret_cast<uintptr_t>(in) & 15) != 0) return false;
  return true;
}

int conv_rows_pairs(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so, int kflip,
                    const int32_t* table, int64_t ld, const int32_t* parent, const uint8_t* kofs, int K, int64_t n_cap,
                    const int32_t* n_dev, int Cin, int Cout, float* out, int64_t ldo, int accumulate, int math_mode,
                    cudaStream_t st) {
  if (n_cap == 0) return 0;
  const int x3 = (math_mode == B200SCN_MATH_TF32X3) ? 1 : 0;
  // input channels in chunks of <= 64 (register budget of the gather ping-pong); later chunks accumulate
  for (int c0 = 0; c0 < Cin; c0 += 64) {
    const int cc = Cin - c0 < 64 ? Cin - c0 : 64;
    PPlan pl = plan_pairs(K, cc, Cout, n_cap, x3);
    if (!pl.ok) {
      set_error("conv_pairs: shape not supported (Cin=%d Cout=%d K=%d)", Cin, Cout, K);
      return B200SCN_ERR_UNSUPPORTED;
    }
    PairParams p;
    p.in = in + c0; p.ldi = (int)ldi; p.W = W + (int64_t)c0 * w_si; p.w_sk = (int)w_sk; p.w_si = (int)w_si;
    p.w_so = (int)w_so; p.kflip = kflip; p.table = table; p.ld = (int)ld; p.parent = parent; p.kofs = kofs; p.K = K;
    p.n_cap = n_cap; p.n_dev = n_dev; p.out = out; p.ldo = (int)ldo; p.accumulate = (c0 > 0) ? 1 : accumulate;
    p.x3 = x3;
    p.warps = pl.warps;
    p.ksplit = pl.ksplit;
    int rc;
    switch (pl.Q) {
      case 1: rc = launch_q<1>(p, pl, st); break;
      case 2: rc = launch_q<2>(p, pl, st); break;
      case 3: rc = launch_q<3>(p, pl, st); break;
      default: rc = launch_q<4>(p, pl, st); break;
    }
    if (rc) return rc;
  }
  return 0;
}

}  // namespace b200scn
