// Grouped pair lists (GPL): the rulebook in the form the tensor-core convolution consumes.
//
// profiles/README.md section 2: in k_conv_pairs a third of the issue slots go to listing the rule
// pairs of every (32-row tile, offset) -- work that depends only on the rulebook and is redone by
// each of the ~16 forward/dgrad uses of a U-Net level per training step.  Here that listing is
// done ONCE per level by the rulebook builder and kept in HBM:
//     tile      = 64 consecutive output rows
//     group     = 8 rule pairs of one kernel offset inside one tile (last group of an offset padded)
//     g_in  [G][8] int32   input row of every pair (padding slots: row 0, routed to a dummy output row)
//     g_meta[G]    int4    {out rows 0..3, out rows 4..7 (local, 1 byte each; 64 = dummy), offset k, 0}
//     tile_off[T+1] int32  group range of every tile (groups sorted by offset inside a tile)
// 48 B per group, ~2.3 groups per site-offset... in practice ~0.9 groups per site at the finest
// level: ~45 B/site instead of the 108 B/site neighbour table.
//
// k_conv_gpl then is only the dense inner loop: prefetched group records -> LDG.128 row gathers
// straight into the mma B fragment -> mma.sync m16n8k8 tf32 (optionally 3-term split) -> float2
// read-modify-write into a per-warp shared accumulator -> one coalesced store per output row.
// Large inputs: one warp per tile, weights resident in shared memory.  Small inputs: one tile per
// CTA, its warps split the tile's groups and their partial accumulators are summed in warp order.
#include <stdlib.h>

#include "common.cuh"

namespace b200scn {

int scan_i32_public(const int32_t* in, int32_t* out, int64_t n_cap, const int32_t* n_dev, int32_t* total_out,
                    int32_t* end_out, int32_t* bsums, cudaStream_t st);

namespace {

constexpr int kTileRows = 64;
constexpr int kMaxSmem = 220 * 1024;

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- builder -------------------------------------------------------------------------------------
// one warp per tile; FILL == false: count groups per tile; FILL == true: write the records
template <bool FILL>
__global__ void k_gpl_build(const int32_t* __restrict__ table, int64_t ld, const int32_t* __restrict__ parent,
                            const uint8_t* __restrict__ kofs, int K, int64_t n_cap, const int32_t* __restrict__ n_dev,
                            int32_t* __restrict__ tile_cnt, const int32_t* __restrict__ tile_off,
                            int32_t* __restrict__ g_in, int4* __restrict__ g_meta, int64_t cap_groups) {
  const int n = live_rows(n_dev, n_cap);
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int64_t n_tiles = ((int64_t)n + kTileRows - 1) / kTileRows;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int64_t tile = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5); tile < n_tiles;
       tile += (int64_t)gridDim.x * warps) {
    const int base = (int)(tile * kTileRows);
    const int r0 = base + lane, r1 = base + 32 + lane;
    int pk0 = -1, pk1 = -1, pp0 = -1, pp1 = -1;
    if (parent != nullptr) {
      if (r0 < n) { pk0 = kofs[r0]; pp0 = __ldg(parent + r0); }
      if (r1 < n) { pk1 = kofs[r1]; pp1 = __ldg(parent + r1); }
    }
    int64_t grp = FILL ? (int64_t)tile_off[tile] : 0;
    int count = 0;
    for (int k = 0; k < K; ++k) {
      int e0 = -1, e1 = -1;
      if (parent != nullptr) {
        e0 = pk0 == k ? pp0 : -1;
        e1 = pk1 == k ? pp1 : -1;
      } else if (table != nullptr) {
        if (r0 < n) e0 = __ldg(table + (int64_t)k * ld + r0);
        if (r1 < n) e1 = __ldg(table + (int64_t)k * ld + r1);
      } else {
        if (r0 < n) e0 = r0;
        if (r1 < n) e1 = r1;
      }
      const uint32_t m0 = __ballot_sync(0xffffffffu, e0 >= 0), m1 = __ballot_sync(0xffffffffu, e1 >= 0);
      const int c0 = __popc(m0), cnt = c0 + __popc(m1);
      const int ng = (cnt + 7) >> 3;
      if (FILL && ng > 0 && grp + ng <= cap_groups) {
        // slot of this lane's pairs inside the offset's padded block of ng*8 slots
        int32_t* gin = g_in + grp * 8;
        unsigned char* gout = reinterpret_cast<unsigned char*>(g_meta + grp);  // 16 B per group: 8 out bytes first
        if (e0 >= 0) {
          const int s = __popc(m0 & lt_mask);
          gin[s] = e0;
          gout[(s >> 3) * 16 + (s & 7)] = (unsigned char)lane;
        }
        if (e1 >= 0) {
          const int s = c0 + __popc(m1 & lt_mask);
          gin[s] = e1;
          gout[(s >> 3) * 16 + (s & 7)] = (unsigned char)(32 + lane);
        }
        const int pad = ng * 8 - cnt;
        if (lane < pad) {
          const int s = cnt + lane;
          gin[s] = 0;  // any valid row: the slot's products only ever reach the dummy accumulator row
          gout[(s >> 3) * 16 + (s & 7)] = (unsigned char)kTileRows;  // dummy accumulator row
        }
        if (lane < ng) {
          int* meta = reinterpret_cast<int*>(g_meta + grp + lane);
          meta[2] = k;
          meta[3] = 0;
        }
      }
      grp += ng;
      count += ng;
    }
    if (!FILL && lane == 0) tile_cnt[tile] = count;
  }
}

// ---- convolution over the lists ------------------------------------------------------------------
struct GplParams {
  const float* in; int ldi;
  const float* W; int w_sk, w_si, w_so; int kflip;
  const int32_t* tile_off; const int32_t* g_in; const int4* g_meta;
  int K;
  int64_t n_cap; const int32_t* n_dev;
  float* out; int ldo; int accumulate; int x3;
  int warps, split;  // split != 0: the CTA's warps share one tile
};

template <int Q, int MT, bool WSMEM, bool X3>
__global__ void __launch_bounds__(WSMEM ? 768 : 256, WSMEM ? 1 : 2) k_conv_gpl(const GplParams p) {
  constexpr int R = kTileRows, KS = 2 * Q, CT = 16 * MT, AS = CT + 8;
  constexpr int DEPTH = Q == 1 ? 4 : 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = live_rows(p.n_dev, p.n_cap);
  const int warps = p.warps;
  const int split = p.split;
  if (split ? ((int64_t)blockIdx.x * R >= n) : ((int64_t)blockIdx.x * warps * R >= n)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = p.K;
  const int cob = blockIdx.y * CT;

  const int w_elems = WSMEM ? K * KS * MT * 128 : 0;
  float* Whi = reinterpret_cast<float*>(smem_raw);
  float* Wlo = Whi + w_elems;
  constexpr int kAccBytes = (R + 1) * AS * 4;
  const int warp_off = w_elems * 4 * (X3 ? 2 : 1) + warp * kAccBytes;
  float* const acc = reinterpret_cast<float*>(smem_raw + warp_off);
  const bool w_vec2 = ((reinterpret_cast<uintptr_t>(p.W) & 7) == 0) && (p.w_sk % 2 == 0) &&
                      (p.w_si % 2 == 0 || p.w_si == 1) && (p.w_so % 2 == 0 || p.w_so == 1);

  if (WSMEM) {
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
      const int mt = col >> 4, c = col & 15, gg = c >> 1;
      const int r = ((rem & 1) << 1) | (c & 1);
      const int dst = ((((k * KS + s) * MT + mt) * 32) + gg * 4 + tt) * 4 + r;
      const float hi = tf32_hi(v);
      Whi[dst] = hi;
      if (X3) Wlo[dst] = v - hi;
    }
    __syncthreads();
  }

  const int64_t n_tiles = ((int64_t)n + R - 1) / R;
  const int64_t wt0 = split ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * warps + warp;
  const int64_t wt_stride = split ? (int64_t)gridDim.x : (int64_t)gridDim.x * warps;
  for (int64_t wt = wt0; wt < n_tiles; wt += wt_stride) {
    const int base = (int)(wt * R);
    __syncwarp();
    for (int e = lane; e < (R + 1) * AS / 4; e += 32)
      reinterpret_cast<float4*>(acc)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    // this warp's group range
    int gb = __ldg(p.tile_off + wt), ge = __ldg(p.tile_off + wt + 1);
    if (split) {
      const int per = (ge - gb + warps - 1) / warps;
      const int b2 = gb + warp * per;
      ge = min(ge, b2 + per);
      gb = min(b2, ge);
    }
    __syncwarp();

    // two-level prefetch: group records (row index, meta) travel 2*DEPTH groups ahead of the MMAs,
    // the feature rows they point to DEPTH groups ahead -- neither load sits on the critical path
    float4 G[DEPTH][Q];
    int4 M[DEPTH], Mn[DEPTH];
    int In[DEPTH];
    // (records and rows past the end of the range are clamped to the last group: redundant but
    // unconditional loads, never consumed)
    const int g_last = max(ge - 1, gb);
    const int32_t* gin_l = p.g_in + g;
    const float4* in_l = reinterpret_cast<const float4*>(p.in + 4 * t);
    const int ld4 = p.ldi >> 2;  // row stride in float4 (ldi % 4 == 0)
    auto fetch_rec = [&](int grp, int d) {
      const int gc = min(grp, g_last);
      In[d] = __ldg(gin_l + (int64_t)gc * 8);
      Mn[d] = __ldg(p.g_meta + gc);
    };
    auto fetch = [&](int grp, int d) {
      const float4* src = in_l + (int64_t)In[d] * ld4;
      M[d] = Mn[d];
#pragma unroll
      for (int q = 0; q < Q; ++q) G[d][q] = __ldg(src + 4 * q);
      fetch_rec(grp + DEPTH, d);
    };
    if (gb < ge) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) fetch_rec(gb + d, d);
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) fetch(gb + d, d);
    }

    float wreg[WSMEM ? 1 : KS][WSMEM ? 1 : MT][4];
    int wreg_k = -1;

    for (int g0 = gb; g0 < ge; g0 += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int grp = g0 + d;
        if (grp < ge) {
          const int kk = M[d].z;
          const uint32_t o_lo = (uint32_t)M[d].x, o_hi = (uint32_t)M[d].y;
          if (!WSMEM && kk != wreg_k) {
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
          }
          const int wk_off = kk * (KS * MT * 128) + lane * 4;
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
              // single-pass tf32: raw fp32 bits (the tensor core ignores the low mantissa bits)
              const float bxh = X3 ? tf32_hi(bx) : bx, byh = X3 ? tf32_hi(by) : by;
              const uint32_t b0 = __float_as_uint(bxh), b1 = __float_as_uint(byh);
              const uint32_t l0 = __float_as_uint(bx - bxh), l1 = __float_as_uint(by - byh);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                uint32_t ah[4], al[4];
                if (WSMEM) {
                  const int off = wk_off + (s * MT + mt) * 128;
                  const float4 a4 = *reinterpret_cast<const float4*>(Whi + off);
                  ah[0] = __float_as_uint(a4.x); ah[1] = __float_as_uint(a4.y);
                  ah[2] = __float_as_uint(a4.z); ah[3] = __float_as_uint(a4.w);
                  if (X3) {
                    const float4 l4 = *reinterpret_cast<const float4*>(Wlo + off);
                    al[0] = __float_as_uint(l4.x); al[1] = __float_as_uint(l4.y);
                    al[2] = __float_as_uint(l4.z); al[3] = __float_as_uint(l4.w);
                  }
                } else {
#pragma unroll
                  for (int r = 0; r < 4; ++r) {
                    const float wv = wreg[WSMEM ? 0 : s][WSMEM ? 0 : mt][r];
                    const float vh = X3 ? tf32_hi(wv) : wv;
                    ah[r] = __float_as_uint(vh);
                    al[r] = __float_as_uint(wv - vh);
                  }
                }
                if (X3) {
                  mma_tf32(dacc[mt], al, b0, b1);
                  mma_tf32(dacc[mt], ah, l0, l1);
                }
                mma_tf32(dacc[mt], ah, b0, b1);
              }
            }
          }
          fetch(grp + DEPTH, d);  // this slot is free again
          // pairs 2t and 2t+1 of the group -> accumulator rows (padding lands in the dummy row)
          const uint32_t ow = t < 2 ? o_lo : o_hi;
          const int o0 = (ow >> (16 * (t & 1))) & 0xff, o1 = (ow >> (16 * (t & 1) + 8)) & 0xff;
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

    __syncwarp();
    if (split) __syncthreads();
    const bool vec_ok = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const int e_begin = split ? tid : lane, e_step = split ? (int)blockDim.x : 32;
    const float* acc0 = reinterpret_cast<const float*>(smem_raw + w_elems * 4 * (X3 ? 2 : 1));
    for (int e = e_begin; e < R * (CT / 4); e += e_step) {
      const int row = e / (CT / 4), c4 = e - row * (CT / 4);
      if (base + row < n) {
        float4 v;
        if (split) {
          v = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int w = 0; w < warps; ++w) {
            const float4 u = *reinterpret_cast<const float4*>(acc0 + w * (kAccBytes / 4) + row * AS + 4 * c4);
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
    if (split) __syncthreads();
  }
}

struct GPlan {
  int Q, MT, passes, warps, split, ctas_per_sm;
  bool wsmem, ok;
  size_t smem;
};

GPlan plan_gpl(int K, int Cin, int Cout, int64_t n_cap, int x3) {
  GPlan pl{};
  pl.ok = false;
  if (Cin % 16 || Cout % 16 || Cin > 64 || Cin < 16) return pl;
  pl.Q = Cin / 16;
  const int cot = Cout / 16;
  const bool large = n_cap >= 65536;
  auto acc_bytes = [](int mt) { return (size_t)(kTileRows + 1) * (16 * mt + 8) * 4; };
  if (large) {
    pl.MT = 1;
    for (int mt = 4; mt >= 1; --mt) {
      if (cot % mt != 0) continue;
      if ((size_t)K * Cin * 16 * mt * 4 * (x3 ? 2 : 1) + 8 * acc_bytes(mt) <= (size_t)kMaxSmem) {
        pl.MT = mt;
        break;
      }
    }
  } else {
    pl.MT = (cot % 2 == 0) ? 2 : 1;
  }
  pl.passes = cot / pl.MT;
  const size_t wbytes = (size_t)K * Cin * 16 * pl.MT * 4 * (x3 ? 2 : 1);
  pl.wsmem = large && (wbytes + 8 * acc_bytes(pl.MT) <= (size_t)kMaxSmem);
  pl.split = large ? 0 : 1;
  const size_t fixed = pl.wsmem ? wbytes : 0;
  static int large_warps = 0;  // B200SCN_GPL_WARPS: warps per CTA of the large-input variant (A/B runs)
  if (large_warps == 0) {
    const char* e = getenv("B200SCN_GPL_WARPS");
    large_warps = e ? atoi(e) : 24;
    if (large_warps < 1 || large_warps > 24) large_warps = 24;
  }
  int warps = large ? large_warps : 8;
  while (warps > 1 && fixed + warps * acc_bytes(pl.MT) > (size_t)kMaxSmem) --warps;
  if (fixed + warps * acc_bytes(pl.MT) > (size_t)kMaxSmem) return pl;
  if (!pl.wsmem && warps > 8) warps = 8;  // launch bound of the direct-weight variant
  pl.warps = warps;
  pl.smem = fixed + warps * acc_bytes(pl.MT);
  int cps = (int)((size_t)(227 * 1024) / (pl.smem + 1024));
  if (cps > 2048 / (warps * 32)) cps = 2048 / (warps * 32);
  if (cps > 8) cps = 8;
  if (cps < 1) cps = 1;
  pl.ctas_per_sm = cps;
  pl.ok = true;
  return pl;
}

template <int Q, int MT, bool WSMEM, bool X3>
int launch_gpl(const GplParams& p, const GPlan& pl, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_conv_gpl<Q, MT, WSMEM, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    configured = true;
  }
  const int64_t tiles = cdiv(p.n_cap, kTileRows);
  int64_t gx = pl.split ? tiles : cdiv(tiles, pl.warps);
  int64_t gmax = WSMEM ? ((int64_t)sm_count() * pl.ctas_per_sm) / pl.passes : (int64_t)sm_count() * pl.ctas_per_sm * 4;
  if (gmax < 1) gmax = 1;
  if (gx > gmax) gx = gmax;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)pl.passes);
  k_conv_gpl<Q, MT, WSMEM, X3><<<grid, pl.warps * 32, pl.smem, st>>>(p);
  count_launches(1);
  B200_CHECK_LAUNCH("conv_gpl");
  return 0;
}

template <int Q, int MT>
int launch_gpl_qm(const GplParams& p, const GPlan& pl, cudaStream_t st) {
  if (pl.wsmem) return p.x3 ? launch_gpl<Q, MT, true, true>(p, pl, st) : launch_gpl<Q, MT, true, false>(p, pl, st);
  return p.x3 ? launch_gpl<Q, MT, false, true>(p, pl, st) : launch_gpl<Q, MT, false, false>(p, pl, st);
}

template <int Q>
int launch_gpl_q(const GplParams& p, const GPlan& pl, cudaStream_t st) {
  switch (pl.MT) {
    case 1: return launch_gpl_qm<Q, 1>(p, pl, st);
    case 2: return launch_gpl_qm<Q, 2>(p, pl, st);
    case 3: return launch_gpl_qm<Q, 3>(p, pl, st);
    default: return launch_gpl_qm<Q, 4>(p, pl, st);
  }
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

extern "C" int b200scn_gpl_tile_rows(void) { return kTileRows; }

/* upper bound of the number of groups for n_cap rows (static regime allocation) */
extern "C" int64_t b200scn_gpl_max_groups(int64_t n_cap, int K) {
  const int64_t tiles = cdiv(n_cap < 1 ? 1 : n_cap, kTileRows);
  return tiles * (int64_t)K * (kTileRows / 8);
}

extern "C" int64_t b200scn_gpl_scratch_bytes(int64_t n_cap) {
  const int64_t tiles = cdiv(n_cap < 1 ? 1 : n_cap, kTileRows);
  return (tiles + 8) * 4 + (cdiv(tiles, 1024) + 16) * 4 + 256;
}

/* stage 0: tile_off[tiles + 1] (exclusive scan of the per-tile group counts) and total_dev[1].
 * stage 1: fill g_in[cap_groups * 8] / g_meta[cap_groups] (int4 records) using tile_off. */
extern "C" int b200scn_gpl_build(int stage, const int32_t* table, int64_t ld, const int32_t* parent,
                                 const uint8_t* kofs, int K, int64_t n_cap, const int32_t* n_dev, int32_t* tile_off,
                                 int32_t* total_dev, int32_t* g_in, void* g_meta, int64_t cap_groups, void* scratch,
                                 int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(K >= 1 && K <= 64, "gpl_build: K must be 1..64");
  B200_CHECK_ARG(scratch_bytes >= b200scn_gpl_scratch_bytes(n_cap), "gpl_build: scratch too small");
  const int64_t tiles = cdiv(n_cap < 1 ? 1 : n_cap, kTileRows);
  int32_t* tile_cnt = reinterpret_cast<int32_t*>(scratch);
  int32_t* bsums = tile_cnt + tiles + 8;
  const int warps = 8;
  int64_t grid = cdiv(tiles, warps);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (stage == 0) {
    B200_CUDA(cudaMemsetAsync(tile_cnt, 0, (size_t)(tiles + 8) * 4, st));
    if (n_cap > 0)
      k_gpl_build<false><<<(int)grid, warps * 32, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, tile_cnt, nullptr,
                                                           nullptr, nullptr, 0);
    // exclusive scan over all `tiles` counts (tiles beyond the live ones hold 0)
    if (int rc = scan_i32_public(tile_cnt, tile_off, tiles, nullptr, total_dev, tile_off, bsums, st)) return rc;
    count_launches(4);
  } else {
    if (n_cap > 0)
      k_gpl_build<true><<<(int)grid, warps * 32, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, nullptr, tile_off,
                                                          g_in, reinterpret_cast<int4*>(g_meta), cap_groups);
    count_launches(1);
  }
  B200_CHECK_LAUNCH("gpl_build");
  return 0;
}

extern "C" int b200scn_conv_gpl(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so,
                                int kflip, const int32_t* tile_off, const int32_t* g_in, const void* g_meta, int K,
                                int64_t n_out_cap, const int32_t* n_out_dev, int Cin, int Cout, float* out, int64_t ldo,
                                int accumulate, int math_mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(Cin % 16 == 0 && Cout % 16 == 0 && Cin >= 16 && Cout >= 16, "conv_gpl: channels must be multiples of 16");
  B200_CHECK_ARG(math_mode == B200SCN_MATH_TF32 || math_mode == B200SCN_MATH_TF32X3, "conv_gpl: tensor-core math modes only");
  B200_CHECK_ARG(ldi % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "conv_gpl: input rows must be 16-byte aligned");
  const int64_t lim = (1ll << 31) - 1;
  B200_CHECK_ARG(ldi < lim && ldo < lim && (int64_t)K * Cin * Cout < lim, "conv_gpl: stride out of 32-bit range");
  if (n_out_cap == 0) return 0;
  const int x3 = math_mode == B200SCN_MATH_TF32X3 ? 1 : 0;
  for (int c0 = 0; c0 < Cin; c0 += 64) {
    const int cc = Cin - c0 < 64 ? Cin - c0 : 64;
    GPlan pl = plan_gpl(K, cc, Cout, n_out_cap, x3);
    B200_CHECK_ARG(pl.ok, "conv_gpl: unsupported shape Cin=%d Cout=%d K=%d", Cin, Cout, K);
    GplParams p;
    p.in = in + c0; p.ldi = (int)ldi; p.W = W + (int64_t)c0 * w_si; p.w_sk = (int)w_sk; p.w_si = (int)w_si;
    p.w_so = (int)w_so; p.kflip = kflip; p.tile_off = tile_off; p.g_in = g_in;
    p.g_meta = reinterpret_cast<const int4*>(g_meta); p.K = K; p.n_cap = n_out_cap; p.n_dev = n_out_dev; p.out = out;
    p.ldo = (int)ldo; p.accumulate = (c0 > 0) ? 1 : accumulate; p.x3 = x3; p.warps = pl.warps; p.split = pl.split;
    int rc;
    switch (pl.Q) {
      case 1: rc = launch_gpl_q<1>(p, pl, st); break;
      case 2: rc = launch_gpl_q<2>(p, pl, st); break;
      case 3: rc = launch_gpl_q<3>(p, pl, st); break;
      default: rc = launch_gpl_q<4>(p, pl, st); break;
    }
    if (rc) return rc;
  }
  return 0;
}
