// Grouped pair lists (GPL): the rulebook in the form the tensor-core convolution consumes, and the
// convolution kernel over them.
//
// profiles/README.md section 2: listing the rule pairs of every (tile, offset) inside the conv kernel
// cost a third of its issue slots -- work that depends only on the rulebook and is redone by each of
// the ~16 forward/dgrad uses of a U-Net level per training step.  Here the listing is done ONCE per
// level by the rulebook builder and kept in HBM:
//     tile      = R consecutive output rows (R = 128, or 64 for levels with few rows)
//     group     = 16 rule pairs of one kernel offset inside one tile (last group of an offset padded)
//     g_rec[G][8] int4   entry j of a group describes its pairs j and j+8:
//                        { in row (j), in row (j+8), out(j) | out(j+8) << 8 | k << 16, 0 }
//                        (out = row inside the tile, R = dummy row of a padding slot, whose in row is 0)
//     tile_off[T+1] int32  group range of every tile (groups of a tile are sorted by offset)
// 128 B per group, ~0.35 groups per site at the finest level: ~45 B/site instead of the 108 B/site
// neighbour table, and one LDG.128 per lane hands the kernel everything it needs to know about a group.
//
// k_conv_gpl is then only the dense inner loop.  One mma.sync m16n8k8 tf32 per (16 pairs, 8 input
// channels, 8 output channels): the PAIRS are the M dimension, so
//   A = gathered feature rows, LDG.128 straight into the fragment (lane (g,t) loads channels 4t..4t+3
//       of the rows of pairs g and g+8; the k index of the MMA is a permutation of the channels),
//   B = W[k], pre-arranged in fragment order in shared memory (large inputs) or read from L1/L2,
//   C = the current value of the 16 output rows in a per-warp shared-memory accumulator, D goes back
//       to the same place: the scatter-add costs one LDS.128 + one STS.128 per row and no FADD.
// Within one offset every output row occurs once, and a warp walks its groups in order, so the
// accumulation order is fixed: offsets ascending, exactly like the reference's rulebook loop.
// tf32x3 (fp32-equivalent): D = C + A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.
// Large inputs: one warp per tile, 20 warps per SM.  Small inputs: one tile per CTA, its warps split
// the tile's groups and their partial accumulators are summed in warp order.
#include <stdlib.h>

#include "common.cuh"

namespace b200scn {

int scan_i32_public(const int32_t* in, int32_t* out, int64_t n_cap, const int32_t* n_dev, int32_t* total_out,
                    int32_t* end_out, int32_t* bsums, cudaStream_t st);

namespace {

constexpr int kMaxSmem = 220 * 1024;
constexpr int kLargeWarps = 20;

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// D(16 pairs x 8 co) += A(16 pairs x 8 ci) * B(8 ci x 8 co)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---- builder -------------------------------------------------------------------------------------
// one warp per tile; FILL == false: count groups per tile; FILL == true: write the records
template <bool FILL, int R>
__global__ void k_gpl_build(const int32_t* __restrict__ table, int64_t ld, const int32_t* __restrict__ parent,
                            const uint8_t* __restrict__ kofs, int K, int64_t n_cap, const int32_t* __restrict__ n_dev,
                            int32_t* __restrict__ tile_cnt, const int32_t* __restrict__ tile_off,
                            int4* __restrict__ g_rec, int64_t cap_groups) {
  constexpr int NR = R / 32;  // rows per lane
  const int n = live_rows(n_dev, n_cap);
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int64_t n_tiles = ((int64_t)n + R - 1) / R;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int64_t tile = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5); tile < n_tiles;
       tile += (int64_t)gridDim.x * warps) {
    const int base = (int)(tile * R);
    int pk[NR], pp[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      pk[i] = -1; pp[i] = -1;
      const int r = base + 32 * i + lane;
      if (parent != nullptr && r < n) { pk[i] = kofs[r]; pp[i] = __ldg(parent + r); }
    }
    int64_t grp = FILL ? (int64_t)tile_off[tile] : 0;
    int count = 0;
    for (int k = 0; k < K; ++k) {
      int e[NR], pre[NR];
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = base + 32 * i + lane;
        e[i] = -1;
        if (parent != nullptr) {
          e[i] = pk[i] == k ? pp[i] : -1;
        } else if (table != nullptr) {
          if (r < n) e[i] = __ldg(table + (int64_t)k * ld + r);
        } else {
          if (r < n) e[i] = r;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, e[i] >= 0);
        pre[i] = cnt + __popc(m & lt_mask);
        cnt += __popc(m);
      }
      const int ng = (cnt + 15) >> 4;
      if (FILL && ng > 0 && grp + ng <= cap_groups) {
        int32_t* rec = reinterpret_cast<int32_t*>(g_rec + grp * 8);
        auto put = [&](int s, int in_row, int out_row) {
          const int w = s & 15;
          int32_t* ent = rec + ((s >> 4) * 8 + (w & 7)) * 4;
          ent[w >> 3] = in_row;
          reinterpret_cast<unsigned char*>(ent + 2)[w >> 3] = (unsigned char)out_row;
        };
#pragma unroll
        for (int i = 0; i < NR; ++i)
          if (e[i] >= 0) put(pre[i], e[i], 32 * i + lane);
        // padding slots: any valid input row, routed to the dummy accumulator row
        if (lane < ng * 16 - cnt) put(cnt + lane, 0, R);
        for (int j = lane; j < ng * 8; j += 32) {
          unsigned char* b = reinterpret_cast<unsigned char*>(rec + j * 4 + 2);
          b[2] = (unsigned char)k;
          b[3] = 0;
          rec[j * 4 + 3] = 0;
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
  const int32_t* tile_off; const int4* g_rec;
  int K;
  int64_t n_cap; const int32_t* n_dev;
  float* out; int ldo; int accumulate;
  int warps, split;  // split != 0: the CTA's warps share one tile
  uint32_t ones;     // 0xffffffff (opaque to the compiler)
};

__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
// one 32-bit word of a group record.  Scalar on purpose: a vector load pins its destination to an aligned
// register quad, and ptxas then copies the fields out of it right behind the load -- a stall for the
// full memory latency on every group
__device__ __forceinline__ int ldg_nc32(const void* ptr) {
  int v;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(ptr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}

// Q = Cin / 16, MT = output channels of this pass / 16, WSMEM: all K weight slices resident in shared
// memory in fragment order, X3: 3-term tf32 split, R: tile rows.
//
// Fragment <-> channel maps (lane = 4 g + t):
//   A (k-step s = 2q + h): a0/a1 = pair g / g+8 at ci = 16q + 4t + 2h, a2/a3 = same pairs at ci + 1
//   B (n-tile j):          b0 = W[16q + 4t + 2h][8j + g], b1 = W[16q + 4t + 2h + 1][8j + g]
//   C/D (n-tile j):        d0,d1 = pair g, co = 8j + 2t, 8j + 2t + 1;  d2,d3 = pair g+8, same co
// Accumulator row r keeps its 8-channel chunk j at position j ^ x(r), x(r) = (r >> 1) & 1 (MT odd) or
// r & 3 (MT even), so that the LDS.64/STS.64 of four consecutive rows fall into distinct banks.
template <int Q, int MT, bool WSMEM, bool X3, int R>
__global__ void __launch_bounds__(WSMEM ? kLargeWarps * 32 : 256, 1) k_conv_gpl(const GplParams p) {
  constexpr int KS = 2 * Q, CT = 16 * MT, C4 = CT / 4, NT = 2 * MT;
  constexpr int DEPTH = Q == 1 ? 2 : 1;  // feature rows are fetched DEPTH groups ahead, records 2*DEPTH
  constexpr int RING = 2 * DEPTH;
  // WREG: the B fragments of the current offset live in registers and are reloaded (raw fp32 from shared
  // memory, low parts recomputed) only when the offset changes -- the L1/shared data pipe is this
  // kernel's bottleneck (profiles/README.md), and the weights were a third of its shared-memory traffic
  constexpr bool WREG = WSMEM && (Q * MT <= 2);
  constexpr bool WLO = WSMEM && X3 && !WREG;  // separate low-part copy in shared memory
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = live_rows(p.n_dev, p.n_cap);
  const int warps = p.warps;
  const int split = p.split;
  if (split ? ((int64_t)blockIdx.x * R >= n) : (n == 0)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = p.K;
  const int cob = blockIdx.y * CT;

  const int w_elems = WSMEM ? K * KS * MT * 128 : 0;
  float* Whi = reinterpret_cast<float*>(smem_raw);
  float* Wlo = Whi + w_elems;
  constexpr int kAccBytes = (R + 1) * CT * 4;
  const int acc_base = w_elems * 4 * (WLO ? 2 : 1);
  float4* const acc = reinterpret_cast<float4*>(smem_raw + acc_base + warp * kAccBytes);
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t s_acc = s_base + acc_base + warp * kAccBytes + t * 8;  // + row * CT * 4 + chunk * 32
  const uint32_t s_whi = s_base + lane * 16, s_wlo = s_whi + w_elems * 4;

  if (WSMEM) {
    // fragment order: [(k, kstep s, mt)][lane][4] = {b0, b1 of n-tile 2mt, b0, b1 of n-tile 2mt+1}
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
      const int mt = col >> 4, odd = (col >> 3) & 1, gg = col & 7;
      const int dst = ((((k * KS + s) * MT + mt) * 32) + gg * 4 + tt) * 4 + odd * 2 + (rem & 1);
      const float hi = tf32_hi(v);
      Whi[dst] = WREG ? v : hi;
      if (WLO) Wlo[dst] = v - hi;
    }
    __syncthreads();
  }

  const int64_t n_tiles = ((int64_t)n + R - 1) / R;
  // Tile -> warp assignment.  Split mode: tile = CTA.  Otherwise every warp takes a CONTIGUOUS range of
  // tiles of (nearly) equal cost, cost(tile) = groups + kTileCost: the work per tile varies 10x between
  // track and shower regions, and a round-robin assignment left the SMs idle for 20% of the kernel.
  // cost is monotone in the tile index (prefix sums), so the range bounds are two binary searches.
  constexpr int kTileCost = 4;
  int64_t wt_begin, wt_end, wt_stride;
  if (split) {
    wt_begin = blockIdx.x; wt_end = n_tiles; wt_stride = gridDim.x;
  } else {
    const int64_t nw = (int64_t)gridDim.x * warps, w = (int64_t)blockIdx.x * warps + warp;
    const int64_t total = (int64_t)__ldg(p.tile_off + n_tiles) + kTileCost * n_tiles;
    const int64_t lo_cost = total * w / nw, hi_cost = total * (w + 1) / nw;
    // first tile t with prefix cost(t) = tile_off[t] + kTileCost * t >= bound
    int64_t a0 = 0, a1 = n_tiles, b0 = 0, b1 = n_tiles;
    while (a0 < a1 || b0 < b1) {
      const int64_t ma = (a0 + a1) >> 1, mb = (b0 + b1) >> 1;
      const int64_t ca = (int64_t)__ldg(p.tile_off + ma) + kTileCost * ma;
      const int64_t cb = (int64_t)__ldg(p.tile_off + mb) + kTileCost * mb;
      if (a0 < a1) { if (ca >= lo_cost) a1 = ma; else a0 = ma + 1; }
      if (b0 < b1) { if (cb >= hi_cost) b1 = mb; else b0 = mb + 1; }
    }
    wt_begin = a0; wt_end = (w == nw - 1) ? n_tiles : b0; wt_stride = 1;
  }
  const char* const in_l = reinterpret_cast<const char*>(p.in) + 16 * t;
  const uint32_t ldb = (uint32_t)p.ldi * 4u;  // row stride in bytes
  const int4* const rec_l = p.g_rec + g;
  for (int64_t wt = wt_begin; wt < wt_end; wt += wt_stride) {
    const int base = (int)(wt * R);
    __syncwarp();
    for (int e = lane; e < (R + 1) * C4; e += 32) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    // this warp's group range
    int gb = __ldg(p.tile_off + wt), ge = __ldg(p.tile_off + wt + 1);
    if (split) {
      const int per = (ge - gb + warps - 1) / warps;
      const int b2 = gb + warp * per;
      ge = min(ge, b2 + per);
      gb = min(b2, ge);
    }
    __syncwarp();

    // Software pipeline over a ring of RING record slots and DEPTH row slots; the loop body is unrolled
    // RING times so that every slot index is static and no value is ever copied between slots (a copy
    // of a just-issued load stalls the warp until the load returns).  Records and rows past the end of
    // the range are clamped to the last group: redundant but unconditional loads, never consumed.
    int Rx[RING], Ry[RING], Rz[RING];
    float4 G[DEPTH][2][Q];
    const int g_last = max(ge - 1, gb);
    auto fetch_rec = [&](int grp, int u) {
      const int4* r = rec_l + (int64_t)min(grp, g_last) * 8;
      Rx[u] = ldg_nc32(r);
      Ry[u] = ldg_nc32(reinterpret_cast<const int*>(r) + 1);
      Rz[u] = ldg_nc32(reinterpret_cast<const int*>(r) + 2);
    };
    auto fetch_rows = [&](int u, int d) {
      const float4* s0 = reinterpret_cast<const float4*>(in_l + (uint64_t)(uint32_t)Rx[u] * ldb);
      const float4* s1 = reinterpret_cast<const float4*>(in_l + (uint64_t)(uint32_t)Ry[u] * ldb);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        G[d][0][q] = __ldg(s0 + 4 * q);
        G[d][1][q] = __ldg(s1 + 4 * q);
      }
    };
    if (gb < ge) {
#pragma unroll
      for (int u = 0; u < RING; ++u) fetch_rec(gb + u, u);
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) fetch_rows(d, d);
    }

    uint32_t wbh[WREG ? KS : 1][WREG ? MT : 1][4], wbl[WREG ? KS : 1][WREG ? MT : 1][4];
    int wcur = -1;
    for (int g0 = gb; g0 < ge; g0 += RING) {
#pragma unroll
      for (int u = 0; u < RING; ++u) {
        const int grp = g0 + u;
        if (grp >= ge) break;
        const int d = u % DEPTH;
        const uint32_t mz = (uint32_t)Rz[u];
        const int kk = (mz >> 16) & 0xff;
        const int o0 = mz & 0xff, o1 = (mz >> 8) & 0xff;
        const int x0 = (MT & 1) ? ((o0 >> 1) & 1) : (o0 & 3), x1 = (MT & 1) ? ((o1 >> 1) & 1) : (o1 & 3);
        const uint32_t a0 = s_acc + o0 * (CT * 4), a1 = s_acc + o1 * (CT * 4);
        // C operand: the running sums of the 16 output rows
        float dacc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float2 c01 = lds64(a0 + ((j ^ x0) << 5)), c23 = lds64(a1 + ((j ^ x1) << 5));
          dacc[j][0] = c01.x; dacc[j][1] = c01.y; dacc[j][2] = c23.x; dacc[j][3] = c23.y;
        }
        const int kwt = p.kflip ? (K - 1 - kk) : kk;
        const float* wk = WSMEM ? nullptr : p.W + (int64_t)kwt * p.w_sk + (int64_t)(4 * t) * p.w_si +
                                                (int64_t)(cob + g) * p.w_so;
        const uint32_t wk_off = (uint32_t)kk * (KS * MT * 512);
        if (WREG && kk != wcur) {
          wcur = kk;
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const float4 b4 = lds128(s_whi + wk_off + (s * MT + mt) * 512);
              const float w4[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                wbh[WREG ? s : 0][WREG ? mt : 0][r] = __float_as_uint(w4[r]);  // the tensor core truncates
                wbl[WREG ? s : 0][WREG ? mt : 0][r] = __float_as_uint(w4[r] - tf32_hi(w4[r]));
              }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float f0[4] = {G[d][0][q].x, G[d][0][q].y, G[d][0][q].z, G[d][0][q].w};
          const float f1[4] = {G[d][1][q].x, G[d][1][q].y, G[d][1][q].z, G[d][1][q].w};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int s = 2 * q + h;
            const float av[4] = {f0[2 * h], f1[2 * h], f0[2 * h + 1], f1[2 * h + 1]};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              // Every fragment register is DEFINED by an instruction ptxas cannot see through: it knows the
              // tensor core ignores the low mantissa bits, drops a plain mask in front of an HMMA and then
              // re-assembles the operand quad with four MOVs in front of every single HMMA that uses it.
              if (X3) {
                const float lo = av[r] - tf32_hi(av[r]);
                al[r] = __float_as_uint(lo);
                ah[r] = __float_as_uint(av[r] - lo);  // == tf32_hi(av[r]), exactly
              } else {
                ah[r] = __float_as_uint(av[r]) & p.ones;  // single pass: raw fp32 bits
              }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint32_t bh[2][2], bl[2][2];
              if (WREG) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  bh[r >> 1][r & 1] = wbh[WREG ? s : 0][WREG ? mt : 0][r];
                  bl[r >> 1][r & 1] = wbl[WREG ? s : 0][WREG ? mt : 0][r];
                }
              } else if (WSMEM) {
                const float4 b4 = lds128(s_whi + wk_off + (s * MT + mt) * 512);
                bh[0][0] = __float_as_uint(b4.x); bh[0][1] = __float_as_uint(b4.y);
                bh[1][0] = __float_as_uint(b4.z); bh[1][1] = __float_as_uint(b4.w);
                if (X3) {
                  const float4 l4 = lds128(s_wlo + wk_off + (s * MT + mt) * 512);
                  bl[0][0] = __float_as_uint(l4.x); bl[0][1] = __float_as_uint(l4.y);
                  bl[1][0] = __float_as_uint(l4.z); bl[1][1] = __float_as_uint(l4.w);
                }
              } else {
                const float* wp = wk + (int64_t)(16 * q + 2 * h) * p.w_si + (int64_t)(16 * mt) * p.w_so;
                const float w4[4] = {__ldg(wp), __ldg(wp + p.w_si), __ldg(wp + 8 * p.w_so),
                                     __ldg(wp + p.w_si + 8 * p.w_so)};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const float hi = X3 ? tf32_hi(w4[r]) : w4[r];
                  bh[r >> 1][r & 1] = __float_as_uint(hi);
                  bl[r >> 1][r & 1] = __float_as_uint(w4[r] - hi);
                }
              }
#pragma unroll
              for (int o = 0; o < 2; ++o) {
                if (X3) {
                  mma_tf32(dacc[2 * mt + o], al, bh[o]);
                  mma_tf32(dacc[2 * mt + o], ah, bl[o]);
                }
                mma_tf32(dacc[2 * mt + o], ah, bh[o]);
              }
            }
          }
        }
        // refill the two slots this group has just released
        fetch_rows((u + DEPTH) % RING, d);
        fetch_rec(grp + RING, u);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          sts64(a0 + ((j ^ x0) << 5), dacc[j][0], dacc[j][1]);
          sts64(a1 + ((j ^ x1) << 5), dacc[j][2], dacc[j][3]);
        }
        __syncwarp();
      }
    }

    __syncwarp();
    if (split) __syncthreads();
    const bool vec_ok = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const int e_begin = split ? tid : lane, e_step = split ? (int)blockDim.x : 32;
    const float4* acc0 = reinterpret_cast<const float4*>(smem_raw + acc_base);
    for (int e = e_begin; e < R * C4; e += e_step) {
      const int row = e / C4, c4 = e - row * C4;
      if (base + row < n) {
        const int xr = (MT & 1) ? ((row >> 1) & 1) : (row & 3);
        const int src = row * C4 + ((((c4 >> 1) ^ xr) << 1) | (c4 & 1));
        float4 v;
        if (split) {
          v = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int w = 0; w < warps; ++w) {
            const float4 u = acc0[w * (kAccBytes / 16) + src];
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
        } else {
          v = acc[src];
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

size_t acc_bytes(int R, int mt) { return (size_t)(R + 1) * (16 * mt) * 4; }
// shared-memory weight copy of the WSMEM variants: raw fp32 when the fragments are register-cached
// (Q * MT <= 2, see WREG), else separate high and low parts for tf32x3
size_t w_smem_bytes(int K, int Cin, int mt, int x3) {
  const bool wreg = (Cin / 16) * mt <= 2;
  return (size_t)K * Cin * 16 * mt * 4 * ((x3 && !wreg) ? 2 : 1);
}

GPlan plan_gpl(int K, int Cin, int Cout, int64_t n_cap, int x3, int R) {
  GPlan pl{};
  pl.ok = false;
  if (Cin % 16 || Cout % 16 || Cin > 64 || Cin < 16) return pl;
  pl.Q = Cin / 16;
  const int cot = Cout / 16;
  // large: enough tiles to give every SM a CTA full of one-tile warps
  const bool large = R == 128 && n_cap >= (int64_t)R * 8 * sm_count();
  pl.MT = 1;
  if (large) {
    for (int mt = 2; mt >= 1; --mt) {
      if (cot % mt != 0) continue;
      if (w_smem_bytes(K, Cin, mt, x3) + 12 * acc_bytes(R, mt) <= (size_t)kMaxSmem) {
        pl.MT = mt;
        break;
      }
    }
  } else {
    // small inputs: narrow passes until the grid covers the SMs
    const int64_t tiles = cdiv(n_cap, R);
    pl.MT = (cot % 2 == 0 && tiles * (cot / 2) >= sm_count()) ? 2 : 1;
  }
  pl.passes = cot / pl.MT;
  const size_t wbytes = w_smem_bytes(K, Cin, pl.MT, x3);
  pl.wsmem = large && (wbytes + 12 * acc_bytes(R, pl.MT) <= (size_t)kMaxSmem);
  pl.split = large ? 0 : 1;
  const size_t fixed = pl.wsmem ? wbytes : 0;
  static int large_warps = 0;  // B200SCN_GPL_WARPS: warps per CTA of the large-input variant (A/B runs)
  if (large_warps == 0) {
    const char* e = getenv("B200SCN_GPL_WARPS");
    large_warps = e ? atoi(e) : kLargeWarps;
    if (large_warps < 1 || large_warps > kLargeWarps) large_warps = kLargeWarps;
  }
  int warps = pl.wsmem ? large_warps : 8;
  while (warps > 1 && fixed + warps * acc_bytes(R, pl.MT) > (size_t)kMaxSmem) --warps;
  if (fixed + warps * acc_bytes(R, pl.MT) > (size_t)kMaxSmem) return pl;
  pl.warps = warps;
  pl.smem = fixed + warps * acc_bytes(R, pl.MT);
  int cps = (int)((size_t)(227 * 1024) / (pl.smem + 1024));
  if (cps > 2048 / (warps * 32)) cps = 2048 / (warps * 32);
  if (cps > 8) cps = 8;
  if (cps < 1) cps = 1;
  pl.ctas_per_sm = cps;
  pl.ok = true;
  return pl;
}

template <int Q, int MT, bool WSMEM, bool X3, int R>
int launch_gpl(const GplParams& p, const GPlan& pl, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_conv_gpl<Q, MT, WSMEM, X3, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kMaxSmem));
    configured = true;
  }
  const int64_t tiles = cdiv(p.n_cap, R);
  int64_t gx = pl.split ? tiles : cdiv(tiles, pl.warps);
  int64_t gmax = WSMEM ? ((int64_t)sm_count() * pl.ctas_per_sm) / pl.passes : (int64_t)sm_count() * pl.ctas_per_sm * 4;
  if (gmax < 1) gmax = 1;
  if (gx > gmax) gx = gmax;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)pl.passes);
  k_conv_gpl<Q, MT, WSMEM, X3, R><<<grid, pl.warps * 32, pl.smem, st>>>(p);
  count_launches(1);
  B200_CHECK_LAUNCH("conv_gpl");
  return 0;
}

template <int Q, int MT, bool X3>
int launch_gpl_qmx(const GplParams& p, const GPlan& pl, int R, cudaStream_t st) {
  if (pl.wsmem) return launch_gpl<Q, MT, true, X3, 128>(p, pl, st);
  return R == 128 ? launch_gpl<Q, MT, false, X3, 128>(p, pl, st) : launch_gpl<Q, MT, false, X3, 64>(p, pl, st);
}

template <int Q>
int launch_gpl_q(const GplParams& p, const GPlan& pl, int R, int x3, cudaStream_t st) {
  if (pl.MT == 1) return x3 ? launch_gpl_qmx<Q, 1, true>(p, pl, R, st) : launch_gpl_qmx<Q, 1, false>(p, pl, R, st);
  return x3 ? launch_gpl_qmx<Q, 2, true>(p, pl, R, st) : launch_gpl_qmx<Q, 2, false>(p, pl, R, st);
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

/* tile rows the library recommends for a level of n_cap rows */
extern "C" int b200scn_gpl_tile_rows(int64_t n_cap) { return n_cap >= 12288 ? 128 : 64; }

/* upper bound of the number of groups for n_cap rows (static regime allocation) */
extern "C" int64_t b200scn_gpl_max_groups(int64_t n_cap, int K, int tile_rows) {
  const int64_t tiles = cdiv(n_cap < 1 ? 1 : n_cap, tile_rows);
  return tiles * (int64_t)K * (tile_rows / 16);
}

extern "C" int64_t b200scn_gpl_scratch_bytes(int64_t n_cap) {
  const int64_t tiles = cdiv(n_cap < 1 ? 1 : n_cap, 64);
  return (tiles + 8) * 4 + (cdiv(tiles, 1024) + 16) * 4 + 256;
}

/* stage 0: tile_off[tiles + 1] (exclusive scan of the per-tile group counts) and total_dev[1].
 * stage 1: fill g_rec[cap_groups * 8] (int4 records) using tile_off. */
extern "C" int b200scn_gpl_build(int stage, const int32_t* table, int64_t ld, const int32_t* parent,
                                 const uint8_t* kofs, int K, int64_t n_cap, const int32_t* n_dev, int tile_rows,
                                 int32_t* tile_off, int32_t* total_dev, void* g_rec, int64_t cap_groups, void* scratch,
                                 int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(K >= 1 && K <= 64, "gpl_build: K must be 1..64");
  B200_CHECK_ARG(tile_rows == 64 || tile_rows == 128, "gpl_build: tile_rows must be 64 or 128");
  B200_CHECK_ARG(scratch_bytes >= b200scn_gpl_scratch_bytes(n_cap), "gpl_build: scratch too small");
  const int64_t tiles = cdiv(n_cap < 1 ? 1 : n_cap, tile_rows);
  int32_t* tile_cnt = reinterpret_cast<int32_t*>(scratch);
  int32_t* bsums = tile_cnt + tiles + 8;
  const int warps = 4;
  int64_t grid = cdiv(tiles, warps);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  int4* rec = reinterpret_cast<int4*>(g_rec);
  if (stage == 0) {
    B200_CUDA(cudaMemsetAsync(tile_cnt, 0, (size_t)(tiles + 8) * 4, st));
    if (n_cap > 0) {
      if (tile_rows == 128)
        k_gpl_build<false, 128><<<(int)grid, warps * 32, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, tile_cnt,
                                                                  nullptr, nullptr, 0);
      else
        k_gpl_build<false, 64><<<(int)grid, warps * 32, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, tile_cnt,
                                                                 nullptr, nullptr, 0);
    }
    // exclusive scan over all `tiles` counts (tiles beyond the live ones hold 0)
    if (int rc = scan_i32_public(tile_cnt, tile_off, tiles, nullptr, total_dev, tile_off, bsums, st)) return rc;
    count_launches(4);
  } else {
    if (n_cap > 0) {
      if (tile_rows == 128)
        k_gpl_build<true, 128><<<(int)grid, warps * 32, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, nullptr,
                                                                 tile_off, rec, cap_groups);
      else
        k_gpl_build<true, 64><<<(int)grid, warps * 32, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, nullptr,
                                                                tile_off, rec, cap_groups);
    }
    count_launches(1);
  }
  B200_CHECK_LAUNCH("gpl_build");
  return 0;
}

extern "C" int b200scn_conv_gpl(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so,
                                int kflip, int tile_rows, const int32_t* tile_off, const void* g_rec, int K,
                                int64_t n_out_cap, const int32_t* n_out_dev, int Cin, int Cout, float* out, int64_t ldo,
                                int accumulate, int math_mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(Cin % 16 == 0 && Cout % 16 == 0 && Cin >= 16 && Cout >= 16, "conv_gpl: channels must be multiples of 16");
  B200_CHECK_ARG(math_mode == B200SCN_MATH_TF32 || math_mode == B200SCN_MATH_TF32X3, "conv_gpl: tensor-core math modes only");
  B200_CHECK_ARG(ldi % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "conv_gpl: input rows must be 16-byte aligned");
  B200_CHECK_ARG(tile_rows == 64 || tile_rows == 128, "conv_gpl: tile_rows must be 64 or 128");
  const int64_t lim = (1ll << 31) - 1;
  B200_CHECK_ARG(ldi < lim && ldo < lim && (int64_t)K * Cin * Cout < lim, "conv_gpl: stride out of 32-bit range");
  if (n_out_cap == 0) return 0;
  const int x3 = math_mode == B200SCN_MATH_TF32X3 ? 1 : 0;
  for (int c0 = 0; c0 < Cin; c0 += 64) {
    const int cc = Cin - c0 < 64 ? Cin - c0 : 64;
    GPlan pl = plan_gpl(K, cc, Cout, n_out_cap, x3, tile_rows);
    B200_CHECK_ARG(pl.ok, "conv_gpl: unsupported shape Cin=%d Cout=%d K=%d", Cin, Cout, K);
    GplParams p;
    p.in = in + c0; p.ldi = (int)ldi; p.W = W + (int64_t)c0 * w_si; p.w_sk = (int)w_sk; p.w_si = (int)w_si;
    p.w_so = (int)w_so; p.kflip = kflip; p.tile_off = tile_off; p.g_rec = reinterpret_cast<const int4*>(g_rec);
    p.K = K; p.n_cap = n_out_cap; p.n_dev = n_out_dev; p.out = out;
    p.ldo = (int)ldo; p.accumulate = (c0 > 0) ? 1 : accumulate; p.warps = pl.warps; p.split = pl.split; p.ones = 0xffffffffu;
    int rc;
    switch (pl.Q) {
      case 1: rc = launch_gpl_q<1>(p, pl, tile_rows, x3, st); break;
      case 2: rc = launch_gpl_q<2>(p, pl, tile_rows, x3, st); break;
      case 3: rc = launch_gpl_q<3>(p, pl, tile_rows, x3, st); break;
      default: rc = launch_gpl_q<4>(p, pl, tile_rows, x3, st); break;
    }
    if (rc) return rc;
  }
  return 0;
}
