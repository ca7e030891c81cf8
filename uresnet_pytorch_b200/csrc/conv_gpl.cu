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
constexpr int kRecRing = 8;  // per-warp shared-memory ring of group records (128 B each)

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// D(16 pairs x 8 co) += A(16 pairs x 8 ci) * B(8 ci x 8 co)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Record flags (byte 3 of word 2), set by the builder's centre-first layout and read by k_conv_gpl
constexpr int kGplCZero = 1;    // C = 0 (first group of a sub-tile: nothing to read)

// ---- builder -------------------------------------------------------------------------------------
// one warp per tile; FILL == false: count groups per tile; FILL == true: write the records
template <bool FILL, int R>
__global__ void k_gpl_build(const int32_t* __restrict__ table, int64_t ld, const int32_t* __restrict__ parent,
                            const uint8_t* __restrict__ kofs, int K, int64_t n_cap, const int32_t* __restrict__ n_dev,
                            int32_t* __restrict__ tile_cnt, const int32_t* __restrict__ tile_off,
                            int32_t* __restrict__ tile_beg_out, int4* __restrict__ g_rec, int64_t cap_groups,
                            int32_t* __restrict__ koff, int layout) {
  constexpr int NR = R / 32;  // rows per lane
  // centre-first layout (layout 1, submanifold tables only): the centre offset is listed FIRST and per 16-row
  // sub-tile -- slot j of such a group is output row 16 s + j -- so that the convolution kernel starts the
  // sub-tile's sums in registers instead of reading the shared-memory accumulator (flag kGplCZero, see k_conv_gpl)
  const int kc = (K - 1) >> 1;
  const int nd = (table != nullptr && parent == nullptr && layout == 1) ? 1 : 0;
  const int kd0 = kc;  // the dense offset
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
    // fixed-stride mode (FILL && tile_cnt): tile t owns the slots [t * stride, (t + 1) * stride) and the
    // builder publishes its begin/end itself -- one pass, no count + scan
    const bool fixed = FILL && tile_beg_out != nullptr;
    int64_t grp = FILL ? (fixed ? tile * (int64_t)K * (R / 16) : (int64_t)tile_off[tile]) : 0;
    const int64_t grp0 = grp;
    int count = 0;
    // table entries of offset k + 1 are requested while offset k is being listed (the per-offset work is a
    // short dependent chain, the L2 round trip of the table column would otherwise dominate it)
    int nxt[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = base + 32 * i + lane;
      nxt[i] = (table != nullptr && parent == nullptr && r < n) ? __ldg(table + r) : -1;
    }
    if (nd > 0) {
      const int steps = (R / 16) * nd;
      if (FILL && grp + steps <= cap_groups) {
        int32_t* rec = reinterpret_cast<int32_t*>(g_rec + grp * 8);
        for (int dk = 0; dk < nd; ++dk) {
          const int k = kd0 + dk;
          const int flags = kGplCZero;
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            const int r = base + 32 * i + lane;
            // (the centre "neighbour" always exists; rows past the end read row 0 into sums nobody stores)
            const int in_row = r < n ? __ldg(table + (int64_t)k * ld + r) : 0;
            const int st = 2 * i + (lane >> 4), j = lane & 15;  // sub-tile, slot
            int32_t* ent = rec + ((st * nd + dk) * 8 + (j & 7)) * 4;
            ent[j >> 3] = in_row;
            reinterpret_cast<unsigned char*>(ent + 2)[j >> 3] = (unsigned char)(16 * st + j);
            if (j < 8) {
              unsigned char* b = reinterpret_cast<unsigned char*>(ent + 2);
              b[2] = (unsigned char)k;
              b[3] = (unsigned char)flags;
              ent[3] = 0;
            }
          }
        }
      }
      grp += steps;
      count += steps;
    }
    for (int k = 0; k < K; ++k) {
      int e[NR], pre[NR];
      int cnt = 0;
      if (FILL && koff != nullptr && lane == 0) koff[tile * (K + 1) + k] = (int32_t)grp;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = base + 32 * i + lane;
        e[i] = -1;
        if (parent != nullptr) {
          e[i] = pk[i] == k ? pp[i] : -1;
        } else if (table != nullptr) {
          e[i] = (k >= kd0 && k < kd0 + nd) ? -1 : nxt[i];
          nxt[i] = (k + 1 < K && r < n) ? __ldg(table + (int64_t)(k + 1) * ld + r) : -1;
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
    if (FILL && koff != nullptr && lane == 0) koff[tile * (K + 1) + K] = (int32_t)grp;
    if (!FILL && lane == 0) tile_cnt[tile] = count;
    if (fixed && lane == 0) {
      tile_beg_out[tile] = (int32_t)grp0;
      tile_cnt[tile] = (int32_t)(grp0 + count);  // = tile_end
    }
  }
}

// ---- convolution over the lists ------------------------------------------------------------------
struct GplParams {
  const float* in; int ldi;
  const float* W; int w_sk, w_si, w_so; int kflip;
  const float* Wp; int cot;  // fragment-ordered copy of W for this direction (b200scn_pack_weights), Cout / 16
  const int32_t* tile_off; const int32_t* tile_end; const int4* g_rec;  // tile_end == NULL: tile_off + 1
  int Q;  // Cin / 16 (runtime value for k_conv_gpl_small)
  int K;
  int64_t n_cap; const int32_t* n_dev;
  float* out; int ldo; int accumulate;
  const float* res; int ldr;  // optional residual added to the result (shortcut of a residual block)
  int warps, split;  // split != 0: the CTA's warps share one tile
  uint32_t ones;     // 0xffffffff (opaque to the compiler)
  uint32_t zero;     // 0 (opaque to the compiler)
  int layout;        // pair-list layout (b200scn_gpl_build)
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
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// Centre-first groups (record flag kGplCZero, byte 3 of word 2): the group's 16 slots are the 16 consecutive
// output rows of one sub-tile and the centre offset is the first one listed, so the sub-tile's sums START in the
// MMA accumulator registers (C = 0) instead of being read from the zero-filled shared-memory accumulator.

// Large inputs.  Q = Cin / 16, MT = output channels of this pass / 16, X3: 3-term tf32 split.  One warp per
// 128-row tile; all K weight slices resident in shared memory in fragment order.
//
// Fragment <-> channel maps (lane = 4 g + t):
//   A (k-step s = 2q + h): a0/a1 = pair g / g+8 at ci = 16q + 4t + 2h, a2/a3 = same pairs at ci + 1
//   B (n-tile j = 2mt + jj, column n = g): channel co = 16mt + 4 (g >> 1) + 2 jj + (g & 1);
//                          b0 = W[16q + 4t + 2h][co], b1 = W[16q + 4t + 2h + 1][co]
//   C/D (n-tile 2mt + jj):  d0,d1 = pair g, co = 16mt + 4t + 2jj, + 1;  d2,d3 = pair g+8, same co
// so lane (g, t) owns the four CONSECUTIVE channels 16mt + 4t .. + 3 of its two rows: the scatter-add is one
// LDS.128 + one STS.128 per (row, mt) on an accumulator kept in natural channel order.  Rows of a 32-channel
// pass (128 B) swap their 64-byte halves on odd rows, so that the two rows a quarter-warp touches in one
// shared-memory wavefront collide only when their parities are equal (the same rule as for 64-byte rows).
//
// Pipeline (profiles/README.md section 0.3): ptxas tracks ALL row gathers of a warp on one scoreboard, and a
// wait on a scoreboard waits for every load on it -- so the first use of a group's rows also waits for
// whatever was requested after them.  With rows requested two groups ahead (round 1) this exposed the full
// memory latency once per loop trip (20% of all stall samples on one instruction).  Here the rows of the next
// BATCH of NB groups are requested right AFTER the first use of the current batch's rows (their address carries
// a data dependence on a loaded value, so ptxas cannot hoist them): every wait then only covers loads that are
// a whole batch old, and NB groups' rows per warp are in flight -- the kernel is bound by memory latency x
// bytes in flight first (0.321 -> 0.297 ms from NB = 1 to 2), by the L1/shared data pipe next (74% busy).
// Measured and rejected: rows through a cp.async shared-memory ring (removes the latency stalls, adds 20
// wavefronts per group: 0.344 ms; L1-bypassing copies 1.1 ms), prefetch.global.L2 (0.8 ms), NB = 3 (spills).
template <int Q, int MT, bool X3, int NB>
__global__ void __launch_bounds__(kLargeWarps * 32, 1) k_conv_gpl(const GplParams p) {
  static_assert(NB >= 1 && NB <= 3, "batch of 1..3 groups");
  constexpr int R = 128, KS = 2 * Q, CT = 16 * MT, C4 = CT / 4, NT = 2 * MT;
  // WREG: the B fragments of the current offset live in registers and are reloaded (raw fp32 from shared
  // memory, low parts recomputed) only when the offset changes -- the L1/shared data pipe is this
  // kernel's bottleneck (profiles/README.md), and the weights were a third of its shared-memory traffic
  constexpr bool WREG = (Q * MT <= 2);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_wait();
  const int n = live_rows(p.n_dev, p.n_cap);
  const int warps = p.warps;
  if (n == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = p.K;
  const int cob = blockIdx.y * CT;

  const int w_elems = K * KS * MT * 128;
  float* Whi = reinterpret_cast<float*>(smem_raw);
  constexpr int kAccBytes = (R + 1) * CT * 4;
  const int acc_base = w_elems * 4;
  float4* const acc = reinterpret_cast<float4*>(smem_raw + acc_base + warp * kAccBytes);
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t s_acc = s_base + acc_base + warp * kAccBytes;  // + row * CT * 4 + chunk * 16
  const uint32_t s_whi = s_base + lane * 16;
  const uint32_t s_ring = s_base + acc_base + warps * kAccBytes + warp * (kRecRing * 128);

  {
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
      const int mt = col >> 4;
      const int jj = (col >> 1) & 1, gg = (((col & 15) >> 2) << 1) | (col & 1);
      const int dst = ((((k * KS + s) * MT + mt) * 32) + gg * 4 + tt) * 4 + jj * 2 + (rem & 1);
      Whi[dst] = v;  // raw fp32: the tensor core truncates the high part itself, the low part is recomputed
    }
    __syncthreads();
  }

  const int64_t n_tiles = ((int64_t)n + R - 1) / R;
  // Tile -> warp assignment: every warp takes a CONTIGUOUS range of tiles of (nearly) equal cost,
  // cost(tile) = groups + kTileCost: the work per tile varies 10x between track and shower regions, and a
  // round-robin assignment left the SMs idle for 20% of the kernel.  cost is monotone in the tile index
  // (prefix sums), so the range bounds are two binary searches.
  constexpr int kTileCost = 4;
  int64_t wt_begin, wt_end;
  {
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
    wt_begin = a0; wt_end = (w == nw - 1) ? n_tiles : b0;
  }
  const char* const in_l = reinterpret_cast<const char*>(p.in) + 16 * t;
  const uint32_t ldb = (uint32_t)p.ldi * 4u;  // row stride in bytes
  const bool vec_ok = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
  // accumulator address of (row, mt) for this lane: 16-byte chunk 4mt + t, halves swapped on odd rows (MT == 2)
  auto acc_addr = [&](int row, int mt) -> uint32_t {
    const int chunk = (4 * mt + t) ^ ((MT == 2) ? ((row & 1) << 2) : 0);
    return s_acc + (uint32_t)row * (CT * 4) + (uint32_t)chunk * 16u;
  };
  for (int64_t wt = wt_begin; wt < wt_end; ++wt) {
    const int base = (int)(wt * R);
    __syncwarp();
    // centre-first lists: every row of the tile is first WRITTEN by its centre group (C = 0), which precedes all
    // other groups of the tile -- nothing to clear (the dummy row of the padding slots may hold anything)
    if (p.layout != 1)
      for (int e = lane; e < (R + 1) * C4; e += 32) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int gb = __ldg(p.tile_off + wt), ge = __ldg(p.tile_off + wt + 1);
    __syncwarp();

    // Group records stream through a per-warp shared-memory ring filled by cp.async kRecRing - 1 groups ahead
    // (an async copy has no destination register, so ptxas cannot sink it next to its use the way it does
    // with plain loads).  Records and rows past the end of the range are clamped to the last group:
    // redundant but unconditional, never consumed.
    float4 G[2][NB][2][Q];
    int Rz[2][NB];
    const int g_last = max(ge - 1, gb);
    const uint32_t ring_cp = s_ring + lane * 4, ring_rd = s_ring + g * 16;
    auto fetch_rec = [&](int grp) {  // lane l copies word l of the 128-byte record
      const int32_t* src = reinterpret_cast<const int32_t*>(p.g_rec + (int64_t)min(grp, g_last) * 8) + lane;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ring_cp + (uint32_t)(grp & (kRecRing - 1)) * 128u),
                   "l"(src) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // rows of group `grp` (its record must have landed in the ring) into register slot d; `dep` is zero, but
    // only at run time: it ties the request to the value it was derived from.
    auto fetch_rows = [&](int grp, int d, int bi, uint32_t dep) {
      const float4 r = lds128(ring_rd + (uint32_t)(grp & (kRecRing - 1)) * 128u);
      Rz[d][bi] = __float_as_int(r.z);
      const int i0 = __float_as_int(r.x), i1 = __float_as_int(r.y);
      const float4* s0 = reinterpret_cast<const float4*>(in_l + dep + (uint64_t)(uint32_t)i0 * ldb);
      const float4* s1 = reinterpret_cast<const float4*>(in_l + dep + (uint64_t)(uint32_t)i1 * ldb);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        G[d][bi][0][q] = __ldg(s0 + 4 * q);
        G[d][bi][1][q] = __ldg(s1 + 4 * q);
      }
    };
    // Record accounting: at the start of the batch that begins at group b0 the records up to b0 + kRecRing - 2
    // have been requested; the batch requests NB more, and needs the NEXT batch's (up to b0 + 2 NB - 1) landed
    // before it requests that batch's rows: at most kRecRing - 1 - NB request groups may still be pending.  (The
    // new requests overwrite the ring slots of b0 - 1 .. b0 + NB - 2, whose fields were copied to registers when
    // their rows were requested.)
    if (gb < ge) {
      __syncwarp();  // the previous tile's last ring reads are done
#pragma unroll
      for (int u = 0; u < kRecRing - 1; ++u) fetch_rec(gb + u);
      asm volatile("cp.async.wait_group %0;" ::"n"(kRecRing - 1 - NB) : "memory");
      __syncwarp();
#pragma unroll
      for (int bi = 0; bi < NB; ++bi) fetch_rows(gb + bi, 0, bi, 0u);
    }

    uint32_t wbh[WREG ? KS : 1][WREG ? MT : 1][4], wbl[WREG ? KS : 1][WREG ? MT : 1][4];
    int wcur = -1;
    float dacc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) dacc[j][0] = dacc[j][1] = dacc[j][2] = dacc[j][3] = 0.f;
    for (int g0 = gb; g0 < ge; g0 += 2 * NB) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int b0 = g0 + u * NB;
        if (b0 >= ge) break;
        const int d = u;
#pragma unroll
        for (int bi = 0; bi < NB; ++bi) fetch_rec(b0 + bi + kRecRing - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(kRecRing - 1 - NB) : "memory");
        __syncwarp();
        // first use of this batch's rows, then the request for the next batch's
        const uint32_t dep = __float_as_uint(G[d][0][0][0].x) & p.zero;
#pragma unroll
        for (int bi = 0; bi < NB; ++bi) fetch_rows(b0 + NB + bi, d ^ 1, bi, dep);
#pragma unroll
        for (int bi = 0; bi < NB; ++bi) {
        const int grp = b0 + bi;
        if (grp >= ge) break;
        const uint32_t mz = (uint32_t)Rz[d][bi];
        const int kk = (mz >> 16) & 0xff, fl = (int)(mz >> 24);
        const int o0 = mz & 0xff, o1 = (mz >> 8) & 0xff;
        // C operand: the running sums of the 16 output rows
        if (fl & kGplCZero) {
#pragma unroll
          for (int j = 0; j < NT; ++j) dacc[j][0] = dacc[j][1] = dacc[j][2] = dacc[j][3] = 0.f;
        } else {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const float4 c0 = lds128(acc_addr(o0, mt)), c1 = lds128(acc_addr(o1, mt));
            dacc[2 * mt][0] = c0.x; dacc[2 * mt][1] = c0.y; dacc[2 * mt + 1][0] = c0.z; dacc[2 * mt + 1][1] = c0.w;
            dacc[2 * mt][2] = c1.x; dacc[2 * mt][3] = c1.y; dacc[2 * mt + 1][2] = c1.z; dacc[2 * mt + 1][3] = c1.w;
          }
        }
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
          const float f0[4] = {G[d][bi][0][q].x, G[d][bi][0][q].y, G[d][bi][0][q].z, G[d][bi][0][q].w};
          const float f1[4] = {G[d][bi][1][q].x, G[d][bi][1][q].y, G[d][bi][1][q].z, G[d][bi][1][q].w};
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
              } else {
                const float4 b4 = lds128(s_whi + wk_off + (s * MT + mt) * 512);
                const float w4[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  bh[r >> 1][r & 1] = __float_as_uint(w4[r]);
                  if (X3) bl[r >> 1][r & 1] = __float_as_uint(w4[r] - tf32_hi(w4[r]));
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
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          sts128(acc_addr(o0, mt), dacc[2 * mt][0], dacc[2 * mt][1], dacc[2 * mt + 1][0], dacc[2 * mt + 1][1]);
          sts128(acc_addr(o1, mt), dacc[2 * mt][2], dacc[2 * mt][3], dacc[2 * mt + 1][2], dacc[2 * mt + 1][3]);
        }
        __syncwarp();
        }
      }
    }

    __syncwarp();
    for (int e = lane; e < R * C4; e += 32) {
      const int row = e / C4, c4 = e - row * C4;
      if (base + row < n) {
        const int src = row * C4 + (c4 ^ ((MT == 2) ? ((row & 1) << 2) : 0));
        float4 v = acc[src];
        float* o = p.out + (int64_t)(base + row) * p.ldo + cob + 4 * c4;
        if (p.res != nullptr) {
          const float* rr = p.res + (int64_t)(base + row) * p.ldr + cob + 4 * c4;
          v.x += __ldg(rr); v.y += __ldg(rr + 1); v.z += __ldg(rr + 2); v.w += __ldg(rr + 3);
        }
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
  }
}

// ---- small inputs ----------------------------------------------------------------------------------
// One CTA per (tile, pass of CT = 16 MT output channels); its 8 warps split the tile's groups, each with
// a private partial accumulator, and the partials are summed in warp order.  Any Cin % 16 == 0 in ONE
// launch: the input channels are walked in 16-channel chunks (runtime loop), chunk q+1 -- feature rows
// and weight fragments, both straight from L1/L2 -- in flight while chunk q is on the tensor pipe.
// Same fragment <-> channel maps as k_conv_gpl.
template <int MT>
struct GplChunk {
  float4 f0, f1;      // channels 16q + 4t .. +3 of the rows of pairs g and g+8
  float w[2][MT][4];  // [h][mt] = {b0, b1 of n-tile 2mt, b0, b1 of n-tile 2mt+1}
};

template <int MT, bool X3, int R, bool PACKED>
__global__ void __launch_bounds__(512, MT == 1 ? 2 : 1) k_conv_gpl_small(const GplParams p) {
  constexpr int CT = 16 * MT, C4 = CT / 4, NT = 2 * MT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_wait();
  const int n = live_rows(p.n_dev, p.n_cap);
  if ((int64_t)blockIdx.x * R >= n) return;
  const int warps = blockDim.x >> 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = p.K, Q = p.Q;
  const int cob = blockIdx.y * CT;
  constexpr int kAccBytes = (R + 1) * CT * 4;
  float4* const acc = reinterpret_cast<float4*>(smem_raw + warp * kAccBytes);
  const uint32_t s_acc = (uint32_t)__cvta_generic_to_shared(smem_raw) + warp * kAccBytes + t * 8;
  const char* const in_l = reinterpret_cast<const char*>(p.in) + 16 * t;
  const uint32_t ldb = (uint32_t)p.ldi * 4u;
  const int4* const rec_l = p.g_rec + g;
  const int64_t n_tiles = ((int64_t)n + R - 1) / R;
  const float* const w_l = p.W + (int64_t)(4 * t) * p.w_si + (int64_t)(cob + g) * p.w_so;

  for (int64_t wt = blockIdx.x; wt < n_tiles; wt += gridDim.x) {
    const int base = (int)(wt * R);
    __syncwarp();
    for (int e = lane; e < (R + 1) * C4; e += 32) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    int gb = __ldg(p.tile_off + wt), ge = __ldg((p.tile_end ? p.tile_end : p.tile_off + 1) + wt);
    // split == 2: warp w takes groups w, w + warps, ... -- the groups of a tile are sorted by offset, so all
    // warps of the CTA then work on the same two or three offsets at any time and their packed weight
    // fragments stay L1-resident (a contiguous range per warp spreads the CTA over all K offsets: 440 KB of
    // weights for a 64 -> 64 layer, L1 hit rate 27-52% in the ncu capture).  split == 1: contiguous ranges.
    int gs = 1;
    if (p.split == 2) {
      gs = warps;
      gb = min(gb + warp, ge);
    } else {
      const int per = (ge - gb + warps - 1) / warps;
      const int b2 = gb + warp * per;
      ge = min(ge, b2 + per);
      gb = min(b2, ge);
    }
    __syncwarp();
    if (gb < ge) {
      const int g_last = gb + ((ge - 1 - gb) / gs) * gs;  // this warp's last group
      auto load_rec = [&](int grp, int& x, int& y, int& z) {
        const int* r = reinterpret_cast<const int*>(rec_l + (int64_t)min(grp, g_last) * 8);
        x = ldg_nc32(r); y = ldg_nc32(r + 1); z = ldg_nc32(r + 2);
      };
      // PACKED: one coalesced LDG.128 per (h, mt) from the fragment-ordered copy (the offset flip is
      // already folded into it); else four scalar loads from the strided original -- 4x the L1 wavefronts
      const float* const wp_l = p.Wp + (size_t)(cob >> 4) * 128 + lane * 4;
      auto w_of = [&](int z) {
        const int kk = (z >> 16) & 0xff;
        if (PACKED) return wp_l + (size_t)kk * Q * 2 * p.cot * 128;
        return w_l + (int64_t)(p.kflip ? (K - 1 - kk) : kk) * p.w_sk;
      };
      auto load_chunk = [&](GplChunk<MT>& c, int x, int y, const float* wk, int q) {
        c.f0 = __ldg(reinterpret_cast<const float4*>(in_l + (uint64_t)(uint32_t)x * ldb) + 4 * q);
        c.f1 = __ldg(reinterpret_cast<const float4*>(in_l + (uint64_t)(uint32_t)y * ldb) + 4 * q);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (PACKED) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(wk + (size_t)((2 * q + h) * p.cot + mt) * 128));
              c.w[h][mt][0] = v.x; c.w[h][mt][1] = v.y; c.w[h][mt][2] = v.z; c.w[h][mt][3] = v.w;
            } else {
              const float* wp = wk + (int64_t)(16 * q + 2 * h) * p.w_si + (int64_t)(16 * mt) * p.w_so;
              c.w[h][mt][0] = __ldg(wp);
              c.w[h][mt][1] = __ldg(wp + p.w_si);
              c.w[h][mt][2] = __ldg(wp + 8 * p.w_so);
              c.w[h][mt][3] = __ldg(wp + p.w_si + 8 * p.w_so);
            }
          }
      };
      // records run two groups ahead of the group on the tensor pipe, chunks one chunk ahead
      int x0, y0, z0, x1, y1, z1, x2, y2, z2;
      load_rec(gb, x0, y0, z0);
      load_rec(gb + gs, x1, y1, z1);
      GplChunk<MT> cur, nxt;
      load_chunk(cur, x0, y0, w_of(z0), 0);
      for (int grp = gb; grp < ge; grp += gs) {
        load_rec(grp + 2 * gs, x2, y2, z2);
        const int o0 = z0 & 0xff, o1 = (z0 >> 8) & 0xff;
        const int xs0 = (MT & 1) ? ((o0 >> 1) & 1) : (o0 & 3), xs1 = (MT & 1) ? ((o1 >> 1) & 1) : (o1 & 3);
        const uint32_t a0 = s_acc + o0 * (CT * 4), a1 = s_acc + o1 * (CT * 4);
        float dacc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float2 c01 = lds64(a0 + ((j ^ xs0) << 5)), c23 = lds64(a1 + ((j ^ xs1) << 5));
          dacc[j][0] = c01.x; dacc[j][1] = c01.y; dacc[j][2] = c23.x; dacc[j][3] = c23.y;
        }
        const float* const wk0 = w_of(z0);
        const float* const wk1 = w_of(z1);
        for (int q = 0; q < Q; ++q) {
          if (q + 1 < Q) load_chunk(nxt, x0, y0, wk0, q + 1);
          else load_chunk(nxt, x1, y1, wk1, 0);  // first chunk of the next group (clamped at the end: unused)
          const float f0[4] = {cur.f0.x, cur.f0.y, cur.f0.z, cur.f0.w};
          const float f1[4] = {cur.f1.x, cur.f1.y, cur.f1.z, cur.f1.w};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float av[4] = {f0[2 * h], f1[2 * h], f0[2 * h + 1], f1[2 * h + 1]};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              if (X3) {
                const float lo = av[r] - tf32_hi(av[r]);
                al[r] = __float_as_uint(lo);
                ah[r] = __float_as_uint(av[r] - lo);
              } else {
                ah[r] = __float_as_uint(av[r]) & p.ones;
              }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint32_t bh[2][2], bl[2][2];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const float wv = cur.w[h][mt][r];
                bh[r >> 1][r & 1] = __float_as_uint(wv);  // the tensor core truncates to tf32 itself
                bl[r >> 1][r & 1] = __float_as_uint(wv - tf32_hi(wv));
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
          cur = nxt;
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          sts64(a0 + ((j ^ xs0) << 5), dacc[j][0], dacc[j][1]);
          sts64(a1 + ((j ^ xs1) << 5), dacc[j][2], dacc[j][3]);
        }
        __syncwarp();
        x0 = x1; y0 = y1; z0 = z1;
        x1 = x2; y1 = y2; z1 = z2;
      }
    }
    __syncthreads();
    const bool vec_ok = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const float4* acc0 = reinterpret_cast<const float4*>(smem_raw);
    for (int e = tid; e < R * C4; e += blockDim.x) {
      const int row = e / C4, c4 = e - row * C4;
      if (base + row < n) {
        const int xr = (MT & 1) ? ((row >> 1) & 1) : (row & 3);
        const int src = row * C4 + ((((c4 >> 1) ^ xr) << 1) | (c4 & 1));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < warps; ++w) {
          const float4 u = acc0[w * (kAccBytes / 16) + src];
          v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        float* o = p.out + (int64_t)(base + row) * p.ldo + cob + 4 * c4;
        if (p.res != nullptr) {
          const float* rr = p.res + (int64_t)(base + row) * p.ldr + cob + 4 * c4;
          v.x += __ldg(rr); v.y += __ldg(rr + 1); v.z += __ldg(rr + 2); v.w += __ldg(rr + 3);
        }
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
    __syncthreads();
  }
}

size_t acc_bytes(int R, int mt) { return (size_t)(R + 1) * (16 * mt) * 4; }
// shared-memory weight copy of k_conv_gpl: raw fp32 (low parts of the tf32x3 split are recomputed)
size_t w_smem_bytes(int K, int Cin, int mt, int x3) {
  (void)x3;
  return (size_t)K * Cin * 16 * mt * 4;
}

int large_warps() {
  static int v = 0;  // B200SCN_GPL_WARPS: warps per CTA of the large-input kernel (A/B runs)
  if (v == 0) {
    const char* e = getenv("B200SCN_GPL_WARPS");
    v = e ? atoi(e) : kLargeWarps;
    if (v < 1 || v > kLargeWarps) v = kLargeWarps;
  }
  return v;
}

// ---- large inputs: k_conv_gpl, one warp per tile, weights resident in shared memory ---------------
template <int Q, int MT, bool X3, int NB>
int launch_large(GplParams p, int passes, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_conv_gpl<Q, MT, X3, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kMaxSmem));
    configured = true;
  }
  const size_t fixed = w_smem_bytes(p.K, 16 * Q, MT, X3);
  int warps = large_warps();
  const size_t per_warp = acc_bytes(128, MT) + kRecRing * 128;
  while (warps > 1 && fixed + warps * per_warp > (size_t)kMaxSmem) --warps;
  p.warps = warps;
  p.split = 0;
  const int64_t tiles = cdiv(p.n_cap, 128);
  int64_t gx = cdiv(tiles, warps);
  int64_t gmax = sm_count() / passes;
  if (gmax < 1) gmax = 1;
  if (gx > gmax) gx = gmax;
  dim3 grid((unsigned)gx, (unsigned)passes);
  launch_pdl(k_conv_gpl<Q, MT, X3, NB>, grid, dim3(warps * 32), fixed + warps * per_warp, st, p);
  count_launches(1);
  B200_CHECK_LAUNCH("conv_gpl");
  return 0;
}

// groups per batch: 2 where the rows of two batches fit the register budget (16 -> 16 channel passes), else 1
template <int Q>
int launch_large_q(const GplParams& p, int MT, int passes, int x3, cudaStream_t st) {
  constexpr int NB = Q == 1 ? 2 : 1;
  if (MT == 1) return x3 ? launch_large<Q, 1, true, NB>(p, passes, st) : launch_large<Q, 1, false, NB>(p, passes, st);
  return x3 ? launch_large<Q, 2, true, 1>(p, passes, st) : launch_large<Q, 2, false, 1>(p, passes, st);
}

// ---- small inputs: k_conv_gpl_small ---------------------------------------------------------------
template <int MT, bool X3, int R, bool PACKED>
int launch_small(GplParams p, int passes, cudaStream_t st) {
  static thread_local bool configured = false;
  static int small_warps = 0;  // B200SCN_GPL_SMALL_WARPS (A/B runs)
  if (small_warps == 0) {
    const char* e = getenv("B200SCN_GPL_SMALL_WARPS");
    small_warps = e ? atoi(e) : 16;
    if (small_warps < 1 || small_warps > 16) small_warps = 16;
  }
  int warps = small_warps;
  while (warps > 4 && warps * acc_bytes(R, MT) > (size_t)200 * 1024) warps >>= 1;
  const size_t smem = warps * acc_bytes(R, MT);
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_conv_gpl_small<MT, X3, R, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(16 * acc_bytes(R, MT) > (size_t)kMaxSmem ? (size_t)kMaxSmem : 16 * acc_bytes(R, MT))));
    configured = true;
  }
  static int interleave = -1;  // B200SCN_GPL_INTERLEAVE=0: contiguous group range per warp (A/B runs)
  if (interleave < 0) {
    const char* e = getenv("B200SCN_GPL_INTERLEAVE");
    interleave = e ? (atoi(e) != 0) : 1;
  }
  p.warps = warps;
  p.split = interleave ? 2 : 1;
  // CTAs loop over tiles: a grid a few waves deep, not one CTA per tile of the row CAPACITY (in the
  // CUDA-graph regime most of those would find no live tile and only cost launch overhead)
  int64_t gx = cdiv(p.n_cap, R);
  int64_t gmax = cdiv((int64_t)sm_count() * 4, passes);
  if (gmax < sm_count()) gmax = sm_count();
  if (gx > gmax) gx = gmax;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)passes);
  launch_pdl(k_conv_gpl_small<MT, X3, R, PACKED>, grid, dim3(warps * 32), smem, st, p);
  count_launches(1);
  B200_CHECK_LAUNCH("conv_gpl_small");
  return 0;
}

template <int MT, bool X3, bool PACKED>
int launch_small_rp(const GplParams& p, int passes, int R, cudaStream_t st) {
  switch (R) {
    case 32: return launch_small<MT, X3, 32, PACKED>(p, passes, st);
    case 64: return launch_small<MT, X3, 64, PACKED>(p, passes, st);
    default: return launch_small<MT, X3, 128, PACKED>(p, passes, st);
  }
}
template <int MT, bool X3>
int launch_small_r(const GplParams& p, int passes, int R, cudaStream_t st) {
  return p.Wp ? launch_small_rp<MT, X3, true>(p, passes, R, st) : launch_small_rp<MT, X3, false>(p, passes, R, st);
}

// ---- weight packing -------------------------------------------------------------------------------
// job = int64[10]: {W, out, K, Cin, Cout, w_sk, w_si, w_so, kflip, 0}; out[((((k' Q + q) 2 + h) CoT + mt) 32 + lane) 4 + r]
// with the fragment maps of k_conv_gpl: lane = 4 (co & 7) + ((ci & 15) >> 2), h = (ci & 3) >> 1,
// r = 2 ((co >> 3) & 1) + (ci & 1), k' = kflip ? K-1-k : k.
__global__ void k_pack_weights(const int64_t* __restrict__ jobs) {
  const int64_t* j = jobs + (size_t)blockIdx.y * 10;
  const float* W = reinterpret_cast<const float*>(j[0]);
  float* out = reinterpret_cast<float*>(j[1]);
  const int K = (int)j[2], Cin = (int)j[3], Cout = (int)j[4];
  const int64_t w_sk = j[5], w_si = j[6], w_so = j[7];
  const int kflip = (int)j[8];
  const int Q = Cin >> 4, CoT = Cout >> 4;
  const bool co_fast = (w_so == 1);
  const int fast_n = co_fast ? Cout : Cin, slow_n = co_fast ? Cin : Cout;
  const int64_t total = (int64_t)K * Cin * Cout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(e % fast_n);
    const int64_t rest = e / fast_n;
    const int sl = (int)(rest % slow_n);
    const int k = (int)(rest / slow_n);
    const int ci = co_fast ? sl : f, co = co_fast ? f : sl;
    const float v = __ldg(W + k * w_sk + ci * w_si + co * w_so);
    const int kk = kflip ? (K - 1 - k) : k;
    const int q = ci >> 4, c16 = ci & 15, tt = c16 >> 2, rem = c16 & 3, h = rem >> 1;
    const int mt = co >> 4, odd = (co >> 3) & 1, gg = co & 7;
    out[((((size_t)(kk * Q + q) * 2 + h) * CoT + mt) * 32 + gg * 4 + tt) * 4 + odd * 2 + (rem & 1)] = v;
  }
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

/* tile rows the library recommends for a level of n rows */
extern "C" int b200scn_gpl_tile_rows(int64_t n) {
  static int64_t t128 = -1, t64 = -1;  // B200SCN_GPL_T128 / B200SCN_GPL_T64 override the row thresholds (A/B runs)
  if (t128 < 0) {
    const char* e = getenv("B200SCN_GPL_T128");
    t128 = e ? atoll(e) : 60000;  // (12288 until round 2: c3 10.9 -> 10.1 ms, c2 4.18 -> 4.16 ms, c5 unchanged)
    e = getenv("B200SCN_GPL_T64");
    t64 = e ? atoll(e) : 5120;
  }
  return n >= t128 ? 128 : (n >= t64 ? 64 : 32);
}

/* upper bound of the number of groups for n_cap rows (static regime allocation, fixed-stride lists) */
extern "C" int64_t b200scn_gpl_max_groups(int64_t n_cap, int K, int tile_rows) {
  const int64_t tiles = cdiv(n_cap < 1 ? 1 : n_cap, tile_rows);
  return tiles * (int64_t)K * (tile_rows / 16);
}

extern "C" int64_t b200scn_gpl_scratch_bytes(int64_t n_cap) {
  const int64_t tiles = cdiv(n_cap < 1 ? 1 : n_cap, 32);
  return (tiles + 8) * 4 + (cdiv(tiles, 1024) + 16) * 4 + 256;
}

template <bool FILL>
static void launch_build(int tile_rows, int grid, int threads, cudaStream_t st, const int32_t* table, int64_t ld,
                         const int32_t* parent, const uint8_t* kofs, int K, int64_t n_cap, const int32_t* n_dev,
                         int32_t* tile_cnt, const int32_t* tile_off, int32_t* tile_beg_out, int4* rec, int64_t cap,
                         int32_t* koff, int layout) {
  switch (tile_rows) {
    case 32:
      k_gpl_build<FILL, 32><<<grid, threads, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, tile_cnt, tile_off,
                                                      tile_beg_out, rec, cap, koff, layout);
      break;
    case 64:
      k_gpl_build<FILL, 64><<<grid, threads, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, tile_cnt, tile_off,
                                                      tile_beg_out, rec, cap, koff, layout);
      break;
    default:
      k_gpl_build<FILL, 128><<<grid, threads, 0, st>>>(table, ld, parent, kofs, K, n_cap, n_dev, tile_cnt, tile_off,
                                                       tile_beg_out, rec, cap, koff, layout);
  }
}

/* stage 0: tile_off[tiles + 1] (exclusive scan of the per-tile group counts) and total_dev[1].
 * stage 1: fill g_rec[cap_groups * 8] (int4 records) using tile_off.
 * stage 2: single pass with fixed-stride slots (cap_groups >= b200scn_gpl_max_groups): writes tile_off[tiles]
 *          (begin) and tile_end[tiles]; no scan, nothing to read back -- the static (CUDA graph) regime. */
extern "C" int b200scn_gpl_build(int stage, const int32_t* table, int64_t ld, const int32_t* parent,
                                 const uint8_t* kofs, int K, int64_t n_cap, const int32_t* n_dev, int tile_rows,
                                 int layout, int32_t* tile_off, int32_t* tile_end, int32_t* total_dev, void* g_rec,
                                 int64_t cap_groups, int32_t* koff, void* scratch, int64_t scratch_bytes,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(K >= 1 && K <= 64, "gpl_build: K must be 1..64");
  B200_CHECK_ARG(tile_rows == 32 || tile_rows == 64 || tile_rows == 128, "gpl_build: tile_rows must be 32, 64 or 128");
  B200_CHECK_ARG(layout == 0 || (layout == 1 && table != nullptr && parent == nullptr && (K & 1) && K >= 3 &&
                                 koff == nullptr),
                 "gpl_build: the centre-first layout (1) needs a submanifold table with an odd offset count and no koff");
  B200_CHECK_ARG(stage == 2 || scratch_bytes >= b200scn_gpl_scratch_bytes(n_cap), "gpl_build: scratch too small");
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
    if (n_cap > 0)
      launch_build<false>(tile_rows, (int)grid, warps * 32, st, table, ld, parent, kofs, K, n_cap, n_dev, tile_cnt,
                          nullptr, nullptr, nullptr, 0, nullptr, layout);
    // exclusive scan over all `tiles` counts (tiles beyond the live ones hold 0)
    if (int rc = scan_i32_public(tile_cnt, tile_off, tiles, nullptr, total_dev, tile_off, bsums, st)) return rc;
    count_launches(4);
  } else if (stage == 1) {
    if (n_cap > 0)
      launch_build<true>(tile_rows, (int)grid, warps * 32, st, table, ld, parent, kofs, K, n_cap, n_dev, nullptr,
                         tile_off, nullptr, rec, cap_groups, koff, layout);
    count_launches(1);
  } else {
    B200_CHECK_ARG(tile_end != nullptr, "gpl_build: stage 2 needs tile_end");
    B200_CHECK_ARG(cap_groups >= b200scn_gpl_max_groups(n_cap, K, tile_rows), "gpl_build: stage 2 needs worst-case capacity");
    if (n_cap > 0)
      launch_build<true>(tile_rows, (int)grid, warps * 32, st, table, ld, parent, kofs, K, n_cap, n_dev, tile_end,
                         nullptr, tile_off, rec, cap_groups, koff, layout);
    count_launches(1);
  }
  B200_CHECK_LAUNCH("gpl_build");
  return 0;
}

// Large inputs (enough 128-row tiles to give every SM a CTA of one-tile warps) with scanned lists and a weight
// slice that fits in shared memory next to >= 12 accumulators take k_conv_gpl; returns its pass width MT
// (output channels / 16 per pass) or 0 for the small-input kernel.  (Wider inputs take the small-input kernel,
// which walks any Cin in one launch: measured equal on the 768^3 / uf=32 configuration, with a third fewer
// launches.)
static int large_pass_width(int tile_rows, bool scanned, int64_t n_out_cap, int Cin, int Cout, int K, int x3) {
  static int64_t large_min = -1;  // B200SCN_GPL_LARGE_MIN_ROWS overrides the row threshold (tests)
  if (large_min < 0) {
    const char* e = getenv("B200SCN_GPL_LARGE_MIN_ROWS");
    large_min = e ? atoll(e) : (int64_t)128 * 8 * sm_count();
  }
  if (!(tile_rows == 128 && scanned && Cin <= 64 && n_out_cap >= large_min)) return 0;
  const int cot = Cout / 16;
  for (int mt = 2; mt >= 1; --mt)
    if (cot % mt == 0 && w_smem_bytes(K, Cin, mt, x3) + 12 * (acc_bytes(128, mt) + kRecRing * 128) <= (size_t)kMaxSmem)
      return mt;
  return 0;
}

/* Pair-list layout the convolution of these shapes wants (b200scn_gpl_build's `layout`): 0 = offset-major groups
 * only (every kernel reads it), 1 = centre offset first (the large-input kernel only).  B200SCN_GPL_DENSE=0 turns
 * the centre-first choice off (A/B runs). */
extern "C" int b200scn_conv_gpl_layout(int tile_rows, int scanned, int64_t n_out_cap, int Cin, int Cout, int K,
                                       int submanifold, int math_mode) {
  static int dense = -1;
  if (dense < 0) {
    const char* e = getenv("B200SCN_GPL_DENSE");
    dense = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!submanifold || !(K & 1) || K < 3 || Cin % 16 || Cout % 16) return 0;
  const int x3 = math_mode == B200SCN_MATH_TF32X3 ? 1 : 0;
  return large_pass_width(tile_rows, scanned != 0, n_out_cap, Cin, Cout, K, x3) ? dense : 0;
}

extern "C" int b200scn_pack_weights(const int64_t* jobs_dev, int n_jobs, int64_t max_elems, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_jobs <= 0) return 0;
  B200_CHECK_ARG(n_jobs <= 65535, "pack_weights: too many jobs");
  int64_t gx = cdiv(max_elems < 1 ? 1 : max_elems, 256 * 4);
  if (gx > 64) gx = 64;
  k_pack_weights<<<dim3((unsigned)gx, (unsigned)n_jobs), 256, 0, st>>>(jobs_dev);
  count_launches(1);
  B200_CHECK_LAUNCH("pack_weights");
  return 0;
}

extern "C" int b200scn_conv_gpl(const float* in, int64_t ldi, const float* W, const float* Wp, int64_t w_sk, int64_t w_si,
                                int64_t w_so, int kflip, int tile_rows, int layout, const int32_t* tile_off,
                                const int32_t* tile_end, const void* g_rec, int K, int64_t n_out_cap, const int32_t* n_out_dev,
                                int64_t n_out_expect, int Cin, int Cout, float* out, int64_t ldo, int accumulate,
                                const float* res, int64_t ldr, int math_mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(Cin % 16 == 0 && Cout % 16 == 0 && Cin >= 16 && Cout >= 16, "conv_gpl: channels must be multiples of 16");
  B200_CHECK_ARG(math_mode == B200SCN_MATH_TF32 || math_mode == B200SCN_MATH_TF32X3, "conv_gpl: tensor-core math modes only");
  B200_CHECK_ARG(ldi % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "conv_gpl: input rows must be 16-byte aligned");
  B200_CHECK_ARG(tile_rows == 32 || tile_rows == 64 || tile_rows == 128, "conv_gpl: tile_rows must be 32, 64 or 128");
  const int64_t lim = (1ll << 31) - 1;
  B200_CHECK_ARG(ldi < lim && ldo < lim && (int64_t)K * Cin * Cout < lim, "conv_gpl: stride out of 32-bit range");
  if (n_out_cap == 0) return 0;
  const int x3 = math_mode == B200SCN_MATH_TF32X3 ? 1 : 0;
  const int cot = Cout / 16;
  GplParams p;
  p.Wp = Wp; p.cot = Cout / 16;
  p.in = in; p.ldi = (int)ldi; p.W = W; p.w_sk = (int)w_sk; p.w_si = (int)w_si; p.w_so = (int)w_so; p.kflip = kflip;
  p.tile_off = tile_off; p.tile_end = tile_end; p.g_rec = reinterpret_cast<const int4*>(g_rec); p.Q = Cin / 16;
  p.K = K; p.n_cap = n_out_cap; p.n_dev = n_out_dev; p.out = out; p.ldo = (int)ldo; p.accumulate = accumulate;
  p.res = res; p.ldr = (int)ldr;
  p.warps = 8; p.split = 1; p.ones = 0xffffffffu; p.zero = 0u; p.layout = layout;

  const int MTL = large_pass_width(tile_rows, tile_end == nullptr, n_out_cap, Cin, Cout, K, x3);
  B200_CHECK_ARG(layout != 3 || x3, "conv_gpl: layout-3 pair lists are served in the tf32x3 mode only");
  if (MTL) {
    switch (p.Q) {
      case 1: return launch_large_q<1>(p, MTL, cot / MTL, x3, st);
      case 2: return launch_large_q<2>(p, MTL, cot / MTL, x3, st);
      case 3: return launch_large_q<3>(p, MTL, cot / MTL, x3, st);
      default: return launch_large_q<4>(p, MTL, cot / MTL, x3, st);
    }
  }
  B200_CHECK_ARG(layout == 0, "conv_gpl: a centre-first pair list (layout 1) only serves the large-input kernel "
                              "(ask b200scn_conv_gpl_layout first)");
  // small inputs: narrow passes unless the tiles alone already cover the SMs twice
  // (n_out_expect: how many rows the level really has when n_out_cap is only a buffer capacity)
  const int64_t n_live = (n_out_expect > 0 && n_out_expect < n_out_cap) ? n_out_expect : n_out_cap;
  const int64_t tiles = cdiv(n_live, tile_rows);
  static int force_mt = -1;  // B200SCN_GPL_SMALL_MT = 1 / 2 forces the pass width (A/B runs)
  if (force_mt < 0) {
    const char* e = getenv("B200SCN_GPL_SMALL_MT");
    force_mt = e ? atoi(e) : 0;
  }
  int MT = (cot % 2 == 0 && tiles * (cot / 2) >= sm_count()) ? 2 : 1;
  if (force_mt == 1 || (force_mt == 2 && cot % 2 == 0)) MT = force_mt;
  if (MT == 1)
    return x3 ? launch_small_r<1, true>(p, cot, tile_rows, st) : launch_small_r<1, false>(p, cot, tile_rows, st);
  return x3 ? launch_small_r<2, true>(p, cot / 2, tile_rows, st) : launch_small_r<2, false>(p, cot / 2, tile_rows, st);
}
