// Sparse convolution on the 5th-generation tensor cores: SubmanifoldConvolution / Convolution /
// NetworkInNetwork forward and dgrad (/root/reference/uresnet/models/uresnet_sparse.py:21-22 through scn.UNet) as a
// rule-table-driven implicit GEMM
//     out[o][co] = (res[o][co]) + sum_{k, ci} in[table[k][o]][ci] * W[k][ci][co]          (table[k][o] = -1: no term)
// with M = 128 consecutive output rows per CTA, N = Cout (<= 128 per pass), the contraction axis = the flattened
// (offset k, input channel) index walked in blocks of 32 (two 16-channel half rows, possibly of different offsets, so
// every Cin % 16 == 0 -- 16, 48, 80 ... -- runs on full 128-byte swizzle rows), issued as tcgen05.mma kind::tf32 by ONE
// thread, accumulator in TMEM, operands in shared memory in the canonical K-major SWIZZLE_128B layout.
//
// Why this shape for a sparse layer (profiles/README.md, round 2): the pair-list kernels (conv_gpl.cu) keep the rule
// pairs dense in the MMA's M dimension but pay for it with a shared-memory read-modify-write of every output row per
// offset and a re-fetch of W[k] per 16 pairs; at the U-Net's deeper levels (12-16 of 27 neighbours present) and for
// every level of a single 512^3 event they are bound by those, not by the tensor pipe.  Here a missing neighbour
// is a zero row of the A tile (the tensor pipe has the headroom), accumulators never leave TMEM, and W[k] is read
// once per (tile, offset).
//
//   producers (256 threads in 2 or 4 groups, one K block per group at a time): the tile's slice of the rule table sits
//       in shared memory; a group loads the block's A rows (8 lanes = one contiguous row: coalesced gathers), splits
//       them into tf32 high / low parts (3xTF32) and stores both tiles (STS.128, swizzled); the block's packed weights
//       arrive by cp.async.  `full[s]` is signalled (one arrival per warp) after cp.async.wait_group +
//       fence.proxy.async; the groups hide each other's load latency.
//   MMA (1 thread): waits `full[s]`, issues 4 k-steps x (A_lo B_hi + A_hi B_lo + A_hi B_hi), tcgen05.commit -> `empty[s]`.
//   epilogue (the 8 producer warps): tcgen05.ld 32x32b -> shared memory -> coalesced 16-byte stores (+ residual).
// Live row count from device memory (n_dev): tiles beyond it exit; rows beyond it are zero / not stored.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200scn {
namespace {

using namespace tc;

constexpr int kBM = 128;
constexpr int kProducers = 256;
constexpr int kThreads = kProducers + 32;
constexpr int kMaxK = 27;

struct SctParams {
  const float* in; int64_t ldi;
  const float* Wt; const float* Wt_lo;  // packed [n_kb][Cout][32]; Wt_lo = NULL: single-pass tf32
  const int32_t* table; int64_t ld_t;   // [K][ld_t] or NULL (identity, K == 1)
  int K;
  int64_t n_cap; const int32_t* n_dev;
  int Cin, Cout, n_pass;
  float* out; int64_t ldo;
  const float* res; int64_t ldr;
  uint32_t tmem_cols;
  int wide;      // 3xTF32: A_hi x [B_hi | B_lo] as one MMA of width 2N (needs 2N <= 256)
  int acc_cols;  // TMEM columns per accumulator tile (N, or 2N in wide mode)
  int n_sum;     // N-column slices the epilogue adds up = nacc * (wide ? 2 : 1)
  int drain;  // 1: group barrier before the proxy fence (A/B knob B200SCN_TC_DRAIN)
  int nacc;  // accumulator tiles in TMEM the MMAs rotate over (breaks the dependent-accumulate chain), summed by the epilogue
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// S: shared-memory stages (4: four producer groups, 2 / 3: two); L: unused (kept in the instantiation names)
template <bool X3, int S, int L>
__global__ void __launch_bounds__(kThreads, 1) k_sparse_conv_tc(const SctParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[S], empty_bar[S], done_bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.n_pass;
  constexpr uint32_t a_bytes = kBM * 128;
  const uint32_t b_bytes = (uint32_t)N * 128;
  const uint32_t stage_bytes = (a_bytes + b_bytes) * (X3 ? 2 : 1);
  const uint32_t a_lo_off = a_bytes, b_off = a_bytes * (X3 ? 2 : 1), b_lo_off = b_off + b_bytes;
  int32_t* sidx = reinterpret_cast<int32_t*>(smem + (size_t)S * stage_bytes);  // [K][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * N;
  const int c16n = p.Cin >> 4;
  const int F = p.K * c16n;          // 16-channel half rows along the contraction axis
  const int n_kb = (F + 1) >> 1;
  const int64_t o0 = (int64_t)blockIdx.x * kBM;

  // ---- prologue without global memory traffic: overlaps the previous kernel's tail (PDL) ----
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kProducers / 32 / ((S >= 4) ? 4 : 2)); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();
  const int n = live_rows(p.n_dev, p.n_cap);
  const bool live = o0 < n;

  if (live && warp < 8) {
    // ===== producers =====
    for (int e = tid; e < p.K * kBM; e += kProducers) {
      const int k = e >> 7, rr = e & 127;
      int v = -1;
      if (o0 + rr < n) v = p.table ? __ldg(p.table + (int64_t)k * p.ld_t + o0 + rr) : (int)(o0 + rr);
      sidx[e] = v;
    }
    named_bar_sync(1, kProducers);
    // The 8 producer warps form G groups; group g fills the K blocks g, g + G, ... on its own: request (rows by
    // LDG.128, weights by cp.async) -> wait -> split + store -> fence.proxy.async -> arrive.  A thread never has
    // loads in flight when it reaches the proxy fence (the first version prefetched rows L blocks ahead and spent
    // ~1500 cycles per block regardless of L: the fence waited for them); the latency of one group's requests is
    // covered by the other groups' blocks.  Inside a group thread t moves the 16-byte piece t & 7 of rows
    // (t >> 3) + (GT / 8) i: the 8 lanes of a row read 128 contiguous bytes (or two 64-byte runs).
    constexpr int G = (S >= 4) ? 4 : 2;
    constexpr int GT = kProducers / G;      // threads per group
    constexpr int NP = kBM * 8 / GT;        // 16-byte pieces per thread per block
    const int g = tid / GT, tg = tid - g * GT;
    const int j16 = tg & 7, h = j16 >> 2, r0 = tg >> 3;
    const int n_chunks = N * 8 * (X3 ? 2 : 1);  // 16-byte pieces of the block's weights
    for (int kb = g; kb < n_kb; kb += G) {
      const int s = kb % S;
      mbar_wait(&empty_bar[s], ((kb / S) & 1) ^ 1);
      {  // weights of the block: cp.async
        const float* whi = p.Wt + ((int64_t)kb * p.Cout + co0) * 32;
        const float* wlo = X3 ? p.Wt_lo + ((int64_t)kb * p.Cout + co0) * 32 : nullptr;
        const uint32_t b_dst = smem_u32(smem + (size_t)s * stage_bytes + b_off);
        for (int c = tg; c < n_chunks; c += GT) {
          const int lo = X3 ? (c >= N * 8) : 0;
          const int cc = lo ? c - N * 8 : c;
          const int nrow = cc >> 3, j = cc & 7;
          const float* src = (lo ? wlo : whi) + nrow * 32 + 4 * j;
          cp_async16_zfill(b_dst + (lo ? b_bytes : 0u) + (uint32_t)nrow * 128 + (((uint32_t)j ^ (uint32_t)(nrow & 7)) << 4), src, 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      const int f = 2 * kb + h;
      const bool in_range = f < F;
      int tap = 0, c16 = 0;
      if (in_range) { tap = f / c16n; c16 = f - tap * c16n; }
      const float* col = p.in + c16 * 16 + (j16 & 3) * 4;
      const int32_t* srow = sidx + tap * kBM + r0;
      float4 v[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int idx = in_range ? srow[(GT / 8) * i] : -1;
        v[i] = idx >= 0 ? __ldg(reinterpret_cast<const float4*>(col + (int64_t)idx * p.ldi)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int r = r0 + (GT / 8) * i;
        const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hi[e] = X3 ? __uint_as_float(__float_as_uint(x[e]) & 0xffffe000u) : x[e];
          lo[e] = x[e] - hi[e];
        }
        const uint32_t d = a_base + (uint32_t)r * 128 + ((((uint32_t)j16) ^ (uint32_t)(r & 7)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
        if (X3)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d + a_lo_off), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]) : "memory");
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");          // the weights of this block have landed
      // fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR's cost grows with the
      // shared-memory stores still in flight; a barrier among the group's threads drains them first
      if (p.drain) named_bar_sync(2 + g, GT);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);                      // one arrival per warp of the group
    }
    // ===== epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31, column half w >> 2; the tile is staged through
    // shared memory (the operand stages are free now) so that the global stores / residual loads are coalesced =====
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int er = (warp & 3) * 32 + lane;
    const int half_n = N >> 1;  // multiple of 8
    const int c_begin = (warp >> 2) * half_n, c_end = c_begin + half_n;
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int lds = N + 4;  // floats per staged row: an odd number of 16-byte units -> conflict-free row-wise STS.128
    float* stage_out = reinterpret_cast<float*>(smem);
    for (int c = c_begin; c < c_end; c += 8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
      for (int a = 0; a < p.n_sum; ++a) {  // accumulator tiles (and, in wide mode, their two halves) are N columns apart
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(t_row + (uint32_t)(a * N + c)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += __uint_as_float(v[j]);
      }
      float4* dst = reinterpret_cast<float4*>(stage_out + (size_t)er * lds + c);
      dst[0] = make_float4(f[0], f[1], f[2], f[3]);
      dst[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    named_bar_sync(1, kProducers);
    const int n4 = N >> 2;
    const bool vec = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                     (!p.res || ((p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0)));
    for (int e = tid; e < kBM * n4; e += kProducers) {
      const int row = e / n4, c4 = e - row * n4;
      const int64_t eo = o0 + row;
      if (eo >= n) break;  // rows are visited in ascending order by every thread
      float4 f = *reinterpret_cast<const float4*>(stage_out + (size_t)row * lds + 4 * c4);
      float* orow = p.out + eo * p.ldo + co0 + 4 * c4;
      const float* rrow = p.res ? p.res + eo * p.ldr + co0 + 4 * c4 : nullptr;
      if (vec) {
        if (rrow) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow));
          f.x += r4.x; f.y += r4.y; f.z += r4.z; f.w += r4.w;
        }
        *reinterpret_cast<float4*>(orow) = f;
      } else {
        if (rrow) { f.x += rrow[0]; f.y += rrow[1]; f.z += rrow[2]; f.w += rrow[3]; }
        orow[0] = f.x; orow[1] = f.y; orow[2] = f.z; orow[3] = f.w;
      }
    }
  } else if (live && lane == 0) {
    // ===== MMA issuer =====
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
    // N >> 3 in bits 17-22, M >> 4 in bits 24-28
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    const uint32_t idesc_wide = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    int rot = 0, issued = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % S;
      mbar_wait(&full_bar[s], (kb / S) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes), b_addr = a_addr + b_off;
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // UMMA K = 8 tf32 = 32 bytes along the swizzled row
        // consecutive MMAs go to different accumulator tiles: each tile's first MMA overwrites (accumulate = 0)
        auto next_acc = [&](uint32_t& d, uint32_t& acc) {
          d = tmem_base + (uint32_t)(rot * p.acc_cols);
          acc = issued >= p.nacc ? 1u : 0u;
          ++issued;
          if (++rot == p.nacc) rot = 0;
        };
        uint32_t d, acc;
        if (X3 && p.wide) {
          // two MMAs per k-step instead of three: the low weight tile sits right behind the high one in shared
          // memory, so A_hi x [B_hi | B_lo] is ONE instruction of width 2N (columns [N, 2N) of the accumulator hold
          // the A_hi B_lo sum, added in the epilogue); A_lo x B_hi goes to columns [0, N).  The 128 x 32-byte A
          // slice is read twice per k-step instead of three times -- the tensor core's shared-memory reads are
          // what bounds this kernel (profiles/README.md round 2).
          next_acc(d, acc);
          umma_tf32(d, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc_wide, acc);
          umma_tf32(d, umma_desc(a_addr + a_lo_off + k * 32), umma_desc(b_addr + k * 32), idesc, 1u);
        } else if (X3) {
          next_acc(d, acc);
          umma_tf32(d, umma_desc(a_addr + a_lo_off + k * 32), umma_desc(b_addr + k * 32), idesc, acc);
          next_acc(d, acc);
          umma_tf32(d, umma_desc(a_addr + k * 32), umma_desc(a_addr + b_lo_off + k * 32), idesc, acc);
          next_acc(d, acc);
          umma_tf32(d, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, acc);
        } else {
          next_acc(d, acc);
          umma_tf32(d, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, acc);
        }
      }
      umma_commit(&empty_bar[s]);
    }
    umma_commit(&done_bar);
  }
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// Packed weights: job = int64[12] {W, Wt, Wt_lo, K, Cin, Cout, w_sk, w_si, w_so, kflip, 0, 0};
//   Wt[kb][co][j] = W[k'(f)][16 c16(f) + (j & 15)][co],  f = 2 kb + (j >> 4),  tap(f) = f / (Cin / 16),
//   k' = kflip ? K - 1 - tap : tap;  f >= K Cin / 16 (odd tail): 0.  Wt = tf32 high part, Wt_lo = exact remainder.
__global__ void k_pack_wt_tc(const int64_t* __restrict__ jobs) {
  const int64_t* j = jobs + (size_t)blockIdx.y * 12;
  const float* W = reinterpret_cast<const float*>(j[0]);
  float* Wt = reinterpret_cast<float*>(j[1]);
  float* Wt_lo = reinterpret_cast<float*>(j[2]);
  const int K = (int)j[3], Cin = (int)j[4], Cout = (int)j[5];
  const int64_t w_sk = j[6], w_si = j[7], w_so = j[8];
  const int kflip = (int)j[9];
  const int c16n = Cin >> 4, F = K * c16n, n_kb = (F + 1) >> 1;
  const int64_t total = (int64_t)n_kb * Cout * 32;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int jj = (int)(e & 31);
    const int64_t rest = e >> 5;
    const int co = (int)(rest % Cout);
    const int kb = (int)(rest / Cout);
    const int f = 2 * kb + (jj >> 4);
    float v = 0.f;
    if (f < F) {
      const int tap = f / c16n, c16 = f - tap * c16n;
      const int kk = kflip ? (K - 1 - tap) : tap;
      v = __ldg(W + kk * w_sk + (int64_t)(16 * c16 + (jj & 15)) * w_si + co * w_so);
    }
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    Wt[e] = hi;
    Wt_lo[e] = v - hi;
  }
}

template <bool X3, int S, int L>
int launch_sct(const SctParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_sparse_conv_tc<X3, S, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    configured = true;
  }
  launch_pdl(k_sparse_conv_tc<X3, S, L>, grid, dim3(kThreads), smem, st, p);
  return 0;
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

/* floats of ONE packed copy (high parts; the low-part copy has the same size) */
extern "C" int64_t b200scn_conv_tc_packed_elems(int K, int Cin, int Cout) {
  const int F = K * (Cin / 16);
  return (int64_t)((F + 1) / 2) * Cout * 32;
}

extern "C" int b200scn_pack_weights_tc(const int64_t* jobs_dev, int n_jobs, int64_t max_elems, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_jobs <= 0) return 0;
  B200_CHECK_ARG(n_jobs <= 65535, "pack_weights_tc: too many jobs");
  int64_t gx = cdiv(max_elems < 1 ? 1 : max_elems, 256 * 4);
  if (gx > 64) gx = 64;
  k_pack_wt_tc<<<dim3((unsigned)gx, (unsigned)n_jobs), 256, 0, st>>>(jobs_dev);
  count_launches(1);
  B200_CHECK_LAUNCH("pack_weights_tc");
  return 0;
}

/* 1 when b200scn_conv_tc takes this layer shape */
extern "C" int b200scn_conv_tc_supported(int K, int Cin, int Cout) {
  return (K >= 1 && K <= kMaxK && Cin >= 16 && Cin % 16 == 0 && Cout >= 16 && Cout % 16 == 0) ? 1 : 0;
}

extern "C" int b200scn_conv_tc(const float* in, int64_t ldi, const float* Wt, const float* Wt_lo, const int32_t* table,
                               int64_t ld_t, int K, int64_t n_cap, const int32_t* n_dev, int64_t n_expect, int Cin,
                               int Cout, float* out, int64_t ldo, const float* res, int64_t ldr, int math_mode,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(b200scn_conv_tc_supported(K, Cin, Cout), "conv_tc: needs K <= 27 and channel counts that are multiples of 16");
  B200_CHECK_ARG(table != nullptr || K == 1, "conv_tc: no table only for K == 1");
  B200_CHECK_ARG(ldi % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "conv_tc: input rows must be 16-byte aligned");
  B200_CHECK_ARG(math_mode == B200SCN_MATH_TF32 || math_mode == B200SCN_MATH_TF32X3, "conv_tc: tensor-core math modes only");
  B200_CHECK_ARG(n_cap < (1ll << 31), "conv_tc: too many rows");
  const bool x3 = math_mode == B200SCN_MATH_TF32X3;
  B200_CHECK_ARG(Wt != nullptr && (!x3 || Wt_lo != nullptr), "conv_tc: packed weights missing");
  if (n_cap == 0) return 0;
  // output channels per CTA: the largest multiple of 16 <= 128 that divides Cout
  int n_pass = Cout;
  if (n_pass > 128) {
    n_pass = 128;
    while (Cout % n_pass != 0) n_pass -= 16;
  }
  SctParams p;
  p.in = in; p.ldi = ldi; p.Wt = Wt; p.Wt_lo = x3 ? Wt_lo : nullptr; p.table = table; p.ld_t = ld_t; p.K = K; p.n_cap = n_cap;
  p.n_dev = n_dev; p.Cin = Cin; p.Cout = Cout; p.n_pass = n_pass; p.out = out; p.ldo = ldo; p.res = res; p.ldr = ldr;
  static int nacc_knob = -1;  // B200SCN_TC_NACC: accumulator tiles (1, 2 or 4; A/B runs)
  if (nacc_knob < 0) {
    const char* e = getenv("B200SCN_TC_NACC");
    nacc_knob = e ? atoi(e) : 1;  // (measured: no effect on the step time; one tile keeps TMEM small)
    if (nacc_knob != 1 && nacc_knob != 2 && nacc_knob != 4) nacc_knob = 1;
  }
  static int drain_knob = -1;
  if (drain_knob < 0) {
    const char* e = getenv("B200SCN_TC_DRAIN");
    drain_knob = e ? (atoi(e) ? 1 : 0) : 1;
  }
  p.drain = drain_knob;
  static int wide_knob = -1;  // B200SCN_TC_WIDE=0: three MMAs per k-step (A/B runs)
  if (wide_knob < 0) {
    const char* e = getenv("B200SCN_TC_WIDE");
    wide_knob = e ? (atoi(e) ? 1 : 0) : 1;
  }
  p.wide = (x3 && wide_knob && 2 * n_pass <= 256) ? 1 : 0;
  p.acc_cols = p.wide ? 2 * n_pass : n_pass;
  p.nacc = nacc_knob;
  while (p.nacc > 1 && p.nacc * p.acc_cols > 512) p.nacc >>= 1;
  p.n_sum = p.nacc * (p.wide ? 2 : 1);
  // (every accumulator tile is written: a K block issues >= 4 MMAs and there are at most 4 tiles)
  uint32_t cols = 32;
  while ((int)cols < p.nacc * p.acc_cols) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage = (size_t)(kBM * 128 + n_pass * 128) * (x3 ? 2 : 1);
  const size_t tail = (size_t)K * kBM * 4 + 1024;
  // rows the level really holds (capacity-sized buffers in the CUDA-graph regime): two CTAs per SM when one wave of
  // 128-row tiles would not cover them otherwise
  const int64_t n_live = (n_expect > 0 && n_expect < n_cap) ? n_expect : n_cap;
  const int64_t tiles_live = cdiv(n_live, kBM) * (Cout / n_pass);
  static int two_knob = -1;  // B200SCN_TC_TWO=0/1: never / whenever it fits (A/B runs)
  if (two_knob < 0) {
    const char* e = getenv("B200SCN_TC_TWO");
    two_knob = e ? (atoi(e) ? 1 : 0) : 2;
  }
  bool two_per_sm = tiles_live > sm_count() && 2 * stage + tail <= 110 * 1024;
  if (two_knob == 0) two_per_sm = false;
  if (two_knob == 1) two_per_sm = 2 * stage + tail <= 110 * 1024;
  dim3 grid((unsigned)cdiv(n_cap, kBM), (unsigned)(Cout / n_pass));
  int rc;
  if (two_per_sm) {
    const size_t smem = 2 * stage + tail;
    rc = x3 ? launch_sct<true, 2, 1>(p, grid, smem, st) : launch_sct<false, 2, 1>(p, grid, smem, st);
  } else if (4 * stage + tail <= 220 * 1024) {
    const size_t smem = 4 * stage + tail;
    rc = x3 ? launch_sct<true, 4, 2>(p, grid, smem, st) : launch_sct<false, 4, 2>(p, grid, smem, st);
  } else {
    const size_t smem = 3 * stage + tail;
    B200_CHECK_ARG(smem <= 224 * 1024, "conv_tc: tile does not fit shared memory");
    rc = x3 ? launch_sct<true, 3, 1>(p, grid, smem, st) : launch_sct<false, 3, 1>(p, grid, smem, st);
  }
  if (rc) return rc;
  count_launches(1);
  B200_CHECK_LAUNCH("conv_tc");
  return 0;
}

// ======================================================================================================
// Weight gradient on the tensor cores:  dW[k][ci][co] (+)= sum_o in[table[k][o]][ci] * dout[o][co]
// (SubmanifoldConvolution / Convolution / NetworkInNetwork backward, R/uresnet/trainval.py:21-30 via autograd).
// Per offset this is a GEMM that CONTRACTS OVER THE ROWS, so both operands are MN-major for the tensor core:
//   A = gathered input rows: [rows][32 flattened (offset, channel) columns] tiles like the forward kernel's, read
//       with the roles of M and K swapped -- M = 128 flattened (offset, channel) elements = four 32-column tiles
//       (leading byte offset = tile stride), K = 8 rows per tcgen05.mma.  MN-major 32-bit operands must use the
//       SWIZZLE_128B_BASE32B layout (4-row atoms, 32-byte swizzle granularity; stride byte offset 512);
//   B = dout rows [rows][Cout], staged the same way (N = Cout padded to a multiple of 32);
//   D = 128 x N fp32 in TMEM, accumulated over ALL row tiles of the CTA's row range.
// grid = (M groups of 128 flattened elements) x (row splits); a CTA writes its 128 x Cout block of the partial
// dW of its row split; k_wgrad_tc_fold sums the row splits in fixed order (deterministic), optionally into a
// persistent gradient buffer.  Producer groups as in the forward kernel; 3xTF32 with high / low tiles of both
// operands.
namespace b200scn {
namespace {

constexpr int kWR = 32;  // rows per stage (4 k-steps of 8 rows)

struct SwtParams {
  const float* in; int64_t ldi;
  const float* dout; int64_t ldd;
  const int32_t* table; int64_t ld_t;
  int K;
  int64_t n_cap; const int32_t* n_dev;
  int Cin, Cout, n_pass, N;  // output channels per pass (blockIdx.z), N = n_pass rounded up to a multiple of 32
  float* part; int64_t part_stride;  // [row splits][K * Cin * Cout]
  uint32_t tmem_cols;
};

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B: the one layout MN-major 32-bit operands may use
  return d;
}
// byte offset of 16-byte piece j of row r inside a [rows][128 B] tile in that layout: Swizzle<2,5,2> -- the 32-byte
// chunk index (bits 5-6) is XORed with the row index modulo 4 (bits 7-8); atoms are 4 rows x 128 bytes
__device__ __forceinline__ uint32_t mn32_off(int r, int j) {
  return (uint32_t)r * 128u + ((((uint32_t)(j >> 1)) ^ ((uint32_t)r & 3u)) << 5) + (((uint32_t)j & 1u) << 4);
}

template <bool X3, int S, int G>
__global__ void __launch_bounds__(kThreads, 1) k_sparse_wgrad_tc(const SwtParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[S], empty_bar[S], done_bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.N, nch = N >> 5;                       // dout chunk tiles of 32 channels
  constexpr uint32_t tile_bytes = kWR * 128;             // one [32 rows][32 columns] tile
  constexpr uint32_t a_bytes = 4 * tile_bytes;           // four tiles = 128 flattened (offset, channel) elements
  const uint32_t b_bytes = (uint32_t)nch * tile_bytes;
  const uint32_t stage_bytes = (a_bytes + b_bytes) * (X3 ? 2 : 1);
  const uint32_t a_lo_off = a_bytes, b_off = a_bytes * (X3 ? 2 : 1), b_lo_off = b_off + b_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mg = blockIdx.x, rs = blockIdx.y, RS = gridDim.y;
  const int co0 = blockIdx.z * p.n_pass;
  const int c16n = p.Cin >> 4;
  const int F = p.K * c16n;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kProducers / 32 / G); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();
  const int n = live_rows(p.n_dev, p.n_cap);
  const int tiles = (n + kWR - 1) / kWR;
  const int per = (tiles + RS - 1) / RS;
  const int t_begin = rs * per, t_end = min(tiles, t_begin + per);
  const int n_it = max(t_end - t_begin, 0);

  if (warp < 8) {
    // ===== producers: group g fills the row tiles t_begin + g, + G, ... =====
    constexpr int GT = kProducers / G;
    constexpr int NPA = 4 * kWR * 8 / GT;  // A pieces (16 bytes) per thread per stage
    const int g = tid / GT, tg = tid - g * GT;
    const int j16 = tg & 7, h = j16 >> 2;
    // A piece i of this thread: cid = tg + GT i -> tile (cid >> 8), row ((cid >> 3) & 31)
    int tapv[NPA], colv[NPA], rowv[NPA], tilev[NPA];
#pragma unroll
    for (int i = 0; i < NPA; ++i) {
      const int cid = tg + GT * i;
      tilev[i] = cid >> 8;
      rowv[i] = (cid >> 3) & 31;
      const int f = 8 * mg + 2 * tilev[i] + h;
      if (f < F) { tapv[i] = f / c16n; colv[i] = (f - tapv[i] * c16n) * 16 + (j16 & 3) * 4; }
      else { tapv[i] = -1; colv[i] = 0; }
    }
    const int nb_pieces = kWR * nch * 8;  // B pieces per stage
    auto idx_of = [&](int i, int64_t o0) -> int {
      const int64_t o = o0 + rowv[i];
      if (tapv[i] < 0 || o >= n) return -1;
      return p.table ? __ldg(p.table + (int64_t)tapv[i] * p.ld_t + o) : (int)o;
    };
    int idx_next[NPA];
    if (g < n_it) {
#pragma unroll
      for (int i = 0; i < NPA; ++i) idx_next[i] = idx_of(i, (int64_t)(t_begin + g) * kWR);
    }
    for (int it = g; it < n_it; it += G) {
      const int s = it % S;
      const int64_t o0 = (int64_t)(t_begin + it) * kWR;
      mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
      float4 va[NPA];
#pragma unroll
      for (int i = 0; i < NPA; ++i) {
        const int idx = idx_next[i];
        va[i] = idx >= 0 ? __ldg(reinterpret_cast<const float4*>(p.in + (int64_t)idx * p.ldi + colv[i]))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (it + G < n_it) {  // table entries of this group's next tile: in flight under this tile's work
#pragma unroll
        for (int i = 0; i < NPA; ++i) idx_next[i] = idx_of(i, (int64_t)(t_begin + it + G) * kWR);
      }
      const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes), b_base = a_base + b_off;
      auto split_store = [&](uint32_t d, uint32_t lo_off, const float4& v4) {
        const float x[4] = {v4.x, v4.y, v4.z, v4.w};
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hi[e] = X3 ? __uint_as_float(__float_as_uint(x[e]) & 0xffffe000u) : x[e];
          lo[e] = x[e] - hi[e];
        }
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
        if (X3)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d + lo_off), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]) : "memory");
      };
      // dout rows: contiguous, zero beyond the live rows / the real channel count
      for (int c = tg; c < nb_pieces; c += GT) {
        const int jj = c & 7, r = (c >> 3) & 31, ch = c >> 8;
        const int64_t o = o0 + r;
        const int col = ch * 32 + jj * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o < n && col < p.n_pass) v = __ldg(reinterpret_cast<const float4*>(p.dout + o * p.ldd + co0 + col));
        split_store(b_base + (uint32_t)ch * tile_bytes + mn32_off(r, jj), b_bytes, v);
      }
#pragma unroll
      for (int i = 0; i < NPA; ++i)
        split_store(a_base + (uint32_t)tilev[i] * tile_bytes + mn32_off(rowv[i], j16), a_lo_off, va[i]);
      named_bar_sync(2 + g, GT);  // drain the group's shared-memory stores before the (MEMBAR-backed) proxy fence
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
    // ===== epilogue: TMEM lane m = flattened element 128 mg + m -> dW[tap][ci][0 .. Cout) =====
    if (n_it > 0) {
      mbar_wait(&done_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int m = (warp & 3) * 32 + lane;
    const int e = 128 * mg + m, f = e >> 4;
    const int half_n = N >> 1;  // multiple of 16
    const int c_begin = (warp >> 2) * half_n, c_end = c_begin + half_n;
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // every lane of the warp takes part in the (warp-wide, aligned) TMEM loads; only the stores are predicated
    const bool valid = f < F;
    const int tap = valid ? f / c16n : 0, ci = valid ? (f - tap * c16n) * 16 + (e & 15) : 0;
    float* orow = p.part + (int64_t)rs * p.part_stride + ((int64_t)tap * p.Cin + ci) * p.Cout + co0;
    for (int c = c_begin; c < c_end; c += 8) {
      uint32_t v[8];
      if (n_it > 0) {  // uniform over the CTA
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(t_row + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (X3) {  // columns [N, 2N): the A_hi x B_lo sums
          uint32_t w[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                       : "r"(t_row + (uint32_t)(N + c)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0u;
      }
      if (valid && c < p.n_pass) {  // n_pass is a multiple of 16: an 8-column piece is entirely inside or outside
        *reinterpret_cast<float4*>(orow + c) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
        *reinterpret_cast<float4*>(orow + c + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (lane == 0 && n_it > 0) {
    // ===== MMA issuer: D[128 x N] += A^T B over the 32 rows of every stage; both operands MN-major =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_wide = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int it = 0; it < n_it; ++it) {
      const int s = it % S;
      mbar_wait(&full_bar[s], (it / S) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes), b_addr = a_addr + b_off;
#pragma unroll
      for (int ks = 0; ks < kWR / 8; ++ks) {
        const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
        const uint32_t ko = ks * 1024;  // 8 rows of 128 bytes
        if (X3) {
          // two MMAs per k-step (see k_sparse_conv_tc): A_hi x [B_hi | B_lo] as one instruction of width 2N --
          // the low dout tiles follow the high ones in shared memory -- then A_lo x B_hi into columns [0, N)
          umma_tf32(tmem_base, umma_desc_mn(a_addr + ko, tile_bytes, 512), umma_desc_mn(b_addr + ko, tile_bytes, 512), idesc_wide, acc);
          umma_tf32(tmem_base, umma_desc_mn(a_addr + a_lo_off + ko, tile_bytes, 512), umma_desc_mn(b_addr + ko, tile_bytes, 512), idesc, 1u);
        } else {
          umma_tf32(tmem_base, umma_desc_mn(a_addr + ko, tile_bytes, 512), umma_desc_mn(b_addr + ko, tile_bytes, 512), idesc, acc);
        }
      }
      umma_commit(&empty_bar[s]);
    }
    umma_commit(&done_bar);
  }
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

__global__ void k_wgrad_tc_fold(const float* __restrict__ part, int64_t part_stride, int RS, int64_t elems,
                                float* __restrict__ dW, int accumulate) {
  pdl_wait();
  const int64_t n4 = elems >> 2;  // elems = K * Cin * Cout is a multiple of 256
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    float4 s = accumulate ? reinterpret_cast<const float4*>(dW)[e] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < RS; ++r) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(part + (int64_t)r * part_stride) + e);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(dW)[e] = s;
  }
}

int wgrad_tc_splits(int64_t n_rows, int K, int Cin) {
  const int n_mg = (K * (Cin / 16) + 7) / 8;
  const int64_t tiles = cdiv(n_rows < 1 ? 1 : n_rows, kWR);
  int64_t rs = cdiv((int64_t)sm_count(), n_mg);
  if (rs > tiles) rs = tiles;
  if (rs > 64) rs = 64;
  if (rs < 1) rs = 1;
  return (int)rs;
}

template <bool X3, int S, int G>
int launch_swt(const SwtParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_sparse_wgrad_tc<X3, S, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    configured = true;
  }
  launch_pdl(k_sparse_wgrad_tc<X3, S, G>, grid, dim3(kThreads), smem, st, p);
  return 0;
}

}  // namespace
}  // namespace b200scn

/* scratch for b200scn_wgrad_tc: the per-row-split partial dW blocks.  n_rows: the rows the level typically holds
 * (n_expect when known, else the capacity) -- the SAME value must be passed to b200scn_wgrad_tc. */
extern "C" int64_t b200scn_wgrad_tc_scratch_bytes(int64_t n_rows, int K, int Cin, int Cout) {
  return (int64_t)wgrad_tc_splits(n_rows, K, Cin) * K * Cin * Cout * 4 + 256;
}

extern "C" int b200scn_wgrad_tc(const float* in, int64_t ldi, const float* dout, int64_t ldd, const int32_t* table,
                                int64_t ld_t, int K, int64_t n_cap, const int32_t* n_dev, int64_t n_expect, int Cin,
                                int Cout, float* dW, int accumulate, int math_mode, void* scratch,
                                int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(b200scn_conv_tc_supported(K, Cin, Cout), "wgrad_tc: needs K <= 27 and channel counts that are multiples of 16");
  B200_CHECK_ARG(table != nullptr || K == 1, "wgrad_tc: no table only for K == 1");
  B200_CHECK_ARG(ldi % 4 == 0 && ldd % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(dout)) & 15) == 0,
                 "wgrad_tc: rows must be 16-byte aligned");
  B200_CHECK_ARG(math_mode == B200SCN_MATH_TF32 || math_mode == B200SCN_MATH_TF32X3, "wgrad_tc: tensor-core math modes only");
  B200_CHECK_ARG(n_cap < (1ll << 31), "wgrad_tc: too many rows");
  const int64_t elems = (int64_t)K * Cin * Cout;
  const int64_t n_rows = (n_expect > 0 && n_expect < n_cap) ? n_expect : n_cap;
  const int RS = wgrad_tc_splits(n_rows, K, Cin);
  B200_CHECK_ARG(scratch_bytes >= (int64_t)RS * elems * 4, "wgrad_tc: scratch too small");
  if (n_cap == 0) {
    if (!accumulate) B200_CUDA(cudaMemsetAsync(dW, 0, (size_t)elems * 4, st));
    return 0;
  }
  const bool x3 = math_mode == B200SCN_MATH_TF32X3;
  SwtParams p;
  p.in = in; p.ldi = ldi; p.dout = dout; p.ldd = ldd; p.table = table; p.ld_t = ld_t; p.K = K; p.n_cap = n_cap; p.n_dev = n_dev;
  int n_pass = Cout;  // output channels per CTA: the largest multiple of 16 <= 128 that divides Cout
  if (n_pass > 128) {
    n_pass = 128;
    while (Cout % n_pass != 0) n_pass -= 16;
  }
  p.Cin = Cin; p.Cout = Cout; p.n_pass = n_pass; p.N = (n_pass + 31) / 32 * 32;
  p.part = reinterpret_cast<float*>(scratch); p.part_stride = elems;
  uint32_t cols = 32;
  while ((int)cols < p.N * (x3 ? 2 : 1)) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage = (size_t)(4 + p.N / 32) * kWR * 128 * (x3 ? 2 : 1);
  const int n_mg = (K * (Cin / 16) + 7) / 8;
  dim3 grid((unsigned)n_mg, (unsigned)RS, (unsigned)(Cout / n_pass));
  int rc;
  if (4 * stage + 1024 <= 200 * 1024) {
    const size_t smem = 4 * stage + 1024;
    rc = x3 ? launch_swt<true, 4, 4>(p, grid, smem, st) : launch_swt<false, 4, 4>(p, grid, smem, st);
  } else {
    const size_t smem = 3 * stage + 1024;
    rc = x3 ? launch_swt<true, 3, 2>(p, grid, smem, st) : launch_swt<false, 3, 2>(p, grid, smem, st);
  }
  if (rc) return rc;
  int64_t gx = cdiv(elems / 4, 256);
  if (gx > 2 * sm_count()) gx = 2 * sm_count();
  launch_pdl(k_wgrad_tc_fold, dim3((unsigned)gx), dim3(256), 0, st, (const float*)p.part, p.part_stride, RS, elems, dW, accumulate);
  count_launches(2);
  B200_CHECK_LAUNCH("wgrad_tc");
  return 0;
}
