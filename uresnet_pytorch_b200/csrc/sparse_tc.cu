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
//   producers (256 threads): thread (r, h) owns half h (16 channels) of A row r.  The tile's slice of the rule table
//       sits in shared memory; per K block the thread loads its half row (LDG.128 x4, issued kLag blocks ahead),
//       splits it into tf32 high / low parts (3xTF32) and stores both tiles (STS.128); the block's packed weights
//       arrive by cp.async.  `full[s]` is signalled after cp.async.wait_group + fence.proxy.async.
//   MMA (1 thread): waits `full[s]`, issues 4 k-steps x (A_lo B_hi + A_hi B_lo + A_hi B_hi), tcgen05.commit -> `empty[s]`.
//   epilogue (the 8 producer warps): tcgen05.ld 32x32b, optional residual add, 16-byte stores.
// Live row count from device memory (n_dev): tiles beyond it exit; rows beyond it are zero / not stored.
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
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// S: shared-memory stages, L: blocks between a producer's requests (rows, weights) and their completion
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
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kProducers); mbar_init(&empty_bar[s], 1); }
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
    const int r = tid & 127, h = tid >> 7;
    const int64_t o = o0 + r;
    for (int e = tid; e < p.K * kBM; e += kProducers) {
      const int k = e >> 7, rr = e & 127;
      int v = -1;
      if (o0 + rr < n) v = p.table ? __ldg(p.table + (int64_t)k * p.ld_t + o0 + rr) : (int)(o0 + rr);
      sidx[e] = v;
    }
    named_bar_sync(1, kProducers);
    const uint32_t sw = (uint32_t)(r & 7);
    const int n_chunks = N * 8 * (X3 ? 2 : 1);  // 16-byte pieces of the block's weights
    auto copy_b = [&](int kb, int s) {
      const float* whi = p.Wt + ((int64_t)kb * p.Cout + co0) * 32;
      const float* wlo = X3 ? p.Wt_lo + ((int64_t)kb * p.Cout + co0) * 32 : nullptr;
      const uint32_t b_dst = smem_u32(smem + (size_t)s * stage_bytes + b_off);
      for (int c = tid; c < n_chunks; c += kProducers) {
        const int lo = X3 ? (c >= N * 8) : 0;
        const int cc = lo ? c - N * 8 : c;
        const int nrow = cc >> 3, j = cc & 7;
        const float* src = (lo ? wlo : whi) + nrow * 32 + 4 * j;
        cp_async16_zfill(b_dst + (lo ? b_bytes : 0u) + (uint32_t)nrow * 128 + (((uint32_t)j ^ (uint32_t)(nrow & 7)) << 4), src, 16);
      }
    };
    auto load_half = [&](int kb, float4 (&v)[4]) {
      const int f = 2 * kb + h;
      const float* src = nullptr;
      if (f < F) {
        const int tap = f / c16n, c16 = f - tap * c16n;
        const int idx = sidx[tap * kBM + r];
        if (idx >= 0) src = p.in + (int64_t)idx * p.ldi + c16 * 16;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = src ? __ldg(reinterpret_cast<const float4*>(src) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 q[L + 1][4];  // q[0] = oldest request
    for (int it = 0; it < n_kb + L; ++it) {
      float4 fresh[4];
      if (it < n_kb) {
        const int s = it % S;
        mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
        copy_b(it, s);
        load_half(it, fresh);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      const int kb = it - L;
      if (kb >= 0) {
        const int s = kb % S;
        const uint32_t a_dst = smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)r * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x[4] = {q[0][j].x, q[0][j].y, q[0][j].z, q[0][j].w};
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[e] = X3 ? __uint_as_float(__float_as_uint(x[e]) & 0xffffe000u) : x[e];
            lo[e] = x[e] - hi[e];
          }
          const uint32_t d = a_dst + ((((uint32_t)(4 * h + j)) ^ sw) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
          if (X3)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d + a_lo_off), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]) : "memory");
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(L) : "memory");  // the weights of block kb have landed
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
        mbar_arrive(&full_bar[s]);
      }
      // age the request queue
#pragma unroll
      for (int u = 0; u < L; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) q[u][j] = (u + 1 < L) ? q[u + 1][j] : fresh[j];
    }
    // ===== epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31, column half w >> 2 =====
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int er = (warp & 3) * 32 + lane;
    const int64_t eo = o0 + er;
    const int half_n = N >> 1;  // multiple of 8
    const int c_begin = (warp >> 2) * half_n, c_end = c_begin + half_n;
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float* orow = p.out + eo * p.ldo + co0;
    const float* rrow = p.res ? p.res + eo * p.ldr + co0 : nullptr;
    const bool vec = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                     (!p.res || ((p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0)));
    for (int c = c_begin; c < c_end; c += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(t_row + (uint32_t)c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (eo < n) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j]);
        if (vec) {
          if (rrow) {
            const float4 r0 = *reinterpret_cast<const float4*>(rrow + c), r1 = *reinterpret_cast<const float4*>(rrow + c + 4);
            f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
            f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
          }
          *reinterpret_cast<float4*>(orow + c) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(orow + c + 4) = make_float4(f[4], f[5], f[6], f[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) orow[c + j] = f[j] + (rrow ? rrow[c + j] : 0.f);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (live && lane == 0) {
    // ===== MMA issuer =====
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
    // N >> 3 in bits 17-22, M >> 4 in bits 24-28
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % S;
      mbar_wait(&full_bar[s], (kb / S) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes), b_addr = a_addr + b_off;
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // UMMA K = 8 tf32 = 32 bytes along the swizzled row
        const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
        if (X3) {
          umma_tf32(tmem_base, umma_desc(a_addr + a_lo_off + k * 32), umma_desc(b_addr + k * 32), idesc, first);
          umma_tf32(tmem_base, umma_desc(a_addr + k * 32), umma_desc(a_addr + b_lo_off + k * 32), idesc, 1u);
          umma_tf32(tmem_base, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, 1u);
        } else {
          umma_tf32(tmem_base, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, first);
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
  uint32_t cols = 32;
  while ((int)cols < n_pass) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage = (size_t)(kBM * 128 + n_pass * 128) * (x3 ? 2 : 1);
  const size_t tail = (size_t)K * kBM * 4 + 1024;
  // rows the level really holds (capacity-sized buffers in the CUDA-graph regime): two CTAs per SM when one wave of
  // 128-row tiles would not cover them otherwise
  const int64_t n_live = (n_expect > 0 && n_expect < n_cap) ? n_expect : n_cap;
  const int64_t tiles_live = cdiv(n_live, kBM) * (Cout / n_pass);
  const bool two_per_sm = tiles_live > sm_count() && 2 * stage + tail <= 110 * 1024;
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
