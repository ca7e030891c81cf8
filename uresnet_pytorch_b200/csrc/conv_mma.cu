// Gather-GEMM sparse convolution on the tensor pipe (warp-level mma.sync, TF32 operands,
// fp32 accumulate) for the narrow 16..64-channel layers of the U-ResNet -- the kernel the
// subm-conv HBM roofline is measured on (north-star job 2).
//
// Formulation (output-stationary, transposed):   D^T[co, site] = sum_k W_k^T[co, ci] . X_k^T[ci, site]
//   M = 16 output channels, N = 8 sites, K = 8 input channels per mma.sync.m16n8k8.
//   Putting the *sites* on the N = 8 side keeps the sparsity skip at 8-row granularity: an
//   (offset, 8-site tile) with no neighbour costs neither a gather nor an MMA.
// Data movement per warp tile (32 sites = 4 n-tiles):
//   - the 27 x 32 slice of the offset-major neighbour table: 27 coalesced 128 B loads;
//   - input rows: one 16 B cp.async per (row, 16 channels) straight into a lane-private slot
//     of a 3-stage per-warp ring (zero-fill for missing neighbours), pipelined over the
//     active offsets; every feature row is touched in full 32 B sectors;
//   - weights: all K offsets of one output-channel tile stay resident in shared memory in
//     mma-fragment order (one conflict-free LDS.128 per A fragment), loaded once per CTA;
//   - each output row is written exactly once, no atomics, offsets accumulated in order.
// The channel -> mma k-slot permutation is chosen so that a lane's B fragment is exactly the
// 16 bytes it gathered: no shuffles, no shared-memory transposes.
// MATH_TF32X3 splits both operands into tf32 hi/lo parts (3 MMAs) for fp32-equivalent
// accuracy; MATH_TF32 is a single pass.
#include "common.cuh"

namespace b200scn {
namespace {

constexpr int kNT = 4;       // n-tiles (8 sites each) per warp tile
constexpr int kStages = 3;   // ring depth (offsets in flight)
constexpr int kMaxSmem = 220 * 1024;

struct MmaParams {
  const float* in; int64_t ldi;
  const float* W; int64_t w_sk, w_si, w_so; int kflip;
  const int32_t* table; int64_t ld;
  const int32_t* parent; const uint8_t* kofs;
  int64_t n_cap; const int32_t* n_dev;
  int Cin, Cout; float* out; int64_t ldo; int accumulate; int x3;
  int warps;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// K = number of kernel offsets (compile time: the table-slice loads are fully unrolled);
// MT = 16-wide output-channel tiles per CTA pass.
template <int K, int MT>
__global__ void __launch_bounds__(MT >= 3 ? 256 : 512) k_conv_mma(const MmaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Cin = p.Cin;
  const int KS = Cin >> 3;   // k-steps of 8 channels
  const int Q = Cin >> 4;    // 16-channel (16 B per lane) chunks
  const int n = live_rows(p.n_dev, p.n_cap);
  const int warps = p.warps;
  if ((int64_t)blockIdx.x * warps * 32 >= n) return;  // nothing for this CTA: skip the weight load too

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int cob = blockIdx.y * (MT * 16);  // first output channel of this pass

  float* Wsm = reinterpret_cast<float*>(smem_raw);
  const int w_elems = K * KS * MT * 128;
  unsigned char* warp_base = smem_raw + (size_t)w_elems * 4;
  const size_t idx_bytes = (size_t)K * 32 * 4;
  const size_t mask_bytes = 128;  // K <= 32 masks
  const size_t ring_bytes = (size_t)kStages * kNT * Q * 32 * 16;
  const size_t per_warp = idx_bytes + mask_bytes + ring_bytes;
  int32_t* idx_s = reinterpret_cast<int32_t*>(warp_base + warp * per_warp);
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(warp_base + warp * per_warp + idx_bytes);
  float4* ring = reinterpret_cast<float4*>(warp_base + warp * per_warp + idx_bytes + mask_bytes);

  // ---- weights of this output-channel tile -> shared memory, mma A-fragment order ----------
  // Global reads run along whichever of (ci, co) is contiguous (forward: co, dgrad: ci) so
  // that a warp reads whole 128 B lines; the scatter into fragment order happens in smem.
  {
    const int CT = MT * 16;
    const bool co_fast = (p.w_so == 1);
    const int fast_n = co_fast ? CT : Cin, slow_n = co_fast ? Cin : CT;
    const int total = K * slow_n * fast_n;
    for (int e = tid; e < total; e += blockDim.x) {
      const int f = e % fast_n;
      int rest = e / fast_n;
      const int sl = rest % slow_n;
      const int k = rest / slow_n;
      const int ci = co_fast ? sl : f, col = co_fast ? f : sl;  // col = output channel within this pass
      const int kk = p.kflip ? (K - 1 - k) : k;
      const float v = __ldg(p.W + (int64_t)kk * p.w_sk + (int64_t)ci * p.w_si + (int64_t)(cob + col) * p.w_so);
      const int c16 = ci & 15, tt = c16 >> 2, rem = c16 & 3;
      const int s = ((ci >> 4) << 1) + (rem >> 1);
      const int mt = col >> 4, c = col & 15, gg = c & 7;
      const int r = ((rem & 1) << 1) | (c >> 3);
      Wsm[((((k * KS + s) * MT + mt) * 32) + gg * 4 + tt) * 4 + r] = v;
    }
  }
  __syncthreads();

  const int n_wt = (n + 31) >> 5;
  for (int wt = blockIdx.x * warps + warp; wt < n_wt; wt += gridDim.x * warps) {
    const int base = wt << 5;
    const int row = base + lane;
    __syncwarp();
    // ---- neighbour-table slice of this warp tile ---------------------------------------
    uint32_t any = 0;
    if (p.parent != nullptr) {
      int pk = -1, pp = -1;
      if (row < n) { pk = p.kofs[row]; pp = __ldg(p.parent + row); }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int e = (pk == k) ? pp : -1;
        idx_s[k * 32 + lane] = e;
        const uint32_t m = __ballot_sync(0xffffffffu, e >= 0);
        if (lane == 0) mask_s[k] = m;
        any |= (m != 0u) ? (1u << k) : 0u;
      }
    } else {
      int ev[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        int e = -1;
        if (row < n) e = (p.table != nullptr) ? __ldg(p.table + (int64_t)k * p.ld + row) : row;
        ev[k] = e;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        idx_s[k * 32 + lane] = ev[k];
        const uint32_t m = __ballot_sync(0xffffffffu, ev[k] >= 0);
        if (lane == 0) mask_s[k] = m;
        any |= (m != 0u) ? (1u << k) : 0u;
      }
    }
    __syncwarp();

    float acc[kNT][MT][4];
#pragma unroll
    for (int j = 0; j < kNT; ++j)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[j][mt][r] = 0.f;

    auto produce = [&](int k, int stage) {
      const uint32_t m = mask_s[k];
#pragma unroll
      for (int j = 0; j < kNT; ++j) {
        if ((m >> (8 * j)) & 0xFFu) {
          const int idx = idx_s[k * 32 + 8 * j + g];
          const float* src = p.in + (idx >= 0 ? (int64_t)idx * p.ldi + 4 * t : 0);
          const int sz = idx >= 0 ? 16 : 0;
          float4* dst = ring + ((size_t)(stage * kNT + j) * Q) * 32 + lane;
          for (int q = 0; q < Q; ++q) cp_async16(dst + q * 32, src + 16 * q, sz);
        }
      }
    };

    auto consume = [&](int k, int stage) {
      const uint32_t m = mask_s[k];
      const float4* wk = reinterpret_cast<const float4*>(Wsm) + (size_t)k * KS * MT * 32 + lane;
      for (int q = 0; q < Q; ++q) {
        float4 b4[kNT];
#pragma unroll
        for (int j = 0; j < kNT; ++j)
          if ((m >> (8 * j)) & 0xFFu) b4[j] = ring[((size_t)(stage * kNT + j) * Q + q) * 32 + lane];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = 2 * q + h;
          uint32_t a_hi[MT][4], a_lo[MT][4];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const float4 a4 = wk[(size_t)(s * MT + mt) * 32];
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              a_hi[mt][r] = to_tf32(av[r]);
              a_lo[mt][r] = p.x3 ? to_tf32(av[r] - __uint_as_float(a_hi[mt][r])) : 0u;
            }
          }
#pragma unroll
          for (int j = 0; j < kNT; ++j) {
            if ((m >> (8 * j)) & 0xFFu) {
              const float bx = h ? b4[j].z : b4[j].x;
              const float by = h ? b4[j].w : b4[j].y;
              const uint32_t b0 = to_tf32(bx), b1 = to_tf32(by);
              if (p.x3) {
                const uint32_t l0 = to_tf32(bx - __uint_as_float(b0)), l1 = to_tf32(by - __uint_as_float(b1));
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  mma_tf32(acc[j][mt], a_lo[mt], b0, b1);
                  mma_tf32(acc[j][mt], a_hi[mt], l0, l1);
                }
              }
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[j][mt], a_hi[mt], b0, b1);
            }
          }
        }
      }
    };

    // ---- software pipeline over the active offsets, in increasing k -------------------------
    uint32_t pend = any, cons = any;
    int pstage = 0, cstage = 0;
#pragma unroll
    for (int i = 0; i < kStages - 1; ++i) {
      if (pend) {
        const int k = __ffs(pend) - 1;
        pend &= pend - 1;
        produce(k, pstage);
        pstage = (pstage + 1 == kStages) ? 0 : pstage + 1;
      }
      cp_async_commit();
    }
    while (cons) {
      const int k = __ffs(cons) - 1;
      cons &= cons - 1;
      if (pend) {
        const int kn = __ffs(pend) - 1;
        pend &= pend - 1;
        produce(kn, pstage);
        pstage = (pstage + 1 == kStages) ? 0 : pstage + 1;
      }
      cp_async_commit();
      cp_async_wait<kStages - 1>();
      consume(k, cstage);
      cstage = (cstage + 1 == kStages) ? 0 : cstage + 1;
    }
    cp_async_wait<0>();

    // ---- epilogue: D^T fragments -> out[site][co] (full 32 B sectors per store) -----------------
#pragma unroll
    for (int j = 0; j < kNT; ++j) {
      const int s0 = base + 8 * j + 2 * t;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int co = cob + mt * 16 + g;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int site = s0 + (r & 1);
          if (site < n) {
            float* o = p.out + (int64_t)site * p.ldo + co + 8 * (r >> 1);
            *o = p.accumulate ? (*o + acc[j][mt][r]) : acc[j][mt][r];
          }
        }
      }
    }
  }
}

struct Plan {
  int MT, warps, passes;
  size_t smem;
  bool ok;
};

Plan make_plan_chunk(int K, int Cin, int Cout);

// Largest input-channel chunk (multiple of 16) whose weights + ring fit; the layer is then
// computed as a sum over chunks (later chunks accumulate into `out`).
int pick_cin_chunk(int K, int Cin, int Cout) {
  for (int c = Cin; c >= 16; c -= 16)
    if (make_plan_chunk(K, c, Cout).ok) return c;
  return 0;
}

Plan make_plan_chunk(int K, int Cin, int Cout) {
  Plan best{0, 0, 0, 0, false};
  const size_t per_warp = (size_t)K * 32 * 4 + 128 + (size_t)kStages * kNT * (Cin / 16) * 32 * 16;
  for (int mt = 4; mt >= 1; --mt) {
    if ((Cout / 16) % mt != 0) continue;
    const size_t wbytes = (size_t)K * Cin * 16 * mt * 4;
    if (wbytes + 4 * per_warp > (size_t)kMaxSmem) continue;
    int warps = (int)(((size_t)kMaxSmem - wbytes) / per_warp);
    const int wmax = mt >= 3 ? 8 : 16;
    if (warps > wmax) warps = wmax;
    // two CTAs per SM when the whole thing is small
    if (wbytes + warps * per_warp > (size_t)kMaxSmem / 2 && wbytes + 8 * per_warp <= (size_t)kMaxSmem / 2) warps = 8;
    best = Plan{mt, warps, (Cout / 16) / mt, wbytes + warps * per_warp, true};
    break;
  }
  return best;
}

template <int K, int MT>
int launch(const MmaParams& p, const Plan& plan, cudaStream_t st) {
  static thread_local size_t configured = 0;
  if (plan.smem > configured) {
    B200_CUDA(cudaFuncSetAttribute(k_conv_mma<K, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    configured = kMaxSmem;
  }
  const int64_t n_wt = cdiv(p.n_cap, 32);
  int64_t gx = cdiv(n_wt, plan.warps);
  const int ctas_per_sm = plan.smem <= (size_t)kMaxSmem / 2 ? 2 : 1;
  const int64_t gmax = (int64_t)sm_count() * ctas_per_sm;
  if (gx > gmax) gx = gmax;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)plan.passes);
  k_conv_mma<K, MT><<<grid, plan.warps * 32, plan.smem, st>>>(p);
  count_launches(1);
  B200_CHECK_LAUNCH("conv_mma");
  return 0;
}

template <int K>
int launch_k(const MmaParams& p, const Plan& plan, cudaStream_t st) {
  switch (plan.MT) {
    case 1: return launch<K, 1>(p, plan, st);
    case 2: return launch<K, 2>(p, plan, st);
    case 3: return launch<K, 3>(p, plan, st);
    default: return launch<K, 4>(p, plan, st);
  }
}

}  // namespace

bool conv_mma_supported(int Cin, int Cout, int K, int64_t ldi, const void* in) {
  if (Cin % 16 != 0 || Cout % 16 != 0) return false;
  if (!(K == 1 || K == 4 || K == 8 || K == 9 || K == 27)) return false;
  if (ldi % 4 != 0 || (reinterpret_cast<uintptr_t>(in) & 15) != 0) return false;
  return pick_cin_chunk(K, Cin, Cout) > 0;
}

int conv_rows_mma(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so, int kflip,
                  const int32_t* table, int64_t ld, const int32_t* parent, const uint8_t* kofs, int K, int64_t n_cap,
                  const int32_t* n_dev, int Cin, int Cout, float* out, int64_t ldo, int accumulate, int math_mode,
                  cudaStream_t st) {
  if (n_cap == 0) return 0;
  const int chunk = pick_cin_chunk(K, Cin, Cout);
  if (chunk == 0) {
    set_error("conv_mma: shape not supported (Cin=%d Cout=%d K=%d)", Cin, Cout, K);
    return B200SCN_ERR_UNSUPPORTED;
  }
  for (int c0 = 0; c0 < Cin; c0 += chunk) {
    const int cc = Cin - c0 < chunk ? Cin - c0 : chunk;
    Plan plan = make_plan_chunk(K, cc, Cout);
    MmaParams p;
    p.in = in + c0; p.ldi = ldi; p.W = W + (int64_t)c0 * w_si; p.w_sk = w_sk; p.w_si = w_si; p.w_so = w_so;
    p.kflip = kflip; p.table = table; p.ld = ld; p.parent = parent; p.kofs = kofs; p.n_cap = n_cap; p.n_dev = n_dev;
    p.Cin = cc; p.Cout = Cout; p.out = out; p.ldo = ldo; p.accumulate = (c0 > 0) ? 1 : accumulate;
    p.x3 = (math_mode == B200SCN_MATH_TF32X3) ? 1 : 0;
    p.warps = plan.warps;
    int rc;
    switch (K) {
      case 1: rc = launch_k<1>(p, plan, st); break;
      case 4: rc = launch_k<4>(p, plan, st); break;
      case 8: rc = launch_k<8>(p, plan, st); break;
      case 9: rc = launch_k<9>(p, plan, st); break;
      default: rc = launch_k<27>(p, plan, st); break;
    }
    if (rc) return rc;
  }
  return 0;
}

}  // namespace b200scn
