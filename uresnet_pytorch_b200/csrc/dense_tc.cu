// Dense-path convolution on the 5th-generation tensor cores (north-star job 4): the Conv{2,3}d /
// ConvTranspose{2,3}d family of /root/reference/uresnet/models/uresnet_dense.py:12-226 as an implicit GEMM
//     out[o][co] = bias[co] + sum_{tap, ci} in[idx(o, tap)][ci] * W[tap][ci][co]
// with M = 128 consecutive output voxels per CTA, N = Cout (<= 256 per pass), K = taps x Cin walked in
// blocks of 32 input channels (one 128-byte swizzle row of tf32), issued as tcgen05.mma kind::tf32 by one
// thread, accumulators in TMEM, operands staged in shared memory in the canonical K-major SWIZZLE_128B
// layout:
//     A stage [128 rows][128 B]: row r = the 32 channels of in[idx(o0 + r, tap)] (zero-filled when the tap
//                                falls outside the grid / on a stride hole), gathered with 16-byte cp.async
//     B stage [N rows][128 B]  : row n = Wt[tap][chunk][n][0..31], a K-major (transposed) copy of the weights
//                                made by k_dense_pack_wt once per call
// Pipeline: 4 shared-memory stages; 128 producer threads (one A row each + a share of B) signal a stage's
// `full` mbarrier after cp.async.wait_group + fence.proxy.async; the MMA thread waits on it, issues 4 MMAs
// (K = 8 each), and tcgen05.commit releases the stage (`empty`) -- or, after the last block, hands the
// accumulator to the epilogue: tcgen05.ld 32x32b, + bias, 16-byte stores.
// Math modes: B200SCN_MATH_TF32 = single pass (10-bit mantissa operands, fp32 accumulate); B200SCN_MATH_TF32X3 =
// 3-term split (fp32-equivalent, see k_dense_conv_tc<true>); B200SCN_MATH_FP32 keeps the exact FFMA path
// (csrc/dense.cu), which also serves channel counts this kernel does not take.
#include "common.cuh"
#include "tc_common.cuh"

namespace b200scn {
namespace {

constexpr int kBM = 128;     // output voxels per CTA (UMMA M)
constexpr int kBK = 32;      // input channels per K block (128 B of tf32)
constexpr int kProducers = 128;
constexpr int kLag = 2;      // producers run this many stages ahead of their own wait_group

using namespace tc;

// K blocks: the contraction axis is the flattened (tap, ci) index, walked 32 at a time.  Cin % 32 == 0: block kb =
// (tap, 32-channel chunk).  Cin == 16: block kb = the 16 channels of taps 2 kb and 2 kb + 1 (the odd last tap is
// paired with zeros), so the 16-channel level runs on the same 128-byte swizzle rows.
__host__ __device__ __forceinline__ int dtc_blocks(int taps, int Cin) {
  return Cin == 16 ? (taps + 1) / 2 : taps * (Cin / kBK);
}

// Wt[kb][co][32] = W[tap(kb, j)][ci(kb, j)][co]   (K-major copy: 128 contiguous bytes per (co, block))
__global__ void k_dense_pack_wt(const float* __restrict__ W, int64_t w_sk, int64_t w_si, int64_t w_so, int taps, int Cin,
                                int Cout, float* __restrict__ Wt, float* __restrict__ Wt_lo) {
  const int64_t total = (int64_t)dtc_blocks(taps, Cin) * kBK * Cout;
  const int chunks = Cin / kBK;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % kBK);
    int64_t rest = e / kBK;
    const int co = (int)(rest % Cout);
    const int kb = (int)(rest / Cout);
    int tap, ci;
    if (Cin == 16) { tap = 2 * kb + (j >> 4); ci = j & 15; }
    else { tap = kb / chunks; ci = (kb - tap * chunks) * kBK + j; }
    const float v = tap < taps ? __ldg(W + tap * w_sk + (int64_t)ci * w_si + co * w_so) : 0.f;
    if (Wt_lo != nullptr) {  // 3xTF32: high part = truncation to tf32, low part = exact remainder
      const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      Wt[e] = hi;
      Wt_lo[e] = v - hi;
    } else {
      Wt[e] = v;
    }
  }
}

struct DtcParams {
  const float* in; int64_t ldi;
  const float* Wt;      // packed weights, [K blocks][Cout][32]
  const float* Wt_lo;   // low parts (3xTF32) or NULL
  const float* bias;
  float* out; int64_t ldo;
  int64_t n_out;
  int Cin, Cout, taps;
  int n_pass;           // output channels of one CTA (blockIdx.y selects the pass), multiple of 16, <= 256
  uint32_t tmem_cols;
  DenseGeom geom;
};

// X3: 3xTF32 (fp32-equivalent products): D += A_lo B_hi + A_hi B_lo + A_hi B_hi.  The A rows then pass through
// registers (LDG.128 -> split -> STS.128 into a high and a low tile, next block's row prefetched), the low
// weight parts come from a second packed copy; 3 shared-memory stages of [A_hi | A_lo | B_hi | B_lo].
template <bool X3>
__global__ void __launch_bounds__(kProducers + 32, 1) k_dense_conv_tc(const DtcParams p) {
  constexpr int kStages = X3 ? 3 : 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_base_s;
  // 1024-byte alignment of the operand tiles (SWIZZLE_128B atoms are 8 rows x 128 B)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.n_pass;
  const uint32_t a_bytes = kBM * 128, b_bytes = (uint32_t)N * 128;
  const uint32_t stage_bytes = (a_bytes + b_bytes) * (X3 ? 2 : 1);
  const uint32_t a_lo_off = a_bytes, b_off = a_bytes * (X3 ? 2 : 1), b_lo_off = b_off + b_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * N;
  const int chunks = p.Cin / kBK;  // (0 in the two-taps-per-block mode of Cin == 16)
  const bool half = p.Cin == 16;
  const int n_kb = dtc_blocks(p.taps, p.Cin);
  const int64_t o0 = (int64_t)blockIdx.x * kBM;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], kProducers); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {  // TMEM allocation by one warp, which also owns the deallocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ===== producers: thread r owns A row r (output voxel o0 + r) and B rows r, r + 128 =====
    const int r = tid;
    const int64_t o = o0 + r;
    const uint32_t sw = (uint32_t)(r & 7);
    int cur_tap = -1, idx = -1;
    // the two 16-channel halves of this voxel's row of block kb (nullptr: zeros)
    auto rows_of = [&](int kb, const float*& ra, const float*& rb) {
      if (half) {
        const int t0 = 2 * kb, t1 = t0 + 1;
        const int i0 = o < p.n_out ? dense_index(p.geom, (int)o, t0) : -1;
        const int i1 = (o < p.n_out && t1 < p.taps) ? dense_index(p.geom, (int)o, t1) : -1;
        ra = i0 >= 0 ? p.in + (int64_t)i0 * p.ldi : nullptr;
        rb = i1 >= 0 ? p.in + (int64_t)i1 * p.ldi : nullptr;
        return;
      }
      const int tap = kb / chunks, chunk = kb - tap * chunks;
      if (tap != cur_tap) {
        cur_tap = tap;
        idx = o < p.n_out ? dense_index(p.geom, (int)o, tap) : -1;
      }
      ra = idx >= 0 ? p.in + (int64_t)idx * p.ldi + chunk * kBK : nullptr;
      rb = ra ? ra + 16 : nullptr;
    };
    auto copy_b = [&](int kb, int s) {
      const int64_t woff = ((int64_t)kb * p.Cout + co0) * kBK;
      for (int n = r; n < N; n += kProducers) {
        const uint32_t b_dst = smem_u32(smem + (size_t)s * stage_bytes + b_off) + (uint32_t)n * 128;
        const uint32_t swn = (uint32_t)(n & 7);
#pragma unroll
        for (int j = 0; j < 8; ++j) cp_async16_zfill(b_dst + ((j ^ swn) << 4), p.Wt + woff + (int64_t)n * kBK + 4 * j, 16);
        if (X3) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            cp_async16_zfill(b_dst + b_bytes + ((j ^ swn) << 4), p.Wt_lo + woff + (int64_t)n * kBK + 4 * j, 16);
        }
      }
    };
    if (!X3) {
      for (int kb = 0; kb < n_kb + kLag; ++kb) {
        if (kb < n_kb) {
          const int s = kb % kStages;
          mbar_wait(&empty_bar[s], ((kb / kStages) & 1) ^ 1);
          const float *ra, *rb;
          rows_of(kb, ra, rb);
          const uint32_t a_dst = smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) cp_async16_zfill(a_dst + ((j ^ sw) << 4), (ra ? ra : p.in) + 4 * j, ra ? 16 : 0);
#pragma unroll
          for (int j = 4; j < 8; ++j) cp_async16_zfill(a_dst + ((j ^ sw) << 4), (rb ? rb : p.in) + 4 * (j - 4), rb ? 16 : 0);
          copy_b(kb, s);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int done_kb = kb - kLag;  // the copies of this block have had kLag iterations to land
        if (done_kb >= 0) {
          asm volatile("cp.async.wait_group %0;" ::"n"(kLag) : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
          mbar_arrive(&full_bar[done_kb % kStages]);
        }
      }
    } else {
      float4 cur[8], nxt[8];
      auto load_row = [&](int kb, float4 (&v)[8]) {
        const float *ra, *rb;
        rows_of(kb, ra, rb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = ra ? __ldg(reinterpret_cast<const float4*>(ra) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 + j] = rb ? __ldg(reinterpret_cast<const float4*>(rb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_row(0, cur);
      for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % kStages;
        if (kb + 1 < n_kb) load_row(kb + 1, nxt);  // in flight while this block is split and stored
        mbar_wait(&empty_bar[s], ((kb / kStages) & 1) ^ 1);
        copy_b(kb, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const uint32_t a_dst = smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
          float hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            hi[q] = __uint_as_float(__float_as_uint(x[q]) & 0xffffe000u);
            lo[q] = x[q] - hi[q];
          }
          const uint32_t d = a_dst + ((j ^ sw) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d + a_lo_off), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]) : "memory");
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&full_bar[s]);
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
      }
    }
    // ===== epilogue: TMEM -> registers -> global; warp w reads TMEM lanes 32w .. 32w+31 =====
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* orow = p.out + o * p.ldo + co0;
    const bool vec = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (co0 % 4 == 0);
    for (int c = 0; c < N; c += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(t_row + (uint32_t)c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (o < p.n_out) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + co0 + c + j) : 0.f);
        if (vec) {
          *reinterpret_cast<float4*>(orow + c) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(orow + c + 4) = make_float4(f[4], f[5], f[6], f[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) orow[c + j] = f[j];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (lane == 0) {
    // ===== MMA issuer: one thread =====
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9,
    // 10-12 = 2), both K-major (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kStages;
      mbar_wait(&full_bar[s], (kb / kStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes), b_addr = a_addr + b_off;
#pragma unroll
      for (int k = 0; k < kBK / 8; ++k) {  // UMMA K = 8 tf32 = 32 bytes along the swizzled row
        const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
        if (X3) {
          umma_tf32(tmem_base, umma_desc(a_addr + a_lo_off + k * 32), umma_desc(b_addr + k * 32), idesc, first);
          umma_tf32(tmem_base, umma_desc(a_addr + k * 32), umma_desc(a_addr + b_lo_off + k * 32), idesc, 1u);
          umma_tf32(tmem_base, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, 1u);
        } else {
          umma_tf32(tmem_base, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, first);
        }
      }
      umma_commit(&empty_bar[s]);  // frees the stage once the MMAs above have read it
    }
    umma_commit(&done_bar);        // accumulator complete
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace

int read_geom_public(const int32_t* g, DenseGeom* out);

}  // namespace b200scn

using namespace b200scn;

extern "C" int64_t b200scn_dense_conv_tc_scratch_bytes(int taps, int Cin, int Cout) {
  return (int64_t)2 * dtc_blocks(taps, Cin) * kBK * Cout * 4 + 512;
}

/* tcgen05 implicit-GEMM variant of b200scn_dense_conv.  Needs Cin % 32 == 0 or Cin == 16, and Cout % 16 == 0;
 * returns B200SCN_ERR_UNSUPPORTED otherwise (the caller falls back to b200scn_dense_conv). */
extern "C" int b200scn_dense_conv_tc(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si,
                                     int64_t w_so, const float* bias, const int32_t* geom_host, int batch, int Cin,
                                     int Cout, float* out, int64_t ldo, int math_mode, void* scratch,
                                     int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DenseGeom g;
  if (int rc = read_geom_public(geom_host, &g)) return rc;
  if ((Cin % 32 != 0 && Cin != 16) || Cout % 16 != 0 || Cin < 16 || Cout < 16) {
    set_error("dense_conv_tc: needs Cin %% 32 == 0 or Cin == 16, and Cout %% 16 == 0 (got %d, %d)", Cin, Cout);
    return B200SCN_ERR_UNSUPPORTED;
  }
  B200_CHECK_ARG(ldi % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "dense_conv_tc: input rows must be 16-byte aligned");
  const int taps = g.dims == 3 ? g.ks * g.ks * g.ks : g.ks * g.ks;
  B200_CHECK_ARG(scratch_bytes >= b200scn_dense_conv_tc_scratch_bytes(taps, Cin, Cout), "dense_conv_tc: scratch too small");
  const int64_t n = (int64_t)batch * g.out_sz[0] * g.out_sz[1] * g.out_sz[2];
  B200_CHECK_ARG(n < (1ll << 31) && n >= 1, "dense_conv_tc: bad voxel count");
  B200_CHECK_ARG(math_mode == B200SCN_MATH_TF32 || math_mode == B200SCN_MATH_TF32X3, "dense_conv_tc: tensor-core math modes only");
  const bool x3 = math_mode == B200SCN_MATH_TF32X3;
  const int64_t total = (int64_t)dtc_blocks(taps, Cin) * kBK * Cout;
  float* Wt = reinterpret_cast<float*>(scratch);
  float* Wt_lo = x3 ? Wt + ((total + 63) / 64) * 64 : nullptr;
  {
    int64_t gsz = cdiv(total, 256);
    if (gsz > 2048) gsz = 2048;
    k_dense_pack_wt<<<(int)gsz, 256, 0, st>>>(W, w_sk, w_si, w_so, taps, Cin, Cout, Wt, Wt_lo);
  }
  // passes over the output channels: the largest multiple of 16 <= 256 (128 with the doubled 3xTF32 tiles)
  // that divides Cout
  const int n_max = x3 ? 128 : 256;
  int n_pass = Cout;
  if (n_pass > n_max) {
    n_pass = n_max;
    while (Cout % n_pass != 0) n_pass -= 16;
  }
  DtcParams p;
  p.in = in; p.ldi = ldi; p.Wt = Wt; p.Wt_lo = Wt_lo; p.bias = bias; p.out = out; p.ldo = ldo; p.n_out = n; p.Cin = Cin; p.Cout = Cout;
  p.taps = taps; p.n_pass = n_pass; p.geom = g;
  uint32_t cols = 32;
  while ((int)cols < n_pass) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = (size_t)(x3 ? 3 * 2 : 4) * (kBM * 128 + (size_t)n_pass * 128) + 1024;
  static thread_local bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_dense_conv_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    B200_CUDA(cudaFuncSetAttribute(k_dense_conv_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)cdiv(n, kBM), (unsigned)(Cout / n_pass));
  if (x3)
    k_dense_conv_tc<true><<<grid, kProducers + 32, smem, st>>>(p);
  else
    k_dense_conv_tc<false><<<grid, kProducers + 32, smem, st>>>(p);
  count_launches(2);
  B200_CHECK_LAUNCH("dense_conv_tc");
  return 0;
}
