// Weight gradients of the gather-GEMM convolutions on the tensor pipe (mma.sync, TF32
// operands, fp32 accumulate):   dW[k][ci][co] = sum over rule pairs (a, b) of offset k of
// A[a][ci] * B[b][co]   (A = layer input rows, B = output-gradient rows).
//
// One CTA owns (offset k, row range).  It (1) streams its slice of the offset-major rule
// table with coalesced loads and compacts the valid (a, b) pairs in row order (ballot +
// prefix, deterministic), (2) gathers the pairs' A and B rows 64 at a time into padded shared
// memory with 16 B cp.async (double buffered), (3) contracts over the pair axis with
// m16n8k8: M = input channels, N = output channels, K = pairs, every warp keeping its
// (ci, co) tiles in registers across the whole row range, and (4) writes one partial
// [Ca x Cb] block to scratch.  A second tiny kernel folds the partials in fixed order, so
// the result is bit-reproducible (no atomics).  Sparse offsets cost what their pair count
// costs: an offset with 12% valid rows does 12% of the gathers and MMAs.
#include <stdlib.h>

#include "common.cuh"

namespace b200scn {
namespace {

constexpr int kBatch = 64;  // pairs gathered per stage
constexpr int kWarpsMax = 8;

// per-(offset, column split) tickets: the last row-chunk CTA of an offset folds that offset's
// partial blocks in chunk order (deterministic) -- no separate fold launch.  Zero at module load,
// reset by the folding CTA.  wgrad launches are stream-ordered among themselves.
__device__ unsigned int g_wgrad_ticket[64 * 8];

struct WgParams {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  const int32_t* table; int64_t ld;
  int K, gather_b;
  int64_t n_cap; const int32_t* n_dev;
  int Ca, Cb, cb_per_z;
  int rows_per_cta;
  float* part;
  int x3;
  float* dW; int accumulate; int fused_fold;
  int use_geom; DenseGeom geom;  // dense path: the rule of (row, offset) is computed, not looked up
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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// rows per partial block when n live rows are split over `parts` CTAs (multiple of 32), and how many of
// those CTAs own at least one row
__host__ __device__ __forceinline__ int wg_rows_per_part(int n, int parts) {
  const int per = (n + parts - 1) / parts;
  return ((per + 31) / 32) * 32;
}
__host__ __device__ __forceinline__ int wg_live_parts(int n, int parts) {
  const int per = wg_rows_per_part(n, parts);
  const int live = per > 0 ? (n + per - 1) / per : 0;
  return live < 1 ? 1 : (live > parts ? parts : live);
}

// TPW = (ci, co) mma tiles per warp.  grid = (K, row parts, column splits)
template <int TPW>
__global__ void __launch_bounds__(kWarpsMax * 32) k_wgrad_mma(const WgParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = live_rows(p.n_dev, p.n_cap);
  const int k = blockIdx.x;
  // The live rows are split evenly over the gridDim.y CTAs of this offset (contiguous ranges, a multiple
  // of 32 rows); a CTA walks its range in sub-chunks of rows_per_cta rows and keeps its accumulators in
  // registers across them, so the grid -- and the number of partial blocks to fold -- does not grow with
  // the row CAPACITY of the buffers (profiles/README.md: capacity-sized grids cost ~1 ms per step).
  const int per_cta = wg_rows_per_part(n, (int)gridDim.y);
  const int live_chunks = wg_live_parts(n, (int)gridDim.y);
  if ((int)blockIdx.y >= live_chunks) return;
  const int cta_begin = blockIdx.y * per_cta, cta_end = min(n, cta_begin + per_cta);
  const int cb0 = blockIdx.z * p.cb_per_z;                 // first output channel of this split
  const int Cbz = min(p.cb_per_z, p.Cb - cb0);             // output channels handled here
  const int Ca = p.Ca;
  const int sa = Ca + 8, sb = Cbz + 8;                     // padded row strides (conflict-free frags)
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int g = lane >> 2, t = lane & 3;

  // smem: pair lists [rows_per_cta] x 2 | warp counters | 2 stages x (A[64][sa] + B[64][sb])
  int32_t* pa = reinterpret_cast<int32_t*>(smem_raw);
  int32_t* pb = pa + p.rows_per_cta;
  int32_t* wsum = pb + p.rows_per_cta;  // [kWarpsMax + 1]
  float* stage_base = reinterpret_cast<float*>(wsum + 16);
  const int stage_floats = kBatch * (sa + sb);

  const int MT = Ca >> 4, NT8 = Cbz >> 3;
  const int T = MT * NT8;
  float acc[TPW][4];
#pragma unroll
  for (int i = 0; i < TPW; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;
  const int qa = Ca >> 2, qb = Cbz >> 2;  // 16 B chunks per row

  for (int r_begin = cta_begin; r_begin < cta_end; r_begin += p.rows_per_cta) {
  const int r_end = min(cta_end, r_begin + p.rows_per_cta);
  // ---- (1) compact the valid pairs of [r_begin, r_end) in row order ---------------------------
  int total = 0;
  for (int c0 = r_begin; c0 < r_end; c0 += nthr) {
    const int r = c0 + tid;
    int e = -1;
    if (r < r_end) e = p.use_geom ? dense_index(p.geom, r, k) : (p.table ? __ldg(p.table + (int64_t)k * p.ld + r) : r);
    const uint32_t m = __ballot_sync(0xffffffffu, e >= 0);
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    int before = total;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    int all = 0;
    for (int w = 0; w < nwarp; ++w) all += wsum[w];
    if (e >= 0) {
      const int pos = before + __popc(m & ((1u << lane) - 1u));
      pa[pos] = p.gather_b ? r : e;
      pb[pos] = p.gather_b ? e : r;
    }
    total += all;
    __syncthreads();
  }

  auto issue = [&](int batch, int stage) {
    float* As = stage_base + stage * stage_floats;
    float* Bs = As + kBatch * sa;
    const int p0 = batch * kBatch;
    const int cnt = min(kBatch, total - p0);
    for (int e = tid; e < cnt * qa; e += nthr) {
      const int pr = e / qa, c = e - pr * qa;
      cp_async16(As + pr * sa + 4 * c, p.A + (int64_t)pa[p0 + pr] * p.lda + 4 * c);
    }
    for (int e = tid; e < cnt * qb; e += nthr) {
      const int pr = e / qb, c = e - pr * qb;
      cp_async16(Bs + pr * sb + 4 * c, p.B + (int64_t)pb[p0 + pr] * p.ldb + cb0 + 4 * c);
    }
    // zero the tail rows of the last k-step so stale data never enters the MMA
    const int cnt8 = (cnt + 7) & ~7;
    for (int e = tid; e < (cnt8 - cnt) * sa; e += nthr) As[cnt * sa + e] = 0.f;
    for (int e = tid; e < (cnt8 - cnt) * sb; e += nthr) Bs[cnt * sb + e] = 0.f;
  };

  const int nbatch = (total + kBatch - 1) / kBatch;
  if (nbatch > 0) issue(0, 0);
  cp_async_commit();
  for (int b = 0; b < nbatch; ++b) {
    if (b + 1 < nbatch) issue(b + 1, (b + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* As = stage_base + (b & 1) * stage_floats;
    const float* Bs = As + kBatch * sa;
    const int cnt = min(kBatch, total - b * kBatch);
    const int ksteps = (cnt + 7) >> 3;
    for (int ks = 0; ks < ksteps; ++ks) {
      const float* a_row0 = As + (ks * 8 + t) * sa;
      const float* a_row1 = a_row0 + 4 * sa;
      const float* b_row0 = Bs + (ks * 8 + t) * sb;
      const float* b_row1 = b_row0 + 4 * sb;
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        const int tile = warp + i * nwarp;
        if (tile < T) {
          const int mt = tile / NT8, nt = tile - mt * NT8;
          const float av[4] = {a_row0[mt * 16 + g], a_row0[mt * 16 + g + 8], a_row1[mt * 16 + g],
                               a_row1[mt * 16 + g + 8]};
          const float bx = b_row0[nt * 8 + g], by = b_row1[nt * 8 + g];
          // hi = truncation to tf32 (one LOP3; cvt.rna expands to ~4 ALU ops), lo = exact remainder
          uint32_t ah[4], al[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) ah[r] = __float_as_uint(av[r]) & 0xffffe000u;
          const uint32_t b0 = __float_as_uint(bx) & 0xffffe000u, b1 = __float_as_uint(by) & 0xffffe000u;
          if (p.x3) {
#pragma unroll
            for (int r = 0; r < 4; ++r) al[r] = __float_as_uint(av[r] - __uint_as_float(ah[r]));
            const uint32_t l0 = __float_as_uint(bx - __uint_as_float(b0));
            const uint32_t l1 = __float_as_uint(by - __uint_as_float(b1));
            mma_tf32(acc[i], al, b0, b1);
            mma_tf32(acc[i], ah, l0, l1);
          }
          mma_tf32(acc[i], ah, b0, b1);
        }
      }
    }
    __syncthreads();  // stage (b & 1) is refilled by the next iteration's issue
  }
  cp_async_wait<0>();
  __syncthreads();  // the pair lists are rewritten by the next sub-chunk
  }

  // ---- (4) partial block -> scratch[(chunk * K + k)][Ca][Cb] ------------------------------------
  float* dst = p.part + ((int64_t)blockIdx.y * p.K + k) * Ca * p.Cb;
#pragma unroll
  for (int i = 0; i < TPW; ++i) {
    const int tile = warp + i * nwarp;
    if (tile < T) {
      const int mt = tile / NT8, nt = tile - mt * NT8;
      const int ci = mt * 16 + g, co = cb0 + nt * 8 + 2 * t;
      *reinterpret_cast<float2*>(dst + (int64_t)ci * p.Cb + co) = make_float2(acc[i][0], acc[i][1]);
      *reinterpret_cast<float2*>(dst + (int64_t)(ci + 8) * p.Cb + co) = make_float2(acc[i][2], acc[i][3]);
    }
  }

  if (!p.fused_fold) return;
  // ---- (5) optional: the last chunk-CTA of this (offset, column split) folds the partial blocks ----
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  unsigned int* ticket = &g_wgrad_ticket[(k & 63) * 8 + (blockIdx.z & 7)];
  if (tid == 0) is_last = (atomicAdd(ticket, 1u) == (unsigned)live_chunks - 1u);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int64_t chunk_stride = (int64_t)p.K * Ca * p.Cb;
  const float* src = p.part + (int64_t)k * Ca * p.Cb;
  float* out = p.dW + (int64_t)k * Ca * p.Cb;
  for (int e = tid; e < Ca * Cbz; e += nthr) {
    const int ci = e / Cbz, co = cb0 + (e - ci * Cbz);
    const int64_t off = (int64_t)ci * p.Cb + co;
    float s = 0.f;
    int c = 0;
    for (; c + 8 <= live_chunks; c += 8) {  // 8 independent loads in flight, summed in chunk order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (c + u) * chunk_stride + off);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; c < live_chunks; ++c) s += __ldcg(src + c * chunk_stride + off);
    out[off] = p.accumulate ? out[off] + s : s;
  }
  if (tid == 0) *ticket = 0u;
}

__global__ void k_wgrad_fold(const float* __restrict__ part, int64_t elems, int rows_per_cta, int nchunks,
                             int64_t n_cap, const int32_t* __restrict__ n_dev, float* __restrict__ dW,
                             int accumulate) {
  const int n = live_rows(n_dev, n_cap);
  const int live = n > 0 ? wg_live_parts(n, nchunks) : 0;
  (void)rows_per_cta;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < elems; e += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < live; ++c) s += part[(int64_t)c * elems + e];
    dW[e] = accumulate ? dW[e] + s : s;
  }
}

struct WgPlan {
  int rows_per_cta, nchunks, zsplits, cb_per_z, warps, tpw;
  size_t smem;
  bool ok;
};

WgPlan wg_plan(int64_t n_cap, int K, int Ca, int Cb) {
  WgPlan pl{};
  pl.ok = false;
  if (Ca % 16 != 0 || Cb % 16 != 0 || Ca > 512 || Cb > 512) return pl;
  // column splits so that a warp holds at most 16 tiles (64 accumulator registers)
  int z = 1;
  while (true) {
    const int cbz = ((Cb / 8 + z - 1) / z) * 8;
    const int T = (Ca / 16) * (cbz / 8);
    if (T <= kWarpsMax * 16 || cbz <= 8) break;
    ++z;
  }
  int cbz = ((Cb / 16 + z - 1) / z) * 16;  // keep 16-channel granularity (cp.async alignment)
  z = (Cb + cbz - 1) / cbz;
  const int T = (Ca / 16) * (cbz / 8);
  int warps = T < kWarpsMax ? T : kWarpsMax;
  if (warps < 4) warps = 4;  // extra warps still help the table scan, compaction, gathers and the fold
  int tpw = (T + warps - 1) / warps;
  int tpw_pow = 1;
  while (tpw_pow < tpw) tpw_pow <<= 1;
  if (tpw_pow > 16) return pl;
  // sub-chunk of rows compacted at a time (pair lists in shared memory), and the number of row parts
  // (= CTAs per offset and column split = partial blocks): enough to cover the SMs about twice, never
  // more than one part per 256 rows of capacity
  static int parts_target = 0;  // tunable for A/B runs: B200SCN_WG_PARTS (CTAs aimed at, default 4 x SMs)
  if (parts_target == 0) {
    const char* e = getenv("B200SCN_WG_PARTS");
    parts_target = e ? atoi(e) : 4 * sm_count();
    if (parts_target < 1) parts_target = 1;
  }
  const int64_t rows = 1024;
  int64_t parts = cdiv(parts_target, (int64_t)K * z);
  const int64_t max_parts = cdiv(n_cap < 1 ? 1 : n_cap, 256);
  if (parts > max_parts) parts = max_parts;
  if (parts > 128) parts = 128;
  if (parts < 1) parts = 1;
  pl.rows_per_cta = (int)rows;
  pl.nchunks = (int)parts;
  pl.zsplits = z;
  pl.cb_per_z = cbz;
  pl.warps = warps;
  pl.tpw = tpw_pow;
  pl.smem = (size_t)rows * 8 + 64 + (size_t)2 * kBatch * (Ca + 8 + cbz + 8) * 4;
  pl.ok = pl.smem <= 220 * 1024;
  return pl;
}

template <int TPW>
int wg_launch(const WgParams& p, const WgPlan& pl, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(k_wgrad_mma<TPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  dim3 grid(p.K, pl.nchunks, pl.zsplits);
  k_wgrad_mma<TPW><<<grid, pl.warps * 32, pl.smem, st>>>(p);
  return 0;
}

}  // namespace

bool wgrad_mma_supported(int64_t n_cap, int K, int Ca, int Cb, int64_t lda, int64_t ldb, const void* A,
                         const void* B) {
  if (lda % 4 != 0 || ldb % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(B) & 15) != 0) return false;
  return wg_plan(n_cap, K, Ca, Cb).ok;
}

int64_t wgrad_scratch_bytes_mma(int64_t n_cap, int K, int Ca, int Cb) {
  WgPlan pl = wg_plan(n_cap, K, Ca, Cb);
  if (!pl.ok) return 0;
  return (int64_t)pl.nchunks * K * Ca * Cb * 4 + 256;
}

int wgrad_table_mma(const float* A, int64_t lda, const float* B, int64_t ldb, const int32_t* table, int64_t ld, int K,
                    int gather_b, int64_t n_cap, const int32_t* n_dev, int Ca, int Cb, float* dW, int accumulate,
                    int math_mode, void* scratch, int64_t scratch_bytes, cudaStream_t st, const DenseGeom* geom) {
  const int64_t elems = (int64_t)K * Ca * Cb;
  if (n_cap == 0) {
    if (!accumulate) B200_CUDA(cudaMemsetAsync(dW, 0, (size_t)elems * 4, st));
    return 0;
  }
  WgPlan pl = wg_plan(n_cap, K, Ca, Cb);
  B200_CHECK_ARG(pl.ok, "wgrad_mma: unsupported shape");
  B200_CHECK_ARG(scratch_bytes >= (int64_t)pl.nchunks * elems * 4, "wgrad_mma: scratch too small (%lld < %lld)",
                 (long long)scratch_bytes, (long long)((int64_t)pl.nchunks * elems * 4));
  WgParams p;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.table = table; p.ld = ld; p.K = K; p.gather_b = gather_b;
  p.n_cap = n_cap; p.n_dev = n_dev; p.Ca = Ca; p.Cb = Cb; p.cb_per_z = pl.cb_per_z;
  p.rows_per_cta = pl.rows_per_cta; p.part = reinterpret_cast<float*>(scratch);
  p.x3 = (math_mode == B200SCN_MATH_TF32X3) ? 1 : 0;
  static int fused = -1;
  if (fused < 0) {
    const char* e = getenv("B200SCN_WG_FUSED_FOLD");
    fused = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  p.dW = dW; p.accumulate = accumulate; p.fused_fold = fused;
  p.use_geom = geom != nullptr ? 1 : 0;
  if (geom != nullptr) p.geom = *geom; else p.geom = DenseGeom{};
  B200_CHECK_ARG(K <= 64 && pl.zsplits <= 8, "wgrad_mma: K > 64 or more than 8 column splits");
  int rc = 0;
  switch (pl.tpw) {
    case 1: rc = wg_launch<1>(p, pl, st); break;
    case 2: rc = wg_launch<2>(p, pl, st); break;
    case 4: rc = wg_launch<4>(p, pl, st); break;
    case 8: rc = wg_launch<8>(p, pl, st); break;
    default: rc = wg_launch<16>(p, pl, st); break;
  }
  if (rc) return rc;
  if (!p.fused_fold) {
    int64_t gsz = cdiv(elems, 256);
    if (gsz > 1024) gsz = 1024;
    k_wgrad_fold<<<(int)gsz, 256, 0, st>>>(reinterpret_cast<const float*>(scratch), elems, pl.rows_per_cta, pl.nchunks,
                                           n_cap, n_dev, dW, accumulate);
  }
  count_launches(p.fused_fold ? 1 : 2);
  B200_CHECK_LAUNCH("wgrad_mma");
  return 0;
}

}  // namespace b200scn
