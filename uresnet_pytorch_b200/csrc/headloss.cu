// Fused segmentation head + loss: Linear(m -> nc) on the OutputLayer rows,
// per-event mean cross-entropy summed over events, per-event accuracy, and the complete backward of
// that chain, without a Python loop over events, host syncs or a dozen elementwise launches.
//
// Reference semantics: /root/reference/uresnet/models/uresnet_sparse.py:25,36 (self.linear) and
// :44-82 (SegmentationLoss.forward): for every event b present in the batch-id column
//     total_loss += mean_{r in b} CE(logits_r, label_r) [* pixel weight_r],   total_acc += hits_b / |b|.
// With c_b = |b| that is a plain sum over rows of CE_r * pw_r / c_{b(r)} -- so after one integer
// counting pass every row is independent:
//     k_hl_count : c_b (integer atomics: exact, order-free)
//     k_hl_main  : per row logits, softmax, CE, hit; d_logits = pw_r / c_b * (p - onehot);
//                  d_feat = W^T d_logits;  per-CTA partial sums of loss, acc, dW, db
//     k_hl_final : folds the partials in CTA order (deterministic)
// All gradients are those of total_loss (upstream gradient 1).
#include "common.cuh"

namespace b200scn {
namespace {

constexpr int kMaxClasses = 8;
constexpr int kMaxFeat = 64;
constexpr int kThreads = 256;

__device__ __forceinline__ int read_id(const void* ids, int dtype, int64_t ld, int64_t r) {
  switch (dtype) {
    case B200SCN_I32: return reinterpret_cast<const int32_t*>(ids)[r * ld];
    case B200SCN_I64: return (int)reinterpret_cast<const int64_t*>(ids)[r * ld];
    case B200SCN_F32: return (int)reinterpret_cast<const float*>(ids)[r * ld];
    default: return (int)reinterpret_cast<const double*>(ids)[r * ld];
  }
}

__global__ void k_hl_count(const void* __restrict__ ids, int id_dtype, int64_t ld_ids, int64_t n_cap,
                           const int32_t* __restrict__ n_dev, int nb, int32_t* __restrict__ cnt) {
  // per-CTA histogram in shared memory first (events are few and a whole event hits ONE counter), then one
  // global atomic per (CTA, event present in it)
  constexpr int kHist = 1024;
  __shared__ int hist[kHist];
  const int n = live_rows(n_dev, n_cap);
  const bool use_hist = nb <= kHist;
  if (use_hist)
    for (int e = threadIdx.x; e < nb; e += blockDim.x) hist[e] = 0;
  __syncthreads();
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int b = read_id(ids, id_dtype, ld_ids, r);
    if (b >= 0 && b < nb) atomicAdd(use_hist ? hist + b : cnt + b, 1);
  }
  __syncthreads();
  if (use_hist)
    for (int e = threadIdx.x; e < nb; e += blockDim.x)
      if (hist[e] != 0) atomicAdd(cnt + e, hist[e]);
}

struct HlParams {
  const float* feat; int64_t ldf;
  const float* W; const float* bias;  // W[nc][m] (torch.nn.Linear layout)
  const float* label; int64_t ld_label;
  const float* pix_w; int64_t ld_pw;  // optional per-row weights
  const void* ids; int id_dtype; int64_t ld_ids;
  int64_t n_cap; const int32_t* n_dev;
  int m, nc, nb;
  const int32_t* cnt;
  float* logits; int64_t ld_logits;  // optional
  float* dfeat; int64_t ld_dfeat;    // optional (training)
  float* part;                       // [grid][4 + nc * m + nc]: loss, acc, bad ids, bad labels, dW, db
};

__global__ void __launch_bounds__(kThreads) k_hl_main(const HlParams p) {
  __shared__ float sW[kMaxClasses * kMaxFeat + kMaxClasses];
  __shared__ float red[kThreads / 32][4 + kMaxClasses * 17];
  const int n = live_rows(p.n_dev, p.n_cap);
  const int m = p.m, nc = p.nc;
  for (int e = threadIdx.x; e < nc * m; e += blockDim.x) sW[e] = p.W[e];
  for (int e = threadIdx.x; e < nc; e += blockDim.x) sW[kMaxClasses * kMaxFeat + e] = p.bias ? p.bias[e] : 0.f;
  __syncthreads();
  const float* sb = sW + kMaxClasses * kMaxFeat;
  const bool train = p.dfeat != nullptr;

  // per-thread accumulators: loss, acc, db[c], and dW in slices of 16 features per pass
  float a_loss = 0.f, a_acc = 0.f, a_bad_id = 0.f, a_bad_lab = 0.f;
  float a_db[kMaxClasses];
#pragma unroll
  for (int c = 0; c < kMaxClasses; ++c) a_db[c] = 0.f;
  const int nvals = 4 + nc * m + nc;
  float* my_part = p.part + (int64_t)blockIdx.x * nvals;

  for (int f0 = 0; f0 < m; f0 += 16) {  // m <= 16 for the reference's models: one pass
    float a_dw[kMaxClasses][16];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
#pragma unroll
      for (int j = 0; j < 16; ++j) a_dw[c][j] = 0.f;
    const bool first = (f0 == 0);
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
      const float* x = p.feat + r * p.ldf;
      float lg[kMaxClasses];
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c) lg[c] = c < nc ? sb[c] : -INFINITY;
      for (int j = 0; j < m; ++j) {
        const float xv = __ldg(x + j);
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
          if (c < nc) lg[c] = fmaf(xv, sW[c * m + j], lg[c]);
      }
      if (first && p.logits != nullptr)
        for (int c = 0; c < nc; ++c) p.logits[r * p.ld_logits + c] = lg[c];
      const int lab = (int)p.label[r * p.ld_label];
      const int b = read_id(p.ids, p.id_dtype, p.ld_ids, r);
      const int cb = (b >= 0 && b < p.nb) ? p.cnt[b] : 0;
      const float inv = cb > 0 ? 1.f / (float)cb : 0.f;
      const float pw = p.pix_w ? p.pix_w[r * p.ld_pw] : 1.f;
      float mx = lg[0];
      int arg = 0;
#pragma unroll
      for (int c = 1; c < kMaxClasses; ++c)
        if (c < nc && lg[c] > mx) { mx = lg[c]; arg = c; }
      float se = 0.f, ex[kMaxClasses];
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c) {
        ex[c] = c < nc ? expf(lg[c] - mx) : 0.f;
        se += ex[c];
      }
      const float lse = mx + logf(se);
      const bool lab_ok = lab >= 0 && lab < nc;
      float lg_lab = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c)
        if (c == lab) lg_lab = lg[c];
      if (first) {
        if (lab_ok) a_loss += (lse - lg_lab) * pw * inv;
        a_acc += (arg == lab) ? inv : 0.f;
        // rows the reference would have crashed on (sample id outside [0, nb), label outside [0, nc)): they
        // contribute nothing here, and are COUNTED so the caller can refuse the step
        a_bad_id += (b >= 0 && b < p.nb) ? 0.f : 1.f;
        a_bad_lab += lab_ok ? 0.f : 1.f;
      }
      if (train) {
        const float scale = lab_ok ? pw * inv : 0.f;
        const float rse = 1.f / se;
        float dl[kMaxClasses];
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
          dl[c] = c < nc ? scale * (ex[c] * rse - (c == lab ? 1.f : 0.f)) : 0.f;
          if (first) a_db[c] += dl[c];
        }
        if (first) {
          for (int j = 0; j < m; ++j) {
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < kMaxClasses; ++c)
              if (c < nc) d = fmaf(dl[c], sW[c * m + j], d);
            p.dfeat[r * p.ld_dfeat + j] = d;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (f0 + j < m) {
            const float xv = __ldg(x + f0 + j);
#pragma unroll
            for (int c = 0; c < kMaxClasses; ++c) a_dw[c][j] = fmaf(dl[c], xv, a_dw[c][j]);
          }
        }
      }
    }
    // CTA reduction in fixed order: lanes by shuffle tree, warps in index order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    auto warp_sum = [&](float v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      return v;
    };
    if (first) {
      const float l = warp_sum(a_loss), a = warp_sum(a_acc), bi = warp_sum(a_bad_id), bl = warp_sum(a_bad_lab);
      if (lane == 0) { red[warp][0] = l; red[warp][1] = a; red[warp][2] = bi; red[warp][3] = bl; }
    }
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) {
      if (c < nc) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = warp_sum(a_dw[c][j]);
          if (lane == 0) red[warp][4 + c * 17 + j] = v;
        }
        if (first) {
          const float v = warp_sum(a_db[c]);
          if (lane == 0) red[warp][4 + c * 17 + 16] = v;
        }
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 4 + kMaxClasses * 17; e += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarp; ++w) s += red[w][e];
      if (e < 4) {
        if (first) my_part[e] = s;
      } else {
        const int c = (e - 4) / 17, j = (e - 4) % 17;
        if (c < nc) {
          if (j < 16) { if (f0 + j < m) my_part[4 + c * m + f0 + j] = s; }
          else if (first) my_part[4 + nc * m + c] = s;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void k_hl_final(const float* __restrict__ part, int nparts, int nvals, int m, int nc, float* __restrict__ loss_acc,
                           float* __restrict__ dW, float* __restrict__ db, int accumulate) {
  for (int e = threadIdx.x; e < nvals; e += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < nparts; ++b) s += part[(int64_t)b * nvals + e];
    if (e < 4) loss_acc[e] = s;
    else if (e < 4 + nc * m) { if (dW) dW[e - 4] = accumulate ? dW[e - 4] + s : s; }
    else if (db) db[e - 4 - nc * m] = accumulate ? db[e - 4 - nc * m] + s : s;
  }
}


// ---- inference outputs ---------------------------------------------------------
// What /root/reference/uresnet/trainval.py:127-130 ships to the host after a forward pass -- 'segmentation'
// (logits) and 'softmax' -- plus the argmax prediction and, when labels are given, the per-event sums behind
// the confusion / accuracy part of compute_metrics_sparse (/root/reference/uresnet/utils.py:127-200):
//   per event b:  [0] pixels, [1] correct pixels, [2] sum of -log softmax[label] (log-loss numerator),
//                 [3] sum of softmax[label] (correct_softmax numerator),
//                 [4 + c] sum of softmax[label] over the pixels of true class c (class_mean_softmax numerator),
//                 [4 + nc + c * nc + c2]        pixels of true class c predicted c2 (confusion_matrix),
//                 [4 + nc + nc*nc + c * nc + c2] their summed energy (energy_confusion_matrix).
// One pass: per-CTA tables in shared memory, per-CTA partials to HBM, folded in CTA order in double.
struct HpParams {
  const float* feat; int64_t ldf;
  const float* W; const float* bias;
  const float* label; int64_t ld_label;   // optional
  const void* ids; int id_dtype; int64_t ld_ids;
  const void* energy; int en_dtype; int64_t ld_en;  // optional
  int64_t n_cap; const int32_t* n_dev;
  int m, nc, nb;
  float* logits; int64_t ld_logits;
  float* softmax; int64_t ld_softmax;
  int32_t* pred;
  float* part;  // [grid][nb * E], E = 4 + nc + 2 nc nc; NULL without labels
};

__device__ __forceinline__ float read_val(const void* p, int dtype, int64_t ld, int64_t r) {
  switch (dtype) {
    case B200SCN_I32: return (float)reinterpret_cast<const int32_t*>(p)[r * ld];
    case B200SCN_I64: return (float)reinterpret_cast<const int64_t*>(p)[r * ld];
    case B200SCN_F32: return reinterpret_cast<const float*>(p)[r * ld];
    default: return (float)reinterpret_cast<const double*>(p)[r * ld];
  }
}

__global__ void __launch_bounds__(kThreads) k_hp_main(const HpParams p) {
  extern __shared__ float hp_sm[];  // [nc * m + nc] weights, then [nb * E] tables
  const int n = live_rows(p.n_dev, p.n_cap);
  const int m = p.m, nc = p.nc;
  const int E = 4 + nc + 2 * nc * nc;
  float* sW = hp_sm;
  float* sb = sW + nc * m;
  float* tab = sb + nc;
  const bool metrics = p.part != nullptr && p.label != nullptr;
  for (int e = threadIdx.x; e < nc * m; e += blockDim.x) sW[e] = p.W[e];
  for (int e = threadIdx.x; e < nc; e += blockDim.x) sb[e] = p.bias ? p.bias[e] : 0.f;
  if (metrics)
    for (int e = threadIdx.x; e < p.nb * E; e += blockDim.x) tab[e] = 0.f;
  __syncthreads();
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const float* x = p.feat + r * p.ldf;
    float lg[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) lg[c] = c < nc ? sb[c] : -INFINITY;
    for (int j = 0; j < m; ++j) {
      const float xv = __ldg(x + j);
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c)
        if (c < nc) lg[c] = fmaf(xv, sW[c * m + j], lg[c]);
    }
    float mx = lg[0];
    int arg = 0;
#pragma unroll
    for (int c = 1; c < kMaxClasses; ++c)
      if (c < nc && lg[c] > mx) { mx = lg[c]; arg = c; }
    float se = 0.f, ex[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) {
      ex[c] = c < nc ? expf(lg[c] - mx) : 0.f;
      se += ex[c];
    }
    const float rse = 1.f / se;
    if (p.logits != nullptr)
      for (int c = 0; c < nc; ++c) p.logits[r * p.ld_logits + c] = lg[c];
    if (p.softmax != nullptr)
      for (int c = 0; c < nc; ++c) p.softmax[r * p.ld_softmax + c] = ex[c] * rse;
    if (p.pred != nullptr) p.pred[r] = arg;
    if (metrics) {
      const int lab = (int)p.label[r * p.ld_label];
      const int b = read_id(p.ids, p.id_dtype, p.ld_ids, r);
      if (b >= 0 && b < p.nb && lab >= 0 && lab < nc) {
        float pl = 0.f, lgl = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
          if (c == lab) { pl = ex[c] * rse; lgl = lg[c]; }
        float* t = tab + b * E;
        atomicAdd(t + 0, 1.f);
        if (arg == lab) atomicAdd(t + 1, 1.f);
        atomicAdd(t + 2, (mx + logf(se)) - lgl);
        atomicAdd(t + 3, pl);
        atomicAdd(t + 4 + lab, pl);
        atomicAdd(t + 4 + nc + lab * nc + arg, 1.f);
        if (p.energy != nullptr)
          atomicAdd(t + 4 + nc + nc * nc + lab * nc + arg, read_val(p.energy, p.en_dtype, p.ld_en, r));
      }
    }
  }
  if (metrics) {
    __syncthreads();
    float* my = p.part + (int64_t)blockIdx.x * p.nb * E;
    for (int e = threadIdx.x; e < p.nb * E; e += blockDim.x) my[e] = tab[e];
  }
}

__global__ void k_hp_final(const float* __restrict__ part, int nparts, int nvals, double* __restrict__ out) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nvals; e += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nparts; ++b) s += (double)part[(int64_t)b * nvals + e];
    out[e] = s;
  }
}

constexpr int kHpMaxTable = 8192;  // floats of per-CTA tables in shared memory

int hl_grid(int64_t n_cap) {
  int64_t g = cdiv(n_cap < 1 ? 1 : n_cap, kThreads);
  const int64_t cap = 2 * (int64_t)sm_count();
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

extern "C" int64_t b200scn_head_loss_scratch_bytes(int64_t n_cap, int m, int nc, int nb) {
  return 256 + (int64_t)nb * 4 + 256 + (int64_t)hl_grid(n_cap) * (4 + nc * m + nc) * 4;
}

extern "C" int b200scn_head_loss(const float* feat, int64_t ldf, const float* W, const float* bias, const float* label,
                                 int64_t ld_label, const float* pix_w, int64_t ld_pw, const void* ids, int id_dtype,
                                 int64_t ld_ids, int64_t n_cap, const int32_t* n_dev, int m, int nc, int nb,
                                 float* logits, int64_t ld_logits, float* loss_acc, float* dfeat, int64_t ld_dfeat,
                                 float* dW, float* db, int accumulate, void* scratch, int64_t scratch_bytes,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(m >= 1 && m <= kMaxFeat && nc >= 1 && nc <= kMaxClasses, "head_loss: need m <= %d, nc <= %d", kMaxFeat,
                 kMaxClasses);
  B200_CHECK_ARG(nb >= 1, "head_loss: nb must be >= 1");
  B200_CHECK_ARG(scratch_bytes >= b200scn_head_loss_scratch_bytes(n_cap, m, nc, nb), "head_loss: scratch too small");
  int32_t* cnt = reinterpret_cast<int32_t*>(scratch);
  const int64_t cnt_bytes = ((int64_t)nb * 4 + 255) / 256 * 256;
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + cnt_bytes);
  const int grid = hl_grid(n_cap);
  B200_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nb * 4, st));
  if (n_cap > 0)
    k_hl_count<<<grid, kThreads, 0, st>>>(ids, id_dtype, ld_ids, n_cap, n_dev, nb, cnt);
  HlParams p;
  p.feat = feat; p.ldf = ldf; p.W = W; p.bias = bias; p.label = label; p.ld_label = ld_label; p.pix_w = pix_w;
  p.ld_pw = ld_pw; p.ids = ids; p.id_dtype = id_dtype; p.ld_ids = ld_ids; p.n_cap = n_cap; p.n_dev = n_dev; p.m = m;
  p.nc = nc; p.nb = nb; p.cnt = cnt; p.logits = logits; p.ld_logits = ld_logits; p.dfeat = dfeat; p.ld_dfeat = ld_dfeat;
  p.part = part;
  k_hl_main<<<grid, kThreads, 0, st>>>(p);
  k_hl_final<<<1, 128, 0, st>>>(part, grid, 4 + nc * m + nc, m, nc, loss_acc, dW, db, accumulate);
  count_launches(3);
  B200_CHECK_LAUNCH("head_loss");
  return 0;
}

extern "C" int64_t b200scn_head_predict_scratch_bytes(int64_t n_cap, int nc, int nb) {
  return 256 + (int64_t)hl_grid(n_cap) * nb * (4 + nc + 2 * nc * nc) * 4;
}

extern "C" int b200scn_head_predict(const float* feat, int64_t ldf, const float* W, const float* bias,
                                    const float* label, int64_t ld_label, const void* ids, int id_dtype,
                                    int64_t ld_ids, const void* energy, int energy_dtype, int64_t ld_energy,
                                    int64_t n_cap, const int32_t* n_dev, int m, int nc, int nb, float* logits,
                                    int64_t ld_logits, float* softmax, int64_t ld_softmax, int32_t* pred,
                                    double* metrics, void* scratch, int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(m >= 1 && m <= kMaxFeat && nc >= 1 && nc <= kMaxClasses, "head_predict: need m <= %d, nc <= %d",
                 kMaxFeat, kMaxClasses);
  const int E = 4 + nc + 2 * nc * nc;
  const bool want = metrics != nullptr && label != nullptr;
  if (want) {
    B200_CHECK_ARG(ids != nullptr && nb >= 1, "head_predict: metrics need the sample-id column and nb >= 1");
    B200_CHECK_ARG((int64_t)nb * E <= kHpMaxTable, "head_predict: %d events x %d classes exceed the metric tables "
                   "(split the batch)", nb, nc);
    B200_CHECK_ARG(scratch_bytes >= b200scn_head_predict_scratch_bytes(n_cap, nc, nb), "head_predict: scratch too small");
  }
  const int grid = hl_grid(n_cap);
  HpParams p;
  p.feat = feat; p.ldf = ldf; p.W = W; p.bias = bias; p.label = label; p.ld_label = ld_label; p.ids = ids;
  p.id_dtype = id_dtype; p.ld_ids = ld_ids; p.energy = energy; p.en_dtype = energy_dtype; p.ld_en = ld_energy;
  p.n_cap = n_cap; p.n_dev = n_dev; p.m = m; p.nc = nc; p.nb = nb; p.logits = logits; p.ld_logits = ld_logits;
  p.softmax = softmax; p.ld_softmax = ld_softmax; p.pred = pred;
  p.part = want ? reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 256) : nullptr;
  const size_t smem = (size_t)(nc * m + nc + (want ? nb * E : 0)) * 4;
  k_hp_main<<<grid, kThreads, smem, st>>>(p);
  count_launches(1);
  if (want) {
    const int nvals = nb * E;
    k_hp_final<<<(nvals + 127) / 128, 128, 0, st>>>(p.part, grid, nvals, metrics);
    count_launches(1);
  }
  B200_CHECK_LAUNCH("head_predict");
  return 0;
}
