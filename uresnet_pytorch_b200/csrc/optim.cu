// Multi-tensor Adam in ONE launch: the optimiser step of
// /root/reference/uresnet/trainval.py:21-30 (torch.optim.Adam defaults: no weight decay, no amsgrad) over
// every parameter of the model at once.  job = int64[4] = {param, grad, exp_avg, exp_avg_sq} pointers plus
// int64 numel; the step counter lives in device memory so that the launch can sit inside a CUDA graph:
// all CTAs read it, the last CTA to finish increments it (ticket counter, reset by that CTA).
#include "common.cuh"

namespace b200scn {
namespace {

__global__ void k_adam(const int64_t* __restrict__ jobs, int32_t* __restrict__ step_dev, unsigned int* __restrict__ ticket,
                       float lr, float beta1, float beta2, float eps) {
  const int64_t* j = jobs + (size_t)blockIdx.y * 5;
  float* p = reinterpret_cast<float*>(j[0]);
  const float* g = reinterpret_cast<const float*>(j[1]);
  float* m = reinterpret_cast<float*>(j[2]);
  float* v = reinterpret_cast<float*>(j[3]);
  const int64_t n = j[4];
  const int t = *step_dev + 1;
  // bias corrections in double (1 - beta^t loses most of its bits in float for small t), once per CTA
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    s_bc[0] = (float)(1.0 - pow((double)beta1, (double)t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, (double)t));
  }
  __syncthreads();
  const float bc2s = s_bc[1];
  const float step_size = lr / s_bc[0];
  auto upd = [&](float& pe, float ge, float& me, float& ve) {
    me = beta1 * me + (1.f - beta1) * ge;
    ve = beta2 * ve + (1.f - beta2) * ge * ge;
    pe -= step_size * (me / (sqrtf(ve) / bc2s + eps));
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = vec ? (n >> 2) : 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    float4 pe = reinterpret_cast<float4*>(p)[e], me = reinterpret_cast<float4*>(m)[e], ve = reinterpret_cast<float4*>(v)[e];
    const float4 ge = reinterpret_cast<const float4*>(g)[e];
    upd(pe.x, ge.x, me.x, ve.x);
    upd(pe.y, ge.y, me.y, ve.y);
    upd(pe.z, ge.z, me.z, ve.z);
    upd(pe.w, ge.w, me.w, ve.w);
    reinterpret_cast<float4*>(m)[e] = me;
    reinterpret_cast<float4*>(v)[e] = ve;
    reinterpret_cast<float4*>(p)[e] = pe;
  }
  for (int64_t e = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float pe = p[e], me = m[e], ve = v[e];
    upd(pe, g[e], me, ve);
    m[e] = me;
    v[e] = ve;
    p[e] = pe;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int total = gridDim.x * gridDim.y;
    if (atomicAdd(ticket, 1u) == total - 1u) {
      *step_dev = t;
      *ticket = 0u;
    }
  }
}

}  // namespace
}  // namespace b200scn

using namespace b200scn;

extern "C" int b200scn_adam_step(const int64_t* jobs_dev, int n_jobs, int64_t max_elems, int32_t* step_dev,
                                 uint32_t* ticket_dev, float lr, float beta1, float beta2, float eps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_jobs <= 0) return 0;
  B200_CHECK_ARG(n_jobs <= 65535, "adam_step: too many tensors");
  int64_t gx = cdiv(max_elems < 1 ? 1 : max_elems, 256 * 4 * 4);  // float4 per thread, ~4 per thread
  if (gx > 32) gx = 32;
  k_adam<<<dim3((unsigned)gx, (unsigned)n_jobs), 256, 0, st>>>(jobs_dev, step_dev, ticket_dev, lr, beta1, beta2, eps);
  count_launches(1);
  B200_CHECK_LAUNCH("adam_step");
  return 0;
}
