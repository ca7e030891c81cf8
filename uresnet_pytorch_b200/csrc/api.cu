// libb200scn: error plumbing, device queries and the math-mode dispatch of the conv entry points.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace b200scn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200SCN_PDL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

// conv_simt.cu
int conv_table_simt(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int, const int32_t*, int64_t, int,
                    int64_t, const int32_t*, int, int, float*, int64_t, int, cudaStream_t);
int conv_parent_simt(const float*, int64_t, const float*, int64_t, int64_t, int64_t, const int32_t*, const uint8_t*,
                     int, int64_t, const int32_t*, int, int, float*, int64_t, cudaStream_t);
int64_t wgrad_scratch_bytes_simt(int64_t, int, int, int);
int wgrad_table_simt(const float*, int64_t, const float*, int64_t, const int32_t*, int64_t, int, int, int64_t,
                     const int32_t*, int, int, float*, int, void*, int64_t, cudaStream_t, const DenseGeom&);
// conv_pairs.cu
bool conv_pairs_supported(int Cin, int Cout, int K, int64_t ldi, const void* in, int64_t n_cap, int64_t ld,
                          int64_t ldo, int64_t n_in_max);
int conv_rows_pairs(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int, const int32_t*, int64_t,
                    const int32_t*, const uint8_t*, int, int64_t, const int32_t*, int, int, float*, int64_t, int, int,
                    cudaStream_t);
// wgrad_mma.cu
bool wgrad_mma_supported(int64_t, int, int, int, int64_t, int64_t, const void*, const void*);
int64_t wgrad_scratch_bytes_mma(int64_t, int, int, int);
int wgrad_table_mma(const float*, int64_t, const float*, int64_t, const int32_t*, int64_t, int, int, int64_t,
                    const int32_t*, int, int, float*, int, int, void*, int64_t, cudaStream_t, const DenseGeom*);

}  // namespace b200scn

using namespace b200scn;

extern "C" const char* b200scn_last_error(void) { return g_err; }
extern "C" int b200scn_version(void) { return 1; }
extern "C" int64_t b200scn_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int b200scn_sm_count(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("sm_count: no CUDA device");
    return B200SCN_ERR_CUDA;
  }
  return sm_count();
}

extern "C" int b200scn_conv_table(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si,
                                  int64_t w_so, int kflip, const int32_t* table, int64_t ld, int K, int64_t n_out_cap,
                                  const int32_t* n_out_dev, int Cin, int Cout, float* out, int64_t ldo, int accumulate,
                                  int math_mode, void* stream) {
  B200_CHECK_ARG(Cin >= 1 && Cout >= 1 && K >= 1, "conv_table: bad shape Cin=%d Cout=%d K=%d", Cin, Cout, K);
  B200_CHECK_ARG(table != nullptr || K == 1, "conv_table: table == NULL requires K == 1");
  B200_CHECK_ARG(table == nullptr || ld >= n_out_cap, "conv_table: ld < n_out_cap");
  B200_CHECK_ARG(ldi >= Cin && ldo >= Cout, "conv_table: row stride smaller than channel count");
  B200_CHECK_ARG(math_mode >= 0 && math_mode <= 2, "conv_table: bad math mode %d", math_mode);
  if (math_mode != B200SCN_MATH_FP32 && conv_pairs_supported(Cin, Cout, K, ldi, in, n_out_cap, ld, ldo, 1ll << 26))
    return conv_rows_pairs(in, ldi, W, w_sk, w_si, w_so, kflip, table, ld, nullptr, nullptr, K, n_out_cap, n_out_dev,
                           Cin, Cout, out, ldo, accumulate, math_mode, (cudaStream_t)stream);
  return conv_table_simt(in, ldi, W, w_sk, w_si, w_so, kflip, table, ld, K, n_out_cap, n_out_dev, Cin, Cout, out, ldo,
                         accumulate, (cudaStream_t)stream);
}

extern "C" int b200scn_conv_parent(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si,
                                   int64_t w_so, const int32_t* parent, const uint8_t* kofs, int K, int64_t n_out_cap,
                                   const int32_t* n_out_dev, int Cin, int Cout, float* out, int64_t ldo, int math_mode,
                                   void* stream) {
  B200_CHECK_ARG(Cin >= 1 && Cout >= 1 && K >= 1, "conv_parent: bad shape Cin=%d Cout=%d K=%d", Cin, Cout, K);
  B200_CHECK_ARG(parent != nullptr && kofs != nullptr, "conv_parent: parent/kofs must not be NULL");
  B200_CHECK_ARG(ldi >= Cin && ldo >= Cout, "conv_parent: row stride smaller than channel count");
  B200_CHECK_ARG(math_mode >= 0 && math_mode <= 2, "conv_parent: bad math mode %d", math_mode);
  if (math_mode != B200SCN_MATH_FP32 && conv_pairs_supported(Cin, Cout, K, ldi, in, n_out_cap, 0, ldo, 1ll << 26))
    return conv_rows_pairs(in, ldi, W, w_sk, w_si, w_so, 0, nullptr, 0, parent, kofs, K, n_out_cap, n_out_dev, Cin,
                           Cout, out, ldo, 0, math_mode, (cudaStream_t)stream);
  return conv_parent_simt(in, ldi, W, w_sk, w_si, w_so, parent, kofs, K, n_out_cap, n_out_dev, Cin, Cout, out, ldo,
                          (cudaStream_t)stream);
}

extern "C" int64_t b200scn_wgrad_scratch_bytes(int64_t n_cap, int K, int Ca, int Cb) {
  const int64_t a = wgrad_scratch_bytes_simt(n_cap, K, Ca, Cb), b = wgrad_scratch_bytes_mma(n_cap, K, Ca, Cb);
  return a > b ? a : b;
}

extern "C" int b200scn_wgrad_table(const float* A, int64_t lda, const float* B, int64_t ldb, const int32_t* table,
                                   int64_t ld, int K, int gather_b, int64_t n_cap, const int32_t* n_dev, int Ca, int Cb,
                                   float* dW, int accumulate, int math_mode, void* scratch, int64_t scratch_bytes,
                                   void* stream) {
  B200_CHECK_ARG(Ca >= 1 && Cb >= 1 && K >= 1, "wgrad_table: bad shape Ca=%d Cb=%d K=%d", Ca, Cb, K);
  B200_CHECK_ARG(table != nullptr || K == 1, "wgrad_table: table == NULL requires K == 1");
  B200_CHECK_ARG(lda >= Ca && ldb >= Cb, "wgrad_table: row stride smaller than channel count");
  B200_CHECK_ARG(math_mode >= 0 && math_mode <= 2, "wgrad_table: bad math mode %d", math_mode);
  if (math_mode != B200SCN_MATH_FP32 && wgrad_mma_supported(n_cap, K, Ca, Cb, lda, ldb, A, B))
    return wgrad_table_mma(A, lda, B, ldb, table, ld, K, gather_b, n_cap, n_dev, Ca, Cb, dW, accumulate, math_mode,
                           scratch, scratch_bytes, (cudaStream_t)stream, nullptr);
  return wgrad_table_simt(A, lda, B, ldb, table, ld, K, gather_b, n_cap, n_dev, Ca, Cb, dW, accumulate, scratch,
                          scratch_bytes, (cudaStream_t)stream, DenseGeom{});
}
