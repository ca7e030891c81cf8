// placeholder -- replaced below by the tensor-core kernels
#include "common.cuh"
namespace b200scn {
bool conv_table_mma_supported(int, int, int, int64_t, int64_t, const void*, const void*) { return false; }
int conv_table_mma(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int, const int32_t*, int64_t, int,
                   int64_t, const int32_t*, int, int, float*, int64_t, int, int, cudaStream_t) {
  set_error("conv_table_mma: not built");
  return B200SCN_ERR_UNSUPPORTED;
}
}  // namespace b200scn
