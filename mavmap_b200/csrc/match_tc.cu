// match_tc.cu — tcgen05 tensor-core candidate selection for the matcher (placeholder until
// the TF32 distance-GEMM kernel lands; AUTO falls back to the exact SIMT path).
#include "common.cuh"
#include "match.cuh"
namespace mm {
int match_tc_pairs(const float*, const float*, int, const int64_t*, const int32_t*, const int32_t*,
                   const PairJob*, int, double, Knn2*, Knn2*, cudaStream_t, bool required) {
  if (required) set_error("tcgen05 matcher path not built");
  return MM_ERR_UNSUPPORTED;
}
}  // namespace mm
