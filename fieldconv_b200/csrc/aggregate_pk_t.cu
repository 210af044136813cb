// k_aggregate<., TRANSPOSE = true, PACK = true, ...> instantiations (K5a writing the packed fp16 operand planes).
#include "aggregate_kernel.cuh"

namespace fcb {
int aggregate_transposed_packed(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out, int64_t N,
                                int C, int B, int R, const float* feat_amax, const float* norm, float* bound, cudaStream_t st) {
    return dispatch_aggregate<true, true>(feat, rowptr, rec, rot, out, N, C, B, R, nullptr, feat_amax, norm, bound, st);
}
}  // namespace fcb
