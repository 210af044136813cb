// k_aggregate<., TRANSPOSE = true, PACK = false, ...> instantiations (K5a: G = sum conj(sten) gy over the by-source CSR).
#include "aggregate_kernel.cuh"

namespace fcb {
int aggregate_transposed_f32(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out, int64_t N,
                             int C, int B, int R, float* amax, cudaStream_t st) {
    return dispatch_aggregate<true, false>(feat, rowptr, rec, rot, out, N, C, B, R, amax, nullptr, nullptr, nullptr, st);
}
}  // namespace fcb
