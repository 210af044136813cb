// F2: the edge aggregations of TransField, the learned 'gradient' that lifts scalar features to tangent
// vectors (nn/trans_field.py:96-110 of the reference) — two scatter_adds over the support edges, replaced by deterministic
// segmented reductions over the CSR rows of the dense-stencil plan (fcb_plan_build_dense), any stencil (E,R,2) accepted:
//   agg[i, c, r] = sum_{e -> i} (x[src(e), c] - x[i, c]) * s1[e, r]   (complex;  s1 = lift_sten[:, :, 1]; contribAng = -agg,
//                                                                      the difference is formed per edge as in :104, so the
//                                                                      result carries no x_i S1 - sum x_j s1 cancellation)
//   agg[i, Ci, r] = sum_{e -> i} s1[e, r]                              (S1: what the adjoint needs for the -x[i] part)
//   mag[i, c, r] = sum_{e -> i} x[src(e), c] * softAbs(s0[e, r])       (real;     s0 = lift_sten[:, :, 0], utils/field.py:29-37)
// The tiny per-vertex weighting (nn/trans_field.py:10-25) is composed by the host module from these.  One thread per
// (row, channel); x is real.  The transposed kernel is the adjoint with respect to x: the +x[src] part over the by-source
// rows, the -x[i] part through S1.
#include "common.cuh"

namespace fcb {

constexpr int LIFT_MAX_R = 8;

__device__ __forceinline__ float soft_abs_c(float2 z) {
    return ((fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f)) ? 0.f : sqrtf(z.x * z.x + z.y * z.y);
}

__global__ void __launch_bounds__(256) k_lift_aggregate(const float* __restrict__ x, const float2* __restrict__ sten,
                                                        const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                                                        const int32_t* __restrict__ perm, float2* __restrict__ agg,
                                                        float* __restrict__ mag, int64_t N, int Ci, int R) {
    const int C1 = Ci + 1;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C1) return;
    const int64_t row = t / C1;
    const int c = (int)(t - row * C1);
    float2 a[LIFT_MAX_R];
    float m[LIFT_MAX_R];
#pragma unroll
    for (int r = 0; r < LIFT_MAX_R; ++r) { a[r] = make_float2(0.f, 0.f); m[r] = 0.f; }
    const float xi = c < Ci ? x[row * Ci + c] : 0.f;
    const int p1 = rowptr[row + 1];
    for (int p = rowptr[row]; p < p1; ++p) {
        const int64_t e = perm[p];
        const float xs = c < Ci ? x[(int64_t)nbr[p] * Ci + c] : 1.0f;
        const float xv = c < Ci ? xs - xi : 1.0f;
        const float2* s = sten + e * R * 2;
#pragma unroll
        for (int r = 0; r < LIFT_MAX_R; ++r) {
            if (r < R) {
                const float2 s0 = s[2 * r], s1 = s[2 * r + 1];
                a[r].x = fmaf(xv, s1.x, a[r].x);
                a[r].y = fmaf(xv, s1.y, a[r].y);
                m[r] = fmaf(xs, soft_abs_c(s0), m[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < LIFT_MAX_R; ++r) {
        if (r < R) {
            agg[(row * C1 + c) * R + r] = a[r];
            if (c < Ci) mag[(row * Ci + c) * R + r] = m[r];
        }
    }
}

// gx[j, c] = sum_{e: src = j} sum_r ( Re(conj(g_agg[tgt, c, r]) s1[e, r]) + g_mag[tgt, c, r] softAbs(s0[e, r]) )
//            - sum_r Re(conj(g_agg[j, c, r]) S1[j, r])
__global__ void __launch_bounds__(256) k_lift_aggregate_T(const float2* __restrict__ g_agg, const float* __restrict__ g_mag,
                                                          const float2* __restrict__ s1sum, int64_t s1_stride,
                                                          const float2* __restrict__ sten, const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ nbr, const int32_t* __restrict__ perm,
                                                          float* __restrict__ gx, int64_t N, int Ci, int R) {
    const int C1 = Ci + 1;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * Ci) return;
    const int64_t row = t / Ci;
    const int c = (int)(t - row * Ci);
    // the out-edge sum and the own-row term nearly cancel (gradient of the per-edge differences x_j - x_i): accumulate in
    // double — this kernel is tiny (N x Ci threads), the precision matters more than the FP64 rate
    double acc = 0.0;
    const int p1 = rowptr[row + 1];
    for (int p = rowptr[row]; p < p1; ++p) {
        const int64_t e = perm[p];
        const int64_t i = nbr[p];
        const float2* s = sten + e * R * 2;
        const float2* ga = g_agg + (i * C1 + c) * R;
        const float* gm = g_mag + (i * Ci + c) * R;
        for (int r = 0; r < R; ++r) {
            const float2 s0 = s[2 * r], s1 = s[2 * r + 1];
            const float2 g = ga[r];
            acc += (double)g.x * s1.x + (double)g.y * s1.y + (double)gm[r] * soft_abs_c(s0);
        }
    }
    double own = 0.0;
    for (int r = 0; r < R; ++r) {
        const float2 g = g_agg[(row * C1 + c) * R + r], s = s1sum[row * s1_stride + r];
        own += (double)g.x * s.x + (double)g.y * s.y;
    }
    gx[t] = (float)(acc - own);
}

}  // namespace fcb

using namespace fcb;

extern "C" int fcb_lift_aggregate_f32(const float* x, const float* lift_sten, const int32_t* rowptr_tgt, const int32_t* nbr_tgt,
                                      const int32_t* perm_tgt, float* agg, float* mag, int64_t N, int Ci, int R, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FCB_REQUIRE(x && lift_sten && rowptr_tgt && nbr_tgt && perm_tgt && agg && mag, FCB_E_ARG, "lift_aggregate: null pointer");
    FCB_REQUIRE(N >= 0 && Ci > 0 && R >= 1 && R <= LIFT_MAX_R, FCB_E_UNSUPPORTED, "lift_aggregate: needs 1 <= n_rings <= %d", LIFT_MAX_R);
    const int64_t tot = N * (Ci + 1);
    if (tot == 0) return FCB_OK;
    FCB_LAUNCH("lift_aggregate", st, k_lift_aggregate<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
                                         x, reinterpret_cast<const float2*>(lift_sten), rowptr_tgt, nbr_tgt, perm_tgt,
                                         reinterpret_cast<float2*>(agg), mag, N, Ci, R));
    return FCB_OK;
}

extern "C" int fcb_lift_aggregate_bwd_f32(const float* g_agg, const float* g_mag, const float* agg, const float* lift_sten, const int32_t* rowptr_src,
                                          const int32_t* nbr_src, const int32_t* perm_src, float* gx, int64_t N, int Ci, int R,
                                          void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FCB_REQUIRE(g_agg && g_mag && agg && lift_sten && rowptr_src && nbr_src && perm_src && gx, FCB_E_ARG, "lift_aggregate_bwd: null pointer");
    FCB_REQUIRE(N >= 0 && Ci > 0 && R >= 1 && R <= LIFT_MAX_R, FCB_E_UNSUPPORTED, "lift_aggregate_bwd: needs 1 <= n_rings <= %d", LIFT_MAX_R);
    const int64_t tot = N * Ci;
    if (tot == 0) return FCB_OK;
    FCB_LAUNCH("lift_aggregate_T", st, k_lift_aggregate_T<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
                                           reinterpret_cast<const float2*>(g_agg), g_mag,
                                           reinterpret_cast<const float2*>(agg) + (int64_t)Ci * R, (int64_t)(Ci + 1) * R,
                                           reinterpret_cast<const float2*>(lift_sten),
                                           rowptr_src, nbr_src, perm_src, gx, N, Ci, R));
    return FCB_OK;
}
