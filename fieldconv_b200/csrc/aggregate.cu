// K1 / K5a — gauge-aligned gather + deterministic segmented reduction over CSR rows.
//
// Forward (transpose = 0), nn/field_conv.py:128-134 + utils/field.py:40-48 of the reference:
//   contrib[i, r, m, c] = sum_{e in row i} sten[e,r,m] * x[src(e),c] * conj(u[src(e),c])^m
// Backward gather (transpose = 1), the adjoint of the same sparse operator applied to gy:
//   G[j, m, r, o]       = sum_{e in row j (by-source)} conj(sten[e,r,m]) * gy[tgt(e),o]
// with sten[e,r,m] = w_r(e) * wxp_e * exp(i m theta_e), only rings f and f+1 non-zero
// (transforms/fc_precomp.py:10-27,83-95).  Rows are sorted by ring floor f, so a lane keeps just
// the two live rings in registers and writes every ring exactly once, in order: no atomics, a
// fixed summation order, bit-identical results run to run.
//
// Thread mapping: one lane owns (row, channel pair): a 128-bit load fetches two complex
// channels of the neighbour's feature row; consecutive lanes read consecutive 16-byte pieces
// of the same row, so a row of C channels is fetched as C/2 coalesced float4 loads.
#include "aggregate_kernel.cuh"

namespace fcb {

// the other three (transpose, packed) instantiation sets live in their own translation units
int aggregate_transposed_f32(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out, int64_t N,
                             int C, int B, int R, float* amax, cudaStream_t st);
int aggregate_forward_packed(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out, int64_t N,
                             int C, int B, int R, const float* feat_amax, const float* norm, float* bound, cudaStream_t st);
int aggregate_transposed_packed(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out, int64_t N,
                                int C, int B, int R, const float* feat_amax, const float* norm, float* bound, cudaStream_t st);

// Dense-stencil variant: arbitrary supp_sten (E,R,M), one lane per (row, ring, channel pair).
template <int B, bool TRANSPOSE>
__global__ void __launch_bounds__(256) k_aggregate_dense(const float4* __restrict__ feat, const float2* __restrict__ sten,
                                                         const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                                                         const int32_t* __restrict__ perm, float4* __restrict__ out,
                                                         int64_t N, int C, int R, uint32_t* __restrict__ amax) {
    constexpr int M = 2 * B + 1;
    const int P = C >> 1;
    const int64_t lane_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t unit = lane_id / P;  // (row, ring)
    const int64_t row = unit / R;
    float mx = 0.f;
    if (row < N) {
    const int ring = (int)(unit - row * R);
    const int cp = (int)(lane_id - unit * P);
    float2 acc[2][M];
#pragma unroll
    for (int m = 0; m < M; ++m) acc[0][m] = acc[1][m] = make_float2(0.f, 0.f);
    const int p0 = rowptr[row], p1 = rowptr[row + 1];
    for (int p = p0; p < p1; ++p) {
        const int64_t e = perm[p];
        const float4 v = __ldg(feat + (int64_t)nbr[p] * P + cp);
        const float2* s = sten + (e * R + ring) * M;
        float2 a[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            a[m] = __ldg(s + m);
            if (TRANSPOSE) a[m].y = -a[m].y;
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const float2 z = ch ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
            float2 xh[M];
            if (!TRANSPOSE) gauge_align<B>(z, xh);
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float2 pr = cmul(a[m], TRANSPOSE ? z : xh[m]);
                acc[ch][m].x += pr.x;
                acc[ch][m].y += pr.y;
            }
        }
    }
    float4* orow = out + row * ((int64_t)R * C * M / 2);
    store_ring<M>(reinterpret_cast<char*>(orow + (TRANSPOSE ? (int64_t)ring * P : (int64_t)ring * M * P) + cp), acc,
                  16u * (uint32_t)(TRANSPOSE ? R * P : P), mx);
    }
    fold_amax(amax, mx);
}

int launch_aggregate(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out, int64_t N,
                     int C, int B, int R, int transpose, float* amax, cudaStream_t st) {
    const int rc = check_aggregate(feat, rec, out, N, C, B, R);
    if (rc) return rc;
    return transpose ? aggregate_transposed_f32(feat, rowptr, rec, rot, out, N, C, B, R, amax, st)
                     : dispatch_aggregate<false, false>(feat, rowptr, rec, rot, out, N, C, B, R, amax, nullptr, nullptr, nullptr, st);
}

// PK output (see k_aggregate): out_pk holds pk_bytes(N, 2*R*M*C) bytes; *bound receives feat_amax * norm.
int launch_aggregate_packed(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, void* out_pk,
                            int64_t N, int C, int B, int R, int transpose, const float* feat_amax, const float* norm,
                            float* bound, cudaStream_t st) {
    const int rc = check_aggregate(feat, rec, static_cast<const float*>(out_pk), N, C, B, R);
    if (rc) return rc;
    FCB_REQUIRE(feat_amax && norm && bound, FCB_E_ARG, "aggregate_packed: null scale pointers");
    FCB_REQUIRE(((2 * (int64_t)R * (2 * B + 1) * C) % PK_COLS) == 0, FCB_E_UNSUPPORTED, "aggregate_packed: 2*R*M*C must be a multiple of 64");
    FCB_REQUIRE((reinterpret_cast<uintptr_t>(out_pk) & 127u) == 0, FCB_E_ALIGN, "aggregate_packed: output must be 128-byte aligned");
    float* o = static_cast<float*>(out_pk);
    return transpose ? aggregate_transposed_packed(feat, rowptr, rec, rot, o, N, C, B, R, feat_amax, norm, bound, st)
                     : aggregate_forward_packed(feat, rowptr, rec, rot, o, N, C, B, R, feat_amax, norm, bound, st);
}

// max over the CSR rows of sum_e |wxp_e| (the l-infinity operator norm of the aggregation, all rings and frequencies
// included: ring weights are in [0,1], |e^{i m theta}| = 1): with positive vertex weights it is <= 1 for the by-target
// order (fc_precomp.py:87 normalises the row mass) and the column mass for the by-source order.
__global__ void __launch_bounds__(256) k_plan_norm(const int32_t* __restrict__ rowptr, const int4* __restrict__ rec, int64_t N,
                                                   uint32_t* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    if (row < N) {
        const int p1 = rowptr[row + 1];
        for (int p = rowptr[row]; p < p1; ++p) {
            const int4 rc = __ldg(rec + p);
            const float a = __int_as_float(rc.z), b = __int_as_float(rc.w);
            s += sqrtf(a * a + b * b);
        }
        s *= 1.0001f;      // the bound only feeds a power-of-two scale with a factor-2 headroom; keep it on the safe side
    }
    fold_amax(out, s);
}

int launch_plan_norm(const int32_t* rowptr, const void* rec, int64_t N, float* out, cudaStream_t st) {
    if (cudaMemsetAsync(out, 0, 4, st) != cudaSuccess) {
        set_error("plan_norm: cudaMemsetAsync failed");
        return FCB_E_CUDA;
    }
    if (N == 0) return FCB_OK;
    FCB_LAUNCH("plan_norm", st, k_plan_norm<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(rowptr, static_cast<const int4*>(rec), N,
                                                                                      reinterpret_cast<uint32_t*>(out)));
    return FCB_OK;
}

template <bool TRANSPOSE>
static int dispatch_aggregate_dense(const float* feat, const float* sten, const int32_t* rowptr, const int32_t* nbr,
                                    const int32_t* perm, float* out, int64_t N, int C, int B, int R, float* amax,
                                    cudaStream_t st) {
    uint32_t* am = reinterpret_cast<uint32_t*>(amax);
    const int64_t lanes = N * R * (C / 2);
    if (lanes == 0) return FCB_OK;
    const unsigned blocks = (unsigned)((lanes + 255) / 256);
    const float4* f4 = reinterpret_cast<const float4*>(feat);
    const float2* s2 = reinterpret_cast<const float2*>(sten);
    float4* o4 = reinterpret_cast<float4*>(out);
    prof_begin(TRANSPOSE ? "aggregate_dense_T" : "aggregate_dense", st);
    switch (B) {
        case 0: k_aggregate_dense<0, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R, am); break;
        case 1: k_aggregate_dense<1, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R, am); break;
        case 2: k_aggregate_dense<2, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R, am); break;
        case 3: k_aggregate_dense<3, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R, am); break;
        case 4: k_aggregate_dense<4, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R, am); break;
        default: set_error("aggregate_dense: band_limit %d unsupported", B); return FCB_E_UNSUPPORTED;
    }
    prof_end(st);
    FCB_CUDA_LAUNCH_CHECK("aggregate_dense");
    return FCB_OK;
}

int launch_aggregate_dense(const float* feat, const float* sten, const int32_t* rowptr, const int32_t* nbr,
                           const int32_t* perm, float* out, int64_t N, int C, int B, int R, int transpose,
                           float* amax, cudaStream_t st) {
    FCB_REQUIRE(N >= 0 && C > 0 && R >= 1 && R <= FCB_MAX_RINGS, FCB_E_ARG, "aggregate_dense: bad sizes");
    FCB_REQUIRE(B >= 0 && B <= FCB_MAX_BAND_LIMIT, FCB_E_UNSUPPORTED, "aggregate_dense: band_limit %d unsupported", B);
    FCB_REQUIRE((C & 1) == 0, FCB_E_ALIGN, "aggregate_dense: channel count must be even");
    FCB_REQUIRE(aligned16(feat) && aligned16(out), FCB_E_ALIGN, "aggregate_dense: pointers must be 16-byte aligned");
    FCB_REQUIRE(N * (int64_t)R * (C / 2) / 256 < 0x7fffffffLL, FCB_E_UNSUPPORTED, "aggregate_dense: grid too large");
    return transpose ? dispatch_aggregate_dense<true>(feat, sten, rowptr, nbr, perm, out, N, C, B, R, amax, st)
                     : dispatch_aggregate_dense<false>(feat, sten, rowptr, nbr, perm, out, N, C, B, R, amax, st);
}

}  // namespace fcb

extern "C" int fcb_plan_norm(const int32_t* rowptr, const void* rec, int64_t N, float* out, void* stream) {
    FCB_REQUIRE(rowptr && rec && out && N >= 0, FCB_E_ARG, "plan_norm: bad arguments");
    return fcb::launch_plan_norm(rowptr, rec, N, out, static_cast<cudaStream_t>(stream));
}

extern "C" int fcb_aggregate_f32(const float* feat, const int32_t* rowptr, const void* rec, const float* rot,
                                 float* out, int64_t N, int C, int band_limit, int R, int transpose, void* stream) {
    FCB_REQUIRE(feat && rowptr && rec && rot && out, FCB_E_ARG, "aggregate: null pointer");
    return fcb::launch_aggregate(feat, rowptr, rec, rot, out, N, C, band_limit, R, transpose, nullptr,
                                 static_cast<cudaStream_t>(stream));
}
