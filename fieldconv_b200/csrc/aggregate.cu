// K1 / K5a — gauge-aligned gather + deterministic segmented reduction over CSR rows.
//
// Forward (transpose = 0), nn/field_conv.py:128-134 + utils/field.py:40-48 of the reference:
//   contrib[i, r, c, m] = sum_{e in row i} sten[e,r,m] * x[src(e),c] * conj(u[src(e),c])^m
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
#include "common.cuh"

namespace fcb {

template <int B>
struct Coef {
    static constexpr int M = 2 * B + 1;
    float2 a[M];
    // a[B+m] = wxp * exp(i m theta), built by recurrence from rot = exp(i theta)
    __device__ __forceinline__ void build(float2 wxp, float2 rot, bool conj_all) {
        a[B] = wxp;
#pragma unroll
        for (int m = 1; m <= B; ++m) {
            a[B + m] = cmul(a[B + m - 1], rot);
            a[B - m] = cmul_conj(a[B - m + 1], rot);
        }
        if (conj_all) {
#pragma unroll
            for (int m = 0; m < M; ++m) a[m].y = -a[m].y;
        }
    }
};

// xh[B+m] = z * conj(u)^m, u = z/|z| (1 at origin entries: utils/field.py:14-16,42-46)
template <int B>
__device__ __forceinline__ void gauge_align(float2 z, float2* xh) {
    const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
    const float n2 = z.x * z.x + z.y * z.y;
    const float ri = rsqrtf(n2);
    const float2 u = origin ? make_float2(1.f, 0.f) : make_float2(z.x * ri, z.y * ri);
    xh[B] = z;
#pragma unroll
    for (int m = 1; m <= B; ++m) {
        xh[B + m] = cmul_conj(xh[B + m - 1], u);
        xh[B - m] = cmul(xh[B - m + 1], u);
    }
}

// store the 2 x M complex values a lane holds for one ring
template <int M, bool TRANSPOSE>
__device__ __forceinline__ void store_ring(float4* __restrict__ orow, const float2 (&acc)[2][M], int ring, int cp, int C,
                                           int R) {
    if (!TRANSPOSE) {
        // out[row][ring][c][m]: channels 2cp, 2cp+1 -> 2M consecutive complex = M float4
        float4* dst = orow + ((int64_t)ring * C + 2 * cp) * M / 2;
        float tmp[4 * M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            tmp[2 * m] = acc[0][m].x;
            tmp[2 * m + 1] = acc[0][m].y;
            tmp[2 * M + 2 * m] = acc[1][m].x;
            tmp[2 * M + 2 * m + 1] = acc[1][m].y;
        }
#pragma unroll
        for (int q = 0; q < M; ++q) dst[q] = make_float4(tmp[4 * q], tmp[4 * q + 1], tmp[4 * q + 2], tmp[4 * q + 3]);
    } else {
        // out[row][m][ring][o]: one float4 (channels 2cp, 2cp+1) per m
#pragma unroll
        for (int m = 0; m < M; ++m)
            orow[((int64_t)(m * R + ring) * C) / 2 + cp] = make_float4(acc[0][m].x, acc[0][m].y, acc[1][m].x, acc[1][m].y);
    }
}

template <int B, bool TRANSPOSE>
__global__ void __launch_bounds__(256) k_aggregate(const float4* __restrict__ feat, const int32_t* __restrict__ rowptr,
                                                   const int4* __restrict__ rec, const float2* __restrict__ rot,
                                                   float4* __restrict__ out, int64_t N, int C, int R) {
    constexpr int M = 2 * B + 1;
    const int P = C >> 1;
    const int64_t lane_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = lane_id / P;
    if (row >= N) return;
    const int cp = (int)(lane_id - row * P);

    float2 accF[2][M], accC[2][M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        accF[0][m] = accF[1][m] = make_float2(0.f, 0.f);
        accC[0][m] = accC[1][m] = make_float2(0.f, 0.f);
    }
    float4* orow = out + row * ((int64_t)R * C * M / 2);

    int fcur = 0;
    const int p0 = rowptr[row], p1 = rowptr[row + 1];
    for (int p = p0; p < p1; ++p) {
        const int4 rc = __ldg(rec + p);
        const float2 rt = __ldg(rot + p);
        const int f = (int)((uint32_t)rc.x >> NBR_BITS);
        const int64_t nbr = (int64_t)((uint32_t)rc.x & NBR_MASK);
        while (fcur < f) {  // ring fcur is complete: write it once, slide the two-ring window
            store_ring<M, TRANSPOSE>(orow, accF, fcur, cp, C, R);
#pragma unroll
            for (int m = 0; m < M; ++m) {
                accF[0][m] = accC[0][m];
                accF[1][m] = accC[1][m];
                accC[0][m] = accC[1][m] = make_float2(0.f, 0.f);
            }
            ++fcur;
        }
        const float4 v = __ldg(feat + nbr * P + cp);
        const float t = __int_as_float(rc.y);
        const float omt = 1.0f - t;  // fc_precomp.py:25
        Coef<B> cf;
        cf.build(make_float2(__int_as_float(rc.z), __int_as_float(rc.w)), rt, TRANSPOSE);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const float2 z = ch ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
            float2 xh[M];
            if (!TRANSPOSE) gauge_align<B>(z, xh);
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float2 pr = cmul(cf.a[m], TRANSPOSE ? z : xh[m]);
                accF[ch][m].x = fmaf(omt, pr.x, accF[ch][m].x);
                accF[ch][m].y = fmaf(omt, pr.y, accF[ch][m].y);
                accC[ch][m].x = fmaf(t, pr.x, accC[ch][m].x);
                accC[ch][m].y = fmaf(t, pr.y, accC[ch][m].y);
            }
        }
    }
    while (fcur < R - 1) {
        store_ring<M, TRANSPOSE>(orow, accF, fcur, cp, C, R);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            accF[0][m] = accC[0][m];
            accF[1][m] = accC[1][m];
            accC[0][m] = accC[1][m] = make_float2(0.f, 0.f);
        }
        ++fcur;
    }
    store_ring<M, TRANSPOSE>(orow, accF, R - 1, cp, C, R);
}

// Dense-stencil variant: arbitrary supp_sten (E,R,M), one lane per (row, ring, channel pair).
template <int B, bool TRANSPOSE>
__global__ void __launch_bounds__(256) k_aggregate_dense(const float4* __restrict__ feat, const float2* __restrict__ sten,
                                                         const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                                                         const int32_t* __restrict__ perm, float4* __restrict__ out,
                                                         int64_t N, int C, int R) {
    constexpr int M = 2 * B + 1;
    const int P = C >> 1;
    const int64_t lane_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t unit = lane_id / P;  // (row, ring)
    const int64_t row = unit / R;
    if (row >= N) return;
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
    store_ring<M, TRANSPOSE>(orow, acc, ring, cp, C, R);
}

template <bool TRANSPOSE>
static int dispatch_aggregate(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out,
                              int64_t N, int C, int B, int R, cudaStream_t st) {
    const int64_t lanes = N * (C / 2);
    if (lanes == 0) return FCB_OK;
    const unsigned blocks = (unsigned)((lanes + 255) / 256);
    const float4* f4 = reinterpret_cast<const float4*>(feat);
    const int4* r4 = static_cast<const int4*>(rec);
    const float2* rt = reinterpret_cast<const float2*>(rot);
    float4* o4 = reinterpret_cast<float4*>(out);
    prof_begin(TRANSPOSE ? "aggregate_T" : "aggregate", st);
    switch (B) {
        case 0: k_aggregate<0, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, rowptr, r4, rt, o4, N, C, R); break;
        case 1: k_aggregate<1, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, rowptr, r4, rt, o4, N, C, R); break;
        case 2: k_aggregate<2, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, rowptr, r4, rt, o4, N, C, R); break;
        case 3: k_aggregate<3, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, rowptr, r4, rt, o4, N, C, R); break;
        case 4: k_aggregate<4, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, rowptr, r4, rt, o4, N, C, R); break;
        default: set_error("aggregate: band_limit %d unsupported", B); return FCB_E_UNSUPPORTED;
    }
    prof_end(st);
    FCB_CUDA_LAUNCH_CHECK("aggregate");
    return FCB_OK;
}

int launch_aggregate(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out, int64_t N,
                     int C, int B, int R, int transpose, cudaStream_t st) {
    FCB_REQUIRE(N >= 0 && C > 0 && R >= 2 && R <= FCB_MAX_RINGS, FCB_E_ARG, "aggregate: bad sizes");
    FCB_REQUIRE(B >= 0 && B <= FCB_MAX_BAND_LIMIT, FCB_E_UNSUPPORTED, "aggregate: band_limit %d unsupported", B);
    FCB_REQUIRE((C & 1) == 0, FCB_E_ALIGN, "aggregate: channel count must be even (16-byte feature rows)");
    FCB_REQUIRE(aligned16(feat) && aligned16(out) && aligned16(rec), FCB_E_ALIGN, "aggregate: pointers must be 16-byte aligned");
    FCB_REQUIRE(N * (int64_t)(C / 2) / 256 < 0x7fffffffLL, FCB_E_UNSUPPORTED, "aggregate: grid too large");
    return transpose ? dispatch_aggregate<true>(feat, rowptr, rec, rot, out, N, C, B, R, st)
                     : dispatch_aggregate<false>(feat, rowptr, rec, rot, out, N, C, B, R, st);
}

template <bool TRANSPOSE>
static int dispatch_aggregate_dense(const float* feat, const float* sten, const int32_t* rowptr, const int32_t* nbr,
                                    const int32_t* perm, float* out, int64_t N, int C, int B, int R, cudaStream_t st) {
    const int64_t lanes = N * R * (C / 2);
    if (lanes == 0) return FCB_OK;
    const unsigned blocks = (unsigned)((lanes + 255) / 256);
    const float4* f4 = reinterpret_cast<const float4*>(feat);
    const float2* s2 = reinterpret_cast<const float2*>(sten);
    float4* o4 = reinterpret_cast<float4*>(out);
    prof_begin(TRANSPOSE ? "aggregate_dense_T" : "aggregate_dense", st);
    switch (B) {
        case 0: k_aggregate_dense<0, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R); break;
        case 1: k_aggregate_dense<1, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R); break;
        case 2: k_aggregate_dense<2, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R); break;
        case 3: k_aggregate_dense<3, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R); break;
        case 4: k_aggregate_dense<4, TRANSPOSE><<<blocks, 256, 0, st>>>(f4, s2, rowptr, nbr, perm, o4, N, C, R); break;
        default: set_error("aggregate_dense: band_limit %d unsupported", B); return FCB_E_UNSUPPORTED;
    }
    prof_end(st);
    FCB_CUDA_LAUNCH_CHECK("aggregate_dense");
    return FCB_OK;
}

int launch_aggregate_dense(const float* feat, const float* sten, const int32_t* rowptr, const int32_t* nbr,
                           const int32_t* perm, float* out, int64_t N, int C, int B, int R, int transpose,
                           cudaStream_t st) {
    FCB_REQUIRE(N >= 0 && C > 0 && R >= 1 && R <= FCB_MAX_RINGS, FCB_E_ARG, "aggregate_dense: bad sizes");
    FCB_REQUIRE(B >= 0 && B <= FCB_MAX_BAND_LIMIT, FCB_E_UNSUPPORTED, "aggregate_dense: band_limit %d unsupported", B);
    FCB_REQUIRE((C & 1) == 0, FCB_E_ALIGN, "aggregate_dense: channel count must be even");
    FCB_REQUIRE(aligned16(feat) && aligned16(out), FCB_E_ALIGN, "aggregate_dense: pointers must be 16-byte aligned");
    FCB_REQUIRE(N * (int64_t)R * (C / 2) / 256 < 0x7fffffffLL, FCB_E_UNSUPPORTED, "aggregate_dense: grid too large");
    return transpose ? dispatch_aggregate_dense<true>(feat, sten, rowptr, nbr, perm, out, N, C, B, R, st)
                     : dispatch_aggregate_dense<false>(feat, sten, rowptr, nbr, perm, out, N, C, B, R, st);
}

}  // namespace fcb

extern "C" int fcb_aggregate_f32(const float* feat, const int32_t* rowptr, const void* rec, const float* rot,
                                 float* out, int64_t N, int C, int band_limit, int R, int transpose, void* stream) {
    FCB_REQUIRE(feat && rowptr && rec && rot && out, FCB_E_ARG, "aggregate: null pointer");
    return fcb::launch_aggregate(feat, rowptr, rec, rot, out, N, C, band_limit, R, transpose,
                                 static_cast<cudaStream_t>(stream));
}
