// K1 / K5a kernel template (gauge-aligned gather + deterministic segmented reduction) and its dispatcher, shared by the
// four translation units that instantiate it (aggregate.cu, aggregate_t.cu, aggregate_pk.cu, aggregate_pk_t.cu: one per
// (transpose, packed-output) pair, so the variants compile in parallel).
#pragma once
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace fcb {


// p[B+m] = a_m * z * q^m for m = -B..B, by recurrence on the (unit-modulus) per-edge, per-channel rotation q.
// Forward: q = e^{i theta} conj(u) with u = z/|z| (1 at origin entries: utils/field.py:14-16,42-46), so
// p[B+m] = sten-factor a_{e,m} * xhat[src,c,m] (nn/field_conv.py:128-130 folded with fc_precomp.py:83-95);
// transposed: q = e^{-i theta}, p[B+m] = conj(a_{e,m}) * gy.
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// Frequencies 0 and +1 by complex products, the others by the three-term recurrence of a unit-modulus rotation,
// z q^(m+1) = 2 Re(q) z q^m - z q^(m-1)  (and z conj(q) = 2 Re(q) z - z q exactly): one FFMA2 per new frequency.
// Deviation from the product form: <= 4e-7 normwise at |m| = 2, 8e-7 at |m| = 3 (|q|^2 = 1 +- 5e-7 enters linearly),
// inside the fp32 reference's own noise.  wxp and q0 arrive decoded for the direction (conjugated for the transposed
// operator, the packed path's operand scale folded into wxp).  (Measured and dropped, r02c: the three complex products as
// FMUL2 + FFMA2 pairs on pre-rotated operands — 3 fewer issue slots but 8 more FMA-pipe cycles per edge, no change in time.)
template <int B, bool TRANSPOSE>
__device__ __forceinline__ void edge_products(float2 z, float2 wxp, float2 q0, float2* p) {
    p[B] = cmul(wxp, z);
    if (B == 0) return;
    float2 q;
    if (!TRANSPOSE) {
        // branch-free: at origin entries (|re|,|im| < 1e-7) the selects discard the inf/NaN of rsqrt(0)
        const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
        const float ri = rsqrt_ftz(fmaf(z.x, z.x, z.y * z.y));
        const float ux = origin ? 1.f : z.x * ri;
        const float uy = origin ? 0.f : z.y * ri;
        q = cmul_conj(q0, make_float2(ux, uy));
    } else {
        q = q0;
    }
    p[B + 1] = cmul(p[B], q);
    const float2 c2 = bc2(2.f * q.x);
    p[B - 1] = __ffma2_rn(c2, p[B], neg2(p[B + 1]));
#pragma unroll
    for (int m = 2; m <= B; ++m) {
        p[B + m] = __ffma2_rn(c2, p[B + m - 1], neg2(p[B + m - 2]));
        p[B - m] = __ffma2_rn(c2, p[B - m + 1], neg2(p[B - m + 2]));
    }
}

// xh[B+m] = z * conj(u)^m (dense-stencil path)
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

// store the 2 x M complex values a lane holds for one ring: one float4 (channels 2cp, 2cp+1) per m, `m_stride`
// float4 apart.  Forward layout out[row][ring][m][c] (m_stride = C/2), transposed layout out[row][m][ring][o]
// (m_stride = R*C/2): in both, consecutive lanes write consecutive 16-byte pieces -> full-line coalesced stores.
// `mx` follows max|value stored| (the operand scale of the 2xFP16 contraction, gemm_h.cu): FMNMX runs on the ALU pipe,
// off the FMA pipe that bounds these kernels.
template <int M>
__device__ __forceinline__ void store_ring(char* __restrict__ dst, const float2 (&acc)[2][M], uint32_t m_stride_bytes, float& mx) {
#pragma unroll
    for (int m = 0; m < M; ++m) {
        *reinterpret_cast<float4*>(dst + (size_t)(m * m_stride_bytes)) = make_float4(acc[0][m].x, acc[0][m].y, acc[1][m].x, acc[1][m].y);
        mx = fmaxf(fmaxf(mx, fabsf(acc[0][m].x)), fmaxf(fabsf(acc[0][m].y), fmaxf(fabsf(acc[1][m].x), fabsf(acc[1][m].y))));
    }
}

// fold a warp's max|value| into *amax (bit pattern of a non-negative float; max is order-independent, so the result is
// deterministic).  Every lane of the warp must call this.  The plain read first keeps the atomics to the handful of
// warps that actually raise the running maximum.
__device__ __forceinline__ void fold_amax(uint32_t* amax, float mx) {
    const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
    if (amax && (threadIdx.x & 31) == 0 && w > *reinterpret_cast<volatile uint32_t*>(amax)) atomicMax(amax, w);
}

// Packed-operand store (PK format, common.cuh): the 2 x M complex values a lane holds for one ring go out as scaled fp16
// (hi, lo) pairs straight into the swizzled tile image the 2xFP16 GEMMs bulk-copy, so the contraction kernels need no
// producer warps at all.  `row_base` = block of (row tile, chunk 0, hi) + r*128; kk = real column of the m = -B entry,
// kk_m = columns between consecutive m.  A lane's 4 reals (two complex channels) are one 8-byte half of a 16-byte unit:
// two adjacent lanes fill a unit, the lanes of a row cover consecutive 8-byte pieces -> full-sector stores.
// FASTM: the columns of consecutive m are a whole number of 64-column chunks apart (kk_m % 64 == 0: 2*R*C for the
// transposed layout, 2*C for the forward one — cfg 1, cfg 2 transposed, cfg 4), so chunk offset and swizzle are those of
// m = -B and every further m is one pointer increment instead of the full address arithmetic (~10 of ~30 instructions
// per stored (ring, m); the packed store is a quarter of the transposed kernel's instruction stream at cfg 2).
template <int M, bool FASTM>
__device__ __forceinline__ void store_ring_packed(uint8_t* __restrict__ row_base, uint32_t rsw, const float2 (&acc)[2][M],
                                                  uint32_t kk, uint32_t kk_m) {
    uint8_t* p = row_base + (size_t)(kk >> 6) * PK_BLOCK_BYTES + ((((kk >> 3) & 7u) ^ rsw) << 4) + (((kk >> 2) & 1u) << 3);
    const size_t step = (size_t)(kk_m >> 6) * PK_BLOCK_BYTES;
#pragma unroll
    for (int m = 0; m < M; ++m, kk += kk_m) {
        // the accumulators already carry the operand scale (folded into wxp when the plan records are decoded)
        const float a0 = acc[0][m].x, a1 = acc[0][m].y, a2 = acc[1][m].x, a3 = acc[1][m].y;
        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
        if (!FASTM) p = row_base + (size_t)(kk >> 6) * PK_BLOCK_BYTES + ((((kk >> 3) & 7u) ^ rsw) << 4) + (((kk >> 2) & 1u) << 3);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(p + PK_PLANE_BYTES) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        if (FASTM) p += step;
    }
}

constexpr int AGG_CAP = 768;           // plan records staged per CTA and chunk (24 KB of shared memory)

// The two live rings sit in two fixed accumulator sets selected by ring
// parity (ring r lives in acc[r & 1]), so sliding the two-ring window costs one store + one clear, no moves.
// PACK: `out` is a PK buffer (common.cuh) of pk_rows_padded(N) rows x 2*R*M*C columns instead of the fp32 matrix;
// the operand scale comes from the a-priori bound  max|out| <= max|feat| * max_row sum_e |wxp_e|  (pk_feat_amax,
// pk_norm: device floats; block 0 publishes the product in *pk_bound for the GEMM's epilogue), and the lanes of the
// rows N .. pk_rows_padded(N)-1 zero-fill the tail of the last row tile (the weight-gradient GEMM reduces over rows).
// MINB: minimum resident CTAs per SM the register allocation is held to (4 at band limit 1: 64 registers; 3 at band
// limits 0 and 2: 80 registers; 2 beyond).  FASTM: the pointer-increment packed store (store_ring_packed).
//
// Records: the CTA's plan records (one contiguous CSR range, ~11 rows x 40 edges at C = 48) are copied into shared
// memory with coalesced loads first and DECODED once per record instead of once per lane and edge (gather offset
// nbr * P, ring floor, ring weights of the even / odd accumulator set, conjugation of the transposed operator, the
// packed path's operand scale folded into wxp); in the edge loop a record is two LDS.128 through one 32-bit
// shared-window pointer and the gather offset of the next edge one LDS.32, so the feature gather never waits on a
// record.  Ranges longer than AGG_CAP records are staged chunk by chunk (any vertex degree).  The feature row of edge
// p+1 is in flight in registers while edge p is accumulated.
// Measured on B200, cfg-2 layer forward / transposed: round-1 kernel (records through registers, product arithmetic)
// 0.473 / 0.412 ms -> staged + decoded + recurrence arithmetic 0.386 / 0.319 ms (fp32 G; 0.371 with PK G)
// (profiles/r02a_, r02b_aggregate_variants.jsonl) -> overhead instructions trimmed (100 -> 87 per edge and lane forward,
// 71.5 -> 58 transposed; profiles/r04d, r04e) 0.358 / 0.326 ms (PK G).  Deeper feature prefetch (two edges ahead in
// registers, a four-deep cp.async ring in shared memory, an L1 prefetch hint) measured 1-5 % slower each time.  ncu (r04e):
// FMA pipe 60 %, issue-active 63-67 %, largest stall long_scoreboard (the gather).
template <int B, bool TRANSPOSE, bool PACK, int MINB, bool FASTM = false>
__global__ void __launch_bounds__(256, MINB) k_aggregate(const float4* __restrict__ feat, const int32_t* __restrict__ rowptr,
                                                      const int4* __restrict__ rec, const float2* __restrict__ rot,
                                                      float4* __restrict__ out, int64_t N, int C, int R,
                                                      uint32_t* __restrict__ amax, const float* __restrict__ pk_feat_amax,
                                                      const float* __restrict__ pk_norm, float* __restrict__ pk_bound) {
    constexpr int M = 2 * B + 1;
    const int P = C >> 1;
    const int64_t lane0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t lane_id = lane0 + threadIdx.x;
    const int64_t row = lane_id / P;
    const int cp = (int)(lane_id - row * P);
    const bool valid = row < N;
    float mx = 0.f;
    // decoded records, two 16-byte halves AGG_CAP entries apart (one 32-bit shared address walks both):
    //   [i]           {nbr * P (float4 index of the neighbour's row), weight of the even ring set, weight of the odd ring set, ring floor f}
    //   [AGG_CAP + i] {wxp, q0}: conjugated for the transposed operator, operand scale folded into wxp
    __shared__ int4 s_rec[2 * AGG_CAP];
    int4* const s_a = s_rec;
    float4* const s_b = reinterpret_cast<float4*>(s_rec + AGG_CAP);

    // PK addressing of this lane: first byte of its row inside (row tile, chunk 0, hi plane); real column of (ring 0, m = -B)
    uint8_t* pk_row = nullptr;
    uint32_t pk_rsw = 0, pk_kk = 0, pk_kk_ring = 0, pk_kk_m = 0;
    float pk_s = 1.f;
    if (PACK) {
        // pk_feat_amax is the largest REAL component; a complex modulus can be sqrt(2) larger
        const float bound = __ldg(pk_feat_amax) * __ldg(pk_norm) * 1.41422f;
        pk_s = __uint_as_float(scale_field(__float_as_uint(fabsf(bound))) << 23);
        if (lane_id == 0) *pk_bound = fabsf(bound);
        const uint32_t nchunks = (uint32_t)(2 * R * M * C) >> 6;
        pk_row = reinterpret_cast<uint8_t*>(out) + (size_t)(row >> 7) * nchunks * PK_BLOCK_BYTES + (size_t)(row & 127) * 128u;
        pk_rsw = (uint32_t)(row & 7);
        pk_kk = 4u * (uint32_t)cp;
        pk_kk_ring = TRANSPOSE ? 2u * (uint32_t)C : 2u * (uint32_t)(M * C);
        pk_kk_m = TRANSPOSE ? 2u * (uint32_t)(R * C) : 2u * (uint32_t)C;
        if (!valid && row < ((N + 127) & ~(int64_t)127)) {      // tail of the last row tile: zeros
            float2 z[2][M];
#pragma unroll
            for (int m = 0; m < M; ++m) z[0][m] = z[1][m] = make_float2(0.f, 0.f);
            for (int ring = 0; ring < R; ++ring) store_ring_packed<M, FASTM>(pk_row, pk_rsw, z, pk_kk + ring * pk_kk_ring, pk_kk_m);
        }
    }

    float2 acc0[2][M], acc1[2][M];   // even rings / odd rings
#pragma unroll
    for (int m = 0; m < M; ++m) {
        acc0[0][m] = acc0[1][m] = make_float2(0.f, 0.f);
        acc1[0][m] = acc1[1][m] = make_float2(0.f, 0.f);
    }
    // where ring 0 of this lane goes, and how far apart rings / frequencies are (bytes)
    char* dst = reinterpret_cast<char*>(out + row * ((int64_t)R * C * M / 2) + cp);
    const uint32_t ring_stride = 16u * (uint32_t)(TRANSPOSE ? P : P * M);
    const uint32_t m_stride = 16u * (uint32_t)(TRANSPOSE ? R * P : P);
    const float4* fbase = feat + cp;

    // ring fcur is complete: write it once and clear its accumulator set (it becomes ring fcur + 2)
    auto retire = [&](int ring) {
        if (ring & 1) {
            if (PACK) store_ring_packed<M, FASTM>(pk_row, pk_rsw, acc1, pk_kk, pk_kk_m);
            else store_ring<M>(dst, acc1, m_stride, mx);
#pragma unroll
            for (int m = 0; m < M; ++m) acc1[0][m] = acc1[1][m] = make_float2(0.f, 0.f);
        } else {
            if (PACK) store_ring_packed<M, FASTM>(pk_row, pk_rsw, acc0, pk_kk, pk_kk_m);
            else store_ring<M>(dst, acc0, m_stride, mx);
#pragma unroll
            for (int m = 0; m < M; ++m) acc0[0][m] = acc0[1][m] = make_float2(0.f, 0.f);
        }
        dst += ring_stride;
        pk_kk += pk_kk_ring;
    };

    int fcur = 0;
    const int p0 = valid ? rowptr[row] : 0, p1 = valid ? rowptr[row + 1] : 0;
    // CSR range of all rows this CTA touches (CTA-uniform)
    const int64_t row_first = lane0 / P;
    int e0 = 0, e1 = 0;
    if (row_first < N) {
        const int64_t row_last = min(N - 1, (lane0 + blockDim.x - 1) / P);
        e0 = rowptr[row_first];
        e1 = rowptr[row_last + 1];
    }
    for (int lo = e0; lo < e1; lo += AGG_CAP) {
        const int hi = min(e1, lo + AGG_CAP);
        if (lo != e0) __syncthreads();              // every lane is done with the previous chunk
        for (int i = threadIdx.x; i < hi - lo; i += blockDim.x) {
            const int4 rc = __ldg(rec + lo + i);
            const float2 rt = __ldg(rot + lo + i);
            const float t = __int_as_float(rc.y);
            const float omt = 1.0f - t;  // fc_precomp.py:25
            const bool odd = ((uint32_t)rc.x >> NBR_BITS) & 1u;
            const float wx = __int_as_float(rc.z) * pk_s, wy = (TRANSPOSE ? -__int_as_float(rc.w) : __int_as_float(rc.w)) * pk_s;
            const float qx = rt.x, qy = TRANSPOSE ? -rt.y : rt.y;
            s_a[i] = make_int4((int)(((uint32_t)rc.x & NBR_MASK) * (uint32_t)P), __float_as_int(odd ? t : omt), __float_as_int(odd ? omt : t),
                               (int)((uint32_t)rc.x >> NBR_BITS));
            s_b[i] = make_float4(wx, wy, qx, qy);
        }
        __syncthreads();
        const int a = max(p0, lo) - lo, b = min(p1, hi) - lo;      // this lane's edges inside the chunk
        if (a < b) {
            const int last = b - 1;
            // explicit shared-window addressing: one register walks the records (the generic form recomputed the window
            // base every iteration: 5 of ~70 instructions per edge)
            uint32_t sp = (uint32_t)__cvta_generic_to_shared(s_rec + a);
            const uint32_t sp_last = sp + 16u * (uint32_t)(last - a);
            auto lds_nbr = [](uint32_t addr) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
                return v;
            };
            float4 vA = __ldg(fbase + lds_nbr(sp));
#pragma unroll 2
            for (; sp <= sp_last; sp += 16u) {
                const float4 v = vA;
                vA = __ldg(fbase + lds_nbr(min(sp + 16u, sp_last)));
                int4 ra;
                float4 rb;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ra.x), "=r"(ra.y), "=r"(ra.z), "=r"(ra.w) : "r"(sp));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                             : "=f"(rb.x), "=f"(rb.y), "=f"(rb.z), "=f"(rb.w)
                             : "r"(sp), "n"(AGG_CAP * 16));
                const int f = ra.w;
                while (fcur < f) retire(fcur++);
                const float2 w00 = bc2(__int_as_float(ra.y)), w11 = bc2(__int_as_float(ra.z));
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const float2 z = ch ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                    float2 pr[M];
                    edge_products<B, TRANSPOSE>(z, make_float2(rb.x, rb.y), make_float2(rb.z, rb.w), pr);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        acc0[ch][m] = __ffma2_rn(w00, pr[m], acc0[ch][m]);
                        acc1[ch][m] = __ffma2_rn(w11, pr[m], acc1[ch][m]);
                    }
                }
            }
        }
    }
    if (valid)
        while (fcur < R) retire(fcur++);
    if (!PACK) fold_amax(amax, mx);
}

template <bool TRANSPOSE, bool PACK>
static int dispatch_aggregate(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out,
                              int64_t N, int C, int B, int R, float* amax, const float* pk_feat_amax, const float* pk_norm,
                              float* pk_bound, cudaStream_t st) {
    uint32_t* am = reinterpret_cast<uint32_t*>(amax);
    const int64_t lanes = (PACK ? pk_rows_padded(N) : N) * (C / 2);
    if (lanes == 0) return FCB_OK;
    const unsigned blocks = (unsigned)((lanes + 255) / 256);
    const float4* f4 = reinterpret_cast<const float4*>(feat);
    const int4* r4 = static_cast<const int4*>(rec);
    const float2* rt = reinterpret_cast<const float2*>(rot);
    float4* o4 = reinterpret_cast<float4*>(out);
    prof_begin(PACK ? (TRANSPOSE ? "aggregate_T_pk" : "aggregate_pk") : (TRANSPOSE ? "aggregate_T" : "aggregate"), st);
#define FCB_AGG_ARGS <<<blocks, 256, 0, st>>>(f4, rowptr, r4, rt, o4, N, C, R, am, pk_feat_amax, pk_norm, pk_bound)
    // resident CTAs per SM: 4 (64 registers) at band limit 1 — r04f: 1-2 % faster than 3 once the edge loop was trimmed —
    // 3 (<= 85 registers) at band limits 0 and 2, 2 beyond (the accumulators alone take 56+ registers)
    // packed output whose per-m column step is a whole number of chunks: the pointer-increment store (store_ring_packed)
    const bool fastm = PACK && ((TRANSPOSE ? 2 * R * C : 2 * C) % (int)PK_COLS) == 0 && B >= 1 && B <= 2;
    if (fastm) {
        if (B == 1) k_aggregate<1, TRANSPOSE, PACK, 4, PACK> FCB_AGG_ARGS;
        else k_aggregate<2, TRANSPOSE, PACK, 3, PACK> FCB_AGG_ARGS;
    } else
    switch (B) {
        case 0: k_aggregate<0, TRANSPOSE, PACK, 3> FCB_AGG_ARGS; break;
        case 1: k_aggregate<1, TRANSPOSE, PACK, 4> FCB_AGG_ARGS; break;
        case 2: k_aggregate<2, TRANSPOSE, PACK, 3> FCB_AGG_ARGS; break;
        case 3: k_aggregate<3, TRANSPOSE, PACK, 2> FCB_AGG_ARGS; break;
        case 4: k_aggregate<4, TRANSPOSE, PACK, 2> FCB_AGG_ARGS; break;
        default: set_error("aggregate: band_limit %d unsupported", B); return FCB_E_UNSUPPORTED;
    }
#undef FCB_AGG_ARGS
    prof_end(st);
    FCB_CUDA_LAUNCH_CHECK("aggregate");
    return FCB_OK;
}

static inline int check_aggregate(const float* feat, const void* rec, const float* out, int64_t N, int C, int B, int R) {
    FCB_REQUIRE(N >= 0 && C > 0 && R >= 2 && R <= FCB_MAX_RINGS, FCB_E_ARG, "aggregate: bad sizes");
    FCB_REQUIRE(B >= 0 && B <= FCB_MAX_BAND_LIMIT, FCB_E_UNSUPPORTED, "aggregate: band_limit %d unsupported", B);
    FCB_REQUIRE((C & 1) == 0, FCB_E_ALIGN, "aggregate: channel count must be even (16-byte feature rows)");
    FCB_REQUIRE(aligned16(feat) && aligned16(out) && aligned16(rec), FCB_E_ALIGN, "aggregate: pointers must be 16-byte aligned");
    FCB_REQUIRE((N + 128) * (int64_t)(C / 2) < 0xffffffffLL, FCB_E_UNSUPPORTED, "aggregate: N*C/2 must fit 32 bits");
    return FCB_OK;
}


}  // namespace fcb
