// K1 / K5a kernel template (gauge-aligned gather + deterministic segmented reduction) and its dispatcher, shared by the
// four translation units that instantiate it (aggregate.cu, aggregate_t.cu, aggregate_pk.cu, aggregate_pk_t.cu: one per
// (transpose, packed-output) pair, so the variants compile in parallel).
#pragma once
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace fcb {


// p[B+m] = conj?(wxp) * z * q^m for m = -B..B, by recurrence on the (unit-modulus) per-edge, per-channel
// rotation q.  Forward: q = e^{i theta} conj(u) with u = z/|z| (1 at origin entries: utils/field.py:14-16,42-46),
// so p[B+m] = sten-factor a_{e,m} * xhat[src,c,m] (nn/field_conv.py:128-130 folded with fc_precomp.py:83-95);
// transposed: q = e^{-i theta}, p[B+m] = conj(a_{e,m}) * gy.  2 + 2B complex products per (edge, channel).
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// FAST: frequencies beyond +1 by the three-term recurrence of a unit-modulus rotation,  z q^(m+1) = 2 Re(q) z q^m - z q^(m-1)
// (and z conj(q) = 2 Re(q) z - z q exactly), one packed FFMA2 per new frequency instead of a 4-instruction complex
// product: 2 + 1 + (2B - 1)/... complex products become 3 products + (2B - 1) FFMA2.  Deviation from the product form:
// <= 4e-7 normwise at |m| = 2, 8e-7 at |m| = 3 (|q|^2 = 1 +- 5e-7 enters linearly), inside the fp32 reference's own noise.
template <int B, bool TRANSPOSE, bool FAST = false>
__device__ __forceinline__ void edge_products(float2 z, float2 wxp, float2 rot, float2* p) {
    float2 q;
    if (!TRANSPOSE) {
        // branch-free: at origin entries (|re|,|im| < 1e-7) the selects discard the inf/NaN of rsqrt(0)
        const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
        const float ri = rsqrt_ftz(fmaf(z.x, z.x, z.y * z.y));
        const float ux = origin ? 1.f : z.x * ri;
        const float uy = origin ? 0.f : z.y * ri;
        q = cmul_conj(rot, make_float2(ux, uy));
        p[B] = cmul(wxp, z);
    } else {
        q = make_float2(rot.x, -rot.y);
        p[B] = cmul_conj(z, wxp);
    }
    if (FAST && B >= 1) {
        const float2 c2 = make_float2(2.f * q.x, 2.f * q.x);
        p[B + 1] = cmul(p[B], q);
        p[B - 1] = __ffma2_rn(c2, p[B], make_float2(-p[B + 1].x, -p[B + 1].y));
#pragma unroll
        for (int m = 2; m <= B; ++m) {
            p[B + m] = __ffma2_rn(c2, p[B + m - 1], make_float2(-p[B + m - 2].x, -p[B + m - 2].y));
            p[B - m] = __ffma2_rn(c2, p[B - m + 1], make_float2(-p[B - m + 2].x, -p[B - m + 2].y));
        }
        return;
    }
#pragma unroll
    for (int m = 1; m <= B; ++m) {
        p[B + m] = cmul(p[B + m - 1], q);
        p[B - m] = cmul_conj(p[B - m + 1], q);
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
__device__ __forceinline__ void store_ring(float4* __restrict__ dst, const float2 (&acc)[2][M], int64_t m_stride, float& mx) {
#pragma unroll
    for (int m = 0; m < M; ++m) {
        dst[m * m_stride] = make_float4(acc[0][m].x, acc[0][m].y, acc[1][m].x, acc[1][m].y);
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
template <int M>
__device__ __forceinline__ void store_ring_packed(uint8_t* __restrict__ row_base, uint32_t rsw, const float2 (&acc)[2][M],
                                                  uint32_t kk, uint32_t kk_m, float s) {
#pragma unroll
    for (int m = 0; m < M; ++m, kk += kk_m) {
        const float a0 = acc[0][m].x * s, a1 = acc[0][m].y * s, a2 = acc[1][m].x * s, a3 = acc[1][m].y * s;
        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
        uint8_t* p = row_base + (size_t)(kk >> 6) * PK_BLOCK_BYTES + ((((kk >> 3) & 7u) ^ rsw) << 4) + (((kk >> 2) & 1u) << 3);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(p + PK_PLANE_BYTES) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
}

// Out-of-line form of the packed ring store for the NI kernel variants: the fp16 split needs a dozen temporaries that are
// live only at a ring transition (6 times per row), but inlined they inflate the register allocation of the whole edge
// loop and keep the packed band_limit-2 kernel at 2 CTAs/SM.  The accumulators travel by value (registers, not memory).
template <int M>
struct RingAcc {
    float2 v[2][M];
};
template <int M>
__device__ __noinline__ void store_ring_packed_ni(uint8_t* row_base, uint32_t rsw, RingAcc<M> a, uint32_t kk, uint32_t kk_m, float s) {
    store_ring_packed<M>(row_base, rsw, a.v, kk, kk_m, s);
}

constexpr int AGG_STAGE_CAP = 768;      // plan records a CTA of the DEPTH-4 variant can stage (18 KB of shared memory)

// The two live rings sit in two fixed accumulator sets selected by ring
// parity (ring r lives in acc[r & 1]), so sliding the two-ring window costs one store + one clear, no moves.
// PACK: `out` is a PK buffer (common.cuh) of pk_rows_padded(N) rows x 2*R*M*C columns instead of the fp32 matrix;
// the operand scale comes from the a-priori bound  max|out| <= max|feat| * max_row sum_e |wxp_e|  (pk_feat_amax,
// pk_norm: device floats; block 0 publishes the product in *pk_bound for the GEMM's epilogue), and the lanes of the
// rows N .. pk_rows_padded(N)-1 zero-fill the tail of the last row tile (the weight-gradient GEMM reduces over rows).
// MINB: minimum resident CTAs per SM the register allocation is held to (2: up to 128 registers; 3: 85 — more warps to
// hide the gather latency at the price of a tighter register budget; chosen per band limit by agg_min_blocks()).
// DEPTH: software-pipeline depth of the edge loop (2: the record of edge p+2 and the feature row of edge p+1 are in flight;
// 1: both of edge p+1 only — 6 registers fewer, for the register-capped high-occupancy variants; 3: as 1 plus the
// neighbour id of edge p+2, see the loop; 4: as 1 with the CTA's records staged in shared memory — experiment variants,
// not measured yet).
// FAST: three-term frequency recurrence (edge_products) and packed FFMA2 ring accumulation — 25 % fewer instructions in
// the edge loop; experiment variant (FIELDCONV_B200_AGG_VARIANT codes >= 100), not a default until measured on B200.
// NI (packed output only): ring stores through the out-of-line store_ring_packed_ni — experiment variant (codes 2xx).
template <int B, bool TRANSPOSE, bool PACK, int MINB, int DEPTH, bool FAST, bool NI = false>
__global__ void __launch_bounds__(256, MINB) k_aggregate(const float4* __restrict__ feat, const int32_t* __restrict__ rowptr,
                                                      const int4* __restrict__ rec, const float2* __restrict__ rot,
                                                      float4* __restrict__ out, int64_t N, int C, int R,
                                                      uint32_t* __restrict__ amax, const float* __restrict__ pk_feat_amax,
                                                      const float* __restrict__ pk_norm, float* __restrict__ pk_bound) {
    constexpr int M = 2 * B + 1;
    const int P = C >> 1;
    const int64_t lane_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = lane_id / P;
    float mx = 0.f;
    // DEPTH 4: the plan records of all rows this CTA touches (a contiguous CSR range, ~11 rows x 40 edges x 24 B) are
    // staged in shared memory with coalesced loads before the edge loops start, so inside the loops a record is an LDS
    // and the feature gather of edge p+1 never waits for a global record load.  CTAs whose range exceeds the buffer
    // (CTA-uniform test) read the records from global memory as the other depths do.
    constexpr int STAGE_CAP = DEPTH == 4 ? AGG_STAGE_CAP : 1;
    __shared__ int4 s_rec[STAGE_CAP];
    __shared__ float2 s_rot[STAGE_CAP];
    int stage_e0 = 0;
    bool staged = false;
    if (DEPTH == 4) {
        const int64_t lane0 = (int64_t)blockIdx.x * blockDim.x;
        const int64_t row_first = lane0 / P;
        if (row_first < N) {
            const int64_t row_last = min(N - 1, (lane0 + blockDim.x - 1) / P);
            const int e0 = rowptr[row_first], e1 = rowptr[row_last + 1];
            staged = (e1 - e0) <= STAGE_CAP;
            stage_e0 = e0;
            if (staged) {
                for (int i = threadIdx.x; i < e1 - e0; i += blockDim.x) {
                    s_rec[i] = __ldg(rec + e0 + i);
                    s_rot[i] = __ldg(rot + e0 + i);
                }
            }
        }
        __syncthreads();
    }
    // PK addressing of this lane: first byte of its row inside (row tile, chunk 0, hi plane); real column of (ring 0, m = -B)
    uint8_t* pk_row = nullptr;
    uint32_t pk_rsw = 0, pk_kk = 0, pk_kk_ring = 0, pk_kk_m = 0;
    float pk_s = 0.f;
    if (PACK) {
        // pk_feat_amax is the largest REAL component; a complex modulus can be sqrt(2) larger
        const float bound = __ldg(pk_feat_amax) * __ldg(pk_norm) * 1.41422f;
        pk_s = __uint_as_float(scale_field(__float_as_uint(fabsf(bound))) << 23);
        if (lane_id == 0) *pk_bound = fabsf(bound);
        const int cp = (int)(lane_id - row * P);
        const uint32_t nchunks = (uint32_t)(2 * R * M * C) >> 6;
        pk_row = reinterpret_cast<uint8_t*>(out) + (size_t)(row >> 7) * nchunks * PK_BLOCK_BYTES + (size_t)(row & 127) * 128u;
        pk_rsw = (uint32_t)(row & 7);
        pk_kk = 4u * (uint32_t)cp;
        pk_kk_ring = TRANSPOSE ? 2u * (uint32_t)C : 2u * (uint32_t)(M * C);
        pk_kk_m = TRANSPOSE ? 2u * (uint32_t)(R * C) : 2u * (uint32_t)C;
        if (row >= N) {
            if (row < ((N + 127) & ~(int64_t)127)) {      // tail of the last row tile: zeros
                float2 z[2][M];
#pragma unroll
                for (int m = 0; m < M; ++m) z[0][m] = z[1][m] = make_float2(0.f, 0.f);
                for (int ring = 0; ring < R; ++ring) store_ring_packed<M>(pk_row, pk_rsw, z, pk_kk + ring * pk_kk_ring, pk_kk_m, 0.f);
            }
            return;
        }
    }
    if (row < N) {
    const int cp = (int)(lane_id - row * P);

    float2 acc0[2][M], acc1[2][M];   // even rings / odd rings
#pragma unroll
    for (int m = 0; m < M; ++m) {
        acc0[0][m] = acc0[1][m] = make_float2(0.f, 0.f);
        acc1[0][m] = acc1[1][m] = make_float2(0.f, 0.f);
    }
    // where ring 0 of this lane goes, and how far apart rings are (float4 units)
    float4* dst = out + row * ((int64_t)R * C * M / 2) + cp;
    const int ring_stride = TRANSPOSE ? P : P * M;
    const int64_t m_stride = TRANSPOSE ? (int64_t)R * P : (int64_t)P;
    const float4* fbase = feat + cp;

    // ring fcur is complete: write it once and clear its accumulator set (it becomes ring fcur + 2)
    auto retire = [&](int ring) {
        if (ring & 1) {
            if (PACK && NI) {
                RingAcc<M> a;
#pragma unroll
                for (int m = 0; m < M; ++m) { a.v[0][m] = acc1[0][m]; a.v[1][m] = acc1[1][m]; }
                store_ring_packed_ni<M>(pk_row, pk_rsw, a, pk_kk, pk_kk_m, pk_s);
            } else if (PACK) store_ring_packed<M>(pk_row, pk_rsw, acc1, pk_kk, pk_kk_m, pk_s);
            else store_ring<M>(dst, acc1, m_stride, mx);
#pragma unroll
            for (int m = 0; m < M; ++m) acc1[0][m] = acc1[1][m] = make_float2(0.f, 0.f);
        } else {
            if (PACK && NI) {
                RingAcc<M> a;
#pragma unroll
                for (int m = 0; m < M; ++m) { a.v[0][m] = acc0[0][m]; a.v[1][m] = acc0[1][m]; }
                store_ring_packed_ni<M>(pk_row, pk_rsw, a, pk_kk, pk_kk_m, pk_s);
            } else if (PACK) store_ring_packed<M>(pk_row, pk_rsw, acc0, pk_kk, pk_kk_m, pk_s);
            else store_ring<M>(dst, acc0, m_stride, mx);
#pragma unroll
            for (int m = 0; m < M; ++m) acc0[0][m] = acc0[1][m] = make_float2(0.f, 0.f);
        }
        dst += ring_stride;
        pk_kk += pk_kk_ring;
    };

    int fcur = 0;
    const int p0 = rowptr[row], p1 = rowptr[row + 1];
    if (p0 < p1) {
        // Software pipeline, two deep: while edge p is accumulated the feature row of edge p+1 and the plan record
        // of edge p+2 are in flight, so the dependent chain record -> neighbour id -> feature row stays off the
        // FMA pipe's critical path.  All prefetches are unconditional (indices clamped to the row's last edge) so
        // the register rotation unrolls away.  (Measured alternatives that were not faster: an explicit L1 prefetch
        // of the record stream 8 edges ahead plus a two-deep feature gather — more registers, same stalls; packed
        // FFMA2/FMUL2 arithmetic over the lane's two channels — halves the FMA-pipe instructions but nvcc 12.9 spends
        // more than it saves on MOVs that build the aligned 64-bit register pairs.)
        const int last = p1 - 1;
        auto rec_at = [&](int e) { return (DEPTH == 4 && staged) ? s_rec[e - stage_e0] : __ldg(rec + e); };
        auto rot_at = [&](int e) { return (DEPTH == 4 && staged) ? s_rot[e - stage_e0] : __ldg(rot + e); };
        int4 rcA = rec_at(p0);
        float2 rtA = rot_at(p0);
        const int pb = min(p0 + 1, last);
        int4 rcB = rcA;
        float2 rtB = rtA;
        if (DEPTH == 2) {
            rcB = __ldg(rec + pb);
            rtB = __ldg(rot + pb);
        }
        // DEPTH 3 ("light two-deep"): only the neighbour id of edge p+1 travels one iteration ahead of its record (one
        // register and one 4-byte load that hits the line the 16-byte record load touches next), so the feature gather of
        // edge p+1 no longer waits for that record — DEPTH 2's latency tolerance at DEPTH 1's register cost.
        int idN = (DEPTH == 3) ? __ldg(reinterpret_cast<const int*>(rec + pb)) : 0;
        float4 vA = __ldg(fbase + ((uint32_t)rcA.x & NBR_MASK) * (uint32_t)P);
#pragma unroll 2
        for (int p = p0; p < p1; ++p) {
            const int4 rc = rcA;
            const float2 rt = rtA;
            const float4 v = vA;
            if (DEPTH == 3) {
                vA = __ldg(fbase + ((uint32_t)idN & NBR_MASK) * (uint32_t)P);          // edge p+1: id loaded last iteration
                idN = __ldg(reinterpret_cast<const int*>(rec + min(p + 2, last)));
                const int pn = min(p + 1, last);
                rcA = __ldg(rec + pn);
                rtA = __ldg(rot + pn);
            } else if (DEPTH == 2) {
                rcA = rcB;
                rtA = rtB;
                vA = __ldg(fbase + ((uint32_t)rcA.x & NBR_MASK) * (uint32_t)P);
                const int pn = min(p + 2, last);
                rcB = __ldg(rec + pn);
                rtB = __ldg(rot + pn);
            } else {
                const int pn = min(p + 1, last);
                rcA = rec_at(pn);
                rtA = rot_at(pn);
                vA = __ldg(fbase + ((uint32_t)rcA.x & NBR_MASK) * (uint32_t)P);
            }

            const int f = (int)((uint32_t)rc.x >> NBR_BITS);
            while (fcur < f) retire(fcur++);
            const float t = __int_as_float(rc.y);
            const float omt = 1.0f - t;  // fc_precomp.py:25
            const float w0 = (f & 1) ? t : omt;   // weight of the even-ring set
            const float w1 = (f & 1) ? omt : t;   // weight of the odd-ring set
            const float2 wxp = make_float2(__int_as_float(rc.z), __int_as_float(rc.w));
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const float2 z = ch ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                float2 pr[M];
                edge_products<B, TRANSPOSE, FAST>(z, wxp, rt, pr);
                if (FAST) {
                    const float2 w00 = make_float2(w0, w0), w11 = make_float2(w1, w1);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        acc0[ch][m] = __ffma2_rn(w00, pr[m], acc0[ch][m]);
                        acc1[ch][m] = __ffma2_rn(w11, pr[m], acc1[ch][m]);
                    }
                } else {
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    acc0[ch][m].x = fmaf(w0, pr[m].x, acc0[ch][m].x);
                    acc0[ch][m].y = fmaf(w0, pr[m].y, acc0[ch][m].y);
                    acc1[ch][m].x = fmaf(w1, pr[m].x, acc1[ch][m].x);
                    acc1[ch][m].y = fmaf(w1, pr[m].y, acc1[ch][m].y);
                }
                }
            }
        }
    }
    while (fcur < R) retire(fcur++);
    }
    if (!PACK) fold_amax(amax, mx);
}

// Register-allocation variant of the aggregation kernel, 10 * (resident CTAs per SM) + (pipeline depth), per band limit and
// output format.  Measured on B200 (profiles/r01f_layers_occ3.jsonl, r01g_ab_*, r01h_aggregate_variants.jsonl): resident
// warps hide the gather latency better than a deeper software pipeline, until the register cap starts to spill —
//   fp32 output,  band_limit <= 1: 41 (64 registers)   1M vertices C=32: fwd 3.80 -> 2.78 ms, transposed 3.57 -> 2.43 ms
//   fp32 output,  band_limit 2   : 31 (80 registers)   cfg-2 layer: 0.520 -> 0.466 ms, 0.438 -> 0.408 ms  (41: 0.715, spills)
//   packed output, band_limit <= 1: 32                  1M vertices C=32: 3.69 -> 3.03 ms, 3.52 -> 2.64 ms  (41: 3.22 / 3.26)
//   packed output, band_limit 2   : 22 (128 registers)  (32: 0.567 -> 0.695 ms — the fp16 split needs the registers)
// FIELDCONV_B200_AGG_VARIANT=<b0>,<b1>,<b2> (e.g. "32,41,31") overrides the variants of band limits 0, 1, 2 for experiments;
// adding 100 selects the FAST arithmetic (three-term recurrence + FFMA2), e.g. "132,141,131"; 2xx / 3xx (packed output
// only) the out-of-line ring store without / with FAST.
static inline int agg_variant(int band_limit, bool pack) {
    static int tab[3] = {0, 0, 0};
    static bool init = false;
    if (!init) {
        const char* e = getenv("FIELDCONV_B200_AGG_VARIANT");
        if (e) sscanf(e, "%d,%d,%d", &tab[0], &tab[1], &tab[2]);
        init = true;
    }
    if (band_limit < 1 || band_limit > 2) return 22;
    if (tab[band_limit]) return tab[band_limit];
    if (band_limit <= 1) return pack ? 32 : 41;
    return pack ? 22 : 31;
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
    // variant = (resident CTAs per SM, pipeline depth) for this band limit: see agg_variant()
#define FCB_AGG_CASE(b)                                                                                       \
    case b: {                                                                                                 \
        const int var = agg_variant(b, PACK);                                                                 \
        constexpr bool lo = (b == 1 || b == 2);   /* only band limits 1 and 2 have the alternative variants compiled */ \
        if (lo && var == 41) k_aggregate<b, TRANSPOSE, PACK, (lo ? 4 : 2), (lo ? 1 : 2), false> FCB_AGG_ARGS; \
        else if (lo && var == 32) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), 2, false> FCB_AGG_ARGS;       \
        else if (lo && var == 31) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), (lo ? 1 : 2), false> FCB_AGG_ARGS; \
        else if (lo && var == 34) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), (lo ? 4 : 2), false> FCB_AGG_ARGS; \
        else if (lo && var == 44) k_aggregate<b, TRANSPOSE, PACK, (lo ? 4 : 2), (lo ? 4 : 2), false> FCB_AGG_ARGS; \
        else if (lo && var == 134) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), (lo ? 4 : 2), lo> FCB_AGG_ARGS; \
        else if (lo && var == 33) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), (lo ? 3 : 2), false> FCB_AGG_ARGS; \
        else if (lo && var == 43) k_aggregate<b, TRANSPOSE, PACK, (lo ? 4 : 2), (lo ? 3 : 2), false> FCB_AGG_ARGS; \
        else if (lo && var == 133) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), (lo ? 3 : 2), lo> FCB_AGG_ARGS; \
        else if (lo && var == 143) k_aggregate<b, TRANSPOSE, PACK, (lo ? 4 : 2), (lo ? 3 : 2), lo> FCB_AGG_ARGS; \
        else if (lo && var == 141) k_aggregate<b, TRANSPOSE, PACK, (lo ? 4 : 2), (lo ? 1 : 2), lo> FCB_AGG_ARGS; \
        else if (lo && var == 132) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), 2, lo> FCB_AGG_ARGS;         \
        else if (lo && var == 131) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), (lo ? 1 : 2), lo> FCB_AGG_ARGS; \
        else if (lo && var == 122) k_aggregate<b, TRANSPOSE, PACK, 2, 2, lo> FCB_AGG_ARGS;                    \
        else if (lo && PACK && var == 231) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), (lo ? 1 : 2), false, lo && PACK> FCB_AGG_ARGS; \
        else if (lo && PACK && var == 232) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), 2, false, lo && PACK> FCB_AGG_ARGS; \
        else if (lo && PACK && var == 331) k_aggregate<b, TRANSPOSE, PACK, (lo ? 3 : 2), (lo ? 1 : 2), lo, lo && PACK> FCB_AGG_ARGS; \
        else k_aggregate<b, TRANSPOSE, PACK, 2, 2, false> FCB_AGG_ARGS;                                       \
    } break;
    switch (B) {
        FCB_AGG_CASE(0)
        FCB_AGG_CASE(1)
        FCB_AGG_CASE(2)
        FCB_AGG_CASE(3)
        FCB_AGG_CASE(4)
        default: set_error("aggregate: band_limit %d unsupported", B); return FCB_E_UNSUPPORTED;
    }
#undef FCB_AGG_CASE
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
