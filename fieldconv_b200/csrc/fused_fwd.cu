// K1 + K2 fused (band_limit <= 1): gather -> shared-memory operand tile -> tcgen05 contraction, contrib never touches HBM.
//
//   y[i, o] = sum_{r,m,c} contrib[i, r, m, c] W[o, c, r, m],   contrib[i, r, m, c] = sum_{e -> i} sten[e, r, m] xhat[src(e), c, m]
//   (nn/field_conv.py:128-137 with the filter folded as in nn/field_conv.py:10-33; utils/field.py:40-48 for xhat)
//
// One CTA = one tile of 62 target rows (UMMA M = 64: the last two operand rows stay zero) and 1024 threads:
//   warps 0-30  aggregation lanes, lane = (row of the tile, channel pair of the current 32-channel pass): the same gather /
//               gauge alignment / ring-sorted segmented reduction as k_aggregate (aggregate_kernel.cuh).  The edge list of a
//               row is sorted by ring, so when a lane's row moves past ring r its (2B+1) x 2 complex values of that ring are
//               final: the lane splits them into scaled fp16 (hi, lo) pairs and stores them straight into the SWIZZLE_128B
//               K-major operand image of stage (ring parity): 2B+1 chunks of [64 rows x 64 reals], one chunk per frequency m
//               (column of (m, c, re/im) inside the stage = 64 m + 2 c + part).  After fence.proxy.async the 16 lanes of a
//               row sync and one of them arrives on the stage's `a_full` barrier (62 arrivals = the ring is complete for
//               the whole tile).  Rows run freely inside a two-ring window: before writing ring r a lane waits until the
//               MMAs of ring r-2 (same stage) have been committed (`a_empty`).
//   warp 31     one thread: streams the pre-packed filter chunks (k_pack_w_fused: the same K order, fp16 (hi, lo) planes,
//               swizzled) through a shared-memory ring with cp.async.bulk (TMA engine) and issues the MMAs of every completed
//               ring:  [main | cross] += A_hi [W_hi | W_lo]  (one MMA of width 2 Npad),  cross += A_lo W_hi  — the 2xFP16
//               scheme of gemm_h.cu, fp32 accumulators in TMEM, at most 400 accumulating MMAs per accumulator.
//   warps 0-3   epilogue (one warp per TMEM lane quarter; the other aggregation warps exit when their rows are done):
//               tcgen05.ld of the two accumulator blocks, sum, 1/(s_A s_W), 64-byte stores of y.
// C_in > 32 runs as C_in/32 passes over the rows' edge lists (the K order is pass-major), accumulating into the same TMEM tile.
// The operand scale s_A comes from the a-priori bound max|contrib| <= sqrt(2) max|x| max_i sum_{e->i} |wxp_e| (plan norm,
// fc_precomp.py:87 makes the row mass <= 1), folded into wxp per edge, exactly as the packed-operand path does.
#include <cuda_fp16.h>

#include <atomic>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fcb {
namespace ff {

using namespace tc;

constexpr int ROWS = 62;            // target rows per CTA
constexpr int LPR = 16;             // lanes per row: 16 channel pairs = 32 channels per pass
constexpr int AGG_THREADS = ROWS * LPR;     // 992 = 31 warps
constexpr int THREADS = 1024;
constexpr uint32_t A_PLANE = 64 * 128;      // [64 rows x 64 fp16]
constexpr uint32_t A_CHUNK = 2 * A_PLANE;   // hi + lo
constexpr int MAX_SLOTS = 6;

struct Params {
    const float4* x;            // [N_src x Ci] complex, as float4 = two channels
    const int32_t* rowptr;      // by-target CSR of the tile rows
    const int4* rec;
    const float2* rot;
    const __half* Wp;           // packed filter chunks (k_pack_w_fused)
    float* y;                   // [N x 2 Co]
    const float *x_amax, *norm, *w_amax;
    int64_t N;
    int Ci, Co, R, Npad, slots, passes;
    uint32_t tmem_cols;
};

__device__ __forceinline__ float rsqrt_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// wait with back-off: the few warps that outlive their rows must not compete for issue slots with the working ones
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(128);
    }
}

__device__ __forceinline__ void mma_f16_m64(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D = f32, A = B = f16, both K-major, M = 64, N = n
__host__ __device__ inline uint32_t make_idesc_f16_m64(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}

template <int B>
__global__ void __launch_bounds__(THREADS, 1) k_fused_fwd(const Params p) {
    constexpr int M = 2 * B + 1;
    constexpr uint32_t STAGE = M * A_CHUNK;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t b_plane = (uint32_t)p.Npad * 128u;
    const uint32_t w_slot = 2u * b_plane;
    const int S = p.slots;
    // layout: A[2 stages][M chunks][hi | lo] | W[S slots][hi | lo] | barriers
    const uint32_t a0 = base, w0 = base + 2u * STAGE;
    const uint32_t bars = w0 + (uint32_t)S * w_slot;
    auto a_full = [&](int s) { return bars + 8u * s; };
    auto a_empty = [&](int s) { return bars + 8u * (2 + s); };
    auto w_full = [&](int s) { return bars + 8u * (4 + s); };
    auto w_empty = [&](int s) { return bars + 8u * (4 + MAX_SLOTS + s); };
    const uint32_t tmem_full = bars + 8u * (4 + 2 * MAX_SLOTS);
    const uint32_t tmem_slot = bars + 8u * (5 + 2 * MAX_SLOTS);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(sm + (tmem_slot - base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * ROWS;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(a_full(s), ROWS);
            mbar_init(a_empty(s), 1);
        }
        for (int s = 0; s < S; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 31) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // operand rows 62 and 63 of every plane are never written by a lane: zero them once (256 bytes per plane)
    for (int i = tid; i < 2 * M * 2 * 16; i += THREADS) {
        const int plane = i >> 4, u = i & 15;       // plane = (stage, chunk, hi/lo) flattened; 16 x 16 bytes = rows 62, 63
        *reinterpret_cast<uint4*>(sm + (a0 - base) + (uint32_t)plane * A_PLANE + 62u * 128u + (uint32_t)u * 16u) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot_ptr;

    // operand scale of contrib (a-priori bound, as the packed path: aggregate_kernel.cuh)
    const float bound = __ldg(p.x_amax) * __ldg(p.norm) * 1.41422f;
    const uint32_t sf = scale_field(__float_as_uint(fabsf(bound)));
    const float s_a = __uint_as_float(sf << 23);

    if (tid < AGG_THREADS) {
        // ------------------------------------------------------------------ aggregation lanes
        const int rl = tid >> 4, cp = tid & 15;
        const int64_t row = row0 + rl;
        const bool valid = row < p.N;
        const int p0 = valid ? __ldg(p.rowptr + row) : 0, p1 = valid ? __ldg(p.rowptr + row + 1) : 0;
        const unsigned row_mask = 0xFFFFu << (lane & 16);
        const int P = p.Ci >> 1;
        // byte offset of this lane's 8-byte piece inside a plane (row rl, 16-byte unit cp/2 swizzled, half cp&1)
        const uint32_t lane_off = (uint32_t)rl * 128u + ((((uint32_t)cp >> 1) ^ ((uint32_t)rl & 7u)) << 4) + (((uint32_t)cp & 1u) << 3);
        uint8_t* const a_ptr = sm + (a0 - base) + lane_off;
        int qn = 0;                                  // stages retired so far by this lane (= pass * R + ring)
        for (int g = 0; g < p.passes; ++g) {
            const float4* fbase = p.x + g * LPR + cp;
            float2 acc0[2][M], acc1[2][M];           // even rings / odd rings
#pragma unroll
            for (int m = 0; m < M; ++m) {
                acc0[0][m] = acc0[1][m] = make_float2(0.f, 0.f);
                acc1[0][m] = acc1[1][m] = make_float2(0.f, 0.f);
            }
            // ring complete: split into fp16 (hi, lo), store into the stage's operand image, signal, clear
            auto retire = [&](int ring) {
                const int s = qn & 1;
                mbar_wait(a_empty(s), (uint32_t)(((qn >> 1) & 1) ^ 1));
                uint8_t* dst = a_ptr + (uint32_t)s * STAGE;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float2 c0 = (ring & 1) ? acc1[0][m] : acc0[0][m];
                    const float2 c1 = (ring & 1) ? acc1[1][m] : acc0[1][m];
                    const __half2 h01 = __floats2half2_rn(c0.x, c0.y), h23 = __floats2half2_rn(c1.x, c1.y);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    const __half2 l01 = __floats2half2_rn(c0.x - f01.x, c0.y - f01.y), l23 = __floats2half2_rn(c1.x - f23.x, c1.y - f23.y);
                    *reinterpret_cast<uint2*>(dst + (uint32_t)m * A_CHUNK) =
                        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                    *reinterpret_cast<uint2*>(dst + (uint32_t)m * A_CHUNK + A_PLANE) =
                        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                }
                if (ring & 1) {
#pragma unroll
                    for (int m = 0; m < M; ++m) acc1[0][m] = acc1[1][m] = make_float2(0.f, 0.f);
                } else {
#pragma unroll
                    for (int m = 0; m < M; ++m) acc0[0][m] = acc0[1][m] = make_float2(0.f, 0.f);
                }
                fence_proxy_async();
                __syncwarp(row_mask);
                if (cp == 0) mbar_arrive(a_full(s));
                ++qn;
            };
            int fcur = 0;
            if (p0 < p1) {
                const int last = p1 - 1;
                int4 rcA = __ldg(p.rec + p0);
                float2 rtA = __ldg(p.rot + p0);
                float4 vA = __ldg(fbase + ((uint32_t)rcA.x & NBR_MASK) * (uint32_t)P);
#pragma unroll 2
                for (int e = p0; e < p1; ++e) {
                    const int4 rc = rcA;
                    const float2 rt = rtA;
                    const float4 v = vA;
                    const int en = min(e + 1, last);
                    rcA = __ldg(p.rec + en);
                    rtA = __ldg(p.rot + en);
                    vA = __ldg(fbase + ((uint32_t)rcA.x & NBR_MASK) * (uint32_t)P);
                    const int f = (int)((uint32_t)rc.x >> NBR_BITS);
                    while (fcur < f) retire(fcur++);
                    const float t = __int_as_float(rc.y);
                    const float omt = 1.0f - t;  // fc_precomp.py:25
                    const float w0 = (f & 1) ? t : omt, w1 = (f & 1) ? omt : t;
                    const float2 w00 = make_float2(w0, w0), w11 = make_float2(w1, w1);
                    const float2 wxp = make_float2(__int_as_float(rc.z) * s_a, __int_as_float(rc.w) * s_a);
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const float2 z = ch ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                        float2 pr[M];
                        pr[B] = cmul(wxp, z);
                        if (B >= 1) {
                            // branch-free: at origin entries (|re|,|im| < 1e-7) the selects discard the inf/NaN of rsqrt(0)
                            const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
                            const float ri = rsqrt_ftz(fmaf(z.x, z.x, z.y * z.y));
                            const float ux = origin ? 1.f : z.x * ri;
                            const float uy = origin ? 0.f : z.y * ri;
                            const float2 q = cmul_conj(rt, make_float2(ux, uy));
                            pr[B + 1] = cmul(pr[B], q);
                            const float2 c2 = make_float2(2.f * q.x, 2.f * q.x);
                            pr[B - 1] = __ffma2_rn(c2, pr[B], make_float2(-pr[B + 1].x, -pr[B + 1].y));
                        }
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            acc0[ch][m] = __ffma2_rn(w00, pr[m], acc0[ch][m]);
                            acc1[ch][m] = __ffma2_rn(w11, pr[m], acc1[ch][m]);
                        }
                    }
                }
            }
            while (fcur < p.R) retire(fcur++);
        }
        if (warp >= 4) return;       // the epilogue needs one warp per TMEM lane quarter: everyone else is done
    } else if (tid == AGG_THREADS) {
        // ------------------------------------------------------------------ filter loader + MMA issuer (one thread)
        const int stages_total = p.passes * p.R;
        const int total = stages_total * M;
        const __half* src = p.Wp;
        const int64_t src_step = (int64_t)2 * p.Npad * 64;          // fp16 elements per chunk (hi + lo planes)
        int next_load = 0;
        auto load = [&](int i) {
            const int slot = i % S;
            mbar_wait(w_empty(slot), (uint32_t)(((i / S) & 1) ^ 1));
            mbar_expect_tx(w_full(slot), w_slot);
            bulk_copy_g2s(w0 + (uint32_t)slot * w_slot, src + (int64_t)i * src_step, w_slot, w_full(slot));
        };
        for (; next_load < total && next_load < S; ++next_load) load(next_load);
        const uint32_t npad = (uint32_t)p.Npad;
        const bool merged = 2 * p.Npad <= 256;
        const uint32_t idesc1 = make_idesc_f16_m64(p.Npad), idesc2 = make_idesc_f16_m64(2 * p.Npad);
        uint32_t first = 1;
        for (int qs = 0; qs < stages_total; ++qs) {
            const int s = qs & 1;
            mbar_wait(a_full(s), (uint32_t)((qs >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int m = 0; m < M; ++m) {
                const int i = qs * M + m, slot = i % S;
                mbar_wait(w_full(slot), (uint32_t)((i / S) & 1));
                tc_fence_after();
                const uint64_t a_hi = make_desc_k_sw128(a0 + (uint32_t)s * STAGE + (uint32_t)m * A_CHUNK);
                const uint64_t a_lo = make_desc_k_sw128(a0 + (uint32_t)s * STAGE + (uint32_t)m * A_CHUNK + A_PLANE);
                const uint64_t b_hi = make_desc_k_sw128(w0 + (uint32_t)slot * w_slot);
                const uint64_t b_lo = make_desc_k_sw128(w0 + (uint32_t)slot * w_slot + b_plane);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);   // 16 fp16 = 32 B inside the 128-byte swizzle row
                    if (merged) {
                        mma_f16_m64(tmem_d, a_hi + adv, b_hi + adv, idesc2, first ? 0u : 1u);       // [main | cross] (+)= A_hi [W_hi | W_lo]
                    } else {
                        mma_f16_m64(tmem_d, a_hi + adv, b_hi + adv, idesc1, first ? 0u : 1u);
                        mma_f16_m64(tmem_d + npad, a_hi + adv, b_lo + adv, idesc1, first ? 0u : 1u);
                    }
                    mma_f16_m64(tmem_d + npad, a_lo + adv, b_hi + adv, idesc1, 1u);                  // cross += A_lo W_hi
                    first = 0;
                }
                tc_commit(w_empty(slot));
                // refill the slot of the PREVIOUS chunk (its MMAs were committed one iteration ago), not the one just issued
                if (i >= 1 && next_load < total && next_load == i - 1 + S) load(next_load++);
            }
            tc_commit(a_empty(s));
        }
        tc_commit(tmem_full);
    }
    __syncwarp();
    // ---------------------------------------------------------------------- epilogue (warps 0-3: one per TMEM lane quarter)
    if (warp < 4) {
        mbar_wait_sleep(tmem_full, 0);
        tc_fence_after();
        const float inv = __uint_as_float((254u - sf) << 23) * inv_scale_of(p.w_amax);
        const int q = warp;                             // TMEM lane quarter
        const int rl = 16 * q + lane;                   // D row of this thread (M = 64: rows 16q..16q+15 sit on lanes 0..15)
        const int64_t row = row0 + rl;
        const int groups = p.Npad / 16;
        const int n2 = 2 * p.Co;
        const uint32_t lane_base = tmem_d + ((uint32_t)(32 * q) << 16);
        for (int g = 0; g < groups; ++g) {
            uint32_t r0[16], r1[16];
            tc_ld16(lane_base + (uint32_t)(16 * g), r0);
            tc_ld16(lane_base + (uint32_t)(p.Npad + 16 * g), r1);
            tc_ld_wait();
            if (lane < 16 && rl < ROWS && row < p.N) {
                float* dst = p.y + row * (int64_t)n2 + 16 * g;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = (__uint_as_float(r0[4 * c4 + e]) + __uint_as_float(r1[4 * c4 + e])) * inv;
                    const int n = 16 * g + 4 * c4;
                    if (n + 3 < n2) {
                        *reinterpret_cast<float4*>(dst + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + e < n2) dst[4 * c4 + e] = o[e];
                    }
                }
            }
        }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 160;" ::: "memory");      // warps 0-3 (TMEM reads done) and warp 31 (owner of the allocation)
    if (warp == 31) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
    }
}

// W (Co,Ci,R,M) complex -> chunk images in the fused kernel's K order: chunk i = ((pass g) * R + r) * M + m holds the 64
// real columns (c = 32 g + kk/2, part = kk & 1) of ring r, frequency m, as [plane hi | lo][Npad rows][64 fp16] with the
// 128-byte swizzle (unit u of row n at u ^ (n & 7)).  Row n = 2 o + b of the real embedding (api.cu k_pack_w_fwd):
// part 0 -> (w.x, w.y)[b], part 1 -> (-w.y, w.x)[b].  One thread per 16-byte piece.
__global__ void k_pack_w_fused(const float2* __restrict__ W, __half* __restrict__ Wp, int Ci, int Co, int R, int M, int Npad,
                               int64_t total_chunks, const float* __restrict__ amax_w) {
    const int64_t per = (int64_t)Npad * 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_chunks * per) return;
    const int64_t chunk = i / per;
    const int rem = (int)(i - chunk * per);
    const int n = rem % Npad, u = rem / Npad;
    const int m = (int)(chunk % M), r = (int)((chunk / M) % R), g = (int)(chunk / ((int64_t)M * R));
    const float s = scale_of(amax_w);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (n < 2 * Co) {
        const int o = n >> 1, b = n & 1;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int kk = 8 * u + e;
            const int c = 32 * g + (kk >> 1), part = kk & 1;
            if (c < Ci) {
                const float2 w = W[(((int64_t)o * Ci + c) * R + r) * M + m];
                v[e] = part == 0 ? (b == 0 ? w.x : w.y) : (b == 0 ? -w.y : w.x);
            }
        }
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a0 = v[2 * j] * s, a1 = v[2 * j + 1] * s;
        const __half2 hh = __floats2half2_rn(a0, a1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
        h[j] = *reinterpret_cast<const uint32_t*>(&hh);
        l[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    __half* dst = Wp + chunk * (2 * (int64_t)Npad * 64) + (int64_t)n * 64 + ((u ^ (n & 7)) << 3);
    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst + (int64_t)Npad * 64) = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace ff

// shapes the fused kernel takes: band_limit <= 1, C_in a multiple of 32, 2 C_out padded to 16 fits one accumulator pair
// in TMEM (C_out <= 128), at most 400 accumulating MMAs per accumulator
bool fused_fwd_ok(int Ci, int Co, int B, int R) {
    if (B < 0 || B > 1 || R < 2 || R > FCB_MAX_RINGS) return false;
    if (Ci < 32 || (Ci % 32) != 0 || Co < 1 || (Co & 1) != 0) return false;
    const int npad = (2 * Co + 15) / 16 * 16;
    if (2 * npad > 512) return false;
    const int64_t ksteps = (int64_t)(Ci / 32) * R * (2 * B + 1) * 4;
    return ksteps <= 400;
}

static int fused_plan(int Co, int B, int* npad_out, int* slots_out, size_t* smem_out) {
    const int M = 2 * B + 1;
    const int npad = (2 * Co + 15) / 16 * 16;
    const size_t stage = (size_t)M * ff::A_CHUNK;
    const size_t w_slot = 2 * (size_t)npad * 128;
    const size_t fixed = 2 * stage + 1024 /* alignment slack */ + 8 * (6 + 2 * ff::MAX_SLOTS) + 64;
    int slots = (int)((227 * 1024 - fixed) / w_slot);
    if (slots > ff::MAX_SLOTS) slots = ff::MAX_SLOTS;
    if (slots < 2) return 0;
    *npad_out = npad;
    *slots_out = slots;
    *smem_out = fixed + (size_t)slots * w_slot;
    return 1;
}

size_t fused_fwd_ws_bytes(int Ci, int Co, int B, int R) {
    const int npad = (2 * Co + 15) / 16 * 16;
    const int64_t chunks = (int64_t)(Ci / 32) * R * (2 * B + 1);
    return 512 + align_up((size_t)chunks * 2 * npad * 64 * 2, 256) + 256;
}

// ws: [0] max|x| (real components), [64 B] max|W|, [512 B ...] packed filter
int launch_fused_fwd(const float* x, const float* W, const int32_t* rowptr, const void* rec, const float* rot, const float* norm,
                     float* y, int64_t N, int64_t n_feat, int Ci, int Co, int B, int R, void* ws, size_t ws_bytes, cudaStream_t st) {
    FCB_REQUIRE(fused_fwd_ok(Ci, Co, B, R), FCB_E_UNSUPPORTED, "fused_fwd: shape not supported (fcb_fused_supported)");
    FCB_REQUIRE(x && W && rowptr && rec && rot && norm && y && ws, FCB_E_ARG, "fused_fwd: null pointer");
    FCB_REQUIRE(ws_bytes >= fused_fwd_ws_bytes(Ci, Co, B, R), FCB_E_WORKSPACE, "fused_fwd: workspace too small");
    FCB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(W) && aligned16(rec) && aligned16(ws), FCB_E_ALIGN, "fused_fwd: pointers must be 16-byte aligned");
    if (N == 0) return FCB_OK;
    int npad = 0, slots = 0;
    size_t smem = 0;
    FCB_REQUIRE(fused_plan(Co, B, &npad, &slots, &smem), FCB_E_UNSUPPORTED, "fused_fwd: tile does not fit shared memory");
    const int M = 2 * B + 1;
    float* x_amax = static_cast<float*>(ws);
    float* w_amax = static_cast<float*>(ws) + 16;
    __half* Wp = reinterpret_cast<__half*>(static_cast<char*>(ws) + 512);
    int rc = launch_absmax_f32(x, n_feat, 2 * Ci, 2 * (int64_t)Ci, 1, 0, x_amax, st);   // every row a target may gather from
    if (rc) return rc;
    rc = launch_absmax_f32(W, (int64_t)Co * Ci * R * M, 2, 2, 1, 0, w_amax, st);
    if (rc) return rc;
    const int64_t chunks = (int64_t)(Ci / 32) * R * M;
    {
        const int64_t items = chunks * npad * 8;
        FCB_LAUNCH("pack_w_fused", st, ff::k_pack_w_fused<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(
                                          reinterpret_cast<const float2*>(W), Wp, Ci, Co, R, M, npad, chunks, w_amax));
    }
    ff::Params p;
    p.x = reinterpret_cast<const float4*>(x);
    p.rowptr = rowptr;
    p.rec = static_cast<const int4*>(rec);
    p.rot = reinterpret_cast<const float2*>(rot);
    p.Wp = Wp;
    p.y = y;
    p.x_amax = x_amax; p.norm = norm; p.w_amax = w_amax;
    p.N = N; p.Ci = Ci; p.Co = Co; p.R = R; p.Npad = npad; p.slots = slots; p.passes = Ci / 32;
    uint32_t cols = 32;
    while ((int)cols < 2 * npad) cols <<= 1;
    p.tmem_cols = cols;
    static std::atomic<bool> attr_set_dev[64];
    std::atomic<bool> attr_unknown_dev{false};
    int attr_dev = 0;
    if (cudaGetDevice(&attr_dev) != cudaSuccess) attr_dev = -1;
    std::atomic<bool>& attr_set = (attr_dev >= 0 && attr_dev < 64) ? attr_set_dev[attr_dev] : attr_unknown_dev;
    if (!attr_set.load(std::memory_order_acquire)) {       // idempotent: racing threads at worst set the attribute twice
        cudaError_t e = cudaFuncSetAttribute(ff::k_fused_fwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ff::k_fused_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("fused_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return FCB_E_CUDA;
        }
        attr_set.store(true, std::memory_order_release);
    }
    const unsigned grid = (unsigned)((N + ff::ROWS - 1) / ff::ROWS);
    if (B == 0) FCB_LAUNCH("fused_fwd", st, ff::k_fused_fwd<0><<<grid, ff::THREADS, smem, st>>>(p));
    else FCB_LAUNCH("fused_fwd", st, ff::k_fused_fwd<1><<<grid, ff::THREADS, smem, st>>>(p));
    return FCB_OK;
}

}  // namespace fcb
