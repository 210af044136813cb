// K2 / K4 / K5b on the 5th-generation tensor cores, "2xFP16" mode: C = A * B with fp32 inputs and fp32 output, computed
// with tcgen05.mma kind::f16 on operands that are split into TWO fp16 planes each.
//
// Why fp16 pairs instead of 3xTF32 (gemm_tc.cu): a tf32 operand occupies 4 bytes of shared memory for 11 bits of
// significand, so the SS-mode MMAs of the 3xTF32 kernel are bound by shared-memory bandwidth (ncu r01c: tensor pipe 36 %).
// An fp16 carries the same 11 bits in 2 bytes and the f16 MMA runs at twice the tf32 rate, so the identical
// three-product scheme  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  moves half the bytes and takes half the tensor time.
//   hi = fp16(s*x), lo = fp16(s*x - hi):  hi + lo carries 22 bits of s*x (relative error 2^-23 while lo is a normal
//   fp16, absolute error <= 2^-25 below that).  fp16 has a 5-bit exponent, so each operand is first scaled by a power
//   of two s = 2^(14 - floor(log2 max|x|)) that puts its largest entry in [2^14, 2^15): entries down to 2^-17 of the
//   largest keep full precision, smaller ones lose bits only in ABSOLUTE terms (<= 2^-40 of the largest entry) — far
//   inside the normwise fp32 parity budget.  The epilogue multiplies by 1/(s_A s_B), exact.  max|x| comes from the
//   kernel that produced the operand (the aggregation kernels track it for free) or from k_absmax.
//   The accumulator behaviour is the one described in gemm_tc.cu: no accumulator receives more than 400 accumulating
//   MMAs of hi*hi products (one f16 MMA covers 16 reals of K, twice a tf32 MMA).
//
// NN kernel (y = contrib W, gxh = G conj(W)): one CTA = one 128-row tile of A and all N <= 128 columns.
//   warps 0-15 producers: 256-bit coalesced loads of the fp32 A tile (64 reals = 256 B per row and stage), scale + hi/lo
//              split in registers, 128-bit stores into the canonical SWIZZLE_128B K-major layout (one 128-byte row =
//              64 fp16 of one A row), fence.proxy.async, mbarrier arrive; afterwards the epilogue.
//   warp 16    single-thread MMA issue.  The two B planes are adjacent in shared memory, so  A_hi * [B_hi | B_lo]  is ONE
//              MMA of width 2*Npad whose accumulator holds the main product in columns [0, Npad) and the hi*lo cross
//              term in [Npad, 2 Npad);  A_lo * B_hi  is a second MMA of width Npad accumulating into the cross block:
//              A_hi is read from shared memory once for two products (-19 % operand bytes).
//   warp 17    B operand: pre-packed images (k_pack_b_h) fetched with one cp.async.bulk per stage.
// TN kernel (P = contrib^T gy, the weight gradient): both operands MN-major (the hardware transposes), SWIZZLE_128B
//   atoms of 8 vertices x 64 fp16; gy is packed once (k_pack_b_h_tn), contrib goes through the producer warps.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fcb {
namespace th {

using tc::BM;
using namespace tc;

constexpr int KC = 64;         // reals per stage (NN): one 128-byte swizzle row of fp16
constexpr int KV = 64;         // vertices per stage (TN): four MMA K-steps of 16
constexpr int N_PROD_WARPS = 16;
constexpr int N_PROD = N_PROD_WARPS * 32;
constexpr int THREADS = (N_PROD_WARPS + 2) * 32;
constexpr uint32_t A_PLANE = BM * 128;   // 16 KB: 128 rows x 64 fp16 (NN) or 64 vertices x 128 columns (TN)
constexpr int UN = 2;          // 16-byte output units (8 reals) per producer thread and stage
constexpr int PF = 3;          // chunks of A in flight in registers per producer thread (96 KB per SM)

__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Instruction descriptor: D = f32 (bit 4), A = B = f16 (format 0), M = 128, N = n; K-major unless the major bits are or-ed in.
__host__ __device__ inline uint32_t make_idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// MN-major SWIZZLE_128B descriptor (16-bit operands): atoms of 64 MN-elements (128 B) x 8 K-rows; LBO = distance between
// atoms along MN, SBO = distance between 8-row K groups.
__device__ __forceinline__ uint64_t make_desc_mn_sw128_h(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}

struct F8 {
    float v[8];
};
__device__ __forceinline__ void ld256(const float* p, F8& r) {   // one 256-bit read-only load (32-byte aligned)
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld2x128(const float* p, F8& r) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
}
__device__ __forceinline__ void zero8(F8& r) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = 0.f;
}
// 8 reals -> 8 fp16 hi + 8 fp16 lo of s*x (element e at the lower address)
__device__ __forceinline__ void split8(const F8& x, float s, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a0 = x.v[2 * i] * s, a1 = x.v[2 * i + 1] * s;
        const __half2 hh = __floats2half2_rn(a0, a1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ------------------------------------------------------------------------------------------------------------ max|x|
// out (uint32 bit pattern of a non-negative float, or NaN bits) = max(out, max |p[b][r][c]|); out is pre-zeroed.
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ p, int64_t rows, int cols, int64_t ld, int batch,
                                                int64_t stride, int flat4, uint32_t* __restrict__ out) {
    uint32_t mx = 0;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    if (flat4) {       // contiguous and 16-byte aligned: float4 sweep
        const int64_t n4 = (int64_t)batch * rows * cols / 4;
        const float4* q = reinterpret_cast<const float4*>(p);
        auto take = [&](const float4 v) {
            mx = max(mx, __float_as_uint(fabsf(v.x)));
            mx = max(mx, __float_as_uint(fabsf(v.y)));
            mx = max(mx, __float_as_uint(fabsf(v.z)));
            mx = max(mx, __float_as_uint(fabsf(v.w)));
        };
        int64_t i = t0;
        for (; i + 3 * nt < n4; i += 4 * nt) {      // four independent loads in flight per thread
            const float4 v0 = __ldg(q + i), v1 = __ldg(q + i + nt), v2 = __ldg(q + i + 2 * nt), v3 = __ldg(q + i + 3 * nt);
            take(v0); take(v1); take(v2); take(v3);
        }
        for (; i < n4; i += nt) take(__ldg(q + i));
    } else {
        const int64_t per = rows * cols, tot = per * batch;
        for (int64_t i = t0; i < tot; i += nt) {
            const int64_t b = i / per, r = (i - b * per) / cols;
            const int c = (int)(i - b * per - r * cols);
            mx = max(mx, __float_as_uint(fabsf(p[b * stride + r * ld + c])));
        }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx > *reinterpret_cast<volatile uint32_t*>(out)) atomicMax(out, mx);
}

static int launch_absmax(const float* p, int64_t rows, int cols, int64_t ld, int batch, int64_t stride, float* out,
                         cudaStream_t st) {
    if (cudaMemsetAsync(out, 0, 4, st) != cudaSuccess) {
        set_error("absmax: cudaMemsetAsync failed");
        return FCB_E_CUDA;
    }
    const int64_t tot = (int64_t)batch * rows * cols;
    if (tot == 0) return FCB_OK;
    const bool contiguous = (ld == cols) && (batch == 1 || stride == rows * (int64_t)cols);
    const int flat4 = (contiguous && (tot % 4) == 0 && aligned16(p)) ? 1 : 0;
    int64_t blocks = (tot / (flat4 ? 4 : 1) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    FCB_LAUNCH("absmax", st, k_absmax<<<(unsigned)blocks, 256, 0, st>>>(p, rows, cols, ld, batch, stride, flat4,
                                                                        reinterpret_cast<uint32_t*>(out)));
    return FCB_OK;
}

// ------------------------------------------------------------------------------------------------------------ NN
struct Params {
    const float* A;
    const __half* Bp;   // packed B: [batch][chunk of 64 k][plane(hi,lo)][Npad][64 fp16] swizzled (k_pack_b_h)
    float* C;
    const float *amax_a, *amax_b;
    int64_t M, K, lda, ldc, sa, sc;
    int64_t bp_batch_stride;   // fp16 elements
    int N, Npad, nchunks, stages;
    int n_pairs;               // (main, cross) accumulator pairs the k-steps are dealt over, round-robin
    int kgroups, cpg;          // grouped-K mode (kgroups > 1): K = kgroups * cpg chunks; group g accumulates all three
                               // products into its own accumulator and lands in C columns [g*N, (g+1)*N)
    int wide;                  // rows are 32-byte aligned: 256-bit loads
    int64_t a_tile_stride;     // PACKED: bytes between consecutive 128-row tiles of the PK buffer
    int rows_per_tile;         // fp32 A: rows a CTA owns (<= BM, multiple of 8); BM for a PK operand
    int reverse;               // row tiles taken last first
    // split-K (blockIdx.z, kgroups == 1 only): split z contracts the chunks [z * cps, min(nchunks, (z + 1) * cps)) and writes
    // its [M x N] partial at C + z * part_stride (ldc = N); the caller sums the partials in split order (deterministic).
    // Used for WIDE outputs (128 < N <= 256): all N columns in one accumulator pair (512 TMEM columns) and the K range cut
    // so that no accumulator takes more than 400 accumulating MMAs — A is read ONCE instead of once per 128-column chunk.
    int cps;
    int64_t part_stride;
    // fused block epilogue (kgroups == 1, split-K off): C receives z = A B (+ epi_res), epi_act = modReLU(z, epi_bias) —
    // nn/fc_resnet_block.py:84-88 with nn/tangent_nonlin.py:24-35 applied in the TMEM -> register epilogue
    const float* epi_res;      // optional [M x N] residual (row stride epi_ld)
    const float* epi_bias;     // N / 2 modReLU biases (one per complex output channel), nullptr: no activation output
    float* epi_act;            // [M x N] activated output (row stride epi_ld)
    int64_t epi_ld;
    uint32_t* epi_bound;       // nullable: bit pattern of max_i |act_i| (modulus), folded in with atomicMax
    // fused softAngle chain rule (grouped-K mode, kgroups = 2 * sa_band + 1): the k-group accumulators ARE gxhat[:, m, :] of one
    // row, so the epilogue turns them into grad x directly (utils/field.py:40-48) — gxhat never
    // goes to memory and k_softangle_bwd is not launched.  sa_x / sa_gx: [M x N/2] complex, row stride N floats.
    const float* sa_x;
    float* sa_gx;
    int sa_band;
    uint32_t tmem_cols;
};

// gx of 4 complex channels from their gxhat accumulators: non-origin  h_m = conj(g_m) u^(1-m),  gx = u (sum Re h_m - i sum (1-m) Im h_m);
// origin  gx = sum_m g_m
template <int BL>
__device__ __forceinline__ void softangle_epilogue(uint32_t lane_base, int npad, int blk, float inv, const float* __restrict__ xrow,
                                                   float* __restrict__ grow) {
    constexpr int M = 2 * BL + 1;
    uint32_t r[M][8];
#pragma unroll
    for (int kg = 0; kg < M; ++kg) tc_ld8(lane_base + (uint32_t)(kg * npad + 8 * blk), r[kg]);
    tc_ld_wait();
    if (!xrow) return;
    const float4 xa = __ldg(reinterpret_cast<const float4*>(xrow + 8 * blk)), xb = __ldg(reinterpret_cast<const float4*>(xrow + 8 * blk) + 1);
    const float2 zs[4] = {make_float2(xa.x, xa.y), make_float2(xa.z, xa.w), make_float2(xb.x, xb.y), make_float2(xb.z, xb.w)};
    float o[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float2 g[M];
#pragma unroll
        for (int kg = 0; kg < M; ++kg) g[kg] = make_float2(__uint_as_float(r[kg][2 * c]) * inv, __uint_as_float(r[kg][2 * c + 1]) * inv);
        const float2 z = zs[c];
        const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
        float2 out;
        if (origin) {
            out = make_float2(0.f, 0.f);
#pragma unroll
            for (int kg = 0; kg < M; ++kg) { out.x += g[kg].x; out.y += g[kg].y; }
        } else {
            const float ri = rsqrtf(z.x * z.x + z.y * z.y);
            const float2 u = make_float2(z.x * ri, z.y * ri);
            float a = 0.f, b = 0.f;
            {
                const float2 h = cmul(make_float2(g[BL].x, -g[BL].y), u);
                a += h.x; b -= h.y;
            }
            float2 pos = u, neg = u;
#pragma unroll
            for (int k = 1; k <= BL; ++k) {
                pos = cmul(pos, u);              // u^(1+k)
                neg = cmul_conj(neg, u);         // u^(1-k)
                const float2 hm = cmul(make_float2(g[BL - k].x, -g[BL - k].y), pos);   // m = -k
                const float2 hp = cmul(make_float2(g[BL + k].x, -g[BL + k].y), neg);   // m = +k
                a += hm.x + hp.x;
                b -= (float)(1 + k) * hm.y + (float)(1 - k) * hp.y;
            }
            out = cmul(u, make_float2(a, b));
        }
        o[2 * c] = out.x; o[2 * c + 1] = out.y;
    }
    *reinterpret_cast<float4*>(grow + 8 * blk) = make_float4(o[0], o[1], o[2], o[3]);
    *(reinterpret_cast<float4*>(grow + 8 * blk) + 1) = make_float4(o[4], o[5], o[6], o[7]);
}

// PACKED: A is a PK buffer (common.cuh) — the (hi, lo) planes of every 64-column chunk are already the swizzled tile
// images, so the loader thread brings them in with two 16 KB bulk copies per stage on the stage's `full_b` barrier and
// the 16 producer warps only run the epilogue.
// PAIRED (fp32 A only): the producers issue the loads of TWO consecutive chunks back to back, so the two 256-byte pieces
// of a row's 512 bytes reach DRAM together (one page activation instead of two) — 4 chunks of registers instead of 3.
template <bool PACKED, bool PAIRED>
__global__ void __launch_bounds__(THREADS, 1) k_gemm_h_nn(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int S = p.stages;
    const uint32_t b_plane = (uint32_t)p.Npad * 128u;
    // layout: A_hi[S] | A_lo[S] | B[S] (hi plane, lo plane adjacent) | barriers
    const uint32_t a_hi0 = base, a_lo0 = base + S * A_PLANE, b0 = base + 2 * S * A_PLANE;
    const uint32_t bars = b0 + S * 2 * b_plane;
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto full_b = [&](int s) { return bars + 8u * (S + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * S + s); };
    const uint32_t tmem_full = bars + 8u * (3 * S);
    const uint32_t tmem_slot = bars + 8u * (3 * S + 1);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(sm + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Row tiles are taken LAST FIRST: the operand was just written front to back by the aggregation kernel, so its tail is
    // what the (write-back) L2 still holds when this kernel starts.
    const int64_t tile_x = p.reverse ? (int64_t)(gridDim.x - 1u - blockIdx.x) : (int64_t)blockIdx.x;
    const int64_t m0 = tile_x * p.rows_per_tile;
    const int64_t m_end = min(p.M, m0 + p.rows_per_tile);
    const int batch = blockIdx.y;
    const int kc_begin = (int)blockIdx.z * p.cps;
    const int n_kc = min(p.nchunks, kc_begin + p.cps) - kc_begin;       // chunks of this CTA (>= 1 by construction)
    const float* A = p.A + batch * p.sa;
    const __half* Bp = p.Bp + batch * p.bp_batch_stride;
    float* C = p.C + batch * p.sc + (int64_t)blockIdx.z * p.part_stride;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_a(s), N_PROD_WARPS);
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == N_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot_ptr;

    if (warp < N_PROD_WARPS) {
        if (!PACKED) {
        // ------------------------------------------------------------------ producers
        const int t = threadIdx.x;
        const float* src[UN];
        uint32_t off[UN];
        int kcol[UN];
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            const int idx = t + N_PROD * i;
            const int row = idx >> 3, j = idx & 7;          // 8 units of 8 reals per row and stage
            const int64_t m = m0 + row;
            src[i] = (m < m_end) ? (A + m * p.lda + 8 * j) : nullptr;
            kcol[i] = 8 * j;
            off[i] = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
        }
        uint8_t* const hi_base = sm + (a_hi0 - base);
        uint8_t* const lo_base = sm + (a_lo0 - base);
        const float s_a = scale_of(p.amax_a);
        const bool wide = p.wide != 0;
        F8 v[PAIRED ? 4 : PF][UN];
        auto issue = [&](int kc, F8(&dst)[UN]) {
            const int64_t k0 = (int64_t)(kc_begin + kc) * KC;
            if (k0 + KC <= p.K) {                        // full chunk (CTA-uniform)
#pragma unroll
                for (int i = 0; i < UN; ++i) {
                    if (!src[i]) zero8(dst[i]);
                    else if (wide) ld256(src[i] + k0, dst[i]);
                    else ld2x128(src[i] + k0, dst[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < UN; ++i) {
                    zero8(dst[i]);
                    if (src[i]) {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (k0 + kcol[i] + e < p.K) dst[i].v[e] = src[i][k0 + e];
                    }
                }
            }
        };
        uint32_t ps = 0, pph = 1;     // producer stage / parity of the `empty` barrier it waits for
        auto commit = [&](const F8(&sv)[UN]) {
            const uint32_t s = ps;
            mbar_wait(empty(s), pph);
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                uint4 hi, lo;
                split8(sv[i], s_a, hi, lo);
                *reinterpret_cast<uint4*>(hi_base + s * A_PLANE + off[i]) = hi;
                *reinterpret_cast<uint4*>(lo_base + s * A_PLANE + off[i]) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(s));
            if (++ps == (uint32_t)S) { ps = 0; pph ^= 1u; }
        };
        if (PAIRED) {
            const int n = n_kc;
            if (0 < n) issue(0, v[0]);
            if (1 < n) issue(1, v[1]);
            for (int kc = 0; kc < n; kc += 4) {          // slots 0,1 hold chunks kc, kc+1; slots 2,3 take kc+2, kc+3
                if (kc + 2 < n) issue(kc + 2, v[2]);
                if (kc + 3 < n) issue(kc + 3, v[3]);
                commit(v[0]);
                if (kc + 1 < n) commit(v[1]);
                if (kc + 4 < n) issue(kc + 4, v[0]);
                if (kc + 5 < n) issue(kc + 5, v[1]);
                if (kc + 2 < n) commit(v[2]);
                if (kc + 3 < n) commit(v[3]);
            }
        } else {
#pragma unroll
        for (int u = 0; u < PF - 1; ++u)
            if (u < n_kc) issue(u, v[u]);
        for (int kc = 0; kc < n_kc; kc += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int k = kc + u;
                if (k < n_kc) {
                    if (k + PF - 1 < n_kc) issue(k + PF - 1, v[(u + PF - 1) % PF]);
                    commit(v[u]);
                }
            }
        }
        }
        }
        // ------------------------------------------------------------------ epilogue
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const float inv_a = inv_scale_of(p.amax_a), inv_b = inv_scale_of(p.amax_b);
        const int q = warp & 3, part = warp >> 2;
        const int64_t m = m0 + 32 * q + lane;
        const int groups = p.Npad / 16;
        const bool grouped = p.kgroups > 1;
        const int items = groups * (grouped ? p.kgroups : 1);       // (k-group, 16-column group) pairs
        const uint32_t lane_base = tmem_d + ((uint32_t)(32 * q) << 16);
        if (grouped && p.sa_gx) {
            // grad x straight from the k-group accumulators (N % 8 == 0 guaranteed by the launcher)
            const float inv = inv_a * inv_b;
            const float* xrow = m < m_end ? p.sa_x + m * (int64_t)p.N : nullptr;
            float* grow = p.sa_gx + m * (int64_t)p.N;
            for (int blk = part; blk < p.N / 8; blk += N_PROD_WARPS / 4) {
                switch (p.sa_band) {
                    case 1: softangle_epilogue<1>(lane_base, p.Npad, blk, inv, xrow, grow); break;
                    case 2: softangle_epilogue<2>(lane_base, p.Npad, blk, inv, xrow, grow); break;
                    default: softangle_epilogue<3>(lane_base, p.Npad, blk, inv, xrow, grow); break;
                }
            }
        } else {
        float act_mx = 0.f;          // largest modulus this thread wrote to epi_act
        for (int it = part; it < items; it += N_PROD_WARPS / 4) {
            const int kg = it / groups, g = it - kg * groups;
            uint32_t r[16];
            float acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0.f;
            if (grouped) {
                tc_ld16(lane_base + (uint32_t)(kg * p.Npad + 16 * g), r);
                tc_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] = __uint_as_float(r[e]);
            } else {
                for (int pass = 1; pass >= 0; --pass)            // cross-term blocks first (small), then the main products
                    for (int pr = 0; pr < p.n_pairs; ++pr) {
                        tc_ld16(lane_base + (uint32_t)((2 * pr + pass) * p.Npad + 16 * g), r);
                        tc_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(r[e]);
                    }
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = acc[e] * inv_a * inv_b;
            if (p.epi_res && m < m_end) {
                const float* rsrc = p.epi_res + m * p.epi_ld + 16 * g;
                if (16 * g + 15 < p.N) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 r4 = __ldg(reinterpret_cast<const float4*>(rsrc) + c4);
                        acc[4 * c4] += r4.x; acc[4 * c4 + 1] += r4.y; acc[4 * c4 + 2] += r4.z; acc[4 * c4 + 3] += r4.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (16 * g + e < p.N) acc[e] += __ldg(rsrc + e);
                }
            }
            if (p.epi_act && m < m_end) {
                // modReLU on the 8 complex values of this piece: y = relu(|z| + b_c) z / |z|, origin entries passed through
                float* adst = p.epi_act + m * p.epi_ld + 16 * g;
                float a[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = 16 * g + 2 * j;
                    const float zx = acc[2 * j], zy = acc[2 * j + 1];
                    const bool origin = (fabsf(zx) < 1e-7f) && (fabsf(zy) < 1e-7f);
                    const float n2 = zx * zx + zy * zy;
                    const float ri = rsqrtf(n2);
                    const float mod = fmaxf(n2 * ri + (n + 1 < p.N ? __ldg(p.epi_bias + (n >> 1)) : 0.f), 0.f);     // |act| of a non-origin entry
                    const float sc = mod * ri;
                    a[2 * j] = origin ? zx : sc * zx;
                    a[2 * j + 1] = origin ? zy : sc * zy;
                    act_mx = fmaxf(act_mx, origin ? 2e-7f : mod);
                }
                if (16 * g + 15 < p.N) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4)
                        *reinterpret_cast<float4*>(adst + 4 * c4) = make_float4(a[4 * c4], a[4 * c4 + 1], a[4 * c4 + 2], a[4 * c4 + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (16 * g + e < p.N) adst[e] = a[e];
                }
            }
            if (m < m_end) {
                float* dst = C + m * p.ldc + (int64_t)kg * p.N + 16 * g;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int n = 16 * g + 4 * c4;
                    if (n + 3 < p.N) {
                        *reinterpret_cast<float4*>(dst + 4 * c4) = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + e < p.N) dst[4 * c4 + e] = acc[4 * c4 + e];
                    }
                }
            }
        }
        if (p.epi_bound) {
            const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(act_mx * 1.000001f));
            if (lane == 0 && w > *reinterpret_cast<volatile uint32_t*>(p.epi_bound)) atomicMax(p.epi_bound, w);
        }
        }
    } else if (warp == N_PROD_WARPS) {
        // ------------------------------------------------------------------ MMA issuer (one thread; counters and pre-built
        // descriptors only — see gemm_tc.cu)
        if (lane == 0) {
            const uint32_t npad = (uint32_t)p.Npad;
            const uint32_t idesc1 = make_idesc_f16(p.Npad), idesc2 = 2 * p.Npad <= 256 ? make_idesc_f16(2 * p.Npad) : 0u;
            const uint64_t a_hi_d = make_desc_k_sw128(a_hi0), a_lo_d = make_desc_k_sw128(a_lo0);
            const uint64_t b_hi_d = make_desc_k_sw128(b0), b_lo_d = make_desc_k_sw128(b0 + b_plane);
            const uint64_t a_step = (uint64_t)(A_PLANE >> 4), b_step = (uint64_t)((2 * b_plane) >> 4);
            const uint32_t n_pairs = (uint32_t)p.n_pairs;
            const bool merged = 2 * p.Npad <= 256;                  // one MMA of width 2 Npad for [main | cross], else two of width Npad
            uint32_t s = 0, ph = 0;
            uint32_t pr = 0, d_pair = tmem_d, first = n_pairs;       // the first n_pairs k-steps overwrite their pair
            const bool grouped = p.kgroups > 1;
            uint32_t g_left = (uint32_t)p.cpg, d_grp = tmem_d;
            for (int kc = 0; kc < n_kc; ++kc) {
                if (!PACKED) mbar_wait(full_a(s), ph);
                mbar_wait(full_b(s), ph);
                tc_fence_after();
                const uint64_t a_hi = a_hi_d + s * a_step, a_lo = a_lo_d + s * a_step;
                const uint64_t b_hi = b_hi_d + s * b_step, b_lo = b_lo_d + s * b_step;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);   // 16 fp16 = 32 B = 2 x 16 B along K inside the swizzle row
                    if (grouped) {
                        tc_mma_f16(d_grp, a_hi + adv, b_hi + adv, idesc1, (ks == 0 && g_left == (uint32_t)p.cpg) ? 0u : 1u);
                        tc_mma_f16(d_grp, a_lo + adv, b_hi + adv, idesc1, 1u);
                        tc_mma_f16(d_grp, a_hi + adv, b_lo + adv, idesc1, 1u);
                        continue;
                    }
                    if (merged) {
                        tc_mma_f16(d_pair, a_hi + adv, b_hi + adv, idesc2, first ? 0u : 1u);    // [main | cross] (+)= A_hi [B_hi | B_lo]
                    } else {
                        tc_mma_f16(d_pair, a_hi + adv, b_hi + adv, idesc1, first ? 0u : 1u);           // main (+)= A_hi B_hi
                        tc_mma_f16(d_pair + npad, a_hi + adv, b_lo + adv, idesc1, first ? 0u : 1u);    // cross (+)= A_hi B_lo
                    }
                    tc_mma_f16(d_pair + npad, a_lo + adv, b_hi + adv, idesc1, 1u);          // cross += A_lo B_hi
                    if (first) --first;
                    if (++pr == n_pairs) { pr = 0; d_pair = tmem_d; } else d_pair += 2 * npad;
                }
                if (grouped && --g_left == 0) { g_left = (uint32_t)p.cpg; d_grp += npad; }
                tc_commit(empty(s));      // frees the stage once these MMAs have read it
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
            tc_commit(tmem_full);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ B loader (one thread, TMA bulk copies)
        if (lane == 0) {
            const uint32_t bytes = 2u * b_plane;
            const int64_t src_step = (int64_t)2 * p.Npad * KC;
            const __half* src = Bp + (int64_t)kc_begin * src_step;
            const uint8_t* a_src = reinterpret_cast<const uint8_t*>(p.A) + tile_x * p.a_tile_stride +
                                   (int64_t)kc_begin * PK_BLOCK_BYTES;
            uint32_t s = 0, ph = 1;
            for (int kc = 0; kc < n_kc; ++kc) {
                mbar_wait(empty(s), ph);
                mbar_expect_tx(full_b(s), PACKED ? bytes + PK_BLOCK_BYTES : bytes);
                bulk_copy_g2s(b0 + s * 2 * b_plane, src, bytes, full_b(s));
                if (PACKED) {
                    bulk_copy_g2s(a_hi0 + s * A_PLANE, a_src, PK_PLANE_BYTES, full_b(s));
                    bulk_copy_g2s(a_lo0 + s * A_PLANE, a_src + PK_PLANE_BYTES, PK_PLANE_BYTES, full_b(s));
                    a_src += PK_BLOCK_BYTES;
                }
                src += src_step;
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == N_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------ NN, short K
// C[M x N] = A B for K <= 128 (TangentLin: K = 2 Ci): the whole contraction of a row tile is one or two chunks, so the
// per-CTA fixed costs of k_gemm_h_nn (barrier set-up, TMEM allocation, the first B fetch, an un-overlapped epilogue) would be
// most of its time.  Persistent form: one CTA per SM walks over row tiles; the packed B operand (<= 2 chunks) is fetched ONCE
// and stays in shared memory; two operand stages and TWO accumulator sets in TMEM, so that
//   warps 0-15  (producers) load / split / store the fp32 rows of tile t+1,
//   warp  16    (one thread) issues the MMAs of tile t+1 as soon as its stage is full and an accumulator set is free,
//   warps 17-20 (epilogue, one per TMEM lane quarter) drain tile t — all at the same time.
constexpr int SK_PROD_WARPS = 16;
constexpr int SK_PROD = SK_PROD_WARPS * 32;
constexpr int SK_THREADS = (SK_PROD_WARPS + 1 + 4) * 32;
constexpr int SK_UNITS = 2;          // 16-byte output units per producer thread and chunk (128 rows x 8 units / 512 threads)

struct ParamsSK {
    const float* A;
    const __half* Bp;          // packed B (k_pack_b_h): [chunk][plane][Npad][64 fp16]
    float* C;
    const float *amax_a, *amax_b;
    int64_t M, K, lda, ldc;
    int N, Npad, nchunks, wide;
    int rows_per_tile;         // <= BM, multiple of 8
    int64_t tiles;
    uint32_t tmem_cols;
};

__global__ void __launch_bounds__(SK_THREADS, 1) k_gemm_h_nn_small(const ParamsSK p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int NC = p.nchunks;
    const uint32_t b_plane = (uint32_t)p.Npad * 128u;
    const uint32_t a_stage = (uint32_t)NC * 2u * A_PLANE;           // [chunk][hi | lo]
    // layout: B[NC][hi | lo] | A[2 stages] | barriers
    const uint32_t b0 = base, a0 = base + (uint32_t)NC * 2u * b_plane;
    const uint32_t bars = a0 + 2u * a_stage;
    const uint32_t b_full = bars;
    auto a_full = [&](int s) { return bars + 8u * (1 + s); };
    auto a_empty = [&](int s) { return bars + 8u * (3 + s); };
    auto t_full = [&](int s) { return bars + 8u * (5 + s); };
    auto t_empty = [&](int s) { return bars + 8u * (7 + s); };
    const uint32_t tmem_slot = bars + 8u * 9;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(sm + (tmem_slot - base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(b_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(a_full(s), SK_PROD_WARPS);
            mbar_init(a_empty(s), 1);
            mbar_init(t_full(s), 1);
            mbar_init(t_empty(s), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == SK_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot_ptr;
    const uint32_t acc_cols = 2u * (uint32_t)p.Npad;                // [main | cross] of one accumulator set

    if (warp < SK_PROD_WARPS) {
        // ------------------------------------------------------------------ producers
        const int t = threadIdx.x;
        uint32_t off[SK_UNITS];
        int urow[SK_UNITS], ucol[SK_UNITS];
#pragma unroll
        for (int i = 0; i < SK_UNITS; ++i) {
            const int idx = t + SK_PROD * i;
            const int row = idx >> 3, j = idx & 7;
            urow[i] = row;
            ucol[i] = 8 * j;
            off[i] = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
        }
        const float s_a = scale_of(p.amax_a);
        uint8_t* const a_base = sm + (a0 - base);
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            const int st = it & 1;
            const int64_t m0 = tile * p.rows_per_tile;
            const int64_t m_end = min(p.M, m0 + p.rows_per_tile);
            F8 v[2][SK_UNITS];                                      // every load of the tile in flight before the first split
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c >= NC) break;
#pragma unroll
                for (int i = 0; i < SK_UNITS; ++i) {
                    const int64_t m = m0 + urow[i];
                    const int64_t k = (int64_t)c * KC + ucol[i];
                    zero8(v[c][i]);
                    if (m < m_end && k < p.K) {
                        const float* src = p.A + m * p.lda + k;
                        if (k + 8 <= p.K) {
                            if (p.wide) ld256(src, v[c][i]);
                            else ld2x128(src, v[c][i]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (k + e < p.K) v[c][i].v[e] = src[e];
                        }
                    }
                }
            }
            mbar_wait(a_empty(st), (uint32_t)(((it >> 1) & 1) ^ 1));
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c >= NC) break;
#pragma unroll
                for (int i = 0; i < SK_UNITS; ++i) {
                    uint4 hi, lo;
                    split8(v[c][i], s_a, hi, lo);
                    uint8_t* dst = a_base + st * a_stage + (uint32_t)c * 2u * A_PLANE + off[i];
                    *reinterpret_cast<uint4*>(dst) = hi;
                    *reinterpret_cast<uint4*>(dst + A_PLANE) = lo;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(st));
        }
    } else if (warp == SK_PROD_WARPS) {
        // ------------------------------------------------------------------ B fetch (once) + MMA issue (one thread)
        if (lane == 0) {
            const uint32_t b_bytes = (uint32_t)NC * 2u * b_plane;
            mbar_expect_tx(b_full, b_bytes);
            for (int c = 0; c < NC; ++c)
                bulk_copy_g2s(b0 + (uint32_t)c * 2u * b_plane, p.Bp + (int64_t)c * 2 * p.Npad * KC, 2u * b_plane, b_full);
            mbar_wait(b_full, 0);
            const uint32_t npad = (uint32_t)p.Npad;
            const bool merged = 2 * p.Npad <= 256;
            const uint32_t idesc1 = make_idesc_f16(p.Npad), idesc2 = merged ? make_idesc_f16(2 * p.Npad) : 0u;
            const int ksteps = (int)((p.K + 15) / 16);              // k-steps that hold data (the rest of the last chunk is zero)
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const int st = it & 1;
                const uint32_t ph = (uint32_t)((it >> 1) & 1);
                mbar_wait(t_empty(st), ph ^ 1u);                    // the epilogue has drained this accumulator set
                mbar_wait(a_full(st), ph);
                tc_fence_after();
                const uint32_t d = tmem_d + (uint32_t)st * acc_cols;
                uint32_t first = 1;
                for (int c = 0; c < NC; ++c) {
                    const uint32_t a_hi0 = a0 + st * a_stage + (uint32_t)c * 2u * A_PLANE;
                    const uint64_t a_hi = make_desc_k_sw128(a_hi0), a_lo = make_desc_k_sw128(a_hi0 + A_PLANE);
                    const uint64_t b_hi = make_desc_k_sw128(b0 + (uint32_t)c * 2u * b_plane);
                    const uint64_t b_lo = make_desc_k_sw128(b0 + (uint32_t)c * 2u * b_plane + b_plane);
                    for (int ks = 0; ks < KC / 16 && c * (KC / 16) + ks < ksteps; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);
                        if (merged) {
                            tc_mma_f16(d, a_hi + adv, b_hi + adv, idesc2, first ? 0u : 1u);
                        } else {
                            tc_mma_f16(d, a_hi + adv, b_hi + adv, idesc1, first ? 0u : 1u);
                            tc_mma_f16(d + npad, a_hi + adv, b_lo + adv, idesc1, first ? 0u : 1u);
                        }
                        tc_mma_f16(d + npad, a_lo + adv, b_hi + adv, idesc1, 1u);
                        first = 0;
                    }
                }
                tc_commit(a_empty(st));
                tc_commit(t_full(st));
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (4 warps: warp % 4 = TMEM lane quarter)
        const int q = warp & 3;
        const float inv = inv_scale_of(p.amax_a) * inv_scale_of(p.amax_b);
        const int groups = p.Npad / 16;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            const int st = it & 1;
            const int64_t m0 = tile * p.rows_per_tile;
            const int64_t m_end = min(p.M, m0 + p.rows_per_tile);
            const int64_t m = m0 + 32 * q + lane;
            mbar_wait(t_full(st), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t lane_base = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)st * acc_cols;
            for (int g = 0; g < groups; ++g) {
                uint32_t r0[16], r1[16];
                tc_ld16(lane_base + (uint32_t)(16 * g), r0);
                tc_ld16(lane_base + (uint32_t)(p.Npad + 16 * g), r1);
                tc_ld_wait();
                if (m < m_end) {
                    float* dst = p.C + m * p.ldc + 16 * g;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = (__uint_as_float(r1[4 * c4 + e]) + __uint_as_float(r0[4 * c4 + e])) * inv;
                        const int n = 16 * g + 4 * c4;
                        if (n + 3 < p.N) {
                            *reinterpret_cast<float4*>(dst + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (n + e < p.N) dst[4 * c4 + e] = o[e];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty(st));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SK_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
    }
}

// B[K x N] row-major (ldb) -> [batch][chunk][plane][Npad][64 fp16] with the 128B swizzle applied (16-byte unit u of row
// n stored at unit u ^ (n & 7)); hi = fp16(s b), lo = fp16(s b - hi); rows n >= N and k >= K are zero.  One thread per
// 16-byte piece (8 consecutive k of one column n); consecutive threads take consecutive n (coalesced reads of B rows).
__global__ void k_pack_b_h(const float* __restrict__ B, __half* __restrict__ Bp, int64_t K, int N, int Npad, int64_t ldb,
                           int nchunks, int64_t sb, int64_t bp_batch_stride, const float* __restrict__ amax_b) {
    const int64_t per = (int64_t)nchunks * Npad * 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const int batch = blockIdx.y;
    const int n = (int)(i % Npad);
    const int u = (int)((i / Npad) & 7);
    const int64_t c = i / ((int64_t)Npad * 8);
    const float s = scale_of(amax_b);
    F8 x;
    zero8(x);
    if (n < N) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int64_t k = c * KC + 8 * u + e;
            if (k < K) x.v[e] = B[batch * sb + k * ldb + n];
        }
    }
    uint4 hi, lo;
    split8(x, s, hi, lo);
    __half* dst = Bp + batch * bp_batch_stride + (c * 2) * (int64_t)Npad * KC + (int64_t)n * KC + ((u ^ (n & 7)) << 3);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + (int64_t)Npad * KC) = lo;
}

// The folded filter W (Co,Ci,R,M) complex straight into the packed B operand of the 2xFP16 NN kernels — what k_pack_w_fwd /
// k_pack_w_bwd (real embedding, api.cu) followed by k_absmax and k_pack_b_h produce, in ONE launch:
//   forward  (BWD = false): B[2k+a][2o+b], k = (r M + m) Ci + c:   a=0: (w.x, w.y)[b];   a=1: (-w.y, w.x)[b]
//   backward (BWD = true):  group m, B_m[2q+a][2c+b], q = r Co + o: a=0: (w.x, -w.y)[b];  a=1: (w.y, w.x)[b]   (conj(W))
// One thread per 16-byte piece (8 consecutive rows k2 of one column n).  amax: bound of max|W| (largest |component|; the
// complex modulus is fine); thread 0 copies it into the workspace's scale slot the contraction kernel reads.
template <bool BWD>
__global__ void k_pack_w_h(const float2* __restrict__ W, __half* __restrict__ Bp, int Ci, int Co, int R, int M, int Npad, int nchunks,
                           int64_t bp_group_stride, const float* __restrict__ amax, float* __restrict__ amax_slot) {
    const int64_t per = (int64_t)nchunks * Npad * 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int grp = blockIdx.y;                          // BWD: frequency m; forward: 0
    if (i == 0 && grp == 0 && amax_slot != amax) *amax_slot = __ldg(amax);
    if (i >= per) return;
    const int n = (int)(i % Npad);
    const int u = (int)((i / Npad) & 7);
    const int64_t c = i / ((int64_t)Npad * 8);
    const float s = scale_of(amax);
    const int N = BWD ? 2 * Ci : 2 * Co;
    const int64_t K2 = BWD ? 2 * (int64_t)R * Co : 2 * (int64_t)R * M * Ci;
    F8 x;
    zero8(x);
    if (n < N) {
        const int b = n & 1, nn = n >> 1;                // forward: nn = o; backward: nn = c
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int64_t k2 = c * KC + 8 * u + e;
            if (k2 >= K2) continue;
            const int a = (int)(k2 & 1);
            const int64_t k = k2 >> 1;
            float2 w;
            if (BWD) {
                const int o = (int)(k % Co), r = (int)(k / Co);
                w = W[(((int64_t)o * Ci + nn) * R + r) * M + grp];
                x.v[e] = a == 0 ? (b == 0 ? w.x : -w.y) : (b == 0 ? w.y : w.x);
            } else {
                const int ch = (int)(k % Ci), m = (int)((k / Ci) % M), r = (int)(k / ((int64_t)M * Ci));
                w = W[(((int64_t)nn * Ci + ch) * R + r) * M + m];
                x.v[e] = a == 0 ? (b == 0 ? w.x : w.y) : (b == 0 ? -w.y : w.x);
            }
        }
    }
    uint4 hi, lo;
    split8(x, s, hi, lo);
    __half* dst = Bp + grp * bp_group_stride + (c * 2) * (int64_t)Npad * KC + (int64_t)n * KC + ((u ^ (n & 7)) << 3);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + (int64_t)Npad * KC) = lo;
}

// ------------------------------------------------------------------------------------------------------------ TN
// P[Mr x N] = A^T B with A = [Kv x Mr] and B = [Kv x N] row-major: both operands MN-major.  A stage holds 64 vertices:
// A planes = 8 K-groups x 2 atoms (64 columns each) x 1 KB, B planes = 8 K-groups x nb_atoms x 1 KB.
struct ParamsTN {
    const float* A;     // [Kv x Mr], lda
    const __half* Bp;   // packed B: per 64-vertex chunk the hi image then the lo image (k_pack_b_h_tn)
    float* C;           // partials [split][Mr][N] or C itself when split == 1 (ldc)
    const float *amax_a, *amax_b;
    int64_t Mr, Kv, lda, ldc, k_per_split, part_stride;
    int N, Npad, nb_atoms, stages, n_main, wide;
    int64_t a_tile_stride;     // PACKED: bytes between consecutive 128-vertex tiles of the PK buffer
    int a_chunks;              // PACKED: 64-column chunks of ONE batch entry (= Mr / 64)
    // batch (blockIdx.y): entry b reads the columns [b*Mr, (b+1)*Mr) of A (fp32: a_batch_off floats; PACKED: a_chunks
    // chunks further), its own packed B operand and writes its own [Mr x N] block of C / of every split's partials
    int64_t a_batch_off, bp_batch_stride, c_batch_stride;
    uint32_t tmem_cols;
};

// byte offset of the 16-byte piece (vertex v of the chunk, columns 8*f8 .. 8*f8+7) in an MN-major SWIZZLE_128B plane
__host__ __device__ __forceinline__ uint32_t tn_off_h(int v, int f8, int atoms) {
    return (uint32_t)((((v >> 3) * atoms + (f8 >> 3)) << 10) + ((v & 7) << 7) + ((((f8 & 7) ^ (v & 7))) << 4));
}

// PACKED: A is the PK buffer of [Kv x Mr] (rows = vertices).  A stage = 64 vertices x 128 columns = for each of the two
// 64-column chunks and each plane the 8 KB half (vertex rows 0-63 or 64-127) of that chunk's tile image: as an MN-major
// operand the image is 8 K-groups (8 vertices, 1 KB atoms) per chunk, so LBO (between the two column atoms) = 8 KB and
// SBO (between K-groups) = 1 KB.  The loader thread issues the 4 bulk copies; the producer warps only run the epilogue.
template <bool PACKED>
__global__ void __launch_bounds__(THREADS, 1) k_gemm_h_tn(const ParamsTN p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int S = p.stages;
    const uint32_t b_plane = (uint32_t)KV * (uint32_t)p.nb_atoms * 128u;
    const uint32_t a_hi0 = base, a_lo0 = base + S * A_PLANE, b0 = base + 2 * S * A_PLANE;
    const uint32_t bars = b0 + S * 2 * b_plane;
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto full_b = [&](int s) { return bars + 8u * (S + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * S + s); };
    const uint32_t tmem_full = bars + 8u * (3 * S);
    const uint32_t tmem_slot = bars + 8u * (3 * S + 1);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(sm + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int split = blockIdx.z;
    const int batch = blockIdx.y;
    const int64_t kb = (int64_t)split * p.k_per_split;       // multiple of KV
    const int64_t ke = min(p.Kv, kb + p.k_per_split);
    const int nchunks = kb < ke ? (int)((ke - kb + KV - 1) / KV) : 0;
    float* C = p.C + (int64_t)split * p.part_stride + (int64_t)batch * p.c_batch_stride;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_a(s), N_PROD_WARPS);
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == N_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot_ptr;

    if (warp < N_PROD_WARPS) {
        if (!PACKED) {
        const int t = threadIdx.x;
        const float* a_src[UN];
        uint32_t a_off[UN];
        int a_v[UN], a_cnt[UN];                            // vertex within the stage, valid columns (0..8)
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            const int idx = t + N_PROD * i;
            const int v = idx >> 4, f8 = idx & 15;          // 16 units of 8 columns per vertex
            const int64_t m = m0 + 8 * f8;
            a_v[i] = v;
            a_cnt[i] = (int)max((int64_t)0, min((int64_t)8, p.Mr - m));
            a_src[i] = p.A + (int64_t)batch * p.a_batch_off + (kb + v) * p.lda + m;
            a_off[i] = tn_off_h(v, f8, 2);
        }
        const bool cols_full = (m0 + BM <= p.Mr);
        const bool wide = p.wide != 0;
        const int64_t a_step = (int64_t)KV * p.lda;
        const float s_a = scale_of(p.amax_a);
        F8 va[PF][UN];
        auto issue = [&](int kc, F8(&da)[UN]) {
            const int64_t v0 = kb + (int64_t)kc * KV;
            if (cols_full && v0 + KV <= ke) {            // fast path: no guards (CTA-uniform)
#pragma unroll
                for (int i = 0; i < UN; ++i) {
                    if (wide) ld256(a_src[i], da[i]);
                    else ld2x128(a_src[i], da[i]);
                    a_src[i] += a_step;
                }
            } else {
#pragma unroll
                for (int i = 0; i < UN; ++i) {
                    zero8(da[i]);
                    if (v0 + a_v[i] < ke) {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (e < a_cnt[i]) da[i].v[e] = a_src[i][e];
                    }
                    a_src[i] += a_step;
                }
            }
        };
        uint8_t* const ahi = sm + (a_hi0 - base);
        uint8_t* const alo = sm + (a_lo0 - base);
        uint32_t ps = 0, pph = 1;
        auto commit = [&](const F8(&sa)[UN]) {
            const uint32_t s = ps;
            mbar_wait(empty(s), pph);
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                uint4 hi, lo;
                split8(sa[i], s_a, hi, lo);
                *reinterpret_cast<uint4*>(ahi + s * A_PLANE + a_off[i]) = hi;
                *reinterpret_cast<uint4*>(alo + s * A_PLANE + a_off[i]) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(s));
            if (++ps == (uint32_t)S) { ps = 0; pph ^= 1u; }
        };
#pragma unroll
        for (int u = 0; u < PF - 1; ++u)
            if (u < nchunks) issue(u, va[u]);
        for (int kc = 0; kc < nchunks; kc += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int k = kc + u;
                if (k < nchunks) {
                    if (k + PF - 1 < nchunks) issue(k + PF - 1, va[(u + PF - 1) % PF]);
                    commit(va[u]);
                }
            }
        }
        }
        // epilogue
        const int q = warp & 3, part = warp >> 2;
        const int64_t m = m0 + 32 * q + lane;
        const int groups = p.Npad / 16;
        const int n_acc = p.n_main + 1;
        const float inv_a = inv_scale_of(p.amax_a), inv_b = inv_scale_of(p.amax_b);
        if (nchunks > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        for (int g = part; g < groups; g += N_PROD_WARPS / 4) {
            uint32_t r[16];
            float acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0.f;
            if (nchunks > 0) {
                for (int a = n_acc - 1; a >= 0; --a) {       // cross-term accumulator (last) first
                    tc_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(a * p.Npad + 16 * g), r);
                    tc_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(r[e]);
                }
            }
            if (m < p.Mr) {
                float* dst = C + m * p.ldc + 16 * g;
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (16 * g + e < p.N) dst[e] = acc[e] * inv_a * inv_b;
            }
        }
    } else if (warp == N_PROD_WARPS) {
        if (lane == 0 && nchunks > 0) {
            // D = f32, A = B = f16, both MN-major (bits 15, 16), M = 128, N = Npad
            const uint32_t idesc = make_idesc_f16(p.Npad) | (1u << 15) | (1u << 16);
            const uint32_t sbo_a = PACKED ? 1024 : 2 * 1024, sbo_b = (uint32_t)p.nb_atoms * 1024;   // between 8-vertex K groups
            const uint32_t lbo_a = PACKED ? 8 * 1024 : 1024;                                       // between the two column atoms
            const uint64_t a_hi_d = make_desc_mn_sw128_h(a_hi0, lbo_a, sbo_a), a_lo_d = make_desc_mn_sw128_h(a_lo0, lbo_a, sbo_a);
            const uint64_t b_hi_d = make_desc_mn_sw128_h(b0, 1024, sbo_b), b_lo_d = make_desc_mn_sw128_h(b0 + b_plane, 1024, sbo_b);
            const uint64_t a_stage = (uint64_t)(A_PLANE >> 4), b_stage = (uint64_t)((2 * b_plane) >> 4);
            const uint64_t a_kg = (uint64_t)((2 * sbo_a) >> 4), b_kg = (uint64_t)((2 * sbo_b) >> 4);   // 16 vertices
            const uint32_t d_x = tmem_d + (uint32_t)(p.n_main * p.Npad);
            const uint32_t n_main = (uint32_t)p.n_main, npad = (uint32_t)p.Npad;
            uint32_t s = 0, ph = 0, acc = 0, d_main = tmem_d, first = n_main, x_acc = 0;
            for (int kc = 0; kc < nchunks; ++kc) {
                if (!PACKED) mbar_wait(full_a(s), ph);
                mbar_wait(full_b(s), ph);
                tc_fence_after();
                uint64_t a_hi = a_hi_d + s * a_stage, a_lo = a_lo_d + s * a_stage;
                uint64_t b_hi = b_hi_d + s * b_stage, b_lo = b_lo_d + s * b_stage;
#pragma unroll
                for (int kg = 0; kg < KV / 16; ++kg) {
                    tc_mma_f16(d_main, a_hi, b_hi, idesc, first ? 0u : 1u);
                    if (first) --first;
                    if (++acc == n_main) { acc = 0; d_main = tmem_d; } else d_main += npad;
                    tc_mma_f16(d_x, a_lo, b_hi, idesc, x_acc);
                    tc_mma_f16(d_x, a_hi, b_lo, idesc, 1u);
                    x_acc = 1u;
                    a_hi += a_kg; a_lo += a_kg; b_hi += b_kg; b_lo += b_kg;
                }
                tc_commit(empty(s));
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
            tc_commit(tmem_full);
        }
        __syncwarp();
    } else {
        if (lane == 0) {
            const uint32_t bytes = 2u * b_plane;
            const int64_t img = (int64_t)b_plane;                          // fp16 elements per chunk image (hi + lo planes)
            const __half* src = p.Bp + (int64_t)batch * p.bp_batch_stride + (kb / KV) * img;
            // PACKED: column chunks c0, c0+1 of this CTA's 128 columns (the second may lie past the matrix end: its D
            // rows are never stored, so it is simply not loaded)
            const int c0 = (int)(m0 / PK_COLS);
            const int n_ca = (c0 + 1 < p.a_chunks) ? 2 : 1;
            const uint8_t* a_base = reinterpret_cast<const uint8_t*>(p.A) + ((int64_t)batch * p.a_chunks + c0) * PK_BLOCK_BYTES;
            uint32_t s = 0, ph = 1;
            for (int kc = 0; kc < nchunks; ++kc) {
                mbar_wait(empty(s), ph);
                mbar_expect_tx(full_b(s), PACKED ? bytes + (uint32_t)n_ca * 2u * 8192u : bytes);
                bulk_copy_g2s(b0 + s * 2 * b_plane, src, bytes, full_b(s));
                if (PACKED) {
                    const int64_t v0 = kb + (int64_t)kc * KV;                       // first vertex of the stage (multiple of 64)
                    const uint8_t* a_src = a_base + (v0 >> 7) * p.a_tile_stride + ((v0 >> 6) & 1) * 8192;
                    for (int j = 0; j < n_ca; ++j) {
                        bulk_copy_g2s(a_hi0 + s * A_PLANE + j * 8192u, a_src + (int64_t)j * PK_BLOCK_BYTES, 8192u, full_b(s));
                        bulk_copy_g2s(a_lo0 + s * A_PLANE + j * 8192u, a_src + (int64_t)j * PK_BLOCK_BYTES + PK_PLANE_BYTES, 8192u,
                                      full_b(s));
                    }
                }
                src += img;
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == N_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
    }
}

// B[Kv x N] row-major (ldb) -> per 64-vertex chunk the MN-major swizzled image, hi plane then lo plane; vertices >= Kv
// and columns >= N are zero.  One thread per 16-byte piece.
__global__ void k_pack_b_h_tn(const float* __restrict__ B, __half* __restrict__ Bp, int64_t Kv, int N, int atoms, int64_t ldb,
                              int64_t nchunks, const float* __restrict__ amax_b) {
    const int u_per_row = atoms * 8;
    const int64_t per_chunk = (int64_t)KV * u_per_row;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks * per_chunk) return;
    const int64_t c = i / per_chunk;
    const int r = (int)(i - c * per_chunk);
    const int v = r / u_per_row, f8 = r - v * u_per_row;
    const int64_t vg = c * KV + v;
    const float s = scale_of(amax_b);
    F8 x;
    zero8(x);
    if (vg < Kv) {
        const float* q = B + vg * ldb + 8 * f8;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (8 * f8 + e < N) x.v[e] = q[e];
    }
    uint4 hi, lo;
    split8(x, s, hi, lo);
    const int64_t plane = (int64_t)KV * atoms * 64;                       // fp16 elements per plane
    __half* dst = Bp + c * 2 * plane + tn_off_h(v, f8, atoms) / 2;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + plane) = lo;
}


// ---- weight gradient from G (no contrib in the backward):  gW[o,c,r,m] = sum_j conj(xhat[j,c,m]) G[j,m,r,o]
// (identity: sum_n conj(contrib[n,c,r,m]) gy[n,o] regrouped by source vertex, nn/field_conv.py:130-137 under autograd).
// Per frequency m one TN product  P_m[2RCo x 2Ci] = G_m^T Xh_m  with Xh_m the real view of xhat[:, :, m]: the B operand
// of batch entry m.  This kernel builds all M packed operands straight from x (gauge alignment xhat = x conj(u)^m,
// utils/field.py:40-48 + nn/field_conv.py:128-130) — one thread per 16-byte piece (4 complex channels of one vertex).
template <int BL>
__global__ void __launch_bounds__(256) k_pack_xhat_tn(const float2* __restrict__ x, __half* __restrict__ Bp, int64_t Kv, int Ci,
                                                      int atoms, int64_t nchunks, int64_t bp_batch_stride,
                                                      const float* __restrict__ amax_b) {
    constexpr int M = 2 * BL + 1;
    const int u_per_row = atoms * 8;
    const int64_t per_chunk = (int64_t)KV * u_per_row;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks * per_chunk) return;
    const int64_t c = i / per_chunk;
    const int r = (int)(i - c * per_chunk);
    const int v = r / u_per_row, f8 = r - v * u_per_row;
    const int64_t vg = c * KV + v;
    const float s = scale_of(amax_b);
    F8 xm[M];
#pragma unroll
    for (int m = 0; m < M; ++m) zero8(xm[m]);
    if (vg < Kv) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ch = 4 * f8 + e;
            if (ch < Ci) {
                const float2 z = x[vg * Ci + ch];
                const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
                const float ri = rsqrtf(z.x * z.x + z.y * z.y);
                const float2 u = origin ? make_float2(1.f, 0.f) : make_float2(z.x * ri, z.y * ri);
                float2 up = z, dn = z;
                xm[BL].v[2 * e] = z.x; xm[BL].v[2 * e + 1] = z.y;
#pragma unroll
                for (int k = 1; k <= BL; ++k) {
                    up = cmul_conj(up, u);        // z conj(u)^k
                    dn = cmul(dn, u);             // z u^k = z conj(u)^(-k)
                    xm[BL + k].v[2 * e] = up.x; xm[BL + k].v[2 * e + 1] = up.y;
                    xm[BL - k].v[2 * e] = dn.x; xm[BL - k].v[2 * e + 1] = dn.y;
                }
            }
        }
    }
    const int64_t plane = (int64_t)KV * atoms * 64;                       // fp16 elements per plane
    __half* dst = Bp + c * 2 * plane + tn_off_h(v, f8, atoms) / 2;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        uint4 hi, lo;
        split8(xm[m], s, hi, lo);
        *reinterpret_cast<uint4*>(dst + m * bp_batch_stride) = hi;
        *reinterpret_cast<uint4*>(dst + m * bp_batch_stride + plane) = lo;
    }
}

// out (bit pattern of a non-negative float) = max(out, max_i |z_i| * (1 + 2^-20)): the bound on every real component of
// xhat = z conj(u)^m.  out is pre-zeroed.
__global__ void __launch_bounds__(256) k_absmax_modulus(const float2* __restrict__ z, int64_t n, uint32_t* __restrict__ out) {
    float m2 = 0.f;      // max |z|^2; the square root is monotone, one per thread at the end
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    if (((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(z) & 15u) == 0)) {
        const float4* q = reinterpret_cast<const float4*>(z);
        const int64_t n2 = n >> 1;
        auto take = [&](const float4 v) { m2 = fmaxf(m2, fmaxf(v.x * v.x + v.y * v.y, v.z * v.z + v.w * v.w)); };
        int64_t i = t0;
        for (; i + 3 * nt < n2; i += 4 * nt) {
            const float4 v0 = __ldg(q + i), v1 = __ldg(q + i + nt), v2 = __ldg(q + i + 2 * nt), v3 = __ldg(q + i + 3 * nt);
            take(v0); take(v1); take(v2); take(v3);
        }
        for (; i < n2; i += nt) take(__ldg(q + i));
    } else {
        for (int64_t i = t0; i < n; i += nt) {
            const float2 v = __ldg(z + i);
            m2 = fmaxf(m2, v.x * v.x + v.y * v.y);
        }
    }
    float mx = sqrtf(m2) * 1.000001f;
    const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
    if ((threadIdx.x & 31) == 0 && w > *reinterpret_cast<volatile uint32_t*>(out)) atomicMax(out, w);
}

}  // namespace th

// ------------------------------------------------------------------------------------------------------------ host side
constexpr int64_t H_MAX_ACC_MMAS = 400;

// NN: column-chunk width, number of (main, cross) accumulator pairs the k-steps of 16 reals are dealt over, and the number of
// K ranges (split-K CTAs) for `ksteps` k-steps (0: not feasible).  N <= 128: one CTA per row tile, up to 512 / (2 Npad) pairs.
// Wider outputs: chunks of up to 256 columns, ONE pair filling the 512 TMEM columns, and the K range cut so that no
// accumulator takes more than 400 accumulating MMAs — A is read once per 256 columns instead of once per 128.
int gemm_h_plan_nn(int N, int64_t ksteps, int* n_pairs_out, int* split_out) {
    if (N <= 0) return 0;
    if (split_out) *split_out = 1;
    const int64_t need = (ksteps + H_MAX_ACC_MMAS - 1) / H_MAX_ACC_MMAS;
    if (N > 128 && split_out) {
        const int nchunks_col = (N + 255) / 256;
        int nc = (N + nchunks_col - 1) / nchunks_col;
        nc = (nc + 15) / 16 * 16;
        if (nc > 256) nc = 256;
        *n_pairs_out = 1;
        *split_out = (int)(need < 1 ? 1 : need);
        return nc;
    }
    const int widths[3] = {128, 64, 32};
    for (int w : widths) {
        const int nc = N < w ? N : w;
        const int npad = (nc + 15) / 16 * 16;
        const int avail = 512 / (2 * npad);
        if (avail >= 1 && need <= avail) {
            int n_pairs = avail < 2 ? avail : 2;
            if (n_pairs < need) n_pairs = (int)need;
            if (n_pairs > ksteps) n_pairs = (int)(ksteps < 1 ? 1 : ksteps);
            *n_pairs_out = n_pairs;
            return nc;
        }
    }
    return 0;
}

static size_t packed_b_bytes(int N, int64_t K, int batch) {
    const int npad = (N + 15) / 16 * 16;
    const int64_t nchunks = (K + th::KC - 1) / th::KC;
    return (size_t)batch * nchunks * 2 * npad * th::KC * 2;
}

// workspace: [256 B of scalars: max|B|, max|A| when computed here][packed B]
size_t gemm_h_ws_bytes(int N, int64_t K, int batch) { return 256 + align_up(packed_b_bytes(N, K, batch), 256) + 256; }

float* gemm_h_amax_slot(void* ws) { return static_cast<float*>(ws) + 1; }

int launch_absmax_f32(const float* p, int64_t rows, int cols, int64_t ld, int batch, int64_t stride, float* out, cudaStream_t st) {
    return th::launch_absmax(p, rows, cols, ld, batch, stride, out, st);
}

int launch_bound_modulus(const float* z, int64_t n, float* out, cudaStream_t st) {
    if (cudaMemsetAsync(out, 0, 4, st) != cudaSuccess) {
        set_error("bound_modulus: cudaMemsetAsync failed");
        return FCB_E_CUDA;
    }
    if (n <= 0) return FCB_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    FCB_LAUNCH("absmax_mod", st, th::k_absmax_modulus<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float2*>(z), n,
                                                                                   reinterpret_cast<uint32_t*>(out)));
    return FCB_OK;
}

// Fills the B side of a launch_gemm_h_nn workspace (scale slot + packed operand) from the folded filter: dir 0 = forward
// operand [2K x 2Co], dir 1 = the M grouped backward operands [2RCo x 2Ci].  w_bound (nullable): device float >= max|W|.
int launch_pack_w_h(const float* W, int dir, int Ci, int Co, int R, int M, const float* w_bound, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
    const int N = dir ? 2 * Ci : 2 * Co;
    const int64_t K = dir ? 2 * (int64_t)R * Co : 2 * (int64_t)R * M * Ci;
    const int groups = dir ? M : 1;
    FCB_REQUIRE(W && ws && ws_bytes >= gemm_h_ws_bytes(N, K, groups), FCB_E_WORKSPACE, "pack_w_h: workspace too small");
    const int npad = (N + 15) / 16 * 16;
    const int nchunks = (int)((K + th::KC - 1) / th::KC);
    float* slot = static_cast<float*>(ws);
    __half* Bp = reinterpret_cast<__half*>(static_cast<char*>(ws) + 256);
    const float* amax = w_bound;
    if (!amax) {
        int rc = th::launch_absmax(W, (int64_t)Co * Ci * R * M, 2, 2, 1, 0, slot, st);
        if (rc) return rc;
        amax = slot;
    }
    const int64_t per = (int64_t)nchunks * npad * 8;
    dim3 grid((unsigned)((per + 255) / 256), (unsigned)groups);
    const int64_t stride = (int64_t)nchunks * 2 * npad * th::KC;
    const float2* W2 = reinterpret_cast<const float2*>(W);
    if (dir) FCB_LAUNCH("pack_w_h", st, (th::k_pack_w_h<true><<<grid, 256, 0, st>>>(W2, Bp, Ci, Co, R, M, npad, nchunks, stride, amax, slot)));
    else FCB_LAUNCH("pack_w_h", st, (th::k_pack_w_h<false><<<grid, 256, 0, st>>>(W2, Bp, Ci, Co, R, M, npad, nchunks, stride, amax, slot)));
    return FCB_OK;
}

int launch_gemm_h_nn(const float* A, const float* B, float* C, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb,
                     int64_t ldc, int batch, int64_t sa, int64_t sb, int64_t sc, int n_pairs, int kgroups,
                     const float* amax_a, void* ws, size_t ws_bytes, int a_packed, cudaStream_t st, int split_k, float* parts,
                     const GemmEpilogue* epi, const float* sa_x, float* sa_gx, int b_prepacked, const float* b_bound) {
    FCB_REQUIRE(A && (B || b_prepacked) && C && ws && amax_a, FCB_E_ARG, "gemm_h: null pointer");
    FCB_REQUIRE(M >= 0 && N > 0 && K > 0 && batch >= 1 && batch <= 65535, FCB_E_ARG, "gemm_h: bad sizes");
    FCB_REQUIRE(N <= 256, FCB_E_UNSUPPORTED, "gemm_h: N=%d > 256 not supported by one accumulator pair", N);
    FCB_REQUIRE(split_k >= 1 && split_k <= 65535 && (split_k == 1 || (parts && kgroups == 1)), FCB_E_ARG, "gemm_h: bad split-K arguments");
    if (a_packed) {
        FCB_REQUIRE(batch == 1 && (K % th::KC) == 0 && (reinterpret_cast<uintptr_t>(A) & 127u) == 0, FCB_E_ARG,
                    "gemm_h: a packed A operand needs batch == 1, K %% 64 == 0 and a 128-byte aligned buffer");
        lda = 4; sa = 0;
    }
    FCB_REQUIRE((lda % 4) == 0 && (ldc % 4) == 0 && (sa % 4) == 0 && (sc % 4) == 0 && aligned16(A) && aligned16(C),
                FCB_E_ALIGN, "gemm_h: A/C leading dimensions and strides must be multiples of 4 floats, 16-byte aligned");
    FCB_REQUIRE(kgroups >= 1 && (kgroups == 1 || (batch == 1 && K % th::KC == 0)), FCB_E_ARG, "gemm_h: bad k-group shape");
    FCB_REQUIRE(ws_bytes >= gemm_h_ws_bytes(N, K, batch * kgroups), FCB_E_WORKSPACE, "gemm_h: workspace too small");
    if (M == 0) return FCB_OK;
    const int npad = (N + 15) / 16 * 16;
    const int nchunks = (int)((K + th::KC - 1) / th::KC);
    const float* amax_b = static_cast<float*>(ws);
    __half* Bp = reinterpret_cast<__half*>(static_cast<char*>(ws) + 256);
    const int64_t bp_stride = (int64_t)nchunks * 2 * npad * th::KC;
    if (!b_prepacked) {      // b_prepacked: launch_pack_w_h already filled the scale slot and Bp of this workspace
        // one common scale for all batches / k-groups of B: the caller's bound, else a pass over B
        if (b_bound) {
            amax_b = b_bound;
        } else {
            const bool contiguous = (ldb == N) && (batch * kgroups == 1 || sb == K * (int64_t)N);
            int rc = contiguous ? th::launch_absmax(B, (int64_t)batch * kgroups * K, N, N, 1, 0, static_cast<float*>(ws), st)
                                : th::launch_absmax(B, K, N, ldb, batch * kgroups, sb, static_cast<float*>(ws), st);
            if (rc) return rc;
        }
        const int64_t per = (int64_t)nchunks * npad * 8;
        dim3 grid((unsigned)((per + 255) / 256), (unsigned)(batch * kgroups));
        FCB_LAUNCH("pack_b_h", st, th::k_pack_b_h<<<grid, 256, 0, st>>>(B, Bp, K, N, npad, ldb, nchunks, sb, bp_stride, amax_b));
    }
    // short K (TangentLin): persistent kernel with a resident B operand and two accumulator sets (k_gemm_h_nn_small).
    // FIELDCONV_B200_GEMM_SMALL=0 keeps the general kernel (A/B switch).
    static const bool small_on = [] { const char* e = getenv("FIELDCONV_B200_GEMM_SMALL"); return !e || atoi(e) != 0; }();
    if (small_on && !a_packed && nchunks <= 2 && batch == 1 && kgroups == 1 && split_k == 1 && !epi && !sa_gx && npad <= 128 &&
        M > th::BM) {
        th::ParamsSK q;
        q.A = A; q.Bp = Bp; q.C = C; q.amax_a = amax_a; q.amax_b = amax_b;
        q.M = M; q.K = K; q.lda = lda; q.ldc = ldc;
        q.N = N; q.Npad = npad; q.nchunks = nchunks;
        q.wide = ((lda % 8) == 0 && (reinterpret_cast<uintptr_t>(A) & 31u) == 0) ? 1 : 0;
        // equal row tiles that give every CTA of the persistent grid the same number of tiles
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t tiles128 = (M + th::BM - 1) / th::BM;
        const int64_t ctas = tiles128 < sms ? tiles128 : sms;
        const int64_t rounds = (tiles128 + ctas - 1) / ctas;
        int64_t rpt = ((M + rounds * ctas - 1) / (rounds * ctas) + 7) / 8 * 8;
        if (rpt > th::BM) rpt = th::BM;
        q.rows_per_tile = (int)rpt;
        q.tiles = (M + rpt - 1) / rpt;
        uint32_t cols = 32;
        while ((int)cols < 4 * npad) cols <<= 1;                    // two [main | cross] sets
        q.tmem_cols = cols;
        const size_t smem = (size_t)nchunks * 2 * npad * 128 + 2 * (size_t)nchunks * 2 * th::A_PLANE + 1024 + 8 * 10 + 64;
        static std::atomic<bool> sk_attr_dev[64];
        if (!(dev >= 0 && dev < 64 && sk_attr_dev[dev].load(std::memory_order_acquire))) {
            if (cudaFuncSetAttribute(th::k_gemm_h_nn_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
                set_error("gemm_h: cudaFuncSetAttribute failed");
                return FCB_E_CUDA;
            }
            if (dev >= 0 && dev < 64) sk_attr_dev[dev].store(true, std::memory_order_release);
        }
        const unsigned grid = (unsigned)(q.tiles < ctas ? q.tiles : ctas);
        FCB_LAUNCH("gemm_h_nn_small", st, (th::k_gemm_h_nn_small<<<grid, th::SK_THREADS, smem, st>>>(q)));
        return FCB_OK;
    }
    th::Params p;
    p.A = A; p.Bp = Bp; p.C = C;
    p.amax_a = amax_a; p.amax_b = amax_b;
    p.M = M; p.K = K; p.lda = lda; p.ldc = ldc; p.sa = sa; p.sc = sc;
    p.bp_batch_stride = bp_stride;
    p.N = N; p.Npad = npad; p.nchunks = nchunks * kgroups;
    p.kgroups = kgroups; p.cpg = nchunks;
    p.wide = ((lda % 8) == 0 && (sa % 8) == 0 && (reinterpret_cast<uintptr_t>(A) & 31u) == 0) ? 1 : 0;
    p.a_tile_stride = (int64_t)nchunks * kgroups * PK_BLOCK_BYTES;
    if (split_k > nchunks) split_k = nchunks;
    p.cps = (nchunks * kgroups + split_k - 1) / split_k;
    split_k = (nchunks * kgroups + p.cps - 1) / p.cps;          // no empty K range
    p.part_stride = 0;
    p.epi_res = nullptr; p.epi_bias = nullptr; p.epi_act = nullptr; p.epi_ld = 0; p.epi_bound = nullptr;
    p.sa_x = nullptr; p.sa_gx = nullptr; p.sa_band = 0;
    if (sa_gx) {
        FCB_REQUIRE(sa_x && kgroups >= 3 && kgroups <= 7 && (kgroups & 1) && (N % 8) == 0 && aligned16(sa_x) && aligned16(sa_gx), FCB_E_ARG,
                    "gemm_h: the fused softAngle epilogue needs the grouped product of band limit 1..3 and N %% 8 == 0");
        p.sa_x = sa_x; p.sa_gx = sa_gx; p.sa_band = (kgroups - 1) / 2;
    }
    if (epi) {
        FCB_REQUIRE(split_k == 1 && kgroups == 1 && batch == 1 && (N & 1) == 0, FCB_E_ARG, "gemm_h: the fused epilogue needs one un-split product");
        p.epi_res = epi->res; p.epi_bias = epi->bias; p.epi_act = epi->act; p.epi_ld = epi->ld;
        p.epi_bound = epi->act ? reinterpret_cast<uint32_t*>(epi->act_bound) : nullptr;
    }
    if (split_k > 1) {                                          // partials [split][batch][M][N]
        p.C = parts;
        p.ldc = N;
        p.sc = M * (int64_t)N;
        p.part_stride = (int64_t)batch * M * N;
    }
    int cols_needed;
    if (kgroups > 1) {
        p.K = K * kgroups;
        n_pairs = 1;
        cols_needed = npad * kgroups;
    } else {
        cols_needed = 2 * npad * n_pairs;
    }
    FCB_REQUIRE(n_pairs >= 1 && cols_needed <= 512, FCB_E_ARG, "gemm_h: accumulators do not fit TMEM");
    p.n_pairs = n_pairs;
    uint32_t cols = 32;
    while ((int)cols < cols_needed) cols <<= 1;
    p.tmem_cols = cols;
    const size_t stage_bytes = 2 * th::A_PLANE + 2 * (size_t)npad * 128;
    int stages = (int)((220 * 1024 - 2048) / stage_bytes);
    if (stages > 6) stages = 6;
    if (stages > p.cps) stages = p.cps < 1 ? 1 : p.cps;
    FCB_REQUIRE(stages >= 1, FCB_E_UNSUPPORTED, "gemm_h: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 1024 /*align slack*/ + 8 * (3 * stages + 2) + 64;
    // the opt-in shared-memory size is a per-device (per-context) function attribute: remember it per device
    static std::atomic<bool> attr_set_dev[64];
    std::atomic<bool> attr_unknown_dev{false};
    int attr_dev = 0;
    if (cudaGetDevice(&attr_dev) != cudaSuccess) attr_dev = -1;
    std::atomic<bool>& attr_set = (attr_dev >= 0 && attr_dev < 64) ? attr_set_dev[attr_dev] : attr_unknown_dev;
    if (!attr_set.load(std::memory_order_acquire)) {       // idempotent: racing threads at worst set the attribute twice
        cudaError_t e = cudaFuncSetAttribute(th::k_gemm_h_nn<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(th::k_gemm_h_nn<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(th::k_gemm_h_nn<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("gemm_h: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return FCB_E_CUDA;
        }
        attr_set.store(true, std::memory_order_release);
    }
    // fp32 operand: the row range is cut into equal tiles of <= 128 rows so that the grid fills whole waves of 148 CTAs (a
    // 128-row MMA on fewer valid rows costs tensor time the kernel has to spare; the DRAM-bound loads shrink with the rows).
    // A PK operand is tiled in 128-row images by its producer.
    static const bool ragged = [] { const char* e = getenv("FIELDCONV_B200_GEMM_RAGGED"); return !e || atoi(e) != 0; }();
    static const bool reverse = [] { const char* e = getenv("FIELDCONV_B200_GEMM_REVERSE"); return !e || atoi(e) != 0; }();
    p.reverse = reverse ? 1 : 0;
    p.rows_per_tile = th::BM;
    if (!a_packed && ragged && batch == 1 && split_k == 1) {
        const int64_t tiles = (M + th::BM - 1) / th::BM;
        const int64_t ctas = (tiles + 147) / 148 * 148;
        int64_t rpt = ((M + ctas - 1) / ctas + 7) / 8 * 8;
        if (tiles > 148 && rpt >= 96 && rpt < th::BM) p.rows_per_tile = (int)rpt;
    }
    dim3 grid((unsigned)((M + p.rows_per_tile - 1) / p.rows_per_tile), (unsigned)batch, (unsigned)split_k);
    // FIELDCONV_B200_GEMM_PAIRED=1: paired chunk loads in the fp32-operand producers (experiment switch, read once)
    static const bool paired = [] { const char* e = getenv("FIELDCONV_B200_GEMM_PAIRED"); return e && atoi(e) != 0; }();
    if (a_packed) FCB_LAUNCH("gemm_p_nn", st, (th::k_gemm_h_nn<true, false><<<grid, th::THREADS, smem, st>>>(p)));
    else if (paired) FCB_LAUNCH("gemm_h_nn", st, (th::k_gemm_h_nn<false, true><<<grid, th::THREADS, smem, st>>>(p)));
    else FCB_LAUNCH("gemm_h_nn", st, (th::k_gemm_h_nn<false, false><<<grid, th::THREADS, smem, st>>>(p)));
    if (split_k > 1) return launch_reduce_splits(parts, C, M, N, ldc, sc, batch, split_k, st);
    return FCB_OK;
}

// packed gy operand of the TN GEMM: [256 B scalars][hi + lo images of every 64-vertex chunk]
size_t gemm_h_tn_ws_bytes(int N, int64_t Kv) {
    const int npad = (N + 15) / 16 * 16;
    const int atoms = (npad + 63) / 64;
    const int64_t nchunks = (Kv + th::KV - 1) / th::KV;
    return 256 + align_up((size_t)nchunks * 2 * th::KV * atoms * 128, 256) + 256;
}

// The TN launch proper.  batch > 1: entry b takes the columns [b*Mr, (b+1)*Mr) of A (whose rows are lda reals long; for a
// packed A, lda = total real columns of the PK buffer), the packed B operand at Bp + b*bp_batch_stride and writes the
// [Mr x N] block b of C (dense, ldc = N when batched) or of every split's partials [split][batch][Mr][N].
static int launch_gemm_h_tn_packed_b(const float* A, const __half* Bp, const float* amax_b, int64_t bp_batch_stride, float* C,
                                     int64_t Mr, int N, int64_t Kv, int64_t lda, int64_t ldc, int batch, int split,
                                     int64_t k_per_split, float* parts, int n_main, const float* amax_a, int a_packed,
                                     cudaStream_t st);

size_t gemm_h_tn_xhat_ws_bytes(int Ci, int64_t Kv, int M) {
    const int npad = (2 * Ci + 15) / 16 * 16;
    const int atoms = (npad + 63) / 64;
    const int64_t nchunks = (Kv + th::KV - 1) / th::KV;
    return 256 + align_up((size_t)M * nchunks * 2 * th::KV * atoms * 128, 256) + 256;
}

// P[m][2RCo][2Ci] (or the split partials) = G_m^T Xh_m for all m in ONE launch; G = [Kv x M*Mr] (fp32, or PK when a_packed)
int launch_gemm_h_tn_xhat(const float* G, const float* x, float* P, int64_t Mr, int Ci, int band_limit, int64_t Kv, int split,
                          int64_t k_per_split, float* parts, int n_main, const float* amax_g, void* bp_ws, size_t bp_bytes,
                          int a_packed, cudaStream_t st, const float* x_bound) {
    const int M = 2 * band_limit + 1;
    const int N = 2 * Ci;
    FCB_REQUIRE(G && x && P && bp_ws && amax_g, FCB_E_ARG, "gemm_h_tn_xhat: null pointer");
    FCB_REQUIRE(bp_bytes >= gemm_h_tn_xhat_ws_bytes(Ci, Kv, M), FCB_E_WORKSPACE, "gemm_h_tn_xhat: packed-operand workspace too small");
    FCB_REQUIRE(band_limit >= 0 && band_limit <= FCB_MAX_BAND_LIMIT, FCB_E_UNSUPPORTED, "gemm_h_tn_xhat: band_limit");
    if (Mr == 0 || Kv == 0) return FCB_OK;
    const int npad = (N + 15) / 16 * 16;
    const int atoms = (npad + 63) / 64;
    const int64_t nchunks = (Kv + th::KV - 1) / th::KV;
    const float* amax_b = x_bound;                  // max_i |x_i|: supplied, else one pass over x
    __half* Bp = reinterpret_cast<__half*>(static_cast<char*>(bp_ws) + 256);
    const int64_t bp_stride = nchunks * 2 * th::KV * atoms * 64;            // fp16 elements per batch entry
    if (!amax_b) {
        int rc = launch_bound_modulus(x, Kv * Ci, static_cast<float*>(bp_ws), st);
        if (rc) return rc;
        amax_b = static_cast<const float*>(bp_ws);
    }
    {
        const int64_t items = nchunks * th::KV * atoms * 8;
        const unsigned pb = (unsigned)((items + 255) / 256);
        const float2* x2 = reinterpret_cast<const float2*>(x);
        prof_begin("pack_xhat_tn", st);
        switch (band_limit) {
            case 0: th::k_pack_xhat_tn<0><<<pb, 256, 0, st>>>(x2, Bp, Kv, Ci, atoms, nchunks, bp_stride, amax_b); break;
            case 1: th::k_pack_xhat_tn<1><<<pb, 256, 0, st>>>(x2, Bp, Kv, Ci, atoms, nchunks, bp_stride, amax_b); break;
            case 2: th::k_pack_xhat_tn<2><<<pb, 256, 0, st>>>(x2, Bp, Kv, Ci, atoms, nchunks, bp_stride, amax_b); break;
            case 3: th::k_pack_xhat_tn<3><<<pb, 256, 0, st>>>(x2, Bp, Kv, Ci, atoms, nchunks, bp_stride, amax_b); break;
            default: th::k_pack_xhat_tn<4><<<pb, 256, 0, st>>>(x2, Bp, Kv, Ci, atoms, nchunks, bp_stride, amax_b); break;
        }
        prof_end(st);
        FCB_CUDA_LAUNCH_CHECK("pack_xhat_tn");
    }
    return launch_gemm_h_tn_packed_b(G, Bp, amax_b, bp_stride, P, Mr, N, Kv, (int64_t)M * Mr, N, M, split, k_per_split, parts, n_main,
                                     amax_g, a_packed, st);
}

int launch_gemm_h_tn(const float* A, const float* B, float* C, int64_t Mr, int N, int64_t Kv, int64_t lda, int64_t ldb,
                     int64_t ldc, int split, int64_t k_per_split, float* parts, int n_main, const float* amax_a, void* bp_ws,
                     size_t bp_bytes, int a_packed, cudaStream_t st, const float* b_bound) {
    FCB_REQUIRE(A && B && C && bp_ws && amax_a, FCB_E_ARG, "gemm_h_tn: null pointer");
    FCB_REQUIRE(N > 0 && N <= 256 && split >= 1 && split <= 65535, FCB_E_UNSUPPORTED, "gemm_h_tn: unsupported shape");
    if (a_packed) {
        FCB_REQUIRE((Mr % PK_COLS) == 0 && (reinterpret_cast<uintptr_t>(A) & 127u) == 0, FCB_E_ARG,
                    "gemm_h_tn: a packed A operand needs Mr %% 64 == 0 and a 128-byte aligned buffer");
        lda = 4;
    }
    FCB_REQUIRE((lda % 4) == 0 && aligned16(A) && aligned16(bp_ws), FCB_E_ALIGN, "gemm_h_tn: alignment");
    FCB_REQUIRE(k_per_split % th::KV == 0 || split == 1, FCB_E_ARG, "gemm_h_tn: vertex ranges must be multiples of 64");
    FCB_REQUIRE(bp_bytes >= gemm_h_tn_ws_bytes(N, Kv), FCB_E_WORKSPACE, "gemm_h_tn: packed-operand workspace too small");
    if (Mr == 0) return FCB_OK;
    const int npad = (N + 15) / 16 * 16;
    const int nb_atoms = (npad + 63) / 64;
    const int64_t nchunks = (Kv + th::KV - 1) / th::KV;
    const float* amax_b = b_bound;                  // bound supplied by the producer of B, else one pass over it
    __half* Bp = reinterpret_cast<__half*>(static_cast<char*>(bp_ws) + 256);
    {
        if (!amax_b) {
            int rc = th::launch_absmax(B, Kv, N, ldb, 1, 0, static_cast<float*>(bp_ws), st);
            if (rc) return rc;
            amax_b = static_cast<const float*>(bp_ws);
        }
        const int64_t items = nchunks * th::KV * nb_atoms * 8;
        if (items > 0)
            FCB_LAUNCH("pack_b_h_tn", st, th::k_pack_b_h_tn<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(B, Bp, Kv, N, nb_atoms, ldb, nchunks, amax_b));
    }
    return launch_gemm_h_tn_packed_b(A, Bp, amax_b, 0, C, Mr, N, Kv, a_packed ? Mr : lda, ldc, 1, split, k_per_split, parts, n_main, amax_a,
                                     a_packed, st);
}

static int launch_gemm_h_tn_packed_b(const float* A, const __half* Bp, const float* amax_b, int64_t bp_batch_stride, float* C,
                                     int64_t Mr, int N, int64_t Kv, int64_t lda, int64_t ldc, int batch, int split,
                                     int64_t k_per_split, float* parts, int n_main, const float* amax_a, int a_packed,
                                     cudaStream_t st) {
    FCB_REQUIRE(batch >= 1 && batch <= 65535, FCB_E_ARG, "gemm_h_tn: bad batch");
    FCB_REQUIRE(batch == 1 || ldc == N, FCB_E_ARG, "gemm_h_tn: a batched product writes dense [Mr x N] blocks");
    if (a_packed) FCB_REQUIRE((lda % PK_COLS) == 0 && lda >= (int64_t)batch * Mr, FCB_E_ARG, "gemm_h_tn: packed A needs lda = total PK columns");
    const int npad = (N + 15) / 16 * 16;
    const int nb_atoms = (npad + 63) / 64;
    th::ParamsTN p;
    p.A = A; p.Bp = Bp;
    p.C = split > 1 ? parts : C;
    p.amax_a = amax_a; p.amax_b = amax_b;
    p.Mr = Mr; p.Kv = Kv; p.lda = a_packed ? 4 : lda;
    p.ldc = split > 1 ? N : ldc;
    p.k_per_split = k_per_split;
    p.part_stride = split > 1 ? (int64_t)batch * Mr * (int64_t)N : 0;
    p.c_batch_stride = Mr * (split > 1 ? (int64_t)N : ldc);
    p.a_batch_off = Mr;
    p.bp_batch_stride = bp_batch_stride;
    p.N = N; p.Npad = npad; p.nb_atoms = nb_atoms;
    p.wide = (!a_packed && (lda % 8) == 0 && (Mr % 8) == 0 && (reinterpret_cast<uintptr_t>(A) & 31u) == 0) ? 1 : 0;
    p.a_chunks = (int)(Mr / PK_COLS);
    p.a_tile_stride = (a_packed ? lda / PK_COLS : (int64_t)p.a_chunks) * PK_BLOCK_BYTES;
    FCB_REQUIRE(n_main >= 1 && npad * (n_main + 1) <= 512, FCB_E_ARG, "gemm_h_tn: accumulators do not fit TMEM");
    p.n_main = n_main;
    uint32_t cols = 32;
    while ((int)cols < npad * (n_main + 1)) cols <<= 1;
    p.tmem_cols = cols;
    const size_t stage_bytes = 2 * th::A_PLANE + 2 * (size_t)th::KV * nb_atoms * 128;
    int stages = (int)((220 * 1024 - 2048) / stage_bytes);
    if (stages > 6) stages = 6;
    FCB_REQUIRE(stages >= 1, FCB_E_UNSUPPORTED, "gemm_h_tn: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 1024 + 8 * (3 * stages + 2) + 64;
    dim3 grid((unsigned)((Mr + th::BM - 1) / th::BM), (unsigned)batch, (unsigned)split);
    // the opt-in shared-memory size is a per-device (per-context) function attribute: remember it per device
    static std::atomic<bool> attr_set_dev[64];
    std::atomic<bool> attr_unknown_dev{false};
    int attr_dev = 0;
    if (cudaGetDevice(&attr_dev) != cudaSuccess) attr_dev = -1;
    std::atomic<bool>& attr_set = (attr_dev >= 0 && attr_dev < 64) ? attr_set_dev[attr_dev] : attr_unknown_dev;
    if (!attr_set.load(std::memory_order_acquire)) {       // idempotent: racing threads at worst set the attribute twice
        cudaError_t e = cudaFuncSetAttribute(th::k_gemm_h_tn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(th::k_gemm_h_tn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("gemm_h_tn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return FCB_E_CUDA;
        }
        attr_set.store(true, std::memory_order_release);
    }
    if (a_packed) FCB_LAUNCH("gemm_p_tn", st, th::k_gemm_h_tn<true><<<grid, th::THREADS, smem, st>>>(p));
    else FCB_LAUNCH("gemm_h_tn", st, th::k_gemm_h_tn<false><<<grid, th::THREADS, smem, st>>>(p));
    return FCB_OK;
}

}  // namespace fcb
