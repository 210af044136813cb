// K2 / K5b on the 5th-generation tensor cores: C[M x N] = A[M x K] * B[K x N], fp32 in / fp32 out, computed
// with tcgen05.mma kind::tf32 and fp32 accumulators in TMEM.
//
// Accuracy modes
//   3xTF32 (default tensor-core mode): every operand is split into hi = tf32(x) and lo = x - hi and the
//   product is accumulated as  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (the dropped lo*lo term is ~2^-22 relative),
//   which removes the operand-rounding error.  What remains is the tensor core's accumulator behaviour: the
//   fp32 accumulate in TMEM truncates (measured: a systematic toward-zero drift of ~6e-8 per accumulating MMA,
//   2e-5 after the 1080 MMAs of a K=2880 contraction).  So the small cross terms go to their OWN accumulator
//   (their truncation error is 2^-11 smaller) and the hi*hi products are dealt round-robin over up to three
//   more accumulators (TMEM has 512 columns); the epilogue adds the accumulators with ordinary fp32 adds.
//   TF32: hi planes only (~1e-3 relative; looser tolerance, stated separately in the tests).
//
// Structure of one CTA (one 128-row tile of A, all N <= 256 columns), 18 warps:
//   warps 0-15 producers: 128-bit coalesced loads of the fp32 A tile (32 reals = 128 B per row and stage),
//              hi/lo split in registers, stores into the canonical 128B-swizzled K-major UMMA layout,
//              fence.proxy.async, mbarrier arrive.  After the K loop the same warps run the epilogue
//              (tcgen05.ld 32x32b -> registers -> 128-bit global stores).
//   warp 16    single-thread MMA issue (tcgen05.mma, tcgen05.commit), TMEM alloc / dealloc.
//   warp 17    B operand: the weights are pre-packed (k_pack_b_tc) into the exact shared-memory image of every
//              stage, so one cp.async.bulk (TMA engine, mbarrier complete_tx) per stage brings hi and lo planes in.
// Shared-memory ring of S stages: {A_hi 16 KB, A_lo 16 KB, B_hi, B_lo (Npad*128 B each)}.
#include <atomic>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fcb {
namespace tc {

constexpr int KC = 32;         // reals per stage == one 128-byte swizzle row
constexpr int N_PROD_WARPS = 16;
constexpr int N_PROD = N_PROD_WARPS * 32;        // producer threads
constexpr int A_F4 = BM * KC / 4 / N_PROD;        // float4 of the A tile per producer thread and stage (2)
constexpr int THREADS = (N_PROD_WARPS + 2) * 32;
constexpr uint32_t A_PLANE = BM * KC * 4;   // 16 KB
constexpr int PF = 6;          // producer prefetch depth: chunks of A in flight in registers (memory-level parallelism)

// debug timeline (tools/trace_gemm.py): when non-null, CTA 0 records clock64() at pipeline events
__device__ long long* g_trace = nullptr;
#define FCB_TRACE(slot, kc, cond)                                                          \
    do {                                                                                   \
        if (trace && (cond) && (kc) < 64) trace[(slot) * 64 + (kc)] = clock64();           \
    } while (0)

struct Params {
    const float* A;
    const float* Bp;   // packed B: [batch][chunk][plane(hi,lo)][Npad][32] swizzled
    float* C;
    int64_t M, K, lda, ldc, sa, sc;
    int64_t bp_batch_stride;   // floats
    int N, Npad, nchunks, stages, mode;
    int n_main;                // accumulators for the hi*hi products (3xTF32: + 1 for the cross terms)
    int kgroups, cpg;          // grouped-K mode (kgroups > 1): K = kgroups * cpg chunks; group g accumulates (all three
                               // product types) into its own TMEM accumulator and lands in C columns [g*N, (g+1)*N)
    uint32_t tmem_cols;
};

__global__ void __launch_bounds__(THREADS, 1) k_gemm_tc_nn(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int S = p.stages;
    const uint32_t b_plane = (uint32_t)p.Npad * 128u;
    // layout: A_hi[S] | A_lo[S] | B[S] (hi,lo) | barriers
    const uint32_t a_hi0 = base, a_lo0 = base + S * A_PLANE, b0 = base + 2 * S * A_PLANE;
    const uint32_t bars = b0 + S * 2 * b_plane;
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto full_b = [&](int s) { return bars + 8u * (S + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * S + s); };
    const uint32_t tmem_full = bars + 8u * (3 * S);
    const uint32_t tmem_slot = bars + 8u * (3 * S + 1);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(sm + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* const trace = (blockIdx.x == 0 && blockIdx.y == 0) ? g_trace : nullptr;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int batch = blockIdx.y;
    const float* A = p.A + batch * p.sa;
    const float* Bp = p.Bp + batch * p.bp_batch_stride;
    float* C = p.C + batch * p.sc;

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
        // ------------------------------------------------------------------ producers
        const int t = threadIdx.x;   // 0..N_PROD-1
        // Software pipeline: the global loads of chunks kc+1 .. kc+PF-1 are in flight (registers) while chunk kc is
        // split and stored: with ~2 us of loaded HBM latency the bytes in flight per SM set the streaming rate.  All addressing is hoisted: per thread A_F4 source pointers and shared-memory offsets.
        const float* src[A_F4];
        uint32_t off[A_F4];
        int kcol[A_F4];
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int idx = t + N_PROD * i;
            const int row = idx >> 3, j = idx & 7;
            const int64_t m = m0 + row;
            src[i] = (m < p.M) ? (A + m * p.lda + 4 * j) : nullptr;
            kcol[i] = 4 * j;
            off[i] = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
        }
        uint8_t* const hi_base = sm + (a_hi0 - base);
        uint8_t* const lo_base = sm + (a_lo0 - base);
        const bool split3 = (p.mode == FCB_GEMM_TC_3XTF32);
        float4 v[PF][A_F4];
        auto issue = [&](int kc, float4(&dst)[A_F4]) {
            const int64_t k0 = (int64_t)kc * KC;
            if (k0 + KC <= p.K) {                        // full chunk (warp-uniform)
#pragma unroll
                for (int i = 0; i < A_F4; ++i)
                    dst[i] = src[i] ? __ldg(reinterpret_cast<const float4*>(src[i] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
                for (int i = 0; i < A_F4; ++i) {
                    dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int64_t k = k0 + kcol[i];
                    if (src[i]) {
                        if (k < p.K) dst[i].x = src[i][k0];
                        if (k + 1 < p.K) dst[i].y = src[i][k0 + 1];
                        if (k + 2 < p.K) dst[i].z = src[i][k0 + 2];
                        if (k + 3 < p.K) dst[i].w = src[i][k0 + 3];
                    }
                }
            }
        };
        uint32_t ps = 0, pph = 1;     // producer stage / parity of the `empty` barrier it waits for
        auto commit = [&](int kc, const float4(&sv)[A_F4]) {
            const uint32_t s = ps;
            FCB_TRACE(0, kc, t == 0);
            mbar_wait(empty(s), pph);
            FCB_TRACE(1, kc, t == 0);
#pragma unroll
            for (int i = 0; i < A_F4; ++i) {
                const float4 hi = make_float4(tf32_hi(sv[i].x), tf32_hi(sv[i].y), tf32_hi(sv[i].z), tf32_hi(sv[i].w));
                *reinterpret_cast<float4*>(hi_base + s * A_PLANE + off[i]) = hi;
                if (split3)
                    *reinterpret_cast<float4*>(lo_base + s * A_PLANE + off[i]) =
                        make_float4(sv[i].x - hi.x, sv[i].y - hi.y, sv[i].z - hi.z, sv[i].w - hi.w);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(s));
            FCB_TRACE(2, kc, t == 0);
            if (++ps == (uint32_t)S) { ps = 0; pph ^= 1u; }
        };
#pragma unroll
        for (int u = 0; u < PF - 1; ++u)
            if (u < p.nchunks) issue(u, v[u]);
        for (int kc = 0; kc < p.nchunks; kc += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int k = kc + u;
                if (k < p.nchunks) {
                    if (k + PF - 1 < p.nchunks) issue(k + PF - 1, v[(u + PF - 1) % PF]);
                    commit(k, v[u]);
                }
            }
        }
        // ------------------------------------------------------------------ epilogue
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3, part = warp >> 2;
        const int64_t m = m0 + 32 * q + lane;
        const int groups = p.Npad / 16;
        const bool grouped = p.kgroups > 1;
        const int n_acc = grouped ? 1 : p.n_main + (p.mode == FCB_GEMM_TC_3XTF32 ? 1 : 0);
        const int items = groups * (grouped ? p.kgroups : 1);       // (k-group, 16-column group) pairs
        for (int it = part; it < items; it += N_PROD_WARPS / 4) {
            const int kg = it / groups, g = it - kg * groups;
            uint32_t r[16];
            float acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0.f;
            for (int a = n_acc - 1; a >= 0; --a) {      // cross-term accumulator (last) first: small + large
                tc_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)((a + kg) * p.Npad + 16 * g), r);
                tc_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(r[e]);
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(acc[e]);
            if (m < p.M) {
                float* dst = C + m * p.ldc + (int64_t)kg * p.N + 16 * g;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int n = 16 * g + 4 * c4;
                    if (n + 3 < p.N) {
                        *reinterpret_cast<float4*>(dst + 4 * c4) =
                            make_float4(__uint_as_float(r[4 * c4]), __uint_as_float(r[4 * c4 + 1]),
                                        __uint_as_float(r[4 * c4 + 2]), __uint_as_float(r[4 * c4 + 3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + e < p.N) dst[4 * c4 + e] = __uint_as_float(r[4 * c4 + e]);
                    }
                }
            }
        }
    } else if (warp == N_PROD_WARPS) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        // Everything this thread needs per MMA is kept in counters / pre-built descriptors: a single thread
        // issues dependent scalar instructions ~5 cycles apart, so divisions or descriptor rebuilds in this loop
        // would cost more than the MMAs themselves (measured: ~270 cycles per k-step before, 48-cycle MMAs).
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(p.Npad);
            const bool x3 = (p.mode == FCB_GEMM_TC_3XTF32);
            const uint64_t a_hi_d = make_desc_k_sw128(a_hi0), a_lo_d = make_desc_k_sw128(a_lo0);
            const uint64_t b_hi_d = make_desc_k_sw128(b0), b_lo_d = make_desc_k_sw128(b0 + b_plane);
            const uint64_t a_step = (uint64_t)(A_PLANE >> 4), b_step = (uint64_t)((2 * b_plane) >> 4);
            const uint32_t d_x = tmem_d + (uint32_t)(p.n_main * p.Npad);
            const uint32_t n_main = (uint32_t)p.n_main, npad = (uint32_t)p.Npad;
            uint32_t s = 0, ph = 0;          // stage, phase
            uint32_t acc = 0, d_main = tmem_d, first = n_main;   // round-robin hi*hi accumulator; `first` MMAs overwrite
            uint32_t x_acc = 0;              // 0 only for the very first cross-term MMA
            const bool grouped = p.kgroups > 1;
            uint32_t g_left = (uint32_t)p.cpg, d_grp = tmem_d;
            for (int kc = 0; kc < p.nchunks; ++kc) {
                FCB_TRACE(3, kc, true);
                mbar_wait(full_a(s), ph);
                FCB_TRACE(4, kc, true);
                mbar_wait(full_b(s), ph);
                FCB_TRACE(5, kc, true);
                tc_fence_after();
                const uint64_t a_hi = a_hi_d + s * a_step, a_lo = a_lo_d + s * a_step;
                const uint64_t b_hi = b_hi_d + s * b_step, b_lo = b_lo_d + s * b_step;
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);   // 8 tf32 = 32 B = 2 x 16 B along K inside the swizzle row
                    if (grouped) {
                        tc_mma_tf32(d_grp, a_hi + adv, b_hi + adv, idesc, (ks == 0 && g_left == (uint32_t)p.cpg) ? 0u : 1u);
                        if (x3) {
                            tc_mma_tf32(d_grp, a_lo + adv, b_hi + adv, idesc, 1u);
                            tc_mma_tf32(d_grp, a_hi + adv, b_lo + adv, idesc, 1u);
                        }
                        continue;
                    }
                    tc_mma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, first ? 0u : 1u);
                    if (first) --first;
                    if (++acc == n_main) { acc = 0; d_main = tmem_d; } else d_main += npad;
                    if (x3) {
                        tc_mma_tf32(d_x, a_lo + adv, b_hi + adv, idesc, x_acc);
                        tc_mma_tf32(d_x, a_hi + adv, b_lo + adv, idesc, 1u);
                        x_acc = 1u;
                    }
                }
                if (grouped && --g_left == 0) { g_left = (uint32_t)p.cpg; d_grp += npad; }
                tc_commit(empty(s));      // frees the stage once these MMAs have read it
                FCB_TRACE(6, kc, true);
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
            tc_commit(tmem_full);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ B loader (one thread, TMA bulk copies)
        if (lane == 0) {
            const uint32_t bytes = (p.mode == FCB_GEMM_TC_3XTF32 ? 2u : 1u) * b_plane;
            const float* src = Bp;
            const int64_t src_step = (int64_t)2 * p.Npad * KC;
            uint32_t s = 0, ph = 1;
            for (int kc = 0; kc < p.nchunks; ++kc) {
                mbar_wait(empty(s), ph);
                FCB_TRACE(7, kc, true);
                mbar_expect_tx(full_b(s), bytes);
                bulk_copy_g2s(b0 + s * 2 * b_plane, src, bytes, full_b(s));
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

// ------------------------------------------------------------------------------------------------------------
// TN variant (K4, weight gradient): P[Mr x N] = A^T B with A = [Kv x Mr] and B = [Kv x N] row-major, i.e. both
// operands are "MN-major" (the reduction index, vertices, is the slow one in memory).  The hardware transposes:
// UMMA descriptors with a_major = b_major = MN.  For 32-bit (tf32) MN-major operands the only legal shared
// memory layout is SWIZZLE_128B_BASE32B: atoms of 4 K-rows x 128 bytes, a row = 32 consecutive features of one
// vertex (exactly a coalesced 128-byte piece of the global row), 32-byte chunks XOR-permuted by the row index.  Both operands are activations,
// so the producer warps split both into hi/lo planes.  grid.z = split over vertex ranges; every split writes its
// own partial tile, summed later in split order (k_reduce_splits) so the result is deterministic.
struct ParamsTN {
    const float* A;    // [Kv x Mr], lda
    const float* Bp;   // packed B: [chunk of 32 vertices][plane(hi,lo)] shared-memory images (k_pack_b_tn)
    float* C;          // partials [split][Mr][N] or C itself when split == 1 (ldc)
    int64_t Mr, Kv, lda, ldc, k_per_split, part_stride;
    int N, Npad, nb_atoms, stages, mode, n_main;
    uint32_t tmem_cols;
};

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // between 32-element MN atoms
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // between 4-row K groups
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                             // SWIZZLE_128B_BASE32B
    return d;
}

constexpr int KV = 32;   // vertices per stage (4 MMA K-steps of 8)
constexpr int PFT = 6;   // producer prefetch depth (chunks of A in flight in registers)

// byte offset of the 16-byte piece holding elements (vertex v of the chunk, features 4*f4 .. 4*f4+3) inside an
// MN-major SWIZZLE_128B_BASE32B operand tile with `atoms` 32-feature atoms per vertex row
__host__ __device__ __forceinline__ uint32_t tn_tile_off(int v, int f4, int atoms) {
    const int kq = v >> 2, row = v & 3, at = f4 >> 3, unit = f4 & 7;
    return (uint32_t)((kq * atoms + at) * 512 + row * 128 + (((unit >> 1) ^ row) << 5) + ((unit & 1) << 4));
}

// The weight gradient's B operand (gy) is the same for every 128-column tile of contrib, so it is split into hi/lo
// and laid out as the exact shared-memory image of every 32-vertex chunk ONCE (k_pack_b_tn); the GEMM CTAs fetch it
// with one cp.async.bulk per stage.  Only the A operand (contrib, read exactly once overall) goes through the
// producer warps' registers.
__global__ void __launch_bounds__(THREADS, 1) k_gemm_tc_tn(const ParamsTN p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int S = p.stages;
    const uint32_t b_plane = (uint32_t)KV * (uint32_t)p.nb_atoms * 128u;   // 32 vertices x nb_atoms x 128 B
    // layout: A_hi[S] | A_lo[S] | B[S] (hi,lo) | barriers
    const uint32_t a_hi0 = base, a_lo0 = base + S * A_PLANE, b0 = base + 2 * S * A_PLANE;
    const uint32_t bars = b0 + S * 2 * b_plane;
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto full_b = [&](int s) { return bars + 8u * (S + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * S + s); };
    const uint32_t tmem_full = bars + 8u * (3 * S);
    const uint32_t tmem_slot = bars + 8u * (3 * S + 1);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(sm + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* const trace = (blockIdx.x == 0 && blockIdx.z == 0) ? g_trace : nullptr;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int split = blockIdx.z;
    const int64_t kb = (int64_t)split * p.k_per_split;       // multiple of KV
    const int64_t ke = min(p.Kv, kb + p.k_per_split);
    const int nchunks = kb < ke ? (int)((ke - kb + KV - 1) / KV) : 0;
    float* C = p.C + (int64_t)split * p.part_stride;

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
    const bool split3 = (p.mode == FCB_GEMM_TC_3XTF32);

    if (warp < N_PROD_WARPS) {
        const int t = threadIdx.x;
        // hoisted addressing: element pointers at the split's first vertex, advanced by KV rows per stage
        const float* a_src[A_F4];
        uint32_t a_off[A_F4];
        int a_v[A_F4], a_cnt[A_F4];                       // vertex within the stage, valid floats (0..4)
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int idx = t + N_PROD * i;
            const int v = idx >> 5, f4 = idx & 31;
            const int64_t m = m0 + 4 * f4;
            a_v[i] = v;
            a_cnt[i] = (int)max((int64_t)0, min((int64_t)4, p.Mr - m));
            a_src[i] = p.A + (kb + v) * p.lda + m;
            a_off[i] = tn_tile_off(v, f4, 4);
        }
        const bool cols_full = (m0 + BM <= p.Mr);        // CTA-uniform: every column piece of this tile is complete
        const int64_t a_step = (int64_t)KV * p.lda;
        float4 va[PFT][A_F4];
        auto issue = [&](int kc, float4(&da)[A_F4]) {
            const int64_t v0 = kb + (int64_t)kc * KV;
            if (cols_full && v0 + KV <= ke) {            // fast path: no guards
#pragma unroll
                for (int i = 0; i < A_F4; ++i) {
                    da[i] = __ldg(reinterpret_cast<const float4*>(a_src[i]));
                    a_src[i] += a_step;
                }
            } else {
#pragma unroll
                for (int i = 0; i < A_F4; ++i) {
                    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (v0 + a_v[i] < ke) {
                        const float* q = a_src[i];
                        if (a_cnt[i] == 4) r = __ldg(reinterpret_cast<const float4*>(q));
                        else {
                            if (a_cnt[i] > 0) r.x = q[0];
                            if (a_cnt[i] > 1) r.y = q[1];
                            if (a_cnt[i] > 2) r.z = q[2];
                        }
                    }
                    da[i] = r;
                    a_src[i] += a_step;
                }
            }
        };
        uint8_t* const ahi = sm + (a_hi0 - base);
        uint8_t* const alo = sm + (a_lo0 - base);
        uint32_t ps = 0, pph = 1;
        auto commit = [&](int kc, const float4(&sa)[A_F4]) {
            const uint32_t s = ps;
            FCB_TRACE(0, kc, t == 0);
            mbar_wait(empty(s), pph);
            FCB_TRACE(1, kc, t == 0);
#pragma unroll
            for (int i = 0; i < A_F4; ++i) {
                const float4 hi = make_float4(tf32_hi(sa[i].x), tf32_hi(sa[i].y), tf32_hi(sa[i].z), tf32_hi(sa[i].w));
                *reinterpret_cast<float4*>(ahi + s * A_PLANE + a_off[i]) = hi;
                if (split3)
                    *reinterpret_cast<float4*>(alo + s * A_PLANE + a_off[i]) =
                        make_float4(sa[i].x - hi.x, sa[i].y - hi.y, sa[i].z - hi.z, sa[i].w - hi.w);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(s));
            FCB_TRACE(2, kc, t == 0);
            if (++ps == (uint32_t)S) { ps = 0; pph ^= 1u; }
        };
#pragma unroll
        for (int u = 0; u < PFT - 1; ++u)
            if (u < nchunks) issue(u, va[u]);
        for (int kc = 0; kc < nchunks; kc += PFT) {
#pragma unroll
            for (int u = 0; u < PFT; ++u) {
                const int k = kc + u;
                if (k < nchunks) {
                    if (k + PFT - 1 < nchunks) issue(k + PFT - 1, va[(u + PFT - 1) % PFT]);
                    commit(k, va[u]);
                }
            }
        }
        // epilogue
        const int q = warp & 3, part = warp >> 2;
        const int64_t m = m0 + 32 * q + lane;
        const int groups = p.Npad / 16;
        const int n_acc = p.n_main + (split3 ? 1 : 0);
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
                for (int a = n_acc - 1; a >= 0; --a) {
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
                    if (16 * g + e < p.N) dst[e] = acc[e];
            }
        }
    } else if (warp == N_PROD_WARPS) {
        if (lane == 0 && nchunks > 0) {
            // D=f32, A=B=tf32, both MN-major (bits 15, 16), M=128, N=Npad.  Counters and pre-built descriptors
            // only: see the NN issuer.
            const uint32_t idesc = make_idesc_tf32(p.Npad) | (1u << 15) | (1u << 16);
            const uint32_t sbo_a = 4 * 512, sbo_b = (uint32_t)p.nb_atoms * 512;   // between 4-vertex K groups
            const uint64_t a_hi_d = make_desc_mn_sw128(a_hi0, 512, sbo_a), a_lo_d = make_desc_mn_sw128(a_lo0, 512, sbo_a);
            const uint64_t b_hi_d = make_desc_mn_sw128(b0, 512, sbo_b), b_lo_d = make_desc_mn_sw128(b0 + b_plane, 512, sbo_b);
            const uint64_t a_stage = (uint64_t)(A_PLANE >> 4), b_stage = (uint64_t)((2 * b_plane) >> 4);
            const uint64_t a_kg = (uint64_t)((2 * sbo_a) >> 4), b_kg = (uint64_t)((2 * sbo_b) >> 4);   // 8 vertices
            const uint32_t d_x = tmem_d + (uint32_t)(p.n_main * p.Npad);
            const uint32_t n_main = (uint32_t)p.n_main, npad = (uint32_t)p.Npad;
            uint32_t s = 0, ph = 0, acc = 0, d_main = tmem_d, first = n_main, x_acc = 0;
            for (int kc = 0; kc < nchunks; ++kc) {
                FCB_TRACE(3, kc, true);
                mbar_wait(full_a(s), ph);
                FCB_TRACE(4, kc, true);
                mbar_wait(full_b(s), ph);
                FCB_TRACE(5, kc, true);
                tc_fence_after();
                uint64_t a_hi = a_hi_d + s * a_stage, a_lo = a_lo_d + s * a_stage;
                uint64_t b_hi = b_hi_d + s * b_stage, b_lo = b_lo_d + s * b_stage;
#pragma unroll
                for (int kg = 0; kg < KV / 8; ++kg) {
                    tc_mma_tf32(d_main, a_hi, b_hi, idesc, first ? 0u : 1u);
                    if (first) --first;
                    if (++acc == n_main) { acc = 0; d_main = tmem_d; } else d_main += npad;
                    if (split3) {
                        tc_mma_tf32(d_x, a_lo, b_hi, idesc, x_acc);
                        tc_mma_tf32(d_x, a_hi, b_lo, idesc, 1u);
                        x_acc = 1u;
                    }
                    a_hi += a_kg; a_lo += a_kg; b_hi += b_kg; b_lo += b_kg;
                }
                tc_commit(empty(s));
                FCB_TRACE(6, kc, true);
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
            tc_commit(tmem_full);
        }
        __syncwarp();
    } else {
        // B loader (one thread, TMA bulk copies of the pre-packed chunk images)
        if (lane == 0) {
            const uint32_t bytes = (split3 ? 2u : 1u) * b_plane;
            const int64_t img = (int64_t)2 * (b_plane / 4);                 // floats per chunk image (hi + lo)
            const float* src = p.Bp + (kb / KV) * img;
            uint32_t s = 0, ph = 1;
            for (int kc = 0; kc < nchunks; ++kc) {
                mbar_wait(empty(s), ph);
                FCB_TRACE(7, kc, true);
                mbar_expect_tx(full_b(s), bytes);
                bulk_copy_g2s(b0 + s * 2 * b_plane, src, bytes, full_b(s));
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

// B[Kv x N] row-major (ldb) -> per 32-vertex chunk the MN-major swizzled shared-memory image, hi plane then lo plane
// (hi = tf32(b), lo = b - hi); vertices >= Kv and features >= N are zero.  One thread per 16-byte piece.
__global__ void k_pack_b_tn(const float* __restrict__ B, float* __restrict__ Bp, int64_t Kv, int N, int atoms, int64_t ldb,
                            int64_t nchunks) {
    const int f4_per_row = atoms * 8;
    const int64_t per_chunk = (int64_t)KV * f4_per_row;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks * per_chunk) return;
    const int64_t c = i / per_chunk;
    const int r = (int)(i - c * per_chunk);
    const int v = r / f4_per_row, f4 = r - v * f4_per_row;
    const int64_t vg = c * KV + v;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (vg < Kv) {
        const float* q = B + vg * ldb + 4 * f4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (4 * f4 + e < N) x[e] = q[e];
    }
    float4 hi = make_float4(tf32_hi(x[0]), tf32_hi(x[1]), tf32_hi(x[2]), tf32_hi(x[3]));
    float4 lo = make_float4(x[0] - hi.x, x[1] - hi.y, x[2] - hi.z, x[3] - hi.w);
    const int64_t plane = (int64_t)KV * atoms * 32;                       // floats per plane
    float* dst = Bp + c * 2 * plane + tn_tile_off(v, f4, atoms) / 4;
    *reinterpret_cast<float4*>(dst) = hi;
    *reinterpret_cast<float4*>(dst + plane) = lo;
}

// B[K x N] row-major (ldb) -> [chunk][plane][Npad][32] with the 128B swizzle applied (16-byte unit j of row n
// stored at unit j ^ (n & 7)), hi = tf32(b), lo = b - hi; rows n >= N and k >= K are zero.
__global__ void k_pack_b_tc(const float* __restrict__ B, float* __restrict__ Bp, int64_t K, int N, int Npad, int64_t ldb,
                            int nchunks, int64_t sb, int64_t bp_batch_stride) {
    const int64_t per = (int64_t)nchunks * Npad * KC;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const int batch = blockIdx.y;
    const int kk = (int)(i % KC);
    const int n = (int)((i / KC) % Npad);
    const int64_t c = i / ((int64_t)KC * Npad);
    const int64_t k = c * KC + kk;
    float v = 0.f;
    if (n < N && k < K) v = B[batch * sb + k * ldb + n];
    const float hi = tf32_hi(v);
    const float lo = v - hi;
    const int unit = (kk >> 2) ^ (n & 7);
    float* dst = Bp + batch * bp_batch_stride + (c * 2) * (int64_t)Npad * KC + (int64_t)n * KC + unit * 4 + (kk & 3);
    dst[0] = hi;
    dst[(int64_t)Npad * KC] = lo;
}

}  // namespace tc

}  // namespace fcb

// debug: device buffer of 8*64 int64 receiving the pipeline timeline of CTA 0 of k_gemm_tc_nn (NULL = off)
extern "C" int fcb_debug_trace(void* dev_buf) {
    long long* p = static_cast<long long*>(dev_buf);
    return cudaMemcpyToSymbol(fcb::tc::g_trace, &p, sizeof(p)) == cudaSuccess ? FCB_OK : FCB_E_CUDA;
}

namespace fcb {

// tcgen05 accumulation into TMEM truncates (measured drift ~6e-8 per accumulating MMA), so no accumulator may
// receive more than TC_MAX_ACC_MMAS of them inside the fp32 parity budget: the hi*hi products of a 3xTF32
// contraction are dealt round-robin over n_main accumulators and wide outputs are cut into column chunks narrow
// enough that n_main + 1 accumulators fit the 512 TMEM columns.  Returns the chunk width (0: not feasible).
constexpr int64_t TC_MAX_ACC_MMAS = 400;
int gemm_tc_plan(int N, int64_t ksteps, int mode, int* n_main_out) {
    if (N <= 0) return 0;
    if (mode == FCB_GEMM_TC_TF32) {
        *n_main_out = 1;
        return N < 256 ? N : 256;
    }
    const int64_t need = (ksteps + TC_MAX_ACC_MMAS - 1) / TC_MAX_ACC_MMAS;
    const int widths[3] = {128, 64, 32};
    for (int w : widths) {
        const int nc = N < w ? N : w;
        const int npad = (nc + 15) / 16 * 16;
        const int avail = 512 / npad - 1;
        if (avail >= 1 && need <= avail) {
            int n_main = avail < 3 ? avail : 3;
            if (n_main < need) n_main = (int)need;
            if (n_main > ksteps) n_main = (int)(ksteps < 1 ? 1 : ksteps);
            *n_main_out = n_main;
            return nc;
        }
    }
    return 0;
}

size_t gemm_tc_ws_bytes(int N, int64_t K, int batch) {
    const int npad = (N + 15) / 16 * 16;
    const int64_t nchunks = (K + tc::KC - 1) / tc::KC;
    return align_up((size_t)batch * nchunks * 2 * npad * tc::KC * 4, 256) + 256;
}

// kgroups == 1: `batch` independent GEMMs (grid.y).  kgroups > 1 (batch must be 1): ONE pass over A = [M x kgroups*K]
// whose k-group g (K columns) is contracted with B_g = B + g*sb and written to C columns [g*N, (g+1)*N) — the
// batched-over-m grad-x contraction as a single long-K pipeline instead of kgroups short ones.
int launch_gemm_tc_nn(const float* A, const float* B, float* C, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb,
                      int64_t ldc, int batch, int64_t sa, int64_t sb, int64_t sc, int mode, int n_main, int kgroups,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
    FCB_REQUIRE(A && B && C && ws, FCB_E_ARG, "gemm_tc: null pointer");
    FCB_REQUIRE(M >= 0 && N > 0 && K > 0 && batch >= 1 && batch <= 65535, FCB_E_ARG, "gemm_tc: bad sizes");
    FCB_REQUIRE(N <= 256, FCB_E_UNSUPPORTED, "gemm_tc: N=%d > 256 not supported by one accumulator tile", N);
    FCB_REQUIRE(mode == FCB_GEMM_TC_3XTF32 || mode == FCB_GEMM_TC_TF32, FCB_E_ARG, "gemm_tc: bad mode");
    FCB_REQUIRE((lda % 4) == 0 && (ldc % 4) == 0 && (sa % 4) == 0 && (sc % 4) == 0 && aligned16(A) && aligned16(C),
                FCB_E_ALIGN, "gemm_tc: A/C leading dimensions and strides must be multiples of 4 floats, 16-byte aligned");
    FCB_REQUIRE(kgroups >= 1 && (kgroups == 1 || (batch == 1 && K % tc::KC == 0)), FCB_E_ARG, "gemm_tc: bad k-group shape");
    FCB_REQUIRE(ws_bytes >= gemm_tc_ws_bytes(N, K, batch * kgroups), FCB_E_WORKSPACE, "gemm_tc: workspace too small");
    if (M == 0) return FCB_OK;
    const int npad = (N + 15) / 16 * 16;
    const int nchunks = (int)((K + tc::KC - 1) / tc::KC);
    float* Bp = static_cast<float*>(ws);
    const int64_t bp_stride = (int64_t)nchunks * 2 * npad * tc::KC;
    {
        const int64_t per = (int64_t)nchunks * npad * tc::KC;
        dim3 grid((unsigned)((per + 255) / 256), (unsigned)(batch * kgroups));
        FCB_LAUNCH("pack_b_tc", st, tc::k_pack_b_tc<<<grid, 256, 0, st>>>(B, Bp, K, N, npad, ldb, nchunks, sb, bp_stride));
    }
    tc::Params p;
    p.A = A; p.Bp = Bp; p.C = C;
    p.M = M; p.K = K; p.lda = lda; p.ldc = ldc; p.sa = sa; p.sc = sc;
    p.bp_batch_stride = bp_stride;
    p.N = N; p.Npad = npad; p.nchunks = nchunks * kgroups; p.mode = mode;
    p.kgroups = kgroups; p.cpg = nchunks;
    if (kgroups > 1) {
        p.K = K * kgroups;
        n_main = kgroups;      // TMEM columns: one accumulator per k-group, cross terms folded in
    }
    FCB_REQUIRE(n_main >= 1 && npad * (n_main + ((mode == FCB_GEMM_TC_3XTF32 && kgroups == 1) ? 1 : 0)) <= 512, FCB_E_ARG,
                "gemm_tc: accumulators do not fit TMEM");
    p.n_main = n_main;
    uint32_t cols = 32;
    while ((int)cols < npad * (n_main + ((mode == FCB_GEMM_TC_3XTF32 && kgroups == 1) ? 1 : 0))) cols <<= 1;
    p.tmem_cols = cols;
    const size_t stage_bytes = 2 * tc::A_PLANE + 2 * (size_t)npad * 128;
    int stages = (int)((220 * 1024 - 2048) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages > p.nchunks) stages = p.nchunks < 1 ? 1 : p.nchunks;
    FCB_REQUIRE(stages >= 1, FCB_E_UNSUPPORTED, "gemm_tc: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 1024 /*align slack*/ + 8 * (3 * stages + 2) + 64;
    // the opt-in shared-memory size is a per-device (per-context) function attribute: remember it per device
    static std::atomic<bool> attr_set_dev[64];
    std::atomic<bool> attr_unknown_dev{false};
    int attr_dev = 0;
    if (cudaGetDevice(&attr_dev) != cudaSuccess) attr_dev = -1;
    std::atomic<bool>& attr_set = (attr_dev >= 0 && attr_dev < 64) ? attr_set_dev[attr_dev] : attr_unknown_dev;
    if (!attr_set.load(std::memory_order_acquire)) {       // idempotent: racing threads at worst set the attribute twice
        cudaError_t e = cudaFuncSetAttribute(tc::k_gemm_tc_nn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return FCB_E_CUDA;
        }
        attr_set.store(true, std::memory_order_release);
    }
    dim3 grid((unsigned)((M + tc::BM - 1) / tc::BM), (unsigned)batch);
    FCB_LAUNCH("gemm_tc_nn", st, tc::k_gemm_tc_nn<<<grid, tc::THREADS, smem, st>>>(p));
    return FCB_OK;
}

// workspace of the packed B operand of the TN GEMM (hi + lo images of every 32-vertex chunk)
size_t gemm_tc_tn_ws_bytes(int N, int64_t Kv) {
    const int npad = (N + 15) / 16 * 16;
    const int atoms = (npad + 31) / 32;
    const int64_t nchunks = (Kv + tc::KV - 1) / tc::KV;
    return align_up((size_t)nchunks * 2 * tc::KV * atoms * 128, 256) + 256;
}

// P[Mr x N] = A^T B on the tensor cores; split >= 1 vertex ranges; `parts` (split*Mr*N floats) is used when split > 1;
// `bp_ws` (gemm_tc_tn_ws_bytes) receives the packed B operand.
int launch_gemm_tc_tn(const float* A, const float* B, float* C, int64_t Mr, int N, int64_t Kv, int64_t lda, int64_t ldb,
                      int64_t ldc, int split, int64_t k_per_split, float* parts, int mode, int n_main, void* bp_ws,
                      size_t bp_bytes, cudaStream_t st) {
    FCB_REQUIRE(A && B && C && bp_ws, FCB_E_ARG, "gemm_tc_tn: null pointer");
    FCB_REQUIRE(N > 0 && N <= 256 && split >= 1 && split <= 65535, FCB_E_UNSUPPORTED, "gemm_tc_tn: unsupported shape");
    FCB_REQUIRE((lda % 4) == 0 && aligned16(A) && aligned16(bp_ws), FCB_E_ALIGN, "gemm_tc_tn: alignment");
    FCB_REQUIRE(k_per_split % tc::KV == 0 || split == 1, FCB_E_ARG, "gemm_tc_tn: vertex ranges must be multiples of 32");
    FCB_REQUIRE(bp_bytes >= gemm_tc_tn_ws_bytes(N, Kv), FCB_E_WORKSPACE, "gemm_tc_tn: packed-operand workspace too small");
    if (Mr == 0) return FCB_OK;
    const int npad = (N + 15) / 16 * 16;
    const int nb_atoms = (npad + 31) / 32;
    const int64_t nchunks = (Kv + tc::KV - 1) / tc::KV;
    float* Bp = static_cast<float*>(bp_ws);
    {
        const int64_t items = nchunks * tc::KV * nb_atoms * 8;
        if (items > 0)
            FCB_LAUNCH("pack_b_tn", st, tc::k_pack_b_tn<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(B, Bp, Kv, N, nb_atoms, ldb, nchunks));
    }
    tc::ParamsTN p;
    p.A = A; p.Bp = Bp;
    p.C = split > 1 ? parts : C;
    p.Mr = Mr; p.Kv = Kv; p.lda = lda;
    p.ldc = split > 1 ? N : ldc;
    p.k_per_split = k_per_split;
    p.part_stride = split > 1 ? Mr * (int64_t)N : 0;
    p.N = N; p.Npad = npad; p.nb_atoms = nb_atoms; p.mode = mode;
    FCB_REQUIRE(n_main >= 1 && npad * (n_main + (mode == FCB_GEMM_TC_3XTF32 ? 1 : 0)) <= 512, FCB_E_ARG,
                "gemm_tc_tn: accumulators do not fit TMEM");
    p.n_main = n_main;
    uint32_t cols = 32;
    while ((int)cols < npad * (n_main + (mode == FCB_GEMM_TC_3XTF32 ? 1 : 0))) cols <<= 1;
    p.tmem_cols = cols;
    const size_t stage_bytes = 2 * tc::A_PLANE + 2 * (size_t)tc::KV * nb_atoms * 128;
    int stages = (int)((220 * 1024 - 2048) / stage_bytes);
    if (stages > 8) stages = 8;
    FCB_REQUIRE(stages >= 1, FCB_E_UNSUPPORTED, "gemm_tc_tn: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 1024 + 8 * (3 * stages + 2) + 64;
    dim3 grid((unsigned)((Mr + tc::BM - 1) / tc::BM), 1, (unsigned)split);
    // the opt-in shared-memory size is a per-device (per-context) function attribute: remember it per device
    static std::atomic<bool> attr_set_dev[64];
    std::atomic<bool> attr_unknown_dev{false};
    int attr_dev = 0;
    if (cudaGetDevice(&attr_dev) != cudaSuccess) attr_dev = -1;
    std::atomic<bool>& attr_set = (attr_dev >= 0 && attr_dev < 64) ? attr_set_dev[attr_dev] : attr_unknown_dev;
    if (!attr_set.load(std::memory_order_acquire)) {       // idempotent: racing threads at worst set the attribute twice
        cudaError_t e = cudaFuncSetAttribute(tc::k_gemm_tc_tn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("gemm_tc_tn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return FCB_E_CUDA;
        }
        attr_set.store(true, std::memory_order_release);
    }
    FCB_LAUNCH("gemm_tc_tn", st, tc::k_gemm_tc_tn<<<grid, tc::THREADS, smem, st>>>(p));
    return FCB_OK;
}

}  // namespace fcb
