// K2 / K4 / K5b (FP32 FMA path) — real fp32 GEMM used for the three dense contractions of
// the layer.  The complex contraction y = contrib (N x K) @ W (K x Co) (nn/field_conv.py:10-33,137)
// is carried as a real GEMM on the interleaved (re,im) storage:
//   [N x 2K] @ [2K x 2Co],  B[(k,re),(o,re)] = Wr, B[(k,re),(o,im)] = Wi,
//                           B[(k,im),(o,re)] = -Wi, B[(k,im),(o,im)] = Wr
// which costs exactly the 8 K Co real flops per vertex of the complex product.
//
// Tiling: 128 x (16*TN) output tile per 256-thread CTA, 8 x TN register tile per thread,
// BK = 8 slices staged through shared memory with register prefetch of the next slice.
// trans_a = 1 reads A as (K x M) row-major (used for gW = contrib^H gy, reduction over vertices),
// split_k > 1 writes per-split partial tiles that are then summed in split order (deterministic).
#include "common.cuh"

namespace fcb {

constexpr int64_t TC_MAX_ACC_MMAS_GROUPED = 400;   // same budget as gemm_tc_plan (gemm_tc.cu)
constexpr int G_BM = 128;
constexpr int G_BK = 8;
constexpr int G_PAD = 4;

template <int TN, bool TRANS_A>
__global__ void __launch_bounds__(256) k_gemm(const float* __restrict__ A, const float* __restrict__ Bm,
                                              float* __restrict__ C, int64_t M, int N, int64_t K, int64_t lda,
                                              int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int split_k,
                                              int64_t k_per_split, int64_t part_stride) {
    constexpr int BN = 16 * TN;
    __shared__ __align__(16) float As[G_BK][G_BM + G_PAD];
    __shared__ __align__(16) float Bs[G_BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int batch = blockIdx.z / split_k;
    const int split = blockIdx.z - batch * split_k;
    const int64_t m0 = (int64_t)blockIdx.x * G_BM;
    const int n0 = blockIdx.y * BN;
    A += batch * sa;
    Bm += batch * sb;
    if (split_k > 1) C += ((int64_t)split * gridDim.z / split_k + batch) * part_stride;  // partial buffer [split][batch]
    else C += batch * sc;
    const int64_t kb = (int64_t)split * k_per_split;
    const int64_t ke = min(K, kb + k_per_split);

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // global->register staging
    float4 ra;          // A: one float4 per thread per slice
    float4 rb;          // B: one float4 per thread (threads < BK*BN/4)
    constexpr int B_F4 = G_BK * BN / 4;
    const int a_row = TRANS_A ? (tid >> 5) : (tid >> 1);        // TRANS_A: k index 0..7 ; else m index 0..127
    const int a_col = TRANS_A ? ((tid & 31) << 2) : ((tid & 1) << 2);  // TRANS_A: m offset ; else k offset
    const int b_k = tid / (BN / 4), b_n = (tid % (BN / 4)) << 2;

    auto load_slice = [&](int64_t k0) {
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TRANS_A) {
            const int64_t k = k0 + a_row;
            const int64_t m = m0 + a_col;
            if (k < ke) {
                const float* p = A + k * lda + m;
                if (m + 3 < M) ra = *reinterpret_cast<const float4*>(p);
                else {
                    if (m < M) ra.x = p[0];
                    if (m + 1 < M) ra.y = p[1];
                    if (m + 2 < M) ra.z = p[2];
                }
            }
        } else {
            const int64_t m = m0 + a_row;
            const int64_t k = k0 + a_col;
            if (m < M) {
                const float* p = A + m * lda + k;
                if (k + 3 < ke) ra = *reinterpret_cast<const float4*>(p);
                else {
                    if (k < ke) ra.x = p[0];
                    if (k + 1 < ke) ra.y = p[1];
                    if (k + 2 < ke) ra.z = p[2];
                }
            }
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < B_F4) {
            const int64_t k = k0 + b_k;
            const int n = n0 + b_n;
            if (k < ke && n < N) {
                const float* p = Bm + k * ldb + n;
                if (n + 3 < N) rb = *reinterpret_cast<const float4*>(p);
                else {
                    rb.x = p[0];
                    if (n + 1 < N) rb.y = p[1];
                    if (n + 2 < N) rb.z = p[2];
                }
            }
        }
    };
    auto store_slice = [&]() {
        if (TRANS_A) {
            *reinterpret_cast<float4*>(&As[a_row][a_col]) = ra;
        } else {
            As[a_col + 0][a_row] = ra.x;
            As[a_col + 1][a_row] = ra.y;
            As[a_col + 2][a_row] = ra.z;
            As[a_col + 3][a_row] = ra.w;
        }
        if (tid < B_F4) *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = rb;
    };

    if (kb < ke) load_slice(kb);
    for (int64_t k0 = kb; k0 < ke; k0 += G_BK) {
        store_slice();
        __syncthreads();
        if (k0 + G_BK < ke) load_slice(k0 + G_BK);
#pragma unroll
        for (int kk = 0; kk < G_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n < N) C[m * ldc + n] = acc[i][j];
        }
    }
}

// C[b][m][n] = sum_s partials[s][b][m][n], fixed order
__global__ void k_reduce_splits(const float* __restrict__ partials, float* __restrict__ C, int64_t M, int N, int64_t ldc,
                                int64_t sc, int batch, int split_k) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = M * N;
    if (i >= per * batch) return;
    const int b = (int)(i / per);
    const int64_t r = i - (int64_t)b * per;
    const int64_t m = r / N;
    const int n = (int)(r - m * N);
    float s = 0.f;
    for (int k = 0; k < split_k; ++k) s += partials[((int64_t)k * batch + b) * per + r];
    C[b * sc + m * ldc + n] = s;
}

// Few outputs, many splits (the TN products over the vertex axis): 8 lanes share an output, lane l sums the splits
// l, l + 8, ... and the eight sums are added in lane order — still a fixed order, 8x shorter dependent chains.
__global__ void __launch_bounds__(256) k_reduce_splits_wide(const float* __restrict__ partials, float* __restrict__ C, int64_t M,
                                                            int N, int64_t ldc, int64_t sc, int batch, int split_k) {
    __shared__ float red[8][33];
    const int ox = threadIdx.x & 31, l = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + ox;
    const int64_t per = M * N;
    const bool live = i < per * batch;
    const int b = live ? (int)(i / per) : 0;
    const int64_t r = i - (int64_t)b * per;
    float s = 0.f;
    if (live)
        for (int k = l; k < split_k; k += 8) s += partials[((int64_t)k * batch + b) * per + r];
    red[l][ox] = s;
    __syncthreads();
    if (l == 0 && live) {
        float t = red[0][ox];
#pragma unroll
        for (int j = 1; j < 8; ++j) t += red[j][ox];
        const int64_t m = r / N;
        C[b * sc + m * ldc + (int)(r - m * N)] = t;
    }
}

static int reduce_splits(const float* partials, float* C, int64_t M, int N, int64_t ldc, int64_t sc, int batch, int split_k,
                         cudaStream_t st) {
    const int64_t tot = M * (int64_t)N * batch;
    if (tot == 0) return FCB_OK;
    if (split_k >= 16 && tot <= 32768) {
        FCB_LAUNCH("reduce_splits", st, k_reduce_splits_wide<<<(unsigned)((tot + 31) / 32), 256, 0, st>>>(partials, C, M, N, ldc, sc, batch, split_k));
    } else {
        FCB_LAUNCH("reduce_splits", st, k_reduce_splits<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(partials, C, M, N, ldc, sc, batch, split_k));
    }
    return FCB_OK;
}

template <bool TRANS_A>
static int dispatch_gemm(const float* A, const float* Bm, float* C, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb,
                         int64_t ldc, int batch, int64_t sa, int64_t sb, int64_t sc, int split_k, int64_t kps,
                         int64_t part_stride, cudaStream_t st) {
    int tn = (N + 15) / 16;
    if (tn > 8) tn = 8;
    if (tn == 5) tn = 6;
    if (tn == 7) tn = 8;
    if (tn == 3) tn = 4;
    const int bn = 16 * tn;
    dim3 grid((unsigned)((M + G_BM - 1) / G_BM), (unsigned)((N + bn - 1) / bn), (unsigned)(batch * split_k));
#define FCB_GEMM_CASE(T) \
    case T: k_gemm<T, TRANS_A><<<grid, 256, 0, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, sa, sb, sc, split_k, kps, part_stride); break;
    prof_begin(TRANS_A ? "gemm_tn" : "gemm_nn", st);
    switch (tn) {
        FCB_GEMM_CASE(1)
        FCB_GEMM_CASE(2)
        FCB_GEMM_CASE(4)
        FCB_GEMM_CASE(6)
        FCB_GEMM_CASE(8)
        default: set_error("gemm: internal tile selection error"); return FCB_E_ARG;
    }
#undef FCB_GEMM_CASE
    prof_end(st);
    FCB_CUDA_LAUNCH_CHECK("gemm");
    return FCB_OK;
}

static bool mode_is_tc(int mode) { return mode == FCB_GEMM_TC_3XTF32 || mode == FCB_GEMM_TC_TF32 || mode == FCB_GEMM_TC_2XF16; }

// vertices (TN) / reals (NN) per stage and per MMA K-step of the tensor-core kernels
static int tc_stage(int mode) { return mode == FCB_GEMM_TC_2XF16 ? 64 : 32; }
static int tc_kstep(int mode) { return mode == FCB_GEMM_TC_2XF16 ? 16 : 8; }

static int64_t tc_ksteps(int64_t K, int trans_a, int split_k, int mode) {
    const int st = tc_stage(mode), ks = tc_kstep(mode);
    if (!trans_a) return (K + st - 1) / st * (st / ks);
    int64_t kps = (K + split_k - 1) / split_k;
    kps = (kps + st - 1) / st * st;
    return kps / ks;
}

// accumulation plan of the selected tensor-core kernel: column-chunk width (0: not feasible) and accumulator count
static int tc_plan(int N, int64_t ksteps, int mode, int trans_a, int* n_acc, int* split = nullptr) {
    int dummy = 1;
    if (mode == FCB_GEMM_TC_2XF16) return trans_a ? gemm_tc_plan(N, ksteps, FCB_GEMM_TC_3XTF32, n_acc) : gemm_h_plan_nn(N, ksteps, n_acc, split ? split : &dummy);
    return gemm_tc_plan(N, ksteps, mode, n_acc);
}

// split-K partials of the 2xFP16 NN kernel's wide-output plan (N > 128): [split][batch][M][chunk] floats, else 0
size_t gemm_h_nn_parts_bytes(int64_t M, int N, int64_t K, int batch) {
    int n_pairs = 1, split = 1;
    const int chunk = gemm_h_plan_nn(N, tc_ksteps(K, 0, 1, FCB_GEMM_TC_2XF16), &n_pairs, &split);
    if (chunk <= 0 || split <= 1) return 0;
    return align_up((size_t)split * batch * M * chunk * 4, 256);
}

// tensor cores are used when the mode asks for them and the accumulation plan fits TMEM (gemm_tc_plan)
static bool use_tc(int N, int64_t K, int trans_a, int batch, int split_k, int flags) {
    const int mode = flags & FCB_GEMM_MASK;
    if (!mode_is_tc(mode) || N <= 0 || K <= 0) return false;
    if (trans_a && batch != 1) return false;
    int n_main;
    return tc_plan(N, tc_ksteps(K, trans_a, split_k, mode), mode, trans_a, &n_main) > 0;
}

size_t gemm_ws_bytes(int64_t M, int N, int64_t K, int trans_a, int batch, int split_k, int flags) {
    const bool h = (flags & FCB_GEMM_MASK) == FCB_GEMM_TC_2XF16;
    if (use_tc(N, K, trans_a, batch, split_k, flags) && !trans_a)
        return h ? gemm_h_ws_bytes(N, K, batch) + gemm_h_nn_parts_bytes(M, N, K, batch) : gemm_tc_ws_bytes(N, K, batch);
    const size_t parts = split_k > 1 ? align_up((size_t)split_k * batch * M * N * 4, 256) : 0;
    if (use_tc(N, K, trans_a, batch, split_k, flags) && trans_a)
        return parts + (h ? gemm_h_tn_ws_bytes(N, K) : gemm_tc_tn_ws_bytes(N, K));   // + packed B
    return parts;
}

int launch_reduce_splits(const float* partials, float* C, int64_t M, int N, int64_t ldc, int64_t sc, int batch, int split_k,
                         cudaStream_t st) {
    return reduce_splits(partials, C, M, N, ldc, sc, batch, split_k, st);
}

// A = [M x groups*Kg] row-major (lda = groups*Kg), Bm = groups matrices [Kg x N] back to back, C = [M x groups*N].
int launch_gemm_grouped(const float* A, const float* Bm, float* C, int64_t M, int N, int64_t Kg, int groups, int flags,
                        const float* a_amax, void* ws, size_t ws_bytes, int* done, cudaStream_t st, const float* sa_x, float* sa_gx) {
    *done = 0;
    const int mode = flags & FCB_GEMM_MASK;
    if (!mode_is_tc(mode) || groups < 2) return FCB_OK;
    const int npad = (N + 15) / 16 * 16;
    if (mode == FCB_GEMM_TC_2XF16) {
        const int64_t mmas_h = Kg / 64 * 4 * 3;                                  // accumulating MMAs per accumulator
        const bool packed = (flags & FCB_FLAG_A_PACKED) != 0;
        if (Kg % 64 != 0 || N > 128 || npad * groups > 512 || mmas_h > TC_MAX_ACC_MMAS_GROUPED || ((groups * (int64_t)N) % 4) != 0 ||
            ((groups * Kg) % 4) != 0 || !aligned16(A) || !aligned16(C) || ws_bytes < gemm_h_ws_bytes(N, Kg, groups)) {
            FCB_REQUIRE(!packed && !(flags & FCB_FLAG_B_PREPACKED), FCB_E_ARG, "gemm_grouped: packed operand with an infeasible grouped shape");
            return FCB_OK;
        }
        FCB_REQUIRE(!packed || a_amax, FCB_E_ARG, "gemm_grouped: a packed A operand needs its scale");
        if (!a_amax) {
            float* slot = gemm_h_amax_slot(ws);
            int rc = launch_absmax_f32(A, M, (int)(groups * Kg), groups * Kg, 1, 0, slot, st);
            if (rc) return rc;
            a_amax = slot;
        }
        const bool sa = sa_gx && sa_x && groups >= 3 && groups <= 7 && (groups & 1) && (N % 8) == 0 && aligned16(sa_x) && aligned16(sa_gx);
        int rc = launch_gemm_h_nn(A, Bm, C, M, N, Kg, groups * Kg, N, groups * (int64_t)N, 1, 0, Kg * N, 0, 1, groups, a_amax, ws,
                                  ws_bytes, packed ? 1 : 0, st, 1, nullptr, nullptr, sa ? sa_x : nullptr, sa ? sa_gx : nullptr,
                                  (flags & FCB_FLAG_B_PREPACKED) ? 1 : 0);
        if (rc == FCB_OK) *done = sa ? 2 : 1;
        return rc;
    }
    const int64_t chunks = Kg / 32;
    const int64_t mmas = chunks * 4 * (mode == FCB_GEMM_TC_3XTF32 ? 3 : 1);     // accumulating MMAs per accumulator
    if (Kg % 32 != 0 || npad * groups > 512 || mmas > TC_MAX_ACC_MMAS_GROUPED || ((groups * (int64_t)N) % 4) != 0 ||
        ((groups * Kg) % 4) != 0 || !aligned16(A) || !aligned16(C) || ws_bytes < gemm_tc_ws_bytes(N, Kg, groups))
        return FCB_OK;
    int rc = launch_gemm_tc_nn(A, Bm, C, M, N, Kg, groups * Kg, N, groups * (int64_t)N, 1, 0, Kg * N, 0, mode, 1, groups, ws,
                               ws_bytes, st);
    if (rc == FCB_OK) *done = 1;
    return rc;
}

int launch_gemm(const float* A, const float* Bm, float* C, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb,
                int64_t ldc, int trans_a, int batch, int64_t sa, int64_t sb, int64_t sc, int split_k, void* ws,
                size_t ws_bytes, int flags, const float* a_amax, cudaStream_t st, const GemmEpilogue* epi, int* epi_fused,
                const float* b_amax) {
    const bool b_pre = (flags & FCB_FLAG_B_PREPACKED) != 0;
    FCB_REQUIRE(A && (Bm || b_pre) && C, FCB_E_ARG, "gemm: null pointer");
    if (epi_fused) *epi_fused = 0;
    FCB_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 1 && split_k >= 1, FCB_E_ARG, "gemm: bad sizes");
    const int mode = flags & FCB_GEMM_MASK;
    const bool h = mode == FCB_GEMM_TC_2XF16;
    const bool packed = (flags & FCB_FLAG_A_PACKED) != 0;
    if (packed) {
        FCB_REQUIRE(h && batch == 1 && a_amax && (trans_a ? gemm_pk_tn_ok(M, N, K, split_k) : gemm_pk_nn_ok(N, K)), FCB_E_ARG,
                    "gemm: packed A operand not supported for this shape / mode");
        lda = 4; sa = 0;      // unused by the packed kernels; keep the alignment tests below neutral
    }
    if (use_tc(N, K, trans_a, batch, split_k, flags) && !trans_a && (ldc % 4) == 0 && (sc % 4) == 0 && aligned16(C) &&
        (!h || ((lda % 4) == 0 && (sa % 4) == 0 && aligned16(A)))) {
        int n_main = 1, h_split = 1;
        const int chunk = tc_plan(N, tc_ksteps(K, 0, 1, mode), mode, 0, &n_main, &h_split);
        float* h_parts = nullptr;
        if (h) {
            const size_t parts_b = gemm_h_nn_parts_bytes(M, N, K, batch);
            FCB_REQUIRE(ws && ws_bytes >= gemm_h_ws_bytes(N, K, batch) + parts_b, FCB_E_WORKSPACE, "gemm: workspace too small (fcb_gemm_workspace_bytes)");
            if (parts_b) h_parts = reinterpret_cast<float*>(static_cast<char*>(ws) + gemm_h_ws_bytes(N, K, batch));
            if (!h_parts) h_split = 1;
            if (M == 0) return FCB_OK;
            if (!a_amax) {       // operand maximum not supplied by the producer of A: one extra pass over A
                float* slot = gemm_h_amax_slot(ws);
                const bool contiguous = (lda == K) && (batch == 1 || sa == M * K);
                int rc = contiguous ? launch_absmax_f32(A, (int64_t)batch * M, (int)K, K, 1, 0, slot, st)
                                    : launch_absmax_f32(A, M, (int)K, lda, batch, sa, slot, st);
                if (rc) return rc;
                a_amax = slot;
            }
        }
        FCB_REQUIRE(!b_pre || (h && chunk >= N && batch == 1), FCB_E_ARG, "gemm: pre-packed B needs the single-launch 2xFP16 path");
        // the block epilogue rides along when the product is ONE un-split 2xFP16 launch
        const bool fuse_epi = h && epi && chunk >= N && h_split == 1 && batch == 1;
        if (fuse_epi && epi_fused) *epi_fused = 1;
        for (int n0 = 0; n0 < N; n0 += chunk) {      // column chunks (one unless N is wide): same A, offset B and C
            const int nc = N - n0 < chunk ? N - n0 : chunk;
            int rc = h ? launch_gemm_h_nn(A, b_pre ? nullptr : Bm + n0, C + n0, M, nc, K, lda, ldb, ldc, batch, sa, sb, sc, n_main, 1, a_amax,
                                          ws, ws_bytes, packed ? 1 : 0, st, h_split, h_parts, fuse_epi ? epi : nullptr, nullptr, nullptr,
                                          b_pre ? 1 : 0, chunk >= N ? b_amax : nullptr)
                       : launch_gemm_tc_nn(A, Bm + n0, C + n0, M, nc, K, lda, ldb, ldc, batch, sa, sb, sc, mode, n_main, 1, ws,
                                           ws_bytes, st);
            if (rc) return rc;
        }
        return FCB_OK;
    }
    FCB_REQUIRE(!b_pre, FCB_E_ARG, "gemm: pre-packed B reached a path that does not take it");
    float* partials = static_cast<float*>(ws);
    if (use_tc(N, K, trans_a, batch, split_k, flags) && trans_a && (lda % 4) == 0 && aligned16(A)) {
        const size_t parts_bytes = split_k > 1 ? align_up((size_t)split_k * M * N * 4, 256) : 0;
        FCB_REQUIRE(ws && ws_bytes >= parts_bytes + (h ? gemm_h_tn_ws_bytes(N, K) : gemm_tc_tn_ws_bytes(N, K)), FCB_E_WORKSPACE,
                    "gemm: workspace too small (fcb_gemm_workspace_bytes)");
        void* bp_ws = static_cast<char*>(ws) + parts_bytes;
        const size_t bp_bytes = ws_bytes - parts_bytes;
        const int stg = tc_stage(mode);
        int64_t kps_tc = (K + split_k - 1) / split_k;
        kps_tc = (kps_tc + stg - 1) / stg * stg;
        int n_main = 1;
        const int chunk = tc_plan(N, kps_tc / tc_kstep(mode), mode, 1, &n_main);
        if (h && !a_amax && M > 0 && K > 0) {
            float* slot = gemm_h_amax_slot(bp_ws);
            int rc = launch_absmax_f32(A, K, (int)M, lda, 1, 0, slot, st);
            if (rc) return rc;
            a_amax = slot;
        }
        for (int n0 = 0; n0 < N; n0 += chunk) {
            const int nc = N - n0 < chunk ? N - n0 : chunk;
            int rc = h ? launch_gemm_h_tn(A, Bm + n0, C + n0, M, nc, K, lda, ldb, ldc, split_k, kps_tc, partials, n_main, a_amax, bp_ws,
                                          bp_bytes, packed ? 1 : 0, st, b_amax)
                       : launch_gemm_tc_tn(A, Bm + n0, C + n0, M, nc, K, lda, ldb, ldc, split_k, kps_tc, partials, mode, n_main,
                                           bp_ws, bp_bytes, st);
            if (rc) return rc;
            if (split_k > 1) {
                rc = launch_reduce_splits(partials, C + n0, M, nc, ldc, 0, 1, split_k, st);
                if (rc) return rc;
            }
        }
        return FCB_OK;
    }
    FCB_REQUIRE(!packed, FCB_E_ARG, "gemm: packed A operand reached the FP32 path");
    FCB_REQUIRE(split_k == 1 || (partials && ws_bytes >= (size_t)split_k * batch * M * N * 4), FCB_E_WORKSPACE,
                "gemm: split_k > 1 needs a workspace of split_k*batch*M*N floats");
    FCB_REQUIRE((lda % 4) == 0 && (ldb % 4) == 0 && (sa % 4) == 0 && (sb % 4) == 0 && aligned16(A) && aligned16(Bm),
                FCB_E_ALIGN, "gemm: A/B leading dimensions and strides must be multiples of 4 floats, 16-byte aligned");
    FCB_REQUIRE((int64_t)batch * split_k <= 65535, FCB_E_UNSUPPORTED, "gemm: batch*split_k too large");
    if (M == 0 || N == 0) return FCB_OK;
    int64_t kps = (K + split_k - 1) / split_k;
    kps = (kps + G_BK - 1) / G_BK * G_BK;  // keep float4 alignment of the K offsets
    float* out = split_k > 1 ? partials : C;
    const int64_t part_stride = M * (int64_t)N;
    int rc;
    if (split_k > 1) {
        rc = trans_a ? dispatch_gemm<true>(A, Bm, out, M, N, K, lda, ldb, N, batch, sa, sb, 0, split_k, kps, part_stride, st)
                     : dispatch_gemm<false>(A, Bm, out, M, N, K, lda, ldb, N, batch, sa, sb, 0, split_k, kps, part_stride, st);
        if (rc) return rc;
        return reduce_splits(partials, C, M, N, ldc, sc, batch, split_k, st);
    }
    rc = trans_a ? dispatch_gemm<true>(A, Bm, out, M, N, K, lda, ldb, ldc, batch, sa, sb, sc, 1, kps, 0, st)
                 : dispatch_gemm<false>(A, Bm, out, M, N, K, lda, ldb, ldc, batch, sa, sb, sc, 1, kps, 0, st);
    return rc;
}

// 2xFP16 TN accumulation plan for Kv vertices cut into `split` ranges, all N columns in one chunk (false: use the generic path)
bool gemm_h_tn_plan(int N, int64_t Kv, int split, int* n_main, int64_t* k_per_split) {
    if (N <= 0 || N > 256 || Kv <= 0 || split < 1) return false;
    const int mode = FCB_GEMM_TC_2XF16;
    int64_t kps = (Kv + split - 1) / split;
    kps = (kps + tc_stage(mode) - 1) / tc_stage(mode) * tc_stage(mode);
    // all N columns in one accumulator tile: (n_main + 1) * Npad TMEM columns, at most 400 accumulating MMAs each
    const int npad = (N + 15) / 16 * 16;
    const int avail = 512 / npad - 1;
    const int64_t need = (kps / tc_kstep(mode) + 399) / 400;
    if (avail < 1 || need > avail) return false;
    int nm = avail < 3 ? avail : 3;
    if (nm < need) nm = (int)need;
    *n_main = nm;
    *k_per_split = kps;
    return true;
}
// vertices one split of that product may cover (its accumulators are limited to 400 accumulating MMAs of 16 vertices)
int64_t gemm_h_tn_max_vertices_per_split(int N) {
    const int npad = (N + 15) / 16 * 16;
    const int avail = 512 / npad - 1;
    return avail < 1 ? 0 : (int64_t)(avail < 3 ? avail : 3) * 400 * 16;
}

// PK operands: the same feasibility tests the dispatchers above apply (shape only; the buffers are the library's own)
bool gemm_pk_nn_ok(int N, int64_t K) {
    return (K % 64) == 0 && (N % 4) == 0 && use_tc(N, K, 0, 1, 1, FCB_GEMM_TC_2XF16);
}
bool gemm_pk_tn_ok(int64_t Mr, int N, int64_t Kv, int split) {
    return (Mr % 64) == 0 && use_tc(N, Kv, 1, 1, split < 1 ? 1 : split, FCB_GEMM_TC_2XF16);
}
bool gemm_h_single_launch(int N, int64_t K, int flags) {
    const int mode = flags & FCB_GEMM_MASK;
    if (mode != FCB_GEMM_TC_2XF16 || !use_tc(N, K, 0, 1, 1, flags)) return false;
    int n_main = 1, h_split = 1;
    return tc_plan(N, tc_ksteps(K, 0, 1, mode), mode, 0, &n_main, &h_split) >= N;
}
bool gemm_pk_grouped_ok(int N, int64_t Kg, int groups) {
    const int npad = (N + 15) / 16 * 16;
    const int64_t mmas_h = Kg / 64 * 4 * 3;
    return groups >= 2 && Kg % 64 == 0 && N <= 128 && npad * groups <= 512 && mmas_h <= TC_MAX_ACC_MMAS_GROUPED &&
           ((groups * (int64_t)N) % 4) == 0;
}

}  // namespace fcb

extern "C" int fcb_gemm_tc_feasible(int N, int64_t K, int trans_a, int split_k, int flags) {
    return fcb::use_tc(N, K, trans_a, 1, split_k < 1 ? 1 : split_k, flags) ? 1 : 0;
}

extern "C" int fcb_gemm_workspace_bytes(int64_t M, int N, int64_t K, int trans_a, int batch, int split_k, int flags,
                                        size_t* bytes) {
    FCB_REQUIRE(bytes && M >= 0 && N >= 0 && K >= 0 && batch >= 1 && split_k >= 1, FCB_E_ARG, "gemm_workspace: bad arguments");
    *bytes = fcb::gemm_ws_bytes(M, N, K, trans_a, batch, split_k, flags);
    return FCB_OK;
}

extern "C" int fcb_gemm_f32(const float* A, const float* B, float* C, int64_t M, int N, int64_t K, int64_t lda,
                            int64_t ldb, int64_t ldc, int trans_a, int batch, int64_t stride_a, int64_t stride_b,
                            int64_t stride_c, int split_k, const float* a_bound, const float* b_bound, void* workspace,
                            size_t workspace_bytes, int flags, void* stream) {
    fcb::prof_scope_lin(true);
    const int rc = fcb::launch_gemm(A, B, C, M, N, K, lda, ldb, ldc, trans_a, batch, stride_a, stride_b, stride_c, split_k,
                                    workspace, workspace_bytes, flags & ~fcb::FCB_FLAG_A_PACKED, a_bound,
                                    static_cast<cudaStream_t>(stream), nullptr, nullptr, b_bound);
    fcb::prof_scope_lin(false);
    return rc;
}
