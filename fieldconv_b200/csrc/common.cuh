// Shared helpers for libfieldconv_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fieldconv_b200.h"

namespace fcb {

void set_error(const char* fmt, ...);
void count_launch();   // bumps the process-wide kernel-launch counter read by fcb_launch_count()
// optional per-launch CUDA-event timing (fcb_profile_*): no-ops unless enabled
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);
void prof_scope_lin(bool on);   // records of this thread's following launches are named "lin_<kernel>" (fcb_gemm_f32)

// FCB_LAUNCH("name", stream, kernel<<<grid, block, smem, stream>>>(args...));
#define FCB_LAUNCH(name, st, ...)            \
    do {                                     \
        fcb::prof_begin((name), (st));       \
        __VA_ARGS__;                         \
        fcb::prof_end((st));                 \
        FCB_CUDA_LAUNCH_CHECK(name);         \
    } while (0)

#define FCB_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            fcb::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

#define FCB_CUDA_LAUNCH_CHECK(what)                                                   \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        fcb::count_launch();                                                          \
        if (e__ != cudaSuccess) {                                                     \
            fcb::set_error("%s: %s", (what), cudaGetErrorString(e__));                \
            return FCB_E_CUDA;                                                        \
        }                                                                             \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Bump allocator over a caller-provided workspace.
struct Arena {
    char* base;
    size_t cap, off;
    Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

constexpr int FCB_FLAG_A_PACKED = 0x10000;    // internal (launch_gemm / launch_gemm_grouped): A operand is a PK buffer
constexpr int FCB_FLAG_B_PREPACKED = 0x20000; // internal: the B side of the 2xFP16 workspace is already filled (launch_pack_w_h)

constexpr int NBR_BITS = 27;
constexpr uint32_t NBR_MASK = (1u << NBR_BITS) - 1u;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// ---- operand scale of the 2xFP16 contraction (gemm_h.cu; also used by the packing aggregation kernels)
// power-of-two scale from max|x| (bit pattern of a non-negative float): largest entry -> [2^14, 2^15)
__host__ __device__ __forceinline__ uint32_t scale_field(uint32_t amax_bits) {
    const int e = (int)((amax_bits >> 23) & 0xffu);
    if (e == 255) return 127u;                 // inf / NaN operand: scale 1, the result is non-finite as in fp32
    int f = 268 - e;                           // (127 + 14) + (127 - e)
    if (f > 253) f = 253;                      // max|x| < 2^-112 (or 0): everything underflows anyway
    if (f < 1) f = 1;
    return (uint32_t)f;
}
__device__ __forceinline__ float scale_of(const float* amax) { return __uint_as_float(scale_field(__float_as_uint(__ldg(amax))) << 23); }
__device__ __forceinline__ float inv_scale_of(const float* amax) {
    return __uint_as_float((254u - scale_field(__float_as_uint(__ldg(amax)))) << 23);
}

// ---- packed operand format ("PK") written by the packing aggregation kernels and bulk-copied by the 2xFP16 GEMMs:
// a real matrix [rows x cols] (cols % 64 == 0) is stored as [row tile of 128][chunk of 64 columns][plane hi, lo] blocks of
// 16 KB, each block the shared-memory image of a SWIZZLE_128B K-major tile: row r at r*128 B, its 16-byte unit u (8 fp16)
// at unit u ^ (r & 7).  hi = fp16(s*v), lo = fp16(s*v - hi).  Rows past the matrix end inside the last tile are zero.
constexpr int PK_ROWS = 128;
constexpr int PK_COLS = 64;
constexpr uint32_t PK_PLANE_BYTES = PK_ROWS * PK_COLS * 2;      // 16 KB
constexpr uint32_t PK_BLOCK_BYTES = 2 * PK_PLANE_BYTES;         // hi + lo
static inline int64_t pk_rows_padded(int64_t rows) { return (rows + PK_ROWS - 1) / PK_ROWS * PK_ROWS; }
static inline size_t pk_bytes(int64_t rows, int64_t cols) { return (size_t)(pk_rows_padded(rows) / PK_ROWS) * (size_t)(cols / PK_COLS) * PK_BLOCK_BYTES; }

// Optional fused epilogue of the forward contraction: C receives z = A B (+ res), act = modReLU(z, bias)
// (nn/fc_resnet_block.py:84-88, nn/tangent_nonlin.py:24-35).  Applied inside the 2xFP16 kernel's TMEM -> register epilogue
// when the product is one un-split launch; launch_gemm reports through *fused whether it was, so the caller can run the
// stand-alone kernel otherwise.
struct GemmEpilogue {
    const float* res;      // [M x N], row stride ld, or nullptr
    const float* bias;     // N / 2 floats, or nullptr (no activation output)
    float* act;            // [M x N], row stride ld
    int64_t ld;
    float* act_bound;      // nullable: max_i |act_i| (complex modulus) is atomically folded into it (pre-zeroed by the caller)
};

// internal entry points shared between translation units
int sort_pairs(uint32_t* k_in, uint32_t* v_in, uint32_t* k_out, uint32_t* v_out, int64_t n, int bits,
               void* ws, size_t ws_bytes, cudaStream_t st);
size_t sort_workspace(int64_t n);
// amax (nullable): device float, bit-pattern max of |out| is atomically folded into it (caller zeroes it first)
int launch_aggregate(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, float* out,
                     int64_t N, int C, int B, int R, int transpose, float* amax, cudaStream_t st);
// PK-format output (packing aggregation): out_pk = pk_bytes(N, 2*R*M*C) bytes, *bound = *feat_amax * *norm (operand scale)
int launch_aggregate_packed(const float* feat, const int32_t* rowptr, const void* rec, const float* rot, void* out_pk,
                            int64_t N, int C, int B, int R, int transpose, const float* feat_amax, const float* norm,
                            float* bound, cudaStream_t st);
int launch_plan_norm(const int32_t* rowptr, const void* rec, int64_t N, float* out, cudaStream_t st);
int launch_aggregate_dense(const float* feat, const float* sten, const int32_t* rowptr, const int32_t* nbr,
                           const int32_t* perm, float* out, int64_t N, int C, int B, int R, int transpose,
                           float* amax, cudaStream_t st);
// Real GEMM dispatcher.  flags & FCB_GEMM_MASK selects FP32 FMA or the tcgen05 path (trans_a == 0, N <= 256 only;
// anything else runs on the FMA path).  ws must hold gemm_ws_bytes(...) bytes.
size_t gemm_ws_bytes(int64_t M, int N, int64_t K, int trans_a, int batch, int split_k, int flags);
// flags & FCB_FLAG_A_PACKED: A is a PK buffer (only with FCB_GEMM_TC_2XF16, batch == 1 and gemm_pk_*_ok shapes)
int launch_gemm(const float* A, const float* Bm, float* C, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb,
                int64_t ldc, int trans_a, int batch, int64_t sa, int64_t sb, int64_t sc, int split_k,
                void* ws, size_t ws_bytes, int flags, const float* a_amax, cudaStream_t st, const GemmEpilogue* epi = nullptr,
                int* epi_fused = nullptr, const float* b_amax = nullptr);
size_t gemm_tc_ws_bytes(int N, int64_t K, int batch);
// 2xFP16 tensor-core kernels (gemm_h.cu).  a_amax: device float holding max|A| (from the kernel that produced A).
int gemm_h_plan_nn(int N, int64_t ksteps, int* n_pairs, int* split_k = nullptr);   // split_k: wide outputs (N > 128) allowed
size_t gemm_h_nn_parts_bytes(int64_t M, int N, int64_t K, int batch);              // split-K partials of the wide-output plan
size_t gemm_h_ws_bytes(int N, int64_t K, int batch);
size_t gemm_h_tn_ws_bytes(int N, int64_t Kv);
float* gemm_h_amax_slot(void* ws);      // scratch float inside a gemm_h workspace for max|A| computed by the dispatcher
int launch_absmax_f32(const float* p, int64_t rows, int cols, int64_t ld, int batch, int64_t stride, float* out, cudaStream_t st);
// a_packed != 0: A is a PK buffer (lda / sa ignored; NN: M x kgroups*K, TN: Kv x Mr) and amax_a the scale it was packed with
int launch_gemm_h_nn(const float* A, const float* B, float* C, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb,
                     int64_t ldc, int batch, int64_t sa, int64_t sb, int64_t sc, int n_pairs, int kgroups,
                     const float* amax_a, void* ws, size_t ws_bytes, int a_packed, cudaStream_t st, int split_k = 1,
                     float* parts = nullptr, const GemmEpilogue* epi = nullptr, const float* sa_x = nullptr, float* sa_gx = nullptr,
                     int b_prepacked = 0, const float* b_bound = nullptr);
// folded filter W (Co,Ci,R,M) -> the B side (scale slot + packed operand) of a launch_gemm_h_nn workspace in one launch
int launch_pack_w_h(const float* W, int dir, int Ci, int Co, int R, int M, const float* w_bound, void* ws, size_t ws_bytes,
                    cudaStream_t st);
// whether launch_gemm / launch_gemm_grouped would run this NN product as ONE 2xFP16 launch_gemm_h_nn call over all N columns
// (the condition under which FCB_FLAG_B_PREPACKED may be passed)
bool gemm_h_single_launch(int N, int64_t K, int flags);
int launch_gemm_h_tn(const float* A, const float* B, float* C, int64_t Mr, int N, int64_t Kv, int64_t lda, int64_t ldb,
                     int64_t ldc, int split, int64_t k_per_split, float* parts, int n_main, const float* amax_a, void* bp_ws,
                     size_t bp_bytes, int a_packed, cudaStream_t st, const float* b_bound = nullptr);
// weight gradient from G: P[m][Mr][2Ci] = G_m^T Xh_m for all 2B+1 frequencies in one batched 2xFP16 TN launch (gemm_h.cu);
// G = [Kv x M*Mr] fp32 or PK; the packed xhat operands are built from x.  gemm_h_tn_plan: accumulation plan of that launch.
size_t gemm_h_tn_xhat_ws_bytes(int Ci, int64_t Kv, int M);
int launch_gemm_h_tn_xhat(const float* G, const float* x, float* P, int64_t Mr, int Ci, int band_limit, int64_t Kv, int split,
                          int64_t k_per_split, float* parts, int n_main, const float* amax_g, void* bp_ws, size_t bp_bytes,
                          int a_packed, cudaStream_t st, const float* x_bound = nullptr);
// out = max_i |z_i| (1 + 2^-20) over n complex numbers (the bound the operand-scale logic accepts for z and for xhat)
int launch_bound_modulus(const float* z, int64_t n, float* out, cudaStream_t st);
bool gemm_h_tn_plan(int N, int64_t Kv, int split, int* n_main, int64_t* k_per_split);
int64_t gemm_h_tn_max_vertices_per_split(int N);
// fused forward for band_limit <= 1 (fused_fwd.cu)
bool fused_fwd_ok(int Ci, int Co, int B, int R);
size_t fused_fwd_ws_bytes(int Ci, int Co, int B, int R);
int launch_fused_fwd(const float* x, const float* W, const int32_t* rowptr, const void* rec, const float* rot, const float* norm,
                     float* y, int64_t N, int64_t n_feat, int Ci, int Co, int B, int R, void* ws, size_t ws_bytes, cudaStream_t st);
// whether the 2xFP16 kernels can consume PK operands for these shapes (same tests the dispatchers apply)
bool gemm_pk_nn_ok(int N, int64_t K);
bool gemm_pk_tn_ok(int64_t Mr, int N, int64_t Kv, int split);
bool gemm_pk_grouped_ok(int N, int64_t Kg, int groups);
// column-chunk width + number of hi*hi accumulators for `ksteps` accumulating MMA steps (0: tensor cores not usable)
int gemm_tc_plan(int N, int64_t ksteps, int mode, int* n_main);
size_t gemm_tc_tn_ws_bytes(int N, int64_t Kv);
int launch_gemm_tc_tn(const float* A, const float* B, float* C, int64_t Mr, int N, int64_t Kv, int64_t lda, int64_t ldb,
                      int64_t ldc, int split, int64_t k_per_split, float* parts, int mode, int n_main, void* bp_ws,
                      size_t bp_bytes, cudaStream_t st);
int launch_reduce_splits(const float* partials, float* C, int64_t M, int N, int64_t ldc, int64_t sc, int batch, int split_k,
                         cudaStream_t st);
int launch_gemm_tc_nn(const float* A, const float* B, float* C, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb,
                      int64_t ldc, int batch, int64_t sa, int64_t sb, int64_t sc, int mode, int n_main, int kgroups,
                      void* ws, size_t ws_bytes, cudaStream_t st);
// grad-x contraction gxh[:, m, :] = G[:, m, :] @ Bt[m] for all m in one pass over G when the tensor-core plan allows
// (returns FCB_OK and sets *done = 1), otherwise leaves *done = 0 for the caller to use the batched path
// sa_x / sa_gx (optional, 2xFP16 grouped product only): the softAngle chain rule runs in the epilogue and writes grad x
// ([M x N/2] complex) instead of C; *done = 2 then
int launch_gemm_grouped(const float* A, const float* Bm, float* C, int64_t M, int N, int64_t Kg, int groups, int flags,
                        const float* a_amax, void* ws, size_t ws_bytes, int* done, cudaStream_t st, const float* sa_x = nullptr,
                        float* sa_gx = nullptr);

}  // namespace fcb
