// C ABI glue: weight packing, the softAngle chain rule, modReLU, and the fwd / bwd drivers that
// chain K1 (aggregate) -> K2 (contract) and K4 / K5 for one FieldConv layer.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace fcb {

static thread_local char g_err[512] = "";

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-launch timing (debug/bench facility; not thread-safe, creates CUDA events)
struct ProfRec { const char* name; cudaEvent_t a, b; };
static ProfRec* g_prof = nullptr;
static int g_prof_cap = 0, g_prof_n = 0;
static bool g_prof_on = false;

// launches made on behalf of fcb_gemm_f32 (TangentLin, not the FieldConv contraction) are recorded as "lin_<name>" so
// per-kernel averages of the contraction kernels are not diluted by the small TangentLin GEMMs
static thread_local bool g_prof_lin = false;
static const char* prof_alias(const char* name) {
    static const char* const tab[][2] = {
        {"gemm_h_nn", "lin_gemm_h_nn"}, {"gemm_h_tn", "lin_gemm_h_tn"}, {"gemm_nn", "lin_gemm_nn"}, {"gemm_tn", "lin_gemm_tn"},
        {"gemm_tc_nn", "lin_gemm_tc_nn"}, {"gemm_tc_tn", "lin_gemm_tc_tn"}, {"absmax", "lin_absmax"}, {"pack_b_h", "lin_pack_b_h"},
        {"pack_b_h_tn", "lin_pack_b_h_tn"}, {"pack_b_tc", "lin_pack_b_tc"}, {"pack_b_tn", "lin_pack_b_tn"},
        {"reduce_splits", "lin_reduce_splits"}};
    for (const auto& t : tab)
        if (strcmp(name, t[0]) == 0) return t[1];
    return name;
}
void prof_scope_lin(bool on) { g_prof_lin = on; }

void prof_begin(const char* name, cudaStream_t st) {
    if (!g_prof_on || g_prof_n >= g_prof_cap) return;
    g_prof[g_prof_n].name = g_prof_lin ? prof_alias(name) : name;
    cudaEventRecord(g_prof[g_prof_n].a, st);
}
void prof_end(cudaStream_t st) {
    if (!g_prof_on || g_prof_n >= g_prof_cap) return;
    cudaEventRecord(g_prof[g_prof_n].b, st);
    ++g_prof_n;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ----------------------------------------------------------------------------- weight packing
// W is complex (Co,Ci,R,M) as folded by the host from (zonal, spherical, phase) — nn/field_conv.py:10-33.
// Forward operand: Bw[2k+a][2o+b], k = (r*M + m)*Ci + c (ring-major, channel fastest: contrib's layout).
__global__ void k_pack_w_fwd(const float2* __restrict__ W, float* __restrict__ Bw, int Ci, int Co, int R, int M) {
    const int64_t K = (int64_t)R * Ci * M;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Co) return;
    const int o = (int)(i % Co);
    const int64_t k = i / Co;
    const int c = (int)(k % Ci);
    const int m = (int)((k / Ci) % M);
    const int r = (int)(k / ((int64_t)M * Ci));
    const float2 w = W[(((int64_t)o * Ci + c) * R + r) * M + m];
    float* row0 = Bw + (2 * k) * (2 * (int64_t)Co) + 2 * o;
    float* row1 = row0 + 2 * (int64_t)Co;
    row0[0] = w.x;  row0[1] = w.y;
    row1[0] = -w.y; row1[1] = w.x;
}

// Backward (grad x) operand, one matrix per m: Bt[m][2q+a][2c+b], q = r*Co + o, value conj(W[o,c,r,m]).
__global__ void k_pack_w_bwd(const float2* __restrict__ W, float* __restrict__ Bt, int Ci, int Co, int R, int M) {
    const int64_t Q = (int64_t)R * Co;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q * Ci * M) return;
    const int c = (int)(i % Ci);
    const int64_t q = (i / Ci) % Q;
    const int m = (int)(i / (Ci * Q));
    const int o = (int)(q % Co), r = (int)(q / Co);
    const float2 w = W[(((int64_t)o * Ci + c) * R + r) * M + m];
    float* base = Bt + (int64_t)m * (2 * Q) * (2 * Ci);
    float* row0 = base + (2 * q) * (2 * (int64_t)Ci) + 2 * c;
    float* row1 = row0 + 2 * (int64_t)Ci;
    row0[0] = w.x; row0[1] = -w.y;
    row1[0] = w.y; row1[1] = w.x;
}

// P = contrib_real^T @ gy_real  ([2K x 2Co])  ->  gW[o,c,r,m] = sum_n conj(contrib) * gy
__global__ void k_combine_gw(const float* __restrict__ P, float2* __restrict__ gW, int Ci, int Co, int R, int M) {
    const int64_t K = (int64_t)R * Ci * M;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Co) return;
    const int o = (int)(i % Co);
    const int64_t k = i / Co;
    const int c = (int)(k % Ci);
    const int m = (int)((k / Ci) % M);
    const int r = (int)(k / ((int64_t)M * Ci));
    const float* row0 = P + (2 * k) * (2 * (int64_t)Co) + 2 * o;
    const float* row1 = row0 + 2 * (int64_t)Co;
    gW[(((int64_t)o * Ci + c) * R + r) * M + m] = make_float2(row0[0] + row1[1], row0[1] - row1[0]);
}

// Weight gradient from G (no contrib in the backward): P[m][(r,o,a)][(c,b)] = sum_j G[j,m,r,o]_a xhat[j,c,m]_b  ->
//   gW[o,c,r,m] = sum_j conj(xhat[j,c,m]) G[j,m,r,o] = (P_rr + P_ii) + i (P_ir - P_ri)   (first index: part of G)
__global__ void k_combine_gw_g(const float* __restrict__ P, float2* __restrict__ gW, int Ci, int Co, int R, int M) {
    const int64_t tot = (int64_t)R * Ci * M * Co;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tot) return;
    const int c = (int)(i % Ci);
    const int o = (int)((i / Ci) % Co);
    const int r = (int)((i / ((int64_t)Ci * Co)) % R);
    const int m = (int)(i / ((int64_t)Ci * Co * R));
    const int64_t q = (int64_t)r * Co + o;
    const float* row0 = P + ((int64_t)m * 2 * R * Co + 2 * q) * (2 * (int64_t)Ci) + 2 * c;   // G real part
    const float* row1 = row0 + 2 * (int64_t)Ci;                                               // G imaginary part
    gW[(((int64_t)o * Ci + c) * R + r) * M + m] = make_float2(row0[0] + row1[1], row1[0] - row0[1]);
}

// xh[m][n][c] = x[n,c] conj(u)^m (nn/field_conv.py:128-130): the explicit operand of the generic (non-2xFP16) gW-from-G path
template <int B>
__global__ void k_xhat(const float2* __restrict__ x, float2* __restrict__ xh, int64_t total) {
    constexpr int M = 2 * B + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float2 z = x[i];
    const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
    const float ri = rsqrtf(z.x * z.x + z.y * z.y);
    const float2 u = origin ? make_float2(1.f, 0.f) : make_float2(z.x * ri, z.y * ri);
    float2 up = z, dn = z;
    xh[(int64_t)B * total + i] = z;
#pragma unroll
    for (int k = 1; k <= B; ++k) {
        up = cmul_conj(up, u);
        dn = cmul(dn, u);
        xh[(int64_t)(B + k) * total + i] = up;
        xh[(int64_t)(B - k) * total + i] = dn;
    }
}

// ----------------------------------------------------------------------------- softAngle chain rule
// gxh is [N][M][Ci] complex (gradient w.r.t. xhat[n,c,m] = x conj(u)^m).  SURVEY.md appendix A.3:
//   non-origin: h_m = conj(g_m) u^(1-m);  gx = u * ( sum_m Re h_m  - i sum_m (1-m) Im h_m )
//   origin    : gx = sum_m g_m            (phi is the constant 0 there: utils/field.py:42-46)
template <int B>
__global__ void k_softangle_bwd(const float2* __restrict__ x, const float2* __restrict__ gxh, float2* __restrict__ gx,
                                int64_t N, int Ci) {
    constexpr int M = 2 * B + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * Ci) return;
    const int64_t n = i / Ci;
    const int c = (int)(i - n * Ci);
    const float2 z = x[i];
    float2 g[M];
#pragma unroll
    for (int m = 0; m < M; ++m) g[m] = gxh[(n * M + m) * Ci + c];
    const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
    float2 out;
    if (origin) {
        out = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < M; ++m) { out.x += g[m].x; out.y += g[m].y; }
    } else {
        const float ri = rsqrtf(z.x * z.x + z.y * z.y);
        const float2 u = make_float2(z.x * ri, z.y * ri);
        // pw[j] = u^(j - B) ... we need u^(1-m) for m = -B..B, i.e. exponents 1+B .. 1-B
        float2 up = u;   // u^1
        float a = 0.f, b = 0.f;
        // m = 0: exponent 1
        {
            const float2 h = cmul(make_float2(g[B].x, -g[B].y), u);
            a += h.x; b -= h.y;
        }
        float2 pos = u;                      // u^(1+k) built upward for m = -k
        float2 neg = u;                      // u^(1-k) built downward for m = +k
#pragma unroll
        for (int k = 1; k <= B; ++k) {
            pos = cmul(pos, u);              // u^(1+k)
            neg = cmul_conj(neg, u);         // u^(1-k)
            const float2 hm = cmul(make_float2(g[B - k].x, -g[B - k].y), pos);   // m = -k, (1-m) = 1+k
            const float2 hp = cmul(make_float2(g[B + k].x, -g[B + k].y), neg);   // m = +k, (1-m) = 1-k
            a += hm.x + hp.x;
            b -= (float)(1 + k) * hm.y + (float)(1 - k) * hp.y;
        }
        (void)up;
        out = cmul(u, make_float2(a, b));
    }
    gx[i] = out;
}

// ----------------------------------------------------------------------------- modReLU (TangentNonLin)
// nn/tangent_nonlin.py:24-35: y = relu(|x| + b_c) * x/|x|, origin entries passed through.
__global__ void k_modrelu_fwd(const float2* __restrict__ x, const float* __restrict__ bias, float2* __restrict__ y,
                              int64_t total, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float2 z = x[i];
    const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
    float2 out = z;
    if (!origin) {
        const float n2 = z.x * z.x + z.y * z.y;
        const float ri = rsqrtf(n2);
        const float mag = n2 * ri;
        const float s = fmaxf(mag + bias[i % C], 0.f);
        out = make_float2(s * z.x * ri, s * z.y * ri);
    }
    y[i] = out;
}

// z = y (+ res) in place, act = modReLU(z, bias): the stand-alone form of the block epilogue for the contraction paths that
// cannot fuse it (FP32-FMA, 3xTF32, split / chunked 2xFP16 products)
__global__ void k_res_modrelu(float2* __restrict__ y, const float2* __restrict__ res, const float* __restrict__ bias,
                              float2* __restrict__ act, int64_t total, int C, uint32_t* __restrict__ act_bound) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float mod = 0.f;                 // |act_i| (the whole warp stays for the reduction below)
    if (i < total) {
        float2 z = y[i];
        if (res) {
            const float2 r = res[i];
            z.x += r.x; z.y += r.y;
            y[i] = z;
        }
        if (act) {
            const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
            float2 out = z;
            mod = 2e-7f;
            if (!origin) {
                const float n2 = z.x * z.x + z.y * z.y;
                const float ri = rsqrtf(n2);
                mod = fmaxf(n2 * ri + bias[i % C], 0.f);
                const float s = mod * ri;
                out = make_float2(s * z.x, s * z.y);
            }
            act[i] = out;
        }
    }
    if (act && act_bound) {
        const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(mod * 1.000001f));
        if ((threadIdx.x & 31) == 0 && w > *reinterpret_cast<volatile uint32_t*>(act_bound)) atomicMax(act_bound, w);
    }
}

constexpr int MR_ROWS = 256;  // rows per block slab in the backward bias reduction

// gx = u (s' Re t + i (s/rho) Im t), t = conj(u) g ; gb_c = sum_n s' Re t   
__global__ void __launch_bounds__(256) k_modrelu_bwd(const float2* __restrict__ x, const float* __restrict__ bias,
                                                     const float2* __restrict__ gy, float2* __restrict__ gx,
                                                     float* __restrict__ gb_part, int64_t N, int C, uint32_t* __restrict__ gx_bound) {
    extern __shared__ float red[];   // [lanes_per_col][C]
    const int rl_count = max(1, 256 / C);
    const int rl = threadIdx.x / C, c = threadIdx.x - rl * C;
    const int64_t r0 = (int64_t)blockIdx.x * MR_ROWS;
    float part = 0.f;
    float gmx = 0.f;                             // max |component| of what this thread writes to gx
    if (rl < rl_count) {
        for (int cc = c; cc < C; cc += 256) {   // C > 256: a thread strides channels
            const float bc = bias[cc];
            float acc = 0.f;
            const int64_t r_end = min(N, r0 + MR_ROWS);
#pragma unroll 4
            for (int64_t r = r0 + rl; r < r_end; r += rl_count) {
                const int64_t i = r * C + cc;
                const float2 z = x[i];
                const float2 g = gy[i];
                const bool origin = (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f);
                float2 out = g;
                if (!origin) {
                    const float n2 = z.x * z.x + z.y * z.y;
                    const float ri = rsqrtf(n2);
                    const float mag = n2 * ri;
                    const float2 u = make_float2(z.x * ri, z.y * ri);
                    const float2 t = cmul_conj(g, u);  // g * conj(u)
                    const float pre = mag + bc;
                    const float sp = pre > 0.f ? 1.f : 0.f;
                    const float s = fmaxf(pre, 0.f);
                    out = cmul(u, make_float2(sp * t.x, s * ri * t.y));
                    acc += sp * t.x;
                }
                gx[i] = out;
                gmx = fmaxf(gmx, fmaxf(fabsf(out.x), fabsf(out.y)));
            }
            if (C > 256) gb_part[(int64_t)blockIdx.x * C + cc] = acc; else part = acc;
        }
    }
    __shared__ uint32_t s_mx;                    // block maximum first: one global atomic per block
    if (gx_bound) {
        if (threadIdx.x == 0) s_mx = 0u;
        __syncthreads();
        const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(gmx));
        if ((threadIdx.x & 31) == 0 && w != 0u) atomicMax(&s_mx, w);
    }
    if (C <= 256) {
        if (rl < rl_count) red[rl * C + c] = part;
        __syncthreads();
        if (threadIdx.x < C) {
            float s = 0.f;
            for (int k = 0; k < rl_count; ++k) s += red[k * C + threadIdx.x];
            gb_part[(int64_t)blockIdx.x * C + threadIdx.x] = s;
        }
    } else {
        __syncthreads();
    }
    if (gx_bound && threadIdx.x == 0 && s_mx > *reinterpret_cast<volatile uint32_t*>(gx_bound)) atomicMax(gx_bound, s_mx);
}

// out[c] = sum_p part[p][c]: 32 lanes of a block share a channel group, lane l of column c sums the slabs l, l + 32, ...,
// the 32 sums are added in lane order (fixed order).
__global__ void __launch_bounds__(1024) k_colsum_parts(const float* __restrict__ part, float* __restrict__ out, int64_t nparts, int C) {
    __shared__ float red[32][33];
    const int cx = threadIdx.x & 31, l = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (c < C)
        for (int64_t p = l; p < nparts; p += 32) s += part[p * C + c];
    red[l][cx] = s;
    __syncthreads();
    if (l == 0 && c < C) {
        float t = red[0][cx];
#pragma unroll
        for (int j = 1; j < 32; ++j) t += red[j][cx];
        out[c] = t;
    }
}

// ----------------------------------------------------------------------------- layer drivers
struct Dims {
    int64_t N;
    int Ci, Co, B, R, M;
    int64_t K;    // complex contraction length R*Ci*M
    int64_t Kt;   // complex length of the transposed gather R*Co*M
};

static int check_dims(const char* who, int64_t N, int Ci, int Co, int B, int R, Dims* d) {
    FCB_REQUIRE(N >= 0 && Ci > 0 && Co > 0, FCB_E_ARG, "%s: bad sizes", who);
    FCB_REQUIRE(B >= 0 && B <= FCB_MAX_BAND_LIMIT, FCB_E_UNSUPPORTED, "%s: band_limit %d unsupported (max %d)", who, B, FCB_MAX_BAND_LIMIT);
    FCB_REQUIRE(R >= 1 && R <= FCB_MAX_RINGS, FCB_E_UNSUPPORTED, "%s: n_rings %d unsupported", who, R);
    FCB_REQUIRE((Ci & 1) == 0 && (Co & 1) == 0, FCB_E_ALIGN, "%s: channel counts must be even (pad with a zero channel)", who);
    FCB_REQUIRE(N <= FCB_MAX_VERTICES, FCB_E_UNSUPPORTED, "%s: N too large", who);
    d->N = N; d->Ci = Ci; d->Co = Co; d->B = B; d->R = R; d->M = 2 * B + 1;
    d->K = (int64_t)R * Ci * d->M;
    d->Kt = (int64_t)R * Co * d->M;
    return FCB_OK;
}

// Vertex ranges of the weight-gradient reduction: enough CTAs for (at most) four full waves on 148 SMs, and short
// enough that each range's accumulating MMA steps fit three TMEM accumulators of the tensor-core plan (400 steps of
// 8 vertices each: gemm_tc_plan) so the 3xTF32 kernel keeps its full 128-column tiles.
static int choose_split(int64_t rows_m, int64_t kdim, int flags) {
    const int64_t tiles = (rows_m + 127) / 128;
    int64_t s = (4 * 148) / tiles;
    const int64_t per_mma = (flags & FCB_GEMM_MASK) == FCB_GEMM_TC_2XF16 ? 16 : 8;     // vertices per accumulating MMA
    const int64_t s_acc = (kdim + 3 * 400 * per_mma - 1) / (3 * 400 * per_mma);
    if (s < s_acc) s = s_acc;
    const int64_t cap = kdim / 512;
    if (s > cap) s = cap;
    if (s > 4096) s = 4096;
    if (s < 1) s = 1;
    return (int)s;
}

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

// gW from G: 2B+1 products P_m[2RCo x 2Ci] = G_m^T Xh_m over the vertices.  Fast path: one batched 2xFP16 TN launch.
struct GwPlan {
    bool batched;          // the 2xFP16 batched launch is feasible
    int split, n_main;
    int64_t kps;
    size_t xhat_ws;        // packed (batched) or explicit fp32 (generic) xhat operand
    size_t parts;          // split partials (+ per-product GEMM workspace on the generic path)
};
static GwPlan gw_from_g_plan(const Dims& d, int flags) {
    GwPlan g;
    const int64_t Mr = 2 * (int64_t)d.R * d.Co;
    const int64_t rows_eff = (int64_t)d.M * ((Mr + 127) / 128) * 128;
    g.split = choose_split(rows_eff, d.N, flags);
    if ((flags & FCB_GEMM_MASK) == FCB_GEMM_TC_2XF16) {        // wide outputs (2Ci > 128) leave room for fewer accumulators
        const int64_t cap = gemm_h_tn_max_vertices_per_split(2 * d.Ci);
        if (cap > 0 && (d.N + g.split - 1) / g.split > cap - 64) g.split = (int)((d.N + cap - 65) / (cap - 64));
    }
    g.n_main = 1;
    g.kps = 0;
    g.batched = (flags & FCB_GEMM_MASK) == FCB_GEMM_TC_2XF16 && (Mr % 4) == 0 &&
                gemm_h_tn_plan(2 * d.Ci, d.N, g.split, &g.n_main, &g.kps);
    if (g.batched) {
        g.xhat_ws = gemm_h_tn_xhat_ws_bytes(d.Ci, d.N, d.M);
        g.parts = g.split > 1 ? align_up((size_t)g.split * d.M * Mr * 2 * d.Ci * 4, 256) : 0;
    } else {
        g.xhat_ws = align_up((size_t)d.M * d.N * d.Ci * 8, 256);
        g.parts = align_up(gemm_ws_bytes(Mr, 2 * d.Ci, d.N, 1, 1, g.split, flags & ~FCB_FLAG_A_PACKED), 256);
    }
    return g;
}

// workspace of the forward contraction: packed filter (+ the split-K partials of a wide output, 2 Co > 128)
static size_t fwd_gemm_ws(const Dims& d) {
    return max_sz(gemm_tc_ws_bytes(2 * d.Co, 2 * d.K, 1), gemm_h_ws_bytes(2 * d.Co, 2 * d.K, 1) + gemm_h_nn_parts_bytes(d.N, 2 * d.Co, 2 * d.K, 1));
}
// workspace of the grad-x contraction (grouped, or one product per frequency when 2 Ci is wide)
static size_t gx_gemm_ws(const Dims& d) {
    const int64_t Q2 = 2 * (int64_t)d.R * d.Co;
    return max_sz(gemm_tc_ws_bytes(2 * d.Ci, Q2, d.M), gemm_h_ws_bytes(2 * d.Ci, Q2, d.M) + gemm_h_nn_parts_bytes(d.N, 2 * d.Ci, Q2, d.M));
}

static size_t fwd_ws(const Dims& d) {
    return 256 /* max|contrib| slot */ + align_up((size_t)(4 * d.K * d.Co) * 4, 256) + fwd_gemm_ws(d) + 512;
}

static size_t gw_parts_bytes(const Dims& d, int flags) {
    const int sp = choose_split(2 * d.K, d.N, flags);
    const size_t b = gemm_ws_bytes(2 * d.K, 2 * d.Co, d.N, 1, 1, sp, flags);
    const size_t fp32_parts = (size_t)(4 * d.K * d.Co) * 4 * sp;
    return align_up(b > fp32_parts ? b : fp32_parts, 256) + 256;
}

// from_g: no contrib from the forward -> the weight gradient is taken from G and xhat (gw_from_g_plan)
static size_t bwd_ws(const Dims& d, bool from_g, int flags) {
    size_t s = 512;                                                         // max|contrib|, max|G| slots
    s += align_up((size_t)(4 * d.K * d.Co) * 4, 256);                       // Bt
    s += align_up((size_t)(4 * d.K * d.Co) * 4, 256);                       // P
    if (from_g) {
        const GwPlan g = gw_from_g_plan(d, flags);
        s += g.xhat_ws + g.parts + 512;
    } else {
        s += gw_parts_bytes(d, flags);                                      // split-K partials (+ the packed gy operand)
    }
    const size_t n_pad = (size_t)pk_rows_padded(d.N);                       // PK buffers hold whole 128-row tiles
    s += align_up(n_pad * d.Kt * 8, 256);                                   // G
    s += align_up((size_t)d.N * d.M * d.Ci * 8, 256);                       // gxh
    // packed operand of the tensor-core grad-x GEMM
    s += gx_gemm_ws(d);
    return s + 2048;
}

// Packed-operand (PK) mode, fcb_*_pk_f32: the aggregation kernels write scaled fp16 (hi, lo) tile images and the 2xFP16
// GEMMs bulk-copy them.  contrib is packed when both its consumers (forward contraction, weight gradient) can take it;
// G when the grouped grad-x contraction can.  Pure functions of the layer dimensions: forward and backward agree.
static bool pk_contrib_ok(const Dims& d) {
    const int64_t K2 = 2 * d.K;
    if ((K2 % PK_COLS) != 0 || d.R < 2) return false;
    return gemm_pk_nn_ok(2 * d.Co, K2) && gemm_pk_tn_ok(K2, 2 * d.Co, d.N, choose_split(K2, d.N, FCB_GEMM_TC_2XF16));
}
static bool pk_g_ok(const Dims& d) { return gemm_pk_grouped_ok(2 * d.Ci, 2 * (int64_t)d.R * d.Co, d.M); }

// The forward drivers aggregate into `contrib` first; *amax (max|contrib|, folded by the aggregation kernel) is the
// operand scale of the 2xFP16 contraction.
static float* fwd_amax_slot(void* ws, float* user_slot, cudaStream_t st) {
    if (user_slot) return user_slot;
    float* slot = static_cast<float*>(ws);
    if (cudaMemsetAsync(slot, 0, 4, st) != cudaSuccess) return nullptr;
    return slot;
}

static int apply_epilogue(const Dims& d, float* y, const GemmEpilogue* epi, cudaStream_t st) {
    const int64_t tot = d.N * d.Co;
    if (tot == 0 || (!epi->res && !epi->act)) return FCB_OK;
    FCB_REQUIRE(epi->ld == 2 * (int64_t)d.Co, FCB_E_ARG, "fwd: the block epilogue needs dense (N, Co) residual / activation buffers");
    FCB_LAUNCH("res_modrelu", st, k_res_modrelu<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
                                      reinterpret_cast<float2*>(y), reinterpret_cast<const float2*>(epi->res), epi->bias,
                                      reinterpret_cast<float2*>(epi->act), tot, d.Co, reinterpret_cast<uint32_t*>(epi->act_bound)));
    return FCB_OK;
}

// epi (optional): y receives z = contraction (+ res), epi->act = modReLU(z, bias) — fused into the 2xFP16 kernel's epilogue
// when the product is a single un-split launch, else applied by k_res_modrelu right after
static int contract_fwd(const Dims& d, const float* contrib, const float* amax, const float* W, float* y, void* ws,
                        size_t ws_bytes, int flags, cudaStream_t st, const GemmEpilogue* epi = nullptr, const float* w_bound = nullptr) {
    FCB_REQUIRE(ws_bytes >= fwd_ws(d), FCB_E_WORKSPACE, "fwd: workspace too small");
    Arena ar(ws, ws_bytes);
    ar.take<char>(256);      // the max|contrib| slot (fwd_amax_slot)
    float* Bw = ar.take<float>((size_t)(4 * d.K * d.Co));
    const int64_t tot = d.K * d.Co;
    const size_t tcb = fwd_gemm_ws(d);
    void* tcw = ar.take<char>(tcb);
    int fused = 0;
    int rc;
    if (gemm_h_single_launch(2 * d.Co, 2 * d.K, flags)) {
        // one launch: W -> packed fp16 (hi, lo) operand in the contraction's workspace (no fp32 embedding, no separate max / pack)
        rc = launch_pack_w_h(W, 0, d.Ci, d.Co, d.R, d.M, w_bound, tcw, tcb, st);
        if (rc) return rc;
        rc = launch_gemm(contrib, nullptr, y, d.N, 2 * d.Co, 2 * d.K, 2 * d.K, 2 * d.Co, 2 * d.Co, 0, 1, 0, 0, 0, 1, tcw, tcb,
                         flags | FCB_FLAG_B_PREPACKED, amax, st, epi, &fused);
        if (rc || !epi || fused) return rc;
        return apply_epilogue(d, y, epi, st);
    }
    FCB_LAUNCH("pack_w_fwd", st, k_pack_w_fwd<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(W), Bw, d.Ci, d.Co, d.R, d.M));
    rc = launch_gemm(contrib, Bw, y, d.N, 2 * d.Co, 2 * d.K, 2 * d.K, 2 * d.Co, 2 * d.Co, 0, 1, 0, 0, 0, 1, tcw, tcb, flags, amax, st,
                     epi, &fused);
    if (rc || !epi || fused) return rc;
    return apply_epilogue(d, y, epi, st);
}

// contrib_packed / g_packed: contrib is (G will be) a PK buffer; gather_transpose(G, g_amax, g_packed) fills G either way.
// contrib == nullptr: the forward kept nothing — the weight gradient comes from G and xhat (k_combine_gw_g), so the
// backward neither stores nor recomputes the N x K contrib.
template <typename GatherT>
static int backward_common(const Dims& d, const float* x, const float* W, const float* gy, const float* contrib,
                           const float* contrib_amax, float* g_amax, GatherT&& gather_transpose, float* gx, float* gW,
                           Arena& ar, int flags, cudaStream_t st, bool contrib_packed = false, bool g_packed = false,
                           const float* x_bound = nullptr, const float* w_bound = nullptr) {
    float* Bt = ar.take<float>((size_t)(4 * d.K * d.Co));
    float* P = ar.take<float>((size_t)(4 * d.K * d.Co));
    const bool from_g = gW && !contrib;
    const GwPlan gp = from_g ? gw_from_g_plan(d, flags) : GwPlan();
    const int gw_split = from_g ? gp.split : choose_split(2 * d.K, d.N, flags);
    const size_t parts_bytes = from_g ? gp.parts + 256 : gw_parts_bytes(d, flags);
    void* xhat_ws = from_g ? static_cast<void*>(ar.take<char>(gp.xhat_ws + 256)) : nullptr;
    float* parts = reinterpret_cast<float*>(ar.take<char>(parts_bytes));
    float* G = ar.take<float>((size_t)pk_rows_padded(d.N) * d.Kt * 2);
    float* gxh = ar.take<float>((size_t)d.N * d.M * d.Ci * 2);
    FCB_REQUIRE(ar.ok(), FCB_E_WORKSPACE, "bwd: workspace too small");
    const int64_t tot = d.K * d.Co;
    if (gW && !from_g) {
        // K4: P[2K x 2Co] = contrib_real^T @ gy_real, split over vertices, fixed-order reduction
        int rc = launch_gemm(contrib, gy, P, 2 * d.K, 2 * d.Co, d.N, 2 * d.K, 2 * d.Co, 2 * d.Co, 1, 1, 0, 0, 0, gw_split, parts,
                             parts_bytes, flags | (contrib_packed ? FCB_FLAG_A_PACKED : 0), contrib_amax, st);
        if (rc) return rc;
        FCB_LAUNCH("combine_gw", st, k_combine_gw<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(P, reinterpret_cast<float2*>(gW), d.Ci, d.Co, d.R, d.M));
    }
    if (gx || from_g) {
        // K5a: G[j][m][r][o] = sum_{e: src=j} conj(sten) gy[tgt]
        if (cudaMemsetAsync(g_amax, 0, 4, st) != cudaSuccess) {
            set_error("bwd: cudaMemsetAsync failed");
            return FCB_E_CUDA;
        }
        int rc = gather_transpose(G, g_amax, g_packed);
        if (rc) return rc;
    }
    if (gx) {
        // K5b: gxh[:, m, :] = G[:, m, :] @ conj(W)[m]  (one real GEMM per m, batched)
        const int64_t Q2 = 2 * (int64_t)d.R * d.Co;
        const size_t tcb = gx_gemm_ws(d);
        void* tcw = ar.take<char>(tcb);
        FCB_REQUIRE(ar.ok(), FCB_E_WORKSPACE, "bwd: workspace too small");
        int grouped = 0;     // all m in one pass over G (one long-K tensor-core pipeline) when the plan allows
        // the grouped 2xFP16 product takes conj(W) packed straight from W (one launch) when it is certain to run
        const bool pre = (flags & FCB_GEMM_MASK) == FCB_GEMM_TC_2XF16 && gemm_pk_grouped_ok(2 * d.Ci, Q2, d.M) &&
                         tcb >= gemm_h_ws_bytes(2 * d.Ci, Q2, d.M) && aligned16(G) && aligned16(gxh);
        int rc;
        if (pre) {
            rc = launch_pack_w_h(W, 1, d.Ci, d.Co, d.R, d.M, w_bound, tcw, tcb, st);
            if (rc) return rc;
        } else {
            FCB_LAUNCH("pack_w_bwd", st, k_pack_w_bwd<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(W), Bt, d.Ci, d.Co, d.R, d.M));
        }
        // 2xFP16 grouped product: the softAngle chain rule runs in its epilogue (grouped == 2) — gxhat never reaches memory
        rc = launch_gemm_grouped(G, pre ? nullptr : Bt, gxh, d.N, 2 * d.Ci, Q2, d.M,
                                 flags | (g_packed ? FCB_FLAG_A_PACKED : 0) | (pre ? FCB_FLAG_B_PREPACKED : 0), g_amax, tcw, tcb, &grouped, st, x, gx);
        if (rc) return rc;
        FCB_REQUIRE(grouped || !pre, FCB_E_ARG, "bwd: pre-packed filter but the grouped contraction did not run");
        FCB_REQUIRE(grouped || !g_packed, FCB_E_ARG, "bwd: packed G but the grouped contraction did not run");
        if (!grouped) {
            rc = launch_gemm(G, Bt, gxh, d.N, 2 * d.Ci, Q2, (int64_t)d.M * Q2, 2 * d.Ci, (int64_t)d.M * 2 * d.Ci, 0, d.M, Q2,
                             Q2 * 2 * d.Ci, 2 * d.Ci, 1, tcw, tcb, flags, g_amax, st);
            if (rc) return rc;
        }
        const int64_t el = d.N * d.Ci;
        const unsigned blocks = (unsigned)((el + 255) / 256);
        const float2* x2 = reinterpret_cast<const float2*>(x);
        const float2* g2 = reinterpret_cast<const float2*>(gxh);
        float2* o2 = reinterpret_cast<float2*>(gx);
        if (el > 0 && grouped != 2) {
            prof_begin("softangle_bwd", st);
            switch (d.B) {
                case 0: k_softangle_bwd<0><<<blocks, 256, 0, st>>>(x2, g2, o2, d.N, d.Ci); break;
                case 1: k_softangle_bwd<1><<<blocks, 256, 0, st>>>(x2, g2, o2, d.N, d.Ci); break;
                case 2: k_softangle_bwd<2><<<blocks, 256, 0, st>>>(x2, g2, o2, d.N, d.Ci); break;
                case 3: k_softangle_bwd<3><<<blocks, 256, 0, st>>>(x2, g2, o2, d.N, d.Ci); break;
                case 4: k_softangle_bwd<4><<<blocks, 256, 0, st>>>(x2, g2, o2, d.N, d.Ci); break;
            }
            prof_end(st);
            FCB_CUDA_LAUNCH_CHECK("softangle_bwd");
        }
    }
    if (from_g) {
        // K4': P[m][2RCo][2Ci] = G_m^T Xh_m, split over vertices, fixed-order reduction; G is read once more instead of contrib
        const int64_t Mr = 2 * (int64_t)d.R * d.Co;
        const int N2 = 2 * d.Ci;
        if (gp.batched) {
            int rc = launch_gemm_h_tn_xhat(G, x, P, Mr, d.Ci, d.B, d.N, gp.split, gp.kps, parts, gp.n_main, g_amax, xhat_ws,
                                           gp.xhat_ws + 256, g_packed ? 1 : 0, st, x_bound);
            if (rc) return rc;
            if (gp.split > 1) {
                rc = launch_reduce_splits(parts, P, (int64_t)d.M * Mr, N2, N2, 0, 1, gp.split, st);
                if (rc) return rc;
            }
        } else {
            FCB_REQUIRE(!g_packed, FCB_E_ARG, "bwd: packed G but the batched weight-gradient product is not feasible");
            float2* xh = static_cast<float2*>(xhat_ws);
            const int64_t el = d.N * d.Ci;
            const unsigned blocks = (unsigned)((el + 255) / 256);
            const float2* x2 = reinterpret_cast<const float2*>(x);
            prof_begin("xhat", st);
            switch (d.B) {
                case 0: k_xhat<0><<<blocks, 256, 0, st>>>(x2, xh, el); break;
                case 1: k_xhat<1><<<blocks, 256, 0, st>>>(x2, xh, el); break;
                case 2: k_xhat<2><<<blocks, 256, 0, st>>>(x2, xh, el); break;
                case 3: k_xhat<3><<<blocks, 256, 0, st>>>(x2, xh, el); break;
                case 4: k_xhat<4><<<blocks, 256, 0, st>>>(x2, xh, el); break;
            }
            prof_end(st);
            FCB_CUDA_LAUNCH_CHECK("xhat");
            for (int m = 0; m < d.M; ++m) {
                int rc = launch_gemm(G + (int64_t)m * Mr, reinterpret_cast<const float*>(xh) + (int64_t)m * d.N * N2, P + (int64_t)m * Mr * N2,
                                     Mr, N2, d.N, (int64_t)d.M * Mr, N2, N2, 1, 1, 0, 0, 0, gp.split, parts, parts_bytes,
                                     flags & ~FCB_FLAG_A_PACKED, g_amax, st);
                if (rc) return rc;
            }
        }
        FCB_LAUNCH("combine_gw", st, k_combine_gw_g<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(P, reinterpret_cast<float2*>(gW), d.Ci, d.Co, d.R, d.M));
    }
    return FCB_OK;
}

}  // namespace fcb

using namespace fcb;

extern "C" int fcb_profile_enable(int max_records) {
    FCB_REQUIRE(max_records > 0 && max_records <= (1 << 20), FCB_E_ARG, "profile_enable: bad record count");
    if (g_prof_cap < max_records) {
        for (int i = 0; i < g_prof_cap; ++i) { cudaEventDestroy(g_prof[i].a); cudaEventDestroy(g_prof[i].b); }
        delete[] g_prof;
        g_prof = new ProfRec[max_records];
        g_prof_cap = max_records;
        for (int i = 0; i < g_prof_cap; ++i) {
            if (cudaEventCreate(&g_prof[i].a) != cudaSuccess || cudaEventCreate(&g_prof[i].b) != cudaSuccess) {
                set_error("profile_enable: cudaEventCreate failed");
                g_prof_cap = i;
                return FCB_E_CUDA;
            }
        }
    }
    g_prof_n = 0;
    g_prof_on = true;
    return FCB_OK;
}

extern "C" int fcb_profile_disable(void) {
    g_prof_on = false;
    return FCB_OK;
}

// After the caller synchronised the stream(s): writes up to `capacity` durations (ms) and the kernel names
// joined by '\n' into names_buf; *count = number of records; resets the record list.
extern "C" int fcb_profile_collect(char* names_buf, size_t names_bytes, float* ms, int capacity, int* count) {
    FCB_REQUIRE(names_buf && ms && count && names_bytes > 0, FCB_E_ARG, "profile_collect: null argument");
    size_t off = 0;
    int n = g_prof_n < capacity ? g_prof_n : capacity;
    names_buf[0] = 0;
    for (int i = 0; i < n; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b) != cudaSuccess) {
            set_error("profile_collect: events not complete (synchronise first)");
            cudaGetLastError();
            return FCB_E_CUDA;
        }
        ms[i] = t;
        const size_t len = strlen(g_prof[i].name);
        if (off + len + 2 > names_bytes) { n = i; break; }
        memcpy(names_buf + off, g_prof[i].name, len);
        off += len;
        names_buf[off++] = '\n';
        names_buf[off] = 0;
    }
    *count = n;
    g_prof_n = 0;
    return FCB_OK;
}

extern "C" const char* fcb_last_error(void) { return g_err; }
extern "C" int fcb_version(void) { return 100; }
extern "C" unsigned long long fcb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int fcb_fwd_workspace_bytes(int64_t N, int Ci, int Co, int band_limit, int R, int flags, size_t* bytes) {
    Dims d;
    (void)flags;
    FCB_REQUIRE(bytes, FCB_E_ARG, "fwd_workspace: null");
    int rc = check_dims("fwd_workspace", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    *bytes = fwd_ws(d);
    return FCB_OK;
}

extern "C" int fcb_bwd_workspace_bytes(int64_t N, int Ci, int Co, int band_limit, int R, int flags, size_t* bytes) {
    Dims d;
    FCB_REQUIRE(bytes, FCB_E_ARG, "bwd_workspace: null");
    int rc = check_dims("bwd_workspace", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    *bytes = bwd_ws(d, !(flags & FCB_FLAG_HAVE_CONTRIB), flags);
    return FCB_OK;
}

static int fwd_impl(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt, const float* rot_tgt, float* y,
                    float* contrib, float* contrib_absmax, int64_t N, int Ci, int Co, int band_limit, int R, int flags, void* ws,
                    size_t ws_bytes, void* stream, const GemmEpilogue* epi, const float* w_bound = nullptr) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Dims d;
    int rc = check_dims("fwd", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    FCB_REQUIRE(R >= 2, FCB_E_UNSUPPORTED, "fwd: n_rings must be >= 2 (reference divides by n_rings-1)");
    FCB_REQUIRE(x && W && rowptr_tgt && rec_tgt && rot_tgt && y && ws, FCB_E_ARG, "fwd: null pointer");
    FCB_REQUIRE(contrib, FCB_E_UNSUPPORTED, "fwd: this path needs a contrib buffer (N*R*Ci*M complex)");
    FCB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(W) && aligned16(contrib), FCB_E_ALIGN, "fwd: pointers must be 16-byte aligned");
    FCB_REQUIRE(ws_bytes >= fwd_ws(d), FCB_E_WORKSPACE, "fwd: workspace too small");
    if (N == 0) return FCB_OK;
    float* amax = fwd_amax_slot(ws, contrib_absmax, st);
    FCB_REQUIRE(amax, FCB_E_CUDA, "fwd: cudaMemsetAsync failed");
    rc = launch_aggregate(x, rowptr_tgt, rec_tgt, rot_tgt, contrib, N, Ci, band_limit, R, 0, amax, st);
    if (rc) return rc;
    return contract_fwd(d, contrib, amax, W, y, ws, ws_bytes, flags, st, epi, w_bound);
}

extern "C" int fcb_fwd_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt,
                           const float* rot_tgt, float* y, float* contrib, float* contrib_absmax, int64_t N, int Ci, int Co,
                           int band_limit, int R, int flags, void* ws, size_t ws_bytes, void* stream) {
    return fwd_impl(x, W, rowptr_tgt, rec_tgt, rot_tgt, y, contrib, contrib_absmax, N, Ci, Co, band_limit, R, flags, ws, ws_bytes,
                    stream, nullptr);
}

static int make_epilogue(const float* res, const float* bias, float* act, int Co, const fcb_bounds* bounds, cudaStream_t st,
                         GemmEpilogue* e) {
    FCB_REQUIRE((bias == nullptr) == (act == nullptr), FCB_E_ARG, "fwd_act: bias and act go together");
    FCB_REQUIRE((!res || aligned16(res)) && (!act || aligned16(act)), FCB_E_ALIGN, "fwd_act: res / act must be 16-byte aligned");
    e->res = res; e->bias = bias; e->act = act; e->ld = 2 * (int64_t)Co;
    e->act_bound = (bounds && act) ? bounds->act : nullptr;
    if (e->act_bound && cudaMemsetAsync(e->act_bound, 0, 4, st) != cudaSuccess) {
        set_error("fwd_act: cudaMemsetAsync failed");
        return FCB_E_CUDA;
    }
    return FCB_OK;
}

// Block epilogue fused into the layer (nn/fc_resnet_block.py:84-88): y = conv(x) + res (res may be NULL),
// act = modReLU(y, bias) (nn/tangent_nonlin.py:24-35; bias / act may both be NULL).  Otherwise as fcb_fwd_f32.
extern "C" int fcb_fwd_act_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt,
                               const float* rot_tgt, float* y, float* contrib, float* contrib_absmax, const float* res,
                               const float* bias, float* act, const fcb_bounds* bounds, int64_t N, int Ci, int Co, int band_limit,
                               int R, int flags, void* ws, size_t ws_bytes, void* stream) {
    GemmEpilogue e;
    int rc = make_epilogue(res, bias, act, Co, bounds, static_cast<cudaStream_t>(stream), &e);
    if (rc) return rc;
    return fwd_impl(x, W, rowptr_tgt, rec_tgt, rot_tgt, y, contrib, contrib_absmax, N, Ci, Co, band_limit, R, flags, ws, ws_bytes,
                    stream, &e, bounds ? bounds->w : nullptr);
}

extern "C" int fcb_bwd_f32(const float* x, const float* W, const float* gy, const float* contrib,
                           const float* contrib_absmax, const int32_t* rowptr_tgt, const void* rec_tgt, const float* rot_tgt,
                           const int32_t* rowptr_src, const void* rec_src, const float* rot_src, float* gx, float* gW,
                           const fcb_bounds* bounds, int64_t N, int Ci, int Co, int band_limit, int R, int flags, void* ws,
                           size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Dims d;
    int rc = check_dims("bwd", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    FCB_REQUIRE(R >= 2, FCB_E_UNSUPPORTED, "bwd: n_rings must be >= 2");
    FCB_REQUIRE(x && W && gy && ws, FCB_E_ARG, "bwd: null pointer");
    FCB_REQUIRE(!gx || (rowptr_src && rec_src && rot_src), FCB_E_ARG, "bwd: grad x needs the by-source plan");
    FCB_REQUIRE(!gW || contrib || (rowptr_src && rec_src && rot_src), FCB_E_ARG, "bwd: grad W needs contrib or the by-source plan");
    (void)rowptr_tgt; (void)rec_tgt; (void)rot_tgt;
    FCB_REQUIRE(aligned16(x) && aligned16(gy) && aligned16(W), FCB_E_ALIGN, "bwd: pointers must be 16-byte aligned");
    const bool from_g = gW && !contrib;          // nothing kept by the forward: gW = xhat^H G, no recompute
    FCB_REQUIRE(ws_bytes >= bwd_ws(d, from_g, flags), FCB_E_WORKSPACE, "bwd: workspace too small");
    if (N == 0) {
        if (gW && cudaMemsetAsync(gW, 0, (size_t)d.K * d.Co * 8, st) != cudaSuccess) {
            set_error("bwd: cudaMemsetAsync failed");
            return FCB_E_CUDA;
        }
        return FCB_OK;
    }
    Arena ar(ws, ws_bytes);
    float* slots = ar.take<float>(128);          // [64] max|G|
    auto gather = [&](float* G, float* g_amax, bool) {
        return launch_aggregate(gy, rowptr_src, rec_src, rot_src, G, N, Co, band_limit, R, 1, g_amax, st);
    };
    return backward_common(d, x, W, gy, contrib, contrib_absmax, slots + 64, gather, gx, gW, ar, flags, st, false, false,
                           bounds ? bounds->x : nullptr, bounds ? bounds->w : nullptr);
}

// ------------------------------------------------------------------ packed-operand (PK) variants
extern "C" int fcb_pk_supported(int64_t N, int Ci, int Co, int band_limit, int R) {
    Dims d;
    if (check_dims("pk_supported", N, Ci, Co, band_limit, R, &d)) return 0;
    return pk_contrib_ok(d) ? 1 : 0;
}

extern "C" int fcb_pk_contrib_bytes(int64_t N, int Ci, int band_limit, int R, size_t* bytes) {
    FCB_REQUIRE(bytes && N >= 0 && Ci > 0 && band_limit >= 0 && R >= 1, FCB_E_ARG, "pk_contrib_bytes: bad arguments");
    const int64_t K2 = 2 * (int64_t)R * Ci * (2 * band_limit + 1);
    FCB_REQUIRE((K2 % PK_COLS) == 0, FCB_E_UNSUPPORTED, "pk_contrib_bytes: 2*R*M*Ci must be a multiple of 64");
    *bytes = pk_bytes(N, K2);
    return FCB_OK;
}

static int fwd_pk_impl(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt, const float* rot_tgt,
                       const float* norm_tgt, float* y, void* contrib_pk, float* contrib_scale, int64_t N, int Ci, int Co,
                       int band_limit, int R, int flags, void* ws, size_t ws_bytes, void* stream, const GemmEpilogue* epi,
                       const float* x_bound, const float* w_bound = nullptr) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Dims d;
    int rc = check_dims("fwd_pk", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    FCB_REQUIRE((flags & FCB_GEMM_MASK) == FCB_GEMM_TC_2XF16, FCB_E_ARG, "fwd_pk: needs FCB_GEMM_TC_2XF16");
    FCB_REQUIRE(x && W && rowptr_tgt && rec_tgt && rot_tgt && norm_tgt && y && contrib_pk && contrib_scale && ws, FCB_E_ARG, "fwd_pk: null pointer");
    FCB_REQUIRE(pk_contrib_ok(d), FCB_E_UNSUPPORTED, "fwd_pk: shape not supported by the packed path (fcb_pk_supported)");
    FCB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(W), FCB_E_ALIGN, "fwd_pk: pointers must be 16-byte aligned");
    FCB_REQUIRE(ws_bytes >= fwd_ws(d), FCB_E_WORKSPACE, "fwd_pk: workspace too small");
    if (N == 0) return FCB_OK;
    const float* x_amax = x_bound;                         // supplied by the producer of x, else one pass over x
    if (!x_amax) {
        float* slot = static_cast<float*>(ws) + 16;        // inside the 256-byte scalar area at the head of the workspace
        rc = launch_absmax_f32(x, N, 2 * Ci, 2 * (int64_t)Ci, 1, 0, slot, st);
        if (rc) return rc;
        x_amax = slot;
    }
    rc = launch_aggregate_packed(x, rowptr_tgt, rec_tgt, rot_tgt, contrib_pk, N, Ci, band_limit, R, 0, x_amax, norm_tgt, contrib_scale, st);
    if (rc) return rc;
    return contract_fwd(d, static_cast<const float*>(contrib_pk), contrib_scale, W, y, ws, ws_bytes, flags | FCB_FLAG_A_PACKED, st, epi,
                        w_bound);
}

extern "C" int fcb_fwd_pk_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt,
                              const float* rot_tgt, const float* norm_tgt, float* y, void* contrib_pk, float* contrib_scale,
                              const fcb_bounds* bounds, int64_t N, int Ci, int Co, int band_limit, int R, int flags, void* ws,
                              size_t ws_bytes, void* stream) {
    return fwd_pk_impl(x, W, rowptr_tgt, rec_tgt, rot_tgt, norm_tgt, y, contrib_pk, contrib_scale, N, Ci, Co, band_limit, R, flags, ws,
                       ws_bytes, stream, nullptr, bounds ? bounds->x : nullptr, bounds ? bounds->w : nullptr);
}

extern "C" int fcb_fwd_act_pk_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt,
                                  const float* rot_tgt, const float* norm_tgt, float* y, void* contrib_pk, float* contrib_scale,
                                  const float* res, const float* bias, float* act, const fcb_bounds* bounds, int64_t N, int Ci,
                                  int Co, int band_limit, int R, int flags, void* ws, size_t ws_bytes, void* stream) {
    GemmEpilogue e;
    int rc = make_epilogue(res, bias, act, Co, bounds, static_cast<cudaStream_t>(stream), &e);
    if (rc) return rc;
    return fwd_pk_impl(x, W, rowptr_tgt, rec_tgt, rot_tgt, norm_tgt, y, contrib_pk, contrib_scale, N, Ci, Co, band_limit, R, flags, ws,
                       ws_bytes, stream, &e, bounds ? bounds->x : nullptr, bounds ? bounds->w : nullptr);
}

extern "C" int fcb_bwd_pk_f32(const float* x, const float* W, const float* gy, const void* contrib_pk,
                              const float* contrib_scale, const int32_t* rowptr_tgt, const void* rec_tgt,
                              const float* rot_tgt, const float* norm_tgt, const int32_t* rowptr_src, const void* rec_src,
                              const float* rot_src, const float* norm_src, float* gx, float* gW, const fcb_bounds* bounds,
                              int64_t N, int Ci, int Co, int band_limit, int R, int flags, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Dims d;
    int rc = check_dims("bwd_pk", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    FCB_REQUIRE((flags & FCB_GEMM_MASK) == FCB_GEMM_TC_2XF16, FCB_E_ARG, "bwd_pk: needs FCB_GEMM_TC_2XF16");
    FCB_REQUIRE(x && W && gy && ws, FCB_E_ARG, "bwd_pk: null pointer");
    FCB_REQUIRE(!gx || (rowptr_src && rec_src && rot_src && norm_src), FCB_E_ARG, "bwd_pk: grad x needs the by-source plan and its norm");
    FCB_REQUIRE(!gW || (contrib_pk && contrib_scale) || (rowptr_src && rec_src && rot_src && norm_src), FCB_E_ARG,
                "bwd_pk: grad W needs the packed contrib + its scale, or the by-source plan and its norm");
    (void)rowptr_tgt; (void)rec_tgt; (void)rot_tgt; (void)norm_tgt;
    FCB_REQUIRE(!(contrib_pk && contrib_scale) || pk_contrib_ok(d), FCB_E_UNSUPPORTED,
                "bwd_pk: shape not supported by the packed path (fcb_pk_supported)");
    FCB_REQUIRE(aligned16(x) && aligned16(gy) && aligned16(W), FCB_E_ALIGN, "bwd_pk: pointers must be 16-byte aligned");
    const bool from_g = gW && !(contrib_pk && contrib_scale);
    FCB_REQUIRE(ws_bytes >= bwd_ws(d, from_g, flags), FCB_E_WORKSPACE, "bwd_pk: workspace too small");
    if (N == 0) {
        if (gW && cudaMemsetAsync(gW, 0, (size_t)d.K * d.Co * 8, st) != cudaSuccess) {
            set_error("bwd_pk: cudaMemsetAsync failed");
            return FCB_E_CUDA;
        }
        return FCB_OK;
    }
    Arena ar(ws, ws_bytes);
    float* slots = ar.take<float>(128);   // [64] G scale / max|G|, [80] max|gy|
    const float* contrib = from_g ? nullptr : static_cast<const float*>(contrib_pk);
    if (from_g) contrib_scale = nullptr;
    const bool g_pk = pk_g_ok(d) && (!from_g || gw_from_g_plan(d, flags).batched);
    auto gather = [&](float* G, float* g_amax, bool packed) {
        if (!packed) return launch_aggregate(gy, rowptr_src, rec_src, rot_src, G, N, Co, band_limit, R, 1, g_amax, st);
        const float* gy_amax = bounds ? bounds->gy : nullptr;       // supplied by the producer of gy, else one pass over gy
        if (!gy_amax) {
            int r2 = launch_absmax_f32(gy, N, 2 * Co, 2 * (int64_t)Co, 1, 0, slots + 80, st);
            if (r2) return r2;
            gy_amax = slots + 80;
        }
        return launch_aggregate_packed(gy, rowptr_src, rec_src, rot_src, G, N, Co, band_limit, R, 1, gy_amax, norm_src, g_amax, st);
    };
    return backward_common(d, x, W, gy, contrib, contrib_scale, slots + 64, gather, gx, gW, ar, flags, st, !from_g, g_pk,
                           bounds ? bounds->x : nullptr, bounds ? bounds->w : nullptr);
}

// ------------------------------------------------------------------ fused forward (band_limit <= 1)
// ------------------------------------------------------------------ operand bounds
extern "C" int fcb_bound_f32(const float* z, int64_t n_complex, float* bound_out, void* stream) {
    FCB_REQUIRE(bound_out && n_complex >= 0 && (z || n_complex == 0), FCB_E_ARG, "bound: bad arguments");
    return launch_bound_modulus(z, n_complex, bound_out, static_cast<cudaStream_t>(stream));
}

extern "C" int fcb_fused_supported(int Ci, int Co, int band_limit, int R) { return fused_fwd_ok(Ci, Co, band_limit, R) ? 1 : 0; }

extern "C" int fcb_fwd_fused_workspace_bytes(int Ci, int Co, int band_limit, int R, size_t* bytes) {
    FCB_REQUIRE(bytes, FCB_E_ARG, "fwd_fused_workspace: null");
    FCB_REQUIRE(fused_fwd_ok(Ci, Co, band_limit, R), FCB_E_UNSUPPORTED, "fwd_fused_workspace: shape not supported (fcb_fused_supported)");
    *bytes = fused_fwd_ws_bytes(Ci, Co, band_limit, R);
    return FCB_OK;
}

extern "C" int fcb_fwd_fused_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt,
                                 const float* rot_tgt, const float* norm_tgt, float* y, int64_t N, int64_t n_feat_rows, int Ci,
                                 int Co, int band_limit, int R, void* ws, size_t ws_bytes, void* stream) {
    Dims d;
    int rc = check_dims("fwd_fused", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    FCB_REQUIRE(n_feat_rows >= N, FCB_E_ARG, "fwd_fused: n_feat_rows must be >= N");
    return launch_fused_fwd(x, W, rowptr_tgt, rec_tgt, rot_tgt, norm_tgt, y, N, n_feat_rows, Ci, Co, band_limit, R, ws, ws_bytes,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int fcb_fwd_dense_f32(const float* x, const float* W, const float* sten, const int32_t* rowptr_tgt,
                                 const int32_t* nbr_tgt, const int32_t* perm_tgt, float* y, float* contrib,
                                 float* contrib_absmax, int64_t N, int Ci, int Co, int band_limit, int R, int flags, void* ws,
                                 size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Dims d;
    int rc = check_dims("fwd_dense", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    FCB_REQUIRE(x && W && rowptr_tgt && nbr_tgt && perm_tgt && y && contrib && ws, FCB_E_ARG, "fwd_dense: null pointer");
    FCB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(W) && aligned16(contrib), FCB_E_ALIGN, "fwd_dense: pointers must be 16-byte aligned");
    FCB_REQUIRE(ws_bytes >= fwd_ws(d), FCB_E_WORKSPACE, "fwd_dense: workspace too small");
    if (N == 0) return FCB_OK;
    float* amax = fwd_amax_slot(ws, contrib_absmax, st);
    FCB_REQUIRE(amax, FCB_E_CUDA, "fwd_dense: cudaMemsetAsync failed");
    rc = launch_aggregate_dense(x, sten, rowptr_tgt, nbr_tgt, perm_tgt, contrib, N, Ci, band_limit, R, 0, amax, st);
    if (rc) return rc;
    return contract_fwd(d, contrib, amax, W, y, ws, ws_bytes, flags, st);
}

extern "C" int fcb_bwd_dense_f32(const float* x, const float* W, const float* gy, const float* contrib,
                                 const float* contrib_absmax, const float* sten, const int32_t* rowptr_src, const int32_t* nbr_src,
                                 const int32_t* perm_src, float* gx, float* gW, int64_t N, int Ci, int Co,
                                 int band_limit, int R, int flags, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Dims d;
    int rc = check_dims("bwd_dense", N, Ci, Co, band_limit, R, &d);
    if (rc) return rc;
    FCB_REQUIRE(x && W && gy && ws, FCB_E_ARG, "bwd_dense: null pointer");
    FCB_REQUIRE(!gW || contrib, FCB_E_ARG, "bwd_dense: grad W needs the contrib saved by the forward");
    FCB_REQUIRE(!gx || (sten && rowptr_src && nbr_src && perm_src), FCB_E_ARG, "bwd_dense: grad x needs the by-source plan");
    FCB_REQUIRE(ws_bytes >= bwd_ws(d, false, flags), FCB_E_WORKSPACE, "bwd_dense: workspace too small");
    if (N == 0) {
        if (gW && cudaMemsetAsync(gW, 0, (size_t)d.K * d.Co * 8, st) != cudaSuccess) {
            set_error("bwd_dense: cudaMemsetAsync failed");
            return FCB_E_CUDA;
        }
        return FCB_OK;
    }
    Arena ar(ws, ws_bytes);
    float* slots = ar.take<float>(128);
    auto gather = [&](float* G, float* g_amax, bool) {
        return launch_aggregate_dense(gy, sten, rowptr_src, nbr_src, perm_src, G, N, Co, band_limit, R, 1, g_amax, st);
    };
    return backward_common(d, x, W, gy, contrib, contrib_absmax, slots + 64, gather, gx, gW, ar, flags, st);
}

extern "C" int fcb_modrelu_fwd_f32(const float* x, const float* bias, float* y, int64_t N, int C, void* stream) {
    FCB_REQUIRE(x && bias && y && N >= 0 && C > 0, FCB_E_ARG, "modrelu_fwd: bad arguments");
    const int64_t tot = N * C;
    if (tot == 0) return FCB_OK;
    FCB_LAUNCH("modrelu_fwd", static_cast<cudaStream_t>(stream), k_modrelu_fwd<<<(unsigned)((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(x), bias, reinterpret_cast<float2*>(y), tot, C));
    return FCB_OK;
}

extern "C" int fcb_modrelu_bwd_workspace_bytes(int64_t N, int C, size_t* bytes) {
    FCB_REQUIRE(bytes && N >= 0 && C > 0, FCB_E_ARG, "modrelu_bwd_workspace: bad arguments");
    *bytes = align_up((size_t)((N + MR_ROWS - 1) / MR_ROWS + 1) * C * 4, 256);
    return FCB_OK;
}

extern "C" int fcb_modrelu_bwd_f32(const float* x, const float* bias, const float* gy, float* gx, float* gb, float* gx_bound,
                                   int64_t N, int C, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (gx_bound && cudaMemsetAsync(gx_bound, 0, 4, st) != cudaSuccess) {
        set_error("modrelu_bwd: cudaMemsetAsync failed");
        return FCB_E_CUDA;
    }
    FCB_REQUIRE(x && bias && gy && gx && gb && ws && N >= 0 && C > 0, FCB_E_ARG, "modrelu_bwd: bad arguments");
    const int64_t slabs = (N + MR_ROWS - 1) / MR_ROWS;
    FCB_REQUIRE(ws_bytes >= (size_t)(slabs + 1) * C * 4, FCB_E_WORKSPACE, "modrelu_bwd: workspace too small");
    float* parts = static_cast<float*>(ws);
    if (slabs > 0) {
        const int rl_count = C <= 256 ? (256 / C > 0 ? 256 / C : 1) : 1;
        const size_t smem = C <= 256 ? (size_t)rl_count * C * 4 : 0;
        FCB_LAUNCH("modrelu_bwd", st, k_modrelu_bwd<<<(unsigned)slabs, 256, smem, st>>>(reinterpret_cast<const float2*>(x), bias,
                                                          reinterpret_cast<const float2*>(gy),
                                                          reinterpret_cast<float2*>(gx), parts, N, C,
                                                          reinterpret_cast<uint32_t*>(gx_bound)));
    }
    FCB_LAUNCH("colsum_parts", st, k_colsum_parts<<<(unsigned)((C + 31) / 32), 1024, 0, st>>>(parts, gb, slabs, C));
    return FCB_OK;
}
