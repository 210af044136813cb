// F2: ECHO descriptors (nn/echo.py:94-148 of the reference).  For every support edge e: j -> i and every
// channel c whose source feature x[j,c] is not an origin entry, the log-map point ln_e, rotated into the feature's own
// frame (ln_e * conj(x/|x|)), votes x[j,c] * wxp_e into the four raster bins around n_bins * that point of the (i, c)
// disk histogram with bilinear weights (nn/echo.py:30-61); the descriptor is softAbs of the histogram.
//
// The reference compacts the (edge, channel) pairs with nonzero() and issues four scatter_adds with float atomics; here
// one thread owns one (target, channel) histogram (kept in shared memory, [bin][thread] so a warp's accesses to one bin
// are conflict-free), walks the row of the by-target CSR in a fixed order and writes each bin once: deterministic, no
// atomics, no host sync.  The backward walks the by-source CSR: one thread owns grad x[j, c].
#include "common.cuh"

namespace fcb {

constexpr int ECHO_MAX_BINS = 40;       // n_bins <= 3: 37 bins
constexpr int ECHO_THREADS = 128;
constexpr int ECHO_MAX_CELLS = 49;      // (2 * 3 + 1)^2 raster cells

struct EchoVote {
    int bin[4];
    float w[4];
    float px, py;      // n_bins * ln * conj(u)
    float fx, fy, cx, cy;
};

// nn/echo.py:38-58: p = n_bins * aligned; pC / pF = clamp(ceil / floor); weights and raster cells of the four votes
__device__ __forceinline__ EchoVote echo_rasterize(float2 ln, float2 u, int nb, const int* __restrict__ dmap) {
    EchoVote v;
    const float2 al = cmul_conj(ln, u);
    v.px = al.x * (float)nb;
    v.py = al.y * (float)nb;
    const float lim = (float)nb;
    v.cx = fminf(fmaxf(ceilf(v.px), -lim), lim);
    v.cy = fminf(fmaxf(ceilf(v.py), -lim), lim);
    v.fx = fminf(fmaxf(floorf(v.px), -lim), lim);
    v.fy = fminf(fmaxf(floorf(v.py), -lim), lim);
    const int w = 2 * nb + 1;
    const int icx = (int)v.cx + nb, icy = (int)v.cy + nb, ifx = (int)v.fx + nb, ify = (int)v.fy + nb;
    v.w[0] = (v.cx - v.px) * (v.cy - v.py);  v.bin[0] = dmap[w * ifx + ify];
    v.w[1] = (v.px - v.fx) * (v.py - v.fy);  v.bin[1] = dmap[w * icx + icy];
    v.w[2] = (v.px - v.fx) * (v.cy - v.py);  v.bin[2] = dmap[w * icx + ify];
    v.w[3] = (v.cx - v.px) * (v.py - v.fy);  v.bin[3] = dmap[w * ifx + icy];
    return v;
}

__device__ __forceinline__ bool is_origin_c(float2 z) { return (fabsf(z.x) < 1e-7f) && (fabsf(z.y) < 1e-7f); }

__global__ void __launch_bounds__(ECHO_THREADS) k_echo_fwd(const float2* __restrict__ x, const float2* __restrict__ ln,
                                                           const float2* __restrict__ wxp, const int32_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ nbr, const int32_t* __restrict__ perm,
                                                           const int32_t* __restrict__ dmap_g, float2* __restrict__ hist,
                                                           float* __restrict__ out, int64_t N, int C, int nb, int ds) {
    __shared__ float2 s_h[ECHO_MAX_BINS * ECHO_THREADS];
    __shared__ int s_map[ECHO_MAX_CELLS];
    const int cells = (2 * nb + 1) * (2 * nb + 1);
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s_map[i] = dmap_g[i];
    for (int b = 0; b < ds; ++b) s_h[b * ECHO_THREADS + threadIdx.x] = make_float2(0.f, 0.f);
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t row = t / C;
    const int c = (int)(t - row * C);
    float2* h = s_h + threadIdx.x;
    const int p1 = rowptr[row + 1];
    for (int p = rowptr[row]; p < p1; ++p) {
        const float2 z = x[(int64_t)nbr[p] * C + c];
        if (is_origin_c(z)) continue;                          // nn/echo.py:118: only non-zero features vote
        const int64_t e = perm[p];
        const float ri = rsqrtf(z.x * z.x + z.y * z.y);
        const EchoVote v = echo_rasterize(ln[e], make_float2(z.x * ri, z.y * ri), nb, s_map);
        const float2 xw = cmul(z, wxp[e]);                     // :131
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 a = h[v.bin[k] * ECHO_THREADS];
            a.x = fmaf(xw.x, v.w[k], a.x);
            a.y = fmaf(xw.y, v.w[k], a.y);
            h[v.bin[k] * ECHO_THREADS] = a;
        }
    }
    for (int b = 0; b < ds; ++b) {
        const float2 a = h[b * ECHO_THREADS];
        hist[t * ds + b] = a;
        out[t * ds + b] = is_origin_c(a) ? 0.f : sqrtf(a.x * a.x + a.y * a.y);       // softAbs, utils/field.py:29-37
    }
}

// grad x[j, c] = sum over the out-edges e: j -> i of
//   conj(wxp_e) * sum_k w_k gH[i,c,bin_k]                      (the vote x * wxp is linear in x)
// + g_phi * i x / |x|^2,  g_phi = gpx * py - gpy * px          (the raster point rotates with the feature's angle)
// with gH = g_out * H / |H| (softAbs backward, 0 at origin bins), g_wk = Re(conj(gH_k) xW) and the bilinear derivatives.
__global__ void __launch_bounds__(ECHO_THREADS) k_echo_bwd(const float2* __restrict__ x, const float2* __restrict__ ln,
                                                           const float2* __restrict__ wxp, const int32_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ nbr, const int32_t* __restrict__ perm,
                                                           const int32_t* __restrict__ dmap_g, const float2* __restrict__ hist,
                                                           const float* __restrict__ g_out, float2* __restrict__ gx, int64_t N,
                                                           int C, int nb, int ds) {
    __shared__ int s_map[ECHO_MAX_CELLS];
    const int cells = (2 * nb + 1) * (2 * nb + 1);
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s_map[i] = dmap_g[i];
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t row = t / C;
    const int c = (int)(t - row * C);
    const float2 z = x[t];
    float2 acc = make_float2(0.f, 0.f);
    if (!is_origin_c(z)) {
        const float n2 = z.x * z.x + z.y * z.y;
        const float ri = rsqrtf(n2);
        const float2 u = make_float2(z.x * ri, z.y * ri);
        float gphi = 0.f;
        const int p1 = rowptr[row + 1];
        for (int p = rowptr[row]; p < p1; ++p) {
            const int64_t e = perm[p];
            const int64_t i = nbr[p];
            const EchoVote v = echo_rasterize(ln[e], u, nb, s_map);
            const float2 wx = wxp[e];
            const float2 xw = cmul(z, wx);
            const float2* H = hist + (i * C + c) * ds;
            const float* go = g_out + (i * C + c) * ds;
            float2 gxw = make_float2(0.f, 0.f);
            float gw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 hk = H[v.bin[k]];
                float2 gh = make_float2(0.f, 0.f);
                if (!is_origin_c(hk)) {
                    const float s = go[v.bin[k]] * rsqrtf(hk.x * hk.x + hk.y * hk.y);
                    gh = make_float2(s * hk.x, s * hk.y);
                }
                gxw.x = fmaf(v.w[k], gh.x, gxw.x);
                gxw.y = fmaf(v.w[k], gh.y, gxw.y);
                gw[k] = gh.x * xw.x + gh.y * xw.y;
            }
            const float2 d = cmul_conj(gxw, wx);               // gxw * conj(wxp)
            acc.x += d.x;
            acc.y += d.y;
            const float gpx = -gw[0] * (v.cy - v.py) + gw[1] * (v.py - v.fy) + gw[2] * (v.cy - v.py) - gw[3] * (v.py - v.fy);
            const float gpy = -gw[0] * (v.cx - v.px) + gw[1] * (v.px - v.fx) - gw[2] * (v.px - v.fx) + gw[3] * (v.cx - v.px);
            gphi += gpx * v.py - gpy * v.px;                   // dp/dphi = -i p
        }
        const float s = gphi / n2;
        acc.x -= s * z.y;                                      // dphi/dx = (-y, x) / |z|^2
        acc.y += s * z.x;
    }
    gx[t] = acc;
}

}  // namespace fcb

using namespace fcb;

static int echo_check(const char* who, int64_t N, int C, int nb, int ds) {
    FCB_REQUIRE(N >= 0 && C > 0, FCB_E_ARG, "%s: bad sizes", who);
    FCB_REQUIRE(nb >= 1 && nb <= 3 && ds >= 1 && ds <= ECHO_MAX_BINS, FCB_E_UNSUPPORTED, "%s: n_bins must be 1..3 (at most %d histogram bins)", who,
                ECHO_MAX_BINS);
    return FCB_OK;
}

extern "C" int fcb_echo_fwd_f32(const float* x, const float* ln, const float* wxp, const int32_t* rowptr_tgt, const int32_t* nbr_tgt,
                                const int32_t* perm_tgt, const int32_t* dmap, float* hist, float* out, int64_t N, int C, int n_bins,
                                int hdim, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = echo_check("echo_fwd", N, C, n_bins, hdim);
    if (rc) return rc;
    FCB_REQUIRE(x && ln && wxp && rowptr_tgt && nbr_tgt && perm_tgt && dmap && hist && out, FCB_E_ARG, "echo_fwd: null pointer");
    const int64_t tot = N * C;
    if (tot == 0) return FCB_OK;
    FCB_LAUNCH("echo_fwd", st, k_echo_fwd<<<(unsigned)((tot + ECHO_THREADS - 1) / ECHO_THREADS), ECHO_THREADS, 0, st>>>(
                                   reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(ln), reinterpret_cast<const float2*>(wxp),
                                   rowptr_tgt, nbr_tgt, perm_tgt, dmap, reinterpret_cast<float2*>(hist), out, N, C, n_bins, hdim));
    return FCB_OK;
}

extern "C" int fcb_echo_bwd_f32(const float* x, const float* ln, const float* wxp, const int32_t* rowptr_src, const int32_t* nbr_src,
                                const int32_t* perm_src, const int32_t* dmap, const float* hist, const float* g_out, float* gx,
                                int64_t N, int C, int n_bins, int hdim, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = echo_check("echo_bwd", N, C, n_bins, hdim);
    if (rc) return rc;
    FCB_REQUIRE(x && ln && wxp && rowptr_src && nbr_src && perm_src && dmap && hist && g_out && gx, FCB_E_ARG, "echo_bwd: null pointer");
    const int64_t tot = N * C;
    if (tot == 0) return FCB_OK;
    FCB_LAUNCH("echo_bwd", st, k_echo_bwd<<<(unsigned)((tot + ECHO_THREADS - 1) / ECHO_THREADS), ECHO_THREADS, 0, st>>>(
                                   reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(ln), reinterpret_cast<const float2*>(wxp),
                                   rowptr_src, nbr_src, perm_src, dmap, reinterpret_cast<const float2*>(hist), g_out,
                                   reinterpret_cast<float2*>(gx), N, C, n_bins, hdim));
    return FCB_OK;
}
