// K0 — edge plan: FCPrecomp arithmetic + the two CSR orders, entirely on the device.
// Follows transforms/fc_precomp.py:10-27,67-74,87,92 of the reference (see include/fieldconv_b200.h).
#include "common.cuh"

namespace fcb {

// ----------------------------------------------------------------------------- stable LSD radix sort
// One warp owns a contiguous chunk of RS_CHUNK elements; per pass: per-chunk digit histogram ->
// exclusive scan over (digit-major, chunk-minor) counts -> stable scatter with warp match ranking.
constexpr int RS_CHUNK = 2048;
constexpr int RS_WARPS = 8;

__global__ void __launch_bounds__(RS_WARPS * 32) k_radix_hist(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                              uint32_t* __restrict__ hist, int64_t nchunks) {
    __shared__ uint32_t cnt[RS_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t chunk = (int64_t)blockIdx.x * RS_WARPS + warp;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    if (chunk < nchunks) {
        const int64_t lo = chunk * RS_CHUNK;
        const int64_t hi = min(n, lo + RS_CHUNK);
        for (int64_t i = lo + lane; i < hi; i += 32) atomicAdd(&cnt[warp][(keys[i] >> shift) & 255u], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[(int64_t)d * nchunks + chunk] = cnt[warp][d];
    }
}

__global__ void __launch_bounds__(RS_WARPS * 32) k_radix_scatter(const uint32_t* __restrict__ keys_in,
                                                                 const uint32_t* __restrict__ vals_in,
                                                                 uint32_t* __restrict__ keys_out,
                                                                 uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                                 const uint32_t* __restrict__ offs, int64_t nchunks) {
    __shared__ uint32_t cnt[RS_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t chunk = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (chunk >= nchunks) return;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = offs[(int64_t)d * nchunks + chunk];
    __syncwarp();
    const int64_t lo = chunk * RS_CHUNK;
    const int64_t hi = min(n, lo + RS_CHUNK);
    for (int64_t base = lo; base < hi; base += 32) {
        const int64_t i = base + lane;
        const bool valid = i < hi;
        uint32_t key = 0, val = 0;
        if (valid) { key = keys_in[i]; val = vals_in[i]; }
        const uint32_t d = valid ? ((key >> shift) & 255u) : (256u + lane);  // invalid lanes: unique groups
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (valid) pos = cnt[warp][d] + rank;
        __syncwarp();
        if (valid && rank == 0) cnt[warp][d] += __popc(peers);
        __syncwarp();
        if (valid) { keys_out[pos] = key; vals_out[pos] = val; }
    }
}

// ----------------------------------------------------------------------------- exclusive scan (uint32)
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS) k_scan_tile(uint32_t* __restrict__ data, int64_t n,
                                                          uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t warp_tot[SC_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0u;
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; ++w) {
        if (w < warp) woff += warp_tot[w];
        total += warp_tot[w];
    }
    uint32_t run = woff + inc - sum;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0 && tile_sums) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_add(uint32_t* __restrict__ data, int64_t n,
                                                         const uint32_t* __restrict__ tile_offs) {
    const uint32_t off = tile_offs[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i)
        if (base + i < n) data[base + i] += off;
}

static size_t scan_scratch_elems(int64_t n) {
    size_t tot = 0;
    while (n > SC_TILE) {
        n = (n + SC_TILE - 1) / SC_TILE;
        tot += align_up((size_t)n, 64);
    }
    return tot + 64;
}

static int exclusive_scan(uint32_t* data, int64_t n, uint32_t* scratch, cudaStream_t st) {
    if (n <= 0) return FCB_OK;
    const int64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    if (tiles == 1) {
        FCB_LAUNCH("scan_tile", st, k_scan_tile<<<1, SC_THREADS, 0, st>>>(data, n, nullptr));
        return FCB_OK;
    }
    FCB_LAUNCH("scan_tile", st, k_scan_tile<<<(unsigned)tiles, SC_THREADS, 0, st>>>(data, n, scratch));
    int rc = exclusive_scan(scratch, tiles, scratch + align_up((size_t)tiles, 64), st);
    if (rc) return rc;
    FCB_LAUNCH("scan_add", st, k_scan_add<<<(unsigned)tiles, SC_THREADS, 0, st>>>(data, n, scratch));
    return FCB_OK;
}

size_t sort_workspace(int64_t n) {
    const int64_t nchunks = (n + RS_CHUNK - 1) / RS_CHUNK;
    const size_t hist = (size_t)256 * (size_t)(nchunks > 0 ? nchunks : 1);
    return align_up(hist * 4, 256) + align_up(scan_scratch_elems((int64_t)hist) * 4, 256) + 512;
}

int sort_pairs(uint32_t* k_in, uint32_t* v_in, uint32_t* k_out, uint32_t* v_out, int64_t n, int bits, void* ws,
               size_t ws_bytes, cudaStream_t st) {
    FCB_REQUIRE(n >= 0 && bits >= 0 && bits <= 32, FCB_E_ARG, "sort: bad n/bits");
    FCB_REQUIRE(ws_bytes >= sort_workspace(n), FCB_E_WORKSPACE, "sort: workspace too small");
    if (n == 0) return FCB_OK;
    const int64_t nchunks = (n + RS_CHUNK - 1) / RS_CHUNK;
    Arena ar(ws, ws_bytes);
    uint32_t* hist = ar.take<uint32_t>((size_t)256 * nchunks);
    uint32_t* scratch = ar.take<uint32_t>(scan_scratch_elems(256 * nchunks));
    int passes = (bits + 7) / 8;
    if (passes == 0) passes = 1;
    const unsigned blocks = (unsigned)((nchunks + RS_WARPS - 1) / RS_WARPS);
    uint32_t *ki = k_in, *vi = v_in, *ko = k_out, *vo = v_out;
    // make the result land in (k_out, v_out): with an even pass count start by copying in -> out
    if (passes % 2 == 0) {
        cudaMemcpyAsync(k_out, k_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(v_out, v_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
        ki = k_out; vi = v_out; ko = k_in; vo = v_in;
    }
    for (int p = 0; p < passes; ++p) {
        FCB_LAUNCH("radix_hist", st, k_radix_hist<<<blocks, RS_WARPS * 32, 0, st>>>(ki, n, 8 * p, hist, nchunks));
        int rc = exclusive_scan(hist, 256 * nchunks, scratch, st);
        if (rc) return rc;
        FCB_LAUNCH("radix_scatter", st, k_radix_scatter<<<blocks, RS_WARPS * 32, 0, st>>>(ki, vi, ko, vo, n, 8 * p, hist, nchunks));
        uint32_t* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    return FCB_OK;
}

// ----------------------------------------------------------------------------- FCPrecomp arithmetic
struct RingTap { int f; float t; bool keep; };

// transforms/fc_precomp.py:67-74 (support filter) and :10-27 (ring floor + two-tap weight).
// IEEE round-to-nearest division/subtraction so (f, t) are bit-identical to the CPU reference.
__device__ __forceinline__ RingTap ring_tap(float log_mag, float eps, const float* __restrict__ radii, int R) {
    RingTap o;
    const float rn = __fdiv_rn(log_mag, eps);
    o.keep = (rn <= 1.0f);
    int c = 0;
    bool found = false;
    for (int k = 0; k < R; ++k) {
        if (!found && radii[k] >= rn) { c = k; found = true; }
    }
    if (c == 0) c = 1;  // also the "not found" fallback of argmin over the all-1e8 row
    o.f = c - 1;
    o.t = __fdiv_rn(__fsub_rn(rn, radii[o.f]), __fsub_rn(radii[c], radii[o.f]));
    return o;
}

__global__ void k_edge_keys(const int64_t* __restrict__ edges, const float* __restrict__ log_mag,
                            const float* __restrict__ radii, float eps, int64_t E, int64_t N, int R,
                            uint32_t* __restrict__ key_t, uint32_t* __restrict__ key_s, uint32_t* __restrict__ idx_t,
                            uint32_t* __restrict__ idx_s) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t j = edges[2 * e], i = edges[2 * e + 1];
    const uint32_t seg = (uint32_t)(R - 1);
    const uint32_t invalid = (uint32_t)N * seg;
    uint32_t kt = invalid, ks = invalid;
    if (j >= 0 && j < N && i >= 0 && i < N) {
        RingTap tap = ring_tap(log_mag[e], eps, radii, R);
        if (tap.keep) {
            kt = (uint32_t)i * seg + (uint32_t)tap.f;
            ks = (uint32_t)j * seg + (uint32_t)tap.f;
        }
    }
    key_t[e] = kt; key_s[e] = ks;
    idx_t[e] = (uint32_t)e; idx_s[e] = (uint32_t)e;
}

__global__ void k_dense_keys(const int64_t* __restrict__ edges, int64_t E, int64_t N, uint32_t* __restrict__ key_t,
                             uint32_t* __restrict__ key_s, uint32_t* __restrict__ idx_t, uint32_t* __restrict__ idx_s) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t j = edges[2 * e], i = edges[2 * e + 1];
    const bool ok = (j >= 0 && j < N && i >= 0 && i < N);
    key_t[e] = ok ? (uint32_t)i : (uint32_t)N;
    key_s[e] = ok ? (uint32_t)j : (uint32_t)N;
    idx_t[e] = (uint32_t)e; idx_s[e] = (uint32_t)e;
}

// rowptr[v] = first sorted position whose key >= v*seg  (v = 0..N); rowptr[N] = number of kept edges
__global__ void k_rowptr(const uint32_t* __restrict__ sorted_keys, int64_t E, int64_t N, uint32_t seg,
                         int32_t* __restrict__ rowptr) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > N) return;
    const uint32_t want = (uint32_t)v * seg;
    int64_t lo = 0, hi = E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    rowptr[v] = (int32_t)lo;
}

// per target row: sum of source masses in CSR order (fc_precomp.py:87 denominator), then the
// per-edge normalised weight, stored by ORIGINAL edge id so both orders use the same value.
__global__ void k_row_weights(const int32_t* __restrict__ rowptr_t, const uint32_t* __restrict__ perm_t,
                              const int64_t* __restrict__ edges, const float* __restrict__ w, int64_t N,
                              float* __restrict__ wn_edge) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int p0 = rowptr_t[i], p1 = rowptr_t[i + 1];
    float sum = 0.f;
    for (int p = p0; p < p1; ++p) sum = __fadd_rn(sum, w[edges[2 * (int64_t)perm_t[p]]]);
    const float den = __fadd_rn(1e-12f, sum);
    for (int p = p0; p < p1; ++p) {
        const uint32_t e = perm_t[p];
        wn_edge[e] = __fdiv_rn(w[edges[2 * (int64_t)e]], den);
    }
}

__global__ void k_emit_records(const uint32_t* __restrict__ sorted_keys, const uint32_t* __restrict__ perm,
                               const int32_t* __restrict__ rowptr, const int64_t* __restrict__ edges,
                               const float* __restrict__ log_mag, const float* __restrict__ log_ang,
                               const float2* __restrict__ xp, const float* __restrict__ wn_edge,
                               const float* __restrict__ radii, float eps, int64_t E, int64_t N, int R, int nbr_col,
                               int4* __restrict__ rec, float2* __restrict__ rot, int32_t* __restrict__ perm_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E || p >= rowptr[N]) return;
    const uint32_t e = perm[p];
    const int64_t nbr = edges[2 * (int64_t)e + nbr_col];
    RingTap tap = ring_tap(log_mag[e], eps, radii, R);
    const float wn = wn_edge[e];
    const float2 t = xp[e];
    int4 r;
    r.x = (int)((uint32_t)nbr | ((uint32_t)tap.f << NBR_BITS));
    r.y = __float_as_int(tap.t);
    r.z = __float_as_int(wn * t.x);   // fc_precomp.py:92
    r.w = __float_as_int(wn * t.y);
    rec[p] = r;
    float s, c;
    sincosf(log_ang[e], &s, &c);      // fc_precomp.py:83-84 (m = 1; higher m by recurrence in K1)
    rot[p] = make_float2(c, s);
    perm_out[p] = (int32_t)e;
}

__global__ void k_emit_dense(const uint32_t* __restrict__ perm, const int32_t* __restrict__ rowptr,
                             const int64_t* __restrict__ edges, int64_t E, int64_t N, int nbr_col,
                             int32_t* __restrict__ nbr_out, int32_t* __restrict__ perm_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E || p >= rowptr[N]) return;
    const uint32_t e = perm[p];
    nbr_out[p] = (int32_t)edges[2 * (int64_t)e + nbr_col];
    perm_out[p] = (int32_t)e;
}

static int key_bits(uint64_t max_key) {
    int b = 0;
    while ((max_key >> b) != 0 && b < 32) ++b;
    return b < 1 ? 1 : b;
}

}  // namespace fcb

using namespace fcb;

extern "C" int fcb_sort_workspace_bytes(int64_t n, size_t* bytes) {
    FCB_REQUIRE(n >= 0 && bytes, FCB_E_ARG, "sort_workspace: bad arguments");
    *bytes = sort_workspace(n);
    return FCB_OK;
}

extern "C" int fcb_sort_pairs_u32(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                                  int64_t n, int bits, void* workspace, size_t workspace_bytes, void* stream) {
    return sort_pairs(keys_in, vals_in, keys_out, vals_out, n, bits, workspace, workspace_bytes,
                      static_cast<cudaStream_t>(stream));
}

static size_t plan_ws(int64_t E) {
    const size_t e = (size_t)(E > 0 ? E : 1);
    // key_t,key_s,idx_t,idx_s + sorted key/idx + wn_edge + sort scratch
    return 7 * align_up(e * 4, 256) + sort_workspace(E) + 1024;
}

extern "C" int fcb_plan_workspace_bytes(int64_t E, int64_t N, int R, size_t* bytes) {
    FCB_REQUIRE(E >= 0 && N >= 0 && bytes, FCB_E_ARG, "plan_workspace: bad arguments");
    (void)R;
    *bytes = plan_ws(E);
    return FCB_OK;
}

extern "C" int fcb_plan_dense_workspace_bytes(int64_t E, int64_t N, size_t* bytes) {
    FCB_REQUIRE(E >= 0 && N >= 0 && bytes, FCB_E_ARG, "plan_workspace: bad arguments");
    *bytes = plan_ws(E);
    return FCB_OK;
}

extern "C" int fcb_plan_build(const int64_t* edges, const float* log_mag, const float* log_ang, const float* xp,
                              const float* w, const float* radii, float epsilon, int64_t E, int64_t N, int R,
                              int32_t* rowptr_tgt, void* rec_tgt, float* rot_tgt, int32_t* perm_tgt,
                              int32_t* rowptr_src, void* rec_src, float* rot_src, int32_t* perm_src, void* ws,
                              size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FCB_REQUIRE(E >= 0 && N >= 0, FCB_E_ARG, "plan_build: negative size");
    FCB_REQUIRE(R >= 2 && R <= FCB_MAX_RINGS, FCB_E_UNSUPPORTED,
                "plan_build: n_rings=%d unsupported (reference divides by n_rings-1; max %d)", R, FCB_MAX_RINGS);
    FCB_REQUIRE(N <= FCB_MAX_VERTICES, FCB_E_UNSUPPORTED, "plan_build: N=%lld exceeds %d", (long long)N, FCB_MAX_VERTICES);
    FCB_REQUIRE(E < (int64_t)0x7fffffff, FCB_E_UNSUPPORTED, "plan_build: E=%lld does not fit int32 row pointers", (long long)E);
    FCB_REQUIRE((uint64_t)N * (uint64_t)(R - 1) < 0xffffffffull, FCB_E_UNSUPPORTED, "plan_build: N*(R-1) overflows the sort key");
    FCB_REQUIRE(rowptr_tgt && rowptr_src && rec_tgt && rec_src && rot_tgt && rot_src && perm_tgt && perm_src,
                FCB_E_ARG, "plan_build: null output");
    FCB_REQUIRE(E == 0 || (edges && log_mag && log_ang && xp), FCB_E_ARG, "plan_build: null edge input");
    FCB_REQUIRE(w && radii, FCB_E_ARG, "plan_build: null input");
    FCB_REQUIRE(aligned16(rec_tgt) && aligned16(rec_src), FCB_E_ALIGN, "plan_build: rec buffers must be 16-byte aligned");
    FCB_REQUIRE(ws_bytes >= plan_ws(E), FCB_E_WORKSPACE, "plan_build: workspace too small");
    FCB_REQUIRE(epsilon > 0.f, FCB_E_ARG, "plan_build: epsilon must be positive");

    Arena ar(ws, ws_bytes);
    const size_t e = (size_t)(E > 0 ? E : 1);
    uint32_t* key_t = ar.take<uint32_t>(e);
    uint32_t* key_s = ar.take<uint32_t>(e);
    uint32_t* idx_t = ar.take<uint32_t>(e);
    uint32_t* idx_s = ar.take<uint32_t>(e);
    uint32_t* skey = ar.take<uint32_t>(e);
    uint32_t* sidx = ar.take<uint32_t>(e);
    float* wn_edge = ar.take<float>(e);
    void* sort_ws = ar.take<char>(sort_workspace(E));
    const size_t sort_ws_bytes = sort_workspace(E);
    const int bits = key_bits((uint64_t)N * (uint64_t)(R - 1));
    const unsigned eb = (unsigned)((E + 255) / 256), nb = (unsigned)((N + 1 + 255) / 256);
    const uint32_t seg = (uint32_t)(R - 1);

    if (E > 0) {
        FCB_LAUNCH("edge_keys", st, k_edge_keys<<<eb, 256, 0, st>>>(edges, log_mag, radii, epsilon, E, N, R, key_t, key_s, idx_t, idx_s));
    }
    // by (target, ring floor)
    int rc = sort_pairs(key_t, idx_t, skey, sidx, E, bits, sort_ws, sort_ws_bytes, st);
    if (rc) return rc;
    FCB_LAUNCH("rowptr", st, k_rowptr<<<nb, 256, 0, st>>>(skey, E, N, seg, rowptr_tgt));
    if (E > 0 && N > 0) {
        FCB_LAUNCH("row_weights", st, k_row_weights<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(rowptr_tgt, sidx, edges, w, N, wn_edge));
        FCB_LAUNCH("emit_records(tgt)", st, k_emit_records<<<eb, 256, 0, st>>>(skey, sidx, rowptr_tgt, edges, log_mag, log_ang,
                                          reinterpret_cast<const float2*>(xp), wn_edge, radii, epsilon, E, N, R, 0,
                                          static_cast<int4*>(rec_tgt), reinterpret_cast<float2*>(rot_tgt), perm_tgt));
    }
    // by (source, ring floor)
    rc = sort_pairs(key_s, idx_s, skey, sidx, E, bits, sort_ws, sort_ws_bytes, st);
    if (rc) return rc;
    FCB_LAUNCH("rowptr", st, k_rowptr<<<nb, 256, 0, st>>>(skey, E, N, seg, rowptr_src));
    if (E > 0 && N > 0) {
        FCB_LAUNCH("emit_records(src)", st, k_emit_records<<<eb, 256, 0, st>>>(skey, sidx, rowptr_src, edges, log_mag, log_ang,
                                          reinterpret_cast<const float2*>(xp), wn_edge, radii, epsilon, E, N, R, 1,
                                          static_cast<int4*>(rec_src), reinterpret_cast<float2*>(rot_src), perm_src));
    }
    return FCB_OK;
}

extern "C" int fcb_plan_build_dense(const int64_t* edges, int64_t E, int64_t N, int32_t* rowptr_tgt, int32_t* nbr_tgt,
                                    int32_t* perm_tgt, int32_t* rowptr_src, int32_t* nbr_src, int32_t* perm_src,
                                    void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FCB_REQUIRE(E >= 0 && N >= 0, FCB_E_ARG, "plan_build_dense: negative size");
    FCB_REQUIRE(N <= FCB_MAX_VERTICES, FCB_E_UNSUPPORTED, "plan_build_dense: N too large");
    FCB_REQUIRE(E < (int64_t)0x7fffffff, FCB_E_UNSUPPORTED, "plan_build_dense: E does not fit int32 row pointers");
    FCB_REQUIRE(rowptr_tgt && rowptr_src && nbr_tgt && nbr_src && perm_tgt && perm_src, FCB_E_ARG, "plan_build_dense: null output");
    FCB_REQUIRE(E == 0 || edges, FCB_E_ARG, "plan_build_dense: null edges");
    FCB_REQUIRE(ws_bytes >= plan_ws(E), FCB_E_WORKSPACE, "plan_build_dense: workspace too small");
    Arena ar(ws, ws_bytes);
    const size_t e = (size_t)(E > 0 ? E : 1);
    uint32_t* key_t = ar.take<uint32_t>(e);
    uint32_t* key_s = ar.take<uint32_t>(e);
    uint32_t* idx_t = ar.take<uint32_t>(e);
    uint32_t* idx_s = ar.take<uint32_t>(e);
    uint32_t* skey = ar.take<uint32_t>(e);
    uint32_t* sidx = ar.take<uint32_t>(e);
    (void)ar.take<float>(e);
    void* sort_ws = ar.take<char>(sort_workspace(E));
    const size_t sort_ws_bytes = sort_workspace(E);
    const int bits = key_bits((uint64_t)N);
    const unsigned eb = (unsigned)((E + 255) / 256), nb = (unsigned)((N + 1 + 255) / 256);
    if (E > 0) {
        FCB_LAUNCH("dense_keys", st, k_dense_keys<<<eb, 256, 0, st>>>(edges, E, N, key_t, key_s, idx_t, idx_s));
    }
    int rc = sort_pairs(key_t, idx_t, skey, sidx, E, bits, sort_ws, sort_ws_bytes, st);
    if (rc) return rc;
    FCB_LAUNCH("rowptr", st, k_rowptr<<<nb, 256, 0, st>>>(skey, E, N, 1u, rowptr_tgt));
    if (E > 0) {
        FCB_LAUNCH("emit_dense(tgt)", st, k_emit_dense<<<eb, 256, 0, st>>>(sidx, rowptr_tgt, edges, E, N, 0, nbr_tgt, perm_tgt));
    }
    rc = sort_pairs(key_s, idx_s, skey, sidx, E, bits, sort_ws, sort_ws_bytes, st);
    if (rc) return rc;
    FCB_LAUNCH("rowptr", st, k_rowptr<<<nb, 256, 0, st>>>(skey, E, N, 1u, rowptr_src));
    if (E > 0) {
        FCB_LAUNCH("emit_dense(src)", st, k_emit_dense<<<eb, 256, 0, st>>>(sidx, rowptr_src, edges, E, N, 1, nbr_src, perm_src));
    }
    return FCB_OK;
}

// ----------------------------------------------------------------------------- FCPrecomp outputs in the reference's dense form
// transforms/fc_precomp.py:53-97 returns (supp_edges', supp_sten (E',R,M), ln (E'), wxp (E')) with the kept edges in
// INPUT order.  The compact plan already holds (f, t, wxp) per kept edge; this expands it.
namespace fcb {

__global__ void k_keep_flags(const int64_t* __restrict__ edges, const float* __restrict__ log_mag, float eps, int64_t E,
                             int64_t N, uint32_t* __restrict__ flag) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t j = edges[2 * e], i = edges[2 * e + 1];
    const bool ok = j >= 0 && j < N && i >= 0 && i < N && __fdiv_rn(log_mag[e], eps) <= 1.0f;   // fc_precomp.py:67-69
    flag[e] = ok ? 1u : 0u;
}

// one thread per kept edge (by-target sorted position p); slot[e] = rank of input edge e among the kept edges
__global__ void k_expand_stencil(const int32_t* __restrict__ rowptr, const int4* __restrict__ rec,
                                 const int32_t* __restrict__ perm, const uint32_t* __restrict__ slot,
                                 const int64_t* __restrict__ edges, const float* __restrict__ log_mag,
                                 const float* __restrict__ log_ang, float eps, int64_t E, int64_t N, int R, int B,
                                 int64_t E_out, int64_t* __restrict__ edges_out, float2* __restrict__ sten,
                                 float2* __restrict__ ln, float2* __restrict__ wxp_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E || p >= rowptr[N]) return;
    const int64_t e = perm[p];
    const int64_t q = slot[e];
    if (q >= E_out) return;
    const int4 rc = rec[p];
    const int f = (int)((uint32_t)rc.x >> NBR_BITS);
    const float t = __int_as_float(rc.y);
    const float2 wxp = make_float2(__int_as_float(rc.z), __int_as_float(rc.w));
    const float theta = log_ang[e];
    const float rn = __fdiv_rn(log_mag[e], eps);
    edges_out[2 * q] = edges[2 * e];
    edges_out[2 * q + 1] = edges[2 * e + 1];
    float s1, c1;
    sincosf(theta, &s1, &c1);
    ln[q] = make_float2(rn * c1, rn * s1);                 // fc_precomp.py:77
    wxp_out[q] = wxp;                                      // fc_precomp.py:92
    const int M = 2 * B + 1;
    float2* dst = sten + q * (int64_t)R * M;
    for (int r = 0; r < R; ++r) {
        const float w = (r == f + 1) ? t : (r == f ? __fsub_rn(1.0f, t) : 0.f);   // fc_precomp.py:24-25
        for (int m = -B; m <= B; ++m) {
            float sm, cm;
            sincosf(__fmul_rn((float)m, theta), &sm, &cm);                         // fc_precomp.py:83-84
            const float ax = __fmul_rn(w, cm), ay = __fmul_rn(w, sm);
            dst[r * M + m + B] = make_float2(__fsub_rn(__fmul_rn(ax, wxp.x), __fmul_rn(ay, wxp.y)),
                                             __fadd_rn(__fmul_rn(ax, wxp.y), __fmul_rn(ay, wxp.x)));   // :95
        }
    }
}

}  // namespace fcb

extern "C" int fcb_precomp_workspace_bytes(int64_t E, size_t* bytes) {
    FCB_REQUIRE(E >= 0 && bytes, FCB_E_ARG, "precomp_workspace: bad arguments");
    *bytes = fcb::align_up((size_t)(E > 0 ? E : 1) * 4, 256) + fcb::align_up(fcb::scan_scratch_elems(E) * 4, 256) + 512;
    return FCB_OK;
}

extern "C" int fcb_precomp_expand_f32(const int64_t* edges, const float* log_mag, const float* log_ang, float epsilon,
                                      int64_t E, int64_t N, int R, int band_limit, const int32_t* rowptr_tgt,
                                      const void* rec_tgt, const int32_t* perm_tgt, int64_t E_kept, int64_t* edges_out,
                                      float* supp_sten, float* ln, float* wxp, void* ws, size_t ws_bytes, void* stream) {
    using namespace fcb;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FCB_REQUIRE(E >= 0 && N >= 0 && E_kept >= 0 && E_kept <= E, FCB_E_ARG, "precomp_expand: bad sizes");
    FCB_REQUIRE(R >= 2 && R <= FCB_MAX_RINGS && band_limit >= 0 && band_limit <= FCB_MAX_BAND_LIMIT, FCB_E_UNSUPPORTED,
                "precomp_expand: unsupported n_rings / band_limit");
    FCB_REQUIRE(epsilon > 0.f, FCB_E_ARG, "precomp_expand: epsilon must be positive");
    if (E == 0 || E_kept == 0) return FCB_OK;
    FCB_REQUIRE(edges && log_mag && log_ang && rowptr_tgt && rec_tgt && perm_tgt && edges_out && supp_sten && ln && wxp && ws,
                FCB_E_ARG, "precomp_expand: null pointer");
    size_t need = 0;
    fcb_precomp_workspace_bytes(E, &need);
    FCB_REQUIRE(ws_bytes >= need, FCB_E_WORKSPACE, "precomp_expand: workspace too small");
    Arena ar(ws, ws_bytes);
    uint32_t* slot = ar.take<uint32_t>((size_t)E);
    uint32_t* scratch = ar.take<uint32_t>(scan_scratch_elems(E));
    const unsigned eb = (unsigned)((E + 255) / 256);
    FCB_LAUNCH("keep_flags", st, k_keep_flags<<<eb, 256, 0, st>>>(edges, log_mag, epsilon, E, N, slot));
    int rc = exclusive_scan(slot, E, scratch, st);
    if (rc) return rc;
    FCB_LAUNCH("expand_stencil", st, k_expand_stencil<<<eb, 256, 0, st>>>(rowptr_tgt, static_cast<const int4*>(rec_tgt), perm_tgt, slot,
                                     edges, log_mag, log_ang, epsilon, E, N, R, band_limit, E_kept, edges_out,
                                     reinterpret_cast<float2*>(supp_sten), reinterpret_cast<float2*>(ln),
                                     reinterpret_cast<float2*>(wxp)));
    return FCB_OK;
}

