// Support-graph construction on the device: the Euclidean radius query of the reference's
// SupportGraph transform (transforms/support_graph.py:56-59: radius(pos, pos, epsilon, max_num_neighbors=512), self loops
// included, rows (query j, found i) grouped by j).  The reference delegates to torch_cluster (a k-d tree on the CPU, a
// brute-force scan on the GPU); FieldConv accepts any edge order, so the contract is the edge SET per query point, capped at
// max_neighbors.
//
// Uniform grid with cell size r, hashed into a power-of-two table: points are sorted by bucket (the library's stable LSD
// radix sort), a query scans the 27 cells around its own and accepts a candidate only if it really lies in the scanned cell
// (two of the 27 cells may share a bucket: without the test their points would be reported twice).  Two passes with the same
// traversal order: count (-> exclusive scan by the caller) and fill.
#include "common.cuh"

namespace fcb {

struct GridParams {
    float ox, oy, oz;       // lower corner of the bounding box
    float inv_cell;         // 1 / r
    uint32_t mask;          // table size - 1
};

__device__ __forceinline__ int3 cell_of(const float* __restrict__ pos, int64_t i, const GridParams g) {
    return make_int3((int)floorf((pos[3 * i] - g.ox) * g.inv_cell), (int)floorf((pos[3 * i + 1] - g.oy) * g.inv_cell),
                     (int)floorf((pos[3 * i + 2] - g.oz) * g.inv_cell));
}
__device__ __forceinline__ uint32_t bucket_of(int3 c, uint32_t mask) {
    return (((uint32_t)c.x * 73856093u) ^ ((uint32_t)c.y * 19349663u) ^ ((uint32_t)c.z * 83492791u)) & mask;
}

__global__ void k_radius_keys(const float* __restrict__ pos, int64_t N, GridParams g, uint32_t* __restrict__ keys,
                              uint32_t* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    keys[i] = bucket_of(cell_of(pos, i, g), g.mask);
    ids[i] = (uint32_t)i;
}

// start[b] = first sorted position whose bucket >= b (b = 0 .. table size)
__global__ void k_radius_starts(const uint32_t* __restrict__ sorted_keys, int64_t N, uint32_t table, int32_t* __restrict__ start) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > table) return;
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < (uint32_t)b) lo = mid + 1; else hi = mid;
    }
    start[b] = (int32_t)lo;
}

// FILL = false: counts[i] = min(max_nbrs, #{j : |p_j - p_i| <= r});  FILL = true: writes the pairs (i, j) at offsets[i]..
template <bool FILL>
__global__ void k_radius_scan(const float* __restrict__ pos, int64_t N, float r2, int max_nbrs, GridParams g,
                              const uint32_t* __restrict__ sorted_ids, const int32_t* __restrict__ start,
                              int32_t* __restrict__ counts, const int64_t* __restrict__ offsets, int64_t* __restrict__ edges) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
    const int3 c = cell_of(pos, i, g);
    int n = 0;
    int64_t* out = FILL ? edges + 2 * offsets[i] : nullptr;
    for (int dz = -1; dz <= 1 && n < max_nbrs; ++dz)
        for (int dy = -1; dy <= 1 && n < max_nbrs; ++dy)
            for (int dx = -1; dx <= 1 && n < max_nbrs; ++dx) {
                const int3 cc = make_int3(c.x + dx, c.y + dy, c.z + dz);
                const uint32_t b = bucket_of(cc, g.mask);
                const int s1 = start[b + 1];
                for (int s = start[b]; s < s1 && n < max_nbrs; ++s) {
                    const int64_t j = sorted_ids[s];
                    const int3 cj = cell_of(pos, j, g);
                    if (cj.x != cc.x || cj.y != cc.y || cj.z != cc.z) continue;       // another cell sharing the bucket
                    const float ex = pos[3 * j] - px, ey = pos[3 * j + 1] - py, ez = pos[3 * j + 2] - pz;
                    if (ex * ex + ey * ey + ez * ez <= r2) {
                        if (FILL) {
                            out[2 * n] = i;
                            out[2 * n + 1] = j;
                        }
                        ++n;
                    }
                }
            }
    if (!FILL) counts[i] = n;
}

static uint32_t table_size(int64_t N) {
    uint32_t t = 1024;
    while ((int64_t)t < 2 * N && t < (1u << 30)) t <<= 1;
    return t;
}

static size_t radius_ws(int64_t N) {
    const size_t n = (size_t)(N > 0 ? N : 1);
    return 4 * align_up(n * 4, 256) + align_up(((size_t)table_size(N) + 1) * 4, 256) + sort_workspace(N) + 1024;
}

}  // namespace fcb

using namespace fcb;

extern "C" int fcb_radius_workspace_bytes(int64_t N, size_t* bytes) {
    FCB_REQUIRE(bytes && N >= 0, FCB_E_ARG, "radius_workspace: bad arguments");
    *bytes = radius_ws(N);
    return FCB_OK;
}

// Pass 1: builds the grid in `workspace` (kept for pass 2) and writes counts[N].
extern "C" int fcb_radius_count(const float* pos, int64_t N, float r, int max_neighbors, float ox, float oy, float oz,
                                int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FCB_REQUIRE(pos && counts && workspace, FCB_E_ARG, "radius_count: null pointer");
    FCB_REQUIRE(N >= 0 && N < (1LL << 31) && r > 0.f && max_neighbors > 0, FCB_E_ARG, "radius_count: bad arguments");
    FCB_REQUIRE(workspace_bytes >= radius_ws(N), FCB_E_WORKSPACE, "radius_count: workspace too small");
    if (N == 0) return FCB_OK;
    Arena ar(workspace, workspace_bytes);
    uint32_t* sk = ar.take<uint32_t>((size_t)N);      // sorted keys   \  kept for fcb_radius_fill
    uint32_t* si = ar.take<uint32_t>((size_t)N);      // sorted ids    /
    const uint32_t table = table_size(N);
    int32_t* start = ar.take<int32_t>((size_t)table + 1);
    uint32_t* k0 = ar.take<uint32_t>((size_t)N);
    uint32_t* i0 = ar.take<uint32_t>((size_t)N);
    const size_t sort_b = sort_workspace(N);
    void* sort_ws = ar.take<char>(sort_b);
    FCB_REQUIRE(ar.ok(), FCB_E_WORKSPACE, "radius_count: workspace too small");
    GridParams g{ox, oy, oz, 1.0f / r, table - 1u};
    const unsigned blocks = (unsigned)((N + 255) / 256);
    FCB_LAUNCH("radius_keys", st, k_radius_keys<<<blocks, 256, 0, st>>>(pos, N, g, k0, i0));
    int bits = 0;
    while ((1u << bits) < table) ++bits;
    int rc = sort_pairs(k0, i0, sk, si, N, bits, sort_ws, sort_b, st);
    if (rc) return rc;
    FCB_LAUNCH("radius_starts", st, k_radius_starts<<<(unsigned)((table + 1 + 255) / 256), 256, 0, st>>>(sk, N, table, start));
    FCB_LAUNCH("radius_count", st, k_radius_scan<false><<<blocks, 256, 0, st>>>(pos, N, r * r, max_neighbors, g, si, start, counts,
                                                                             nullptr, nullptr));
    return FCB_OK;
}

// Pass 2: offsets[N+1] = exclusive scan of counts (int64); edges[E x 2] int64 rows (query, found), grouped by query.
extern "C" int fcb_radius_fill(const float* pos, int64_t N, float r, int max_neighbors, float ox, float oy, float oz,
                               const int64_t* offsets, int64_t* edges, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FCB_REQUIRE(pos && offsets && edges && workspace, FCB_E_ARG, "radius_fill: null pointer");
    FCB_REQUIRE(N >= 0 && N < (1LL << 31) && r > 0.f && max_neighbors > 0, FCB_E_ARG, "radius_fill: bad arguments");
    FCB_REQUIRE(workspace_bytes >= radius_ws(N), FCB_E_WORKSPACE, "radius_fill: workspace too small");
    if (N == 0) return FCB_OK;
    Arena ar(workspace, workspace_bytes);
    ar.take<uint32_t>((size_t)N);
    uint32_t* si = ar.take<uint32_t>((size_t)N);
    const uint32_t table = table_size(N);
    int32_t* start = ar.take<int32_t>((size_t)table + 1);
    GridParams g{ox, oy, oz, 1.0f / r, table - 1u};
    FCB_LAUNCH("radius_fill", st, k_radius_scan<true><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(pos, N, r * r, max_neighbors, g, si, start,
                                                                                                nullptr, offsets, edges));
    return FCB_OK;
}
