/* fieldconv_b200 — C ABI of the B200-native FieldConv hot path (libfieldconv_b200.so).
 *
 * The reference (twmitchel/FieldConv) has NO native code on this path: the boundary it
 * exposes is the Python operator `FieldConv.forward(x, supp_edges, supp_sten)`
 * (nn/field_conv.py:104-137) built from ATen ops plus `torch_scatter.scatter_add`
 * (nn/field_conv.py:134), fed by `FCPrecomp.__call__` (transforms/fc_precomp.py:53-97).
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated otherwise; complex = interleaved
 *    (re, im) float pairs exactly like torch.complex64; all buffers are owned by the caller;
 *  - the library allocates no device memory and never synchronises: all work is enqueued on `stream` (a cudaStream_t
 *    passed as void*).  Its only process state: a thread-local error string, an atomic launch counter
 *    (fcb_launch_count), per-device atomic "opt-in shared-memory attribute already set" flags (idempotent), and the
 *    optional per-launch timing facility fcb_profile_* (a debugging aid: NOT thread-safe, creates CUDA events — leave it
 *    off in concurrent use);
 *  - return 0 on success, a negative FCB_E_* code on failure (fcb_last_error() explains);
 *  - feature rows must be 16-byte aligned: channel counts must be even (pad with a zero
 *    channel otherwise — the Python layer does this).
 *
 * Edge record layout (16 bytes, "rec"): { int32 nbr | ring_floor << 27, float t, float wxp_re,
 * float wxp_im } where nbr is the source (by-target plan) or the target (by-source plan),
 * ring_floor = f and t the two-tap radial weights (1-t on ring f, t on ring f+1) of
 * fc_precomp.py:10-27, and wxp the normalised integration weight times transport
 * (fc_precomp.py:87,92).  "rot" = (cos theta, sin theta) of the log-map angle (fc_precomp.py:83).
 */
#ifndef FIELDCONV_B200_H
#define FIELDCONV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCB_OK 0
#define FCB_E_ARG (-1)       /* bad shape / size / null pointer */
#define FCB_E_ALIGN (-2)     /* pointer or channel-count alignment */
#define FCB_E_WORKSPACE (-3) /* workspace too small */
#define FCB_E_CUDA (-4)      /* a CUDA runtime call failed (message has the CUDA error) */
#define FCB_E_UNSUPPORTED (-5)

#define FCB_MAX_BAND_LIMIT 4
#define FCB_MAX_RINGS 32
#define FCB_MAX_VERTICES (1 << 27)

/* flags for fcb_fwd_f32 / fcb_bwd_f32 */
#define FCB_GEMM_SIMT_FP32 0  /* contraction on FP32 FMA (bit-for-bit fp32 semantics) */
#define FCB_GEMM_TC_3XTF32 1  /* tcgen05 tensor cores, error-compensated 3xTF32 (fp32-grade) */
#define FCB_GEMM_TC_TF32 2    /* tcgen05 tensor cores, plain TF32 (looser tolerance) */
#define FCB_GEMM_TC_2XF16 3   /* tcgen05 tensor cores, operands as scaled fp16 (hi, lo) pairs (fp32-grade, fastest) */
#define FCB_GEMM_MASK 0xff
#define FCB_FLAG_HAVE_CONTRIB 0x100 /* fcb_bwd_workspace_bytes: contrib will be supplied, no recompute buffer */

const char* fcb_last_error(void);
int fcb_version(void);
/* number of CUDA kernels this library has launched in this process (monotonic counter; the only
 * other process-wide state besides the thread-local error string) */
unsigned long long fcb_launch_count(void);

/* Optional per-launch timing used by bench.py for the roofline numbers: between enable and
 * disable every kernel launch of this library is bracketed by CUDA events on its stream.
 * Debug facility: not thread-safe, owns its events (the one exception to "allocates nothing").
 * Collect after synchronising: names are '\n'-joined in launch order, ms[i] the durations. */
int fcb_profile_enable(int max_records);
int fcb_profile_disable(void);
int fcb_profile_collect(char* names_buf, size_t names_bytes, float* ms, int capacity, int* count);

/* ------------------------------------------------------------------ plan (K0)
 * Replaces, on the device and without host synchronisation:
 *   transforms/fc_precomp.py:67-74  r = logMag/epsilon, drop edges with r > 1
 *   transforms/fc_precomp.py:10-27  radial two-tap interpolation (ring floor f, weight t)
 *   transforms/fc_precomp.py:87,92  w_j / (1e-12 + sum_{e'->i} w_j') * xp
 *   the implicit grouping that scatter_add's index performs (nn/field_conv.py:134) and that
 *   autograd's gather-backward performs for grad x — as two CSR orders of the kept edges:
 *   by (target, ring floor) and by (source, ring floor), both STABLE in the input order.
 * rowptr_*[N] holds the number of kept edges; rec/rot/perm entries beyond it are untouched.
 * perm_*[p] is the index into the caller's edge list of sorted edge p.
 * ring_radii: R floats = sqrt(k/(R-1)) computed by the caller exactly as fc_precomp.py:12. */
int fcb_plan_workspace_bytes(int64_t E, int64_t N, int R, size_t* bytes);
int fcb_plan_build(const int64_t* edges_ji, const float* log_mag, const float* log_ang,
                   const float* xp /*complex (E)*/, const float* w /*(N)*/,
                   const float* ring_radii, float epsilon, int64_t E, int64_t N, int R,
                   int32_t* rowptr_tgt, void* rec_tgt, float* rot_tgt, int32_t* perm_tgt,
                   int32_t* rowptr_src, void* rec_src, float* rot_src, int32_t* perm_src,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Plan for the dense-stencil drop-in signature forward(x, supp_edges, supp_sten)
 * (nn/field_conv.py:104): CSR by target and by source only; stencil rows are addressed through
 * perm_*.  nbr_*[p] = source (by-target order) / target (by-source order) of sorted edge p. */
int fcb_plan_dense_workspace_bytes(int64_t E, int64_t N, size_t* bytes);
int fcb_plan_build_dense(const int64_t* edges_ji, int64_t E, int64_t N,
                         int32_t* rowptr_tgt, int32_t* nbr_tgt, int32_t* perm_tgt,
                         int32_t* rowptr_src, int32_t* nbr_src, int32_t* perm_src,
                         void* workspace, size_t workspace_bytes, void* stream);

/* FCPrecomp.__call__ outputs in the reference's own dense form (transforms/fc_precomp.py:53-97):
 * expands a plan built by fcb_plan_build from the same inputs into
 *   edges_out (E_kept,2) int64, supp_sten (E_kept,R,2B+1) complex, ln (E_kept) complex = polar(r/eps, theta),
 *   wxp (E_kept) complex, the kept edges in INPUT order (what boolean-mask indexing produces, :69-74).
 * E_kept = rowptr_tgt[N], read back by the caller (the reference synchronises at torch.nonzero, :69). */
int fcb_precomp_workspace_bytes(int64_t E, size_t* bytes);
int fcb_precomp_expand_f32(const int64_t* edges_ji, const float* log_mag, const float* log_ang, float epsilon,
                           int64_t E, int64_t N, int R, int band_limit, const int32_t* rowptr_tgt,
                           const void* rec_tgt, const int32_t* perm_tgt, int64_t E_kept, int64_t* edges_out,
                           float* supp_sten, float* ln, float* wxp, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ------------------------------------------------------------------ operand bounds (optional)
 * The tensor-core paths scale every fp32 operand by a power of two taken from an upper bound of its largest magnitude.
 * For the intermediates (contrib, G) the kernel that writes them tracks the bound for free; for an operand that comes
 * from OUTSIDE a call — the layer input x, the output gradient gy, the operands of fcb_gemm_f32 — the library otherwise
 * spends one extra pass over it, per call.  A caller that runs a whole block knows better: the same x feeds the
 * convolution, the residual TangentLin and both their backwards, and the kernels that write the next layer's input can
 * report its bound as they go.  Every pointer is a device pointer to ONE float and may be NULL (= not available); an
 * upper bound within a factor of ~2 of the true maximum costs no accuracy (the scale is a power of two).
 *   x    in : >= max_i |x_i|, complex modulus, over every row of x the call may gather
 *   gy   in : >= the largest |real or imaginary part| of gy
 *   act  out: max_i |act_i| (modulus) of the activation fcb_fwd_act_* writes — the `x` bound of the next layer
 *   w    in : >= max |W| of the folded filter (a caller that folds the filters of a whole network in one batch gets all
 *             their maxima from one reduction); with it the filter goes from W to the packed tensor-core operand in ONE launch
 * fcb_bound_f32 computes the modulus bound of n complex numbers (what a caller uses when no producer reported one). */
typedef struct fcb_bounds {
    const float* x;
    const float* gy;
    float* act;
    const float* w;
} fcb_bounds;
int fcb_bound_f32(const float* z, int64_t n_complex, float* bound_out, void* stream);

/* ------------------------------------------------------------------ forward (K1 + K2)
 * Replaces nn/field_conv.py:128-137 (+ utils/field.py:40-48, + the weightContrib* reduction
 * :10-33 given the folded weight W[o,c,r,m] = coeff/(2B+1)):
 *   contrib[i, r, m, c] = sum_{e: tgt(e)=i} x[src(e),c] conj(u)^m * sten[e,r,m]   (deterministic
 *                         segmented reduction over the CSR row, fixed order)
 *   y[i, o]             = sum_{c,r,m} contrib[i,r,m,c] * W[o,c,r,m]
 * contrib (N x R*M*Ci complex, k = (r*M + m)*Ci + c) is an OUTPUT the caller may keep for backward.
 * contrib_absmax (one device float, may be NULL): max|contrib| is folded into it with an atomic max
 * (the caller zero-initialises it; several calls on row sub-ranges may share one slot).  It is the
 * operand scale of the FCB_GEMM_TC_2XF16 contraction; hand it back to fcb_bwd_f32 with contrib.
 * W is complex (Co,Ci,R,M) contiguous.  M = 2*band_limit+1. */
int fcb_fwd_workspace_bytes(int64_t N, int Ci, int Co, int band_limit, int R, int flags, size_t* bytes);
int fcb_fwd_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt,
                const float* rot_tgt, float* y, float* contrib, float* contrib_absmax, int64_t N, int Ci,
                int Co, int band_limit, int R, int flags, void* workspace, size_t workspace_bytes,
                void* stream);

/* ------------------------------------------------------------------ backward (K4 + K5)
 * Replaces torch autograd through nn/field_conv.py:128-137 (PyTorch complex convention
 * g = dL/dRe + i dL/dIm):
 *   gW[o,c,r,m] = sum_n conj(contrib[n,r,c,m]) gy[n,o]     (deterministic split reduction)
 *   gx          = softAngle chain rule applied to the transposed gather over the by-source
 *                 CSR of gy Wh conj(sten)                   
 * contrib may be NULL (the forward kept nothing): the weight gradient is then taken from the transposed aggregation G
 * and xhat,  gW[o,c,r,m] = sum_j conj(xhat[j,c,m]) G[j,m,r,o]  (the same sum regrouped by source vertex; needs the
 * by-source plan) — nothing of size N x K is stored or recomputed.  contrib_absmax: the slot filled by the forward (any
 * upper bound of max|contrib| within a factor 2^10 works), or NULL — the 2xFP16 mode then spends one extra pass over
 * contrib on it.  bounds (may be NULL): see fcb_bounds.  gW is complex (Co,Ci,R,M); either of gx / gW may be NULL. */
int fcb_bwd_workspace_bytes(int64_t N, int Ci, int Co, int band_limit, int R, int flags, size_t* bytes);
int fcb_bwd_f32(const float* x, const float* W, const float* gy, const float* contrib,
                const float* contrib_absmax, const int32_t* rowptr_tgt, const void* rec_tgt, const float* rot_tgt,
                const int32_t* rowptr_src, const void* rec_src, const float* rot_src,
                float* gx, float* gW, const fcb_bounds* bounds, int64_t N, int Ci, int Co, int band_limit, int R,
                int flags, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ packed-operand variants (FCB_GEMM_TC_2XF16 only)
 * Same results as fcb_fwd_f32 / fcb_bwd_f32 (nn/field_conv.py:128-137 and its autograd), different intermediate
 * format: the aggregation kernels write contrib (and, in the backward, G) directly as the scaled fp16 (hi, lo) tile
 * images the tensor-core contractions consume ("PK": 128-row x 64-column blocks, SWIZZLE_128B, hi plane then lo
 * plane), so the contraction kernels fetch their A operand with bulk copies instead of converting fp32 in
 * producer warps.  The operand scale must be known BEFORE the aggregation runs, so it comes from the a-priori bound
 *   max|contrib| <= max|x| * max_i sum_{e->i} |wxp_e|        (by-target plan;  <= max|x| for positive vertex weights:
 *   max|G|       <= max|gy| * max_j sum_{e: src=j} |wxp_e|     fc_precomp.py:87 normalises the row mass)
 * fcb_plan_norm computes max_row sum_e |wxp_e| of one CSR order of a plan (one device float, once per mesh).
 * contrib_pk: fcb_pk_contrib_bytes(...) bytes, 128-byte aligned, opaque; contrib_scale (one device float, written
 * by the forward) travels with it to the backward.  fcb_pk_supported: 1 when the layer shape is taken by this path
 * (2*R*M*Ci a multiple of 64 and the 2xFP16 accumulation plans of gemm_h.cu feasible), else 0 — use fcb_fwd_f32.
 * Workspaces: fcb_fwd_workspace_bytes / fcb_bwd_workspace_bytes with the same flags. */
int fcb_plan_norm(const int32_t* rowptr, const void* rec, int64_t N, float* norm_out, void* stream);
int fcb_pk_supported(int64_t N, int Ci, int Co, int band_limit, int R);
int fcb_pk_contrib_bytes(int64_t N, int Ci, int band_limit, int R, size_t* bytes);
int fcb_fwd_pk_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt,
                   const float* rot_tgt, const float* norm_tgt, float* y, void* contrib_pk, float* contrib_scale,
                   const fcb_bounds* bounds, int64_t N, int Ci, int Co, int band_limit, int R, int flags,
                   void* workspace, size_t workspace_bytes, void* stream);
int fcb_bwd_pk_f32(const float* x, const float* W, const float* gy, const void* contrib_pk,
                   const float* contrib_scale, const int32_t* rowptr_tgt, const void* rec_tgt,
                   const float* rot_tgt, const float* norm_tgt, const int32_t* rowptr_src, const void* rec_src,
                   const float* rot_src, const float* norm_src, float* gx, float* gW, const fcb_bounds* bounds,
                   int64_t N, int Ci, int Co, int band_limit, int R, int flags, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- FCResNetBlock epilogue fused into the layer (nn/fc_resnet_block.py:84-88 of the reference):
 *   y   = FieldConv(x) + res        (res: (N, Co) complex64 — the TangentLin residual nn/tangent_lin.py:27-29 — or NULL)
 *   act = modReLU(y, bias)          (nn/tangent_nonlin.py:24-35; bias: Co floats; bias and act may both be NULL)
 * Same arguments as fcb_fwd_f32 / fcb_fwd_pk_f32 plus (res, bias, act).  With the 2xFP16 contraction and an output of at
 * most 128 real columns the epilogue runs inside the contraction kernel (TMEM -> registers -> two stores); otherwise one
 * pointwise kernel follows.  y keeps the PRE-activation values (what the modReLU backward needs). */
int fcb_fwd_act_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt, const float* rot_tgt,
                    float* y, float* contrib, float* contrib_absmax, const float* res, const float* bias, float* act,
                    const fcb_bounds* bounds, int64_t N, int Ci, int Co, int band_limit, int R, int flags, void* workspace,
                    size_t workspace_bytes, void* stream);
int fcb_fwd_act_pk_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt, const float* rot_tgt,
                       const float* norm_tgt, float* y, void* contrib_pk, float* contrib_scale, const float* res,
                       const float* bias, float* act, const fcb_bounds* bounds, int64_t N, int Ci, int Co, int band_limit,
                       int R, int flags, void* workspace, size_t workspace_bytes, void* stream);

/* ---- TransField / LiftBlock aggregations (nn/trans_field.py:96-110 of the reference; csrc/lift.cu).  x: (N, Ci) float32
 * scalar features; lift_sten: (E, R, 2) complex64 (frequencies 0 and 1 of FCPrecomp's stencil) in the caller's edge order;
 * the CSR orders come from fcb_plan_build_dense.  agg: (N, Ci+1, R) complex64 — channels 0..Ci-1 hold
 * sum_{e->i} (x[src] - x[i]) s1[e,r] (= -contribAng), channel Ci the channel-independent sum_{e->i} s1[e,r]; mag: (N, Ci, R)
 * float32 = sum_{e->i} x[src] softAbs(s0[e,r]).  The backward is the adjoint with respect to x (by-source order; `agg` is
 * the forward's output, read for its channel Ci). */
int fcb_lift_aggregate_f32(const float* x, const float* lift_sten, const int32_t* rowptr_tgt, const int32_t* nbr_tgt,
                           const int32_t* perm_tgt, float* agg, float* mag, int64_t N, int Ci, int R, void* stream);
int fcb_lift_aggregate_bwd_f32(const float* g_agg, const float* g_mag, const float* agg, const float* lift_sten, const int32_t* rowptr_src,
                               const int32_t* nbr_src, const int32_t* perm_src, float* gx, int64_t N, int Ci, int R,
                               void* stream);

/* ---- ECHO descriptors (nn/echo.py:94-148 of the reference; csrc/echo.cu).  x: (N, C) complex64 tangent features; ln, wxp:
 * (E,) complex64 as returned by FCPrecomp (transforms/fc_precomp.py:77,92) in the caller's edge order; CSR orders from
 * fcb_plan_build_dense; dmap: the (2 n_bins + 1)^2 raster-cell -> bin table of nn/echo.py:11-27 (int32); hdim = number of
 * bins (n_bins 1..3).  hist: (N, C, hdim) complex64 accumulated histogram (kept for the backward); out: (N, C, hdim)
 * float32 = softAbs(hist).  Deterministic (one thread per histogram, fixed edge order); no host sync.  The backward
 * returns grad x for a real upstream gradient g_out of `out`. */
int fcb_echo_fwd_f32(const float* x, const float* ln, const float* wxp, const int32_t* rowptr_tgt, const int32_t* nbr_tgt,
                     const int32_t* perm_tgt, const int32_t* dmap, float* hist, float* out, int64_t N, int C, int n_bins,
                     int hdim, void* stream);
int fcb_echo_bwd_f32(const float* x, const float* ln, const float* wxp, const int32_t* rowptr_src, const int32_t* nbr_src,
                     const int32_t* perm_src, const int32_t* dmap, const float* hist, const float* g_out, float* gx,
                     int64_t N, int C, int n_bins, int hdim, void* stream);

/* ---- Support-graph construction (transforms/support_graph.py:56-59 of the reference: radius(pos, pos, epsilon,
 * max_num_neighbors=512), self loops included, rows (query j, found i) grouped by j).  Two passes over a hashed uniform grid
 * of cell size r (csrc/radius.cu): fcb_radius_count builds the grid in `workspace` and writes counts[N] (capped at
 * max_neighbors); the caller turns them into offsets[N+1] (exclusive scan, int64) and calls fcb_radius_fill with the SAME
 * workspace, which writes edges[E x 2] int64.  (ox, oy, oz) = lower corner of the bounding box of pos (N x 3 float32). */
int fcb_radius_workspace_bytes(int64_t N, size_t* bytes);
int fcb_radius_count(const float* pos, int64_t N, float r, int max_neighbors, float ox, float oy, float oz, int32_t* counts,
                     void* workspace, size_t workspace_bytes, void* stream);
int fcb_radius_fill(const float* pos, int64_t N, float r, int max_neighbors, float ox, float oy, float oz,
                    const int64_t* offsets, int64_t* edges, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Fused forward (band_limit <= 1): gather -> shared-memory operand tile -> tcgen05 contraction in ONE kernel; the
 * N x K `contrib` of nn/field_conv.py:130-134 never exists in device memory (csrc/fused_fwd.cu).  Same arguments as
 * fcb_fwd_pk_f32 minus the contrib buffers; norm_tgt = fcb_plan_norm of the by-target order; n_feat_rows = rows of x
 * (>= N; larger when the N target rows gather from halo rows behind them, fieldconv_b200/partition.py).  The backward of a layer run
 * this way is fcb_bwd_f32 / fcb_bwd_pk_f32 with contrib == NULL (gW from G and xhat).
 * fcb_fused_supported: 1 when the shape is taken (band_limit <= 1, Ci a multiple of 32, Co even and <= 128, at most 400
 * accumulating MMAs per TMEM accumulator), else 0 — use fcb_fwd_f32 / fcb_fwd_pk_f32. */
int fcb_fused_supported(int Ci, int Co, int band_limit, int R);
int fcb_fwd_fused_workspace_bytes(int Ci, int Co, int band_limit, int R, size_t* bytes);
int fcb_fwd_fused_f32(const float* x, const float* W, const int32_t* rowptr_tgt, const void* rec_tgt, const float* rot_tgt,
                      const float* norm_tgt, float* y, int64_t N, int64_t n_feat_rows, int Ci, int Co, int band_limit, int R,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ dense-stencil forward/backward
 * Exact drop-in for arbitrary supp_sten (E,R,M) complex (nn/field_conv.py:114-116), any number
 * of non-zero rings per edge.  Same outputs as above. */
int fcb_fwd_dense_f32(const float* x, const float* W, const float* sten, const int32_t* rowptr_tgt,
                      const int32_t* nbr_tgt, const int32_t* perm_tgt, float* y, float* contrib,
                      float* contrib_absmax, int64_t N, int Ci, int Co, int band_limit, int R, int flags,
                      void* workspace, size_t workspace_bytes, void* stream);
int fcb_bwd_dense_f32(const float* x, const float* W, const float* gy, const float* contrib,
                      const float* contrib_absmax, const float* sten, const int32_t* rowptr_src, const int32_t* nbr_src,
                      const int32_t* perm_src, float* gx, float* gW, int64_t N, int Ci, int Co,
                      int band_limit, int R, int flags, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------ building blocks (exported for tests)
 * Gauge-aligned aggregation only (K1) / its transpose (first half of K5). */
int fcb_aggregate_f32(const float* feat, const int32_t* rowptr, const void* rec, const float* rot,
                      float* out, int64_t N, int C, int band_limit, int R, int transpose, void* stream);
/* Real fp32 GEMM C[MxN] = A*B (trans_a=0: A is MxK row-major; trans_a=1: A is KxM row-major),
 * B is KxN row-major; batch >= 1 with element strides; split_k >= 1 writes per-split partials into
 * the workspace and reduces them in a fixed order (deterministic).  flags & FCB_GEMM_MASK selects
 * the FP32-FMA kernel or a tcgen05 tensor-core kernel (3xTF32 / TF32 / 2xFP16) when its accumulation plan fits.
 * The workspace size comes from fcb_gemm_workspace_bytes with the same arguments.
 * a_bound / b_bound (device floats, may be NULL): upper bounds of max|A|, max|B| for the 2xFP16 operand scales (see
 * fcb_bounds); each one given saves the pass the library would otherwise take over that operand. */
int fcb_gemm_workspace_bytes(int64_t M, int N, int64_t K, int trans_a, int batch, int split_k, int flags,
                             size_t* bytes);
/* 1 if fcb_gemm_f32 with these flags would run on the tensor cores inside the fp32 parity budget (the
 * accumulation plan — column chunks x TMEM accumulators — fits), 0 if it would fall back to FP32 FMA. */
int fcb_gemm_tc_feasible(int N, int64_t K, int trans_a, int split_k, int flags);
int fcb_gemm_f32(const float* A, const float* B, float* C, int64_t M, int N, int64_t K,
                 int64_t lda, int64_t ldb, int64_t ldc, int trans_a, int batch, int64_t stride_a,
                 int64_t stride_b, int64_t stride_c, int split_k, const float* a_bound, const float* b_bound,
                 void* workspace, size_t workspace_bytes, int flags, void* stream);
/* Stable LSD radix sort of (key,value) uint32 pairs on the low `bits` bits of the key.
 * Result lands in keys_out/vals_out; keys_in/vals_in are clobbered. */
int fcb_sort_workspace_bytes(int64_t n, size_t* bytes);
int fcb_sort_pairs_u32(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                       int64_t n, int bits, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ block epilogue pieces
 * TangentNonLin / modReLU, nn/tangent_nonlin.py:24-35: y = relu(|x|+b_c) x/|x|, origin entries
 * passed through.  Backward returns gx and per-block partial bias gradients reduced in fixed
 * order into gb (C floats); gx_bound (device float, may be NULL) receives the largest |component| of gx — the `gy`
 * bound of the layer whose output gradient gx is (fcb_bounds). */
int fcb_modrelu_fwd_f32(const float* x, const float* bias, float* y, int64_t N, int C, void* stream);
int fcb_modrelu_bwd_workspace_bytes(int64_t N, int C, size_t* bytes);
int fcb_modrelu_bwd_f32(const float* x, const float* bias, const float* gy, float* gx, float* gb, float* gx_bound,
                        int64_t N, int C, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FIELDCONV_B200_H */
