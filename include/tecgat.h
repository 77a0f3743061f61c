/* tecgat.h -- C ABI of libtecgat.so: the B200 (sm_100a) GATv2 SpatialEncoder + haversine graph path.
 *
 * The reference (PANXIONG-CN/TEC-MoLLM) is pure Python and has NO native ABI for this path; each
 * entry point below replaces one step that the reference reaches through torch_geometric / ATen /
 * scikit-learn / scipy, cited as reference file:line (paths relative to /root/reference) or as the
 * step of PyG's GATv2Conv.forward.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a device pointer, *_host a host pointer;
 *   - all device work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden syncs
 *     except where stated (plan creation, graph builder: one-time set-up calls);
 *   - the caller owns every buffer (PyTorch tensors on the Python side); the only library-owned
 *     object is the immutable graph plan;
 *   - every function returns 0 on success, a negative TECGAT_E* code otherwise, and never throws or
 *     exits; tecgat_last_error() returns a thread-local message for the last failure;
 *   - deterministic: no floating-point atomics anywhere.
 *
 * Row-major layouts, R = S*N rows (snapshot-major: row = s*N + node):
 *   x  (R, F)   fp32        xl, xr (R, H*C) storage dtype (fp32, or bf16 for the autocast contract)
 *   y  (R, H*C) fp32        stat   (R, H)   fp32  (log2-sum-exp of the row's scores: alpha = exp2(e - stat), saved for backward)
 */
#ifndef TECGAT_H_
#define TECGAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TECGAT_ABI_VERSION 4

/* error codes */
#define TECGAT_OK 0
#define TECGAT_EINVAL (-1)  /* bad argument (shape, dtype, null pointer, unsupported size) */
#define TECGAT_ECUDA (-2)   /* CUDA runtime error (message carries cudaGetErrorString) */
#define TECGAT_ENOMEM (-3)  /* host allocation failure */
#define TECGAT_ENOSUP (-4)  /* configuration not supported by the compiled kernels */

/* storage dtype of xl / xr / dxl / dxr */
#define TECGAT_F32 0
#define TECGAT_BF16 1

/* snapshot modes */
#define TECGAT_MODE_SHARED 0  /* the N-node graph applies to every snapshot (intended semantics)   */
#define TECGAT_MODE_LITERAL 1 /* modules.py:353-356 exactly as written: only snapshot 0 sees edges */

/* projection implementations */
#define TECGAT_PROJ_TC 0   /* tcgen05 tensor-core GEMM fed by bulk-TMA (product path)               */
#define TECGAT_PROJ_FFMA 1 /* CUDA-core kernel; used to cross-check the tensor-core path in tests   */

typedef struct tecgat_plan tecgat_plan_t;

int tecgat_abi_version(void);
const char *tecgat_last_error(void);

/* ---- graph plan: replaces remove_self_loops + add_self_loops done on EVERY forward by PyG
 *      (GATv2Conv.forward, called at src/model/modules.py:356) with a one-time,
 *      cached structure: destination-sorted CSR + source-sorted CSR (atomic-free backward), and for the
 *      forward and the backward kernel one tiling each of the node axis (tile_nodes_fwd / tile_nodes_bwd
 *      consecutive nodes per tile, multiples of 8) with the tile's row window and its ELL slab.
 *      Synchronises `stream` (the edge list is read back to the host once). */
int tecgat_plan_create(const int64_t *edge_index_dev, int64_t num_edges, int32_t num_nodes,
                       int32_t tile_nodes_fwd, int32_t tile_nodes_bwd, void *stream,
                       tecgat_plan_t **plan_out);
int tecgat_plan_destroy(tecgat_plan_t *plan);
/* info[0]=E (kept edges + N self loops) [1]=max in-degree [2]=max out-degree [3]=forward tiles
 * [4]=tile_nodes_fwd [5]=max forward window rows [6]=num_nodes [7]=kept (non-self) edges
 * [8]=backward tiles [9]=tile_nodes_bwd [10]=max backward window rows
 * [11]=1 when the graph is banded and the sliding-window backward tiling was built (edge_bwd_sw.cu) */
int tecgat_plan_info(const tecgat_plan_t *plan, int64_t *info12_host);
/* Host copies for tests: rowptr (N+1), col (E) = source node of CSR slot k, eid (E) = index of that
 * edge in PyG's post-surgery edge order (kept edges in input order, then self loops).  Every CSR row
 * starts with the node's self loop. */
int tecgat_plan_export(const tecgat_plan_t *plan, int32_t *rowptr_host, int32_t *col_host,
                       int32_t *eid_host);

/* ---- projections: replace lin_l / lin_r (two torch.nn.functional.linear -> cuBLAS addmm, SURVEY.md
 *      K1) and their autograd backward (K10).  [xl|xr] = x [Wl;Wr]^T + [bl|br] in ONE pass over x. */
int tecgat_project_fwd(const float *x_dev, const float *wl_dev, const float *bl_dev,
                       const float *wr_dev, const float *br_dev, void *xl_dev, void *xr_dev,
                       int64_t rows, int32_t in_channels, int32_t hc, int32_t dtype, int32_t impl,
                       void *stream);
/* workspace bytes needed by tecgat_project_bwd */
int64_t tecgat_project_bwd_workspace(int64_t rows, int32_t in_channels, int32_t hc, int32_t impl);
/* dx = dxl Wl + dxr Wr (skipped when dx_dev is NULL); dWl = dxl^T x; dbl = sum dxl; same for r.   */
int tecgat_project_bwd(const void *dxl_dev, const void *dxr_dev, const float *x_dev,
                       const float *wl_dev, const float *wr_dev, float *dx_dev, float *dwl_dev,
                       float *dbl_dev, float *dwr_dev, float *dbr_dev, void *workspace_dev,
                       int64_t rows, int32_t in_channels, int32_t hc, int32_t dtype, int32_t impl,
                       void *stream);

/* dx += dxl Wl + dxr Wr (dx_dev pre-loaded by the caller, e.g. with the residual branch's gradient from
 * tecgat_residual_permute_bwd: replaces autograd's separate accumulation pass); otherwise as tecgat_project_bwd
 * with impl = TECGAT_PROJ_TC and the same workspace.  tecgat_project_bwd_acc_supported: 1 when (F, hc) is in range. */
int tecgat_project_bwd_acc_supported(int32_t in_channels, int32_t hc);
int tecgat_project_bwd_acc(const void *dxl_dev, const void *dxr_dev, const float *x_dev,
                           const float *wl_dev, const float *wr_dev, float *dx_dev, float *dwl_dev,
                           float *dbl_dev, float *dwr_dev, float *dbr_dev, void *workspace_dev,
                           int64_t rows, int32_t in_channels, int32_t hc, int32_t dtype, void *stream);

/* ---- fused edge phase: replaces gather + LeakyReLU*att + segment softmax + dropout + scatter-add
 *      + bias (~20 ATen launches in PyG) with one kernel over all snapshots.
 *      dropout_p == 0 disables dropout; otherwise the keep bit of (snapshot s, CSR slot k, head h) is the
 *      counter-based hash restated by tecgat_dropout_mask_host.                                          */
int tecgat_edge_fwd(const tecgat_plan_t *plan, const void *xl_dev, const void *xr_dev,
                    const float *att_dev, const float *bias_dev, float *y_dev, float *stat_dev,
                    int32_t snapshots, int32_t heads, int32_t out_channels, float negative_slope,
                    float dropout_p, uint64_t seed, int32_t mode, int32_t dtype, void *stream);
int64_t tecgat_edge_bwd_workspace(const tecgat_plan_t *plan, int32_t snapshots, int32_t heads,
                                  int32_t out_channels);
/* backward of the edge phase, atomic-free: every node's lane reduces its incoming edges (d xr) and its
 * outgoing edges (d xl) from the two CSR orientations; d att / d bias via deterministic two-stage sums. */
int tecgat_edge_bwd(const tecgat_plan_t *plan, const void *xl_dev, const void *xr_dev,
                    const float *att_dev, const float *bias_dev, const float *y_dev,
                    const float *stat_dev, const float *gy_dev, void *dxl_dev, void *dxr_dev,
                    float *datt_dev, float *dbias_dev, void *workspace_dev, int32_t snapshots,
                    int32_t heads, int32_t out_channels, float negative_slope, float dropout_p,
                    uint64_t seed, int32_t mode, int32_t dtype, void *stream);

/* ---- one call per direction: what GATv2Conv.forward (src/model/modules.py:356) and its autograd backward are for the
 *      reference.  tecgat_forward = tecgat_project_fwd + tecgat_edge_fwd.  tecgat_backward = tecgat_edge_bwd +
 *      tecgat_project_bwd with ONE fixed-order finish of all six parameter gradients (5 launches per training step in
 *      total, against ~45 ATen launches in PyG).
 *        seed_dev      : when non-NULL the dropout seed is read from device memory (see tecgat_seed_advance) and `seed` is
 *                        ignored -- a captured CUDA graph then draws a fresh mask on every replay;
 *        dx_accumulate : dx += (dx pre-loaded with the residual branch's gradient, see tecgat_project_bwd_acc);
 *        grad_accumulate : the six parameter gradients are ADDED to what the buffers hold (the caller's .grad storage:
 *                        replaces autograd's six accumulation kernels); needs tecgat_backward_fused_supported() == 1.
 *      workspace: tecgat_backward_workspace() bytes. */
int tecgat_forward(const tecgat_plan_t *plan, const float *x_dev, const float *wl_dev, const float *bl_dev,
                   const float *wr_dev, const float *br_dev, const float *att_dev, const float *bias_dev,
                   void *xl_dev, void *xr_dev, float *y_dev, float *stat_dev, int32_t snapshots,
                   int32_t in_channels, int32_t heads, int32_t out_channels, float negative_slope,
                   float dropout_p, uint64_t seed, const uint64_t *seed_dev, int32_t mode, int32_t dtype,
                   int32_t impl, void *stream);
/* tecgat_forward that ALSO stores the output rows into a wider row-major tensor:
 *   y_wide_dev[row * ld_wide + c] = y[row, c],  c < heads * out_channels
 * (y_wide_dev points at the first column of this call's slice; 8-byte aligned, ld_wide even, in floats).  A layer with more
 * than two heads runs as independent head pairs on parameter slices -- the heads of GATv2Conv (modules.py:329-336) only meet in
 * the concatenation -- and each pair writes its columns of the (rows, H*C) result in place of a concatenation pass. */
int tecgat_forward_into(const tecgat_plan_t *plan, const float *x_dev, const float *wl_dev, const float *bl_dev,
                        const float *wr_dev, const float *br_dev, const float *att_dev, const float *bias_dev,
                        void *xl_dev, void *xr_dev, float *y_dev, float *stat_dev, int32_t snapshots,
                        int32_t in_channels, int32_t heads, int32_t out_channels, float negative_slope,
                        float dropout_p, uint64_t seed, const uint64_t *seed_dev, int32_t mode, int32_t dtype,
                        int32_t impl, float *y_wide_dev, int64_t ld_wide, void *stream);
int64_t tecgat_backward_workspace(const tecgat_plan_t *plan, int32_t snapshots, int32_t in_channels,
                                  int32_t heads, int32_t out_channels, int32_t impl);
int tecgat_backward_fused_supported(int32_t in_channels, int32_t hc, int32_t impl);
int tecgat_backward(const tecgat_plan_t *plan, const float *x_dev, const float *wl_dev, const float *wr_dev,
                    const float *att_dev, const float *bias_dev, const void *xl_dev, const void *xr_dev,
                    const float *y_dev, const float *stat_dev, const float *gy_dev, void *dxl_dev,
                    void *dxr_dev, float *dx_dev, int32_t dx_accumulate, float *dwl_dev, float *dbl_dev,
                    float *dwr_dev, float *dbr_dev, float *datt_dev, float *dbias_dev,
                    int32_t grad_accumulate, void *workspace_dev, int32_t snapshots, int32_t in_channels,
                    int32_t heads, int32_t out_channels, float negative_slope, float dropout_p, uint64_t seed,
                    const uint64_t *seed_dev, int32_t mode, int32_t dtype, int32_t impl, void *stream);

/* Device-resident dropout seed (replaces the Philox state torch.nn.functional.dropout advances inside PyG's message
 * step): state_dev = {seed, counter}; writes a fresh 64-bit seed to *seed_out_dev and bumps the counter, on `stream`. */
int tecgat_seed_advance(uint64_t *state_dev, uint64_t *seed_out_dev, void *stream);

/* Instrumentation.  tecgat_launch_count: kernels launched by this library since it was loaded.  tecgat_phase_timing(1):
 * tecgat_forward / tecgat_backward record CUDA events between their phases on the caller's stream; tecgat_phase_times
 * synchronises them and returns (then clears) the accumulated milliseconds of {proj_fwd, edge_fwd, edge_bwd, proj_bwd}. */
int64_t tecgat_launch_count(void);
int tecgat_phase_timing(int32_t enable);
int tecgat_phase_times(double *ms_out4_host);

/* Host restatement of the kernels' counter-based dropout RNG (pure integer arithmetic), so tests can
 * hand the oracle exactly the mask the kernels used.  Global slot g = snapshot * edges_per_snapshot + CSR slot;
 * keep_host: (count, heads) bytes for g = first_slot .. first_slot + count - 1.                      */
int tecgat_dropout_mask_host(uint64_t seed, int64_t first_slot, int64_t count, int32_t heads,
                             float dropout_p, int64_t edges_per_snapshot, uint8_t *keep_host);

/* ---- caller glue either side of the encoder: replaces the residual add and the
 *      (L*B, N, C) -> (B, N, L, C) permute copy of TEC_MoLLM.forward (src/model/tec_mollm.py:94,100) with one
 *      pass, and their autograd backward with one pass.  The permute of :84 needs no kernel: snapshots are
 *      independent, so the encoder takes the (B, L, N, C) tensor as it is.  All tensors fp32, contiguous.
 *        fwd:  z[b, n, l, :] = x[b, l, n, :] + y[b, l, n, :]     (y_dev may be NULL: pure transposition)
 *        bwd:  g[b, l, n, :] = gz[b, n, l, :]                    (gradient of x's residual branch AND of y)  */
int tecgat_residual_permute_fwd(const float *x_dev, const float *y_dev, float *z_dev, int32_t batch,
                                int32_t steps, int32_t nodes, int32_t channels, void *stream);
int tecgat_residual_permute_bwd(const float *gz_dev, float *g_dev, int32_t batch, int32_t steps,
                                int32_t nodes, int32_t channels, void *stream);

/* ---- producer side of the spatial block: replaces SpatioTemporalEmbedding.forward
 *      (src/model/modules.py:230-264: five embedding gathers, four adds, one concat) and its autograd backward.
 *        out[s, n, 0:Cr]     = x[s, n, :]
 *        out[s, n, Cr:Cr+De] = node[n] + (((tod[tf[s,0]] + doy[tf[s,1]]) + year[tf[s,2]]) + season[tf[s,3]])
 *      (the reference's association order: bit-identical to torch).  tf_dev: (S, 4) int32, one row per snapshot -- the
 *      reference's time features are per (batch, step) and only EXPANDED over the nodes (train.py:64-65).  Indices are
 *      clamped to the table.  x (S, N, Cr), out (S, N, Cr+De), tables (rows, De): fp32, contiguous.
 *      tecgat_embed_bwd: ge = d out; d node[n] = sum_s ge[s, n, Cr:], d tab[i] = sum over the snapshots that index row i of
 *      sum_n ge[s, n, Cr:]; fixed-order fp64 reductions, no atomics (torch's embedding backward is atomic);
 *      accumulate = 1 adds onto the buffers.  Needs De = 16 (the reference's d_emb) and an even Cr. */
int tecgat_embed_fwd(const float *x_dev, const int32_t *tf_dev, const float *node_dev, const float *tod_dev,
                     const float *doy_dev, const float *year_dev, const float *season_dev, float *out_dev,
                     int32_t snapshots, int32_t nodes, int32_t raw_channels, int32_t emb_dim, int32_t n_tod,
                     int32_t n_doy, int32_t n_year, int32_t n_season, void *stream);
int64_t tecgat_embed_bwd_workspace(int32_t snapshots, int32_t nodes, int32_t emb_dim);
int tecgat_embed_bwd(const float *ge_dev, const int32_t *tf_dev, float *dnode_dev, float *dtod_dev,
                     float *ddoy_dev, float *dyear_dev, float *dseason_dev, void *workspace_dev,
                     int32_t snapshots, int32_t nodes, int32_t raw_channels, int32_t emb_dim, int32_t n_tod,
                     int32_t n_doy, int32_t n_year, int32_t n_season, int32_t accumulate, void *stream);

/* ---- TemporalEncoder, the part between the convolutions of a Multi_Scale_Conv_Block (src/model/modules.py:13-60; SURVEY.md
 *      8f N3): replaces, per branch k = 3 / 5 / 7, GroupNorm(1, C) (:26) + GELU (:27), the channel concat (:52) and the strided
 *      read of the 1x1 convolution (:36-41, :55) with one pass.  y: (samples, branches*C, L) -- the three branch convolutions'
 *      outputs in concatenated channel order; z: (samples, branches*C, ceil(L / stride)) = only the positions the strided 1x1
 *      convolution reads, already normalised, affine-mapped (gamma, beta: (branches, C)) and passed through the exact GELU.
 *      mean / rstd: (samples, branches) saved for backward.  Backward: d y from d z (GELU', affine, GroupNorm Jacobian) and
 *      d gamma / d beta by fixed-order two-stage sums (no atomics).  dtypes: TECGAT_F32 / TECGAT_BF16 for y (and d y) and for
 *      z (and d z); arithmetic in fp32.  channels * L <= 8192. */
int tecgat_gn_gelu_fwd(const void *y_dev, const float *gamma_dev, const float *beta_dev, void *z_dev,
                       float *mean_dev, float *rstd_dev, int64_t samples, int32_t branches, int32_t channels,
                       int32_t length, int32_t stride, float eps, int32_t y_dtype, int32_t z_dtype, void *stream);
int64_t tecgat_gn_gelu_bwd_workspace(int64_t samples, int32_t branches, int32_t channels);
int tecgat_gn_gelu_bwd(const void *y_dev, const float *gamma_dev, const float *beta_dev, const float *mean_dev,
                       const float *rstd_dev, const void *dz_dev, void *dy_dev, float *dgamma_dev,
                       float *dbeta_dev, void *workspace_dev, int64_t samples, int32_t branches, int32_t channels,
                       int32_t length, int32_t stride, int32_t y_dtype, int32_t z_dtype, void *stream);

/* ---- graph builder: replaces calculate_haversine_distance_matrix (src/graph/graph_constructor.py:34-59,
 *      sklearn haversine_distances in fp64), construct_binary_adjacency (:61-81, inclusive `<=`, zero
 *      diagonal), symmetrically_normalize_adjacency (:99-128) and the COO extraction of
 *      convert_to_pyg_and_save (:141-144).  Inputs are per-node coordinates in RADIANS (device, fp64).
 *
 *  tecgraph_distance_rows : dense rows D[r0:r1, 0:n] in km (fp64), for the dense-API mirror.
 *  tecgraph_edges_count   : pass 1 -- per-row neighbour counts and their device-side scan, total in *total_host.
 *                           One warp per row over 32-column chunks; a chunk is skipped when its latitude or
 *                           longitude range alone puts it beyond the threshold; cos(lat) is evaluated once per
 *                           node; "d <= thr" is decided on the haversine argument r against sin^2(thr / 2R)
 *                           (no asin / sqrt per pair).  Pairs inside a relative guard band of the threshold
 *                           are re-evaluated on the host with the reference's exact libm formula so the edge
 *                           SET is bit-exact (synchronises the stream).
 *  tecgraph_edges_fill    : pass 2 -- edge_index (2, E) int64 row-major (row ascending, column ascending
 *                           inside a row, exactly scipy's COO order) and edge_weight fp32 =
 *                           fp32((1/sqrt(deg_r) * 1.0) * 1/sqrt(deg_c)).                              */
typedef struct tecgraph_ctx tecgraph_ctx_t;
int tecgraph_distance_rows(const double *lat_dev, const double *lon_dev, int64_t n, int64_t r0,
                           int64_t r1, double radius_km, double *out_dev, void *stream);
int tecgraph_edges_count(const double *lat_dev, const double *lon_dev, int64_t n, double thr_km,
                         double radius_km, void *stream, tecgraph_ctx_t **ctx_out,
                         int64_t *total_host, int64_t *ambiguous_host);
int tecgraph_edges_fill(tecgraph_ctx_t *ctx, int64_t *edge_index_dev, float *edge_weight_dev,
                        void *stream);
int tecgraph_ctx_destroy(tecgraph_ctx_t *ctx);
/* stats4 = {count-kernel ms, fill-kernel ms (0 before tecgraph_edges_fill), pairs evaluated by the count pass, guard-band
 * pairs re-checked on the host}: CUDA-event times of the two edge kernels (synchronises them). */
int tecgraph_ctx_stats(tecgraph_ctx_t *ctx, double *stats4_host);

#ifdef __cplusplus
}
#endif
#endif /* TECGAT_H_ */
