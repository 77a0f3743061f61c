// edge_fwd.cu -- fused GATv2 edge phase, forward: for every snapshot and destination node
//     e_ij = att . LeakyReLU(xl_j + xr_i),  alpha = softmax_j(e_ij),  y_i = sum_j alpha_ij q_ij xl_j + bias
// in ONE kernel (PyG: gather, add, leaky_relu, mul, sum, scatter-max, exp, scatter-add, div, dropout, mul,
// scatter-add, bias = ~20 ATen launches and 4-6 materialised (S*E, H, C) tensors; SURVEY.md K4-K9, reached
// from /root/reference/src/model/modules.py:356).  Online softmax per lane, no atomics, no cross-lane traffic.
#include "edge_common.cuh"

namespace tg {

struct EdgeFwdArgs {
    const void *xl, *xr;
    const float *att, *bias;
    float *y, *m, *den;
    const int32_t *rowptr, *col, *tile_lo, *tile_hi;
    int32_t N, T, num_tiles, S, H;
    int64_t E;
    float slope, inv_keep;
    uint32_t drop_thr;
    uint64_t seed;
    int32_t literal;
    int32_t win_rows_smem;  // rows of the xl window that fit in the shared-memory slab
};

template <int C, typename ST, bool SM>
__device__ __forceinline__ void edge_fwd_body(const EdgeFwdArgs &a, unsigned char *smem_raw) {
    constexpr int CP = (C + 1) / 2;
    const int tid = threadIdx.x;
    const int tile = blockIdx.x % a.num_tiles;
    const int snap = blockIdx.x / a.num_tiles;
    const int H = a.H, HC = H * C;
    const int n0 = tile * a.T;
    const int n1 = min(a.N, n0 + a.T);
    const int nt = n1 - n0;
    const bool self_only = a.literal && snap > 0;  // modules.py:353-356 as written: rows >= N have no edges
    int lo = a.tile_lo[tile], hi = a.tile_hi[tile];
    if (self_only) { lo = n0; hi = n1; }
    const int win = hi - lo;

    const ST *xl_g = static_cast<const ST *>(a.xl) + (static_cast<int64_t>(snap) * a.N + lo) * HC;
    const ST *xr_g = static_cast<const ST *>(a.xr) + (static_cast<int64_t>(snap) * a.N + n0) * HC;

    // shared layout: [mbarrier 16][xr slab + 16][y tile fp32][xl window + 16]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned char *xr_base = smem_raw + 16;
    float *y_s = reinterpret_cast<float *>(xr_base + round16(a.T * HC * sizeof(ST)) + 16);
    unsigned char *xl_base = reinterpret_cast<unsigned char *>(y_s) + round16(a.T * HC * sizeof(float));

    const CopyPlan cr = plan_copy(xr_g, xr_base, nt * HC * (uint32_t)sizeof(ST));
    const CopyPlan cl = plan_copy(xl_g, xl_base, SM ? win * HC * (uint32_t)sizeof(ST) : 0u);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, cr.mid + cl.mid);
        issue_copy_bulk(cr, bar);
        issue_copy_bulk(cl, bar);
    }
    copy_ragged(cr, tid);
    copy_ragged(cl, tid);
    __syncthreads();
    mbar_wait(bar, 0);

    const int node_l = tid / H;
    const int h = tid - node_l * H;
    if (node_l < nt) {
        const int d = n0 + node_l;
        float2 xr_i[CP], att_h[CP], acc[CP];
        load_row<C>(reinterpret_cast<const ST *>(cr.s) + node_l * HC + h * C, xr_i);
        load_row<C>(a.att + h * C, att_h);
#pragma unroll
        for (int i = 0; i < CP; ++i) acc[i] = make_float2(0.f, 0.f);
        const int k1 = __ldg(a.rowptr + d + 1);
        const int k0 = self_only ? k1 - 1 : __ldg(a.rowptr + d);  // the self loop is the row's last slot
        // neighbour rows: shared window (SM) or global/L2 gather (window too large for shared memory)
        const ST *src_base = (SM ? reinterpret_cast<const ST *>(cl.s) : xl_g) + h * C;
        const uint32_t key = a.drop_thr ? dropout_snapshot_key(a.seed, (uint32_t)snap) : 0u;
        const float2 slope2 = make_float2(a.slope, a.slope);
        constexpr float kLog2e = 1.4426950408889634f;
        float mx = -INFINITY, l = 0.f;  // running max (log2 domain) and running sum
        for (int k = k0; k < k1; ++k) {
            const int j = __ldg(a.col + k) - lo;
            float2 xj[CP], s[CP], z[CP];
            load_row<C>(src_base + j * HC, xj);
            const float e = edge_score<C, ST>(att_h, xj, xr_i, slope2, s, z) * kLog2e;
            float q = 1.f;
            if (a.drop_thr) q = dropout_bits16(key, (uint32_t)k, (uint32_t)h) >= a.drop_thr ? a.inv_keep : 0.f;
            const float mn = fmaxf(mx, e);
            const float sc = fast_exp2(mx - mn);
            const float pe = fast_exp2(e - mn);
            l = fmaf(l, sc, pe);
            const float2 sc2 = make_float2(sc, sc), w2 = make_float2(pe * q, pe * q);
#pragma unroll
            for (int i = 0; i < CP; ++i) acc[i] = __ffma2_rn(acc[i], sc2, __fmul2_rn(w2, xj[i]));
            mx = mn;
        }
        const float den = l + 1e-16f;
        const float inv = 1.f / den;
        float2 bias_h[CP], out[CP];
        load_row<C>(a.bias + h * C, bias_h);
        const float2 inv2 = make_float2(inv, inv);
#pragma unroll
        for (int i = 0; i < CP; ++i) out[i] = __ffma2_rn(acc[i], inv2, bias_h[i]);
        store_row<C>(y_s + node_l * HC + h * C, out);
        const int64_t r = (static_cast<int64_t>(snap) * a.N + d) * H + h;
        a.m[r] = mx;   // softmax shift in the log2 domain (m * log2 e); consumed only by edge_bwd
        a.den[r] = den;
    }
    __syncthreads();
    float *y_g = a.y + (static_cast<int64_t>(snap) * a.N + n0) * HC;
    for (int i = tid; i < nt * HC; i += blockDim.x) y_g[i] = y_s[i];
}

template <int C, typename ST>
__global__ void __launch_bounds__(256) edge_fwd_kernel(const EdgeFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile = blockIdx.x % a.num_tiles;
    const int snap = blockIdx.x / a.num_tiles;
    int win = a.tile_hi[tile] - a.tile_lo[tile];
    if (a.literal && snap > 0) win = min(a.N, (tile + 1) * a.T) - tile * a.T;
    if (win <= a.win_rows_smem)
        edge_fwd_body<C, ST, true>(a, smem_raw);
    else
        edge_fwd_body<C, ST, false>(a, smem_raw);
}

template <int C, typename ST>
static int launch_fwd(const EdgeFwdArgs &a, int threads, size_t fixed_smem, int max_win, cudaStream_t st) {
    EdgeFwdArgs b = a;
    const int HC = a.H * C;
    const size_t row_bytes = size_t(HC) * sizeof(ST);
    int rows_fit = int((size_t(kSmemBudget) - fixed_smem - 16) / row_bytes);
    b.win_rows_smem = rows_fit < max_win ? rows_fit : max_win;
    if (b.win_rows_smem < 0) b.win_rows_smem = 0;
    const size_t smem = fixed_smem + round16(uint32_t(b.win_rows_smem * row_bytes)) + 16;
    auto kern = edge_fwd_kernel<C, ST>;
    TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = int64_t(a.num_tiles) * a.S;
    kern<<<(unsigned)grid, threads, smem, st>>>(b);
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

}  // namespace tg

extern "C" int tecgat_edge_fwd(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att,
                               const float *bias, float *y, float *m, float *den, int32_t snapshots, int32_t heads,
                               int32_t out_channels, float negative_slope, float dropout_p, uint64_t seed, int32_t mode,
                               int32_t dtype, void *stream) {
    using namespace tg;
    TG_REQUIRE(plan && xl && xr && att && bias && y && m && den, TECGAT_EINVAL, "edge_fwd: NULL argument");
    TG_REQUIRE(snapshots > 0 && heads > 0 && out_channels > 0, TECGAT_EINVAL, "edge_fwd: non-positive size");
    TG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, TECGAT_EINVAL, "edge_fwd: dropout_p %f outside [0, 1)", dropout_p);
    TG_REQUIRE(negative_slope >= 0.f && negative_slope <= 1.f, TECGAT_ENOSUP, "edge_fwd: negative_slope %f outside [0, 1]", negative_slope);
    TG_REQUIRE(mode == TECGAT_MODE_SHARED || mode == TECGAT_MODE_LITERAL, TECGAT_EINVAL, "edge_fwd: bad mode %d", mode);
    TG_REQUIRE(dtype == TECGAT_F32 || dtype == TECGAT_BF16, TECGAT_EINVAL, "edge_fwd: bad dtype %d", dtype);
    const int threads = ((plan->tile_nodes * heads + 31) / 32) * 32;
    TG_REQUIRE(threads <= 256, TECGAT_ENOSUP, "edge_fwd: tile_nodes (%d) * heads (%d) exceeds 256 lanes; build the plan with a smaller tile",
               plan->tile_nodes, heads);
    TG_REQUIRE(int64_t(plan->num_tiles) * snapshots < (int64_t(1) << 31), TECGAT_ENOSUP, "edge_fwd: grid too large");
    EdgeFwdArgs a;
    a.xl = xl; a.xr = xr; a.att = att; a.bias = bias; a.y = y; a.m = m; a.den = den;
    a.rowptr = plan->rowptr_in; a.col = plan->col_in; a.tile_lo = plan->tile_lo; a.tile_hi = plan->tile_hi;
    a.N = plan->num_nodes; a.T = plan->tile_nodes; a.num_tiles = plan->num_tiles; a.S = snapshots; a.H = heads;
    a.E = plan->num_edges;
    a.slope = negative_slope;
    a.drop_thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u;
    a.inv_keep = 1.f / (1.f - dropout_p);
    a.seed = seed;
    a.literal = (mode == TECGAT_MODE_LITERAL);
    a.win_rows_smem = 0;
    const int HC = heads * out_channels;
    const size_t esz = dtype == TECGAT_F32 ? 4 : 2;
    const size_t fixed = 16 + round16(uint32_t(a.T * HC * esz)) + 16 + round16(uint32_t(a.T * HC * 4));
    TG_REQUIRE(fixed + 64 < size_t(kSmemBudget), TECGAT_ENOSUP, "edge_fwd: tile of %d nodes x %d channels does not fit shared memory",
               a.T, HC);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TG_CASE(CC)                                                                                   \
    case CC:                                                                                          \
        return dtype == TECGAT_F32 ? launch_fwd<CC, float>(a, threads, fixed, plan->max_window, st)   \
                                   : launch_fwd<CC, __nv_bfloat16>(a, threads, fixed, plan->max_window, st);
    switch (out_channels) {
        TG_FOR_EACH_C(TG_CASE)
        default:
            break;
    }
#undef TG_CASE
    tecgat_set_error("edge_fwd: out_channels=%d is not among the compiled channel counts", out_channels);
    return TECGAT_ENOSUP;
}
