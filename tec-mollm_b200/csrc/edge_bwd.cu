// edge_bwd.cu -- fused GATv2 edge phase, backward, atomic-free (formulas section 8a-3, restated and
// gradient-checked in oracle/gatv2_oracle.py::gatv2_backward_manual).  With g = dL/dy, delta_i = g_i . (y_i - bias),
// q = keep/(1-p):
//     alpha_ij = exp(e_ij - m_i) / den_i                 (recomputed from xl_j, xr_i and the saved m, den)
//     de_ij    = alpha_ij (q_ij g_i.xl_j - delta_i)
//     d xr_i   = sum_{j -> i} de_ij att * lrelu'(s_ij)                          (lane's role as DESTINATION)
//     d xl_j   = sum_{j -> i} alpha_ij q_ij g_i + de_ij att * lrelu'(s_ij)      (lane's role as SOURCE)
//     d att    = sum_ij de_ij lrelu(s_ij),   d bias = sum_i g_i
// One lane owns one (node, head) and walks BOTH CSR orientations, so each gradient row has exactly one writer:
// no atomics, bit-reproducible.  d att / d bias: per-CTA partials + a fixed-order fp64 second stage.
#include "edge_common.cuh"
#include "reduce.cuh"

namespace tg {

struct EdgeBwdArgs {
    const void *xl, *xr;
    const float *att, *bias, *y, *m, *den, *gy;
    void *dxl, *dxr;
    float *partials;  // (grid, 2*HC): [d att | d bias] per CTA
    const int32_t *rowptr_in, *col_in, *rowptr_out, *col_out, *slot_out, *tile_lo, *tile_hi;
    int32_t N, T, num_tiles, S, H;
    int64_t E;
    float slope, inv_keep;
    uint32_t drop_thr;
    uint64_t seed;
    int32_t literal;
    int32_t win_rows_smem;  // windows up to this many rows are staged in shared memory
    int32_t region_rows;    // rows each shared region is sized for (>= T)
};

template <int C, typename ST, bool SM>
__device__ __forceinline__ void edge_bwd_body(const EdgeBwdArgs &a, unsigned char *smem_raw) {
    const int tid = threadIdx.x;
    const int tile = blockIdx.x % a.num_tiles;
    const int snap = blockIdx.x / a.num_tiles;
    const int H = a.H, HC = H * C;
    const int n0 = tile * a.T;
    const int n1 = min(a.N, n0 + a.T);
    const int nt = n1 - n0;
    const bool self_only = a.literal && snap > 0;
    int lo = a.tile_lo[tile], hi = a.tile_hi[tile];
    if (self_only) { lo = n0; hi = n1; }
    const int win = hi - lo;
    const int64_t row0 = static_cast<int64_t>(snap) * a.N + lo;  // global row of window row 0

    const ST *xl_g = static_cast<const ST *>(a.xl) + row0 * HC;
    const ST *xr_g = static_cast<const ST *>(a.xr) + row0 * HC;
    const float *g_g = a.gy + row0 * HC;
    const float *y_g = a.y + row0 * HC;
    const float *m_g = a.m + row0 * H;
    const float *den_g = a.den + row0 * H;

    // shared layout: [mbarrier 16][xl region][xr region][g region][y region][m | inv_den | delta]
    const uint32_t reg_st = round16(a.region_rows * HC * (uint32_t)sizeof(ST)) + 16;
    const uint32_t reg_f = round16(a.region_rows * HC * (uint32_t)sizeof(float)) + 16;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned char *xl_base = smem_raw + 16;
    unsigned char *xr_base = xl_base + reg_st;
    unsigned char *g_base = xr_base + reg_st;
    unsigned char *y_base = g_base + reg_f;
    float *m_s = reinterpret_cast<float *>(y_base + reg_f);
    float *inv_s = m_s + a.region_rows * H;
    float *delta_s = inv_s + a.region_rows * H;

    const ST *xl_w = xl_g, *xr_w = xr_g;  // window accessors: shared (SM) or global (fallback)
    const float *g_w = g_g;
    if (SM) {
        const uint32_t nb_st = win * HC * (uint32_t)sizeof(ST), nb_f = win * HC * (uint32_t)sizeof(float);
        const CopyPlan c0 = plan_copy(xl_g, xl_base, nb_st);
        const CopyPlan c1 = plan_copy(xr_g, xr_base, nb_st);
        const CopyPlan c2 = plan_copy(g_g, g_base, nb_f);
        const CopyPlan c3 = plan_copy(y_g, y_base, nb_f);
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(bar, c0.mid + c1.mid + c2.mid + c3.mid);
            issue_copy_bulk(c0, bar);
            issue_copy_bulk(c1, bar);
            issue_copy_bulk(c2, bar);
            issue_copy_bulk(c3, bar);
        }
        copy_ragged(c0, tid);
        copy_ragged(c1, tid);
        copy_ragged(c2, tid);
        copy_ragged(c3, tid);
        __syncthreads();
        mbar_wait(bar, 0);
        xl_w = reinterpret_cast<const ST *>(c0.s);
        xr_w = reinterpret_cast<const ST *>(c1.s);
        g_w = reinterpret_cast<const float *>(c2.s);
        const float *y_w = reinterpret_cast<const float *>(c3.s);
        // per (window row, head): m, 1/den, delta = g . (y - bias)
        for (int i = tid; i < win * H; i += blockDim.x) {
            const int r = i / H, hh = i - r * H;
            const float *gp = g_w + r * HC + hh * C, *yp = y_w + r * HC + hh * C;
            float dl = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) dl = fmaf(gp[c], yp[c] - __ldg(a.bias + hh * C + c), dl);
            delta_s[i] = dl;
            m_s[i] = m_g[i];
            inv_s[i] = 1.f / den_g[i];
        }
        __syncthreads();
    }

    constexpr int CP = (C + 1) / 2;
    constexpr float kLog2e = 1.4426950408889634f;
    const int node_l = tid / H;
    const int h = tid - node_l * H;
    const bool active = node_l < nt;
    float2 dxl[CP], dxr[CP], datt[CP];
#pragma unroll
    for (int i = 0; i < CP; ++i) dxl[i] = dxr[i] = datt[i] = make_float2(0.f, 0.f);

    if (active) {
        const int v = n0 + node_l;
        const int vl = v - lo;
        float2 xl_v[CP], xr_v[CP], g_v[CP], att_h[CP];
        load_row<C>(xl_w + vl * HC + h * C, xl_v);
        load_row<C>(xr_w + vl * HC + h * C, xr_v);
        load_row<C>(g_w + vl * HC + h * C, g_v);
        load_row<C>(a.att + h * C, att_h);
        float m_v, inv_v, delta_v;
        if (SM) {
            m_v = m_s[vl * H + h];
            inv_v = inv_s[vl * H + h];
            delta_v = delta_s[vl * H + h];
        } else {
            m_v = m_g[vl * H + h];
            inv_v = 1.f / den_g[vl * H + h];
            float2 yv[CP], bh[CP], d2 = make_float2(0.f, 0.f);
            load_row<C>(y_g + vl * HC + h * C, yv);
            load_row<C>(a.bias + h * C, bh);
#pragma unroll
            for (int i = 0; i < CP; ++i) d2 = __ffma2_rn(g_v[i], __fadd2_rn(yv[i], make_float2(-bh[i].x, -bh[i].y)), d2);
            delta_v = hsum(d2);
        }
        const uint32_t key = a.drop_thr ? dropout_snapshot_key(a.seed, (uint32_t)snap) : 0u;
        const float2 slope2 = make_float2(a.slope, a.slope);
        const float dslope = 1.f - a.slope;  // lrelu'(s) = slope + (1 - slope) * [s > 0]
        const float2 dslope2 = make_float2(dslope, dslope);

        // ---- role 1: v as DESTINATION, in-edges (u -> v): d xr_v, d att ---------------------------------------
        {
            const int k1 = __ldg(a.rowptr_in + v + 1);
            const int k0 = self_only ? k1 - 1 : __ldg(a.rowptr_in + v);
            const ST *base = xl_w + h * C;
            for (int k = k0; k < k1; ++k) {
                const int u = __ldg(a.col_in + k) - lo;
                float2 xu[CP], s[CP], z[CP];
                load_row<C>(base + u * HC, xu);
                const float e = edge_score<C, ST>(att_h, xu, xr_v, slope2, s, z);
                float2 gx2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < CP; ++i) gx2 = __ffma2_rn(g_v[i], xu[i], gx2);
                float q = 1.f;
                if (a.drop_thr) q = dropout_bits16(key, (uint32_t)k, (uint32_t)h) >= a.drop_thr ? a.inv_keep : 0.f;
                const float alpha = fast_exp2(fmaf(e, kLog2e, -m_v)) * inv_v;
                const float de = alpha * fmaf(q, hsum(gx2), -delta_v);
                const float2 de2 = make_float2(de, de);
#pragma unroll
                for (int i = 0; i < CP; ++i) {
                    datt[i] = __ffma2_rn(de2, z[i], datt[i]);
                    const float2 step = make_float2(s[i].x > 0.f ? 1.f : 0.f, s[i].y > 0.f ? 1.f : 0.f);
                    const float2 d = __ffma2_rn(step, dslope2, slope2);
                    dxr[i] = __ffma2_rn(__fmul2_rn(de2, att_h[i]), d, dxr[i]);
                }
            }
        }
        // ---- role 2: v as SOURCE, out-edges (v -> u): d xl_v ---------------------------------------------------
        {
            const int k1 = __ldg(a.rowptr_out + v + 1);
            const int k0 = self_only ? k1 - 1 : __ldg(a.rowptr_out + v);
            const ST *base_r = xr_w + h * C;
            const float *base_g = g_w + h * C;
            for (int k2 = k0; k2 < k1; ++k2) {
                const int u = __ldg(a.col_out + k2) - lo;
                float2 xru[CP], gu[CP], s[CP], z[CP];
                load_row<C>(base_r + u * HC, xru);
                load_row<C>(base_g + u * HC, gu);
                const float e = edge_score<C, ST>(att_h, xl_v, xru, slope2, s, z);
                float2 gx2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < CP; ++i) gx2 = __ffma2_rn(gu[i], xl_v[i], gx2);
                float m_u, inv_u, delta_u;
                if (SM) {
                    m_u = m_s[u * H + h];
                    inv_u = inv_s[u * H + h];
                    delta_u = delta_s[u * H + h];
                } else {
                    m_u = m_g[u * H + h];
                    inv_u = 1.f / den_g[u * H + h];
                    float2 yu[CP], bh[CP], d2 = make_float2(0.f, 0.f);
                    load_row<C>(y_g + u * HC + h * C, yu);
                    load_row<C>(a.bias + h * C, bh);
#pragma unroll
                    for (int i = 0; i < CP; ++i) d2 = __ffma2_rn(gu[i], __fadd2_rn(yu[i], make_float2(-bh[i].x, -bh[i].y)), d2);
                    delta_u = hsum(d2);
                }
                float q = 1.f;
                if (a.drop_thr) {
                    const uint32_t kin = (uint32_t)__ldg(a.slot_out + k2);
                    q = dropout_bits16(key, kin, (uint32_t)h) >= a.drop_thr ? a.inv_keep : 0.f;
                }
                const float alpha = fast_exp2(fmaf(e, kLog2e, -m_u)) * inv_u;
                const float aq = alpha * q;
                const float de = alpha * fmaf(q, hsum(gx2), -delta_u);
                const float2 aq2 = make_float2(aq, aq), de2 = make_float2(de, de);
#pragma unroll
                for (int i = 0; i < CP; ++i) {
                    const float2 step = make_float2(s[i].x > 0.f ? 1.f : 0.f, s[i].y > 0.f ? 1.f : 0.f);
                    const float2 d = __ffma2_rn(step, dslope2, slope2);
                    dxl[i] = __ffma2_rn(__fmul2_rn(de2, att_h[i]), d, __ffma2_rn(aq2, gu[i], dxl[i]));
                }
            }
        }
    }

    // ---- d bias partial straight from the g tile (before any region is recycled) ------------------------
    float dbias_j = 0.f;
    if (tid < HC) {
        const float *gp = g_w + (n0 - lo) * HC + tid;
        for (int r = 0; r < nt; ++r) dbias_j += gp[r * HC];
    }
    __syncthreads();  // every lane is done reading the windows: recycle xl/xr regions as output staging
    ST *dxl_s = reinterpret_cast<ST *>(xl_base);
    ST *dxr_s = reinterpret_cast<ST *>(xr_base);
    float *red = reinterpret_cast<float *>(y_base);  // (C, nt*H) d att scratch
    const int ntl = nt * H;
    if (active) {
        store_row<C>(dxl_s + node_l * HC + h * C, dxl);
        store_row<C>(dxr_s + node_l * HC + h * C, dxr);
#pragma unroll
        for (int c = 0; c < C; ++c) red[c * ntl + tid] = (c & 1) ? datt[c / 2].y : datt[c / 2].x;
    }
    __syncthreads();
    ST *dxl_g = static_cast<ST *>(a.dxl) + (static_cast<int64_t>(snap) * a.N + n0) * HC;
    ST *dxr_g = static_cast<ST *>(a.dxr) + (static_cast<int64_t>(snap) * a.N + n0) * HC;
    for (int i = tid; i < nt * HC; i += blockDim.x) {
        dxl_g[i] = dxl_s[i];
        dxr_g[i] = dxr_s[i];
    }
    if (tid < HC) {
        const int hh = tid / C, c = tid - hh * C;
        const float *rp = red + c * ntl + hh;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // fixed order: deterministic
        int n = 0;
        for (; n + 4 <= nt; n += 4) {
            s0 += rp[(n + 0) * H];
            s1 += rp[(n + 1) * H];
            s2 += rp[(n + 2) * H];
            s3 += rp[(n + 3) * H];
        }
        for (; n < nt; ++n) s0 += rp[n * H];
        float *out = a.partials + static_cast<int64_t>(blockIdx.x) * 2 * HC;
        out[tid] = (s0 + s1) + (s2 + s3);
        out[HC + tid] = dbias_j;
    }
}

template <int C, typename ST>
__global__ void __launch_bounds__(256, 2) edge_bwd_kernel(const EdgeBwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile = blockIdx.x % a.num_tiles;
    const int snap = blockIdx.x / a.num_tiles;
    int win = a.tile_hi[tile] - a.tile_lo[tile];
    if (a.literal && snap > 0) win = min(a.N, (tile + 1) * a.T) - tile * a.T;
    if (win <= a.win_rows_smem)
        edge_bwd_body<C, ST, true>(a, smem_raw);
    else
        edge_bwd_body<C, ST, false>(a, smem_raw);
}

template <int C, typename ST>
static int launch_bwd(const EdgeBwdArgs &a, int threads, int max_win, cudaStream_t st) {
    EdgeBwdArgs b = a;
    const int H = a.H, HC = H * C;
    // bytes per window row across the four regions and the three stat arrays
    const size_t per_row = size_t(HC) * (2 * sizeof(ST) + 2 * sizeof(float)) + size_t(H) * 12;
    const size_t fixed = 16 + 4 * 32 + 64;
    int rows_fit = int((size_t(kSmemBudget) - fixed) / per_row);
    if (rows_fit < a.T) {
        tecgat_set_error("edge_bwd: a tile of %d nodes x %d channels does not fit shared memory", a.T, HC);
        return TECGAT_ENOSUP;
    }
    b.win_rows_smem = rows_fit < max_win ? rows_fit : max_win;
    b.region_rows = b.win_rows_smem > a.T ? b.win_rows_smem : a.T;
    const size_t smem = 16 + 2 * (round16(uint32_t(b.region_rows * HC * sizeof(ST))) + 16) +
                        2 * (round16(uint32_t(b.region_rows * HC * sizeof(float))) + 16) + size_t(b.region_rows) * H * 12 + 16;
    auto kern = edge_bwd_kernel<C, ST>;
    TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = int64_t(a.num_tiles) * a.S;
    kern<<<(unsigned)grid, threads, smem, st>>>(b);
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

}  // namespace tg

extern "C" int64_t tecgat_edge_bwd_workspace(const tecgat_plan_t *plan, int32_t snapshots, int32_t heads,
                                             int32_t out_channels) {
    if (!plan || snapshots <= 0 || heads <= 0 || out_channels <= 0) return 0;
    return int64_t(plan->num_tiles) * snapshots * 2 * heads * out_channels * (int64_t)sizeof(float);
}

extern "C" int tecgat_edge_bwd(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att,
                               const float *bias, const float *y, const float *m, const float *den, const float *gy,
                               void *dxl, void *dxr, float *datt, float *dbias, void *workspace, int32_t snapshots,
                               int32_t heads, int32_t out_channels, float negative_slope, float dropout_p, uint64_t seed,
                               int32_t mode, int32_t dtype, void *stream) {
    using namespace tg;
    TG_REQUIRE(plan && xl && xr && att && bias && y && m && den && gy && dxl && dxr && datt && dbias && workspace,
               TECGAT_EINVAL, "edge_bwd: NULL argument");
    TG_REQUIRE(snapshots > 0 && heads > 0 && out_channels > 0, TECGAT_EINVAL, "edge_bwd: non-positive size");
    TG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, TECGAT_EINVAL, "edge_bwd: dropout_p %f outside [0, 1)", dropout_p);
    TG_REQUIRE(negative_slope >= 0.f && negative_slope <= 1.f, TECGAT_ENOSUP, "edge_bwd: negative_slope %f outside [0, 1]", negative_slope);
    TG_REQUIRE(mode == TECGAT_MODE_SHARED || mode == TECGAT_MODE_LITERAL, TECGAT_EINVAL, "edge_bwd: bad mode %d", mode);
    TG_REQUIRE(dtype == TECGAT_F32 || dtype == TECGAT_BF16, TECGAT_EINVAL, "edge_bwd: bad dtype %d", dtype);
    const int threads = ((plan->tile_nodes * heads + 31) / 32) * 32;
    TG_REQUIRE(threads <= 256, TECGAT_ENOSUP, "edge_bwd: tile_nodes (%d) * heads (%d) exceeds 256 lanes", plan->tile_nodes, heads);
    const int HC = heads * out_channels;
    TG_REQUIRE(HC <= threads, TECGAT_ENOSUP, "edge_bwd: heads*out_channels (%d) exceeds the CTA size (%d)", HC, threads);
    const int64_t grid = int64_t(plan->num_tiles) * snapshots;
    TG_REQUIRE(grid < (int64_t(1) << 31), TECGAT_ENOSUP, "edge_bwd: grid too large");
    EdgeBwdArgs a;
    a.xl = xl; a.xr = xr; a.att = att; a.bias = bias; a.y = y; a.m = m; a.den = den; a.gy = gy;
    a.dxl = dxl; a.dxr = dxr; a.partials = static_cast<float *>(workspace);
    a.rowptr_in = plan->rowptr_in; a.col_in = plan->col_in; a.rowptr_out = plan->rowptr_out; a.col_out = plan->col_out;
    a.slot_out = plan->slot_out; a.tile_lo = plan->tile_lo; a.tile_hi = plan->tile_hi;
    a.N = plan->num_nodes; a.T = plan->tile_nodes; a.num_tiles = plan->num_tiles; a.S = snapshots; a.H = heads;
    a.E = plan->num_edges;
    a.slope = negative_slope;
    a.drop_thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u;
    a.inv_keep = 1.f / (1.f - dropout_p);
    a.seed = seed;
    a.literal = (mode == TECGAT_MODE_LITERAL);
    a.win_rows_smem = 0;
    a.region_rows = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = TECGAT_ENOSUP;
#define TG_CASE(CC)                                                                                  \
    case CC:                                                                                         \
        rc = dtype == TECGAT_F32 ? launch_bwd<CC, float>(a, threads, plan->max_window, st)           \
                                 : launch_bwd<CC, __nv_bfloat16>(a, threads, plan->max_window, st);  \
        break;
    switch (out_channels) {
        TG_FOR_EACH_C(TG_CASE)
        default:
            tecgat_set_error("edge_bwd: out_channels=%d is not among the compiled channel counts", out_channels);
            return TECGAT_ENOSUP;
    }
#undef TG_CASE
    if (rc != TECGAT_OK) return rc;
    ReduceSegs segs = {{datt, dbias, nullptr, nullptr}, {0, HC, 0, 0}, {HC, 2 * HC, 0, 0}};
    return reduce_columns(a.partials, grid, 2 * HC, segs, st);
}
