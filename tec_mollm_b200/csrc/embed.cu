// embed.cu -- the producer side of the spatial block: SpatioTemporalEmbedding.forward
// (/root/reference/src/model/modules.py:230-264) fused into one pass, so that only the RAW (B, L, N, C_raw) features and the
// per-snapshot time indices cross the host link (train.py:58-65 copies x and the (B, L, 4) time features, then expands them
// with stride 0 over the nodes):
//     out[s, n, 0:Cr]      = x[s, n, :]
//     out[s, n, Cr:Cr+De]  = node_emb[n] + (((tod[i0] + doy[i1]) + year[i2]) + season[i3])        (s = b*L + l, i* = tf[s, :])
// in the reference's own association order (tod + doy + year + season, then node + temporal: modules.py:259-260), i.e.
// bit-identical to torch.  The reference gathers and adds five (B, L, N, De) tensors and concatenates (7 ATen launches,
// ~6 full-tensor round trips).
//
// Backward: the gradient of the embedded tensor, ge (S, N, Cr+De), is reduced over snapshots (d node_emb) and over nodes
// (per-snapshot temporal gradient, then per table row) without atomics: warp-granular partial sums in a fixed order, fp64
// second stages -> bit-reproducible (torch's embedding backward uses atomics).
#include "common.cuh"

namespace tg {

constexpr int kEmbThreads = 256;

__global__ void __launch_bounds__(kEmbThreads) embed_fwd_kernel(const float *__restrict__ x, const int32_t *__restrict__ tf,
                                                                const float *__restrict__ node, const float *__restrict__ tod,
                                                                const float *__restrict__ doy, const float *__restrict__ year,
                                                                const float *__restrict__ season, float *__restrict__ out, int N,
                                                                int Cr, int De, int n_tod, int n_doy, int n_year, int n_season,
                                                                int nodes_per_block) {
    __shared__ float T[64];
    const int s = blockIdx.y;
    const int n0 = blockIdx.x * nodes_per_block;
    const int nn = min(nodes_per_block, N - n0);
    if ((int)threadIdx.x < De) {
        const int c = threadIdx.x;
        auto clampi = [](int v, int hi) { return v < 0 ? 0 : (v >= hi ? hi - 1 : v); };
        const int i0 = clampi(tf[s * 4 + 0], n_tod), i1 = clampi(tf[s * 4 + 1], n_doy), i2 = clampi(tf[s * 4 + 2], n_year),
                  i3 = clampi(tf[s * 4 + 3], n_season);
        // modules.py:259: temporal_emb = tod_emb + doy_emb + year_emb + season_emb (left to right)
        T[c] = __fadd_rn(__fadd_rn(__fadd_rn(tod[i0 * De + c], doy[i1 * De + c]), year[i2 * De + c]), season[i3 * De + c]);
    }
    __syncthreads();
    const int F = Cr + De;
    const int64_t row0 = (int64_t)s * N + n0;
    const float *xs = x + row0 * Cr;
    float *os = out + row0 * F;
    for (int i = threadIdx.x; i < nn * F; i += kEmbThreads) {
        const int n = i / F, c = i - n * F;
        float v;
        if (c < Cr) v = xs[n * Cr + c];
        else v = __fadd_rn(node[(int64_t)(n0 + n) * De + (c - Cr)], T[c - Cr]);  // modules.py:260: node_emb + temporal_emb
        os[i] = v;
    }
}

// Same pass for even channel counts known at compile time (the reference's 6 raw + 16 embedding channels): one float2 per
// thread and iteration, the row / channel split is a division by a constant.  Bit-identical to the kernel above (same adds).
template <int CR2, int DE2>
__global__ void __launch_bounds__(kEmbThreads) embed_fwd_pairs_kernel(const float2 *__restrict__ x, const int32_t *__restrict__ tf,
                                                                      const float2 *__restrict__ node, const float *__restrict__ tod,
                                                                      const float *__restrict__ doy, const float *__restrict__ year,
                                                                      const float *__restrict__ season, float2 *__restrict__ out, int N,
                                                                      int n_tod, int n_doy, int n_year, int n_season, int nodes_per_block) {
    constexpr int F2 = CR2 + DE2, De = 2 * DE2;
    __shared__ float2 T2[DE2];
    const int s = blockIdx.y;
    const int n0 = blockIdx.x * nodes_per_block;
    const int nn = min(nodes_per_block, N - n0);
    if ((int)threadIdx.x < De) {
        const int c = threadIdx.x;
        auto clampi = [](int v, int hi) { return v < 0 ? 0 : (v >= hi ? hi - 1 : v); };
        const int i0 = clampi(tf[s * 4 + 0], n_tod), i1 = clampi(tf[s * 4 + 1], n_doy), i2 = clampi(tf[s * 4 + 2], n_year),
                  i3 = clampi(tf[s * 4 + 3], n_season);
        reinterpret_cast<float *>(T2)[c] =
            __fadd_rn(__fadd_rn(__fadd_rn(tod[i0 * De + c], doy[i1 * De + c]), year[i2 * De + c]), season[i3 * De + c]);
    }
    __syncthreads();
    const int64_t row0 = (int64_t)s * N + n0;
    const float2 *xs = x + row0 * CR2;
    const float2 *ns = node + (int64_t)n0 * DE2;
    float2 *os = out + row0 * F2;
    for (int i = threadIdx.x; i < nn * F2; i += kEmbThreads) {
        const int n = i / F2, c = i - n * F2;
        float2 v;
        if (c < CR2) {
            v = __ldcs(xs + n * CR2 + c);  // read once
        } else {
            const float2 e = __ldg(ns + n * DE2 + (c - CR2)), t = T2[c - CR2];
            v = make_float2(__fadd_rn(e.x, t.x), __fadd_rn(e.y, t.y));
        }
        os[i] = v;
    }
}

// ---- backward ------------------------------------------------------------------------------------------------------
// One warp owns (a tile of 128 nodes, a group of snapshots).  Lane = (row r = lane / 8, channel pair k = lane % 8): one load
// instruction covers four rows' contiguous 64-byte embedding halves.  For De != 16 the generic scalar path below is used.
constexpr int kTileNodes = 128;

__global__ void __launch_bounds__(128) embed_bwd_partial_kernel(const float *__restrict__ ge, float *__restrict__ pnode /* (G, N, 16) */,
                                                                float *__restrict__ ptime /* (S, tiles, 16) */, int S, int N, int Cr,
                                                                int tiles, int G) {
    const int warp_global = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp_global >= tiles * G) return;
    const int tile = warp_global % tiles, g = warp_global / tiles;
    const int s0 = (int)((int64_t)S * g / G), s1 = (int)((int64_t)S * (g + 1) / G);
    const int r = lane >> 3, k = lane & 7;
    const int F = Cr + 16;
    const int n0 = tile * kTileNodes;
    float2 accn[32];  // node sums: nodes n0 + 4 j + r, j = 0..31, channel pair k
#pragma unroll
    for (int j = 0; j < 32; ++j) accn[j] = make_float2(0.f, 0.f);
    for (int s = s0; s < s1; ++s) {
        const float *base = ge + ((int64_t)s * N + n0) * F + Cr + 2 * k;
        float2 t = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int n = 4 * j + r;
            float2 v = make_float2(0.f, 0.f);
            if (n0 + n < N) v = *reinterpret_cast<const float2 *>(base + (int64_t)n * F);
            accn[j].x += v.x;
            accn[j].y += v.y;
            t.x += v.x;
            t.y += v.y;
        }
        t.x += __shfl_xor_sync(0xFFFFFFFFu, t.x, 8);
        t.y += __shfl_xor_sync(0xFFFFFFFFu, t.y, 8);
        t.x += __shfl_xor_sync(0xFFFFFFFFu, t.x, 16);
        t.y += __shfl_xor_sync(0xFFFFFFFFu, t.y, 16);
        if (r == 0) *reinterpret_cast<float2 *>(ptime + ((int64_t)s * tiles + tile) * 16 + 2 * k) = t;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n = n0 + 4 * j + r;
        if (n < N) *reinterpret_cast<float2 *>(pnode + ((int64_t)g * N + n) * 16 + 2 * k) = accn[j];
    }
}

// out[j] (+)= sum_p part[p * width + j] in fp64, p ascending: thread per column (coalesced across threads)
__global__ void __launch_bounds__(256) embed_reduce_rows_kernel(const float *__restrict__ part, int64_t P, int64_t width,
                                                                float *__restrict__ out, int accumulate) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= width) return;
    double v = 0.0;
    for (int64_t p = 0; p < P; ++p) v += (double)part[p * width + j];
    out[j] = accumulate ? (float)((double)out[j] + v) : (float)v;
}

// per-snapshot temporal gradient gT[s, c] = sum_tiles ptime[s, tile, c]  (fixed order, fp64)
__global__ void __launch_bounds__(256) embed_time_rows_kernel(const float *__restrict__ ptime, float *__restrict__ gT, int64_t S,
                                                              int tiles, int De) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= S * De) return;
    const int64_t s = i / De;
    const int c = (int)(i - s * De);
    double v = 0.0;
    for (int t = 0; t < tiles; ++t) v += (double)ptime[(s * tiles + t) * De + c];
    gT[i] = (float)v;
}

// table gradient: d tab[i, c] (+)= sum_{s : tf[s, which] == i} gT[s, c]; one CTA per table row, fixed-order tree over s
struct TableJob {
    float *out[4];
    int rows[4];
};
__global__ void __launch_bounds__(256) embed_table_grad_kernel(const float *__restrict__ gT, const int32_t *__restrict__ tf, int S,
                                                               int De, TableJob job, int accumulate) {
    __shared__ double sh[256];
    int row = blockIdx.x, which = 0;
    while (which < 3 && row >= job.rows[which]) row -= job.rows[which++];
    const int hi = job.rows[which];
    for (int c = 0; c < De; ++c) {
        double v = 0.0;
        for (int s = threadIdx.x; s < S; s += 256) {
            int i = tf[s * 4 + which];
            i = i < 0 ? 0 : (i >= hi ? hi - 1 : i);
            if (i == row) v += (double)gT[(int64_t)s * De + c];
        }
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            float *o = job.out[which] + row * De + c;
            *o = accumulate ? (float)((double)*o + sh[0]) : (float)sh[0];
        }
        __syncthreads();
    }
}

static int embed_groups(int S) { return S < 64 ? S : 64; }

}  // namespace tg

extern "C" int tecgat_embed_fwd(const float *x_dev, const int32_t *tf_dev, const float *node_dev, const float *tod_dev,
                                const float *doy_dev, const float *year_dev, const float *season_dev, float *out_dev,
                                int32_t snapshots, int32_t nodes, int32_t raw_channels, int32_t emb_dim, int32_t n_tod,
                                int32_t n_doy, int32_t n_year, int32_t n_season, void *stream) {
    using namespace tg;
    TG_REQUIRE(x_dev && tf_dev && node_dev && tod_dev && doy_dev && year_dev && season_dev && out_dev, TECGAT_EINVAL, "embed_fwd: NULL argument");
    TG_REQUIRE(snapshots > 0 && nodes > 0 && raw_channels > 0 && emb_dim > 0 && emb_dim <= 64, TECGAT_EINVAL,
               "embed_fwd: bad size (emb_dim must be 1..64)");
    TG_REQUIRE(n_tod > 0 && n_doy > 0 && n_year > 0 && n_season > 0, TECGAT_EINVAL, "embed_fwd: empty table");
    TG_REQUIRE(snapshots <= 65535, TECGAT_ENOSUP, "embed_fwd: more than 65535 snapshots per call");
    const char *knob = tg_env("TECGAT_EMBED_NPB");  // tuning knob: nodes per block
    const int npb = knob ? std::max(32, atoi(knob)) : 384;  // measured at B = 128: 128 -> 0.56 ms, 384 -> 0.49 ms (1.5 GB of stores: write-bound)
    dim3 grid((nodes + npb - 1) / npb, snapshots);
    const bool aligned8 = ((reinterpret_cast<uintptr_t>(x_dev) | reinterpret_cast<uintptr_t>(node_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 7) == 0;
    if (raw_channels == 6 && emb_dim == 16 && aligned8)  // the reference's shape (modules.py:211-266 with d_emb = 16)
        embed_fwd_pairs_kernel<3, 8><<<grid, kEmbThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float2 *>(x_dev), tf_dev, reinterpret_cast<const float2 *>(node_dev), tod_dev, doy_dev, year_dev,
            season_dev, reinterpret_cast<float2 *>(out_dev), nodes, n_tod, n_doy, n_year, n_season, npb);
    else
        embed_fwd_kernel<<<grid, kEmbThreads, 0, static_cast<cudaStream_t>(stream)>>>(x_dev, tf_dev, node_dev, tod_dev, doy_dev, year_dev, season_dev,
                                                                                  out_dev, nodes, raw_channels, emb_dim, n_tod, n_doy,
                                                                                  n_year, n_season, npb);
    tg_count_launch();
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

extern "C" int64_t tecgat_embed_bwd_workspace(int32_t snapshots, int32_t nodes, int32_t emb_dim) {
    if (snapshots <= 0 || nodes <= 0 || emb_dim != 16) return 0;
    const int64_t tiles = (nodes + tg::kTileNodes - 1) / tg::kTileNodes, G = tg::embed_groups(snapshots);
    return (G * nodes * 16 + int64_t(snapshots) * tiles * 16 + int64_t(snapshots) * 16) * (int64_t)sizeof(float);
}

extern "C" int tecgat_embed_bwd(const float *ge_dev, const int32_t *tf_dev, float *dnode_dev, float *dtod_dev, float *ddoy_dev,
                                float *dyear_dev, float *dseason_dev, void *workspace_dev, int32_t snapshots, int32_t nodes,
                                int32_t raw_channels, int32_t emb_dim, int32_t n_tod, int32_t n_doy, int32_t n_year,
                                int32_t n_season, int32_t accumulate, void *stream) {
    using namespace tg;
    TG_REQUIRE(ge_dev && tf_dev && dnode_dev && dtod_dev && ddoy_dev && dyear_dev && dseason_dev && workspace_dev, TECGAT_EINVAL,
               "embed_bwd: NULL argument");
    TG_REQUIRE(emb_dim == 16 && (raw_channels % 2) == 0, TECGAT_ENOSUP,
               "embed_bwd: emb_dim %d / raw_channels %d unsupported (the reference's d_emb = 16, even raw channel count)", emb_dim, raw_channels);
    TG_REQUIRE(reinterpret_cast<uintptr_t>(ge_dev) % 8 == 0, TECGAT_EINVAL, "embed_bwd: gradient must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = snapshots, N = nodes, tiles = (N + kTileNodes - 1) / kTileNodes, G = embed_groups(S);
    float *pnode = static_cast<float *>(workspace_dev);
    float *ptime = pnode + int64_t(G) * N * 16;
    float *gT = ptime + int64_t(S) * tiles * 16;
    const int warps = tiles * G;
    embed_bwd_partial_kernel<<<(warps + 3) / 4, 128, 0, st>>>(ge_dev, pnode, ptime, S, N, raw_channels, tiles, G);
    tg_count_launch();
    const int64_t wn = int64_t(N) * 16;
    embed_reduce_rows_kernel<<<(unsigned)((wn + 255) / 256), 256, 0, st>>>(pnode, G, wn, dnode_dev, accumulate);
    tg_count_launch();
    embed_time_rows_kernel<<<(unsigned)((int64_t(S) * 16 + 255) / 256), 256, 0, st>>>(ptime, gT, S, tiles, 16);
    tg_count_launch();
    TableJob job = {{dtod_dev, ddoy_dev, dyear_dev, dseason_dev}, {n_tod, n_doy, n_year, n_season}};
    embed_table_grad_kernel<<<n_tod + n_doy + n_year + n_season, 256, 0, st>>>(gT, tf_dev, S, 16, job, accumulate);
    tg_count_launch();
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}
