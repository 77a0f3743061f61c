// permute.cu -- the caller glue either side of the encoder (reference src/model/tec_mollm.py:94,100), one pass each:
//
//   forward :  z[b, n, l, :] = x[b, l, n, :] + y[b, l, n, :]      residual add + (B,L,N,C) -> (B,N,L,C) for the
//                                                                  temporal encoder (reshaped (B*N, L, C) at :106)
//   backward:  g[b, l, n, :] = gz[b, n, l, :]                      the gradient of both the residual branch and y
//
// The input permute of tec_mollm.py:84 needs no kernel at all: snapshots are independent, so the encoder runs on the
// (B, L, N, C) tensor in place, snapshot s = b*L + l instead of l*B + b.
//
// HBM-bound.  A CTA owns a tile (b, TN consecutive nodes, all L steps): on the (B,L,N,C) side that is L contiguous
// chunks of TN*C floats, on the (B,N,L,C) side ONE contiguous run of TN*L*C floats; the tile is transposed through shared
// memory so that both sides move as full-width vector accesses.  Algorithmic bytes per row: 3*C*4 forward, 2*C*4 backward.
#include "common.cuh"

namespace tg {

template <typename V, bool FWD>  // V = float2 (C even) or float
__global__ void __launch_bounds__(768) residual_permute_kernel(const V *__restrict__ a, const V *__restrict__ b2, V *__restrict__ out,
                                                              int L, int N, int CV, int TN, int tiles_per_b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V *tile = reinterpret_cast<V *>(smem_raw);  // [node][l][c]
    const int bidx = blockIdx.x / tiles_per_b, n0 = (blockIdx.x - bidx * tiles_per_b) * TN;
    const int tn = min(TN, N - n0);
    const int chunk = tn * CV;                       // vectors of one (l, node range) chunk
    const int64_t base_bl = (int64_t)bidx * L * N;   // row index of (b, l = 0, node 0) on the (B,L,N,C) side
    const int64_t base_bn = ((int64_t)bidx * N + n0) * L * CV;  // vector index of the tile on the (B,N,L,C) side
    const int total = tn * L * CV;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    // threadIdx.x walks a chunk (its (node, channel) is fixed when the chunk fits blockDim.x, the usual case), threadIdx.y the steps
    if (!FWD) {
        for (int i = tid; i < total; i += nthr) tile[i] = a[base_bn + i];
        __syncthreads();
    }
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) {
        const int nl = i / CV, c = i - nl * CV;
        V *t = tile + (size_t)nl * L * CV + c;
        const int64_t g0 = (base_bl + n0) * CV + i;
#pragma unroll 4
        for (int l = threadIdx.y; l < L; l += blockDim.y) {
            const int64_t g = g0 + (int64_t)l * N * CV;
            if (FWD) {
                V v = a[g];
                if (b2) {
                    const V w = b2[g];
                    if constexpr (sizeof(V) == 8) { v.x += w.x; v.y += w.y; } else { v += w; }
                }
                t[l * CV] = v;
            } else {
                out[g] = t[l * CV];
            }
        }
    }
    if (FWD) {
        __syncthreads();
        for (int i = tid; i < total; i += nthr) out[base_bn + i] = tile[i];
    }
}

template <bool FWD>
static int launch_permute(const float *a, const float *b2, float *out, int B, int L, int N, int C, cudaStream_t st) {
    TG_REQUIRE(a && out, TECGAT_EINVAL, "residual_permute: NULL argument");
    TG_REQUIRE(B > 0 && L > 0 && N > 0 && C > 0, TECGAT_EINVAL, "residual_permute: non-positive size");
    const size_t row_bytes = size_t(L) * C * 4;
    TG_REQUIRE(row_bytes <= 96 * 1024, TECGAT_ENOSUP, "residual_permute: L*C = %d floats does not fit a shared-memory tile", L * C);
    int TN = (int)std::min<size_t>(32, std::max<size_t>(1, (72 * 1024) / row_bytes));
    if (TN >= 8) TN &= ~7;  // chunk starts stay sector aligned
    TN = std::min(TN, N);
    const int tiles_per_b = (N + TN - 1) / TN;
    const int64_t grid = int64_t(B) * tiles_per_b;
    TG_REQUIRE(grid < (int64_t(1) << 31), TECGAT_ENOSUP, "residual_permute: too many tiles");
    const size_t smem = size_t(TN) * row_bytes;
    const bool vec = (C % 2) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b2) | reinterpret_cast<uintptr_t>(out)) & 7) == 0;
    const int cv = vec ? C / 2 : C;
    const int bx = std::min(256, (TN * cv + 31) / 32 * 32), by = std::max(1, std::min(L, 768 / bx));
    const dim3 block(bx, by);
    if (vec) {
        auto kern = residual_permute_kernel<float2, FWD>;
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)smem));
        kern<<<(unsigned)grid, block, smem, st>>>(reinterpret_cast<const float2 *>(a), reinterpret_cast<const float2 *>(b2),
                                                reinterpret_cast<float2 *>(out), L, N, C / 2, TN, tiles_per_b); tg_count_launch();
    } else {
        auto kern = residual_permute_kernel<float, FWD>;
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)smem));
        kern<<<(unsigned)grid, block, smem, st>>>(a, b2, out, L, N, C, TN, tiles_per_b); tg_count_launch();
    }
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

}  // namespace tg

extern "C" int tecgat_residual_permute_fwd(const float *x_dev, const float *y_dev, float *z_dev, int32_t batch, int32_t steps,
                                           int32_t nodes, int32_t channels, void *stream) {
    return tg::launch_permute<true>(x_dev, y_dev, z_dev, batch, steps, nodes, channels, static_cast<cudaStream_t>(stream));
}

extern "C" int tecgat_residual_permute_bwd(const float *gz_dev, float *g_dev, int32_t batch, int32_t steps, int32_t nodes,
                                           int32_t channels, void *stream) {
    return tg::launch_permute<false>(gz_dev, nullptr, g_dev, batch, steps, nodes, channels, static_cast<cudaStream_t>(stream));
}
