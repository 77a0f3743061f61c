// temporal.cu -- the normalisation / activation / concat / stride part of the TemporalEncoder's Multi_Scale_Conv_Block
// (/root/reference/src/model/modules.py:13-60), fused.  Per block the reference runs, for each of the three
// branches k = 3 / 5 / 7:  Conv1d -> GroupNorm(1, C) -> GELU, then torch.cat over channels, then a 1x1 Conv1d with stride s.
// The convolutions are dense contractions and stay library calls (the host side issues ONE 7-tap convolution with the three
// kernels zero-padded and stacked, which lands in the concatenated layout directly).  Everything between them is this file:
//     z[n, b*C + c, t'] = gelu( gamma[b,c] * (y[n, b*C + c, s t'] - mean[n,b]) * rstd[n,b] + beta[b,c] )
// i.e. GroupNorm statistics over the (C, L) block of branch b of sample n, affine, exact (erf) GELU, and ONLY the positions the
// strided 1x1 convolution reads are written (the odd positions matter for the statistics alone): one pass over y instead of the
// reference's GroupNorm + GELU + cat + strided read (4 round trips of the (n, 3C, L) tensor).
// Backward: d y from d z through GELU', the affine map and the GroupNorm Jacobian, with d gamma / d beta as per-CTA partials
// finished in a fixed order (reduce.cu): no atomics.
#include <algorithm>

#include "common.cuh"

namespace tg {

constexpr int kGnThreads = 256;
constexpr int kGnMaxElems = 8192;  // C * L of one (sample, branch) block kept in shared memory (fp32)

__device__ __forceinline__ float gn_ld(const float *p, int i) { return p[i]; }
__device__ __forceinline__ float gn_ld(const __nv_bfloat16 *p, int i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void gn_st(float *p, int i, float v) { p[i] = v; }
__device__ __forceinline__ void gn_st(__nv_bfloat16 *p, int i, float v) { p[i] = __float2bfloat16_rn(v); }

__device__ __forceinline__ float block_sum(float v, float *red) {  // fixed-order tree: bit-reproducible
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kGnThreads / 32; ++i) t += red[i];
    return t;
}

__device__ __forceinline__ float gelu_f(float a) { return 0.5f * a * (1.f + erff(a * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float a) {
    return 0.5f * (1.f + erff(a * 0.70710678118654752440f)) + a * 0.39894228040143267794f * __expf(-0.5f * a * a);
}

// gamma / beta of the block's branch = block % branches
struct GnParams {
    const float *gamma, *beta;  // (branches, C)
    int branches;
    uint32_t magic_L, magic_Lo;  // floor(2^32 / L) + 1: i / L == __umulhi(i, magic_L) for i * L < 2^32 (here i < 8192)
};
__device__ __forceinline__ int fdiv(int i, uint32_t magic) { return (int)__umulhi((uint32_t)i, magic); }

template <typename TY, typename TZ>
__global__ void __launch_bounds__(kGnThreads) gn_gelu_fwd_kernel(const TY *__restrict__ y, GnParams prm, TZ *__restrict__ z,
                                                                 float *__restrict__ mean_out, float *__restrict__ rstd_out, int C, int L,
                                                                 int stride, float eps) {
    extern __shared__ float sh[];
    float *red = sh + C * L;
    const int blk = blockIdx.x, branch = blk % prm.branches;
    const float *gamma = prm.gamma + (size_t)branch * C, *beta = prm.beta + (size_t)branch * C;
    const int M = C * L;
    const TY *src = y + (size_t)blk * M;
    float s = 0.f;
    for (int i = threadIdx.x; i < M; i += kGnThreads) {
        const float v = gn_ld(src, i);
        sh[i] = v;
        s += v;
    }
    const float mean = block_sum(s, red) / (float)M;
    float q = 0.f;
    for (int i = threadIdx.x; i < M; i += kGnThreads) {
        const float d = sh[i] - mean;
        q += d * d;
    }
    const float rstd = rsqrtf(block_sum(q, red) / (float)M + eps);
    if (threadIdx.x == 0) {
        mean_out[blk] = mean;
        rstd_out[blk] = rstd;
    }
    const int Lo = (L + stride - 1) / stride;
    TZ *dst = z + (size_t)blk * C * Lo;
    for (int i = threadIdx.x; i < C * Lo; i += kGnThreads) {
        const int c = fdiv(i, prm.magic_Lo), t = i - c * Lo;
        const float a = fmaf((sh[c * L + t * stride] - mean) * rstd, gamma[c], beta[c]);
        gn_st(dst, i, gelu_f(a));
    }
}

// d y[c, t] = rstd * (dyh - mean(dyh) - yh * mean(dyh * yh)),  dyh = da * gamma[c],  da = dz * gelu'(a) at the strided positions, 0 else.
// Persistent: CTA g owns branch g % branches and the samples g / branches, + gridDim.x / branches, ..; its d gamma / d beta sums
// stay in registers (thread c = channel c) across those samples: one partial row per CTA.
template <typename TY, typename TZ>
__global__ void __launch_bounds__(kGnThreads) gn_gelu_bwd_kernel(const TY *__restrict__ y, GnParams prm, const float *__restrict__ mean_in,
                                                                 const float *__restrict__ rstd_in, const TZ *__restrict__ dz,
                                                                 TY *__restrict__ dy, float *__restrict__ partials /* (gridDim.x, 2C) */,
                                                                 int64_t samples, int C, int L, int stride) {
    extern __shared__ float sh[];  // yh [C*L], da [C*L], red
    const int M = C * L;
    float *yh = sh, *da = sh + M, *red = sh + 2 * M;
    const int branch = blockIdx.x % prm.branches;
    const float *gamma = prm.gamma + (size_t)branch * C, *beta = prm.beta + (size_t)branch * C;
    const int Lo = (L + stride - 1) / stride;
    float dg = 0.f, db = 0.f;
    for (int64_t n = blockIdx.x / prm.branches; n < samples; n += gridDim.x / prm.branches) {
        const int64_t blk = n * prm.branches + branch;
        const float mean = mean_in[blk], rstd = rstd_in[blk];
        const TY *src = y + (size_t)blk * M;
        const TZ *dsrc = dz + (size_t)blk * C * Lo;
        float s1 = 0.f, s2 = 0.f;
        __syncthreads();  // the previous sample's per-channel pass is done with yh / da
        for (int i = threadIdx.x; i < M; i += kGnThreads) {
            const int c = fdiv(i, prm.magic_L), t = i - c * L;
            const float h = (gn_ld(src, i) - mean) * rstd;
            float d = 0.f;
            const int tq = stride == 2 ? (t >> 1) : (stride == 1 ? t : t / stride);
            if (tq * stride == t) d = gn_ld(dsrc, c * Lo + tq) * gelu_grad(fmaf(h, gamma[c], beta[c]));
            yh[i] = h;
            da[i] = d;
            s1 += d * gamma[c];
            s2 += d * gamma[c] * h;
        }
        const float m1 = block_sum(s1, red) / (float)M;
        const float m2 = block_sum(s2, red) / (float)M;
        TY *dst = dy + (size_t)blk * M;
        for (int i = threadIdx.x; i < M; i += kGnThreads) {
            const int c = fdiv(i, prm.magic_L);
            gn_st(dst, i, rstd * (da[i] * gamma[c] - m1 - yh[i] * m2));
        }
        for (int c = threadIdx.x; c < C; c += kGnThreads)  // C <= kGnThreads: thread c owns channel c
            for (int t = 0; t < L; t += stride) {
                const float d = da[c * L + t];
                dg += d * yh[c * L + t];
                db += d;
            }
    }
    if ((int)threadIdx.x < C) {
        partials[(size_t)blockIdx.x * 2 * C + threadIdx.x] = dg;
        partials[(size_t)blockIdx.x * 2 * C + C + threadIdx.x] = db;
    }
}

// second stage: d gamma[b, c] = sum over the CTAs of branch b, fp64, ascending
__global__ void __launch_bounds__(256) gn_param_reduce_kernel(const float *__restrict__ partials, int ctas, int branches, int C,
                                                              float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= branches * 2 * C) return;
    const int b = i / (2 * C), j = i - b * 2 * C;
    double v = 0.0;
    for (int g = b; g < ctas; g += branches) v += (double)partials[(size_t)g * 2 * C + j];
    if (j < C) dgamma[b * C + j] = (float)v;
    else dbeta[b * C + (j - C)] = (float)v;
}

static GnParams gn_params(const float *gamma, const float *beta, int branches, int L, int stride) {
    const int Lo = (L + stride - 1) / stride;
    return GnParams{gamma, beta, branches, (uint32_t)(4294967296ull / (uint64_t)L) + 1u, (uint32_t)(4294967296ull / (uint64_t)Lo) + 1u};
}
static int gn_bwd_grid(int64_t samples, int branches) {
    const int64_t per_branch = std::min<int64_t>(samples, std::max(1, 8 * tg_sm_count() / branches));
    return (int)(per_branch * branches);
}

template <typename TY, typename TZ>
static int launch_gn_fwd(const void *y, const float *gamma, const float *beta, void *z, float *mean, float *rstd, int64_t samples, int branches,
                         int C, int L, int stride, float eps, cudaStream_t st) {
    const size_t smem = (size_t)(C * L + 8) * 4;
    auto kern = gn_gelu_fwd_kernel<TY, TZ>;
    TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)smem));
    kern<<<(unsigned)(samples * branches), kGnThreads, smem, st>>>(static_cast<const TY *>(y), gn_params(gamma, beta, branches, L, stride),
                                                                  static_cast<TZ *>(z), mean, rstd, C, L, stride, eps);
    tg_count_launch();
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

template <typename TY, typename TZ>
static int launch_gn_bwd(const void *y, const float *gamma, const float *beta, const float *mean, const float *rstd, const void *dz, void *dy,
                         float *partials, float *dgamma, float *dbeta, int64_t samples, int branches, int C, int L, int stride, cudaStream_t st) {
    const size_t smem = (size_t)(2 * C * L + 8) * 4;
    auto kern = gn_gelu_bwd_kernel<TY, TZ>;
    TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)smem));
    const int grid = gn_bwd_grid(samples, branches);
    kern<<<grid, kGnThreads, smem, st>>>(static_cast<const TY *>(y), gn_params(gamma, beta, branches, L, stride), mean, rstd,
                                         static_cast<const TZ *>(dz), static_cast<TY *>(dy), partials, samples, C, L, stride);
    tg_count_launch();
    gn_param_reduce_kernel<<<(branches * 2 * C + 255) / 256, 256, 0, st>>>(partials, grid, branches, C, dgamma, dbeta);
    tg_count_launch();
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

}  // namespace tg

#define TG_GN_CHECK(name)                                                                                                       \
    TG_REQUIRE(samples > 0 && branches > 0 && channels > 0 && length > 0 && stride > 0, TECGAT_EINVAL, name ": non-positive size"); \
    TG_REQUIRE((int64_t)channels * length <= tg::kGnMaxElems, TECGAT_ENOSUP, name ": channels * length = %lld exceeds %d",         \
               (long long)channels * length, tg::kGnMaxElems);                                                                  \
    TG_REQUIRE(samples * branches < (int64_t(1) << 31), TECGAT_ENOSUP, name ": too many (sample, branch) blocks");                 \
    TG_REQUIRE((y_dtype == TECGAT_F32 || y_dtype == TECGAT_BF16) && (z_dtype == TECGAT_F32 || z_dtype == TECGAT_BF16), TECGAT_EINVAL, name ": bad dtype")

extern "C" int tecgat_gn_gelu_fwd(const void *y_dev, const float *gamma_dev, const float *beta_dev, void *z_dev, float *mean_dev,
                                  float *rstd_dev, int64_t samples, int32_t branches, int32_t channels, int32_t length, int32_t stride,
                                  float eps, int32_t y_dtype, int32_t z_dtype, void *stream) {
    using namespace tg;
    TG_REQUIRE(y_dev && gamma_dev && beta_dev && z_dev && mean_dev && rstd_dev, TECGAT_EINVAL, "gn_gelu_fwd: NULL argument");
    TG_GN_CHECK("gn_gelu_fwd");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y_dtype == TECGAT_F32 && z_dtype == TECGAT_F32)
        return launch_gn_fwd<float, float>(y_dev, gamma_dev, beta_dev, z_dev, mean_dev, rstd_dev, samples, branches, channels, length, stride, eps, st);
    if (y_dtype == TECGAT_BF16 && z_dtype == TECGAT_F32)
        return launch_gn_fwd<__nv_bfloat16, float>(y_dev, gamma_dev, beta_dev, z_dev, mean_dev, rstd_dev, samples, branches, channels, length, stride, eps, st);
    if (y_dtype == TECGAT_BF16 && z_dtype == TECGAT_BF16)
        return launch_gn_fwd<__nv_bfloat16, __nv_bfloat16>(y_dev, gamma_dev, beta_dev, z_dev, mean_dev, rstd_dev, samples, branches, channels, length, stride, eps, st);
    return launch_gn_fwd<float, __nv_bfloat16>(y_dev, gamma_dev, beta_dev, z_dev, mean_dev, rstd_dev, samples, branches, channels, length, stride, eps, st);
}

extern "C" int64_t tecgat_gn_gelu_bwd_workspace(int64_t samples, int32_t branches, int32_t channels) {
    if (samples <= 0 || branches <= 0 || channels <= 0) return 0;
    return int64_t(tg::gn_bwd_grid(samples, branches)) * 2 * channels * (int64_t)sizeof(float);
}

extern "C" int tecgat_gn_gelu_bwd(const void *y_dev, const float *gamma_dev, const float *beta_dev, const float *mean_dev,
                                  const float *rstd_dev, const void *dz_dev, void *dy_dev, float *dgamma_dev, float *dbeta_dev,
                                  void *workspace_dev, int64_t samples, int32_t branches, int32_t channels, int32_t length, int32_t stride,
                                  int32_t y_dtype, int32_t z_dtype, void *stream) {
    using namespace tg;
    TG_REQUIRE(y_dev && gamma_dev && beta_dev && mean_dev && rstd_dev && dz_dev && dy_dev && dgamma_dev && dbeta_dev && workspace_dev,
               TECGAT_EINVAL, "gn_gelu_bwd: NULL argument");
    TG_GN_CHECK("gn_gelu_bwd");
    TG_REQUIRE(channels <= tg::kGnThreads, TECGAT_ENOSUP, "gn_gelu_bwd: more than %d channels per branch", tg::kGnThreads);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *part = static_cast<float *>(workspace_dev);
    if (y_dtype == TECGAT_F32 && z_dtype == TECGAT_F32)
        return launch_gn_bwd<float, float>(y_dev, gamma_dev, beta_dev, mean_dev, rstd_dev, dz_dev, dy_dev, part, dgamma_dev, dbeta_dev, samples, branches, channels, length, stride, st);
    if (y_dtype == TECGAT_BF16 && z_dtype == TECGAT_F32)
        return launch_gn_bwd<__nv_bfloat16, float>(y_dev, gamma_dev, beta_dev, mean_dev, rstd_dev, dz_dev, dy_dev, part, dgamma_dev, dbeta_dev, samples, branches, channels, length, stride, st);
    if (y_dtype == TECGAT_BF16 && z_dtype == TECGAT_BF16)
        return launch_gn_bwd<__nv_bfloat16, __nv_bfloat16>(y_dev, gamma_dev, beta_dev, mean_dev, rstd_dev, dz_dev, dy_dev, part, dgamma_dev, dbeta_dev, samples, branches, channels, length, stride, st);
    return launch_gn_bwd<float, __nv_bfloat16>(y_dev, gamma_dev, beta_dev, mean_dev, rstd_dev, dz_dev, dy_dev, part, dgamma_dev, dbeta_dev, samples, branches, channels, length, stride, st);
}
