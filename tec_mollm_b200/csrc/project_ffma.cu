// project_ffma.cu -- CUDA-core (FFMA) implementation of the lin_l / lin_r projections and their backward.
// It exists to cross-check the tcgen05 tensor-core path (project_tc.cu) in the parity tests and to serve shapes
// the tensor-core kernel does not cover; same semantics, same layouts.
//   fwd : [xl | xr] = x [Wl; Wr]^T + [bl | br]            x (R, F) fp32 -> xl, xr (R, HC) storage dtype
//   bwd : dx = dxl Wl + dxr Wr;  dWl = dxl^T x, dbl = sum dxl (same for r): per-CTA partials + fixed-order stage 2
// Under the bf16 contract operands are rounded to bf16 first and accumulated in fp32 (autocast Linear).
#include "common.cuh"
#include "project.cuh"
#include "reduce.cuh"

namespace tg {

constexpr int kRows = 128;  // rows per CTA tile == threads per CTA

template <typename ST>
__device__ __forceinline__ float opnd(float v) { return v; }
template <>
__device__ __forceinline__ float opnd<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// shared: W_s[2HC][F] | b_s[2HC] | x_s[kRows][FP] | o_s[kRows][2HC+1]
template <typename ST>
__global__ void __launch_bounds__(kRows) project_fwd_ffma_kernel(const float *__restrict__ x, const float *__restrict__ wl,
                                                                 const float *__restrict__ bl, const float *__restrict__ wr,
                                                                 const float *__restrict__ br, ST *__restrict__ xl,
                                                                 ST *__restrict__ xr, int64_t R, int F, int HC) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int O = 2 * HC, FP = F | 1, OP = O | 1;
    float *W_s = reinterpret_cast<float *>(smem_raw);
    float *b_s = W_s + O * F;
    float *x_s = b_s + O;
    float *o_s = x_s + kRows * FP;
    const int tid = threadIdx.x;
    for (int i = tid; i < HC * F; i += kRows) {
        W_s[i] = opnd<ST>(wl[i]);
        W_s[HC * F + i] = opnd<ST>(wr[i]);
    }
    for (int i = tid; i < HC; i += kRows) {
        b_s[i] = opnd<ST>(bl[i]);
        b_s[HC + i] = opnd<ST>(br[i]);
    }
    const int64_t num_tiles = (R + kRows - 1) / kRows;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t r0 = tile * kRows;
        const int nr = (int)((R - r0) < (int64_t)kRows ? (R - r0) : (int64_t)kRows);
        __syncthreads();
        for (int i = tid; i < nr * F; i += kRows) x_s[(i / F) * FP + (i % F)] = opnd<ST>(x[r0 * F + i]);
        __syncthreads();
        if (tid < nr) {
            const float *xp = x_s + tid * FP;
            for (int o = 0; o < O; ++o) {
                float acc = 0.f;
                const float *wp = W_s + o * F;
                for (int k = 0; k < F; ++k) acc = fmaf(xp[k], wp[k], acc);
                o_s[tid * OP + o] = acc + b_s[o];
            }
        }
        __syncthreads();
        for (int i = tid; i < nr * HC; i += kRows) {
            const int r = i / HC, c = i - r * HC;
            st_elem(xl + r0 * HC + i, o_s[r * OP + c]);
            st_elem(xr + r0 * HC + i, o_s[r * OP + HC + c]);
        }
    }
}

// bwd.  shared: W_s[2HC][F] | d_s[kRows][2HC+1] | x_s[kRows][FP] | dx_s[kRows][FP]
// per-thread parameter-gradient accumulators: outputs q = tid + i*kRows over [dW (2HC*F) | db (2HC)], i < kMaxAcc
constexpr int kMaxAcc = 48;  // 6144 outputs: e.g. F = 44 with H*C = 44, or F = 64 with H*C = 44
template <typename ST>
__global__ void __launch_bounds__(kRows) project_bwd_ffma_kernel(const ST *__restrict__ dxl, const ST *__restrict__ dxr,
                                                                 const float *__restrict__ x, const float *__restrict__ wl,
                                                                 const float *__restrict__ wr, float *__restrict__ dx,
                                                                 float *__restrict__ partials, int64_t R, int F, int HC) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int O = 2 * HC, FP = F | 1, OP = O | 1;
    float *W_s = reinterpret_cast<float *>(smem_raw);
    float *d_s = W_s + O * F;
    float *x_s = d_s + kRows * OP;
    float *dx_s = x_s + kRows * FP;
    const int tid = threadIdx.x;
    for (int i = tid; i < HC * F; i += kRows) {
        W_s[i] = opnd<ST>(wl[i]);
        W_s[HC * F + i] = opnd<ST>(wr[i]);
    }
    const int nq = O * F + O;
    float acc[kMaxAcc];
#pragma unroll
    for (int i = 0; i < kMaxAcc; ++i) acc[i] = 0.f;
    const int64_t num_tiles = (R + kRows - 1) / kRows;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t r0 = tile * kRows;
        const int nr = (int)((R - r0) < (int64_t)kRows ? (R - r0) : (int64_t)kRows);
        __syncthreads();
        for (int i = tid; i < nr * F; i += kRows) x_s[(i / F) * FP + (i % F)] = opnd<ST>(x[r0 * F + i]);
        for (int i = tid; i < nr * HC; i += kRows) {
            const int r = i / HC, c = i - r * HC;
            d_s[r * OP + c] = ld_elem(dxl + r0 * HC + i);
            d_s[r * OP + HC + c] = ld_elem(dxr + r0 * HC + i);
        }
        __syncthreads();
        if (dx != nullptr && tid < nr) {
            const float *dp = d_s + tid * OP;
            for (int k = 0; k < F; ++k) {
                float a = 0.f;
                for (int o = 0; o < O; ++o) a = fmaf(dp[o], W_s[o * F + k], a);
                dx_s[tid * FP + k] = a;
            }
        }
#pragma unroll
        for (int i = 0; i < kMaxAcc; ++i) {
            const int q = tid + i * kRows;
            if (q < nq) {
                float a = acc[i];
                if (q < O * F) {
                    const int o = q / F, k = q - o * F;
                    for (int r = 0; r < nr; ++r) a = fmaf(d_s[r * OP + o], x_s[r * FP + k], a);
                } else {
                    const int o = q - O * F;
                    for (int r = 0; r < nr; ++r) a += d_s[r * OP + o];
                }
                acc[i] = a;
            }
        }
        __syncthreads();
        if (dx != nullptr)
            for (int i = tid; i < nr * F; i += kRows) dx[r0 * F + i] = dx_s[(i / F) * FP + (i % F)];
    }
#pragma unroll
    for (int i = 0; i < kMaxAcc; ++i) {
        const int q = tid + i * kRows;
        if (q < nq) partials[static_cast<int64_t>(blockIdx.x) * nq + q] = acc[i];
    }
}

static int ffma_grid(int64_t R) {
    const int64_t tiles = (R + kRows - 1) / kRows;
    return (int)(tiles < 148 * 4 ? tiles : 148 * 4);
}

int project_fwd_ffma(const float *x, const float *wl, const float *bl, const float *wr, const float *br, void *xl, void *xr,
                     int64_t R, int F, int HC, int dtype, cudaStream_t st) {
    const int O = 2 * HC;
    const size_t smem = sizeof(float) * (size_t(O) * F + O + size_t(kRows) * (F | 1) + size_t(kRows) * (O | 1));
    TG_REQUIRE(smem <= size_t(kSmemBudget), TECGAT_ENOSUP, "project_fwd(ffma): F=%d, HC=%d needs %zu B shared memory", F, HC, smem);
    const int grid = ffma_grid(R);
    if (dtype == TECGAT_F32) {
        auto k = project_fwd_ffma_kernel<float>;
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(k), (int)smem));
        k<<<grid, kRows, smem, st>>>(x, wl, bl, wr, br, static_cast<float *>(xl), static_cast<float *>(xr), R, F, HC); tg_count_launch();
    } else {
        auto k = project_fwd_ffma_kernel<__nv_bfloat16>;
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(k), (int)smem));
        k<<<grid, kRows, smem, st>>>(x, wl, bl, wr, br, static_cast<__nv_bfloat16 *>(xl), static_cast<__nv_bfloat16 *>(xr), R, F, HC); tg_count_launch();
    }
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

int64_t project_bwd_ffma_workspace(int64_t R, int F, int HC) {
    return int64_t(ffma_grid(R)) * (2 * HC * F + 2 * HC) * (int64_t)sizeof(float);
}

int project_bwd_ffma(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx,
                     float *dwl, float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype,
                     cudaStream_t st) {
    const int O = 2 * HC;
    const int nq = O * F + O;
    TG_REQUIRE(nq <= kMaxAcc * kRows, TECGAT_ENOSUP, "project_bwd(ffma): 2*HC*(F+1) = %d exceeds %d", nq, kMaxAcc * kRows);
    const size_t smem = sizeof(float) * (size_t(O) * F + size_t(kRows) * (O | 1) + 2 * size_t(kRows) * (F | 1));
    TG_REQUIRE(smem <= size_t(kSmemBudget), TECGAT_ENOSUP, "project_bwd(ffma): F=%d, HC=%d needs %zu B shared memory", F, HC, smem);
    const int grid = ffma_grid(R);
    float *partials = static_cast<float *>(workspace);
    if (dtype == TECGAT_F32) {
        auto k = project_bwd_ffma_kernel<float>;
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(k), (int)smem));
        k<<<grid, kRows, smem, st>>>(static_cast<const float *>(dxl), static_cast<const float *>(dxr), x, wl, wr, dx, partials, R, F, HC); tg_count_launch();
    } else {
        auto k = project_bwd_ffma_kernel<__nv_bfloat16>;
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(k), (int)smem));
        k<<<grid, kRows, smem, st>>>(static_cast<const __nv_bfloat16 *>(dxl), static_cast<const __nv_bfloat16 *>(dxr), x, wl, wr, dx, partials, R, F, HC); tg_count_launch();
    }
    TG_LAUNCH_CHECK();
    ReduceSegs segs = {{dwl, dwr, dbl, dbr}, {0, HC * F, O * F, O * F + HC}, {HC * F, O * F, O * F + HC, O * F + O}};
    return reduce_columns(partials, grid, nq, segs, st);
}

}  // namespace tg
