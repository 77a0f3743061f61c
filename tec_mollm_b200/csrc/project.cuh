// project.cuh -- internal launchers of the two projection implementations (see project.cu for the C ABI).
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace tg {
// CUDA-core cross-check path (project_ffma.cu)
int project_fwd_ffma(const float *x, const float *wl, const float *bl, const float *wr, const float *br, void *xl, void *xr,
                     int64_t R, int F, int HC, int dtype, cudaStream_t st);
int64_t project_bwd_ffma_workspace(int64_t R, int F, int HC);
int project_bwd_ffma(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx,
                     float *dwl, float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype,
                     cudaStream_t st);
// tcgen05 + bulk-TMA product path (project_tc.cu)
bool project_tc_supported(int F, int HC);
int project_fwd_tc(const float *x, const float *wl, const float *bl, const float *wr, const float *br, void *xl, void *xr,
                   int64_t R, int F, int HC, int dtype, cudaStream_t st);
int64_t project_bwd_tc_workspace(int64_t R, int F, int HC);
bool project_bwd_tc_supported(int F, int HC, int dtype);
int project_bwd_tc(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx,
                   float *dwl, float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype,
                   cudaStream_t st, bool accumulate = false, ReduceJob *defer = nullptr);
// Which backward projection runs on the product path: bf16 contract -> tcgen05 (single bf16 product, two CTAs per SM: the gradient
// rows already are bf16 and x is rounded to bf16 exactly as autocast does for the reference's Linear backward); fp32 contract ->
// the register-tiled packed-fp32 kernel (the three-term bf16 split costs the tensor-core version 6 MMAs and 3 re-layouts per
// tile: measured 4.4 ms against 1.63 ms).  TECGAT_PROJ_BWD=tc / rt force one.
bool project_bwd_use_tc(int F, int HC, int dtype, const void *dxl, const void *dxr, const void *x, const void *dx);
// register-tiled FFMA2 backward (project_bwd_rt.cu): the product path of the backward projection
bool project_bwd_rt_supported(int F, int HC, const void *dxl, const void *dxr, const void *x, const void *dx);
int64_t project_bwd_rt_workspace(int64_t R, int F, int HC);
int project_bwd_rt(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx, float *dwl,
                   float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype, cudaStream_t st,
                   bool accumulate = false, ReduceJob *defer = nullptr);
}  // namespace tg
