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
int project_bwd_tc(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx,
                   float *dwl, float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype,
                   cudaStream_t st);
// register-tiled FFMA2 backward (project_bwd_rt.cu): the product path of the backward projection
bool project_bwd_rt_supported(int F, int HC, const void *dxl, const void *dxr, const void *x, const void *dx);
int64_t project_bwd_rt_workspace(int64_t R, int F, int HC);
int project_bwd_rt(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx, float *dwl,
                   float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype, cudaStream_t st,
                   bool accumulate = false, ReduceJob *defer = nullptr);
}  // namespace tg
