// project.cu -- C ABI of the lin_l / lin_r projections: argument checks + implementation choice.
#include <cstdlib>

#include "project.cuh"

extern "C" int tecgat_project_fwd(const float *x, const float *wl, const float *bl, const float *wr, const float *br,
                                  void *xl, void *xr, int64_t rows, int32_t F, int32_t hc, int32_t dtype, int32_t impl,
                                  void *stream) {
    TG_REQUIRE(x && wl && bl && wr && br && xl && xr, TECGAT_EINVAL, "project_fwd: NULL argument");
    TG_REQUIRE(rows > 0 && F > 0 && hc > 0, TECGAT_EINVAL, "project_fwd: non-positive size");
    TG_REQUIRE(dtype == TECGAT_F32 || dtype == TECGAT_BF16, TECGAT_EINVAL, "project_fwd: bad dtype %d", dtype);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (impl == TECGAT_PROJ_FFMA) return tg::project_fwd_ffma(x, wl, bl, wr, br, xl, xr, rows, F, hc, dtype, st);
    TG_REQUIRE(impl == TECGAT_PROJ_TC, TECGAT_EINVAL, "project_fwd: bad impl %d", impl);
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xl) | reinterpret_cast<uintptr_t>(xr)) & 15) == 0;
    if (!aligned || !tg::project_tc_supported(F, hc))  // shapes / alignments outside the tensor-core kernel's range
        return tg::project_fwd_ffma(x, wl, bl, wr, br, xl, xr, rows, F, hc, dtype, st);
    return tg::project_fwd_tc(x, wl, bl, wr, br, xl, xr, rows, F, hc, dtype, st);
}

bool tg::project_bwd_use_tc(int F, int HC, int dtype, const void *dxl, const void *dxr, const void *x, const void *dx) {
    const char *env = tg_env("TECGAT_PROJ_BWD");
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dxl) | reinterpret_cast<uintptr_t>(dxr) |
                           reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    if (!aligned || !tg::project_bwd_tc_supported(F, HC, dtype)) return false;
    if (env && env[0] == 't') return true;
    if (env && env[0] == 'r') return false;
    return dtype == TECGAT_BF16 || !tg::project_bwd_rt_supported(F, HC, dxl, dxr, x, dx);
}

extern "C" int64_t tecgat_project_bwd_workspace(int64_t rows, int32_t F, int32_t hc, int32_t impl) {
    if (rows <= 0 || F <= 0 || hc <= 0) return 0;
    if (impl == TECGAT_PROJ_FFMA) return tg::project_bwd_ffma_workspace(rows, F, hc);
    const int64_t a = tg::project_bwd_tc_workspace(rows, F, hc), b = tg::project_bwd_rt_workspace(rows, F, hc);
    return a > b ? a : b;
}

extern "C" int tecgat_project_bwd(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr,
                                  float *dx, float *dwl, float *dbl, float *dwr, float *dbr, void *workspace, int64_t rows,
                                  int32_t F, int32_t hc, int32_t dtype, int32_t impl, void *stream) {
    TG_REQUIRE(dxl && dxr && x && wl && wr && dwl && dbl && dwr && dbr && workspace, TECGAT_EINVAL, "project_bwd: NULL argument");
    TG_REQUIRE(rows > 0 && F > 0 && hc > 0, TECGAT_EINVAL, "project_bwd: non-positive size");
    TG_REQUIRE(dtype == TECGAT_F32 || dtype == TECGAT_BF16, TECGAT_EINVAL, "project_bwd: bad dtype %d", dtype);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (impl == TECGAT_PROJ_FFMA)
        return tg::project_bwd_ffma(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, workspace, rows, F, hc, dtype, st);
    TG_REQUIRE(impl == TECGAT_PROJ_TC, TECGAT_EINVAL, "project_bwd: bad impl %d", impl);
    // Product path (project.cuh): tcgen05 for the bf16 contract, the register-tiled packed-fp32 kernel for the fp32 contract.
    if (!tg::project_bwd_use_tc(F, hc, dtype, dxl, dxr, x, dx) && tg::project_bwd_rt_supported(F, hc, dxl, dxr, x, dx))
        return tg::project_bwd_rt(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, workspace, rows, F, hc, dtype, st);
    return tg::project_bwd_tc(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, workspace, rows, F, hc, dtype, st);
}

// dx += dxl Wl + dxr Wr: the caller pre-loaded dx (the spatial block's residual-branch gradient, permute.cu), so autograd's
// separate accumulation pass disappears.  Only the register-tiled kernel implements it (a bulk-TMA reduction store).
extern "C" int tecgat_project_bwd_acc_supported(int32_t F, int32_t hc) {
    return tg::project_bwd_rt_supported(F, hc, nullptr, nullptr, nullptr, nullptr) ? 1 : 0;
}

extern "C" int tecgat_project_bwd_acc(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr,
                                      float *dx, float *dwl, float *dbl, float *dwr, float *dbr, void *workspace,
                                      int64_t rows, int32_t F, int32_t hc, int32_t dtype, void *stream) {
    TG_REQUIRE(dxl && dxr && x && wl && wr && dx && dwl && dbl && dwr && dbr && workspace, TECGAT_EINVAL, "project_bwd_acc: NULL argument");
    TG_REQUIRE(rows > 0 && F > 0 && hc > 0, TECGAT_EINVAL, "project_bwd_acc: non-positive size");
    TG_REQUIRE(dtype == TECGAT_F32 || dtype == TECGAT_BF16, TECGAT_EINVAL, "project_bwd_acc: bad dtype %d", dtype);
    TG_REQUIRE(tg::project_bwd_rt_supported(F, hc, dxl, dxr, x, dx), TECGAT_ENOSUP,
               "project_bwd_acc: F=%d, H*C=%d (or a pointer that is not 16-byte aligned) is outside the accumulating kernel's range; "
               "call tecgat_project_bwd and add", F, hc);
    if (tg::project_bwd_use_tc(F, hc, dtype, dxl, dxr, x, dx))
        return tg::project_bwd_tc(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, workspace, rows, F, hc, dtype, static_cast<cudaStream_t>(stream), true);
    return tg::project_bwd_rt(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, workspace, rows, F, hc, dtype,
                              static_cast<cudaStream_t>(stream), true);
}
