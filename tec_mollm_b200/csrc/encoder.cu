// encoder.cu -- one C entry per direction of the GATv2 layer (what torch_geometric.nn.GATv2Conv.forward and its autograd
// backward are for the reference, /root/reference/src/model/modules.py:356): forward = projection + fused edge kernel,
// backward = fused edge backward + projection backward + ONE fixed-order finish of all six parameter gradients, optionally
// accumulated in place onto the caller's .grad storage.  Also here: the device-resident dropout seed (CUDA-graph replays draw
// fresh masks) and the optional per-phase CUDA-event timing the benchmark reads.
#include <mutex>
#include <vector>

#include "edge_common.cuh"
#include "project.cuh"
#include "reduce.cuh"

namespace tg {

// ---- per-phase timing: library-owned events recorded on the caller's stream between the phases of the fused entries -------
struct PhaseTimer {
    std::mutex mu;
    bool enabled = false;
    std::vector<cudaEvent_t> pool;          // recycled events
    std::vector<std::pair<int, cudaEvent_t>> marks;  // (phase id, event) in record order; id -1 = start of a direction
    double ms[4] = {0, 0, 0, 0};
};
static PhaseTimer g_timer;

static void phase_mark(int id, cudaStream_t st) {
    if (!g_timer.enabled) return;
    std::lock_guard<std::mutex> lk(g_timer.mu);
    cudaEvent_t ev;
    if (!g_timer.pool.empty()) {
        ev = g_timer.pool.back();
        g_timer.pool.pop_back();
    } else if (cudaEventCreate(&ev) != cudaSuccess) {
        return;
    }
    cudaEventRecord(ev, st);
    g_timer.marks.push_back({id, ev});
}

__global__ void seed_advance_kernel(uint64_t *state, uint64_t *seed_out) {
    // splitmix64 of (seed, counter): one fresh 64-bit seed per training forward, drawn ON THE DEVICE
    uint64_t z = state[0] + 0x9E3779B97F4A7C15ull * (state[1] + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    *seed_out = z ^ (z >> 31);
    state[1] += 1ull;
}

}  // namespace tg

extern "C" int tecgat_seed_advance(uint64_t *state_dev, uint64_t *seed_out_dev, void *stream) {
    TG_REQUIRE(state_dev && seed_out_dev, TECGAT_EINVAL, "seed_advance: NULL argument");
    tg::seed_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(state_dev, seed_out_dev);
    tg_count_launch();
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

extern "C" int tecgat_phase_timing(int32_t enable) {
    std::lock_guard<std::mutex> lk(tg::g_timer.mu);
    tg::g_timer.enabled = enable != 0;
    return TECGAT_OK;
}

// Synchronises the recorded events, adds the elapsed time of every phase (0 proj_fwd, 1 edge_fwd, 2 edge_bwd, 3 proj_bwd incl.
// the gradient finish) since the last call to ms_out4 and forgets the events.
extern "C" int tecgat_phase_times(double *ms_out4) {
    TG_REQUIRE(ms_out4, TECGAT_EINVAL, "phase_times: NULL argument");
    std::lock_guard<std::mutex> lk(tg::g_timer.mu);
    auto &m = tg::g_timer.marks;
    for (size_t i = 0; i < m.size(); ++i) {
        if (m[i].first >= 0 && i > 0) {
            TG_CUDA(cudaEventSynchronize(m[i].second));
            float ms = 0.f;
            TG_CUDA(cudaEventElapsedTime(&ms, m[i - 1].second, m[i].second));
            tg::g_timer.ms[m[i].first] += ms;
        }
    }
    for (auto &e : m) tg::g_timer.pool.push_back(e.second);
    m.clear();
    for (int i = 0; i < 4; ++i) {
        ms_out4[i] = tg::g_timer.ms[i];
        tg::g_timer.ms[i] = 0.0;
    }
    return TECGAT_OK;
}

extern "C" int tecgat_forward(const tecgat_plan_t *plan, const float *x, const float *wl, const float *bl, const float *wr,
                              const float *br, const float *att, const float *bias, void *xl, void *xr, float *y, float *stat,
                              int32_t snapshots, int32_t in_channels, int32_t heads, int32_t out_channels, float negative_slope,
                              float dropout_p, uint64_t seed, const uint64_t *seed_dev, int32_t mode, int32_t dtype, int32_t impl,
                              void *stream) {
    return tecgat_forward_into(plan, x, wl, bl, wr, br, att, bias, xl, xr, y, stat, snapshots, in_channels, heads, out_channels,
                               negative_slope, dropout_p, seed, seed_dev, mode, dtype, impl, nullptr, 0, stream);
}

extern "C" int tecgat_forward_into(const tecgat_plan_t *plan, const float *x, const float *wl, const float *bl, const float *wr,
                                   const float *br, const float *att, const float *bias, void *xl, void *xr, float *y, float *stat,
                                   int32_t snapshots, int32_t in_channels, int32_t heads, int32_t out_channels,
                                   float negative_slope, float dropout_p, uint64_t seed, const uint64_t *seed_dev, int32_t mode,
                                   int32_t dtype, int32_t impl, float *y_wide, int64_t ld_wide, void *stream) {
    TG_REQUIRE(plan, TECGAT_EINVAL, "forward: NULL plan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t rows = int64_t(snapshots) * plan->num_nodes;
    tg::phase_mark(-1, st);
    int rc = tecgat_project_fwd(x, wl, bl, wr, br, xl, xr, rows, in_channels, heads * out_channels, dtype, impl, stream);
    if (rc != TECGAT_OK) return rc;
    tg::phase_mark(0, st);
    rc = tg::edge_fwd_run(plan, xl, xr, att, bias, y, stat, snapshots, heads, out_channels, negative_slope, dropout_p, seed, seed_dev,
                          mode, dtype, stream, y_wide, ld_wide);
    tg::phase_mark(1, st);
    return rc;
}

static int64_t align256(int64_t b) { return (b + 255) & ~int64_t(255); }

extern "C" int64_t tecgat_backward_workspace(const tecgat_plan_t *plan, int32_t snapshots, int32_t in_channels, int32_t heads,
                                             int32_t out_channels, int32_t impl) {
    if (!plan || snapshots <= 0) return 0;
    const int64_t rows = int64_t(snapshots) * plan->num_nodes;
    return align256(tecgat_edge_bwd_workspace(plan, snapshots, heads, out_channels)) +
           align256(tecgat_project_bwd_workspace(rows, in_channels, heads * out_channels, impl));
}

// 1: tecgat_backward finishes all six parameter gradients with one launch and can accumulate them onto existing storage
extern "C" int tecgat_backward_fused_supported(int32_t in_channels, int32_t hc, int32_t impl) {
    if (impl != TECGAT_PROJ_TC) return 0;
    return tecgat_project_bwd_acc_supported(in_channels, hc);
}

extern "C" int tecgat_backward(const tecgat_plan_t *plan, const float *x, const float *wl, const float *wr, const float *att,
                               const float *bias, const void *xl, const void *xr, const float *y, const float *stat,
                               const float *gy, void *dxl, void *dxr, float *dx, int32_t dx_accumulate, float *dwl, float *dbl,
                               float *dwr, float *dbr, float *datt, float *dbias, int32_t grad_accumulate, void *workspace,
                               int32_t snapshots, int32_t in_channels, int32_t heads, int32_t out_channels, float negative_slope,
                               float dropout_p, uint64_t seed, const uint64_t *seed_dev, int32_t mode, int32_t dtype,
                               int32_t impl, void *stream) {
    using namespace tg;
    TG_REQUIRE(plan && dwl && dbl && dwr && dbr && datt && dbias && workspace, TECGAT_EINVAL, "backward: NULL argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int F = in_channels, HC = heads * out_channels;
    const int64_t rows = int64_t(snapshots) * plan->num_nodes;
    unsigned char *ws1 = static_cast<unsigned char *>(workspace);
    unsigned char *ws2 = ws1 + align256(tecgat_edge_bwd_workspace(plan, snapshots, heads, out_channels));
    const bool fused = impl == TECGAT_PROJ_TC && project_bwd_rt_supported(F, HC, dxl, dxr, x, dx);  // (implies the tc kernel's range too)
    TG_REQUIRE(fused || (!grad_accumulate && !dx_accumulate), TECGAT_ENOSUP,
               "backward: in-place accumulation needs the register-tiled projection backward (F=%d, H*C=%d, 16-byte aligned "
               "buffers); call without accumulation", F, HC);
    phase_mark(-1, st);
    int64_t edge_rows = 0;
    int rc = edge_bwd_run(plan, xl, xr, att, bias, y, stat, gy, dxl, dxr, datt, dbias, ws1, snapshots, heads, out_channels,
                          negative_slope, dropout_p, seed, seed_dev, mode, dtype, stream, /*reduce=*/!fused, &edge_rows);
    if (rc != TECGAT_OK) return rc;
    phase_mark(2, st);
    if (!fused) {
        rc = tecgat_project_bwd(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, ws2, rows, F, HC, dtype, impl, stream);
        phase_mark(3, st);
        return rc;
    }
    ReduceJob jobs[2];
    jobs[0] = ReduceJob{reinterpret_cast<const float *>(ws1), edge_rows, 2 * HC, {{datt, dbias, nullptr, nullptr}, {0, HC, 0, 0}, {HC, 2 * HC, 0, 0}}};
    if (project_bwd_use_tc(F, HC, dtype, dxl, dxr, x, dx))
        rc = project_bwd_tc(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, ws2, rows, F, HC, dtype, st, dx_accumulate != 0, &jobs[1]);
    else
        rc = project_bwd_rt(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, ws2, rows, F, HC, dtype, st, dx_accumulate != 0, &jobs[1]);
    if (rc != TECGAT_OK) return rc;
    rc = reduce_columns_multi(jobs, 2, grad_accumulate, st);
    phase_mark(3, st);
    return rc;
}
