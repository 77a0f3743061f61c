// project_bwd_rt.cu -- backward of the lin_l / lin_r projections on the packed fp32 pipe, register tiled:
//     dx = dxl Wl + dxr Wr            (rows x 2HC) . (2HC x F)
//     dWl = dxl^T x, dbl = sum dxl    (same for r)                     one pass over dxl, dxr, x
// Why not tensor cores here: both GEMMs have K or N = 22..44; the fp32 contract (1e-5) needs every operand split into
// 3xTF32 / 3xbf16 terms AND re-laid into the UMMA core-matrix layout by CUDA cores, and the weight-gradient GEMM reduces
// over ROWS, i.e. wants MN-major operands that tcgen05 only takes for 16-bit types.  That re-layout costs more issue
// slots than the 2,000 FMAs per row it would off-load (measured: 4.4 ms for the tcgen05 version of this kernel at B=128,
// against a 0.96 ms HBM floor).  The forward projection, whose only large operand is x, stays on tcgen05 (project_tc.cu).
//
// Persistent CTA = 16 warps: one bulk-TMA producer warp streaming 128-row tiles of (dxl, dxr, x) through a 5-stage
// mbarrier ring, and three TEAMS of five consumer warps; team t owns tiles t, t+3, ..  Inside a team
//   * 2 "dx" warps: thread = 4 rows x 12 outputs; per pair of gradient columns 4 x LDS.64 (rows) + 6 x LDS.128 (weights,
//     broadcast) feed 48 FFMA2; the tile leaves through shared memory as one bulk-TMA store;
//   * 3 "dW" warps: thread = (every 8th row, 8 gradient columns, 12 input columns): 10 x LDS.64 feed 48 FFMA2 per row;
//     its 96 fp32 accumulators live in registers for the whole kernel (column 23 of the padded x tile is a constant 1:
//     the bias gradients come out of the same accumulators).
// End of kernel: the 24 (team, slice) partials are summed in fp64 in a fixed order -> one partial row per CTA, then the
// fixed-order fp64 second stage (reduce.cu).  Deterministic, no atomics.
#include <type_traits>

#include "common.cuh"
#include "project.cuh"
#include "reduce.cuh"

namespace tg {

namespace rt {
constexpr int kRows = 128;   // rows per tile
constexpr int kFP = 24;      // padded input channels (F <= 22: column 23 carries the constant 1)
#ifndef TG_PBWD_DXW
#define TG_PBWD_DXW 2
#endif
// team = kDxWarps "dx" warps (each 24 / kDxWarps outputs of all 128 rows) + 3 "dW" warps.  Per tile a dx thread issues
// 4 rows x (kDxOut / 2) x HC FFMA2 and a dW thread 16 rows x 48: with 2 dx warps the dx side is 1.4x the dW side and the dW
// warps wait on the ring's full barrier (18 % of the samples).  Measured alternative (-DTG_PBWD_DXW=3: two teams of 3 dx
// warps x 8 outputs + 3 dW warps, balanced 704 vs 768 FFMA2, 12 consumer warps): 1.67 ms against 1.63 ms for the default
// three teams of 2 + 3 (15 consumer warps) at B = 128 -- the extra warps hide more latency than the balance recovers.
constexpr int kDxWarps = TG_PBWD_DXW, kDwWarps = 3;
constexpr int kTeams = kDxWarps == 2 ? 3 : 2, kTeamWarps = kDxWarps + kDwWarps;
constexpr int kDxOut = 24 / kDxWarps, kDxPairs = kDxOut / 2, kDxVec = kDxOut / 4;
constexpr int kConsumers = kTeams * kTeamWarps * 32;
constexpr int kThreads = kConsumers + 32;
constexpr int kPad = 128;    // zeroed bytes after every staged tile (threads read up to 2 elements past a row)
// HP = padded gradient columns per array: 24 (HC <= 24: 8 row slices, 5 ring stages) or 48 (HC <= 48: 4 slices, 3 stages);
// either way the 3 dW warps of a team are 96 threads = slices x (2 HP / 8 column tiles) x 2 input tiles
template <int HP> struct Cfg {
    static constexpr int kSlices = 192 / HP;
    static constexpr int kStages = HP == 24 ? 5 : 3;
    static constexpr int kOT = 2 * HP / 8;  // 8-column tiles over [dxl | dxr]
};

struct Args {
    const void *dxl, *dxr;
    const float *x, *wl, *wr;
    float *dx;        // may be NULL
    int32_t accumulate;  // dx += (the caller pre-loaded dx, e.g. with a residual branch's gradient) instead of dx =
    float *partials;  // (grid, 2*HC*F + 2*HC): [dWl | dWr | dbl | dbr] per CTA
    int64_t R;
    int32_t F, HC;
};

struct Smem {
    uint32_t bars, w, stage0, stage_bytes, off_dr, off_x, dxst, dxst_bytes, total;
};
template <typename ST, int HP>
__host__ __device__ inline Smem layout(int F, int HC) {
    constexpr int kStages = Cfg<HP>::kStages, kHP = HP;
    Smem s;
    uint32_t o = 0;
    s.bars = o; o += 128;
    s.w = o; o += 2 * kHP * kFP * 4;
    o = (o + 127) & ~127u;
    const uint32_t tile_d = ((kRows * HC * (uint32_t)sizeof(ST) + kPad + 127) / 128) * 128;
    const uint32_t tile_x = ((kRows * F * 4u + kPad + 127) / 128) * 128;
    s.off_dr = tile_d;
    s.off_x = 2 * tile_d;
    s.stage_bytes = 2 * tile_d + tile_x;
    s.stage0 = o;
    {   // the ring doubles as the end-of-kernel scratch: one fp32 partial block [2*HP][kFP] per (team, slice)
        const uint32_t ring = kStages * s.stage_bytes, scratch = kTeams * Cfg<HP>::kSlices * 2u * kHP * kFP * 4u;
        o += ring > scratch ? ring : scratch;
    }
    s.dxst_bytes = ((kRows * F * 4u + 127) / 128) * 128;
    s.dxst = o; o += kTeams * s.dxst_bytes;
    s.total = o;
    return s;
}

__device__ __forceinline__ float2 ld_pair(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ float2 ld_pair(const __nv_bfloat16 *p) {
    const uint32_t u = *reinterpret_cast<const uint32_t *>(p);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

template <typename ST, int HP>
__global__ void __launch_bounds__(kThreads, 1) project_bwd_rt_kernel(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kHP = HP, kSlices = Cfg<HP>::kSlices, kStages = Cfg<HP>::kStages, kOT = Cfg<HP>::kOT;
    const int F = a.F, HC = a.HC;
    const Smem L = layout<ST, HP>(F, HC);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bars), *empty = full + kStages;
    float *W_s = reinterpret_cast<float *>(smem + L.w);  // [2][kHP][kFP], zero padded
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (a.R + kRows - 1) / kRows;
    const int n_local = blockIdx.x < num_tiles ? (int)((num_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
    const bool need_dx = a.dx != nullptr;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTeamWarps);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < 2 * kHP * kFP; i += kThreads) {
        const int arr = i / (kHP * kFP), c = (i / kFP) % kHP, f = i % kFP;
        float v = 0.f;
        if (c < HC && f < F) {
            v = (arr ? a.wr : a.wl)[c * F + f];
            if (std::is_same<ST, __nv_bfloat16>::value) v = __bfloat162float(__float2bfloat16_rn(v));  // autocast rounds W
        }
        W_s[i] = v;
    }
    for (int i = tid; i < kStages * (int)(L.stage_bytes / 4); i += kThreads) reinterpret_cast<uint32_t *>(smem + L.stage0)[i] = 0u;
    fence_proxy_async();
    __syncthreads();

    if (warp == kTeams * kTeamWarps) {
        // ================================ producer warp ================================
        if (lane == 0) {
            for (int it = 0; it < n_local; ++it) {
                const int s = it % kStages;
                if (it >= kStages) mbar_wait(&empty[s], ((it / kStages) - 1) & 1);
                const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x, r0 = tile * kRows;
                const int nr = (int)((a.R - r0) < (int64_t)kRows ? (a.R - r0) : (int64_t)kRows);
                unsigned char *st = smem + L.stage0 + (size_t)s * L.stage_bytes;
                const uint32_t bd = (uint32_t)nr * HC * (uint32_t)sizeof(ST), bx = (uint32_t)nr * F * 4u;
                const unsigned char *srcs[3] = {reinterpret_cast<const unsigned char *>(static_cast<const ST *>(a.dxl) + r0 * HC),
                                                reinterpret_cast<const unsigned char *>(static_cast<const ST *>(a.dxr) + r0 * HC),
                                                reinterpret_cast<const unsigned char *>(a.x + r0 * F)};
                unsigned char *dsts[3] = {st, st + L.off_dr, st + L.off_x};
                const uint32_t lens[3] = {bd, bd, bx};
                uint32_t tx = 0;
                for (int q = 0; q < 3; ++q) {  // <16-byte ragged ends (last tile only): plain 2-byte copies
                    const uint32_t mid = lens[q] & ~15u;
                    for (uint32_t b = mid; b < lens[q]; b += 2)
                        *reinterpret_cast<uint16_t *>(dsts[q] + b) = *reinterpret_cast<const uint16_t *>(srcs[q] + b);
                    tx += mid;
                }
                mbar_arrive_expect_tx(&full[s], tx);
                for (int q = 0; q < 3; ++q) {
                    const uint32_t mid = lens[q] & ~15u;
                    if (mid) bulk_g2s(dsts[q], srcs[q], mid, &full[s]);
                }
            }
        }
        return;
    }

    // ================================ consumer teams ================================
    const int team = warp / kTeamWarps, tw = warp % kTeamWarps;
    const int HCe = (HC + 1) & ~1;
    // end of kernel (the ring is free by then): [kTeams * kSlices][2*kHP][kFP] fp32 register partials, one block per (team, slice)
    float *scratch = reinterpret_cast<float *>(smem + L.stage0);
    constexpr int kBlock = 2 * kHP * kFP;

    if (tw < kDxWarps) {
        // -------- dx warps: thread = rows {lane, lane+32, lane+64, lane+96} x outputs 12u .. 12u+11 ------------------------
        const int u = tw;
        float *dxst = reinterpret_cast<float *>(smem + L.dxst + (size_t)team * L.dxst_bytes);
        for (int it = team; it < n_local; it += kTeams) {
            const int s = it % kStages;
            const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x, r0 = tile * kRows;
            const int nr = (int)((a.R - r0) < (int64_t)kRows ? (a.R - r0) : (int64_t)kRows);
            const unsigned char *st = smem + L.stage0 + (size_t)s * L.stage_bytes;
            mbar_wait(&full[s], (it / kStages) & 1);
            if (!need_dx) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                continue;
            }
            float2 o[4][kDxPairs];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < kDxPairs; ++k) o[i][k] = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int arr = 0; arr < 2; ++arr) {
                const ST *d = reinterpret_cast<const ST *>(st + (arr ? L.off_dr : 0u)) + lane * HC;
                const float *w = W_s + arr * kHP * kFP + kDxOut * u;
#pragma unroll 2
                for (int c = 0; c < HCe; c += 2) {
                    float2 dv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) dv[i] = ld_pair(d + 32 * i * HC + c);
                    float4 w0[kDxVec], w1[kDxVec];
#pragma unroll
                    for (int k = 0; k < kDxVec; ++k) {
                        w0[k] = *reinterpret_cast<const float4 *>(w + c * kFP + 4 * k);
                        w1[k] = *reinterpret_cast<const float4 *>(w + (c + 1) * kFP + 4 * k);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 a0 = make_float2(dv[i].x, dv[i].x), a1 = make_float2(dv[i].y, dv[i].y);
#pragma unroll
                        for (int k = 0; k < kDxVec; ++k) {
                            o[i][2 * k] = __ffma2_rn(a0, make_float2(w0[k].x, w0[k].y), o[i][2 * k]);
                            o[i][2 * k + 1] = __ffma2_rn(a0, make_float2(w0[k].z, w0[k].w), o[i][2 * k + 1]);
                            o[i][2 * k] = __ffma2_rn(a1, make_float2(w1[k].x, w1[k].y), o[i][2 * k]);
                            o[i][2 * k + 1] = __ffma2_rn(a1, make_float2(w1[k].z, w1[k].w), o[i][2 * k + 1]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (tw == 0 && lane == 0) bulk_wait_read0();  // the previous tile's bulk store has drained dxst
            named_bar(1 + team, kDxWarps * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < kDxPairs; ++k) {
                    const int f = kDxOut * u + 2 * k;
                    if (f < F) *reinterpret_cast<float2 *>(dxst + (lane + 32 * i) * F + f) = o[i][k];
                }
            fence_proxy_async();
            named_bar(1 + team, kDxWarps * 32);
            float *gdx = a.dx + r0 * F;
            if (nr == kRows) {
                if (tw == 0 && lane == 0) {
                    if (a.accumulate) bulk_s2g_add_f32(gdx, dxst, kRows * F * 4u);
                    else bulk_s2g(gdx, dxst, kRows * F * 4u);
                    bulk_commit();
                }
            } else {
                for (int i = tw * 32 + lane; i < nr * F; i += kDxWarps * 32) gdx[i] = a.accumulate ? gdx[i] + dxst[i] : dxst[i];
            }
        }
        if (need_dx && tw == 0 && lane == 0) bulk_wait0();
        named_bar(8, kConsumers);  // all tiles consumed: the ring is free
        named_bar(8, kConsumers);  // the dW warps have written their partial blocks
    } else {
        // -------- dW warps: thread = (16-row slice, gradient columns 8 ot .. 8 ot+7, input columns 12 u .. 12 u+11) -------------
        const int q = (tw - kDxWarps) * 32 + lane;  // 0 .. 95
        const int slice = q / (2 * kOT), ot = (q % (2 * kOT)) >> 1, u = q & 1;
        float2 acc[8][6];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[i][k] = make_float2(0.f, 0.f);
        const uint32_t d_off = (ot < kOT / 2 ? 0u : L.off_dr) + (uint32_t)(8 * (ot < kOT / 2 ? ot : ot - kOT / 2)) * (uint32_t)sizeof(ST);
        for (int it = team; it < n_local; it += kTeams) {
            const int s = it % kStages;
            const int64_t r0 = (blockIdx.x + (int64_t)it * gridDim.x) * kRows;
            const int nr = (int)((a.R - r0) < (int64_t)kRows ? (a.R - r0) : (int64_t)kRows);
            const unsigned char *st = smem + L.stage0 + (size_t)s * L.stage_bytes;
            const ST *d = reinterpret_cast<const ST *>(st + d_off);
            const float *xp = reinterpret_cast<const float *>(st + L.off_x) + 12 * u;
            mbar_wait(&full[s], (it / kStages) & 1);
            const ST *dp = d + slice * HC;   // pointer increments (integer adds on the ALU pipe) instead of r * HC (IMAD: FMA pipe)
            const float *xq = xp + slice * F;
#pragma unroll 1
            for (int r = slice; r < nr; r += kSlices, dp += kSlices * HC, xq += kSlices * F) {
                float2 dv[4], xv[6];
#pragma unroll
                for (int i = 0; i < 4; ++i) dv[i] = ld_pair(dp + 2 * i);
#pragma unroll
                for (int k = 0; k < 6; ++k) xv[k] = ld_pair(xq + 2 * k);
                if (u) xv[5].y = 1.f;  // column 23: the constant 1 of the bias gradient
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 a0 = make_float2(dv[i].x, dv[i].x), a1 = make_float2(dv[i].y, dv[i].y);
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        acc[2 * i][k] = __ffma2_rn(a0, xv[k], acc[2 * i][k]);
                        acc[2 * i + 1][k] = __ffma2_rn(a1, xv[k], acc[2 * i + 1][k]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // ---- CTA partial: every (team, slice) stores its register partials as one fp32 block; after ONE barrier each output is
        //      summed over the 24 blocks in fp64 in a fixed order (formerly 24 barrier-separated read-modify-write rounds in
        //      shared memory: ~45 us per launch, half of the kernel at B = 2) ------------------------------------------------
        named_bar(8, kConsumers);  // all tiles consumed: the ring is free
        {
            float *blk = scratch + (team * kSlices + slice) * kBlock;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    *reinterpret_cast<float2 *>(blk + (8 * ot + i) * kFP + 12 * u + 2 * k) = acc[i][k];
        }
        named_bar(8, kConsumers);
    }
    const int O = 2 * HC;
    float *out = a.partials + (int64_t)blockIdx.x * (O * F + O);
    for (int i = tid; i < O * F + O; i += kConsumers) {
        int o, f;
        if (i < O * F) { o = i / F; f = i - o * F; } else { o = i - O * F; f = kFP - 1; }
        const int arr = o >= HC, c = o - arr * HC;
        const float *src = scratch + (arr * kHP + c) * kFP + f;
        double v = 0.0;
#pragma unroll 4
        for (int g = 0; g < kTeams * kSlices; ++g) v += (double)src[g * kBlock];
        out[i] = (float)v;
    }
}

static int grid_for(int64_t R) {
    const int sms = tg_sm_count();
    const int64_t tiles = (R + kRows - 1) / kRows;
    return (int)(tiles < sms ? tiles : sms);
}
}  // namespace rt

bool project_bwd_rt_supported(int F, int HC, const void *dxl, const void *dxr, const void *x, const void *dx) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dxl) | reinterpret_cast<uintptr_t>(dxr) |
                           reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    return aligned && F >= 2 && F <= rt::kFP - 2 && (F % 2) == 0 && HC >= 2 && HC <= 48 && (HC % 2) == 0;
}

int64_t project_bwd_rt_workspace(int64_t R, int F, int HC) { return int64_t(rt::grid_for(R)) * (2 * HC * F + 2 * HC) * (int64_t)sizeof(float); }

int project_bwd_rt(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx, float *dwl,
                   float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype, cudaStream_t st,
                   bool accumulate, ReduceJob *defer) {
    rt::Args a;
    a.accumulate = accumulate ? 1 : 0;
    a.dxl = dxl; a.dxr = dxr; a.x = x; a.wl = wl; a.wr = wr; a.dx = dx; a.partials = static_cast<float *>(workspace);
    a.R = R; a.F = F; a.HC = HC;
    const int grid = rt::grid_for(R);
    auto launch = [&](auto kern, const rt::Smem &L) -> cudaError_t {
        cudaError_t e = tg_set_smem(reinterpret_cast<const void *>(kern), (int)L.total);
        if (e != cudaSuccess) return e;
        kern<<<grid, rt::kThreads, L.total, st>>>(a);
        tg_count_launch();
        return cudaSuccess;
    };
    if (HC <= 24) {
        if (dtype == TECGAT_F32) TG_CUDA(launch(rt::project_bwd_rt_kernel<float, 24>, rt::layout<float, 24>(F, HC)));
        else TG_CUDA(launch(rt::project_bwd_rt_kernel<__nv_bfloat16, 24>, rt::layout<__nv_bfloat16, 24>(F, HC)));
    } else {
        if (dtype == TECGAT_F32) TG_CUDA(launch(rt::project_bwd_rt_kernel<float, 48>, rt::layout<float, 48>(F, HC)));
        else TG_CUDA(launch(rt::project_bwd_rt_kernel<__nv_bfloat16, 48>, rt::layout<__nv_bfloat16, 48>(F, HC)));
    }
    TG_LAUNCH_CHECK();
    const int O = 2 * HC;
    ReduceSegs segs = {{dwl, dwr, dbl, dbr}, {0, HC * F, O * F, O * F + HC}, {HC * F, O * F, O * F + HC, O * F + O}};
    if (defer) {  // the caller finishes these partials together with the edge kernel's in one launch
        *defer = ReduceJob{a.partials, grid, O * F + O, segs};
        return TECGAT_OK;
    }
    return reduce_columns(a.partials, grid, O * F + O, segs, st);
}

}  // namespace tg
