// edge_common.cuh -- pieces shared by the fused edge forward / backward kernels.
//
// Work decomposition (both kernels): one CTA = (tile of `T` consecutive destination nodes, one snapshot).
// Because the graph is identical across snapshots, the CSR, the tile windows and the parameters are the same
// for every CTA column; only the row slab changes.  Inside a CTA one LANE owns one (node, head) pair -- thread
// id = node_local * H + head -- so that a lane's C channels sit at shared-memory word stride C across lanes
// (C = 11: conflict-free), and the per-destination softmax needs no cross-lane traffic at all.
//
// The rows a tile touches (its nodes plus all their in- and out-neighbours) form the contiguous window
// [lo, hi) computed by the plan; the window's slab of xl (and xr / g / y in backward) is staged in shared
// memory by ONE bulk-TMA copy per array (cp.async.bulk, 16-byte aligned middle) plus a ragged <16-byte head
// and tail.  Windows that do not fit the shared-memory budget (arbitrary, non-banded graphs) fall back to
// gathering neighbour rows straight from global memory (L2) -- same code, template flag SM = false.
//
// Arithmetic: the kernels are instruction-issue bound (ncu: 79 % issue-active, DRAM traffic == algorithmic bytes), so
// the per-channel math runs on Blackwell's packed fp32 pipe: a lane's C channels are held as ceil(C/2) float2 pairs
// (zero padded) and every add / mul / fma is an FADD2 / FMUL2 / FFMA2.  LeakyReLU is max(s, slope*s), valid for the
// slopes the C ABI accepts (0 <= slope <= 1; PyG's default 0.2).
#pragma once
#include "common.cuh"

namespace tg {

__host__ __device__ __forceinline__ uint32_t round16(uint32_t b) { return (b + 15u) & ~15u; }

// Channel counts the edge kernels are instantiated for (C = out_channels per head).
#define TG_FOR_EACH_C(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(11) X(12) X(16) X(24) X(32)

template <typename ST>
struct Round {  // rounding applied by the reference's dtype flow to (xl_j + xr_i) and to leaky_relu(.)
    static __device__ __forceinline__ float2 r(float2 v) { return v; }
};
template <>
struct Round<__nv_bfloat16> {  // under autocast both are bf16 tensors
    static __device__ __forceinline__ float2 r(float2 v) { return __bfloat1622float2(__float22bfloat162_rn(v)); }
};

// C elements of storage type ST -> ceil(C/2) float2 pairs (odd C: the last .y is zero)
template <int C>
__device__ __forceinline__ void load_row(const float *__restrict__ p, float2 (&v)[(C + 1) / 2]) {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) v[i] = make_float2(p[2 * i], p[2 * i + 1]);
    if (C & 1) v[C / 2] = make_float2(p[C - 1], 0.f);
}
template <int C>
__device__ __forceinline__ void load_row(const __nv_bfloat16 *__restrict__ p, float2 (&v)[(C + 1) / 2]) {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) v[i] = make_float2(__bfloat162float(p[2 * i]), __bfloat162float(p[2 * i + 1]));
    if (C & 1) v[C / 2] = make_float2(__bfloat162float(p[C - 1]), 0.f);
}
template <int C>
__device__ __forceinline__ void store_row(float *p, const float2 (&v)[(C + 1) / 2]) {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) {
        p[2 * i] = v[i].x;
        p[2 * i + 1] = v[i].y;
    }
    if (C & 1) p[C - 1] = v[C / 2].x;
}
template <int C>
__device__ __forceinline__ void store_row(__nv_bfloat16 *p, const float2 (&v)[(C + 1) / 2]) {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) {
        p[2 * i] = __float2bfloat16_rn(v[i].x);
        p[2 * i + 1] = __float2bfloat16_rn(v[i].y);
    }
    if (C & 1) p[C - 1] = __float2bfloat16_rn(v[C / 2].x);
}

// s = xj + xr, z = LeakyReLU(s), returns e = att . z   (identical instruction sequence in forward and backward, so the
// alpha recomputed in backward matches the saved softmax statistics)
template <int C, typename ST>
__device__ __forceinline__ float edge_score(const float2 (&att)[(C + 1) / 2], const float2 (&xj)[(C + 1) / 2],
                                            const float2 (&xr)[(C + 1) / 2], float2 slope2, float2 (&s)[(C + 1) / 2],
                                            float2 (&z)[(C + 1) / 2]) {
    float2 e2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < (C + 1) / 2; ++i) {
        s[i] = Round<ST>::r(__fadd2_rn(xj[i], xr[i]));
        const float2 t = __fmul2_rn(s[i], slope2);
        z[i] = Round<ST>::r(make_float2(fmaxf(s[i].x, t.x), fmaxf(s[i].y, t.y)));
        e2 = __ffma2_rn(att[i], z[i], e2);
    }
    return e2.x + e2.y;
}

__device__ __forceinline__ float hsum(float2 v) { return v.x + v.y; }

// 2^x for x <= 0 (softmax weights): bare MUFU.EX2, flush-to-zero -- no denormal rescue sequence
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace tg
