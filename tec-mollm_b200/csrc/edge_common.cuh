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
// gathering neighbour rows straight from global memory (L2) -- same code, different base pointer.
#pragma once
#include "common.cuh"

namespace tg {

__host__ __device__ __forceinline__ uint32_t round16(uint32_t b) { return (b + 15u) & ~15u; }

// Channel counts the edge kernels are instantiated for (C = out_channels per head).
#define TG_FOR_EACH_C(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(11) X(12) X(16) X(24) X(32)

__device__ __forceinline__ float leaky(float s, float slope) { return s > 0.f ? s : s * slope; }

template <typename ST>
struct Round {  // rounding applied by the reference's dtype flow to (xl_j + xr_i) and to leaky_relu(.)
    static __device__ __forceinline__ float r(float v) { return v; }
};
template <>
struct Round<__nv_bfloat16> {  // under autocast both are bf16 tensors
    static __device__ __forceinline__ float r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

}  // namespace tg
