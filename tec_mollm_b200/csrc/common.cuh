// common.cuh -- shared helpers: error plumbing for the C ABI and the sm_100a PTX wrappers
// (mbarrier, bulk-TMA copies, proxy fences, tcgen05) used by every kernel in this library.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tecgat.h"

// ------------------------------------------------------------------------------------------------
// host-side error plumbing
// ------------------------------------------------------------------------------------------------
void tecgat_set_error(const char *fmt, ...);

#define TG_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            tecgat_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return TECGAT_ECUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define TG_REQUIRE(cond, code, ...)                                                                \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            tecgat_set_error(__VA_ARGS__);                                                         \
            return (code);                                                                         \
        }                                                                                          \
    } while (0)

#define TG_LAUNCH_CHECK() TG_CUDA(cudaGetLastError())

// ---- per-launch host overhead, paid once (plan.cu) ------------------------------------------------
// Number of kernels this library has launched since load (tecgat_launch_count): every `<<<>>>` is followed by
// tg_count_launch(), so a benchmark can report how many of OUR kernels ran inside its timed region.
void tg_count_launch(int n = 1);
// SM count of the current device (queried once per device).
int tg_sm_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device, size) instead of on every launch.
cudaError_t tg_set_smem(const void *kernel, int bytes);
// environment tuning knobs, read once per process (getenv on every launch shows up at B = 2)
const char *tg_env(const char *name);
// item schedule (plan.cu): (grid + 1) host array, valid for the plan's lifetime
struct tecgat_plan;
const int64_t *tg_item_bounds(const struct tecgat_plan *plan, bool bwd, int64_t snapshots, int grid, int64_t *max_items);

// One tiling of the node axis: tiles of T consecutive nodes, each with the contiguous row window [lo, hi) it touches and
// a "slab" -- the tile's slice of the graph in ELL form, laid out so that ONE bulk-TMA copy brings it into shared memory:
//   (Ts = T rounded up to a multiple of 8 is the row stride of every section below)
//   header  16 B : int32 {lo, hi, kin | kout << 16, slot_base}
//   k0      T x int32 : first in-CSR slot of node t        (dropout counter of its k-th in-edge = k0 + k)
//   deg     T x int32 : in-degree | out-degree << 16        (0 for the padding nodes of a ragged last tile)
//   ell_in  kin  x T x uint16 : window-relative row (col - lo) of the k-th in-neighbour   (self loop is k = 0; kin and
//                               kout are padded to ODD counts, unused entries are 0 = a valid window row)
//   -- backward tilings only --
//   ell_out kout x T x uint16 : window-relative row of the k-th out-neighbour
//   slot    kout x T x uint16 : in-CSR slot of that edge, relative to slot_base = rowptr_in[lo]
struct tg_tile_meta {  // 32 bytes; the kernels keep the table in shared memory
    int32_t lo, hi;
    int32_t kin_kout;  // kin | kout << 16
    int32_t eligible;  // 1: the slab exists (indices fit uint16)
    int64_t slab_off;  // byte offset of the tile's slab
    int32_t slab_bytes;
    int32_t pad;
};
struct tg_tiling {
    int32_t T = 0, num_tiles = 0, max_window = 0;
    bool bwd = false;
    tg_tile_meta *meta = nullptr;  // device (tiles)
    unsigned char *slabs = nullptr;  // device
    std::vector<tg_tile_meta> h_meta;
    std::vector<int64_t> h_slab_off;  // (tiles + 1)
    // Item schedule: the persistent CTAs take contiguous item ranges of equal estimated WORK, not equal count (a polar tile of
    // the 64,800-node grid costs 50 x an equatorial one).  h_cost_prefix[t] = sum of the cost estimates of tiles < t.
    std::vector<int64_t> h_cost_prefix;  // (tiles + 1)
};

// Sliding-window backward tiling (edge_bwd_sw.cu): chunks of T consecutive nodes whose in- and out-neighbours all lie within
// R rows (banded graph).  Two slabs per chunk, both with compile-time-free uniform sizes (plan-wide ELL row counts):
//   slabD (destination role): int32 hdr[4] = {wlo, whi, 0, 0}; int32 k0[T] (first in-CSR slot: dropout counters);
//         int32 deg[T] (in-degree | out-degree << 16, self loop included); uint16 ell_in[kinp][T]: window-relative row
//         (col - wlo) of in-slot k = 1 + row (self loop excluded), padding = the node's own row
//   slabS (source role): uint16 ell_out[koutp][T]: window-relative row of out-slot k (self loop included, k = 0);
//         uint16 st_out[koutp][T]: stash address of that edge's (alpha q, d e) pair relative to the window start, in 8-byte
//         units = u_rel * (stash_stride / 8) + in_slot * 2; padding = a slot the destination role zero-fills
// window of chunk c: [wlo, whi) = [max(0, cT - R), min(N, (c+1)T + R))
struct tg_sw_plan {
    int32_t T = 0, J = 0, R = 0;
    int32_t kin = 0, kout = 0;      // max in- / out-degree, self loop included
    int32_t kinp = 0, koutp = 0;    // ELL rows: in (self excluded), out (self included), both padded to even counts
    int32_t stash_stride = 0;       // bytes per node of the (alpha q, d e) stash: kin entries of 16 B + 8 B (bank spread)
    int32_t slabD_bytes = 0, slabS_bytes = 0;
    unsigned char *slabD = nullptr, *slabS = nullptr;  // device: J slabs each
};

struct tecgat_plan {
    int32_t num_nodes = 0;
    int64_t num_edges = 0;   // kept edges + N self loops
    int64_t kept_edges = 0;  // non-self edges of the input
    int32_t max_in_deg = 0;
    int32_t max_out_deg = 0;
    // device arrays (int32): the two CSR orientations; every row starts with the node's self loop
    int32_t *rowptr_in = nullptr;  // (N+1) destination-sorted CSR
    int32_t *col_in = nullptr;     // (E)   source node of slot k
    int32_t *rowptr_out = nullptr; // (N+1) source-sorted CSR
    int32_t *col_out = nullptr;    // (E)   destination node of out-slot k2
    int32_t *slot_out = nullptr;   // (E)   in-CSR slot k of out-slot k2 (dropout counter)
    tg_tiling fwd, bwd;
    // host copies kept for export / tests
    int32_t *h_rowptr_in = nullptr;
    int32_t *h_col_in = nullptr;
    int32_t *h_eid_in = nullptr;
    int device = 0;
    // launch geometry of the edge kernels (ring depth, staged caps, shared-memory map): derived once per
    // (kernel, heads, channels, dtype) from the tilings instead of on every launch; POD blobs keyed by the launcher
    mutable std::mutex cache_mu;
    mutable std::map<uint64_t, std::vector<unsigned char>> geom_cache;
    // item schedule of the persistent edge kernels per (direction, snapshots, grid): bounds[b] .. bounds[b + 1] = the items of
    // CTA b, contiguous ranges of equal estimated work (tg_tiling::h_cost_prefix); computed once, handed to the kernels as a
    // kernel-parameter array (constant bank: the CTA's range stays in uniform registers, which a global load would not)
    struct ItemBounds {
        std::vector<int64_t> host;  // (grid + 1)
        int64_t max_items = 0;      // largest range
    };
    mutable std::map<uint64_t, ItemBounds> bounds_cache;
    // sliding-window backward tiling (edge_bwd_sw.cu), built by plan_create when the graph is banded
    struct tg_sw_plan *sw = nullptr;
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)

#ifndef TG_PRODUCER_SLEEP_NS
#define TG_PRODUCER_SLEEP_NS 400
#endif

namespace tg {

constexpr int kSmemBudget = 200 * 1024;  // per-CTA dynamic shared memory we are willing to request

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    // suspend-time hint (ns): the thread sleeps in hardware until the phase completes (wake-up ~60 cycles after the
    // arrive) instead of spinning through issue slots the working warps need
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end as a trapped kernel (sticky error the host reports), never
// as a hung GPU.
// A sleeping waiter is woken by every update of the barrier (each arrive, each bulk-copy chunk): measured 7 - 18 polls per
// wait in the edge kernels, i.e. the poll loop is 8 - 14 % of their warp instructions -- so the clock is only read every
// 256th poll (a poll is then try_wait + sleep + re-check + counter: 7 instructions instead of 10).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 255u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000LL) {
                printf("tecgat: mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x,
                       (int)threadIdx.x, parity);
                __trap();
            }
        }
    }
}

// Producer-side wait: the producer warp runs a whole item ahead of the consumers, so wake-up latency is irrelevant; sleeping
// between polls leaves the issue slots of its scheduler to the two consumer warps it shares it with (try_wait's own suspend
// hint returns within ~20 ns: a bare poll loop issues ~5 instructions every ~40 cycles).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(TG_PRODUCER_SLEEP_NS);
        if ((++polls & 255u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) {
                printf("tecgat: producer mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, parity);
                __trap();
            }
        }
    }
}

// ---- bulk TMA (cp.async.bulk): 1-D, 16-byte aligned, size a multiple of 16 ------------------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
// bulk reduction: global fp32 += shared fp32, element-wise, performed at the L2 (one add per element: deterministic)
__device__ __forceinline__ void bulk_s2g_add_f32(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
// L2 prefetch of a contiguous global range (16-byte aligned, size a multiple of 16): no destination, no completion
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to smem -> visible to the async proxy (bulk stores, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Contiguous global -> shared copy of `bytes` bytes starting at an arbitrarily aligned global address.
// The shared destination is `smem_base + (gaddr & 15)` so that the 16-byte aligned middle can go through
// one bulk-TMA copy (issued by `leader`, completing on `bar`), while the <16-byte ragged head and tail are
// moved with plain 2-byte loads by a few threads.  Returns the shared address of byte 0 of the range.
// Every thread of the CTA must call it with identical arguments; `*tx_bytes` accumulates what the leader
// must announce with expect_tx BEFORE the copies are issued, so the call is split in two phases.
struct CopyPlan {
    const char *g;       // global start
    char *s;             // shared address of byte 0
    uint32_t head, mid, tail;
};
__device__ __forceinline__ CopyPlan plan_copy(const void *gsrc, void *smem_base16, uint32_t bytes) {
    CopyPlan c;
    c.g = static_cast<const char *>(gsrc);
    const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(gsrc) & 15u);
    c.s = static_cast<char *>(smem_base16) + a;
    uint32_t head = (16u - a) & 15u;
    if (head > bytes) head = bytes;
    const uint32_t rest = bytes - head;
    c.head = head;
    c.mid = rest & ~15u;
    c.tail = rest - c.mid;
    return c;
}
__device__ __forceinline__ void issue_copy_bulk(const CopyPlan &c, uint64_t *bar) {
    if (c.mid) bulk_g2s(c.s + c.head, c.g + c.head, c.mid, bar);
}
// ragged ends: 2-byte granularity (all our element types are >= 2 bytes and 2-byte aligned)
__device__ __forceinline__ void copy_ragged(const CopyPlan &c, int tid) {
    const uint32_t nh = c.head >> 1, nt = c.tail >> 1;
    if (tid < (int)nh) {
        reinterpret_cast<uint16_t *>(c.s)[tid] = reinterpret_cast<const uint16_t *>(c.g)[tid];
    } else if (tid < (int)(nh + nt)) {
        const uint32_t off = c.head + c.mid + ((tid - nh) << 1);
        *reinterpret_cast<uint16_t *>(c.s + off) = *reinterpret_cast<const uint16_t *>(c.g + off);
    }
}

// ---- element load helpers (storage dtype -> fp32) ---------------------------------------------
__device__ __forceinline__ float ld_elem(const float *p) { return *p; }
__device__ __forceinline__ float ld_elem(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_elem(float *p, float v) { *p = v; }
__device__ __forceinline__ void st_elem(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// ---- counter-based dropout RNG: integer-only, restated on the host (tecgat_dropout_mask_host) -----------------
// keep(seed, snapshot, head, CSR slot).  Every draw of one launch has its OWN 32-bit counter
//     n = (snapshot * heads + head) * edges_per_snapshot + slot
// (unique while snapshots * heads * edges <= 2^32: 2.9e8 at B = 128 on the 2911-node graph), so no two (snapshot, head)
// streams can be shifted copies of one another.  h = n * kDropMul + base(seed) is a Weyl sequence whose start
// mixes all 64 seed bits; one xor-shift and one multiply finish it into 32 uniform bits that are compared against
// round(p * 2^32).  Per (item, lane) the kernels form key = base + stream * (edges * kDropMul) once; per edge they spend one
// add (h advances by kDropMul), one shift, one xor, one multiply, one compare, one select.
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
constexpr uint32_t kDropMul = 0x9E3779B1u;
__host__ __device__ __forceinline__ uint32_t dropout_base(uint64_t seed) {  // once per thread
    const uint32_t lo = static_cast<uint32_t>(seed), hi = static_cast<uint32_t>(seed >> 32);
    return fmix32(lo ^ fmix32(hi + 0x9E3779B9u));
}
// stream_stride = edges_per_snapshot * kDropMul (mod 2^32): the counter distance between consecutive (snapshot, head) streams
__host__ __device__ __forceinline__ uint32_t dropout_stream_stride(int64_t edges_per_snapshot) {
    return static_cast<uint32_t>(static_cast<uint64_t>(edges_per_snapshot)) * kDropMul;
}
__host__ __device__ __forceinline__ uint32_t dropout_key(uint32_t base, uint32_t snapshot, uint32_t heads, uint32_t head,
                                                         uint32_t stream_stride) {  // once per (snapshot, lane)
    return base + (snapshot * heads + head) * stream_stride;
}
// second half of the hash; `h` = slot * kDropMul + key (consecutive slots: h advances by kDropMul, one integer add)
__host__ __device__ __forceinline__ uint32_t dropout_finish(uint32_t h) {
    h ^= h >> 15;
    h *= 0x85EBCA6Bu;
    return h;
}
__host__ __device__ __forceinline__ uint32_t dropout_bits(uint32_t key, uint32_t slot) { return dropout_finish(slot * kDropMul + key); }
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
    // keep iff bits >= thr;  P(drop) = thr / 2^32
    double t = static_cast<double>(p) * 4294967296.0;
    if (t < 0) t = 0;
    if (t > 4294967295.0) t = 4294967295.0;
    return static_cast<uint32_t>(t + 0.5);
}

}  // namespace tg
#endif  // __CUDACC__
