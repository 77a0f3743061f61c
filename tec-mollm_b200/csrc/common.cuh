// common.cuh -- shared helpers: error plumbing for the C ABI and the sm_100a PTX wrappers
// (mbarrier, bulk-TMA copies, proxy fences, tcgen05) used by every kernel in this library.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

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

struct tecgat_plan {
    int32_t num_nodes = 0;
    int32_t tile_nodes = 0;
    int32_t num_tiles = 0;
    int64_t num_edges = 0;   // kept edges + N self loops
    int64_t kept_edges = 0;  // non-self edges of the input
    int32_t max_in_deg = 0;
    int32_t max_out_deg = 0;
    int32_t max_window = 0;  // max over tiles of (hi - lo)
    // device arrays (int32)
    int32_t *rowptr_in = nullptr;  // (N+1) destination-sorted CSR
    int32_t *col_in = nullptr;     // (E)   source node of slot k
    int32_t *rowptr_out = nullptr; // (N+1) source-sorted CSR
    int32_t *col_out = nullptr;    // (E)   destination node of out-slot k2
    int32_t *slot_out = nullptr;   // (E)   in-CSR slot k of out-slot k2 (dropout counter)
    int32_t *tile_lo = nullptr;    // (tiles) first row of the tile's source/destination window
    int32_t *tile_hi = nullptr;    // (tiles) one past the last row of the window
    // host copies kept for export / tests
    int32_t *h_rowptr_in = nullptr;
    int32_t *h_col_in = nullptr;
    int32_t *h_eid_in = nullptr;
    int device = 0;
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)

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
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end as a trapped kernel (sticky error the host reports), never
// as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) {
            printf("tecgat: mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, parity);
            __trap();
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
// keep(seed, snapshot, CSR slot, head).  The 64-bit seed and the snapshot index are folded ONCE per CTA into a 32-bit
// per-snapshot key; per (slot, head pair) one murmur3-style 32-bit finaliser (3 multiplies, 3 xor-shifts) yields two
// 16-bit uniforms, one per head of the pair.  P(drop) = round(p * 65536) / 65536.
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ uint32_t dropout_snapshot_key(uint64_t seed, uint32_t snapshot) {
    return fmix32(static_cast<uint32_t>(seed) ^ fmix32(static_cast<uint32_t>(seed >> 32) + 0x9E3779B9u * (snapshot + 1u)));
}
__host__ __device__ __forceinline__ uint32_t dropout_bits16(uint32_t key, uint32_t slot, uint32_t head) {
    const uint32_t h = fmix32((slot * 0x9E3779B1u) ^ key ^ ((head >> 1) * 0x7FEB352Du));
    return (h >> ((head & 1u) * 16u)) & 0xFFFFu;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
    // keep iff bits16 >= thr;  P(drop) = thr / 65536
    double t = static_cast<double>(p) * 65536.0;
    if (t < 0) t = 0;
    if (t > 65536.0) t = 65536.0;
    return static_cast<uint32_t>(t + 0.5);
}

}  // namespace tg
#endif  // __CUDACC__
