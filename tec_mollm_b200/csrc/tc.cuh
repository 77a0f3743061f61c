// tc.cuh -- tcgen05 / TMEM PTX wrappers and UMMA descriptor builders (sm_100a) for the projection GEMMs.
//
// Shared-memory operand layout used everywhere in project_tc.cu ("canonical, no swizzle"): a (rows x cols) tile of
// 4-byte (tf32) or 2-byte (bf16) elements is stored as 8-row x 16-byte core matrices, each 128 contiguous bytes:
//     byte_offset(r, c) = (r / 8) * P + (c / EPC) * Q + (r % 8) * 16 + (c % EPC) * ELEM      EPC = 16 / ELEM
// with Q = stride between 16-byte column chunks and P = stride between 8-row blocks.  One layout serves both roles:
//   * K-major operand  (MN = rows, K = cols):  SBO = P, LBO = Q; one MMA consumes 32 bytes of K = 2 chunks -> += 2Q
//   * MN-major operand (MN = cols, K = rows):  SBO = Q, LBO = P; one MMA consumes 8 rows of K (tf32)       -> += P
// (SBO = stride between core matrices along MN, LBO = along K, both >> 4 in the descriptor; descriptor version 1.)
#pragma once
#include "common.cuh"

namespace tg {

// ---- shared-memory matrix descriptor (64-bit) -----------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;  // descriptor version 1 (Blackwell); base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
    return d;
}

// ---- instruction descriptor (32-bit), kind::tf32 / kind::f16, fp32 accumulate ------------------------------
constexpr uint32_t kFmtBF16 = 1, kFmtTF32 = 2;
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                   uint32_t b_mn_major) {
    return (1u << 4)                 // c_format = F32
           | (fmt << 7)              // a_format
           | (fmt << 10)             // b_format
           | (a_mn_major << 15)      // 0 = K-major, 1 = MN-major
           | (b_mn_major << 16)
           | ((N >> 3) << 17)        // n_dim
           | ((M >> 4) << 24);       // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread -> one arrival on `bar` when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp), columns: power of two >= 32 ---------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns ---------------------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace tg
