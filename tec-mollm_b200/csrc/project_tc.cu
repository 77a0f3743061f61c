// project_tc.cu -- the lin_l / lin_r projections as ONE tensor-core GEMM per direction:
//     fwd : [xl | xr] = x [Wl; Wr]^T + [bl | br]      (M = rows, K = in_channels, N = 2*H*C)
// tcgen05.mma (UTCHMMA) with the accumulator in TMEM, the x tiles streamed by bulk-TMA (cp.async.bulk / UBLKCP)
// through an mbarrier ring, results read back with tcgen05.ld and written with bulk-TMA stores.
//
// The reference's rows are 88 bytes (F = 22 fp32): not a legal TMA tensor-map stride (multiple of 16 required) and not a
// UMMA operand layout, so a tile travels as ONE contiguous 1-D bulk copy (128 rows are contiguous in memory) and the
// four worker warps re-lay it into the canonical core-matrix layout (tc.cuh) while splitting it for precision:
//   * fp32 contract : 3xTF32 -- x = hi + lo, W = hi + lo (hi = tf32(x), lo = x - hi exactly);
//                     D = hi*hi + lo*hi + hi*lo in fp32 -> ~2^-21 relative, inside the 1e-5 parity gate that plain TF32
//                     (2^-11) would miss.  The GEMM is HBM-bound (7 flop/B), so the 3x MMA count is free.
//   * bf16 contract : operands rounded to bf16 (what autocast's Linear does), one kind::f16 pass.
//
// CTA = 6 warps, persistent over row tiles of 128:  warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = workers (one row per thread: operand re-layout for tile i, then epilogue of tile i-1, so the tensor pipe,
// the TMA engine and the LSU overlap); two CTAs per SM.
#include <type_traits>

#include "project.cuh"
#include "tc.cuh"

namespace tg {

constexpr int kTileM = 128;
constexpr int kMaxStages = 4;
constexpr uint32_t kSmemTwoCtas = 113 * 1024;  // dynamic shared memory that still lets two CTAs share an SM
constexpr int kTcThreads = 192;

struct TcFwdArgs {
    const float *x, *wl, *bl, *wr, *br;
    void *xl, *xr;
    int64_t R;
    int32_t F, HC;
    int32_t KP;  // K padded: multiple of 8 (tf32) / 16 (bf16)
    int32_t NP;  // N padded: multiple of 16
    int32_t acc_stride;  // TMEM columns between the two accumulator stages (power of two >= NP)
    int32_t stages;      // depth of the x-tile ring (2..kMaxStages)
    int32_t tmem_cols;
};

struct TcFwdSmem {  // byte offsets into dynamic shared memory
    uint32_t bars, tmem_ptr, bias, b_hi, b_lo, a_hi, a_lo, xs, out_l, out_r, total;
    uint32_t stage_bytes, P_a, P_b;
};

template <bool BF16>
__host__ __device__ inline TcFwdSmem tc_fwd_smem(int F, int HC, int KP, int NP, int stages) {
    const uint32_t elem = BF16 ? 2 : 4;
    const uint32_t chunks = KP * elem / 16;  // 16-byte K chunks per row
    TcFwdSmem s;
    uint32_t o = 0;
    s.bars = o; o += 128;
    s.tmem_ptr = o; o += 16;
    s.bias = o; o += ((NP * 4 + 15) / 16) * 16;
    o = (o + 127) & ~127u;
    s.P_b = chunks * 128;
    s.P_a = chunks * 128;
    s.b_hi = o; o += (NP / 8) * s.P_b;
    s.b_lo = o; o += BF16 ? 0 : (NP / 8) * s.P_b;
    s.a_hi = o; o += (kTileM / 8) * s.P_a;
    s.a_lo = o; o += BF16 ? 0 : (kTileM / 8) * s.P_a;
    s.stage_bytes = ((kTileM * F * 4 + 127) / 128) * 128;
    s.xs = o; o += stages * s.stage_bytes;
    const uint32_t out_bytes = ((kTileM * HC * elem + 127) / 128) * 128;
    s.out_l = o; o += out_bytes;
    s.out_r = o; o += out_bytes;
    s.total = o;
    return s;
}

// Canonical-layout writer for one 16-byte chunk (see tc.cuh): row r, chunk index kc.
__device__ __forceinline__ uint32_t canon_off(int r, int kc, uint32_t P) {
    return (uint32_t)(r >> 3) * P + (uint32_t)kc * 128u + (uint32_t)(r & 7) * 16u;
}

// Convert `n` fp32 values of one row (zero padded to the chunk grid) into canonical chunks.
//   tf32: 4 values per chunk, hi -> base_hi, lo -> base_lo;   bf16: 8 values per chunk -> base_hi only.
template <bool BF16, typename LoadFn>
__device__ __forceinline__ void write_row_canonical(unsigned char *base_hi, unsigned char *base_lo, int r, uint32_t P, int n,
                                                    int KP, LoadFn load) {
    if constexpr (BF16) {
        for (int kc = 0; kc < KP / 8; ++kc) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = kc * 8 + 2 * i;
                const float v0 = k < n ? load(k) : 0.f;
                const float v1 = k + 1 < n ? load(k + 1) : 0.f;
                __nv_bfloat162 b = __floats2bfloat162_rn(v0, v1);
                w[i] = *reinterpret_cast<uint32_t *>(&b);
            }
            *reinterpret_cast<uint4 *>(base_hi + canon_off(r, kc, P)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        for (int kc = 0; kc < KP / 4; ++kc) {
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = kc * 4 + i;
                const float v = k < n ? load(k) : 0.f;
                hi[i] = tf32_rna(v);
                lo[i] = v - hi[i];
            }
            *reinterpret_cast<float4 *>(base_hi + canon_off(r, kc, P)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(base_lo + canon_off(r, kc, P)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

template <bool BF16>
__global__ void __launch_bounds__(kTcThreads, 2) project_fwd_tc_kernel(const TcFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    using ST = typename std::conditional<BF16, __nv_bfloat16, float>::type;
    const int F = a.F, HC = a.HC, KP = a.KP, NP = a.NP;
    const int kStages = a.stages;
    const TcFwdSmem L = tc_fwd_smem<BF16>(F, HC, KP, NP, kStages);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint64_t *x_full = bars, *x_empty = bars + kMaxStages;
    uint64_t *a_ready = bars + 2 * kMaxStages, *a_free = a_ready + 1;
    uint64_t *t_full = a_ready + 2, *t_empty = a_ready + 4;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + L.tmem_ptr);
    float *b_s = reinterpret_cast<float *>(smem + L.bias);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (a.R + kTileM - 1) / kTileM;
    const int n_local = blockIdx.x < num_tiles ? (int)((num_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

    // ---- one-time set-up: barriers, TMEM, B operand = [Wl; Wr] (N x K, K-major), bias ---------------------------
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&x_full[s], 1);
            mbar_init(&x_empty[s], 128);
        }
        mbar_init(a_ready, 128);
        mbar_init(a_free, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, a.tmem_cols);
    for (int n = tid; n < NP; n += kTcThreads) {
        const float *wrow = n < HC ? a.wl + (int64_t)n * F : a.wr + (int64_t)(n - HC) * F;
        const int valid = n < 2 * HC ? F : 0;
        write_row_canonical<BF16>(smem + L.b_hi, smem + L.b_lo, n, L.P_b, valid, KP, [&](int k) { return wrow[k]; });
        float b = n < HC ? a.bl[n] : (n < 2 * HC ? a.br[n - HC] : 0.f);
        if (BF16) b = __bfloat162float(__float2bfloat16_rn(b));
        b_s[n] = b;
    }
    fence_proxy_async();  // B operand written through the generic proxy, read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer: one contiguous bulk copy per 128-row tile ==========================================
        if (lane == 0) {
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int s = it % kStages;
                if (it >= kStages) mbar_wait(&x_empty[s], ((it / kStages) - 1) & 1);
                const int64_t r0 = tile * kTileM;
                const int nr = (int)((a.R - r0) < (int64_t)kTileM ? (a.R - r0) : (int64_t)kTileM);
                const uint32_t bytes = (uint32_t)nr * F * 4u, mid = bytes & ~15u;
                const float *src = a.x + r0 * F;
                float *dst = reinterpret_cast<float *>(smem + L.xs + s * L.stage_bytes);
                for (uint32_t w = mid / 4; w < bytes / 4; ++w) dst[w] = src[w];  // <16-byte ragged end of the last tile
                mbar_arrive_expect_tx(&x_full[s], mid);
                if (mid) bulk_g2s(dst, src, mid, &x_full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ==========================================================================================
        const uint32_t idesc = umma_idesc(BF16 ? kFmtBF16 : kFmtTF32, kTileM, NP, 0, 0);
        const uint32_t a_hi = smem_u32(smem + L.a_hi), a_lo = smem_u32(smem + L.a_lo);
        const uint32_t b_hi = smem_u32(smem + L.b_hi), b_lo = smem_u32(smem + L.b_lo);
        const int ksteps = BF16 ? KP / 16 : KP / 8;
        for (int it = 0; it < n_local; ++it) {
            mbar_wait(a_ready, it & 1);
            if (it >= 2) mbar_wait(&t_empty[it & 1], ((it >> 1) - 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t d = tmem_base + (uint32_t)(it & 1) * a.acc_stride;
                uint32_t acc = 0;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t ko = ks * 256;  // two 16-byte K chunks per MMA
                    if (BF16) {
                        umma_bf16(d, umma_desc(a_hi + ko, 128, L.P_a), umma_desc(b_hi + ko, 128, L.P_b), idesc, acc);
                        acc = 1;
                    } else {
                        umma_tf32(d, umma_desc(a_lo + ko, 128, L.P_a), umma_desc(b_hi + ko, 128, L.P_b), idesc, acc);
                        umma_tf32(d, umma_desc(a_hi + ko, 128, L.P_a), umma_desc(b_lo + ko, 128, L.P_b), idesc, 1);
                        umma_tf32(d, umma_desc(a_hi + ko, 128, L.P_a), umma_desc(b_hi + ko, 128, L.P_b), idesc, 1);
                        acc = 1;
                    }
                }
                umma_commit(a_free);             // operand buffer may be overwritten
                umma_commit(&t_full[it & 1]);    // accumulator ready for the epilogue
            }
            __syncwarp();
        }
    } else {
        // ===== workers: re-layout of tile `it`, then epilogue of tile `it - 1` ================================
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;
        const bool issuer = (warp == 2 && lane == 0);
        ST *out_l = reinterpret_cast<ST *>(smem + L.out_l), *out_r = reinterpret_cast<ST *>(smem + L.out_r);
        for (int it = 0; it <= n_local; ++it) {
            if (it < n_local) {
                const int s = it % kStages;
                mbar_wait(&x_full[s], (it / kStages) & 1);
                if (it >= 1) mbar_wait(a_free, (it - 1) & 1);
                const float *xrow = reinterpret_cast<const float *>(smem + L.xs + s * L.stage_bytes) + row * F;
                write_row_canonical<BF16>(smem + L.a_hi, smem + L.a_lo, row, L.P_a, F, KP, [&](int k) { return xrow[k]; });
                fence_proxy_async();
                mbar_arrive(a_ready);
                mbar_arrive(&x_empty[s]);
            }
            if (it >= 1) {
                const int j = it - 1, acc = j & 1;
                const int64_t tile = blockIdx.x + (int64_t)j * gridDim.x;
                const int64_t r0 = tile * kTileM;
                const int nr = (int)((a.R - r0) < (int64_t)kTileM ? (a.R - r0) : (int64_t)kTileM);
                mbar_wait(&t_full[acc], (j >> 1) & 1);
                tc_fence_after();
                if (issuer) bulk_wait_read0();   // previous tile's bulk stores have drained the staging buffers
                named_bar_sync(1, 128);
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * a.acc_stride;
                for (int cb = 0; cb < NP / 16; ++cb) {
                    float v[16];
                    tmem_ld16(taddr + cb * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int n = cb * 16 + i;
                        if (n < HC) st_elem(out_l + row * HC + n, v[i] + b_s[n]);
                        else if (n < 2 * HC) st_elem(out_r + row * HC + (n - HC), v[i] + b_s[n]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&t_empty[acc]);
                fence_proxy_async();
                named_bar_sync(1, 128);
                ST *gl = static_cast<ST *>(a.xl) + r0 * HC, *gr = static_cast<ST *>(a.xr) + r0 * HC;
                if (nr == kTileM) {
                    if (issuer) {
                        bulk_s2g(gl, out_l, kTileM * HC * (uint32_t)sizeof(ST));
                        bulk_s2g(gr, out_r, kTileM * HC * (uint32_t)sizeof(ST));
                        bulk_commit();
                    }
                } else {  // ragged last tile: plain coalesced stores
                    const int wt = (warp - 2) * 32 + lane;
                    for (int i = wt; i < nr * HC; i += 128) {
                        gl[i] = out_l[i];
                        gr[i] = out_r[i];
                    }
                }
            }
        }
        if (issuer) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

static int pow2_cols(int need) {
    int c = 32;
    while (c < need) c <<= 1;
    return c;
}

bool project_tc_supported(int F, int HC) {
    if (F < 1 || F > 256 || HC < 1 || 2 * HC > 256) return false;
    const int NP = ((2 * HC + 15) / 16) * 16;
    if (2 * pow2_cols(NP) > 256) return false;  // two CTAs per SM share the 512 TMEM columns
    const int KPt = ((F + 7) / 8) * 8;
    return tc_fwd_smem<false>(F, HC, KPt, NP, 2).total <= 200 * 1024;
}

template <bool BF16>
static int launch_fwd_tc(TcFwdArgs &a, cudaStream_t st) {
    a.KP = BF16 ? ((a.F + 15) / 16) * 16 : ((a.F + 7) / 8) * 8;
    a.NP = ((2 * a.HC + 15) / 16) * 16;
    a.acc_stride = pow2_cols(a.NP);
    a.tmem_cols = 2 * a.acc_stride;
    a.stages = kMaxStages;  // deepest ring that keeps two CTAs per SM; otherwise one CTA per SM with the full ring
    while (a.stages > 2 && tc_fwd_smem<BF16>(a.F, a.HC, a.KP, a.NP, a.stages).total > kSmemTwoCtas) --a.stages;
    if (tc_fwd_smem<BF16>(a.F, a.HC, a.KP, a.NP, a.stages).total > kSmemTwoCtas) a.stages = kMaxStages;
    while (a.stages > 2 && tc_fwd_smem<BF16>(a.F, a.HC, a.KP, a.NP, a.stages).total > 200u * 1024u) --a.stages;
    const TcFwdSmem L = tc_fwd_smem<BF16>(a.F, a.HC, a.KP, a.NP, a.stages);
    TG_REQUIRE(L.total <= 200u * 1024u, TECGAT_ENOSUP, "project_fwd(tc): F=%d, HC=%d needs %u B shared memory", a.F, a.HC, L.total);
    auto kern = project_fwd_tc_kernel<BF16>;
    TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    const int64_t tiles = (a.R + kTileM - 1) / kTileM;
    const int grid = (int)(tiles < 2 * 148 ? tiles : 2 * 148);
    kern<<<grid, kTcThreads, L.total, st>>>(a);
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

int project_fwd_tc(const float *x, const float *wl, const float *bl, const float *wr, const float *br, void *xl, void *xr,
                   int64_t R, int F, int HC, int dtype, cudaStream_t st) {
    TG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(xl) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(xr) & 15) == 0,
               TECGAT_EINVAL, "project_fwd(tc): x, xl and xr must be 16-byte aligned");
    TcFwdArgs a;
    a.x = x; a.wl = wl; a.bl = bl; a.wr = wr; a.br = br; a.xl = xl; a.xr = xr;
    a.R = R; a.F = F; a.HC = HC;
    return dtype == TECGAT_BF16 ? launch_fwd_tc<true>(a, st) : launch_fwd_tc<false>(a, st);
}

// ---- backward: tensor-core version lands next; until then the C ABI routes TC requests for the backward to the
//      CUDA-core kernels (same results, checked by the same tests) ------------------------------------------------------
int64_t project_bwd_tc_workspace(int64_t R, int F, int HC) { return project_bwd_ffma_workspace(R, F, HC); }
int project_bwd_tc(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx,
                   float *dwl, float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype,
                   cudaStream_t st) {
    return project_bwd_ffma(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, workspace, R, F, HC, dtype, st);
}

}  // namespace tg
