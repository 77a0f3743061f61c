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
#include "reduce.cuh"
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

// Compile-time F: the row is read with 8-byte loads and every bound folds away (straight-line code).
template <bool BF16, int FT>
__device__ __forceinline__ void write_row_canonical_fast(unsigned char *base_hi, unsigned char *base_lo, int r, uint32_t P,
                                                         const float *xrow) {
    constexpr int KP = BF16 ? ((FT + 15) / 16) * 16 : ((FT + 7) / 8) * 8;
    float v[KP];
#pragma unroll
    for (int i = 0; i < FT / 2; ++i) {
        const float2 t = *reinterpret_cast<const float2 *>(xrow + 2 * i);
        v[2 * i] = t.x;
        v[2 * i + 1] = t.y;
    }
#pragma unroll
    for (int i = FT; i < KP; ++i) v[i] = 0.f;
    if constexpr (BF16) {
#pragma unroll
        for (int kc = 0; kc < KP / 8; ++kc) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 b = __floats2bfloat162_rn(v[kc * 8 + 2 * i], v[kc * 8 + 2 * i + 1]);
                w[i] = *reinterpret_cast<uint32_t *>(&b);
            }
            *reinterpret_cast<uint4 *>(base_hi + canon_off(r, kc, P)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
#pragma unroll
        for (int kc = 0; kc < KP / 4; ++kc) {
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                hi[i] = tf32_rna(v[kc * 4 + i]);
                lo[i] = v[kc * 4 + i] - hi[i];
            }
            *reinterpret_cast<float4 *>(base_hi + canon_off(r, kc, P)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(base_lo + canon_off(r, kc, P)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// FT / HT > 0: compile-time in_channels / heads*out_channels (both even): straight-line worker code; 0: generic.
template <bool BF16, int FT, int HT>
__global__ void __launch_bounds__(kTcThreads, 2) project_fwd_tc_kernel(const TcFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    using ST = typename std::conditional<BF16, __nv_bfloat16, float>::type;
    constexpr bool kFast = FT > 0;
    const int F = kFast ? FT : a.F, HC = kFast ? HT : a.HC, KP = a.KP, NP = a.NP;
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
                if constexpr (kFast) write_row_canonical_fast<BF16, FT>(smem + L.a_hi, smem + L.a_lo, row, L.P_a, xrow);
                else write_row_canonical<BF16>(smem + L.a_hi, smem + L.a_lo, row, L.P_a, F, KP, [&](int k) { return xrow[k]; });
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
                if constexpr (kFast) {
                    constexpr int NPc = ((2 * HT + 15) / 16) * 16;
#pragma unroll
                    for (int cb = 0; cb < NPc / 16; ++cb) {
                        float v[16];
                        tmem_ld16(taddr + cb * 16, v);
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {  // pairs never straddle the two outputs (HT is even)
                            const int n = cb * 16 + i;
                            if (n < 2 * HT) {
                                ST *dst = n < HT ? out_l + row * HT + n : out_r + row * HT + (n - HT);
                                const float2 bb = *reinterpret_cast<const float2 *>(b_s + n);
                                const float2 o = make_float2(v[i] + bb.x, v[i + 1] + bb.y);
                                if constexpr (BF16) {
                                    __nv_bfloat162 h2 = __float22bfloat162_rn(o);
                                    *reinterpret_cast<__nv_bfloat162 *>(dst) = h2;
                                } else {
                                    *reinterpret_cast<float2 *>(dst) = o;
                                }
                            }
                        }
                    }
                } else {
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

template <bool BF16, int FT, int HT>
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
    auto kern = project_fwd_tc_kernel<BF16, FT, HT>;
    TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)L.total));
    const int64_t tiles = (a.R + kTileM - 1) / kTileM;
    const int grid = (int)(tiles < 2 * 148 ? tiles : 2 * 148);
    kern<<<grid, kTcThreads, L.total, st>>>(a); tg_count_launch();
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
    // straight-line specialisations for the reference's shapes (train.py:263-266 default, README variant); generic otherwise
    if (F == 22 && HC == 22) return dtype == TECGAT_BF16 ? launch_fwd_tc<true, 22, 22>(a, st) : launch_fwd_tc<false, 22, 22>(a, st);
    if (F == 10 && HC == 10) return dtype == TECGAT_BF16 ? launch_fwd_tc<true, 10, 10>(a, st) : launch_fwd_tc<false, 10, 10>(a, st);
    if (F == 22 && HC == 44) return dtype == TECGAT_BF16 ? launch_fwd_tc<true, 22, 44>(a, st) : launch_fwd_tc<false, 22, 44>(a, st);
    return dtype == TECGAT_BF16 ? launch_fwd_tc<true, 0, 0>(a, st) : launch_fwd_tc<false, 0, 0>(a, st);
}

// =====================================================================================================================
// backward:  dx = [dxl | dxr] [Wl; Wr]            (GEMM1: M = rows, K = 2*H*C, N = in_channels)
//            [dWl; dWr | dbl; dbr] = [dxl | dxr]^T [x | 1]   (GEMM2: M = 2*H*C, K = rows, N = in_channels + 1)
// One pass over dxl, dxr, x.  The row tile of [dxl | dxr] is laid out ONCE in shared memory (canonical layout, tc.cuh) and
// read twice by the tensor core: as the K-major A operand of GEMM1 and as the MN-major A operand of GEMM2 -- no
// transposition pass.  The x tile gets a constant-one column, so the bias gradients fall out of GEMM2 for free.
// GEMM2 accumulates in TMEM across all the CTA's tiles; per-CTA partials + the fixed-order fp64 second stage
// (reduce.cu) keep the parameter gradients deterministic.
//
// Precision.  tcgen05 returns zeros for MN-major kind::tf32 operands in the no-swizzle layout (measured with
// tools/umma_probe.cu; kind::f16 MN-major works), so the fp32 contract cannot reuse the forward's 3xTF32 here.  Instead
// every fp32 value is split into THREE bf16 terms, v = b0 + b1 + b2 (exact to 2^-27), and each GEMM issues the six
// products of order <= 2 (b0b0, b0b1, b1b0, b1b1, b0b2, b2b0): error ~2^-26 per product, fp32 accumulation in TMEM --
// tighter than 3xTF32, same MMA count (K = 16 per instruction instead of 8).  bf16 contract: one product.
// CTA = 10 warps, one per SM: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = workers (two threads per row: one
// re-lays the gradient row, the other the x row; they split the epilogue columns).
// =====================================================================================================================
constexpr int kBwdThreads = 320;
constexpr int kBwdWorkers = 256;
constexpr int kBwdMaxStages = 3;

struct TcBwdArgs {
    const void *dxl, *dxr;
    const float *x, *wl, *wr;
    float *dx;        // may be NULL (input does not require grad): GEMM1 is skipped
    int32_t accumulate;  // dx += (the caller pre-loaded dx with the residual branch's gradient) instead of dx =
    float *partials;  // (grid, 2*HC*F + 2*HC): [dWl | dWr | dbl | dbr] per CTA
    int64_t R;
    int32_t F, HC;
    int32_t OP;   // 2*HC padded to 16 (K of GEMM1)
    int32_t N1;   // F padded to 16 (GEMM1 N)
    int32_t NP2;  // F + 1 padded to 16 (GEMM2 N)
    int32_t d1_stride, d2_col, tmem_cols, stages;
};

struct TcBwdSmem {
    uint32_t bars, tmem_ptr, w, d, x, stage0, dx_st, total;  // w/d/x: first split; split q lives at + q * {w,d,x}_bytes
    uint32_t P_w, P_d, P_x, w_bytes, d_bytes, x_bytes, tile_d, tile_x, stage_bytes;
};

// SPLIT = 1: bf16 contract (gradients already bf16);  SPLIT = 3: fp32 contract through the three-term bf16 split
template <int SPLIT>
__host__ __device__ inline TcBwdSmem tc_bwd_smem(int F, int HC, int OP, int N1, int NP2, int stages) {
    const uint32_t esz = SPLIT == 1 ? 2 : 4;  // bytes of a dxl / dxr element in global memory
    TcBwdSmem s;
    uint32_t o = 0;
    s.bars = o; o += 256;
    s.tmem_ptr = o; o += 128;
    s.P_w = (OP / 8) * 128; s.P_d = (OP / 8) * 128; s.P_x = (NP2 / 8) * 128;
    s.w_bytes = (N1 / 8) * s.P_w; s.d_bytes = (kTileM / 8) * s.P_d; s.x_bytes = (kTileM / 8) * s.P_x;
    s.w = o; o += SPLIT * s.w_bytes;
    s.d = o; o += SPLIT * s.d_bytes;
    s.x = o; o += SPLIT * s.x_bytes;          // also absorbs the MN-major over-read of the gradient operand (M = 128 > 2*HC)
    s.tile_d = ((kTileM * HC * esz + 127) / 128) * 128;
    s.tile_x = ((kTileM * F * 4 + 127) / 128) * 128;
    s.stage_bytes = 2 * s.tile_d + s.tile_x;
    s.stage0 = o; o += stages * s.stage_bytes;
    s.dx_st = o; o += SPLIT == 1 ? s.tile_x : 0;  // fp32 contract: dx is stored straight from registers
    s.total = o;
    return s;
}

// One row -> canonical bf16 chunks; SPLIT = 3 writes the three split terms to base, base + stride, base + 2*stride.
template <int SPLIT, typename LoadFn>
__device__ __forceinline__ void write_row_bf16(unsigned char *base, uint32_t split_stride, int r, uint32_t P, int n, int KP,
                                               LoadFn load) {
    for (int kc = 0; kc < KP / 8; ++kc) {
        uint32_t w0[4], w1[4], w2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = kc * 8 + 2 * i;
            const float v0 = k < n ? load(k) : 0.f;
            const float v1 = k + 1 < n ? load(k + 1) : 0.f;
            __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
            w0[i] = *reinterpret_cast<uint32_t *>(&h);
            if (SPLIT == 3) {
                const float2 hf = __bfloat1622float2(h);
                const float r0 = v0 - hf.x, r1 = v1 - hf.y;                 // exact
                __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
                const float2 mf = __bfloat1622float2(m);
                __nv_bfloat162 l = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
                w1[i] = *reinterpret_cast<uint32_t *>(&m);
                w2[i] = *reinterpret_cast<uint32_t *>(&l);
            }
        }
        const uint32_t off = canon_off(r, kc, P);
        *reinterpret_cast<uint4 *>(base + off) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        if (SPLIT == 3) {
            *reinterpret_cast<uint4 *>(base + split_stride + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
            *reinterpret_cast<uint4 *>(base + 2 * split_stride + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        }
    }
}

// the products kept by the split GEMM, smallest first: (a-term, b-term)
__device__ __constant__ int kSplitPairs[6][2] = {{0, 2}, {2, 0}, {1, 1}, {0, 1}, {1, 0}, {0, 0}};

// <= 110 KB of shared memory and 128 TMEM columns per CTA -> TWO CTAs per SM, so one CTA's TMA -> re-layout -> MMA -> epilogue
// latency chain overlaps the other's (bf16 contract: three ring stages; fp32 contract: the three split terms of both operand
// buffers leave room for one stage, and dx leaves straight from registers instead of through a staging tile).
// fp32 contract: GEMM2's TMEM accumulator is flushed into per-thread fp64 sums every kD2FlushTiles tiles (4,096 rows), so the
// weight gradients do not accumulate fp32 rounding over the CTA's whole share of the rows (~60 k).
constexpr int kD2FlushTiles = 32;
template <int SPLIT>
__global__ void __launch_bounds__(kBwdThreads, 2) project_bwd_tc_kernel(const TcBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    using ST = typename std::conditional<SPLIT == 1, __nv_bfloat16, float>::type;
    const int F = a.F, HC = a.HC, O = 2 * a.HC, OP = a.OP, N1 = a.N1, NP2 = a.NP2, nst = a.stages;
    const bool need_dx = a.dx != nullptr;
    const TcBwdSmem L = tc_bwd_smem<SPLIT>(F, HC, OP, N1, NP2, nst);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint64_t *full = bars, *empty = bars + kBwdMaxStages;
    uint64_t *ops_ready = bars + 2 * kBwdMaxStages, *ops_free = ops_ready + 1;
    uint64_t *t_full = ops_ready + 2, *t_empty = ops_ready + 4;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + L.tmem_ptr);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (a.R + kTileM - 1) / kTileM;
    const int n_local = blockIdx.x < num_tiles ? (int)((num_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < kBwdMaxStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kBwdWorkers);
        }
        mbar_init(ops_ready, kBwdWorkers);
        mbar_init(ops_free, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], kBwdWorkers);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, a.tmem_cols);
    // B operand of GEMM1: B1[n = f][k = o] = Wcat[o][f]  (K-major), zero padded
    for (int n = tid; n < N1; n += kBwdThreads) {
        write_row_bf16<SPLIT>(smem + L.w, L.w_bytes, n, L.P_w, n < F ? O : 0, OP, [&](int o) {
            return o < HC ? a.wl[(int64_t)o * F + n] : a.wr[(int64_t)(o - HC) * F + n];
        });
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer: three contiguous bulk copies per tile (dxl, dxr, x) ==================================
        if (lane == 0) {
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int s = it % nst;
                if (it >= nst) mbar_wait(&empty[s], ((it / nst) - 1) & 1);
                const int64_t r0 = tile * kTileM;
                const int nr = (int)((a.R - r0) < (int64_t)kTileM ? (a.R - r0) : (int64_t)kTileM);
                unsigned char *st = smem + L.stage0 + s * L.stage_bytes;
                const uint32_t bd = (uint32_t)nr * HC * (uint32_t)sizeof(ST), bx = (uint32_t)nr * F * 4u;
                const unsigned char *srcs[3] = {reinterpret_cast<const unsigned char *>(static_cast<const ST *>(a.dxl) + r0 * HC),
                                                reinterpret_cast<const unsigned char *>(static_cast<const ST *>(a.dxr) + r0 * HC),
                                                reinterpret_cast<const unsigned char *>(a.x + r0 * F)};
                unsigned char *dsts[3] = {st, st + L.tile_d, st + 2 * L.tile_d};
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
    } else if (warp == 1) {
        // ===== MMA issuer ==============================================================================================
        const uint32_t idesc1 = umma_idesc(kFmtBF16, kTileM, N1, 0, 0);   // A, B K-major
        const uint32_t idesc2 = umma_idesc(kFmtBF16, kTileM, NP2, 1, 1);  // A, B MN-major (the same buffers, transposed view)
        const uint32_t w0 = smem_u32(smem + L.w), d0 = smem_u32(smem + L.d), x0 = smem_u32(smem + L.x);
        const int ksteps1 = OP / 16;      // K = gradient columns, 16 per MMA = two 16-byte chunks
        const int ksteps2 = kTileM / 16;  // K = rows, 16 per MMA = two 8-row blocks
        const int nprod = SPLIT == 3 ? 6 : 1;
        const uint32_t d2 = tmem_base + a.d2_col;
        uint32_t acc2 = 0;
        for (int it = 0; it < n_local; ++it) {
            mbar_wait(ops_ready, it & 1);
            if (need_dx && it >= 2) mbar_wait(&t_empty[it & 1], ((it >> 1) - 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                if (need_dx) {
                    const uint32_t d1 = tmem_base + (uint32_t)(it & 1) * a.d1_stride;
                    uint32_t acc = 0;
                    for (int ks = 0; ks < ksteps1; ++ks) {
                        const uint32_t ko = ks * 256;
                        for (int q = 0; q < nprod; ++q) {
                            const int ia = SPLIT == 3 ? kSplitPairs[q][0] : 0, ib = SPLIT == 3 ? kSplitPairs[q][1] : 0;
                            umma_bf16(d1, umma_desc(d0 + ia * L.d_bytes + ko, 128, L.P_d),
                                      umma_desc(w0 + ib * L.w_bytes + ko, 128, L.P_w), idesc1, acc);
                            acc = 1;
                        }
                    }
                    umma_commit(&t_full[it & 1]);
                }
                if (SPLIT == 3 && NP2 <= 32 && (it % kD2FlushTiles) == 0) acc2 = 0;  // the workers flushed D2 into their fp64 sums
                for (int ks = 0; ks < ksteps2; ++ks) {
                    const uint32_t da = ks * 2 * L.P_d, dbx = ks * 2 * L.P_x;
                    for (int q = 0; q < nprod; ++q) {
                        const int ia = SPLIT == 3 ? kSplitPairs[q][0] : 0, ib = SPLIT == 3 ? kSplitPairs[q][1] : 0;
                        umma_bf16(d2, umma_desc(d0 + ia * L.d_bytes + da, L.P_d, 128),
                                  umma_desc(x0 + ib * L.x_bytes + dbx, L.P_x, 128), idesc2, acc2);
                        acc2 = 1;
                    }
                }
                umma_commit(ops_free);  // both GEMMs have consumed the operand buffers
            }
            __syncwarp();
        }
    } else {
        // ===== workers ===================================================================================================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int wt = (warp - 2) * 32 + lane;  // 0..255
        const bool issuer = (wt == 0);
        float *dx_st = reinterpret_cast<float *>(smem + L.dx_st);
        const bool d2_flush = SPLIT == 3 && NP2 <= 32;  // one 16-column block of D2 per thread
        double acc2d[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc2d[i] = 0.0;
        const uint32_t taddr2 = tmem_base + ((uint32_t)(quarter * 32) << 16) + a.d2_col;
        for (int it = 0; it <= n_local; ++it) {
            if (it < n_local) {
                const int s = it % nst;
                const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
                const int64_t r0 = tile * kTileM;
                const int nr = (int)((a.R - r0) < (int64_t)kTileM ? (a.R - r0) : (int64_t)kTileM);
                mbar_wait(&full[s], (it / nst) & 1);
                if (it >= 1) mbar_wait(ops_free, (it - 1) & 1);
                if (d2_flush && it >= 1 && (it % kD2FlushTiles) == 0) {
                    // every MMA up to tile it-1 has completed and none of tile it can start before all workers arrive below:
                    // D2 is quiescent.  Move it into the fp64 sums; the issuer restarts the accumulator with this tile.
                    tc_fence_after();
                    float v[16];
                    tmem_ld16(taddr2 + half * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc2d[i] += (double)v[i];
                    tc_fence_before();
                }
                const unsigned char *st = smem + L.stage0 + s * L.stage_bytes;
                const bool live = row < nr;  // rows past the end of the last tile must contribute ZERO to GEMM2
                if (SPLIT == 1 && F == 22 && HC == 22) {
                    // the reference's shape, bf16 contract: the gradient rows already are bf16 (11 words each), so the re-layout
                    // is a copy into the 16-byte chunks of the canonical layout; x is converted pairwise.  Straight-line.
                    if (half == 0) {
                        const uint32_t *gl = reinterpret_cast<const uint32_t *>(st + (size_t)row * 44);
                        const uint32_t *gr = reinterpret_cast<const uint32_t *>(st + L.tile_d + (size_t)row * 44);
                        uint32_t w[24];
#pragma unroll
                        for (int i = 0; i < 11; ++i) {
                            w[i] = live ? gl[i] : 0u;
                            w[11 + i] = live ? gr[i] : 0u;
                        }
                        w[22] = w[23] = 0u;
#pragma unroll
                        for (int kc = 0; kc < 6; ++kc)
                            *reinterpret_cast<uint4 *>(smem + L.d + canon_off(row, kc, L.P_d)) = make_uint4(w[4 * kc], w[4 * kc + 1], w[4 * kc + 2], w[4 * kc + 3]);
                    } else {
                        const float2 *xrow = reinterpret_cast<const float2 *>(st + 2 * L.tile_d + (size_t)row * 88);
                        uint32_t w[16];
#pragma unroll
                        for (int i = 0; i < 11; ++i) {
                            const float2 v = xrow[i];
                            const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
                            w[i] = live ? *reinterpret_cast<const uint32_t *>(&h) : 0u;
                        }
                        w[11] = live ? 0x00003F80u : 0u;  // bf16(1.0) in column F: the bias gradients fall out of GEMM2
                        w[12] = w[13] = w[14] = w[15] = 0u;
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc)
                            *reinterpret_cast<uint4 *>(smem + L.x + canon_off(row, kc, L.P_x)) = make_uint4(w[4 * kc], w[4 * kc + 1], w[4 * kc + 2], w[4 * kc + 3]);
                    }
                } else if (SPLIT == 3 && F == 22 && HC == 22) {
                    // the reference's shape, fp32 contract: v = b0 + b1 + b2 in three bf16 terms, straight-line
                    auto split3 = [&](float v0, float v1, uint32_t &o0, uint32_t &o1, uint32_t &o2) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                        const float2 hf = __bfloat1622float2(h);
                        const float r0 = v0 - hf.x, r1 = v1 - hf.y;  // exact
                        const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
                        const float2 mf = __bfloat1622float2(m);
                        const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
                        o0 = *reinterpret_cast<const uint32_t *>(&h);
                        o1 = *reinterpret_cast<const uint32_t *>(&m);
                        o2 = *reinterpret_cast<const uint32_t *>(&l);
                    };
                    if (half == 0) {
                        const float2 *gl = reinterpret_cast<const float2 *>(st + (size_t)row * 88);
                        const float2 *gr = reinterpret_cast<const float2 *>(st + L.tile_d + (size_t)row * 88);
                        uint32_t w0[24], w1[24], w2[24];
#pragma unroll
                        for (int i = 0; i < 22; ++i) {
                            const float2 v = i < 11 ? gl[i] : gr[i - 11];
                            split3(live ? v.x : 0.f, live ? v.y : 0.f, w0[i], w1[i], w2[i]);
                        }
                        w0[22] = w0[23] = w1[22] = w1[23] = w2[22] = w2[23] = 0u;
#pragma unroll
                        for (int kc = 0; kc < 6; ++kc) {
                            const uint32_t off = canon_off(row, kc, L.P_d);
                            *reinterpret_cast<uint4 *>(smem + L.d + off) = make_uint4(w0[4 * kc], w0[4 * kc + 1], w0[4 * kc + 2], w0[4 * kc + 3]);
                            *reinterpret_cast<uint4 *>(smem + L.d + L.d_bytes + off) = make_uint4(w1[4 * kc], w1[4 * kc + 1], w1[4 * kc + 2], w1[4 * kc + 3]);
                            *reinterpret_cast<uint4 *>(smem + L.d + 2 * L.d_bytes + off) = make_uint4(w2[4 * kc], w2[4 * kc + 1], w2[4 * kc + 2], w2[4 * kc + 3]);
                        }
                    } else {
                        const float2 *xrow = reinterpret_cast<const float2 *>(st + 2 * L.tile_d + (size_t)row * 88);
                        uint32_t w0[16], w1[16], w2[16];
#pragma unroll
                        for (int i = 0; i < 11; ++i) {
                            const float2 v = xrow[i];
                            split3(live ? v.x : 0.f, live ? v.y : 0.f, w0[i], w1[i], w2[i]);
                        }
                        w0[11] = live ? 0x00003F80u : 0u;  // the constant-one column: exact in the first term
                        w1[11] = w2[11] = 0u;
#pragma unroll
                        for (int i = 12; i < 16; ++i) w0[i] = w1[i] = w2[i] = 0u;
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc) {
                            const uint32_t off = canon_off(row, kc, L.P_x);
                            *reinterpret_cast<uint4 *>(smem + L.x + off) = make_uint4(w0[4 * kc], w0[4 * kc + 1], w0[4 * kc + 2], w0[4 * kc + 3]);
                            *reinterpret_cast<uint4 *>(smem + L.x + L.x_bytes + off) = make_uint4(w1[4 * kc], w1[4 * kc + 1], w1[4 * kc + 2], w1[4 * kc + 3]);
                            *reinterpret_cast<uint4 *>(smem + L.x + 2 * L.x_bytes + off) = make_uint4(w2[4 * kc], w2[4 * kc + 1], w2[4 * kc + 2], w2[4 * kc + 3]);
                        }
                    }
                } else if (half == 0) {
                    const ST *gl = reinterpret_cast<const ST *>(st) + row * HC;
                    const ST *gr = reinterpret_cast<const ST *>(st + L.tile_d) + row * HC;
                    write_row_bf16<SPLIT>(smem + L.d, L.d_bytes, row, L.P_d, live ? O : 0, OP,
                                          [&](int o) { return o < HC ? ld_elem(gl + o) : ld_elem(gr + (o - HC)); });
                } else {
                    const float *xrow = reinterpret_cast<const float *>(st + 2 * L.tile_d) + row * F;
                    write_row_bf16<SPLIT>(smem + L.x, L.x_bytes, row, L.P_x, live ? F + 1 : 0, NP2,
                                          [&](int f) { return f < F ? xrow[f] : 1.f; });
                }
                fence_proxy_async();
                mbar_arrive(ops_ready);
                mbar_arrive(&empty[s]);
            }
            if (need_dx && it >= 1) {
                const int j = it - 1, acc = j & 1;
                const int64_t tile = blockIdx.x + (int64_t)j * gridDim.x;
                const int64_t r0 = tile * kTileM;
                const int nr = (int)((a.R - r0) < (int64_t)kTileM ? (a.R - r0) : (int64_t)kTileM);
                mbar_wait(&t_full[acc], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * a.d1_stride;
                if (SPLIT == 3) {
                    // fp32 contract: no staging tile (shared memory is spent on the split operands): each thread owns 16 columns of
                    // its row and stores them as aligned float2 (rows are F * 4 bytes, F even on this path)
                    float *grow = a.dx + (r0 + row) * F;
                    for (int cb = half; cb < N1 / 16; cb += 2) {
                        float v[16];
                        tmem_ld16(taddr + cb * 16, v);
                        if (row < nr) {
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                const int f = cb * 16 + i;
                                if (f + 1 < F && (F & 1) == 0) {  // even F: every row starts 8-byte aligned
                                    float2 *dst = reinterpret_cast<float2 *>(grow + f);
                                    float2 o = make_float2(v[i], v[i + 1]);
                                    if (a.accumulate) {
                                        const float2 old = *dst;
                                        o.x += old.x;
                                        o.y += old.y;
                                    }
                                    *dst = o;
                                } else {
                                    if (f < F) grow[f] = a.accumulate ? grow[f] + v[i] : v[i];
                                    if (f + 1 < F) grow[f + 1] = a.accumulate ? grow[f + 1] + v[i + 1] : v[i + 1];
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&t_empty[acc]);
                    continue;
                }
                if (issuer) bulk_wait_read0();
                named_bar_sync(1, kBwdWorkers);
                for (int cb = half; cb < N1 / 16; cb += 2) {
                    float v[16];
                    tmem_ld16(taddr + cb * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int f = cb * 16 + i;
                        if (f < F) dx_st[row * F + f] = v[i];
                    }
                }
                tc_fence_before();
                mbar_arrive(&t_empty[acc]);
                fence_proxy_async();
                named_bar_sync(1, kBwdWorkers);
                float *gdx = a.dx + r0 * F;
                if (nr == kTileM) {
                    if (issuer) {
                        if (a.accumulate) bulk_s2g_add_f32(gdx, dx_st, kTileM * F * 4u);  // one add per element: deterministic
                        else bulk_s2g(gdx, dx_st, kTileM * F * 4u);
                        bulk_commit();
                    }
                } else if (a.accumulate) {
                    for (int i = wt; i < nr * F; i += kBwdWorkers) gdx[i] += dx_st[i];
                } else {
                    for (int i = wt; i < nr * F; i += kBwdWorkers) gdx[i] = dx_st[i];
                }
            }
        }
        // ---- parameter-gradient partials: D2[m = o][n = f | bias] ------------------------------------------------
        if (n_local > 0) {
            mbar_wait(ops_free, (n_local - 1) & 1);  // every MMA of this CTA has completed
            tc_fence_after();
            float *out = a.partials + (int64_t)blockIdx.x * (O * F + O);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a.d2_col;
            for (int cb = half; cb < NP2 / 16; cb += 2) {
                float v[16];
                tmem_ld16(taddr + cb * 16, v);
                if (row < O) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int n = cb * 16 + i;
                        const float t = d2_flush ? (float)(acc2d[i] + (double)v[i]) : v[i];
                        if (n < F) out[row * F + n] = t;
                        else if (n == F) out[O * F + row] = t;
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

template <int SPLIT>
static bool bwd_tc_config(TcBwdArgs &a) {
    a.OP = ((2 * a.HC + 15) / 16) * 16;
    a.N1 = ((a.F + 15) / 16) * 16;
    a.NP2 = ((a.F + 1 + 15) / 16) * 16;
    if (2 * a.HC > 128 || a.N1 > 256 || a.NP2 > 256) return false;
    a.d1_stride = pow2_cols(a.N1);
    a.d2_col = 2 * a.d1_stride;
    a.tmem_cols = pow2_cols(a.d2_col + a.NP2);
    if (a.tmem_cols > 512) return false;
    const uint32_t budget = 110u * 1024u;  // two CTAs per SM
    for (a.stages = kBwdMaxStages; a.stages >= 1; --a.stages)
        if (tc_bwd_smem<SPLIT>(a.F, a.HC, a.OP, a.N1, a.NP2, a.stages).total <= budget) return true;
    return false;
}

bool project_bwd_tc_supported(int F, int HC, int dtype) {
    TcBwdArgs a{};
    a.F = F; a.HC = HC;
    return dtype == TECGAT_BF16 ? bwd_tc_config<1>(a) : bwd_tc_config<3>(a);
}

static int bwd_tc_grid(int64_t R, int ctas_per_sm = 2) {
    const int64_t tiles = (R + kTileM - 1) / kTileM, cap = int64_t(ctas_per_sm) * tg_sm_count();
    return (int)(tiles < cap ? tiles : cap);
}

int64_t project_bwd_tc_workspace(int64_t R, int F, int HC) {
    const int64_t tc = int64_t(bwd_tc_grid(R)) * (2 * HC * F + 2 * HC) * (int64_t)sizeof(float);
    const int64_t ff = project_bwd_ffma_workspace(R, F, HC);
    return tc > ff ? tc : ff;
}

template <int SPLIT>
static int launch_bwd_tc(TcBwdArgs &a, float *dwl, float *dbl, float *dwr, float *dbr, cudaStream_t st, ReduceJob *defer) {
    bwd_tc_config<SPLIT>(a);
    const TcBwdSmem L = tc_bwd_smem<SPLIT>(a.F, a.HC, a.OP, a.N1, a.NP2, a.stages);
    auto kern = project_bwd_tc_kernel<SPLIT>;
    TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)L.total));
    const int grid = bwd_tc_grid(a.R, 2);
    kern<<<grid, kBwdThreads, L.total, st>>>(a); tg_count_launch();
    TG_LAUNCH_CHECK();
    const int O = 2 * a.HC, F = a.F, HC = a.HC;
    ReduceSegs segs = {{dwl, dwr, dbl, dbr}, {0, HC * F, O * F, O * F + HC}, {HC * F, O * F, O * F + HC, O * F + O}};
    if (defer) {  // the caller finishes these partials together with the edge kernel's in one launch
        *defer = ReduceJob{a.partials, grid, O * F + O, segs};
        return TECGAT_OK;
    }
    return reduce_columns(a.partials, grid, O * F + O, segs, st);
}

int project_bwd_tc(const void *dxl, const void *dxr, const float *x, const float *wl, const float *wr, float *dx,
                   float *dwl, float *dbl, float *dwr, float *dbr, void *workspace, int64_t R, int F, int HC, int dtype,
                   cudaStream_t st, bool accumulate, ReduceJob *defer) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dxl) | reinterpret_cast<uintptr_t>(dxr) |
                           reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    if (!aligned || !project_bwd_tc_supported(F, HC, dtype)) {  // shapes beyond the tensor-core kernel's shared-memory budget
        TG_REQUIRE(!accumulate && !defer, TECGAT_ENOSUP, "project_bwd(tc): F=%d, H*C=%d unsupported with accumulation", F, HC);
        return project_bwd_ffma(dxl, dxr, x, wl, wr, dx, dwl, dbl, dwr, dbr, workspace, R, F, HC, dtype, st);
    }
    TcBwdArgs a{};
    a.dxl = dxl; a.dxr = dxr; a.x = x; a.wl = wl; a.wr = wr; a.dx = dx; a.partials = static_cast<float *>(workspace);
    a.accumulate = accumulate ? 1 : 0;
    a.R = R; a.F = F; a.HC = HC;
    return dtype == TECGAT_BF16 ? launch_bwd_tc<1>(a, dwl, dbl, dwr, dbr, st, defer) : launch_bwd_tc<3>(a, dwl, dbl, dwr, dbr, st, defer);
}

}  // namespace tg
