// edge_common.cuh -- machinery shared by the fused edge forward / backward kernels (edge_fwd.cu, edge_bwd.cu).
//
// Execution model (both kernels): PERSISTENT, warp-specialised CTAs, one per SM.  The work list is every
// (snapshot, tile-of-T-destination-nodes) item in snapshot-major order; CTA c owns the contiguous slice
// [c W / G, (c+1) W / G) so consecutive items of a CTA are neighbouring tiles of one snapshot and their overlapping
// row windows are re-read from L2, not HBM.  One PRODUCER warp streams each item's inputs -- the tile's ELL slab
// (plan.cu), the tile's own rows and the contiguous window of neighbour rows -- into a ring of shared-memory stages
// with bulk-TMA copies (cp.async.bulk, mbarrier complete_tx); the CONSUMER warps compute out of shared memory and
// hand the stage back through a second mbarrier.  No __syncthreads after set-up: warps drift freely.
//
// Lanes: a consumer warp owns 32 / Hp consecutive nodes (Hp = heads padded to a power of two); lane = (node, head)
// with the head index in the HIGH lane bits, so that the 8-byte shared-memory loads of a half-warp hit 16 different
// rows of one head: conflict-free at the row stride of 22 floats.  The per-destination softmax needs no cross-lane
// traffic and nothing is atomic.
//
// Arithmetic: a lane's C channels are held as C/2 float2 pairs plus (odd C) one scalar.  Which channels pair up is
// chosen per lane from the PARITY of its chunk's element offset, so every pair is an aligned 8-byte (fp32) / 4-byte
// (bf16) shared-memory load; all math runs position-wise on the packed fp32 pipe (FADD2 / FFMA2).  LeakyReLU is
// folded into the attention dot product:  att . lrelu(s) = att_p . s + att_m . |s|  with att_p = att (1+slope)/2,
// att_m = att (1-slope)/2 (FFMA2 takes |.| as a free source modifier), and log2(e) is folded into att_p / att_m so the
// scores live in the exp2 domain.
#pragma once
#include "common.cuh"

namespace tg {

__host__ __device__ __forceinline__ uint32_t round16(uint32_t b) { return (b + 15u) & ~15u; }

// Channel counts the edge kernels are instantiated for (C = out_channels per head).
#define TG_FOR_EACH_C(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(11) X(12) X(16) X(24) X(32)

constexpr int kMaxStages = 4;
constexpr int kEdgeSmemBudget = 226 * 1024;  // dynamic shared memory per CTA we are willing to request (1 CTA / SM)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kOverflowGuard = 1e24f;  // softmax sums beyond this re-run the row with the exact maximum as shift

// ---- channel vector ----------------------------------------------------------------------------------------------
template <int C>
struct CV {
    static constexpr int NP = C / 2;
    static constexpr bool ODD = (C & 1) != 0;
    float2 p[NP > 0 ? NP : 1];
    float s;
};
template <int C>
__device__ __forceinline__ void cv_zero(CV<C> &v) {
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) v.p[i] = make_float2(0.f, 0.f);
    v.s = 0.f;
}
// position of channel c in the lane's arrangement: par = 0 -> pairs (0,1)(2,3).. single C-1; par = 1 -> single 0, pairs (1,2)(3,4)..
template <int C>
__device__ __forceinline__ int cv_channel_of_pair(int i, int par) { return 2 * i + par; }
template <int C>
__device__ __forceinline__ int cv_channel_of_single(int par) { return par ? 0 : C - 1; }

__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}
__device__ __forceinline__ uint32_t float2_to_bf16x2(float2 v) {
    const __nv_bfloat162 b = __float22bfloat162_rn(v);
    return *reinterpret_cast<const uint32_t *>(&b);
}

// Load the C channels starting at `chunk` (channel 0).  VEC: `chunk + par` is pair-aligned -> vector loads.
template <int C, bool VEC>
__device__ __forceinline__ void cv_load(CV<C> &v, const float *chunk, int par) {
    const float *q = chunk + par;
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        if (VEC) v.p[i] = *reinterpret_cast<const float2 *>(q + 2 * i);
        else v.p[i] = make_float2(q[2 * i], q[2 * i + 1]);
    }
    if (CV<C>::ODD) v.s = chunk[par ? 0 : C - 1];
}
template <int C, bool VEC>
__device__ __forceinline__ void cv_load(CV<C> &v, const __nv_bfloat16 *chunk, int par) {
    const __nv_bfloat16 *q = chunk + par;
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        if (VEC) v.p[i] = bf16x2_to_float2(*reinterpret_cast<const uint32_t *>(q + 2 * i));
        else v.p[i] = make_float2(__bfloat162float(q[2 * i]), __bfloat162float(q[2 * i + 1]));
    }
    if (CV<C>::ODD) v.s = __bfloat162float(chunk[par ? 0 : C - 1]);
}
template <int C, bool VEC>
__device__ __forceinline__ void cv_store(float *chunk, const CV<C> &v, int par) {
    float *q = chunk + par;
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        if (VEC) *reinterpret_cast<float2 *>(q + 2 * i) = v.p[i];
        else { q[2 * i] = v.p[i].x; q[2 * i + 1] = v.p[i].y; }
    }
    if (CV<C>::ODD) chunk[par ? 0 : C - 1] = v.s;
}
template <int C, bool VEC>
__device__ __forceinline__ void cv_store(__nv_bfloat16 *chunk, const CV<C> &v, int par) {
    __nv_bfloat16 *q = chunk + par;
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        if (VEC) *reinterpret_cast<uint32_t *>(q + 2 * i) = float2_to_bf16x2(v.p[i]);
        else { q[2 * i] = __float2bfloat16_rn(v.p[i].x); q[2 * i + 1] = __float2bfloat16_rn(v.p[i].y); }
    }
    if (CV<C>::ODD) chunk[par ? 0 : C - 1] = __float2bfloat16_rn(v.s);
}
// parameters (att, bias): element-wise loads in the lane's arrangement, scaled
template <int C>
__device__ __forceinline__ void cv_load_param(CV<C> &v, const float *p, int par, float scale) {
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) v.p[i] = make_float2(p[2 * i + par] * scale, p[2 * i + 1 + par] * scale);
    v.s = CV<C>::ODD ? p[par ? 0 : C - 1] * scale : 0.f;
}
template <int C>
__device__ __forceinline__ float cv_dot(const CV<C> &a, const CV<C> &b) {
    float2 d = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) d = __ffma2_rn(a.p[i], b.p[i], d);
    float r = d.x + d.y;
    if (CV<C>::ODD) r = fmaf(a.s, b.s, r);
    return r;
}
__device__ __forceinline__ float2 abs2(float2 v) { return make_float2(fabsf(v.x), fabsf(v.y)); }
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

// score of one edge in the exp2 domain, WITHOUT the destination constant att_p . xr_i:
//     att_p . xj + att_m . |xj + xr|        (s = xj + xr is returned for the backward)
template <int C>
__device__ __forceinline__ float edge_score(const CV<C> &attp, const CV<C> &attm, const CV<C> &xj, const CV<C> &xr, CV<C> &s) {
    float2 el = make_float2(0.f, 0.f), ea = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        s.p[i] = __fadd2_rn(xj.p[i], xr.p[i]);
        el = __ffma2_rn(attp.p[i], xj.p[i], el);
        ea = __ffma2_rn(attm.p[i], abs2(s.p[i]), ea);
    }
    float e = (el.x + ea.x) + (el.y + ea.y);
    if (CV<C>::ODD) {
        s.s = xj.s + xr.s;
        e += fmaf(attm.s, fabsf(s.s), attp.s * xj.s);
    }
    return e;
}

// 2^x, bare MUFU.EX2 with flush-to-zero (softmax weights; no denormal rescue sequence)
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_log2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- lane <-> (node, head) ----------------------------------------------------------------------------------------
__host__ __device__ constexpr int pad_heads(int H) {  // next power of two (<= 32)
    int hp = 1;
    while (hp < H) hp <<= 1;
    return hp;
}

// ---- item schedule ---------------------------------------------------------------------------------------------
struct ItemRange {
    int64_t w0, w1;
};
__device__ __forceinline__ ItemRange cta_items(int64_t items) {
    ItemRange r;
    r.w0 = items * blockIdx.x / gridDim.x;
    r.w1 = items * (blockIdx.x + 1) / gridDim.x;
    return r;
}
// Contiguous item ranges of equal estimated work: computed on the host (tg_item_bounds, plan.cu) and passed INSIDE the kernel
// parameters -- read from the constant bank with a uniform index, so the CTA's (snapshot, tile, count) bookkeeping stays in
// uniform registers exactly as with the closed-form split (loading the two bounds from global memory instead cost 2.5 % of
// edge_bwd: the loop counters became vector registers).  Grids beyond kMaxSchedGrid use equal item counts.
constexpr int kMaxSchedGrid = 191;
struct ItemSchedule {
    int32_t use;                        // 0: equal item counts (closed form)
    int32_t pad;
    int64_t bounds[kMaxSchedGrid + 1];  // CTA b: items [bounds[b], bounds[b + 1])
};
__device__ __forceinline__ ItemRange cta_items_scheduled(const ItemSchedule &s, int64_t items) {
    if (!s.use) return cta_items(items);
    ItemRange r;
    r.w0 = s.bounds[blockIdx.x];
    r.w1 = s.bounds[blockIdx.x + 1];
    return r;
}
inline void fill_schedule(ItemSchedule &s, const tecgat_plan_t *plan, bool bwd, int64_t snapshots, int grid) {
    s.use = 0;
    s.pad = 0;
    const char *knob = tg_env("TECGAT_SCHEDULE");  // A/B knob: "closed" = the closed-form equal-count split computed in the kernel
    if (grid > kMaxSchedGrid || (knob && knob[0] == 'c' && knob[1] == 'l')) return;
    const int64_t *b = tg_item_bounds(plan, bwd, snapshots, grid, nullptr);
    for (int i = 0; i <= grid; ++i) s.bounds[i] = b[i];
    s.use = 1;
}

// ring position: stage index and mbarrier phase parity, advanced without divisions
struct Ring {
    int st;
    uint32_t ph;
    __device__ __forceinline__ void advance(int num_stages) {
        if (++st == num_stages) {
            st = 0;
            ph ^= 1u;
        }
    }
};

// tile table: shared memory when it fits (kMetaSmemTiles), global memory (L1 / L2) otherwise
constexpr int kMetaSmemTiles = 64;
__device__ __forceinline__ tg_tile_meta load_meta(const tg_tile_meta *smem_tab, const tg_tile_meta *gmem_tab, int num_tiles, int tile) {
    int4 a, b;  // two explicit paths: a pointer select would turn both into generic-address loads
    if (num_tiles <= kMetaSmemTiles) {
        a = reinterpret_cast<const int4 *>(smem_tab + tile)[0];
        b = reinterpret_cast<const int4 *>(smem_tab + tile)[1];
    } else {
        a = __ldg(reinterpret_cast<const int4 *>(gmem_tab + tile));
        b = __ldg(reinterpret_cast<const int4 *>(gmem_tab + tile) + 1);
    }
    tg_tile_meta m;
    m.lo = a.x; m.hi = a.y; m.kin_kout = a.z; m.eligible = a.w;
    m.slab_off = (int64_t)(uint32_t)b.x | ((int64_t)b.y << 32);
    m.slab_bytes = b.z; m.pad = b.w;
    return m;
}

// Copy of the rows [r0, r0 + nrows) of a row-major array (row = RB bytes, 16-byte aligned base, Rtot rows) as ONE
// bulk-TMA transfer: the range is widened to the array's 16-byte row period `per` (a power of two: 16 / gcd(16, RB)),
// clamped to the array.  Row r0 then sits `skip` bytes into the shared region.  `tail` (< 16 bytes) is non-zero only
// when the clamped end of the array is not 16-byte aligned (last rows of the last snapshot): plain copies.
struct WinCopy {
    const unsigned char *src;
    uint32_t mid, tail, skip;
};
__device__ __forceinline__ WinCopy win_copy(const void *base, int64_t r0, int nrows, uint32_t RB, int per, int64_t Rtot) {
    const int64_t ra = r0 & ~(int64_t)(per - 1);
    int64_t rb = (r0 + nrows + per - 1) & ~(int64_t)(per - 1);
    if (rb > Rtot) rb = Rtot;
    const uint32_t bytes = (uint32_t)(rb - ra) * RB;
    WinCopy c;
    c.src = static_cast<const unsigned char *>(base) + ra * RB;
    c.mid = bytes & ~15u;
    c.tail = bytes - c.mid;
    c.skip = (uint32_t)(r0 - ra) * RB;
    return c;
}
__device__ __forceinline__ uint32_t win_skip(int64_t r0, uint32_t RB, int per) { return (uint32_t)(r0 & (int64_t)(per - 1)) * RB; }
__device__ __forceinline__ void win_copy_tail(const WinCopy &c, unsigned char *dst, int lane) {  // rare (see WinCopy)
    for (uint32_t b = 2u * lane; b < c.tail; b += 64u)
        *reinterpret_cast<uint16_t *>(dst + c.mid + b) = *reinterpret_cast<const uint16_t *>(c.src + c.mid + b);
}
__host__ __device__ __forceinline__ int row_period(uint32_t RB) {  // rows after which the byte offset is 16-byte aligned again
    int per = 1;
    while ((uint64_t(per) * RB) % 16u) per <<= 1;
    return per;
}

// ---- pieces shared by the two backward kernels (edge_bwd.cu, edge_bwd_sw.cu) ---------------------------------------------
constexpr int kBarConsumers = 1;  // named barrier id used by the consumer warps
constexpr int kFlushItems = 256;   // items between two flushes of the per-lane fp32 parameter-gradient sums

__device__ __forceinline__ void bar_sync_named(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// sgn-weighted accumulate:  B[c] += de * [s[c] > 0]
template <int C>
__device__ __forceinline__ void acc_step(CV<C> &B, const CV<C> &s, float de) {
    const float2 de2 = splat(de);
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        // [s > 0] as saturate(s * 1e38): FMUL.SAT on the FMA pipe instead of FSET on the (busier per instruction) ALU pipe --
        // 0.9 % of edge_bwd; exact for |s| >= 1e-38, i.e. everywhere but inside the kink
        const float2 step = make_float2(__saturatef(s.p[i].x * 1.0e38f), __saturatef(s.p[i].y * 1.0e38f));
        B.p[i] = __ffma2_rn(de2, step, B.p[i]);
    }
    if (CV<C>::ODD) B.s = fmaf(de, s.s > 0.f ? 1.f : 0.f, B.s);  // (measured: the compare is the faster form for the odd channel)
}
template <int C>
__device__ __forceinline__ void cv_axpy(CV<C> &y, float a, const CV<C> &x) {
    const float2 a2 = splat(a);
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) y.p[i] = __ffma2_rn(a2, x.p[i], y.p[i]);
    if (CV<C>::ODD) y.s = fmaf(a, x.s, y.s);
}

template <bool DROP>  // DROP = false: inference / p = 0 instantiation without the hash
struct DropCfg {
    uint32_t thr, key;
    float inv_keep;
    __device__ __forceinline__ float q(uint32_t slot) const { return qh(slot * kDropMul + key); }
    __device__ __forceinline__ float qh(uint32_t h) const {  // h = slot * kDropMul + key (consecutive slots: one add)
        if (!DROP) return 1.f;
        // no branch on "dropout off": threshold 0 keeps every edge and inv_keep is 1 (the loop bodies stay one basic block)
        return dropout_finish(h) >= thr ? inv_keep : 0.f;
    }
};

// internal entry points behind the C ABI (seed_dev != NULL: the dropout seed is read from device memory)
int edge_fwd_run(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att, const float *bias, float *y,
                 float *stat, int32_t snapshots, int32_t heads, int32_t out_channels, float negative_slope, float dropout_p,
                 uint64_t seed, const uint64_t *seed_dev, int32_t mode, int32_t dtype, void *stream, float *y_wide = nullptr,
                 int64_t ld_wide = 0);
// reduce = false: leave the per-CTA partial rows of d att / d bias in `workspace` (*partial_rows of them, width 2*H*C)
// for the caller's own second stage
int edge_bwd_run(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att, const float *bias, const float *y,
                 const float *stat, const float *gy, void *dxl, void *dxr, float *datt, float *dbias, void *workspace,
                 int32_t snapshots, int32_t heads, int32_t out_channels, float negative_slope, float dropout_p, uint64_t seed,
                 const uint64_t *seed_dev, int32_t mode, int32_t dtype, void *stream, bool reduce, int64_t *partial_rows);

// sliding-window backward (edge_bwd_sw.cu): *used = false (and TECGAT_OK) when the plan / shape does not qualify
int edge_bwd_sw_try(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att, const float *bias, const float *y,
                    const float *stat, const float *gy, void *dxl, void *dxr, float *partials, int grid, int max_flushes,
                    int32_t snapshots, int32_t heads, int32_t out_channels, float negative_slope, float dropout_p, uint32_t drop_thr,
                    uint64_t seed, const uint64_t *seed_dev, int32_t mode, int32_t dtype, cudaStream_t st, bool *used);

}  // namespace tg
