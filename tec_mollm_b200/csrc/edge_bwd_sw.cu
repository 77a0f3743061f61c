// edge_bwd_sw.cu -- fused GATv2 edge phase, backward, SLIDING-WINDOW formulation for banded graphs (the TEC grids: every edge
// joins nodes at most R <= T rows apart).  Same mathematics as edge_bwd.cu (header there), but every edge's score is evaluated
// ONCE:
//   D role (node v as destination): walks the in-edges (u -> v) exactly like edge_bwd.cu, produces d xr_v, and STASHES the two
//     per-edge scalars (alpha q, d e) in a shared-memory ring indexed by (destination, in-slot);
//   S role (node v as source): walks the out-edges (v -> u) and only reads the stashed pair, xr_u (for the LeakyReLU branch)
//     and g_u:   d xl_v = sum_u (alpha q)_vu g_u + att . [slope A + (1 - slope) sum_u de_vu [xl_v + xr_u > 0]]
//     -- no score, no exp2, no dropout hash, no g.xl dot product, no delta / stat lookup (about half of the source role's
//     instructions in edge_bwd.cu).
// A CTA walks its contiguous range of chunks (T = 128 nodes) of a snapshot in order; phase k runs D(chunk k) and S(chunk k - 2),
// so S always finds the stash of its destinations (chunks k-3 .. k-1) complete; ONE CTA-wide mbarrier per phase.  Rows are
// staged ONCE each (edge_bwd.cu stages every row 2.1 times): two producer warps stream T-row blocks through circular row
// buffers -- xl for the D role's neighbour window, xr and g for the S role's -- with bulk-TMA copies; own rows (xl_v, xr_v,
// g_v, y_v, stat_v) are plain coalesced global loads.  Chunk ranges that start or end inside a snapshot run the D role of
// the neighbouring chunk as a halo (stash only).  No atomics; d att / d bias as in edge_bwd.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "edge_common.cuh"
#include "reduce.cuh"

namespace tg {

struct SwArgs {
    const void *xl, *xr;
    const float *att, *bias, *y, *stat, *gy;
    void *dxl, *dxr;
    float *partials;
    const unsigned char *slabD, *slabS;
    int32_t N, J, S, R;
    int32_t kinp, koutp;
    int32_t slabD_bytes, slabS_bytes, stash_stride;
    float slope, inv_keep;
    uint32_t drop_thr;
    uint32_t stream_stride;  // dropout counter distance between consecutive (snapshot, head) streams
    uint64_t seed;
    const uint64_t *seed_dev;
    int32_t P, Ps;  // ring capacities: rows (xl / xr / g), nodes (stash)
    int32_t per_st, per_f, per_max;
    uint32_t off_red, off_out, off_slabD, off_slabS, off_xl, off_xr, off_g, off_stash;
    int32_t max_flushes;
    int64_t chunks;  // J * S
};

// 384 threads = 8 consumer warps (two warpgroups) + one producer warpgroup (warps 8 and 9 work, 10 and 11 only donate their
// registers): setmaxnreg moves the consumers to 224 registers (a 10-warp CTA would be capped at 168 by the per-scheduler split
// of the register file)
constexpr int kSwT = 128, kSwWarps = 8, kSwThreads = 384, kSwConsumerRegs = 224, kSwProducerRegs = 56;

struct SwSeg {
    int snap, ca, cb, dfirst, dlast, phases;
};
__device__ __forceinline__ bool sw_next_seg(int64_t &g, int64_t g1, int J, SwSeg &s) {
    if (g >= g1) return false;
    s.snap = (int)(g / J);
    s.ca = (int)(g % J);
    const int64_t left = g1 - g;
    s.cb = (int)min((int64_t)(J - 1), (int64_t)s.ca + left - 1);
    s.dfirst = max(0, s.ca - 1);
    s.dlast = min(J - 1, s.cb + 1);
    s.phases = s.cb - s.dfirst + 3;  // D(dfirst + k) for k <= dlast - dfirst, S(dfirst + k - 2); the last S is S(cb)
    g += s.cb - s.ca + 1;
    return true;
}
__device__ __forceinline__ int sw_wlo(int c, int R) { return max(0, c * kSwT - R); }
__device__ __forceinline__ int sw_whi(int c, int R, int N) { return min(N, (c + 1) * kSwT + R); }

// Row rings: node n of the current segment lives at ring row (vb + n) mod P.  vb is chosen so that the segment's first row
// follows the rows loaded so far (V) and ring rows keep the 16-byte phase of the global rows (vb = snap * N mod per_max).
__device__ __forceinline__ int sw_ring_base(int V, int n0, int snap, int N, int P, int pm) {
    int t = (V - n0) % P;
    if (t < 0) t += P;
    int want = (int)(((int64_t)snap * N) % pm);
    int adj = (want - t) % pm;
    if (adj < 0) adj += pm;
    return (t + adj) % P;
}

// One array's rows [na, nb) of the snapshot whose first global row is row0 -> ring rows starting at pos_na (the ring row of
// node na): widened to the array's 16-byte row period (ring rows keep the global rows' phase), split at the ring's end.
struct SwCopy {
    const unsigned char *src;
    uint32_t pos, first, rest;  // ring row of the first copied row; rows before / after the ring's wrap-around
};
__device__ __forceinline__ SwCopy sw_plan_copy(const void *base, uint32_t RB, int per, int64_t row0, int64_t Rtot, int na, int nb, int pos_na,
                                               int P) {
    const int64_t g0 = row0 + na, g1 = row0 + nb;
    const int64_t ga = g0 & ~(int64_t)(per - 1);
    int64_t gb = (g1 + per - 1) & ~(int64_t)(per - 1);
    if (gb > Rtot) gb = Rtot;
    int pos = pos_na - (int)(g0 - ga);
    if (pos < 0) pos += P;
    const int rows = (int)(gb - ga);
    SwCopy c;
    c.src = static_cast<const unsigned char *>(base) + ga * RB;
    c.pos = (uint32_t)pos;
    c.first = (uint32_t)min(rows, P - pos);
    c.rest = (uint32_t)rows - c.first;
    return c;
}
__device__ __forceinline__ uint32_t sw_copy_bytes(const SwCopy &c, uint32_t RB) { return ((c.first * RB) & ~15u) + ((c.rest * RB) & ~15u); }
// the (< 16 byte) ragged tails with plain stores -- only the last rows of the whole array can have one; done BEFORE the barrier is
// armed so that the arrive orders them
__device__ __forceinline__ void sw_copy_tails(const SwCopy &c, unsigned char *ring, uint32_t RB, int lane) {
    const uint32_t b1 = c.first * RB, m1 = b1 & ~15u, b2 = c.rest * RB, m2 = b2 & ~15u;
    for (uint32_t b = m1 + 2u * lane; b < b1; b += 64u)
        *reinterpret_cast<uint16_t *>(ring + (size_t)c.pos * RB + b) = *reinterpret_cast<const uint16_t *>(c.src + b);
    for (uint32_t b = m2 + 2u * lane; b < b2; b += 64u)
        *reinterpret_cast<uint16_t *>(ring + b) = *reinterpret_cast<const uint16_t *>(c.src + b1 + b);
}
__device__ __forceinline__ void sw_copy_issue(const SwCopy &c, unsigned char *ring, uint32_t RB, uint64_t *bar) {
    const uint32_t b1 = c.first * RB, m1 = b1 & ~15u, m2 = (c.rest * RB) & ~15u;
    if (m1) bulk_g2s(ring + (size_t)c.pos * RB, c.src, m1, bar);
    if (m2) bulk_g2s(ring, c.src + b1, m2, bar);
}

// B[c] += de * [x[c] > thr[c]]   (the S role: [xl_v + xr_u > 0] = [xr_u > -xl_v], no addition needed)
template <int C>
__device__ __forceinline__ void acc_step_gt(CV<C> &B, const CV<C> &x, const CV<C> &thr, float de) {
    const float2 de2 = splat(de);
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        const float2 step = make_float2(x.p[i].x > thr.p[i].x ? 1.f : 0.f, x.p[i].y > thr.p[i].y ? 1.f : 0.f);
        B.p[i] = __ffma2_rn(de2, step, B.p[i]);
    }
    if (CV<C>::ODD) B.s = fmaf(de, x.s > thr.s ? 1.f : 0.f, B.s);
}

template <int C, typename ST, bool DROP>
__global__ void __launch_bounds__(kSwThreads, 1) edge_bwd_sw_kernel(const SwArgs a) {
    constexpr int H = 2, HC = 2 * C, NPW = 16, T = kSwT;
    constexpr uint32_t RB_ST = (uint32_t)HC * sizeof(ST), RB_F = (uint32_t)HC * 4u;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *lfull = reinterpret_cast<uint64_t *>(smem);  // [4] xl block + slabD landed
    uint64_t *rfull = lfull + 4;                            // [4] xr + g block + slabS landed
    // [16] every consumer warp finished phase G: the consumers arrive and wait on phase[G & 15]; the producers observe EVERY phase
    // in order (a parity wait is only meaningful within one reuse of the barrier, and a producer can run many phases ahead of the
    // consumers -- the first segment needs no waits -- or, by the data dependencies, at most ~4 behind)
    uint64_t *phase = rfull + 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N, J = a.J, R = a.R, P = a.P, Ps = a.Ps;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&lfull[i], 1);
            mbar_init(&rfull[i], 1);
        }
        for (int i = 0; i < 16; ++i) mbar_init(&phase[i], kSwWarps);
        fence_mbar_init();
    }
    __syncthreads();
    const int64_t g_begin = a.chunks * blockIdx.x / gridDim.x, g_end = a.chunks * (blockIdx.x + 1) / gridDim.x;
    const int64_t Rtot = (int64_t)a.S * N;
    int seen = 0;  // producers: phases observed so far
    auto wait_phase = [&](int G) {  // until phase G has completed (every consumer warp is past its reads and stash writes)
        while (seen <= G) {
            // try_wait's suspend hint returns within ~20 ns here, so a bare poll loop burns ~150 iterations per phase of the
            // issue slots the consumers need: sleep between polls (the producers run a phase ahead; latency does not matter)
            while (!mbar_try_wait(&phase[seen & 15], (uint32_t)(seen >> 4) & 1u)) __nanosleep(256);
            ++seen;
        }
    };

    if (warp >= kSwWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kSwProducerRegs));
        if (warp > kSwWarps + 1) return;
    }
    if (warp == kSwWarps) {
        // ================================ producer L: xl blocks + slabD for the D role ================================
        int64_t g = g_begin;
        SwSeg sg;
        int V = 0, q = 0, Gbase = 0, prev_last_d = -1;
        unsigned char *ring = smem + a.off_xl;
        while (sw_next_seg(g, g_end, J, sg)) {
            const int n0 = sw_wlo(sg.dfirst, R);
            const int vb = sw_ring_base(V, n0, sg.snap, N, P, a.per_max);
            const int64_t row0 = (int64_t)sg.snap * N;
            int pos = (vb + n0) % P, na = n0;
            for (int k = 0; k <= sg.dlast - sg.dfirst; ++k, ++q) {
                const int d = sg.dfirst + k;
                const int nb = sw_whi(d, R, N);
                // ring space and the slab buffer of block q - 2: free once the D role two blocks back has finished
                if (k >= 2) wait_phase(Gbase + k - 2);
                else if (prev_last_d >= 0) wait_phase(prev_last_d);
                uint64_t *bar = &lfull[q & 3];
                const SwCopy c = sw_plan_copy(a.xl, RB_ST, a.per_st, row0, Rtot, na, nb, pos, P);
                sw_copy_tails(c, ring, RB_ST, lane);
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(bar, (uint32_t)a.slabD_bytes + sw_copy_bytes(c, RB_ST));
                    bulk_g2s(smem + a.off_slabD + (size_t)(q & 1) * a.slabD_bytes, a.slabD + (size_t)d * a.slabD_bytes, (uint32_t)a.slabD_bytes, bar);
                    sw_copy_issue(c, ring, RB_ST, bar);
                    // the D role reads its OWN xr / g / y / stat rows with plain loads one phase from now: pull them into L2
                    const int64_t r0 = row0 + (int64_t)d * T, r1 = row0 + min(N, (d + 1) * T);
                    auto pf = [&](const void *base, uint32_t RB) {
                        const uint64_t b0 = ((uint64_t)r0 * RB + 15u) & ~(uint64_t)15u, b1 = ((uint64_t)r1 * RB) & ~(uint64_t)15u;
                        if (b1 > b0) bulk_prefetch_l2(static_cast<const unsigned char *>(base) + b0, (uint32_t)(b1 - b0));
                    };
                    pf(a.xr, RB_ST);
                    pf(a.gy, RB_F);
                    pf(a.y, RB_F);
                    pf(a.stat, (uint32_t)H * 4u);
                }
                pos += nb - na;
                if (pos >= P) pos -= P;
                na = nb;
            }
            V = (vb + sw_whi(sg.dlast, R, N)) % P;
            prev_last_d = Gbase + (sg.dlast - sg.dfirst);
            Gbase += sg.phases;
        }
        return;
    }
    if (warp == kSwWarps + 1) {
        // ================================ producer R: xr + g blocks + slabS for the S role =============================
        int64_t g = g_begin;
        SwSeg sg;
        int V = 0, q = 0, Gbase = 0;
        unsigned char *ring_r = smem + a.off_xr, *ring_g = smem + a.off_g;
        while (sw_next_seg(g, g_end, J, sg)) {
            const int n0 = sw_wlo(sg.ca, R);
            const int vb = sw_ring_base(V, n0, sg.snap, N, P, a.per_max);
            const int64_t row0 = (int64_t)sg.snap * N;
            int pos = (vb + n0) % P, na = n0;
            for (int j = 0; j <= sg.cb - sg.ca; ++j, ++q) {
                const int s = sg.ca + j;
                const int nb = sw_whi(s, R, N);
                const int ks = s - sg.dfirst + 2;  // the phase that runs S(s)
                if (j >= 2) wait_phase(Gbase + ks - 2);
                else if (Gbase > 0) wait_phase(Gbase - 1);  // the previous segment's last phase (its last S)
                uint64_t *bar = &rfull[q & 3];
                const SwCopy cr = sw_plan_copy(a.xr, RB_ST, a.per_st, row0, Rtot, na, nb, pos, P);
                const SwCopy cg = sw_plan_copy(a.gy, RB_F, a.per_f, row0, Rtot, na, nb, pos, P);
                sw_copy_tails(cr, ring_r, RB_ST, lane);
                sw_copy_tails(cg, ring_g, RB_F, lane);
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(bar, (uint32_t)a.slabS_bytes + sw_copy_bytes(cr, RB_ST) + sw_copy_bytes(cg, RB_F));
                    bulk_g2s(smem + a.off_slabS + (size_t)(q & 1) * a.slabS_bytes, a.slabS + (size_t)s * a.slabS_bytes, (uint32_t)a.slabS_bytes, bar);
                    sw_copy_issue(cr, ring_r, RB_ST, bar);
                    sw_copy_issue(cg, ring_g, RB_F, bar);
                }
                pos += nb - na;
                if (pos >= P) pos -= P;
                na = nb;
            }
            V = (vb + sw_whi(sg.cb, R, N)) % P;
            Gbase += sg.phases;
        }
        return;
    }

    // ================================ consumer warps ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSwConsumerRegs));
    const int nw = lane & (NPW - 1), h = lane >> 4;
    const int node_l = warp * NPW + nw;
    const int par = (h * C) & 1;
    const int ctid = threadIdx.x;
    constexpr int nct = kSwWarps * 32;
    CV<C> attp, attm;
    cv_load_param<C>(attp, a.att + h * C, par, 0.5f * (1.f + a.slope) * kLog2e);
    cv_load_param<C>(attm, a.att + h * C, par, 0.5f * (1.f - a.slope) * kLog2e);
    // att and -bias are only needed once per role: kept in shared memory (the flush scratch's neighbour), not in 22 registers
    float *prm = reinterpret_cast<float *>(smem + a.off_red) + kSwWarps * H * 2 * C;  // [att (HC) | -bias (HC)]
    if (ctid < HC) {
        prm[ctid] = a.att[ctid];
        prm[HC + ctid] = -a.bias[ctid];
    }
    bar_sync_named(kBarConsumers, nct);
    const float *att_s = prm + h * C, *nbias_s = prm + HC + h * C;
    const uint32_t drop_base = dropout_base((DROP && a.seed_dev) ? __ldg(a.seed_dev) : a.seed);
    CV<C> acc_att, acc_bias;
    cv_zero(acc_att);
    cv_zero(acc_bias);
    float *red = reinterpret_cast<float *>(smem + a.off_red);
    int since_flush = 0, flushes = 0;
    auto flush = [&]() {
        auto put = [&](int which, int c, float v) {
            for (int off = NPW >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
            if (nw == 0) red[((warp * H + h) * 2 + which) * C + c] = v;
        };
#pragma unroll
        for (int i = 0; i < CV<C>::NP; ++i) {
            put(0, 2 * i + par, acc_att.p[i].x);
            put(0, 2 * i + 1 + par, acc_att.p[i].y);
            put(1, 2 * i + par, acc_bias.p[i].x);
            put(1, 2 * i + 1 + par, acc_bias.p[i].y);
        }
        if (CV<C>::ODD) {
            put(0, par ? 0 : C - 1, acc_att.s);
            put(1, par ? 0 : C - 1, acc_bias.s);
        }
        bar_sync_named(kBarConsumers, nct);
        if (ctid < 2 * HC) {
            const int which = ctid / HC, j = ctid - which * HC, hd = j / C, c = j - hd * C;
            double v = 0.0;
            for (int wq = 0; wq < kSwWarps; ++wq) v += (double)red[((wq * H + hd) * 2 + which) * C + c];
            a.partials[((int64_t)blockIdx.x * a.max_flushes + flushes) * 2 * HC + which * HC + j] = (float)v;
        }
        bar_sync_named(kBarConsumers, nct);
        cv_zero(acc_att);
        cv_zero(acc_bias);
        ++flushes;
        since_flush = 0;
    };
    ST *out_w = reinterpret_cast<ST *>(smem + a.off_out) + warp * NPW * HC;  // the warp's output rows (d xl, then d xr)
    auto store_rows = [&](const CV<C> &v, bool active, ST *dst_g, int nv) {
        if (active) cv_store<C, true>(out_w + nw * HC + h * C, v, par);
        __syncwarp();
        using W = typename std::conditional<sizeof(ST) == 4, uint2, uint32_t>::type;
        const W *src = reinterpret_cast<const W *>(out_w);
        constexpr int kWords = NPW * HC / 2;
        if (nv == NPW) {
#pragma unroll
            for (int i = 0; i < (kWords + 31) / 32; ++i)
                if (i * 32 + lane < kWords) reinterpret_cast<W *>(dst_g)[i * 32 + lane] = src[i * 32 + lane];
        } else {
            for (int i = lane; i < nv * HC / 2; i += 32) reinterpret_cast<W *>(dst_g)[i] = src[i];
        }
        __syncwarp();
    };
    const float slope = a.slope;
    const uint32_t SS = (uint32_t)a.stash_stride;
    unsigned char *const stash = smem + a.off_stash;
    const ST *const xl_g = static_cast<const ST *>(a.xl) + h * C;
    const ST *const xr_g = static_cast<const ST *>(a.xr) + h * C;
    const float *const gy_g = a.gy + h * C, *const y_g = a.y + h * C, *const stat_g = a.stat + h;
    const ST *const ringL = reinterpret_cast<const ST *>(smem + a.off_xl) + h * C;
    const ST *const ringR = reinterpret_cast<const ST *>(smem + a.off_xr) + h * C;
    const float *const ringG = reinterpret_cast<const float *>(smem + a.off_g) + h * C;
    const int kinp = a.kinp, koutp = a.koutp;

    int64_t g = g_begin;
    SwSeg sg;
    int VL = 0, VR = 0, SV = 0, qL = 0, qR = 0, G = 0;
    while (sw_next_seg(g, g_end, J, sg)) {
        const int vbL = sw_ring_base(VL, sw_wlo(sg.dfirst, R), sg.snap, N, P, a.per_max);
        const int vbR = sw_ring_base(VR, sw_wlo(sg.ca, R), sg.snap, N, P, a.per_max);
        int svb = (SV - sg.dfirst * T) % Ps;
        if (svb < 0) svb += Ps;
        // running ring positions (one modulo per segment, additions with a conditional wrap per phase):
        int posD = (vbL + sw_wlo(sg.dfirst, R)) % P;   // ring row of the window start of the next D chunk
        int stD = (svb + sg.dfirst * T) % Ps;          // stash row of the first node of the next D chunk
        int posS = (vbR + sw_wlo(sg.ca, R)) % P;       // ring row of the window start of the next S chunk
        int stS = (svb + sw_wlo(sg.ca, R)) % Ps;       // stash row of that window start
        DropCfg<DROP> drop;
        drop.thr = a.drop_thr;
        drop.key = dropout_key(drop_base, (uint32_t)sg.snap, (uint32_t)H, (uint32_t)h, a.stream_stride);
        drop.inv_keep = a.inv_keep;
        const int64_t snap0 = (int64_t)sg.snap * N;
        CV<C> xl_next;
        cv_zero(xl_next);
        for (int k = 0; k < sg.phases; ++k, ++G) {
            const int d = sg.dfirst + k, s = d - 2;
            const bool doD = d <= sg.dlast, doS = s >= sg.ca && s <= sg.cb;
            // Own rows are plain global loads with a whole role between request and use: the D role's xr / g / y / stat rows
            // are requested here and used after the S role ran; the S role's xl row was requested one phase ago.
            CV<C> xr_v, g_v, y_v;
            float st_v = 0.f;
#ifndef TG_SW_FAKE_OWN
            if (doD) {
                const int n0 = d * T;
                const int64_t row = snap0 + n0 + (n0 + node_l < N ? node_l : 0);
                cv_load<C, true>(xr_v, xr_g + row * HC, par);
                cv_load<C, true>(g_v, gy_g + row * HC, par);
                cv_load<C, true>(y_v, y_g + row * HC, par);
                st_v = stat_g[row * H];
            }
            const CV<C> xl_s = xl_next;
            if (s + 1 >= sg.ca && s + 1 <= sg.cb) {  // next phase's S role
                const int n0 = (s + 1) * T;
                const int64_t row = snap0 + n0 + (n0 + node_l < N ? node_l : 0);
                cv_load<C, true>(xl_next, xl_g + row * HC, par);
            }
#else
            CV<C> xl_s;
#endif
            if (doS) {
                // ------------------------------------------- S role: chunk s ------------------------------------------------
                const int n0 = s * T, nt = min(N, n0 + T) - n0;
                const bool active = node_l < nt;
                CV<C> B_out, Gacc, nxl;
#pragma unroll
                for (int i = 0; i < CV<C>::NP; ++i) nxl.p[i] = make_float2(-xl_s.p[i].x, -xl_s.p[i].y);
                nxl.s = -xl_s.s;
                const unsigned char *slab = smem + a.off_slabS + (size_t)(qR & 1) * a.slabS_bytes;
                const uint16_t *dego = reinterpret_cast<const uint16_t *>(slab);
                const uint16_t *ell_out = dego + T + node_l;
                const uint16_t *st_out = ell_out + koutp * T;
                mbar_wait(&rfull[qR & 3], (uint32_t)(qR >> 2) & 1u);
                const int deg_out = active ? (int)dego[node_l] : 0;
                const int kmax_out = __reduce_max_sync(0xFFFFFFFFu, deg_out);
                const ST *xr_lo = ringR + posS * HC, *xr_hi = xr_lo - P * HC;
                const float *g_lo = ringG + posS * HC, *g_hi = g_lo - P * HC;
                const int wrapR = P - posS;
                const unsigned char *st_lo = stash + (uint32_t)stS * SS + h * 8, *st_hi = st_lo - (uint32_t)Ps * SS;
                const int wrapS = Ps - stS;
#ifdef TG_SW_FAKE_OWN
                {
                    const int rel = n0 + (active ? node_l : 0) - sw_wlo(s, R);
                    cv_load<C, true>(xl_s, (rel >= wrapR ? xr_hi : xr_lo) + rel * HC, par);
#pragma unroll
                    for (int i = 0; i < CV<C>::NP; ++i) nxl.p[i] = make_float2(-xl_s.p[i].x, -xl_s.p[i].y);
                    nxl.s = -xl_s.s;
                }
#endif
                float A_out = 0.f;
                cv_zero(B_out);
                cv_zero(Gacc);
                int ua = ell_out[0], ub = ell_out[T];
                uint32_t sa_i = st_out[0], sb_i = st_out[T];
#pragma unroll 1
                for (int kk = 0; kk < kmax_out; kk += 2) {
                    const int na = ell_out[(kk + 2) * T], nb = ell_out[(kk + 3) * T];  // look-ahead (never used past koutp)
                    const uint32_t nsa = st_out[(kk + 2) * T], nsb = st_out[(kk + 3) * T];
                    CV<C> ra, rb, ga, gb;
                    const bool wa = ua >= wrapR, wb = ub >= wrapR;
                    cv_load<C, true>(ra, (wa ? xr_hi : xr_lo) + ua * HC, par);
                    cv_load<C, true>(rb, (wb ? xr_hi : xr_lo) + ub * HC, par);
                    cv_load<C, true>(ga, (wa ? g_hi : g_lo) + ua * HC, par);
                    cv_load<C, true>(gb, (wb ? g_hi : g_lo) + ub * HC, par);
                    const float2 pa = *reinterpret_cast<const float2 *>((ua >= wrapS ? st_hi : st_lo) + sa_i * 8u);
                    const float2 pb = *reinterpret_cast<const float2 *>((ub >= wrapS ? st_hi : st_lo) + sb_i * 8u);
                    A_out += pa.y + pb.y;
                    acc_step_gt<C>(B_out, ra, nxl, pa.y);
                    acc_step_gt<C>(B_out, rb, nxl, pb.y);
                    cv_axpy<C>(Gacc, pa.x, ga);
                    cv_axpy<C>(Gacc, pb.x, gb);
                    ua = na;
                    ub = nb;
                    sa_i = nsa;
                    sb_i = nsb;
                }
                ++qR;
                {   // advance the S window to chunk s + 1
                    const int dw = sw_wlo(s + 1, R) - sw_wlo(s, R);
                    posS += dw;
                    if (posS >= P) posS -= P;
                    stS += dw;
                    if (stS >= Ps) stS -= Ps;
                }
                CV<C> dxl, tatt, att_h;
                cv_load<C, false>(att_h, att_s, par);
                const float2 k1 = splat(1.f - slope), aout = splat(slope * A_out);
#pragma unroll
                for (int i = 0; i < CV<C>::NP; ++i) {
                    const float2 DL = __ffma2_rn(k1, B_out.p[i], aout);
                    dxl.p[i] = __ffma2_rn(att_h.p[i], DL, Gacc.p[i]);
                    tatt.p[i] = __fmul2_rn(xl_s.p[i], DL);
                }
                {
                    const float DL = fmaf(1.f - slope, B_out.s, slope * A_out);
                    dxl.s = fmaf(att_h.s, DL, Gacc.s);
                    tatt.s = xl_s.s * DL;
                }
                const int nv = max(0, min(NPW, nt - warp * NPW));
                store_rows(dxl, active, static_cast<ST *>(a.dxl) + (snap0 + n0 + warp * NPW) * HC, nv);
                if (active) {
#pragma unroll
                    for (int i = 0; i < CV<C>::NP; ++i) acc_att.p[i] = __fadd2_rn(acc_att.p[i], tatt.p[i]);
                    if (CV<C>::ODD) acc_att.s += tatt.s;
                }
            }
            if (doD) {
                // ------------------------------------------- D role: chunk d ------------------------------------------------
                const bool main_chunk = d >= sg.ca && d <= sg.cb;
                const int n0 = d * T, nt = min(N, n0 + T) - n0;
                const bool active = node_l < nt;
                CV<C> xl_v, B_in;
                float2 dv;
                {   // delta_v = g_v . (y_v - bias) from the own rows requested before the S role ran
                    float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
                    CV<C> bias_h;
                    cv_load<C, false>(bias_h, nbias_s, par);  // -bias
                    for (int i = 0; i < CV<C>::NP; ++i) d2 = __ffma2_rn(g_v.p[i], __fadd2_rn(y_v.p[i], bias_h.p[i]), d2);
                    float dl = d2.x + d2.y;
                    if (CV<C>::ODD) dl = fmaf(g_v.s, y_v.s + bias_h.s, dl);
                    dv = make_float2(dl, st_v);
                }
                const unsigned char *slab = smem + a.off_slabD + (size_t)(qL & 1) * a.slabD_bytes;
                const int32_t *k0s = reinterpret_cast<const int32_t *>(slab) + 4, *degs = k0s + T;
                const uint16_t *ell_in = reinterpret_cast<const uint16_t *>(degs + T) + node_l;
                mbar_wait(&lfull[qL & 3], (uint32_t)(qL >> 2) & 1u);
                int deg_in = 0;
                uint32_t slot0 = 0;
                if (active) {
                    deg_in = degs[node_l] & 0xFFFF;
                    slot0 = (uint32_t)k0s[node_l];
                }
                const int kmax_in = __reduce_max_sync(0xFFFFFFFFu, deg_in);
                const int wlo = sw_wlo(d, R);
                const ST *xl_lo = ringL + posD * HC, *xl_hi = xl_lo - P * HC;
                const int wrapL = P - posD;
                {   // own xl row: part of the window the producer staged
                    const int rel = n0 + (active ? node_l : 0) - wlo;
                    cv_load<C, true>(xl_v, (rel >= wrapL ? xl_hi : xl_lo) + rel * HC, par);
#ifdef TG_SW_FAKE_OWN
                    cv_load<C, true>(xr_v, (rel >= wrapL ? xl_hi : xl_lo) + rel * HC, par);
                    g_v = xr_v;
                    y_v = xr_v;
                    st_v = 1.f;
#endif
                }
#ifdef TG_SW_FAKE_OWN
                dv = make_float2(cv_dot<C>(g_v, y_v), st_v);
#endif
                int spos = stD + node_l;
                if (spos >= Ps) spos -= Ps;
                float2 *st_own = reinterpret_cast<float2 *>(stash + (uint32_t)spos * SS + h * 8);  // slot j at st_own[2 j]
                float A_in;
                cv_zero(B_in);
                {   // self loop (slot 0 of both CSR rows): evaluated here, read back by the S role like any other out-edge
                    CV<C> sv;
                    const float e = edge_score<C>(attp, attm, xl_v, xr_v, sv);
                    const float gx = cv_dot<C>(g_v, xl_v);
                    const float q = drop.q(slot0);
                    const float alpha = (active ? 1.f : 0.f) * fast_exp2(fminf(e - dv.y, 100.f));
                    const float de = alpha * fmaf(q, gx, -dv.x);
                    A_in = de;
                    acc_step<C>(B_in, sv, de);
                    // (lanes past the last node write zeros into ring rows nobody reads: rows just below the live span)
                    st_own[2 * kinp] = make_float2(0.f, 0.f);  // the zero slot (overwritten below when it is a real edge)
                    st_own[0] = make_float2(alpha * q, de);
                }
                uint32_t hk = (slot0 + 1u) * kDropMul + drop.key;
                int ua = ell_in[0], ub = ell_in[T];
#pragma unroll 1
                for (int kk = 1; kk < kmax_in; kk += 2, hk += 2u * kDropMul) {
                    const int na = ell_in[(kk + 1) * T], nb = ell_in[(kk + 2) * T];  // look-ahead (rows past kinp are never used)
                    CV<C> xa, xb, sa, sb;
                    cv_load<C, true>(xa, (ua >= wrapL ? xl_hi : xl_lo) + ua * HC, par);
                    cv_load<C, true>(xb, (ub >= wrapL ? xl_hi : xl_lo) + ub * HC, par);
                    const float ea = edge_score<C>(attp, attm, xa, xr_v, sa);
                    const float eb = edge_score<C>(attp, attm, xb, xr_v, sb);
                    const float ga = cv_dot<C>(g_v, xa), gb = cv_dot<C>(g_v, xb);
                    const float va = kk < deg_in ? 1.f : 0.f, vb = kk + 1 < deg_in ? 1.f : 0.f;
                    const float aa = va * fast_exp2(fminf(ea - dv.y, 100.f)), ab = vb * fast_exp2(fminf(eb - dv.y, 100.f));
                    const float qa = drop.qh(hk), qb = drop.qh(hk + kDropMul);
                    const float da = aa * fmaf(qa, ga, -dv.x), db = ab * fmaf(qb, gb, -dv.x);
                    A_in += da + db;
                    acc_step<C>(B_in, sa, da);
                    acc_step<C>(B_in, sb, db);
                    st_own[2 * kk] = make_float2(aa * qa, da);
                    st_own[2 * kk + 2] = make_float2(ab * qb, db);  // kk + 1 <= kinp always (kinp = kin rounded down to even)
                    ua = na;
                    ub = nb;
                }
                ++qL;
                {   // advance the D window to chunk d + 1
                    posD += sw_wlo(d + 1, R) - wlo;
                    if (posD >= P) posD -= P;
                    stD += T;
                    if (stD >= Ps) stD -= Ps;
                }
                if (main_chunk) {
                    CV<C> dxr, tatt, att_h;
                    cv_load<C, false>(att_h, att_s, par);
                    const float2 k1 = splat(1.f - slope), ain = splat(slope * A_in);
#pragma unroll
                    for (int i = 0; i < CV<C>::NP; ++i) {
                        const float2 DR = __ffma2_rn(k1, B_in.p[i], ain);
                        dxr.p[i] = __fmul2_rn(att_h.p[i], DR);
                        tatt.p[i] = __fmul2_rn(xr_v.p[i], DR);
                    }
                    {
                        const float DR = fmaf(1.f - slope, B_in.s, slope * A_in);
                        dxr.s = att_h.s * DR;
                        tatt.s = xr_v.s * DR;
                    }
                    const int nv = max(0, min(NPW, nt - warp * NPW));
                    store_rows(dxr, active, static_cast<ST *>(a.dxr) + (snap0 + n0 + warp * NPW) * HC, nv);
                    if (active) {
#pragma unroll
                        for (int i = 0; i < CV<C>::NP; ++i) {
                            acc_att.p[i] = __fadd2_rn(acc_att.p[i], tatt.p[i]);
                            acc_bias.p[i] = __fadd2_rn(acc_bias.p[i], g_v.p[i]);
                        }
                        if (CV<C>::ODD) {
                            acc_att.s += tatt.s;
                            acc_bias.s += g_v.s;
                        }
                    }
                }
            }
            // ---- end of phase: the stash written above is complete for the next phases, the rows read above are free ----
            __syncwarp();
            if (lane == 0) mbar_arrive(&phase[G & 15]);
            if (doS && ++since_flush == kFlushItems) flush();
            mbar_wait(&phase[G & 15], (uint32_t)(G >> 4) & 1u);
        }
        VL = (vbL + sw_whi(sg.dlast, R, N)) % P;
        VR = (vbR + sw_whi(sg.cb, R, N)) % P;
        SV = (svb + min(N, (sg.dlast + 1) * T)) % Ps;
    }
    flush();
    for (; flushes < a.max_flushes; ++flushes)
        if (ctid < 2 * HC) a.partials[((int64_t)blockIdx.x * a.max_flushes + flushes) * 2 * HC + ctid] = 0.f;
}

struct SwGeom {
    int32_t ok, P, Ps, per_st, per_f, per_max;
    uint32_t off_red, off_out, off_slabD, off_slabS, off_xl, off_xr, off_g, off_stash, smem;
};

template <int C, typename ST>
static int launch_sw(SwArgs a, const tecgat_plan_t *plan, int grid, cudaStream_t st, bool *used) {
    const tg_sw_plan *sw = plan->sw;
    *used = false;
    const uint64_t key = (uint64_t(3) << 56) | (uint64_t(C) << 40) | (uint64_t(sizeof(ST)) << 24);
    SwGeom g;
    bool hit = false;
    {
        std::lock_guard<std::mutex> lk(plan->cache_mu);
        auto it = plan->geom_cache.find(key);
        if (it != plan->geom_cache.end()) {
            memcpy(&g, it->second.data(), sizeof(g));
            hit = true;
        }
    }
    if (!hit) {
        constexpr int HC = 2 * C, T = kSwT;
        g.per_st = row_period(uint32_t(HC * sizeof(ST)));
        g.per_f = row_period(uint32_t(HC * 4));
        g.per_max = std::max(g.per_st, g.per_f);
        const int pm = g.per_max;
        // rows: the role's window (T + 2R) + the block in flight (T) + widening of both ends + the segment re-alignment
        g.P = ((2 * T + 2 * sw->R + 3 * pm + pm - 1) / pm) * pm;
        g.Ps = 3 * T + sw->R;  // S(k-2) reads chunks k-3 .. k-1 (+R) while D(k) writes chunk k
        uint32_t o = 256;  // barriers
        g.off_red = o; o += (uint32_t)(((kSwWarps * 2 * 2 * C + 4 * C) * sizeof(float) + 15) & ~size_t(15));  // flush scratch + [att | -bias]
        g.off_out = o; o += (uint32_t)((size_t(T) * HC * sizeof(ST) + 15) & ~size_t(15));
        g.off_slabD = o; o += 2u * sw->slabD_bytes;
        g.off_slabS = o; o += 2u * sw->slabS_bytes;
        o = (o + 127u) & ~127u;
        g.off_xl = o; o += (uint32_t)((size_t(g.P) * HC * sizeof(ST) + 127) & ~size_t(127));
        g.off_xr = o; o += (uint32_t)((size_t(g.P) * HC * sizeof(ST) + 127) & ~size_t(127));
        g.off_g = o; o += (uint32_t)((size_t(g.P) * HC * 4 + 127) & ~size_t(127));
        g.off_stash = o; o += (uint32_t)((size_t(g.Ps) * sw->stash_stride + 127) & ~size_t(127));
        g.smem = o;
        g.ok = o <= 227u * 1024u;
        std::lock_guard<std::mutex> lk(plan->cache_mu);
        auto &blob = plan->geom_cache[key];
        blob.resize(sizeof(g));
        memcpy(blob.data(), &g, sizeof(g));
    }
    if (!g.ok) return TECGAT_OK;  // does not fit: the caller falls back to edge_bwd.cu
    a.P = g.P; a.Ps = g.Ps; a.per_st = g.per_st; a.per_f = g.per_f; a.per_max = g.per_max;
    a.off_red = g.off_red; a.off_out = g.off_out; a.off_slabD = g.off_slabD; a.off_slabS = g.off_slabS;
    a.off_xl = g.off_xl; a.off_xr = g.off_xr; a.off_g = g.off_g; a.off_stash = g.off_stash;
    auto go = [&](auto kern) -> int {
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)g.smem));
        kern<<<(unsigned)grid, kSwThreads, g.smem, st>>>(a);
        tg_count_launch();
        return TECGAT_OK;
    };
    int rc = a.drop_thr ? go(edge_bwd_sw_kernel<C, ST, true>) : go(edge_bwd_sw_kernel<C, ST, false>);
    if (rc != TECGAT_OK) return rc;
    TG_LAUNCH_CHECK();
    *used = true;
    return TECGAT_OK;
}

// Tries the sliding-window kernel; *used = false (and TECGAT_OK) when the plan / shape does not qualify.
int edge_bwd_sw_try(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att, const float *bias, const float *y,
                    const float *stat, const float *gy, void *dxl, void *dxr, float *partials, int grid, int max_flushes,
                    int32_t snapshots, int32_t heads, int32_t out_channels, float negative_slope, float dropout_p, uint32_t drop_thr,
                    uint64_t seed, const uint64_t *seed_dev, int32_t mode, int32_t dtype, cudaStream_t st, bool *used) {
    *used = false;
    const tg_sw_plan *sw = plan->sw;
    // Opt-in (TECGAT_BWD=sw).  Measured on B200 at the default shape (profiles/r02_edge_bwd_sw.md): correct, and each edge's score
    // really is evaluated once (loops 145 instead of 170 instructions per edge and head), but the per-chunk overhead of the
    // ring bookkeeping and the own-row global loads eat the saving: 4.8 ms against edge_bwd.cu's 3.2 ms at B = 128, and parity
    // even with the own rows taken from shared memory.  edge_bwd.cu therefore stays the product path.
    const char *env = tg_env("TECGAT_BWD");
    if (!env || env[0] != 's') return TECGAT_OK;
    if (!sw || sw->T != kSwT || heads != 2 || (out_channels != 11 && out_channels != 5) || mode != TECGAT_MODE_SHARED) return TECGAT_OK;
    SwArgs a;
    a.xl = xl; a.xr = xr; a.att = att; a.bias = bias; a.y = y; a.stat = stat; a.gy = gy; a.dxl = dxl; a.dxr = dxr; a.partials = partials;
    a.slabD = sw->slabD; a.slabS = sw->slabS;
    a.N = plan->num_nodes; a.J = sw->J; a.S = snapshots; a.R = sw->R;
    a.kinp = sw->kinp; a.koutp = sw->koutp;
    a.slabD_bytes = sw->slabD_bytes; a.slabS_bytes = sw->slabS_bytes; a.stash_stride = sw->stash_stride;
    a.slope = negative_slope; a.inv_keep = 1.f / (1.f - dropout_p); a.drop_thr = drop_thr; a.seed = seed; a.seed_dev = seed_dev;
    a.stream_stride = dropout_stream_stride(plan->num_edges);
    a.max_flushes = max_flushes;
    a.chunks = int64_t(sw->J) * snapshots;
    if (out_channels == 11)
        return dtype == TECGAT_F32 ? launch_sw<11, float>(a, plan, grid, st, used) : launch_sw<11, __nv_bfloat16>(a, plan, grid, st, used);
    return dtype == TECGAT_F32 ? launch_sw<5, float>(a, plan, grid, st, used) : launch_sw<5, __nv_bfloat16>(a, plan, grid, st, used);
}

}  // namespace tg
