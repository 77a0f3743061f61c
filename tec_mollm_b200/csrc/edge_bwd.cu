// edge_bwd.cu -- fused GATv2 edge phase, backward, atomic-free (formulas section 8a-3, restated and
// gradient-checked in oracle/gatv2_oracle.py::gatv2_backward_manual).  With g = dL/dy, delta_i = g_i . (y_i - bias),
// q = keep/(1-p), s_ij = xl_j + xr_i, sgn(s) = +1 if s > 0 else -1:
//     alpha_ij = exp2(e_ij - stat_i)                        (recomputed from xl_j, xr_i and the saved log-sum-exp)
//     de_ij    = alpha_ij (q_ij g_i.xl_j - delta_i)
//     d xr_i   = att_p A_i + att_m Bs_i,        A_i = sum_j de_ij,  Bs_i[c] = sum_j de_ij sgn(s_ij[c])   (DESTINATION role)
//     d xl_j   = sum_i alpha_ij q_ij g_i + att_p A'_j + att_m Bs'_j                                    (SOURCE role)
//     d att    = (1+slope)/2 sum_rows (xl_v A'_v + xr_v A_v) + (1-slope)/2 sum_ij de_ij |s_ij|,   d bias = sum_i g_i
// (att_p = att (1+slope)/2, att_m = att (1-slope)/2: LeakyReLU' = (1+slope)/2 + (1-slope)/2 sgn(s), slope at s = 0.)
// One lane owns one (node, head) and walks BOTH CSR orientations, so every gradient row has exactly one writer: no
// atomics, bit-reproducible.  d att / d bias: per-lane fp64 accumulators across the CTA's items, one fixed-order
// cross-lane reduction per CTA at the end, then a fixed-order fp64 second stage over the CTAs (reduce.cu).
// Execution model, lane mapping and arithmetic: edge_common.cuh.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <utility>
#include <vector>

#include "edge_common.cuh"
#include "reduce.cuh"

namespace tg {

struct EdgeBwdArgs {
    const void *xl, *xr;
    const float *att, *bias, *y, *stat, *gy;
    void *dxl, *dxr;
    float *partials;  // (grid, 2*HC): [d att | d bias] per CTA
    const tg_tile_meta *meta;
    const unsigned char *slabs;
    const int32_t *rowptr_in, *col_in, *rowptr_out, *col_out, *slot_out;  // tiles that are not staged
    int32_t N, T, num_tiles, S, H, npw;
    float slope, inv_keep;
    uint32_t drop_thr;
    uint32_t stream_stride;  // dropout counter distance between consecutive (snapshot, head) streams
    uint64_t seed;
    const uint64_t *seed_dev;  // non-NULL: the seed lives in device memory (the one the forward used)
    int32_t literal;
    int32_t cap_rows, cap_kin, cap_kout, num_stages;
    int32_t semi;  // 1: rows too wide for shared memory -- stage only slab + stat + delta planes, gather the rows from L2
    int32_t per_st, per_f, per_stat;  // 16-byte row periods (rows) of xl/xr, of g/y and of stat
    // shared-memory map (bytes): [barriers 128][tile table][y window][out staging][stage 0][stage 1]..
    // stage: [slab][stat window raw][delta|stat planes (H x cap_rows float2)][xl window][xr window][g window]
    uint32_t stage_bytes, off_meta, off_red, off_stage0, off_y, off_out, off_statraw, off_ds, off_xl, off_xr, off_g;
    int32_t max_flushes;  // partial rows per CTA
    int64_t items;
    ItemSchedule sched;  // which items each CTA takes (equal estimated work)
};

// score of an out-edge (v -> u) from the source's side:  c1 + att_m . |xl_v + xr_u|   (c1 = att_p . xl_v)
template <int C>
__device__ __forceinline__ float out_score(const CV<C> &attm, float c1, const CV<C> &xl_v, const CV<C> &xru, CV<C> &s) {
    float2 ea = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        s.p[i] = __fadd2_rn(xl_v.p[i], xru.p[i]);
        ea = __ffma2_rn(attm.p[i], abs2(s.p[i]), ea);
    }
    float e = c1 + (ea.x + ea.y);
    if (CV<C>::ODD) {
        s.s = xl_v.s + xru.s;
        e = fmaf(attm.s, fabsf(s.s), e);
    }
    return e;
}

// Everything one lane = (node v, head) does for one item.  The self loop (slot 0 of both CSR rows, the same edge) is
// evaluated once and feeds both roles; the remaining in- and out-slots are walked two per iteration, branch-free
// (slots past the degree read a valid row and get weight 0).
//   NbrIn(k) / NbrOut(k): row index of the k-th in- / out-neighbour relative to the base pointers
//   SlotOut(k): in-CSR slot (dropout counter) of the k-th out-edge;  DsOf(u, gu): (delta_u, stat_u)
template <int C, typename ST, bool VEC, class Drop, class NbrIn, class NbrOut, class SlotOut, class DsOf, class WaitDs>
__device__ __forceinline__ void bwd_lane(const CV<C> &attp, const CV<C> &attm, const CV<C> &att, float slope, const Drop &drop,
                                         const ST *xl_base, const ST *xr_base, const float *g_base, int HC, int par, ptrdiff_t vrow,
                                         float2 dv, int deg_in, int kmax_in, int deg_out, int kmax_out, uint32_t slot0,
                                         NbrIn nbr_in, NbrOut nbr_out, SlotOut slot_out, DsOf ds_of, WaitDs wait_ds, bool active, CV<C> &dxl,
                                         CV<C> &dxr, CV<C> &tatt, CV<C> &g_v) {
    CV<C> xl_v, xr_v, B_in, B_out, G;
    if (active) {
        cv_load<C, VEC>(xl_v, xl_base + vrow * HC, par);
        cv_load<C, VEC>(xr_v, xr_base + vrow * HC, par);
        cv_load<C, VEC>(g_v, g_base + vrow * HC, par);
    } else {
        cv_zero(xl_v);
        cv_zero(xr_v);
        cv_zero(g_v);
    }
    cv_zero(B_in);
    cv_zero(B_out);
    cv_zero(G);
    float A_in, A_out;
    {   // ---- self loop --------------------------------------------------------------------------------------------
        CV<C> s;
        const float e = edge_score<C>(attp, attm, xl_v, xr_v, s);
        const float gx = cv_dot<C>(g_v, xl_v);
        const float q = drop.q(slot0);
        const float alpha = (active ? 1.f : 0.f) * fast_exp2(fminf(e - dv.y, 100.f));
        const float de = alpha * fmaf(q, gx, -dv.x);
        A_in = A_out = de;
        acc_step<C>(B_in, s, de);
        B_out = B_in;
        cv_axpy<C>(G, alpha * q, g_v);
    }
    // ---- role 1: v as DESTINATION, in-edges (u -> v): A_in, B_in ------------------------------------------------
    uint32_t hk = (slot0 + 1u) * kDropMul + drop.key;
    // neighbour indices are fetched one iteration ahead (index load -> address -> row load is the loop's longest dependent
    // chain); the look-ahead slots past the row's padded length are never used
    ptrdiff_t ua = nbr_in(1), ub = nbr_in(2);
#pragma unroll 1
    for (int k = 1; k < kmax_in; k += 2, hk += 2u * kDropMul) {
        const ptrdiff_t na = nbr_in(k + 2), nb = nbr_in(k + 3);
        CV<C> xa, xb, sa, sb;
        cv_load<C, VEC>(xa, xl_base + ua * HC, par);
        cv_load<C, VEC>(xb, xl_base + ub * HC, par);
        const float ea = edge_score<C>(attp, attm, xa, xr_v, sa);
        const float eb = edge_score<C>(attp, attm, xb, xr_v, sb);
        const float ga = cv_dot<C>(g_v, xa), gb = cv_dot<C>(g_v, xb);
        const float va = k < deg_in ? 1.f : 0.f, vb = k + 1 < deg_in ? 1.f : 0.f;
        const float aa = va * fast_exp2(fminf(ea - dv.y, 100.f)), ab = vb * fast_exp2(fminf(eb - dv.y, 100.f));
        const float da = aa * fmaf(drop.qh(hk), ga, -dv.x);
        const float db = ab * fmaf(drop.qh(hk + kDropMul), gb, -dv.x);
        A_in += da + db;
        acc_step<C>(B_in, sa, da);
        acc_step<C>(B_in, sb, db);
        ua = na;
        ub = nb;
    }
    // ---- role 2: v as SOURCE, out-edges (v -> u): A_out, B_out, G = sum alpha q g_u ------------------------------
    wait_ds();  // the (delta, stat) of every window row has been written (other warps' pre-pass shares)
    const float c1 = cv_dot<C>(attp, xl_v);
    ua = nbr_out(1);
    ub = nbr_out(2);
    uint32_t sla = slot_out(1), slb = slot_out(2);
#pragma unroll 1
    for (int k = 1; k < kmax_out; k += 2) {
        const ptrdiff_t na = nbr_out(k + 2), nb = nbr_out(k + 3);
        const uint32_t nsa = slot_out(k + 2), nsb = slot_out(k + 3);
        CV<C> ra, rb, ga, gb, sa, sb;
        cv_load<C, VEC>(ra, xr_base + ua * HC, par);
        cv_load<C, VEC>(rb, xr_base + ub * HC, par);
        cv_load<C, VEC>(ga, g_base + ua * HC, par);
        cv_load<C, VEC>(gb, g_base + ub * HC, par);
        const float2 dua = ds_of(ua, ga), dub = ds_of(ub, gb);
        const float ea = out_score<C>(attm, c1, xl_v, ra, sa);
        const float eb = out_score<C>(attm, c1, xl_v, rb, sb);
        const float gxa = cv_dot<C>(ga, xl_v), gxb = cv_dot<C>(gb, xl_v);
        const float va = k < deg_out ? 1.f : 0.f, vb = k + 1 < deg_out ? 1.f : 0.f;
        const float aa = va * fast_exp2(fminf(ea - dua.y, 100.f)), ab = vb * fast_exp2(fminf(eb - dub.y, 100.f));
        const float qa = drop.q(sla), qb = drop.q(slb);
        const float da = aa * fmaf(qa, gxa, -dua.x), db = ab * fmaf(qb, gxb, -dub.x);
        A_out += da + db;
        acc_step<C>(B_out, sa, da);
        acc_step<C>(B_out, sb, db);
        cv_axpy<C>(G, aa * qa, ga);
        cv_axpy<C>(G, ab * qb, gb);
        ua = na; ub = nb;
        sla = nsa; slb = nsb;
    }
    // ---- rows:  DR = slope A_in + (1-slope) B_in,  d xr = att DR;   DL likewise,  d xl = G + att DL;
    //      d att partial of this row = xr_v DR + xl_v DL   (lrelu(s) = s lrelu'(s), s = xl + xr) ---------------------
    const float2 k1 = splat(1.f - slope), ain = splat(slope * A_in), aout = splat(slope * A_out);
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) {
        const float2 DR = __ffma2_rn(k1, B_in.p[i], ain), DL = __ffma2_rn(k1, B_out.p[i], aout);
        dxr.p[i] = __fmul2_rn(att.p[i], DR);
        dxl.p[i] = __ffma2_rn(att.p[i], DL, G.p[i]);
        tatt.p[i] = __ffma2_rn(xr_v.p[i], DR, __fmul2_rn(xl_v.p[i], DL));
    }
    {
        const float DR = fmaf(1.f - slope, B_in.s, slope * A_in), DL = fmaf(1.f - slope, B_out.s, slope * A_out);
        dxr.s = att.s * DR;
        dxl.s = fmaf(att.s, DL, G.s);
        tatt.s = fmaf(xr_v.s, DR, xl_v.s * DL);
    }
}

// HT > 0: compile-time number of heads (address arithmetic folds); 0: runtime.  SEMI: the semi-staged variant (rows too
// wide for shared memory) is a separate instantiation so that the fully staged kernel's code stays compact.
// GATHER: the launch contains tiles that are not staged (false drops the gather-from-global code: every item is staged).
// WG: 384-thread CTA = 8 consumer warps (two warpgroups) + one producer warpgroup; the producer group hands its registers to
// the consumers (setmaxnreg), which lifts the consumers to 8 warps x 240 registers -- the register file's split per scheduler
// (16 K registers each) would otherwise cap a 9-warp CTA at 168 registers per thread.
constexpr int kWgConsumerRegs = 232, kWgProducerRegs = 40;
template <int C, typename ST, bool VEC, int HT, bool SEMI, bool GATHER, bool DROP, bool WG = false>
__global__ void __launch_bounds__(WG ? 384 : 256, 1) edge_bwd_kernel(const __grid_constant__ EdgeBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + kMaxStages;
    uint64_t *yfull = empty + kMaxStages;
    uint64_t *yempty = yfull + 1;
    uint64_t *dready = yempty + 1;  // per stage: every consumer warp has written its share of the (delta, stat) planes
    const tg_tile_meta *meta_s = reinterpret_cast<const tg_tile_meta *>(smem + a.off_meta);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = WG ? 8 : (blockDim.x >> 5) - 1;
    const int nct = ncw * 32;  // consumer threads
    // WG kernels exist only for 8 consumer warps x (32 / heads) nodes: the tile size is a compile-time constant there (the ELL
    // strides and the row-to-lane map of the delta pre-pass fold into immediates)
    const int T = (WG && HT > 0) ? 8 * (32 / pad_heads(HT > 0 ? HT : 1)) : a.T;
    const int H = HT > 0 ? HT : a.H, HC = H * C, N = a.N;
    const int Ts = (T + 7) & ~7;  // row stride of the slab sections
    const int NS = a.num_stages;
    const uint32_t RB_ST = (uint32_t)HC * sizeof(ST), RB_F = (uint32_t)HC * 4u, RB_STAT = (uint32_t)H * 4u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ncw);
            mbar_init(&dready[s], ncw);
        }
        mbar_init(yfull, 1);
        mbar_init(yempty, ncw);
        fence_mbar_init();
    }
    if (a.num_tiles <= kMetaSmemTiles)
        for (int i = threadIdx.x; i < a.num_tiles * 8; i += blockDim.x)
            reinterpret_cast<int32_t *>(smem + a.off_meta)[i] = reinterpret_cast<const int32_t *>(a.meta)[i];
    __syncthreads();
    const ItemRange R = cta_items_scheduled(a.sched, a.items);
    int snap = (int)(R.w0 / a.num_tiles), tile = (int)(R.w0 % a.num_tiles);
    const int n_items = (int)(R.w1 - R.w0);  // 32-bit loop counter (a CTA never owns 2^31 items)
    const int64_t Rtot = (int64_t)a.S * N;
    Ring ring{0, 0u};
    uint32_t yph = 0;  // phase parity of the single y window

    if (warp >= ncw) {
        // ================================ producer warp ================================
        if (WG) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWgProducerRegs));
            if (warp > ncw + 1 || (SEMI && warp > ncw)) return;  // the group's remaining warps only donate their registers
        }
        // WG: two producer warps -- `ncw` streams the stages, `ncw + 1` the single y window (gated by the consumers' delta
        // pre-pass, not by a stage): neither waits on the other's barrier, and each keeps few values live in its 40 registers
        const bool do_stage = !WG || warp == ncw, do_y = !WG || warp == ncw + 1;
        for (int w = 0; w < n_items; ++w) {
            const tg_tile_meta m = load_meta(meta_s, a.meta, a.num_tiles, tile);
            const int n0 = tile * T, nt = min(N, n0 + T) - n0;
            const bool lit = a.literal && snap > 0;
            const int lo = lit ? n0 : m.lo, win = lit ? nt : m.hi - m.lo;
            const bool staged = !GATHER || (m.eligible && win <= a.cap_rows && (m.kin_kout & 0xFFFF) <= a.cap_kin && (m.kin_kout >> 16) <= a.cap_kout);
            if (staged) {
                const int64_t row0 = (int64_t)snap * N + lo;
                unsigned char *stage = smem + a.off_stage0 + (size_t)ring.st * a.stage_bytes;
                const WinCopy c3 = win_copy(a.stat, row0, win, RB_STAT, a.per_stat, Rtot);
                if (SEMI) {  // slab + stat window only; the consumers gather xl / xr / g / y rows from global memory (L2)
                    if (lane == 0) mbar_wait_relaxed(&empty[ring.st], ring.ph ^ 1u);
                    __syncwarp();
                    if (c3.tail) {
                        win_copy_tail(c3, stage + a.off_statraw, lane);
                        __syncwarp();
                    }
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full[ring.st], (uint32_t)m.slab_bytes + c3.mid);
                        bulk_g2s(stage, a.slabs + m.slab_off, (uint32_t)m.slab_bytes, &full[ring.st]);
                        if (c3.mid) bulk_g2s(stage + a.off_statraw, c3.src, c3.mid, &full[ring.st]);
                    }
                } else {
                if (do_stage) {
                    const WinCopy c0 = win_copy(a.xl, row0, win, RB_ST, a.per_st, Rtot);
                    const WinCopy c1 = win_copy(a.xr, row0, win, RB_ST, a.per_st, Rtot);
                    const WinCopy c2 = win_copy(a.gy, row0, win, RB_F, a.per_f, Rtot);
                    if (lane == 0) mbar_wait_relaxed(&empty[ring.st], ring.ph ^ 1u);
                    __syncwarp();
                    if (c0.tail | c1.tail | c2.tail | c3.tail) {  // only the last rows of the last snapshot
                        win_copy_tail(c0, stage + a.off_xl, lane);
                        win_copy_tail(c1, stage + a.off_xr, lane);
                        win_copy_tail(c2, stage + a.off_g, lane);
                        win_copy_tail(c3, stage + a.off_statraw, lane);
                        __syncwarp();
                    }
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full[ring.st], (uint32_t)m.slab_bytes + c0.mid + c1.mid + c2.mid + c3.mid);
                        bulk_g2s(stage, a.slabs + m.slab_off, (uint32_t)m.slab_bytes, &full[ring.st]);
                        if (c0.mid) bulk_g2s(stage + a.off_xl, c0.src, c0.mid, &full[ring.st]);
                        if (c1.mid) bulk_g2s(stage + a.off_xr, c1.src, c1.mid, &full[ring.st]);
                        if (c2.mid) bulk_g2s(stage + a.off_g, c2.src, c2.mid, &full[ring.st]);
                        if (c3.mid) bulk_g2s(stage + a.off_statraw, c3.src, c3.mid, &full[ring.st]);
                    }
                }
                if (do_y) {
                    const WinCopy c4 = win_copy(a.y, row0, win, RB_F, a.per_f, Rtot);
                    if (lane == 0) mbar_wait_relaxed(yempty, yph ^ 1u);  // the single y window: released right after the delta pre-pass
                    __syncwarp();
                    if (c4.tail) {
                        win_copy_tail(c4, smem + a.off_y, lane);
                        __syncwarp();
                    }
                    if (lane == 0) {
                        mbar_arrive_expect_tx(yfull, c4.mid);
                        if (c4.mid) bulk_g2s(smem + a.off_y, c4.src, c4.mid, yfull);
                    }
                }
                }
                ring.advance(NS);
                if (!SEMI) yph ^= 1u;
            }
            if (++tile == a.num_tiles) { tile = 0; ++snap; }
        }
        return;
    }

    // ================================ consumer warps ================================
    if (WG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWgConsumerRegs));
    const int npw = HT > 0 ? 32 / pad_heads(HT) : a.npw;
    const int nw = lane & (npw - 1);
    const int h = lane / npw;
    const int node_l = warp * npw + nw;
    const bool head_ok = h < H;
    const int hh = head_ok ? h : 0;
    const int par = VEC ? ((hh * C) & 1) : 0;
    const int ctid = threadIdx.x;  // consumer thread id (consumer warps come first)
    CV<C> attp, attm, att_h, bias_h;
    cv_load_param<C>(attp, a.att + hh * C, par, 0.5f * (1.f + a.slope) * kLog2e);
    cv_load_param<C>(attm, a.att + hh * C, par, 0.5f * (1.f - a.slope) * kLog2e);
    cv_load_param<C>(att_h, a.att + hh * C, par, 1.f);
    cv_load_param<C>(bias_h, a.bias + hh * C, par, 1.f);
    uint32_t key = 0;
    int key_snap = -1;
    const uint32_t drop_base = dropout_base((DROP && a.seed_dev) ? __ldg(a.seed_dev) : a.seed);
    // per-lane fp32 partial sums of d att and d bias: one term per item, flushed to a CTA partial row every kFlushItems
    // items (the second stage sums all rows in fp64)
    CV<C> acc_att, acc_bias;
    cv_zero(acc_att);
    cv_zero(acc_bias);
    float *red = reinterpret_cast<float *>(smem + a.off_red);  // (ncw, H, 2, C) flush scratch
    int since_flush = 0, flushes = 0;
    auto flush = [&]() {
        // fixed-order reduction: lanes of a head inside the warp (xor tree), then the warps; one partial row per flush
        auto put = [&](int which, int c, float v) {
            for (int off = npw >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
            if (nw == 0 && head_ok) red[((warp * H + hh) * 2 + which) * C + c] = v;
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
            for (int wq = 0; wq < ncw; ++wq) v += (double)red[((wq * H + hd) * 2 + which) * C + c];
            a.partials[((int64_t)blockIdx.x * a.max_flushes + flushes) * 2 * HC + which * HC + j] = (float)v;
        }
        bar_sync_named(kBarConsumers, nct);
        cv_zero(acc_att);
        cv_zero(acc_bias);
        ++flushes;
        since_flush = 0;
    };
    ST *out_l = reinterpret_cast<ST *>(smem + a.off_out) + warp * npw * HC;               // d xl rows of this warp
    ST *out_r = reinterpret_cast<ST *>(smem + a.off_out) + (size_t)T * HC + warp * npw * HC;  // d xr rows

    for (int w = 0; w < n_items; ++w) {
        const tg_tile_meta m = load_meta(meta_s, a.meta, a.num_tiles, tile);
        const int n0 = tile * T, nt = min(N, n0 + T) - n0;
        const bool lit = a.literal && snap > 0;
        const int lo = lit ? n0 : m.lo, win = lit ? nt : m.hi - m.lo;
        const bool staged = !GATHER || (m.eligible && win <= a.cap_rows && (m.kin_kout & 0xFFFF) <= a.cap_kin && (m.kin_kout >> 16) <= a.cap_kout);
        const bool active = head_ok && node_l < nt;
        const int64_t row = (int64_t)snap * N + n0 + node_l;
        if (DROP && a.drop_thr && snap != key_snap) {
            key = dropout_key(drop_base, (uint32_t)snap, (uint32_t)H, (uint32_t)hh, a.stream_stride);
            key_snap = snap;
        }
        DropCfg<DROP> drop;
        drop.thr = a.drop_thr;
        drop.key = key;
        drop.inv_keep = a.inv_keep;
        CV<C> dxl, dxr, tatt, g_v;
        if (staged) {
            const int64_t row0 = (int64_t)snap * N + lo;
            unsigned char *stage = smem + a.off_stage0 + (size_t)ring.st * a.stage_bytes;
            const int32_t *hdr = reinterpret_cast<const int32_t *>(stage);
            const int32_t *k0s = hdr + 4;
            const int32_t *degs = k0s + Ts;
            const int kin_t = m.kin_kout & 0xFFFF, kout_t = m.kin_kout >> 16;
            const uint16_t *ell_in = reinterpret_cast<const uint16_t *>(degs + Ts) + node_l;
            const uint16_t *ell_out = ell_in + (size_t)kin_t * Ts;
            const uint16_t *slot_rel = ell_out + (size_t)kout_t * Ts;
            const float *stat_s = reinterpret_cast<const float *>(stage + a.off_statraw + win_skip(row0, RB_STAT, a.per_stat));
            float2 *ds = reinterpret_cast<float2 *>(stage + a.off_ds);  // [head][window row] -> (delta, stat)
            mbar_wait(&full[ring.st], ring.ph);
            // The row windows are in shared memory (full staging) or stay in global memory (semi staging, rows too wide):
            // the same code is inlined once per address space.
            auto item_body = [&](auto semi_tag, const ST *xl_s, const ST *xr_s, const float *g_s, const float *y_s) {
                constexpr bool kSemi = decltype(semi_tag)::value;
                // ---- pre-pass: delta = g . (y - bias).  Each lane takes the window rows node_l, node_l + T, .. of its head
                //      (needed by the SOURCE role of every lane) and its own row (DESTINATION role: no waiting on other warps)
                auto delta_of = [&](int r) -> float2 {
                    CV<C> gg, yy;
                    cv_load<C, VEC>(gg, g_s + (ptrdiff_t)r * HC + hh * C, par);
                    cv_load<C, VEC>(yy, y_s + (ptrdiff_t)r * HC + hh * C, par);
                    float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < CV<C>::NP; ++q) d2 = __ffma2_rn(gg.p[q], __fadd2_rn(yy.p[q], make_float2(-bias_h.p[q].x, -bias_h.p[q].y)), d2);
                    float dl = d2.x + d2.y;
                    if (CV<C>::ODD) dl = fmaf(gg.s, yy.s - bias_h.s, dl);
                    return make_float2(dl, stat_s[r * H + hh]);
                };
                // A lane's share of the window starts at its OWN row's residue mod T, so the own row (DESTINATION role: no
                // waiting on other warps) is one of the rows it computes anyway.
                const int vl = n0 + node_l - lo;  // own row inside the window
                float2 dv = make_float2(0.f, 0.f);
                if (head_ok)
                    for (int r = vl % T; r < win; r += T) {
                        const float2 d = delta_of(r);
                        ds[hh * a.cap_rows + r] = d;
                        if (r == vl) dv = d;
                    }
                if (!active) dv = make_float2(0.f, 0.f);
                __syncwarp();
                if (lane == 0) {
                    if (!kSemi) mbar_arrive(yempty);  // y window may be refilled for the next item
                    mbar_arrive(&dready[ring.st]);
                }
                int deg_in = 0, deg_out = 0;
                uint32_t slot0 = 0;
                if (active) {
                    const int d = degs[node_l];
                    deg_in = d & 0xFFFF;
                    deg_out = d >> 16;
                    if (lit) { deg_in = min(deg_in, 1); deg_out = min(deg_out, 1); }
                    slot0 = (uint32_t)k0s[node_l];
                }
                const uint32_t slot_base = (uint32_t)hdr[3];
                const float2 *ds_h = ds + hh * a.cap_rows;
                uint64_t *dr_bar = &dready[ring.st];
                const uint32_t dr_ph = ring.ph;
                const int kmax_in = __reduce_max_sync(0xFFFFFFFFu, deg_in), kmax_out = __reduce_max_sync(0xFFFFFFFFu, deg_out);
                bwd_lane<C, ST, VEC>(
                    attp, attm, att_h, a.slope, drop, xl_s + hh * C, xr_s + hh * C, g_s + hh * C, HC, par, (ptrdiff_t)vl, dv, deg_in, kmax_in,
                    deg_out, kmax_out, slot0, [&](int k) -> ptrdiff_t { return (ptrdiff_t)ell_in[k * Ts]; },
                    [&](int k) -> ptrdiff_t { return (ptrdiff_t)ell_out[k * Ts]; },
                    [&](int k) -> uint32_t { return slot_base + (uint32_t)slot_rel[k * Ts]; },
                    [&](ptrdiff_t u, const CV<C> &) -> float2 { return ds_h[u]; }, [&]() { mbar_wait(dr_bar, dr_ph); }, active, dxl, dxr,
                    tatt, g_v);
            };
            if constexpr (SEMI) {
                item_body(std::true_type{}, static_cast<const ST *>(a.xl) + row0 * HC, static_cast<const ST *>(a.xr) + row0 * HC,
                          a.gy + row0 * HC, a.y + row0 * HC);
            } else {
                mbar_wait(yfull, yph);
                yph ^= 1u;
                item_body(std::false_type{}, reinterpret_cast<const ST *>(stage + a.off_xl + win_skip(row0, RB_ST, a.per_st)),
                          reinterpret_cast<const ST *>(stage + a.off_xr + win_skip(row0, RB_ST, a.per_st)),
                          reinterpret_cast<const float *>(stage + a.off_g + win_skip(row0, RB_F, a.per_f)),
                          reinterpret_cast<const float *>(smem + a.off_y + win_skip(row0, RB_F, a.per_f)));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ring.st]);
            ring.advance(NS);
            // ---- outputs: stage the warp's rows, write them as contiguous runs -----------------------------------
            if (active) {
                cv_store<C, VEC>(out_l + nw * HC + hh * C, dxl, par);
                cv_store<C, VEC>(out_r + nw * HC + hh * C, dxr, par);
            }
            __syncwarp();
            const int nv = max(0, min(npw, nt - warp * npw));
            const int64_t r0 = ((int64_t)snap * N + n0 + warp * npw) * HC;
            ST *dl_g = static_cast<ST *>(a.dxl) + r0, *dr_g = static_cast<ST *>(a.dxr) + r0;
            if (VEC) {  // rows are pair aligned (8 B fp32 / 4 B bf16) in shared and in global memory
                using W = typename std::conditional<sizeof(ST) == 4, uint2, uint32_t>::type;
                const W *sl = reinterpret_cast<const W *>(out_l), *sr = reinterpret_cast<const W *>(out_r);
                if (HT > 0 && nv == npw) {  // full warp slice, compile-time trip count: straight-line copy
                    constexpr int kWords = HT > 0 ? (32 / pad_heads(HT > 0 ? HT : 1)) * (HT > 0 ? HT : 1) * C / 2 : 0;
#pragma unroll
                    for (int i = 0; i < (kWords + 31) / 32; ++i)
                        if (i * 32 + lane < kWords) {
                            reinterpret_cast<W *>(dl_g)[i * 32 + lane] = sl[i * 32 + lane];
                            reinterpret_cast<W *>(dr_g)[i * 32 + lane] = sr[i * 32 + lane];
                        }
                } else {
                    for (int i = lane; i < nv * HC / 2; i += 32) {
                        reinterpret_cast<W *>(dl_g)[i] = sl[i];
                        reinterpret_cast<W *>(dr_g)[i] = sr[i];
                    }
                }
            } else {
                for (int i = lane; i < nv * HC; i += 32) {
                    dl_g[i] = out_l[i];
                    dr_g[i] = out_r[i];
                }
            }
            __syncwarp();
        } else if constexpr (GATHER) {
            // ---- window too large for a stage: everything straight from global memory (L2) --------------------------
            const int64_t snap0 = (int64_t)snap * N;
            const ST *xl_snap = static_cast<const ST *>(a.xl) + snap0 * HC + hh * C;
            const ST *xr_snap = static_cast<const ST *>(a.xr) + snap0 * HC + hh * C;
            const float *g_snap = a.gy + snap0 * HC + hh * C;
            const float *y_snap = a.y + snap0 * HC + hh * C;
            const float *stat_snap = a.stat + snap0 * H + hh;
            auto ds_of = [&](ptrdiff_t node, const CV<C> &gn) -> float2 {
                CV<C> yy;
                cv_load<C, VEC>(yy, y_snap + node * HC, par);
                float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < CV<C>::NP; ++q) d2 = __ffma2_rn(gn.p[q], __fadd2_rn(yy.p[q], make_float2(-bias_h.p[q].x, -bias_h.p[q].y)), d2);
                float dl = d2.x + d2.y;
                if (CV<C>::ODD) dl = fmaf(gn.s, yy.s - bias_h.s, dl);
                return make_float2(dl, stat_snap[node * H]);
            };
            const int v = n0 + node_l;
            int k0i = 0, deg_in = 0, k0o = 0, deg_out = 0;
            float2 dv = make_float2(0.f, 0.f);
            if (active) {
                k0i = __ldg(a.rowptr_in + v);
                deg_in = __ldg(a.rowptr_in + v + 1) - k0i;
                k0o = __ldg(a.rowptr_out + v);
                deg_out = __ldg(a.rowptr_out + v + 1) - k0o;
                if (lit) { deg_in = min(deg_in, 1); deg_out = min(deg_out, 1); }
                CV<C> gv0;
                cv_load<C, VEC>(gv0, g_snap + (ptrdiff_t)v * HC, par);
                dv = ds_of(v, gv0);
            }
            const int kmax_in = __reduce_max_sync(0xFFFFFFFFu, deg_in), kmax_out = __reduce_max_sync(0xFFFFFFFFu, deg_out);
            const int vsafe = active ? v : 0;
            bwd_lane<C, ST, VEC>(
                attp, attm, att_h, a.slope, drop, xl_snap, xr_snap, g_snap, HC, par, (ptrdiff_t)vsafe, dv, deg_in, kmax_in, deg_out, kmax_out,
                (uint32_t)k0i, [&](int k) -> ptrdiff_t { return k < deg_in ? (ptrdiff_t)__ldg(a.col_in + k0i + k) : (ptrdiff_t)vsafe; },
                [&](int k) -> ptrdiff_t { return k < deg_out ? (ptrdiff_t)__ldg(a.col_out + k0o + k) : (ptrdiff_t)vsafe; },
                [&](int k) -> uint32_t { return k < deg_out ? (uint32_t)__ldg(a.slot_out + k0o + k) : 0u; }, ds_of, []() {}, active, dxl, dxr, tatt,
                g_v);
            if (active) {
                cv_store<C, VEC>(static_cast<ST *>(a.dxl) + row * HC + hh * C, dxl, par);
                cv_store<C, VEC>(static_cast<ST *>(a.dxr) + row * HC + hh * C, dxr, par);
            }
        }
        // ---- parameter-gradient partials of this item -------------------------------------------------------------------
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
        if (++since_flush == kFlushItems) flush();
        if (++tile == a.num_tiles) { tile = 0; ++snap; }
    }

    flush();
    // rows this CTA did not use (fewer items than the largest CTA): zeros
    for (; flushes < a.max_flushes; ++flushes)
        if (ctid < 2 * HC) a.partials[((int64_t)blockIdx.x * a.max_flushes + flushes) * 2 * HC + ctid] = 0.f;
}

struct StagePickBwd {
    int num_stages, cap_rows, cap_kin, cap_kout;
    uint32_t stage_bytes, off_statraw, off_ds, off_xl, off_xr, off_g;
};
struct BwdGeom {  // everything the stage sizes depend on
    int T, H, HC;
    size_t es;
    int per_st, per_f, per_stat;
    size_t slab(int kin, int kout) const { return size_t(round16(16 + 8 * ((T + 7) & ~7) + 2 * ((T + 7) & ~7) * (kin + 2 * kout))); }
    size_t win_st(int rows) const { return round16(uint32_t((rows + 2 * (per_st - 1)) * HC * es)); }
    size_t win_f(int rows) const { return round16(uint32_t((rows + 2 * (per_f - 1)) * HC * 4)); }
    size_t win_stat(int rows) const { return round16(uint32_t((rows + 2 * (per_stat - 1)) * H * 4)); }
    size_t ds(int rows) const { return round16(uint32_t(rows * H * 8)); }
    size_t stage(int rows, int kin, int kout) const { return slab(kin, kout) + win_stat(rows) + ds(rows) + 2 * win_st(rows) + win_f(rows); }
    size_t stage_semi(int rows, int kin, int kout) const { return slab(kin, kout) + win_stat(rows) + ds(rows); }
};
static StagePickBwd pick_stages_bwd(const tg_tiling &tl, const BwdGeom &g, size_t fixed_bytes, int want_stages, bool semi = false) {
    StagePickBwd best{0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double best_score = -1.0;
    for (int ns = want_stages; ns >= 1; --ns) {
        int cap_rows = 0, kin = 0, kout = 0, staged = 0;
        std::vector<std::pair<size_t, int>> need;
        for (int t = 0; t < tl.num_tiles; ++t) {
            const tg_tile_meta &m = tl.h_meta[t];
            if (m.eligible)
                need.push_back({semi ? g.stage_semi(m.hi - m.lo, m.kin_kout & 0xFFFF, m.kin_kout >> 16)
                                     : g.stage(m.hi - m.lo, m.kin_kout & 0xFFFF, m.kin_kout >> 16), t});
        }
        std::sort(need.begin(), need.end());
        for (auto &nt : need) {
            const tg_tile_meta &m = tl.h_meta[nt.second];
            const int r = std::max(cap_rows, std::max(m.hi - m.lo, g.T)), ki = std::max(kin, m.kin_kout & 0xFFFF), ko = std::max(kout, m.kin_kout >> 16);
            const size_t total = semi ? fixed_bytes + ns * g.stage_semi(r, ki, ko) + 256
                                      : fixed_bytes + g.win_f(r) /* the single y window */ + ns * g.stage(r, ki, ko) + 256;
            if (total > size_t(kEdgeSmemBudget)) break;
            cap_rows = r; kin = ki; kout = ko;
            ++staged;
        }
        if (!staged) continue;
        const double frac = double(staged) / tl.num_tiles;
        const double score = frac * (ns >= 2 ? 1.0 : 0.6);
        if (score > best_score) {
            best_score = score;
            best.num_stages = ns;
            best.cap_rows = cap_rows;
            best.cap_kin = kin;
            best.cap_kout = kout;
            best.off_statraw = (uint32_t)g.slab(kin, kout);
            best.off_ds = best.off_statraw + (uint32_t)g.win_stat(cap_rows);
            best.off_xl = best.off_ds + (uint32_t)g.ds(cap_rows);
            best.off_xr = best.off_xl + (uint32_t)g.win_st(cap_rows);
            best.off_g = best.off_xr + (uint32_t)g.win_st(cap_rows);
            best.stage_bytes = (uint32_t)(semi ? g.stage_semi(cap_rows, kin, kout) : g.stage(cap_rows, kin, kout));
        }
        if (frac >= 0.9) break;
    }
    return best;
}

static int bwd_grid(const tecgat_plan_t *plan, int32_t snapshots) {
    const int64_t items = int64_t(plan->bwd.num_tiles) * snapshots;
    return (int)std::min<int64_t>(items, tg_sm_count());
}

static int bwd_max_flushes(const tecgat_plan_t *plan, int32_t snapshots) {  // partial rows per CTA
    const int g = bwd_grid(plan, snapshots);
    int64_t max_items = 0;
    tg_item_bounds(plan, true, snapshots, g, &max_items);
    const int64_t items = int64_t(plan->bwd.num_tiles) * snapshots;
    max_items = std::max(max_items, (items + g - 1) / g);  // the equal-count fallback
    return (int)(max_items / kFlushItems + 1);
}

struct BwdGeomCached {  // everything launch_bwd derives from (tiling, heads, channels, dtype): cached in the plan
    int32_t npw, semi, num_stages, cap_rows, cap_kin, cap_kout, per_st, per_f, per_stat, all_staged;
    uint32_t stage_bytes, off_meta, off_red, off_stage0, off_y, off_out, off_statraw, off_ds, off_xl, off_xr, off_g, smem;
};

template <int C, typename ST, bool VEC, int HT = 0>
static int launch_bwd(EdgeBwdArgs a, const tecgat_plan_t *plan, int grid, cudaStream_t st) {
    const tg_tiling &tl = plan->bwd;
    const int H = a.H, HC = H * C, T = tl.T;
    const int hp = pad_heads(H);
    const char *env = tg_env("TECGAT_BWD_STAGES");
    const int want = env ? std::max(1, std::min(kMaxStages, atoi(env))) : 2;
    const bool force_semi = tg_env("TECGAT_EDGE_SEMI") != nullptr, nostage = tg_env("TECGAT_EDGE_NOSTAGE") != nullptr;  // tests
    const uint64_t key = (uint64_t(2) << 56) | (uint64_t(C) << 40) | (uint64_t(H) << 32) | (uint64_t(sizeof(ST)) << 24) |
                         (uint64_t(VEC) << 16) | (uint64_t(want) << 8) | (uint64_t(force_semi) << 1) | uint64_t(nostage);
    BwdGeomCached c;
    bool hit = false;
    {
        std::lock_guard<std::mutex> lk(plan->cache_mu);
        auto it = plan->geom_cache.find(key);
        if (it != plan->geom_cache.end()) {
            memcpy(&c, it->second.data(), sizeof(c));
            hit = true;
        }
    }
    if (!hit) {
        c.npw = 32 / hp;
        const int ncw0 = T / c.npw;
        BwdGeom g{T, H, HC, sizeof(ST), row_period(uint32_t(HC * sizeof(ST))), row_period(uint32_t(HC * 4)), row_period(uint32_t(H * 4))};
        c.per_st = g.per_st; c.per_f = g.per_f; c.per_stat = g.per_stat;
        const size_t out_bytes = (2 * size_t(T) * HC * sizeof(ST) + 15) & ~size_t(15);
        const size_t red_bytes = (size_t(ncw0) * H * 2 * C * sizeof(float) + 15) & ~size_t(15);
        const size_t meta_bytes = tl.num_tiles <= kMetaSmemTiles ? size_t(tl.num_tiles) * 32 : 0;
        const size_t fixed = 128 + meta_bytes + red_bytes + out_bytes;
        StagePickBwd sp = pick_stages_bwd(tl, g, fixed, want);
        c.semi = 0;
        {   // rows too wide for full staging (e.g. H*C = 44 on the 300 km graph): stage slab + stat + delta planes only
            int full_tiles = 0;
            for (int t = 0; t < tl.num_tiles; ++t) {
                const tg_tile_meta &m = tl.h_meta[t];
                full_tiles += sp.num_stages > 0 && m.eligible && m.hi - m.lo <= sp.cap_rows && (m.kin_kout & 0xFFFF) <= sp.cap_kin &&
                              (m.kin_kout >> 16) <= sp.cap_kout;
            }
            if ((2 * full_tiles < tl.num_tiles || force_semi) && !nostage) {
                const StagePickBwd ss = pick_stages_bwd(tl, g, fixed, want, true);
                if (ss.num_stages > 0) {
                    sp = ss;
                    c.semi = 1;
                }
            }
        }
        c.num_stages = sp.num_stages > 0 ? sp.num_stages : 1;
        c.cap_rows = sp.cap_rows;
        c.cap_kin = sp.num_stages > 0 ? sp.cap_kin : -1;
        if (nostage) c.cap_kin = -1;  // tests: force the gather-from-global path
        c.cap_kout = sp.cap_kout;
        const size_t ybytes = c.semi ? 16 : g.win_f(c.cap_rows);
        c.off_meta = 128;
        c.off_red = (uint32_t)(128 + meta_bytes);
        c.off_y = (uint32_t)(c.off_red + red_bytes);
        c.off_out = (uint32_t)(c.off_y + ybytes);
        c.off_stage0 = (uint32_t)((c.off_out + out_bytes + 127) & ~size_t(127));
        c.stage_bytes = sp.stage_bytes;
        c.off_statraw = sp.off_statraw; c.off_ds = sp.off_ds; c.off_xl = sp.off_xl; c.off_xr = sp.off_xr; c.off_g = sp.off_g;
        c.smem = (uint32_t)(c.off_stage0 + size_t(sp.num_stages) * c.stage_bytes);
        bool all_staged = c.cap_kin >= 0;
        for (int t = 0; t < tl.num_tiles && all_staged; ++t) {
            const tg_tile_meta &m = tl.h_meta[t];
            all_staged = m.eligible && std::max(m.hi - m.lo, 0) <= c.cap_rows && (m.kin_kout & 0xFFFF) <= c.cap_kin && (m.kin_kout >> 16) <= c.cap_kout;
        }
        c.all_staged = all_staged;
        std::lock_guard<std::mutex> lk(plan->cache_mu);
        auto &blob = plan->geom_cache[key];
        blob.resize(sizeof(c));
        memcpy(blob.data(), &c, sizeof(c));
    }
    a.npw = c.npw; a.semi = c.semi; a.num_stages = c.num_stages; a.cap_rows = c.cap_rows; a.cap_kin = c.cap_kin; a.cap_kout = c.cap_kout;
    a.per_st = c.per_st; a.per_f = c.per_f; a.per_stat = c.per_stat;
    a.stage_bytes = c.stage_bytes; a.off_meta = c.off_meta; a.off_red = c.off_red; a.off_stage0 = c.off_stage0; a.off_y = c.off_y;
    a.off_out = c.off_out; a.off_statraw = c.off_statraw; a.off_ds = c.off_ds; a.off_xl = c.off_xl; a.off_xr = c.off_xr; a.off_g = c.off_g;
    const size_t smem = c.smem;
    const bool all_staged = c.all_staged != 0;
    const int ncw = T / a.npw;
    TG_REQUIRE(smem <= 227 * 1024, TECGAT_ENOSUP, "edge_bwd: %zu B shared memory needed (tile %d x %d channels)", smem, T, HC);
    const bool wg = ncw == 8;  // 8 consumer warps: only as the warpgroup-split kernels (specialised shapes)
    auto go = [&](auto kern) -> int {
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)smem));
        fill_schedule(a.sched, plan, true, a.S, grid);
        kern<<<(unsigned)grid, wg ? 384 : (ncw + 1) * 32, smem, st>>>(a);
        tg_count_launch();
        return TECGAT_OK;
    };
    int rc;
    if (wg) {
        rc = TECGAT_ENOSUP;
        if constexpr (HT > 0) {
#ifdef TG_TUNE_DEFAULT_ONLY
            if (!a.semi && all_staged && a.drop_thr != 0) rc = go(edge_bwd_kernel<C, ST, VEC, HT, false, false, true, true>);
#else
            if (a.semi) rc = go(edge_bwd_kernel<C, ST, VEC, HT, true, true, true, true>);
            else if (all_staged && a.drop_thr == 0) rc = go(edge_bwd_kernel<C, ST, VEC, HT, false, false, false, true>);
            else if (all_staged) rc = go(edge_bwd_kernel<C, ST, VEC, HT, false, false, true, true>);
            else rc = go(edge_bwd_kernel<C, ST, VEC, HT, false, true, true, true>);
#endif
        }
        if (rc == TECGAT_ENOSUP)
            tecgat_set_error("edge_bwd: a backward tile of %d nodes x %d heads (8 consumer warps) exists only for the specialised shapes "
                             "(heads = 2, out_channels 5 or 11); build the plan with tile_nodes_bwd = %d", T, H, 7 * a.npw);
    }
#ifndef TG_TUNE_DEFAULT_ONLY
    else if (a.semi) rc = go(edge_bwd_kernel<C, ST, VEC, HT, true, true, true>);
    else if (HT > 0 && all_staged && a.drop_thr == 0) rc = go(edge_bwd_kernel<C, ST, VEC, HT, false, HT == 0, HT == 0>);  // inference: no hash either
    else if (HT > 0 && all_staged) rc = go(edge_bwd_kernel<C, ST, VEC, HT, false, HT == 0, true>);  // compact: no gather code
    else rc = go(edge_bwd_kernel<C, ST, VEC, HT, false, true, true>);
#else
    else rc = TECGAT_ENOSUP;
#endif
    if (rc != TECGAT_OK) return rc;
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

}  // namespace tg

extern "C" int64_t tecgat_edge_bwd_workspace(const tecgat_plan_t *plan, int32_t snapshots, int32_t heads,
                                             int32_t out_channels) {
    if (!plan || snapshots <= 0 || heads <= 0 || out_channels <= 0) return 0;
    return int64_t(tg::bwd_grid(plan, snapshots)) * tg::bwd_max_flushes(plan, snapshots) * 2 * heads * out_channels * (int64_t)sizeof(float);
}

extern "C" int tecgat_edge_bwd(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att,
                               const float *bias, const float *y, const float *stat, const float *gy, void *dxl,
                               void *dxr, float *datt, float *dbias, void *workspace, int32_t snapshots, int32_t heads,
                               int32_t out_channels, float negative_slope, float dropout_p, uint64_t seed, int32_t mode,
                               int32_t dtype, void *stream) {
    TG_REQUIRE(datt && dbias, TECGAT_EINVAL, "edge_bwd: NULL argument");
    return tg::edge_bwd_run(plan, xl, xr, att, bias, y, stat, gy, dxl, dxr, datt, dbias, workspace, snapshots, heads, out_channels,
                            negative_slope, dropout_p, seed, nullptr, mode, dtype, stream, true, nullptr);
}

int tg::edge_bwd_run(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att, const float *bias, const float *y,
                     const float *stat, const float *gy, void *dxl, void *dxr, float *datt, float *dbias, void *workspace,
                     int32_t snapshots, int32_t heads, int32_t out_channels, float negative_slope, float dropout_p, uint64_t seed,
                     const uint64_t *seed_dev, int32_t mode, int32_t dtype, void *stream, bool reduce, int64_t *partial_rows) {
    using namespace tg;
    TG_REQUIRE(plan && xl && xr && att && bias && y && stat && gy && dxl && dxr && workspace,
               TECGAT_EINVAL, "edge_bwd: NULL argument");
    TG_REQUIRE(snapshots > 0 && heads > 0 && out_channels > 0, TECGAT_EINVAL, "edge_bwd: non-positive size");
    TG_REQUIRE(heads <= 32, TECGAT_ENOSUP, "edge_bwd: heads %d > 32", heads);
    TG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, TECGAT_EINVAL, "edge_bwd: dropout_p %f outside [0, 1)", dropout_p);
    TG_REQUIRE(mode == TECGAT_MODE_SHARED || mode == TECGAT_MODE_LITERAL, TECGAT_EINVAL, "edge_bwd: bad mode %d", mode);
    TG_REQUIRE(dtype == TECGAT_F32 || dtype == TECGAT_BF16, TECGAT_EINVAL, "edge_bwd: bad dtype %d", dtype);
    const int hp = pad_heads(heads);
    const tg_tiling &tl = plan->bwd;
    TG_REQUIRE(tl.T % (32 / hp) == 0 && tl.T * hp <= 256, TECGAT_ENOSUP,
               "edge_bwd: backward tile of %d nodes x %d heads does not map onto <= 8 consumer warps; build the plan with "
               "tile_nodes_bwd = a multiple of %d and <= %d", tl.T, heads, 32 / hp, 256 / hp);
    const int HC = heads * out_channels;
    TG_REQUIRE(2 * HC <= tl.T * hp, TECGAT_ENOSUP, "edge_bwd: 2*heads*out_channels (%d) exceeds the consumer threads (%d)", 2 * HC, tl.T * hp);
    const bool vec = (HC % 2) == 0;
    {
        auto al = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
        TG_REQUIRE(al(xl) && al(xr) && al(dxl) && al(dxr) && al(y) && al(gy) && al(stat), TECGAT_EINVAL,
                   "edge_bwd: xl / xr / dxl / dxr / y / gy / stat must be 16-byte aligned");
    }
    EdgeBwdArgs a;
    a.xl = xl; a.xr = xr; a.att = att; a.bias = bias; a.y = y; a.stat = stat; a.gy = gy;
    a.dxl = dxl; a.dxr = dxr; a.partials = static_cast<float *>(workspace);
    a.meta = tl.meta; a.slabs = tl.slabs;
    a.rowptr_in = plan->rowptr_in; a.col_in = plan->col_in; a.rowptr_out = plan->rowptr_out; a.col_out = plan->col_out;
    a.slot_out = plan->slot_out;
    a.N = plan->num_nodes; a.T = tl.T; a.num_tiles = tl.num_tiles; a.S = snapshots; a.H = heads; a.npw = 0;
    a.slope = negative_slope;
    a.drop_thr = dropout_p > 0.f ? std::max(1u, dropout_threshold(dropout_p)) : 0u;
    a.inv_keep = 1.f / (1.f - dropout_p);
    a.stream_stride = dropout_stream_stride(plan->num_edges);
    a.seed = seed;
    a.seed_dev = seed_dev;
    a.literal = (mode == TECGAT_MODE_LITERAL);
    a.items = int64_t(tl.num_tiles) * snapshots;
    const int grid = bwd_grid(plan, snapshots);
    a.max_flushes = bwd_max_flushes(plan, snapshots);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = TECGAT_ENOSUP;
    {   // banded graphs at the reference's shapes: the sliding-window kernel (every edge's score evaluated once)
        bool used = false;
        rc = edge_bwd_sw_try(plan, xl, xr, att, bias, y, stat, gy, dxl, dxr, a.partials, grid, a.max_flushes, snapshots, heads, out_channels,
                             negative_slope, dropout_p, a.drop_thr, seed, seed_dev, mode, dtype, st, &used);
        if (rc != TECGAT_OK) return rc;
        if (used) {
            if (partial_rows) *partial_rows = int64_t(grid) * a.max_flushes;
            if (!reduce) return TECGAT_OK;
            ReduceSegs segs = {{datt, dbias, nullptr, nullptr}, {0, HC, 0, 0}, {HC, 2 * HC, 0, 0}};
            return reduce_columns(a.partials, int64_t(grid) * a.max_flushes, 2 * HC, segs, st);
        }
        rc = TECGAT_ENOSUP;
    }
#define TG_CASE(CC)                                                                                                          \
    case CC:                                                                                                                 \
        if (vec) rc = dtype == TECGAT_F32 ? launch_bwd<CC, float, true>(a, plan, grid, st) : launch_bwd<CC, __nv_bfloat16, true>(a, plan, grid, st); \
        else if constexpr ((CC % 2) == 1) rc = dtype == TECGAT_F32 ? launch_bwd<CC, float, false>(a, plan, grid, st) : launch_bwd<CC, __nv_bfloat16, false>(a, plan, grid, st); \
        break;
#ifdef TG_TUNE_DEFAULT_ONLY  // tools/tune_edge_bwd.sh: compile the default workload's kernel alone (seconds instead of minutes)
    if (heads == 2 && vec && out_channels == 11 && dtype == TECGAT_F32) rc = launch_bwd<11, float, true, 2>(a, plan, grid, st);
    else return TECGAT_ENOSUP;
#else
    if (heads == 2 && vec && (out_channels == 11 || out_channels == 5)) {  // compile-time heads for the reference's shapes
        if (out_channels == 11) rc = dtype == TECGAT_F32 ? launch_bwd<11, float, true, 2>(a, plan, grid, st) : launch_bwd<11, __nv_bfloat16, true, 2>(a, plan, grid, st);
        else rc = dtype == TECGAT_F32 ? launch_bwd<5, float, true, 2>(a, plan, grid, st) : launch_bwd<5, __nv_bfloat16, true, 2>(a, plan, grid, st);
    } else
    switch (out_channels) {
        TG_FOR_EACH_C(TG_CASE)
        default:
            tecgat_set_error("edge_bwd: out_channels=%d is not among the compiled channel counts", out_channels);
            return TECGAT_ENOSUP;
    }
#endif
#undef TG_CASE
    if (rc != TECGAT_OK) return rc;
    if (partial_rows) *partial_rows = int64_t(grid) * a.max_flushes;
    if (!reduce) return TECGAT_OK;
    ReduceSegs segs = {{datt, dbias, nullptr, nullptr}, {0, HC, 0, 0}, {HC, 2 * HC, 0, 0}};
    return reduce_columns(a.partials, int64_t(grid) * a.max_flushes, 2 * HC, segs, st);
}
