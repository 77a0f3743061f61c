// edge_fwd.cu -- fused GATv2 edge phase, forward: for every snapshot and destination node
//     e_ij = att . LeakyReLU(xl_j + xr_i),  alpha = softmax_j(e_ij),  y_i = sum_j alpha_ij q_ij xl_j + bias
// in ONE persistent kernel (PyG: gather, add, leaky_relu, mul, sum, scatter-max, exp, scatter-add, div, dropout, mul,
// scatter-add, bias = ~20 ATen launches and 4-6 materialised (S*E, H, C) tensors; SURVEY.md K4-K9, reached
// from /root/reference/src/model/modules.py:356).  Execution model, lane mapping and arithmetic: edge_common.cuh.
//
// Softmax without a running maximum: softmax is shift invariant, so the row's SELF-LOOP score (CSR slot 0, always
// present) is the shift -- one exp2 per edge, no rescaling of the accumulator.  A row whose sum leaves the safe fp32
// range (scores more than ~80 log2-units above the self loop: never on sane data) is redone with the exact maximum.
// Saved for backward: ONE float per (row, head), stat = shift + log2(sum), i.e. alpha_ij = exp2(e_ij - stat).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "edge_common.cuh"

namespace tg {

struct EdgeFwdArgs {
    const void *xl, *xr;
    const float *att, *bias;
    float *y, *stat;
    const tg_tile_meta *meta;
    ItemSchedule sched;  // which items each CTA takes (equal estimated work)
    const unsigned char *slabs;
    const int32_t *rowptr, *col;  // destination-sorted CSR (tiles that are not staged)
    int32_t N, T, num_tiles, S, H, npw;
    float slope, inv_keep;
    float *y2;     // optional second copy of the output rows inside a wider row-major tensor (row stride ld2 floats), or NULL
    int64_t ld2;
    uint32_t drop_thr;
    uint32_t stream_stride;  // dropout counter distance between consecutive (snapshot, head) streams
    uint64_t seed;
    const uint64_t *seed_dev;  // non-NULL: the seed lives in device memory (CUDA-graph replays draw fresh masks)
    int32_t literal;
    int32_t cap_rows, cap_k, num_stages;
    int32_t per;  // 16-byte row period of xl / xr (rows)
    // shared-memory map (bytes): [barriers 128][tile table][y staging][stage 0][stage 1]..;  stage: [slab][xr tile][xl window]
    uint32_t stage_bytes, off_meta, off_stage0, off_xr, off_xl, off_y;
    int64_t items;
};

// one (destination node, head): everything between "inputs are readable" and "y / stat are known".
// The self loop (CSR slot 0) is peeled: it uses the lane's own xl row and its score is the softmax shift.  The other
// slots are walked TWO per iteration, branch-free (slots past the lane's degree read a valid row and get weight 0), so
// the two edges' instruction streams interleave.
template <int C, typename ST, bool VEC, bool FAST, bool DROP>
__device__ __forceinline__ void fwd_lane(const int Ts /* slab row stride */, const EdgeFwdArgs &a, const CV<C> &attp, const CV<C> &attm, const float *bias_ptr,
                                         const ST *xr_chunk, const ST *xl_self /* own row, + h*C */,
                                         const ST *xl_lane /* window row 0 (FAST) or snapshot row 0, + h*C */, int HC, int par,
                                         const uint16_t *ell /* + node_l */, const int32_t *col /* + k0 */, int deg, int kmax_w,
                                         uint32_t slot0, uint32_t key, CV<C> &out, float &stat) {
    CV<C> xr_i, xl_i, acc;
    if (deg > 0) {
        cv_load<C, VEC>(xr_i, xr_chunk, par);
        cv_load<C, VEC>(xl_i, xl_self, par);
    } else {
        cv_zero(xr_i);
        cv_zero(xl_i);
    }
    const float wself = deg > 0 ? 1.f : 0.f;
    float shift, l = 0.f;
    {
        CV<C> s;
        shift = edge_score<C>(attp, attm, xl_i, xr_i, s);
    }
    const float e_self = shift;
    auto nbr = [&](int k) -> int {
        if (FAST) return (int)ell[k * Ts];
        return k < deg ? __ldg(col + k) : 0;
    };
    const uint32_t h0 = slot0 * kDropMul + key;  // hash input of slot k: h0 + k * kDropMul (one add per edge)
    auto keep = [&](float w, uint32_t hk) -> float {
        if (!DROP) return w;  // inference / p = 0 instantiation: no hash
        // no branch on "dropout off": threshold 0 keeps every edge and inv_keep is 1, so the loop body stays ONE basic
        // block and the two edges of an iteration interleave freely
        return dropout_finish(hk) >= a.drop_thr ? w * a.inv_keep : 0.f;
    };
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        {
            const float w = wself * fast_exp2(e_self - shift);
            l = w;
            const float wq = keep(w, h0);
            const float2 wq2 = splat(wq);
#pragma unroll
            for (int i = 0; i < CV<C>::NP; ++i) acc.p[i] = __fmul2_rn(wq2, xl_i.p[i]);
            acc.s = wq * xl_i.s;
        }
        uint32_t hk = h0 + kDropMul;
#pragma unroll 1
        for (int k = 1; k < kmax_w; k += 2, hk += 2u * kDropMul) {
            const int ja = nbr(k), jb = nbr(k + 1);
            CV<C> xa, xb, sa, sb;
            cv_load<C, VEC>(xa, xl_lane + (FAST ? (ptrdiff_t)(ja * HC) : (ptrdiff_t)ja * HC), par);
            cv_load<C, VEC>(xb, xl_lane + (FAST ? (ptrdiff_t)(jb * HC) : (ptrdiff_t)jb * HC), par);
            const float ea = edge_score<C>(attp, attm, xa, xr_i, sa);
            const float eb = edge_score<C>(attp, attm, xb, xr_i, sb);
            // branch-free validity: slots past the degree read a real row (finite score), weight 0.  The clamp keeps
            // 0 * exp2(.) finite there and still trips the overflow guard (2^100 > kOverflowGuard) for real edges.
            const float va = k < deg ? 1.f : 0.f, vb = k + 1 < deg ? 1.f : 0.f;
            const float wa = va * fast_exp2(fminf(ea - shift, 100.f));
            const float wb = vb * fast_exp2(fminf(eb - shift, 100.f));
            l += wa + wb;
            const float2 qa = splat(keep(wa, hk)), qb = splat(keep(wb, hk + kDropMul));
#pragma unroll
            for (int i = 0; i < CV<C>::NP; ++i) acc.p[i] = __ffma2_rn(qb, xb.p[i], __ffma2_rn(qa, xa.p[i], acc.p[i]));
            if (CV<C>::ODD) acc.s = fmaf(qb.x, xb.s, fmaf(qa.x, xa.s, acc.s));
        }
        const bool ok = l < kOverflowGuard;  // false for inf / nan as well
        if (__all_sync(0xFFFFFFFFu, ok) || attempt == 1) break;
        float mx = e_self;  // rare: exact maximum, then redo the row
        for (int k = 1; k < deg; ++k) {
            CV<C> xj, s;
            const int j = nbr(k);
            cv_load<C, VEC>(xj, xl_lane + (FAST ? (ptrdiff_t)(j * HC) : (ptrdiff_t)j * HC), par);
            mx = fmaxf(mx, edge_score<C>(attp, attm, xj, xr_i, s));
        }
        shift = mx;
    }
    const float inv = l > 0.f ? fast_rcp(l) : 0.f;  // MUFU.RCP (1 ulp): the IEEE sequence is a range check + two Newton steps
    const float2 inv2 = splat(inv);
    CV<C> bias_h;  // read here (L1 hit) instead of living in 11 registers through the edge loop
    cv_load_param<C>(bias_h, bias_ptr, par, 1.f);
#pragma unroll
    for (int i = 0; i < CV<C>::NP; ++i) out.p[i] = __ffma2_rn(acc.p[i], inv2, bias_h.p[i]);
    out.s = fmaf(acc.s, inv, bias_h.s);
    stat = l > 0.f ? shift + fast_log2(l) : 0.f;
}

// HT > 0: compile-time number of heads (address arithmetic folds); 0: runtime
// GATHER: the launch contains tiles that are not staged (false drops the gather-from-global code: every item is staged).
// TT > 0: compile-time tile size (nodes) -- the default 15 consumer warps x 32 / heads: the slab strides become immediates
// WIDE: the output rows are also stored into a wider row-major tensor (a.y2, head pairs of a layer with more than two heads);
// a separate instantiation so that the default kernels carry none of it
template <int C, typename ST, bool VEC, int HT, bool GATHER, bool DROP, int TT = 0, bool WIDE = false>
__global__ void __launch_bounds__(512, 1) edge_fwd_kernel(const __grid_constant__ EdgeFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + kMaxStages;
    const tg_tile_meta *meta_s = reinterpret_cast<const tg_tile_meta *>(smem + a.off_meta);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = TT > 0 ? 15 : (blockDim.x >> 5) - 1;  // consumer warps; the last warp is the producer
    const int H = HT > 0 ? HT : a.H, HC = H * C, T = TT > 0 ? TT : a.T, N = a.N;
    const int Ts = (T + 7) & ~7;  // row stride of the slab sections
    const int NS = a.num_stages;
    constexpr uint32_t ES = sizeof(ST);
    const uint32_t RB = (uint32_t)HC * ES;  // bytes of one xl / xr row
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ncw);
        }
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

    if (warp == ncw) {
        // ================================ producer warp ================================
        for (int w = 0; w < n_items; ++w) {
            const tg_tile_meta m = load_meta(meta_s, a.meta, a.num_tiles, tile);
            const int n0 = tile * T, nt = min(N, n0 + T) - n0;
            const bool lit = a.literal && snap > 0;
            const int lo = lit ? n0 : m.lo, win = lit ? nt : m.hi - m.lo;
            const bool staged = !GATHER || (m.eligible && win <= a.cap_rows && (m.kin_kout & 0xFFFF) <= a.cap_k);
            if (staged) {
                unsigned char *stage = smem + a.off_stage0 + (size_t)ring.st * a.stage_bytes;
                const WinCopy cr = win_copy(a.xr, (int64_t)snap * N + n0, nt, RB, a.per, Rtot);
                const WinCopy cl = win_copy(a.xl, (int64_t)snap * N + lo, win, RB, a.per, Rtot);
                if (lane == 0) mbar_wait_relaxed(&empty[ring.st], ring.ph ^ 1u);
                __syncwarp();
                if (cr.tail | cl.tail) {  // only the last rows of the last snapshot
                    win_copy_tail(cr, stage + a.off_xr, lane);
                    win_copy_tail(cl, stage + a.off_xl, lane);
                    __syncwarp();
                }
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[ring.st], (uint32_t)m.slab_bytes + cr.mid + cl.mid);
                    bulk_g2s(stage, a.slabs + m.slab_off, (uint32_t)m.slab_bytes, &full[ring.st]);
                    if (cr.mid) bulk_g2s(stage + a.off_xr, cr.src, cr.mid, &full[ring.st]);
                    if (cl.mid) bulk_g2s(stage + a.off_xl, cl.src, cl.mid, &full[ring.st]);
                }
                ring.advance(NS);
            }
            if (++tile == a.num_tiles) { tile = 0; ++snap; }
        }
        return;
    }

    // ================================ consumer warps ================================
    const int npw = HT > 0 ? 32 / pad_heads(HT) : a.npw;                 // nodes per warp = 32 / padded heads
    const int nw = lane & (npw - 1);       // node within the warp (npw is a power of two)
    const int h = lane / npw;              // head (>= H: padding lane)
    const int node_l = warp * npw + nw;
    const bool head_ok = h < H;
    const int hh = head_ok ? h : 0;
    const int par = VEC ? ((hh * C) & 1) : 0;
    CV<C> attp, attm;
    cv_load_param<C>(attp, a.att + hh * C, par, 0.5f * (1.f + a.slope) * kLog2e);
    cv_load_param<C>(attm, a.att + hh * C, par, 0.5f * (1.f - a.slope) * kLog2e);
    const float *bias_h = a.bias + hh * C;
    float *ybuf = reinterpret_cast<float *>(smem + a.off_y) + warp * npw * HC;
    uint32_t key = 0;
    int key_snap = -1;
    const uint32_t drop_base = dropout_base((DROP && a.seed_dev) ? __ldg(a.seed_dev) : a.seed);

    for (int w = 0; w < n_items; ++w) {
        const tg_tile_meta m = load_meta(meta_s, a.meta, a.num_tiles, tile);
        const int n0 = tile * T, nt = min(N, n0 + T) - n0;
        const bool lit = a.literal && snap > 0;
        const int lo = lit ? n0 : m.lo, win = lit ? nt : m.hi - m.lo;
        const bool staged = !GATHER || (m.eligible && win <= a.cap_rows && (m.kin_kout & 0xFFFF) <= a.cap_k);
        const bool active = head_ok && node_l < nt;
        const int64_t row = (int64_t)snap * N + n0 + node_l;
        if (DROP && a.drop_thr && snap != key_snap) {
            key = dropout_key(drop_base, (uint32_t)snap, (uint32_t)H, (uint32_t)hh, a.stream_stride);
            key_snap = snap;
        }
        CV<C> out;
        float stat = 0.f;
        if (staged) {
            const unsigned char *stage = smem + a.off_stage0 + (size_t)ring.st * a.stage_bytes;
            const int32_t *k0s = reinterpret_cast<const int32_t *>(stage + 16);
            const int32_t *degs = k0s + Ts;
            const uint16_t *ell = reinterpret_cast<const uint16_t *>(degs + Ts) + node_l;
            const ST *xr_s = reinterpret_cast<const ST *>(stage + a.off_xr + win_skip((int64_t)snap * N + n0, RB, a.per));
            const ST *xl_s = reinterpret_cast<const ST *>(stage + a.off_xl + win_skip((int64_t)snap * N + lo, RB, a.per));
            mbar_wait(&full[ring.st], ring.ph);
            int deg = active ? (degs[node_l] & 0xFFFF) : 0;
            if (lit) deg = min(deg, 1);
            const uint32_t slot0 = active ? (uint32_t)k0s[node_l] : 0u;
            const int kmax_w = __reduce_max_sync(0xFFFFFFFFu, deg);
            fwd_lane<C, ST, VEC, true, DROP>(Ts, a, attp, attm, bias_h, xr_s + node_l * HC + hh * C, xl_s + (n0 + node_l - lo) * HC + hh * C,
                                       xl_s + hh * C, HC, par, ell, nullptr, deg, kmax_w, slot0, key, out, stat);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ring.st]);  // this warp no longer reads the stage
            ring.advance(NS);
            // stage y through the warp's private buffer, then write the warp's rows as one contiguous run
            if (active) cv_store<C, VEC>(ybuf + nw * HC + hh * C, out, par);
            __syncwarp();
            const int nv = max(0, min(npw, nt - warp * npw));  // valid nodes of this warp
            float *y_g = a.y + ((int64_t)snap * N + n0 + warp * npw) * HC;
            if (VEC) {
                const float2 *src = reinterpret_cast<const float2 *>(ybuf);
                float2 *dst = reinterpret_cast<float2 *>(y_g);
                if (HT > 0 && nv == npw) {  // full warp slice, compile-time trip count: straight-line copy
                    constexpr int kWords = HT > 0 ? (32 / pad_heads(HT > 0 ? HT : 1)) * (HT > 0 ? HT : 1) * C / 2 : 0;
#pragma unroll
                    for (int i = 0; i < (kWords + 31) / 32; ++i)
                        if (i * 32 + lane < kWords) dst[i * 32 + lane] = src[i * 32 + lane];
                } else {
                    for (int i = lane; i < nv * HC / 2; i += 32) dst[i] = src[i];
                }
            } else {
                for (int i = lane; i < nv * HC; i += 32) y_g[i] = ybuf[i];
            }
            if constexpr (WIDE) {  // the same rows into the caller's wider tensor (one HC-float run per row)
                float *y_w = a.y2 + ((int64_t)snap * N + n0 + warp * npw) * a.ld2;
                if (VEC) {
                    const int hw = HC / 2;
                    for (int i = lane; i < nv * hw; i += 32) {
                        const int r = i / hw, c = i - r * hw;
                        reinterpret_cast<float2 *>(y_w + (int64_t)r * a.ld2)[c] = reinterpret_cast<const float2 *>(ybuf)[i];
                    }
                } else {
                    for (int i = lane; i < nv * HC; i += 32) {
                        const int r = i / HC, c = i - r * HC;
                        y_w[(int64_t)r * a.ld2 + c] = ybuf[i];
                    }
                }
            }
            __syncwarp();
        } else if constexpr (GATHER) {
            // window or degree too large for a stage: gather straight from global memory (L2)
            int deg = 0, k0 = 0;
            if (active) {
                k0 = __ldg(a.rowptr + n0 + node_l);
                deg = __ldg(a.rowptr + n0 + node_l + 1) - k0;
                if (lit) deg = min(deg, 1);
            }
            const int kmax_w = __reduce_max_sync(0xFFFFFFFFu, deg);
            const ST *xl_snap = static_cast<const ST *>(a.xl) + (int64_t)snap * N * HC + hh * C;
            const ST *xr_chunk = static_cast<const ST *>(a.xr) + row * HC + hh * C;
            fwd_lane<C, ST, VEC, false, DROP>(Ts, a, attp, attm, bias_h, xr_chunk, xl_snap + (int64_t)(n0 + node_l) * HC, xl_snap, HC, par, nullptr,
                                        a.col + k0, deg, kmax_w, (uint32_t)k0, key, out, stat);
            if (active) cv_store<C, VEC>(a.y + row * HC + hh * C, out, par);
            if constexpr (WIDE) {
                if (active) cv_store<C, VEC>(a.y2 + row * a.ld2 + hh * C, out, par);
            }
        }
        if (active) a.stat[row * H + hh] = stat;
        if (++tile == a.num_tiles) { tile = 0; ++snap; }
    }
}

// Choose how many ring stages and which tiles are staged: a tile is staged when its window and degree fit the stage.
// Prefers 2+ stages (prefetch overlaps compute); falls back to 1 stage when that stages far more tiles.
struct StagePick {
    int num_stages, cap_rows, cap_k;
    uint32_t stage_bytes, off_xr, off_xl;
};
static StagePick pick_stages(const tg_tiling &tl, int T, int HC, size_t es, int per, size_t fixed_bytes, int want_stages) {
    const int Ts = (T + 7) & ~7;
    StagePick best{0, 0, 0, 0, 0, 0};
    double best_score = -1.0;
    for (int ns = want_stages; ns >= 1; --ns) {
        const size_t budget = (size_t(kEdgeSmemBudget) - fixed_bytes) / ns;
        // candidate caps: every distinct (window, kin) of the tiles, largest first; take the largest that fits
        int cap_rows = 0, cap_k = 0, staged = 0;
        std::vector<std::pair<size_t, int>> need;  // (bytes, tile)
        auto stage_bytes = [&](int rows, int k) {
            const int slack = 2 * (per - 1);  // the copies are widened to the 16-byte row period
            return size_t(round16(16 + 8 * Ts + 2 * Ts * k)) + round16(uint32_t((T + slack) * HC * es)) + round16(uint32_t((rows + slack) * HC * es));
        };
        for (int t = 0; t < tl.num_tiles; ++t)
            if (tl.h_meta[t].eligible) need.push_back({stage_bytes(tl.h_meta[t].hi - tl.h_meta[t].lo, tl.h_meta[t].kin_kout & 0xFFFF), t});
        std::sort(need.begin(), need.end());
        for (auto &nt : need) {  // grow the caps tile by tile (cheapest first) while the joint stage still fits
            const tg_tile_meta &m = tl.h_meta[nt.second];
            const int r = std::max(cap_rows, std::max(m.hi - m.lo, T)), k = std::max(cap_k, m.kin_kout & 0xFFFF);
            if (stage_bytes(r, k) > budget) break;
            cap_rows = r;
            cap_k = k;
            ++staged;
        }
        if (!staged) continue;
        const double frac = double(staged) / tl.num_tiles;
        const double score = frac * (ns >= 2 ? 1.0 : 0.6);  // a single stage cannot overlap load and compute
        if (score > best_score) {
            best_score = score;
            best.num_stages = ns;
            best.cap_rows = cap_rows;
            best.cap_k = cap_k;
            best.off_xr = round16(16 + 8 * Ts + 2 * Ts * cap_k);
            best.off_xl = best.off_xr + round16(uint32_t((T + 2 * (per - 1)) * HC * es));
            best.stage_bytes = (uint32_t)stage_bytes(cap_rows, cap_k);
        }
        if (frac >= 0.9) break;
    }
    return best;
}

struct FwdGeom {  // everything launch_fwd derives from (tiling, heads, channels, dtype): cached in the plan
    int32_t npw, per, num_stages, cap_rows, cap_k;
    uint32_t stage_bytes, off_meta, off_stage0, off_xr, off_xl, off_y;
    uint32_t smem;
    int32_t all_staged;
};

template <int C, typename ST, bool VEC, int HT = 0>
static int launch_fwd(EdgeFwdArgs a, const tecgat_plan_t *plan, cudaStream_t st) {
    const tg_tiling &tl = plan->fwd;
    const int HC = a.H * C, T = tl.T;
    const int hp = pad_heads(a.H);
    const char *env = tg_env("TECGAT_FWD_STAGES");  // tuning knob: ring depth wanted
    const int want = env ? std::max(1, std::min(kMaxStages, atoi(env))) : 3;
    const bool nostage = tg_env("TECGAT_EDGE_NOSTAGE") != nullptr;  // tests: force the gather-from-global path
    const uint64_t key = (uint64_t(1) << 56) | (uint64_t(C) << 40) | (uint64_t(a.H) << 32) | (uint64_t(sizeof(ST)) << 24) |
                         (uint64_t(VEC) << 16) | (uint64_t(want) << 8) | uint64_t(nostage);
    FwdGeom g;
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
        g.npw = 32 / hp;
        const size_t ybytes = size_t(T) * HC * sizeof(float);
        g.per = row_period(uint32_t(HC * sizeof(ST)));
        g.off_meta = 128;
        g.off_y = 128 + (tl.num_tiles <= kMetaSmemTiles ? tl.num_tiles * 32 : 0);
        g.off_stage0 = (uint32_t)((g.off_y + ybytes + 127) & ~size_t(127));
        const StagePick sp = pick_stages(tl, T, HC, sizeof(ST), g.per, g.off_stage0, want);
        g.num_stages = sp.num_stages > 0 ? sp.num_stages : 1;
        g.cap_rows = sp.cap_rows;
        g.cap_k = sp.num_stages > 0 ? sp.cap_k : -1;  // -1: nothing is staged
        if (nostage) g.cap_k = -1;
        g.stage_bytes = sp.stage_bytes;
        g.off_xr = sp.off_xr;
        g.off_xl = sp.off_xl;
        g.smem = (uint32_t)(g.off_stage0 + size_t(g.num_stages) * g.stage_bytes);
        bool all_staged = g.cap_k >= 0;
        for (int t = 0; t < tl.num_tiles && all_staged; ++t) {
            const tg_tile_meta &m = tl.h_meta[t];
            all_staged = m.eligible && m.hi - m.lo <= g.cap_rows && (m.kin_kout & 0xFFFF) <= g.cap_k;
        }
        g.all_staged = all_staged;
        std::lock_guard<std::mutex> lk(plan->cache_mu);
        auto &blob = plan->geom_cache[key];
        blob.resize(sizeof(g));
        memcpy(blob.data(), &g, sizeof(g));
    }
    a.npw = g.npw; a.per = g.per; a.num_stages = g.num_stages; a.cap_rows = g.cap_rows; a.cap_k = g.cap_k;
    a.stage_bytes = g.stage_bytes; a.off_meta = g.off_meta; a.off_stage0 = g.off_stage0; a.off_xr = g.off_xr; a.off_xl = g.off_xl;
    a.off_y = g.off_y;
    const size_t smem = g.smem;
    const bool all_staged = g.all_staged != 0;
    const int ncw = T / a.npw;  // consumer warps
    const int sms = tg_sm_count();
    auto go = [&](auto kern) -> int {
        TG_CUDA(tg_set_smem(reinterpret_cast<const void *>(kern), (int)smem));
        // one persistent CTA per SM (512 threads and > 113 KB of shared memory: never two)
        const int64_t grid = std::min<int64_t>(a.items, int64_t(sms));
        fill_schedule(a.sched, plan, false, a.S, (int)grid);
        kern<<<(unsigned)grid, (ncw + 1) * 32, smem, st>>>(a);
        tg_count_launch();
        return TECGAT_OK;
    };
    int rc;
    constexpr int kTT = HT > 0 ? 15 * (32 / pad_heads(HT > 0 ? HT : 1)) : 0;  // the default tile of the fixed-head kernels
    if (a.y2) {  // wide output: compiled for the fixed-head shapes only (dropout off = threshold 0 in the hashing kernels)
        rc = TECGAT_ENOSUP;
        if constexpr (HT > 0) {
            if (all_staged && T == kTT) rc = go(edge_fwd_kernel<C, ST, VEC, HT, false, true, kTT, true>);
            else if (all_staged) rc = go(edge_fwd_kernel<C, ST, VEC, HT, false, true, 0, true>);
            else rc = go(edge_fwd_kernel<C, ST, VEC, HT, true, true, 0, true>);
        }
        if (rc == TECGAT_ENOSUP) tecgat_set_error("edge_fwd: the wide output exists only for heads = 2, out_channels 5 or 11");
    } else
    if (HT > 0 && all_staged && T == kTT && a.drop_thr == 0) rc = go(edge_fwd_kernel<C, ST, VEC, HT, HT == 0, HT == 0, kTT>);
    else if (HT > 0 && all_staged && T == kTT) rc = go(edge_fwd_kernel<C, ST, VEC, HT, HT == 0, true, kTT>);
    else if (HT > 0 && all_staged && a.drop_thr == 0) rc = go(edge_fwd_kernel<C, ST, VEC, HT, HT == 0, HT == 0>);  // inference: no hash either
    else if (HT > 0 && all_staged) rc = go(edge_fwd_kernel<C, ST, VEC, HT, HT == 0, true>);  // compact: no gather code
    else rc = go(edge_fwd_kernel<C, ST, VEC, HT, true, true>);
    if (rc != TECGAT_OK) return rc;
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

}  // namespace tg

extern "C" int tecgat_edge_fwd(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att,
                               const float *bias, float *y, float *stat, int32_t snapshots, int32_t heads,
                               int32_t out_channels, float negative_slope, float dropout_p, uint64_t seed, int32_t mode,
                               int32_t dtype, void *stream) {
    return tg::edge_fwd_run(plan, xl, xr, att, bias, y, stat, snapshots, heads, out_channels, negative_slope, dropout_p, seed,
                            nullptr, mode, dtype, stream);
}

int tg::edge_fwd_run(const tecgat_plan_t *plan, const void *xl, const void *xr, const float *att, const float *bias, float *y,
                     float *stat, int32_t snapshots, int32_t heads, int32_t out_channels, float negative_slope, float dropout_p,
                     uint64_t seed, const uint64_t *seed_dev, int32_t mode, int32_t dtype, void *stream, float *y_wide,
                     int64_t ld_wide) {
    using namespace tg;
    TG_REQUIRE(!y_wide || (reinterpret_cast<uintptr_t>(y_wide) % 8 == 0 && ld_wide % 2 == 0 && ld_wide >= int64_t(heads) * out_channels),
               TECGAT_EINVAL, "edge_fwd: the wide output must be 8-byte aligned with an even row stride >= heads * out_channels");
    TG_REQUIRE(plan && xl && xr && att && bias && y && stat, TECGAT_EINVAL, "edge_fwd: NULL argument");
    TG_REQUIRE(snapshots > 0 && heads > 0 && out_channels > 0, TECGAT_EINVAL, "edge_fwd: non-positive size");
    TG_REQUIRE(heads <= 32, TECGAT_ENOSUP, "edge_fwd: heads %d > 32", heads);
    TG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, TECGAT_EINVAL, "edge_fwd: dropout_p %f outside [0, 1)", dropout_p);
    TG_REQUIRE(mode == TECGAT_MODE_SHARED || mode == TECGAT_MODE_LITERAL, TECGAT_EINVAL, "edge_fwd: bad mode %d", mode);
    TG_REQUIRE(dtype == TECGAT_F32 || dtype == TECGAT_BF16, TECGAT_EINVAL, "edge_fwd: bad dtype %d", dtype);
    const int hp = pad_heads(heads);
    const tg_tiling &tl = plan->fwd;
    TG_REQUIRE(tl.T % (32 / hp) == 0 && tl.T * hp <= 480, TECGAT_ENOSUP,
               "edge_fwd: forward tile of %d nodes x %d heads does not map onto <= 15 consumer warps; build the plan with "
               "tile_nodes_fwd = a multiple of %d and <= %d", tl.T, heads, 32 / hp, 480 / hp);
    const int HC = heads * out_channels;
    const bool vec = (HC % 2) == 0;
    TG_REQUIRE(reinterpret_cast<uintptr_t>(xl) % 16 == 0 && reinterpret_cast<uintptr_t>(xr) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(y) % 16 == 0,
               TECGAT_EINVAL, "edge_fwd: xl / xr / y must be 16-byte aligned");
    EdgeFwdArgs a;
    a.xl = xl; a.xr = xr; a.att = att; a.bias = bias; a.y = y; a.stat = stat;
    a.y2 = y_wide; a.ld2 = ld_wide;
    a.meta = tl.meta; a.slabs = tl.slabs;
    a.rowptr = plan->rowptr_in; a.col = plan->col_in;
    a.N = plan->num_nodes; a.T = tl.T; a.num_tiles = tl.num_tiles; a.S = snapshots; a.H = heads; a.npw = 0;
    a.slope = negative_slope;
    a.drop_thr = dropout_p > 0.f ? std::max(1u, dropout_threshold(dropout_p)) : 0u;
    a.inv_keep = 1.f / (1.f - dropout_p);
    a.stream_stride = dropout_stream_stride(plan->num_edges);
    a.seed = seed;
    a.seed_dev = seed_dev;
    a.literal = (mode == TECGAT_MODE_LITERAL);
    a.items = int64_t(tl.num_tiles) * snapshots;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TG_CASE(CC)                                                                                                   \
    case CC:                                                                                                          \
        if (vec) return dtype == TECGAT_F32 ? launch_fwd<CC, float, true>(a, plan, st) : launch_fwd<CC, __nv_bfloat16, true>(a, plan, st); \
        if constexpr ((CC % 2) == 1) return dtype == TECGAT_F32 ? launch_fwd<CC, float, false>(a, plan, st) : launch_fwd<CC, __nv_bfloat16, false>(a, plan, st); \
        break;
    if (heads == 2 && vec) {  // the reference's shapes (train.py:263-266, README variant): compile-time heads
        if (out_channels == 11) return dtype == TECGAT_F32 ? launch_fwd<11, float, true, 2>(a, plan, st) : launch_fwd<11, __nv_bfloat16, true, 2>(a, plan, st);
        if (out_channels == 5) return dtype == TECGAT_F32 ? launch_fwd<5, float, true, 2>(a, plan, st) : launch_fwd<5, __nv_bfloat16, true, 2>(a, plan, st);
    }
    switch (out_channels) {
        TG_FOR_EACH_C(TG_CASE)
        default:
            break;
    }
#undef TG_CASE
    tecgat_set_error("edge_fwd: out_channels=%d is not among the compiled channel counts", out_channels);
    return TECGAT_ENOSUP;
}
