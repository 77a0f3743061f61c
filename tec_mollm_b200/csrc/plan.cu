// plan.cu -- one-time graph plan: what PyG redoes on every forward (remove_self_loops + add_self_loops,
// boolean-mask nonzero with a device->host sync; GATv2Conv.forward called at
// /root/reference/src/model/modules.py:356) becomes a cached, immutable structure:
//   * destination-sorted CSR (stable: a destination's in-edges keep the input order; the self loop comes FIRST in every
//     row -- its score is the softmax shift -- so the summation order is not PyG's: parity is to tolerance, not by order),
//   * source-sorted CSR with, per out-slot, the in-CSR slot of the same edge (dropout counter),
//   * per tile of `tile_nodes` consecutive nodes, the window [lo, hi) of rows the tile touches either as
//     in-neighbours or out-neighbours -- the contiguous slab the edge kernels stage in shared memory.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

static thread_local std::string g_last_error;

void tecgat_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

extern "C" const char *tecgat_last_error(void) { return g_last_error.c_str(); }

#include <atomic>
#include <unordered_map>
static std::atomic<long long> g_launches{0};
void tg_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t tecgat_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int tg_sm_count() {
    static std::mutex mu;
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lk(mu);
    if (!cached[dev]) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cached[dev] = sms;
    }
    return cached[dev];
}

cudaError_t tg_set_smem(const void *kernel, int bytes) {
    static std::mutex mu;
    static std::unordered_map<uint64_t, int> have;  // (kernel, device) -> largest size already granted
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t key = reinterpret_cast<uint64_t>(kernel) * 67u + uint64_t(dev);
    std::lock_guard<std::mutex> lk(mu);
    auto it = have.find(key);
    if (it != have.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) have[key] = bytes;
    return e;
}

const char *tg_env(const char *name) {
    // the knobs are test / tuning switches that tests flip between calls (monkeypatch.setenv), so this is a plain getenv:
    // ~50 ns per lookup, and the launchers only consult it on a geometry-cache miss
    return getenv(name);
}
extern "C" int tecgat_abi_version(void) { return TECGAT_ABI_VERSION; }

template <typename T>
static int upload(T **dst, const std::vector<T> &src, cudaStream_t st) {
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    TG_CUDA(cudaMalloc(reinterpret_cast<void **>(dst), bytes));
    if (!src.empty()) TG_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return TECGAT_OK;
}

// Item schedule of a persistent edge kernel: CTA b of `grid` takes items [bounds[b], bounds[b + 1]) of the snapshot-major item
// list, ranges of equal estimated work (one snapshot's work = h_cost_prefix.back()).  TECGAT_SCHEDULE=count (A/B knob): equal
// item counts.
const int64_t *tg_item_bounds(const tecgat_plan_t *plan, bool bwd, int64_t S, int grid, int64_t *max_items) {
    const tg_tiling &tl = bwd ? plan->bwd : plan->fwd;
    const bool by_count = tg_env("TECGAT_SCHEDULE") != nullptr;
    const uint64_t key = (uint64_t(bwd) << 60) | (uint64_t(by_count) << 59) | (uint64_t(S) << 20) | uint64_t(grid);
    std::lock_guard<std::mutex> lk(plan->cache_mu);
    auto &e = plan->bounds_cache[key];
    if (e.host.empty()) {
        const int64_t W = tl.h_cost_prefix.back(), nt = tl.num_tiles;
        e.host.resize(size_t(grid) + 1);
        for (int b = 0; b <= grid; ++b) {
            if (b == grid || W <= 0 || by_count) {
                e.host[b] = b == grid ? S * nt : S * nt * b / grid;
                continue;
            }
            const int64_t X = S * W, u = X / grid * b + (X % grid) * b / grid;  // floor(S W b / grid)
            const int64_t snap = u / W, rem = u - snap * W;
            const int64_t t = std::upper_bound(tl.h_cost_prefix.begin() + 1, tl.h_cost_prefix.end(), rem) - (tl.h_cost_prefix.begin() + 1);
            e.host[b] = snap * nt + t;
        }
        for (int b = 0; b < grid; ++b) e.max_items = std::max(e.max_items, e.host[b + 1] - e.host[b]);
    }
    if (max_items) *max_items = e.max_items;
    return e.host.data();
}

static void free_tiling(tg_tiling &t) {
    cudaFree(t.meta);
    cudaFree(t.slabs);
    t.meta = nullptr;
    t.slabs = nullptr;
}

extern "C" int tecgat_plan_destroy(tecgat_plan_t *p) {
    if (!p) return TECGAT_OK;
    cudaFree(p->rowptr_in);
    cudaFree(p->col_in);
    cudaFree(p->rowptr_out);
    cudaFree(p->col_out);
    cudaFree(p->slot_out);
    free_tiling(p->fwd);
    free_tiling(p->bwd);
    if (p->sw) {
        cudaFree(p->sw->slabD);
        cudaFree(p->sw->slabS);
        delete p->sw;
    }
    free(p->h_rowptr_in);
    free(p->h_col_in);
    free(p->h_eid_in);
    delete p;
    return TECGAT_OK;
}

// Build one tiling (see tg_tiling in common.cuh).  Forward tilings window the in-neighbours only; backward tilings
// window in- and out-neighbours and carry the out-edge ELL + dropout slots as well.
static int build_tiling(tg_tiling &tl, bool bwd, int32_t T, int64_t N, const std::vector<int32_t> &rp_in,
                        const std::vector<int32_t> &col_in, const std::vector<int32_t> &rp_out,
                        const std::vector<int32_t> &col_out, const std::vector<int32_t> &slot_out, cudaStream_t st) {
    tl.T = T;
    tl.bwd = bwd;
    const int32_t Ts = (T + 7) & ~7;  // row stride of the slab sections: keeps every section 16-byte aligned
    tl.num_tiles = static_cast<int32_t>((N + T - 1) / T);
    tl.h_meta.resize(tl.num_tiles);
    tl.h_slab_off.assign(tl.num_tiles + 1, 0);
    tl.h_cost_prefix.assign(tl.num_tiles + 1, 0);
    for (int32_t t = 0; t < tl.num_tiles; ++t) {
        const int64_t n0 = int64_t(t) * T, n1 = std::min<int64_t>(N, n0 + T);
        int32_t l = static_cast<int32_t>(n0), h = static_cast<int32_t>(n1 - 1), kin = 0, kout = 0;
        for (int64_t n = n0; n < n1; ++n) {
            kin = std::max(kin, rp_in[n + 1] - rp_in[n]);
            kout = std::max(kout, rp_out[n + 1] - rp_out[n]);
        }
        for (int32_t k = rp_in[n0]; k < rp_in[n1]; ++k) {
            l = std::min(l, col_in[k]);
            h = std::max(h, col_in[k]);
        }
        if (bwd)
            for (int32_t k = rp_out[n0]; k < rp_out[n1]; ++k) {
                l = std::min(l, col_out[k]);
                h = std::max(h, col_out[k]);
            }
        // the kernels peel the self loop (slot 0) and walk the remaining slots two at a time: odd padded row counts
        kin |= 1;
        kout |= 1;
        if (!bwd) kout = 0;
        tg_tile_meta &m = tl.h_meta[t];
        m.lo = l;
        m.hi = h + 1;
        bool ok = (h + 1 - l) <= 65535 && kin <= 32767 && kout <= 32767;  // kin | kout << 16 and deg = in | out << 16 live in signed int32
        if (ok && bwd) ok = int64_t(rp_in[h + 1]) - rp_in[l] <= 65535;  // relative dropout slots fit uint16
        m.eligible = ok ? 1 : 0;
        m.kin_kout = ok ? (kin | (kout << 16)) : 0;
        tl.max_window = std::max(tl.max_window, h + 1 - l);
        {   // cost estimate in warp instructions per item (the lanes of a tile walk to the tile's longest row): per-item part +
            // per-slot part, from the kernels' SASS; tiles that can never be staged (gathered from L2) count double
            int64_t c = bwd ? 930 + 84 * int64_t(kin + kout) : 325 + 56 * int64_t(kin);
            if (!ok || (h + 1 - l) > 2048 || kin > 96) c *= 2;
            tl.h_cost_prefix[t + 1] = tl.h_cost_prefix[t] + c;
        }
        const int64_t bytes = ok ? 16 + 8 * int64_t(Ts) + 2 * int64_t(Ts) * (kin + 2 * kout) : 0;
        tl.h_slab_off[t + 1] = tl.h_slab_off[t] + ((bytes + 15) & ~int64_t(15));
        m.slab_off = tl.h_slab_off[t];
        m.slab_bytes = static_cast<int32_t>(tl.h_slab_off[t + 1] - tl.h_slab_off[t]);
        m.pad = 0;
    }
    std::vector<unsigned char> slabs(static_cast<size_t>(std::max<int64_t>(tl.h_slab_off[tl.num_tiles], 16)), 0);
    for (int32_t t = 0; t < tl.num_tiles; ++t) {
        const tg_tile_meta &m = tl.h_meta[t];
        if (!m.eligible) continue;
        const int32_t kin = m.kin_kout & 0xFFFF, kout = m.kin_kout >> 16;
        const int64_t n0 = int64_t(t) * T, n1 = std::min<int64_t>(N, n0 + T);
        unsigned char *base = slabs.data() + tl.h_slab_off[t];
        int32_t *hdr = reinterpret_cast<int32_t *>(base);
        int32_t *k0 = hdr + 4, *deg = k0 + Ts;
        uint16_t *ell_in = reinterpret_cast<uint16_t *>(deg + Ts);
        uint16_t *ell_out = ell_in + size_t(kin) * Ts, *slot = ell_out + size_t(kout) * Ts;
        hdr[0] = m.lo;
        hdr[1] = m.hi;
        hdr[2] = m.kin_kout;
        hdr[3] = rp_in[m.lo];
        for (int64_t n = n0; n < n1; ++n) {
            const int32_t i = static_cast<int32_t>(n - n0), di = rp_in[n + 1] - rp_in[n], dout = bwd ? rp_out[n + 1] - rp_out[n] : 0;
            k0[i] = rp_in[n];
            deg[i] = di | (dout << 16);
            for (int32_t k = 0; k < di; ++k) ell_in[size_t(k) * Ts + i] = static_cast<uint16_t>(col_in[rp_in[n] + k] - m.lo);
            for (int32_t k = 0; k < dout; ++k) {
                ell_out[size_t(k) * Ts + i] = static_cast<uint16_t>(col_out[rp_out[n] + k] - m.lo);
                slot[size_t(k) * Ts + i] = static_cast<uint16_t>(slot_out[rp_out[n] + k] - rp_in[m.lo]);
            }
        }
    }
    int rc;
    if ((rc = upload(&tl.meta, tl.h_meta, st)) || (rc = upload(&tl.slabs, slabs, st)))
        return rc;
    TG_CUDA(cudaStreamSynchronize(st));  // `slabs` is a local
    return TECGAT_OK;
}

// Sliding-window backward tiling (tg_sw_plan in common.cuh).  Returns with plan->sw == nullptr (and no error) when the graph
// does not qualify: an edge longer than T rows, degrees beyond the slab limits, or a padding entry with no zero stash slot.
static int build_sw(tecgat_plan_t *p, int32_t T, int64_t N, const std::vector<int32_t> &rp_in, const std::vector<int32_t> &col_in,
                    const std::vector<int32_t> &rp_out, const std::vector<int32_t> &col_out, const std::vector<int32_t> &slot_out,
                    cudaStream_t st) {
    p->sw = nullptr;
    if (T < 8 || (T % 8) != 0) return TECGAT_OK;
    int32_t R = 0, kin = 0, kout = 0;
    for (int64_t n = 0; n < N; ++n) {
        kin = std::max(kin, rp_in[n + 1] - rp_in[n]);
        kout = std::max(kout, rp_out[n + 1] - rp_out[n]);
        for (int32_t k = rp_in[n]; k < rp_in[n + 1]; ++k) R = std::max<int32_t>(R, std::abs(col_in[k] - int32_t(n)));
    }
    if (R > T || kin > 48 || kout > 48) return TECGAT_OK;
    tg_sw_plan *sw = new (std::nothrow) tg_sw_plan();
    TG_REQUIRE(sw != nullptr, TECGAT_ENOMEM, "plan_create: out of host memory");
    sw->T = T;
    sw->J = int32_t((N + T - 1) / T);
    sw->R = R;
    sw->kin = kin;
    sw->kout = kout;
    sw->kinp = std::max(2, kin & ~1);  // kin - 1 in-slots besides the self loop, rounded up to an even count
    sw->koutp = std::max(2, (kout + 1) & ~1);
    // stash row of a node: slots 0 .. kinp (slot kinp is the one the destination role zero-fills whenever it is not a real
    // edge) + 8 bytes so that consecutive nodes spread over the shared-memory banks (stride = 2 mod 4 words)
    sw->stash_stride = (sw->kinp + 1) * 16 + 8;
    const int32_t su = sw->stash_stride / 8;  // stash units (8 B) per node
    sw->slabD_bytes = int32_t((16 + 8 * T + 2 * T * sw->kinp + 15) & ~15);
    sw->slabS_bytes = int32_t((2 * T + 4 * T * sw->koutp + 15) & ~15);
    std::vector<unsigned char> sd(size_t(sw->J) * sw->slabD_bytes, 0), ss(size_t(sw->J) * sw->slabS_bytes, 0);
    bool ok = (int64_t(T) + 2 * R) * su <= 65535;
    for (int32_t c = 0; c < sw->J && ok; ++c) {
        const int64_t n0 = int64_t(c) * T, n1 = std::min<int64_t>(N, n0 + T);
        const int32_t wlo = int32_t(std::max<int64_t>(0, n0 - R)), whi = int32_t(std::min<int64_t>(N, n0 + T + R));
        int32_t *hdr = reinterpret_cast<int32_t *>(sd.data() + size_t(c) * sw->slabD_bytes);
        int32_t *k0 = hdr + 4, *deg = k0 + T;
        uint16_t *ell_in = reinterpret_cast<uint16_t *>(deg + T);
        uint16_t *dego = reinterpret_cast<uint16_t *>(ss.data() + size_t(c) * sw->slabS_bytes);  // out-degree per node
        uint16_t *ell_out = dego + T;
        uint16_t *st_out = ell_out + size_t(sw->koutp) * T;
        hdr[0] = wlo;
        hdr[1] = whi;
        // a stash slot that is zero for sure: slot kinp of any window node whose in-degree does not reach it
        int32_t zero_node = -1;
        for (int32_t u = wlo; u < whi && zero_node < 0; ++u)
            if (rp_in[u + 1] - rp_in[u] <= sw->kinp) zero_node = u;
        for (int64_t n = n0; n < n1; ++n) {
            const int32_t i = int32_t(n - n0), di = rp_in[n + 1] - rp_in[n], dout = rp_out[n + 1] - rp_out[n];
            k0[i] = rp_in[n];
            deg[i] = di | (dout << 16);
            dego[i] = uint16_t(dout);
            for (int32_t k = 1; k <= sw->kinp; ++k)
                ell_in[size_t(k - 1) * T + i] = uint16_t((k < di ? col_in[rp_in[n] + k] : int32_t(n)) - wlo);
            const int32_t own_zero = di <= sw->kinp ? int32_t(n) : zero_node;
            for (int32_t k = 0; k < sw->koutp; ++k) {
                if (k < dout) {
                    const int32_t u = col_out[rp_out[n] + k], slot = slot_out[rp_out[n] + k] - rp_in[u];  // in-slot of the edge at u
                    ell_out[size_t(k) * T + i] = uint16_t(u - wlo);
                    st_out[size_t(k) * T + i] = uint16_t((u - wlo) * su + slot * 2);
                } else {
                    if (own_zero < 0) { ok = false; break; }
                    ell_out[size_t(k) * T + i] = uint16_t(own_zero - wlo);  // a valid row; its contribution is multiplied by zero
                    st_out[size_t(k) * T + i] = uint16_t((own_zero - wlo) * su + sw->kinp * 2);
                }
            }
        }
        for (int64_t i = n1 - n0; i < T; ++i) {  // padding nodes of the last chunk: degree 0, entries point at the window start
            k0[i] = 0;
            deg[i] = 0;
        }
    }
    if (!ok) {
        delete sw;
        return TECGAT_OK;
    }
    int rc;
    if ((rc = upload(&sw->slabD, sd, st)) || (rc = upload(&sw->slabS, ss, st))) {
        cudaFree(sw->slabD);
        cudaFree(sw->slabS);
        delete sw;
        return rc;
    }
    TG_CUDA(cudaStreamSynchronize(st));  // the uploads read locals
    p->sw = sw;
    return TECGAT_OK;
}

extern "C" int tecgat_plan_create(const int64_t *edge_index_dev, int64_t num_edges, int32_t num_nodes,
                                  int32_t tile_nodes_fwd, int32_t tile_nodes_bwd, void *stream, tecgat_plan_t **plan_out) {
    TG_REQUIRE(plan_out != nullptr, TECGAT_EINVAL, "plan_create: plan_out is NULL");
    *plan_out = nullptr;
    TG_REQUIRE(num_nodes > 0, TECGAT_EINVAL, "plan_create: num_nodes must be positive (got %d)", num_nodes);
    TG_REQUIRE(num_edges >= 0, TECGAT_EINVAL, "plan_create: negative edge count");
    TG_REQUIRE(num_edges == 0 || edge_index_dev != nullptr, TECGAT_EINVAL, "plan_create: edge_index is NULL");
    for (int32_t t : {tile_nodes_fwd, tile_nodes_bwd})
        TG_REQUIRE(t >= 1 && t <= 512, TECGAT_EINVAL, "plan_create: tile_nodes %d outside [1, 512]", t);
    TG_REQUIRE(num_edges + num_nodes < (int64_t(1) << 31), TECGAT_ENOSUP, "plan_create: more than 2^31 edges");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t N = num_nodes, E0 = num_edges;

    std::vector<int64_t> ei(static_cast<size_t>(2 * E0));
    if (E0 > 0) {
        TG_CUDA(cudaMemcpyAsync(ei.data(), edge_index_dev, sizeof(int64_t) * 2 * E0, cudaMemcpyDeviceToHost, st));
        TG_CUDA(cudaStreamSynchronize(st));
    }
    const int64_t *src = ei.data(), *dst = ei.data() + E0;
    int64_t kept = 0;
    for (int64_t e = 0; e < E0; ++e) {
        if (src[e] < 0 || src[e] >= N || dst[e] < 0 || dst[e] >= N) {
            tecgat_set_error("plan_create: edge %lld = (%lld -> %lld) outside [0, %d)", (long long)e, (long long)src[e],
                             (long long)dst[e], num_nodes);
            return TECGAT_EINVAL;
        }
        kept += (src[e] != dst[e]);
    }
    const int64_t E = kept + N;

    tecgat_plan_t *p = new (std::nothrow) tecgat_plan_t();
    TG_REQUIRE(p != nullptr, TECGAT_ENOMEM, "plan_create: out of host memory");
    cudaGetDevice(&p->device);
    p->num_nodes = num_nodes;
    p->num_edges = E;
    p->kept_edges = kept;

    // ---- both CSR orientations (stable counting sort; every row STARTS with the node's self loop, then the kept
    //      edges in input order).  The self loop's score is the softmax shift of the row (edge_fwd.cu). -------------
    std::vector<int32_t> rp_in(N + 1, 0), rp_out(N + 1, 0);
    for (int64_t e = 0; e < E0; ++e)
        if (src[e] != dst[e]) {
            rp_in[dst[e] + 1]++;
            rp_out[src[e] + 1]++;
        }
    for (int64_t i = 0; i < N; ++i) {
        rp_in[i + 1] += 1;   // self loop
        rp_out[i + 1] += 1;
    }
    for (int64_t i = 0; i < N; ++i) {
        p->max_in_deg = std::max(p->max_in_deg, rp_in[i + 1]);
        p->max_out_deg = std::max(p->max_out_deg, rp_out[i + 1]);
        rp_in[i + 1] += rp_in[i];
        rp_out[i + 1] += rp_out[i];
    }
    std::vector<int32_t> col_in(E), eid_in(E), col_out(E), slot_out(E);
    {
        std::vector<int32_t> cur_in(rp_in.begin(), rp_in.end() - 1), cur_out(rp_out.begin(), rp_out.end() - 1);
        std::vector<int32_t> slot_of_edge(E);  // PyG edge id -> in-CSR slot
        for (int64_t i = 0; i < N; ++i) {
            const int32_t k = cur_in[i]++;
            col_in[k] = static_cast<int32_t>(i);
            eid_in[k] = static_cast<int32_t>(kept + i);  // PyG appends the self loops after the kept edges
            slot_of_edge[kept + i] = k;
            const int32_t k2 = cur_out[i]++;
            col_out[k2] = static_cast<int32_t>(i);
            slot_out[k2] = k;
        }
        int32_t id = 0;
        for (int64_t e = 0; e < E0; ++e)
            if (src[e] != dst[e]) {
                const int32_t k = cur_in[dst[e]]++;
                col_in[k] = static_cast<int32_t>(src[e]);
                eid_in[k] = id;
                slot_of_edge[id] = k;
                ++id;
            }
        id = 0;
        for (int64_t e = 0; e < E0; ++e)
            if (src[e] != dst[e]) {
                const int32_t k2 = cur_out[src[e]]++;
                col_out[k2] = static_cast<int32_t>(dst[e]);
                slot_out[k2] = slot_of_edge[id];
                ++id;
            }
    }

    int rc = TECGAT_OK;
    if ((rc = upload(&p->rowptr_in, rp_in, st)) || (rc = upload(&p->col_in, col_in, st)) ||
        (rc = upload(&p->rowptr_out, rp_out, st)) || (rc = upload(&p->col_out, col_out, st)) ||
        (rc = upload(&p->slot_out, slot_out, st)) ||
        (rc = build_tiling(p->fwd, false, tile_nodes_fwd, N, rp_in, col_in, rp_out, col_out, slot_out, st)) ||
        (rc = build_tiling(p->bwd, true, tile_nodes_bwd, N, rp_in, col_in, rp_out, col_out, slot_out, st)) ||
        (rc = build_sw(p, tile_nodes_bwd, N, rp_in, col_in, rp_out, col_out, slot_out, st))) {
        tecgat_plan_destroy(p);
        return rc;
    }
    p->h_rowptr_in = static_cast<int32_t *>(malloc(sizeof(int32_t) * (N + 1)));
    p->h_col_in = static_cast<int32_t *>(malloc(sizeof(int32_t) * std::max<int64_t>(E, 1)));
    p->h_eid_in = static_cast<int32_t *>(malloc(sizeof(int32_t) * std::max<int64_t>(E, 1)));
    if (!p->h_rowptr_in || !p->h_col_in || !p->h_eid_in) {
        tecgat_plan_destroy(p);
        tecgat_set_error("plan_create: out of host memory");
        return TECGAT_ENOMEM;
    }
    memcpy(p->h_rowptr_in, rp_in.data(), sizeof(int32_t) * (N + 1));
    memcpy(p->h_col_in, col_in.data(), sizeof(int32_t) * E);
    memcpy(p->h_eid_in, eid_in.data(), sizeof(int32_t) * E);
    // the uploads read from the std::vectors above: finish them before the vectors go away
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        tecgat_plan_destroy(p);
        tecgat_set_error("plan_create: upload failed: %s", cudaGetErrorString(e));
        return TECGAT_ECUDA;
    }
    *plan_out = p;
    return TECGAT_OK;
}

extern "C" int tecgat_plan_info(const tecgat_plan_t *p, int64_t *info) {
    TG_REQUIRE(p && info, TECGAT_EINVAL, "plan_info: NULL argument");
    info[0] = p->num_edges;
    info[1] = p->max_in_deg;
    info[2] = p->max_out_deg;
    info[3] = p->fwd.num_tiles;
    info[4] = p->fwd.T;
    info[5] = p->fwd.max_window;
    info[6] = p->num_nodes;
    info[7] = p->kept_edges;
    info[8] = p->bwd.num_tiles;
    info[9] = p->bwd.T;
    info[10] = p->bwd.max_window;
    info[11] = p->sw ? 1 : 0;  // the sliding-window backward tiling exists (banded graph)
    return TECGAT_OK;
}

extern "C" int tecgat_plan_export(const tecgat_plan_t *p, int32_t *rowptr, int32_t *col, int32_t *eid) {
    TG_REQUIRE(p, TECGAT_EINVAL, "plan_export: NULL plan");
    if (rowptr) memcpy(rowptr, p->h_rowptr_in, sizeof(int32_t) * (p->num_nodes + 1));
    if (col) memcpy(col, p->h_col_in, sizeof(int32_t) * p->num_edges);
    if (eid) memcpy(eid, p->h_eid_in, sizeof(int32_t) * p->num_edges);
    return TECGAT_OK;
}

extern "C" int tecgat_dropout_mask_host(uint64_t seed, int64_t first_slot, int64_t count, int32_t heads,
                                        float dropout_p, int64_t edges_per_snapshot, uint8_t *keep) {
    // slot = snapshot * edges_per_snapshot + CSR slot (the kernels' numbering)
    TG_REQUIRE(keep && heads > 0 && count >= 0 && edges_per_snapshot > 0, TECGAT_EINVAL, "dropout_mask_host: bad argument");
    const uint32_t thr = tg::dropout_threshold(dropout_p);
    const uint32_t base = tg::dropout_base(seed), stride = tg::dropout_stream_stride(edges_per_snapshot);
    for (int64_t i = 0; i < count; ++i) {
        const int64_t g = first_slot + i;
        const uint32_t snap = uint32_t(g / edges_per_snapshot), slot = uint32_t(g % edges_per_snapshot);
        for (int32_t h = 0; h < heads; ++h)
            keep[i * heads + h] = tg::dropout_bits(tg::dropout_key(base, snap, uint32_t(heads), uint32_t(h), stride), slot) >= thr;
    }
    return TECGAT_OK;
}
