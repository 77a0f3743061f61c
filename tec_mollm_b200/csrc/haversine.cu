// haversine.cu -- graph builder: pairwise haversine distances -> thresholded, row-major edge list + sym-normalised
// weights.  Replaces /root/reference/src/graph/graph_constructor.py:34-59 (sklearn haversine_distances, fp64,
// single thread), :61-81 (inclusive `<=` threshold, zero diagonal), :99-128 (D^-1/2 A D^-1/2 through scipy.sparse)
// and the COO extraction at :141-144, without ever forming the (N, N) matrices (the reference cannot build the
// 64,800-node grid: 2 x 33.6 GB).
//
// Arithmetic is the reference's, in its operation order, in fp64:
//     d = 2 asin(sqrt(sin^2((lat1-lat2)/2) + cos(lat1) cos(lat2) sin^2((lon1-lon2)/2))) * radius
// CUDA's fp64 sin/cos/asin are not bit-identical to glibc's, so the distance itself is reproduced to ~1 ulp, not bit
// for bit; the EDGE SET is made bit-exact: any pair whose distance lies within a relative guard band of the threshold
// is re-evaluated on the host with glibc's libm in exactly sklearn's operation order and that verdict is used.
//
// Layout of the work: one warp per row i, lanes over 32-column chunks; a chunk is skipped when the latitude gap between
// row i and every node of the chunk already exceeds the threshold (great-circle distance >= radius * |dlat|), which on
// lat-major grids leaves ~3 latitude rows of candidates per row instead of N.  Two passes (count -> host scan -> fill)
// so the output is ordered exactly like scipy's COO (row ascending, column ascending inside a row).
#include <math.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"

struct tecgraph_ctx {
    int64_t n = 0;
    double thr = 0, radius = 0, r_lo = 0, r_hi = 0;
    unsigned char *slab = nullptr;          // ONE device allocation holding everything below
    double *lat = nullptr, *lon = nullptr;  // device copies (n)
    double *clat = nullptr;                 // cos(lat) per node: evaluated once, not once per pair
    double *cmin = nullptr, *cmax = nullptr, *lmin = nullptr, *lmax = nullptr, *ccos = nullptr;  // per 32-column chunk
    int32_t *deg = nullptr, *evals = nullptr;  // (n) neighbour counts (ambiguous pairs resolved), pairs evaluated
    int64_t *rowptr = nullptr;              // (n+1), device-side scan
    int64_t *amb = nullptr;                 // guard-band pair keys i*n+j
    int64_t *totals = nullptr;              // {edges, evaluated pairs, guard-band pairs}
    int64_t *extra = nullptr;               // sorted keys of guard-band pairs that ARE edges (host verdict)
    int64_t num_extra = 0;
    int64_t total = 0, evaluated = 0, ambiguous = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // count kernel begin/end, fill kernel begin/end
    bool filled = false;
    cudaStream_t stream = nullptr;          // the stream the slab was allocated on (stream-ordered allocation)
};

namespace tg {

constexpr double kGuard = 1e-9;          // relative half-width of the guard band around the threshold
constexpr int kMaxAmbiguous = 1 << 20;   // capacity of the guard-band pair list

__device__ __forceinline__ double hav_km(double lat1, double lon1, double lat2, double lon2, double radius) {
    // explicit _rn intrinsics: no FMA contraction, the reference's association order
    const double s0 = sin(__dmul_rn(0.5, __dsub_rn(lat1, lat2)));
    const double s1 = sin(__dmul_rn(0.5, __dsub_rn(lon1, lon2)));
    const double cc = __dmul_rn(cos(lat1), cos(lat2));
    const double r = __dadd_rn(__dmul_rn(s0, s0), __dmul_rn(__dmul_rn(cc, s1), s1));
    return __dmul_rn(__dmul_rn(2.0, asin(sqrt(r))), radius);
}

__global__ void distance_rows_kernel(const double *__restrict__ lat, const double *__restrict__ lon, int64_t n, int64_t r0,
                                     int64_t r1, double radius, double *__restrict__ out) {
    const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t i = r0 + blockIdx.y;
    if (j < n && i < r1) out[(i - r0) * n + j] = hav_km(lat[i], lon[i], lat[j], lon[j], radius);
}

__global__ void node_cos_kernel(const double *__restrict__ lat, int64_t n, double *__restrict__ clat) {
    const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (j < n) clat[j] = cos(lat[j]);
}

__global__ void chunk_range_kernel(const double *__restrict__ lat, const double *__restrict__ lon, const double *__restrict__ clat,
                                   int64_t n, double *__restrict__ cmin, double *__restrict__ cmax, double *__restrict__ lmin,
                                   double *__restrict__ lmax, double *__restrict__ ccos) {
    const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t nchunks = (n + 31) / 32;
    if (c >= nchunks) return;
    double lo = INFINITY, hi = -INFINITY, llo = INFINITY, lhi = -INFINITY, cc = INFINITY;
    for (int64_t j = c * 32; j < min(n, c * 32 + 32); ++j) {
        lo = fmin(lo, lat[j]);
        hi = fmax(hi, lat[j]);
        llo = fmin(llo, lon[j]);
        lhi = fmax(lhi, lon[j]);
        cc = fmin(cc, clat[j]);
    }
    cmin[c] = lo; cmax[c] = hi; lmin[c] = llo; lmax[c] = lhi;
    ccos[c] = fmax(cc, 0.0);
}

// The reference's haversine argument in its operation order (sklearn: sin_0 * sin_0 + cos(lat1) * cos(lat2) * sin_1 * sin_1):
//     r = sin^2((lat1 - lat2) / 2) + ((cos(lat1) cos(lat2)) sin((lon1 - lon2) / 2)) sin((lon1 - lon2) / 2)
// d = 2 R asin(sqrt(r)) is monotone in r, so "d <= thr" is decided on r against r_thr = sin^2(thr / 2R) with a guard band:
// no asin / sqrt per pair.  0 = not an edge, 1 = edge, 2 = inside the guard band (the host's libm decides on d itself).
__device__ __forceinline__ int classify_pair(double lat_i, double lon_i, double clat_i, double lat_j, double lon_j, double clat_j,
                                             double r_lo, double r_hi) {
    const double s0 = sin(__dmul_rn(0.5, __dsub_rn(lat_i, lat_j)));
    const double s1 = sin(__dmul_rn(0.5, __dsub_rn(lon_i, lon_j)));
    const double cc = __dmul_rn(clat_i, clat_j);
    const double r = __dadd_rn(__dmul_rn(s0, s0), __dmul_rn(__dmul_rn(cc, s1), s1));
    if (r > r_hi) return 0;
    if (r < r_lo) return 1;
    return 2;
}

__device__ __forceinline__ bool extra_contains(const int64_t *__restrict__ extra, int64_t num, int64_t key) {
    int64_t lo = 0, hi = num;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t v = extra[mid];
        if (v == key) return true;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return false;
}

struct EdgeArgs {
    const double *lat, *lon, *clat, *cmin, *cmax, *lmin, *lmax, *ccos;
    int64_t n;
    double gap, r_lo, r_hi;
    int32_t *deg, *evals;
    int64_t *amb;
    unsigned int *amb_count;
    const int64_t *extra;
    int64_t num_extra;
    const int64_t *rowptr;
    int64_t *edge_index;
    float *edge_weight;
    int64_t total;
};

// One warp per row i, lanes over 32-column chunks.  A chunk is skipped when no node of it can be within the threshold:
// latitude gap alone (d >= R |dlat|), or the longitude gap at the most favourable latitudes of the pair
// (r >= cos(lat_i) min_j cos(lat_j) sin^2(dlon_min / 2), dlon_min taken around the +-pi seam).  On lat-major grids that
// leaves a few dozen candidates per row instead of N.
// FILL == false: count certain edges per row, append guard-band pairs to `amb` (keys i*n+j).
// FILL == true : write (row, col, weight) in column order; guard-band pairs are edges iff their key is in `extra`.
template <bool FILL>
__global__ void __launch_bounds__(256) edges_kernel(const EdgeArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int64_t n = a.n;
    if (i >= n) return;
    const double lat_i = a.lat[i], lon_i = a.lon[i], clat_i = a.clat[i];
    const int64_t nchunks = (n + 31) / 32;
    int count = 0, evals = 0;
    int64_t pos = FILL ? a.rowptr[i] : 0;
    double wi = 0.0;
    if (FILL) wi = a.deg[i] > 0 ? 1.0 / sqrt(static_cast<double>(a.deg[i])) : 0.0;
    for (int64_t cb = 0; cb < nchunks; cb += 32) {
        const int64_t c = cb + lane;
        bool cand = false;
        if (c < nchunks && !(a.cmin[c] - lat_i > a.gap || lat_i - a.cmax[c] > a.gap)) {
            // distance from lon_i to the chunk's longitude interval, also across the seam
            const double lo = a.lmin[c], hi = a.lmax[c];
            const double two_pi = 6.283185307179586476925286766559;
            auto gap_to = [&](double x) { return fmax(0.0, fmax(lo - x, x - hi)); };
            const double dl = fmin(gap_to(lon_i), fmin(gap_to(lon_i + two_pi), gap_to(lon_i - two_pi)));
            const double sh = sin(0.5 * fmin(dl, 3.14159265358979323846));
            cand = !(clat_i * a.ccos[c] * sh * sh > a.r_hi * (1.0 + 1e-6));
        }
        unsigned todo = __ballot_sync(0xffffffffu, cand);
        while (todo) {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t j = (cb + b) * 32 + lane;
            int cls = 0;
            if (j < n && j != i) {
                cls = classify_pair(lat_i, lon_i, clat_i, a.lat[j], a.lon[j], a.clat[j], a.r_lo, a.r_hi);
                ++evals;
            }
            if (!FILL) {
                count += (cls == 1);
                if (cls == 2) {
                    const unsigned slot = atomicAdd(a.amb_count, 1u);  // integer append; order is irrelevant (sorted on host)
                    if (slot < (unsigned)kMaxAmbiguous) a.amb[slot] = i * n + j;
                }
            } else {
                const bool is_edge = cls == 1 || (cls == 2 && extra_contains(a.extra, a.num_extra, i * n + j));
                const unsigned mask = __ballot_sync(0xffffffffu, is_edge);
                if (is_edge) {
                    const int64_t p = pos + __popc(mask & ((1u << lane) - 1u));
                    a.edge_index[p] = i;
                    a.edge_index[a.total + p] = j;
                    const double wj = a.deg[j] > 0 ? 1.0 / sqrt(static_cast<double>(a.deg[j])) : 0.0;
                    a.edge_weight[p] = static_cast<float>((wi * 1.0) * wj);
                }
                pos += __popc(mask);
            }
        }
    }
    if (!FILL) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            count += __shfl_xor_sync(0xffffffffu, count, off);
            evals += __shfl_xor_sync(0xffffffffu, evals, off);
        }
        if (lane == 0) {
            a.deg[i] = count;
            a.evals[i] = evals;
        }
    }
}

// device-side exclusive scan of the degrees (one CTA: n is at most a few 10^5) + totals {edges, evaluated pairs}
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t *__restrict__ deg, const int32_t *__restrict__ evals, int64_t n,
                                                    int64_t *__restrict__ rowptr, int64_t *__restrict__ totals) {
    __shared__ int64_t part[1024], epart[1024];
    const int t = threadIdx.x;
    const int64_t per = (n + 1023) / 1024, b0 = min(n, t * per), b1 = min(n, b0 + per);
    int64_t s = 0, e = 0;
    for (int64_t i = b0; i < b1; ++i) {
        s += deg[i];
        e += evals ? evals[i] : 0;
    }
    part[t] = s;
    epart[t] = e;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0, erun = 0;
        for (int k = 0; k < 1024; ++k) {
            const int64_t v = part[k];
            part[k] = run;
            run += v;
            erun += epart[k];
        }
        rowptr[n] = run;
        totals[0] = run;
        if (evals) totals[1] = erun;
    }
    __syncthreads();
    int64_t run = part[t];
    for (int64_t i = b0; i < b1; ++i) {
        rowptr[i] = run;
        run += deg[i];
    }
}

// the reference's arithmetic with the host libm (glibc), sklearn's operation order (see oracle/haversine_ref.c)
static double hav_km_host(double lat1, double lon1, double lat2, double lon2, double radius) {
    volatile double s0 = sin(0.5 * (lat1 - lat2));
    volatile double s1 = sin(0.5 * (lon1 - lon2));
    volatile double t0 = s0 * s0;
    volatile double t1 = cos(lat1) * cos(lat2);
    volatile double t2 = t1 * s1;
    volatile double t3 = t2 * s1;
    volatile double r = t0 + t3;
    volatile double a = 2.0 * asin(sqrt(r));
    return a * radius;
}

}  // namespace tg

extern "C" int tecgraph_distance_rows(const double *lat, const double *lon, int64_t n, int64_t r0, int64_t r1,
                                      double radius_km, double *out, void *stream) {
    TG_REQUIRE(lat && lon && out, TECGAT_EINVAL, "distance_rows: NULL argument");
    TG_REQUIRE(n > 0 && r0 >= 0 && r1 >= r0 && r1 <= n, TECGAT_EINVAL, "distance_rows: bad row range [%lld, %lld) of %lld",
               (long long)r0, (long long)r1, (long long)n);
    if (r1 == r0) return TECGAT_OK;
    TG_REQUIRE(r1 - r0 <= 65535, TECGAT_EINVAL, "distance_rows: at most 65535 rows per call");
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(r1 - r0));
    tg::distance_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(lat, lon, n, r0, r1, radius_km, out); tg_count_launch();
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

extern "C" int tecgraph_ctx_destroy(tecgraph_ctx_t *c) {
    if (!c) return TECGAT_OK;
    if (c->slab) cudaFreeAsync(c->slab, c->stream);  // back to the device's memory pool: the next build reuses it
    cudaFree(c->extra);
    for (cudaEvent_t e : c->ev)
        if (e) cudaEventDestroy(e);
    delete c;
    return TECGAT_OK;
}

static tg::EdgeArgs edge_args(const tecgraph_ctx_t *c) {
    tg::EdgeArgs a;
    a.lat = c->lat; a.lon = c->lon; a.clat = c->clat; a.cmin = c->cmin; a.cmax = c->cmax; a.lmin = c->lmin; a.lmax = c->lmax; a.ccos = c->ccos;
    a.n = c->n;
    a.gap = (c->thr * (1.0 + 2.0 * tg::kGuard)) / c->radius;  // chunks farther than this in latitude hold no candidates
    a.r_lo = c->r_lo; a.r_hi = c->r_hi;
    a.deg = c->deg; a.evals = c->evals; a.amb = c->amb; a.amb_count = reinterpret_cast<unsigned int *>(c->totals + 2);
    a.extra = c->extra; a.num_extra = c->num_extra; a.rowptr = c->rowptr;
    a.edge_index = nullptr; a.edge_weight = nullptr; a.total = c->total;
    return a;
}

extern "C" int tecgraph_edges_count(const double *lat, const double *lon, int64_t n, double thr_km, double radius_km,
                                    void *stream, tecgraph_ctx_t **ctx_out, int64_t *total_host, int64_t *ambiguous_host) {
    using namespace tg;
    TG_REQUIRE(lat && lon && ctx_out && total_host, TECGAT_EINVAL, "edges_count: NULL argument");
    TG_REQUIRE(n > 0 && n < (int64_t(1) << 31), TECGAT_EINVAL, "edges_count: node count %lld out of range", (long long)n);
    TG_REQUIRE(thr_km >= 0 && radius_km > 0, TECGAT_EINVAL, "edges_count: bad threshold / radius");
    *ctx_out = nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    tecgraph_ctx_t *c = new (std::nothrow) tecgraph_ctx_t();
    TG_REQUIRE(c, TECGAT_ENOMEM, "edges_count: out of host memory");
    c->n = n; c->thr = thr_km; c->radius = radius_km;
    {   // d <= thr  <=>  r <= sin^2(thr / 2R); a relative band of 4e-9 in r covers the 1e-9 band in d (d ~ 2R sqrt(r))
        const double half = thr_km / (2.0 * radius_km);
        const double rt = half >= 1.5707963267948966 ? 2.0 : sin(half) * sin(half);
        c->r_lo = rt * (1.0 - 4.0 * kGuard);
        c->r_hi = rt * (1.0 + 4.0 * kGuard);
    }
    const int64_t nchunks = (n + 31) / 32;
    auto fail = [&](int code) {
        tecgraph_ctx_destroy(c);
        return code;
    };
#define TG_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            tecgat_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
            return fail(TECGAT_ECUDA);                                                                \
        }                                                                                             \
    } while (0)
    {   // one allocation for every device array of the build (the per-call cudaMalloc x8 used to dominate small graphs)
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
        const size_t o_lat = take(8 * n), o_lon = take(8 * n), o_clat = take(8 * n), o_cmin = take(8 * nchunks), o_cmax = take(8 * nchunks),
                     o_lmin = take(8 * nchunks), o_lmax = take(8 * nchunks), o_ccos = take(8 * nchunks), o_deg = take(4 * n),
                     o_ev = take(4 * n), o_rp = take(8 * (n + 1)), o_tot = take(64), o_amb = take(8 * size_t(kMaxAmbiguous));
        {   // stream-ordered allocation from the device's default pool, kept warm: a build is a few hundred microseconds of
            // kernels, and cudaMalloc / cudaFree (device-wide synchronisation, page mapping) used to cost 10x that
            static std::once_flag pool_once[64];
            int dev = 0;
            TG_TRY(cudaGetDevice(&dev));
            if (dev >= 0 && dev < 64)
                std::call_once(pool_once[dev], [dev] {
                    cudaMemPool_t pool;
                    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                        uint64_t keep = UINT64_MAX;
                        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                    }
                });
        }
        c->stream = st;
        TG_TRY(cudaMallocAsync(reinterpret_cast<void **>(&c->slab), off, st));
        unsigned char *b = c->slab;
        c->lat = reinterpret_cast<double *>(b + o_lat); c->lon = reinterpret_cast<double *>(b + o_lon); c->clat = reinterpret_cast<double *>(b + o_clat);
        c->cmin = reinterpret_cast<double *>(b + o_cmin); c->cmax = reinterpret_cast<double *>(b + o_cmax);
        c->lmin = reinterpret_cast<double *>(b + o_lmin); c->lmax = reinterpret_cast<double *>(b + o_lmax); c->ccos = reinterpret_cast<double *>(b + o_ccos);
        c->deg = reinterpret_cast<int32_t *>(b + o_deg); c->evals = reinterpret_cast<int32_t *>(b + o_ev);
        c->rowptr = reinterpret_cast<int64_t *>(b + o_rp); c->totals = reinterpret_cast<int64_t *>(b + o_tot); c->amb = reinterpret_cast<int64_t *>(b + o_amb);
    }
    for (cudaEvent_t &e : c->ev) TG_TRY(cudaEventCreate(&e));
    TG_TRY(cudaMemcpyAsync(c->lat, lat, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    TG_TRY(cudaMemcpyAsync(c->lon, lon, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    TG_TRY(cudaMemsetAsync(c->totals, 0, 64, st));
    node_cos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->lat, n, c->clat); tg_count_launch();
    chunk_range_kernel<<<(unsigned)((nchunks + 255) / 256), 256, 0, st>>>(c->lat, c->lon, c->clat, n, c->cmin, c->cmax, c->lmin, c->lmax, c->ccos);
    tg_count_launch();
    TG_TRY(cudaGetLastError());
    const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    TG_TRY(cudaEventRecord(c->ev[0], st));
    edges_kernel<false><<<blocks, 256, 0, st>>>(edge_args(c)); tg_count_launch();
    TG_TRY(cudaEventRecord(c->ev[1], st));
    TG_TRY(cudaGetLastError());
    scan_kernel<<<1, 1024, 0, st>>>(c->deg, c->evals, n, c->rowptr, c->totals); tg_count_launch();
    int64_t totals[3] = {0, 0, 0};
    TG_TRY(cudaMemcpyAsync(totals, c->totals, sizeof(totals), cudaMemcpyDeviceToHost, st));
    TG_TRY(cudaStreamSynchronize(st));
    const unsigned int namb = static_cast<unsigned int>(totals[2] & 0xFFFFFFFFll);
    c->total = totals[0];
    c->evaluated = totals[1];
    c->ambiguous = namb;
    if (namb > (unsigned)kMaxAmbiguous) {
        tecgat_set_error("edges_count: %u pairs fall in the guard band (capacity %d)", namb, kMaxAmbiguous);
        return fail(TECGAT_ENOSUP);
    }
    if (ambiguous_host) *ambiguous_host = namb;
    // ---- guard band: the host libm decides, in the reference's operation order ----------------------------------
    if (namb > 0) {
        std::vector<int64_t> keys(namb), extra;
        std::vector<double> hlat(n), hlon(n);
        std::vector<int32_t> deg(n);
        TG_TRY(cudaMemcpyAsync(keys.data(), c->amb, sizeof(int64_t) * namb, cudaMemcpyDeviceToHost, st));
        TG_TRY(cudaMemcpyAsync(hlat.data(), c->lat, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        TG_TRY(cudaMemcpyAsync(hlon.data(), c->lon, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        TG_TRY(cudaMemcpyAsync(deg.data(), c->deg, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        TG_TRY(cudaStreamSynchronize(st));
        std::sort(keys.begin(), keys.end());
        for (int64_t key : keys) {
            const int64_t i = key / n, j = key % n;
            if (hav_km_host(hlat[i], hlon[i], hlat[j], hlon[j], radius_km) <= thr_km) {
                extra.push_back(key);
                deg[i] += 1;
            }
        }
        if (!extra.empty()) {
            TG_TRY(cudaMalloc(&c->extra, sizeof(int64_t) * extra.size()));
            TG_TRY(cudaMemcpyAsync(c->extra, extra.data(), sizeof(int64_t) * extra.size(), cudaMemcpyHostToDevice, st));
            TG_TRY(cudaMemcpyAsync(c->deg, deg.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
            scan_kernel<<<1, 1024, 0, st>>>(c->deg, nullptr, n, c->rowptr, c->totals); tg_count_launch();
            TG_TRY(cudaMemcpyAsync(totals, c->totals, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            TG_TRY(cudaStreamSynchronize(st));
            c->total = totals[0];
        }
        c->num_extra = (int64_t)extra.size();
    }
#undef TG_TRY
    *total_host = c->total;
    *ctx_out = c;
    return TECGAT_OK;
}

extern "C" int tecgraph_edges_fill(tecgraph_ctx_t *c, int64_t *edge_index, float *edge_weight, void *stream) {
    using namespace tg;
    TG_REQUIRE(c, TECGAT_EINVAL, "edges_fill: NULL context");
    if (c->total == 0) return TECGAT_OK;
    TG_REQUIRE(edge_index && edge_weight, TECGAT_EINVAL, "edges_fill: NULL output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((c->n * 32 + 255) / 256);
    EdgeArgs a = edge_args(c);
    a.edge_index = edge_index;
    a.edge_weight = edge_weight;
    TG_CUDA(cudaEventRecord(c->ev[2], st));
    edges_kernel<true><<<blocks, 256, 0, st>>>(a); tg_count_launch();
    TG_CUDA(cudaEventRecord(c->ev[3], st));
    c->filled = true;
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

// stats4 = {count-kernel ms, fill-kernel ms (0 before tecgraph_edges_fill), pairs evaluated by the count pass, guard-band pairs};
// synchronises the fill kernel's event
extern "C" int tecgraph_ctx_stats(tecgraph_ctx_t *c, double *stats4_host) {
    TG_REQUIRE(c && stats4_host, TECGAT_EINVAL, "ctx_stats: NULL argument");
    float ms = 0.f;
    TG_CUDA(cudaEventSynchronize(c->ev[1]));
    TG_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    stats4_host[0] = ms;
    stats4_host[1] = 0.0;
    if (c->filled) {
        TG_CUDA(cudaEventSynchronize(c->ev[3]));
        TG_CUDA(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
        stats4_host[1] = ms;
    }
    stats4_host[2] = (double)c->evaluated;
    stats4_host[3] = (double)c->ambiguous;
    return TECGAT_OK;
}
