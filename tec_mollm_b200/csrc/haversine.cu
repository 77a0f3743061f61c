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
#include <new>
#include <vector>

#include "common.cuh"

struct tecgraph_ctx {
    int64_t n = 0;
    double thr = 0, radius = 0;
    double *lat = nullptr, *lon = nullptr;  // device copies (n)
    double *cmin = nullptr, *cmax = nullptr;  // per 32-column chunk latitude range
    int32_t *deg = nullptr;                 // (n) neighbour counts, ambiguous pairs resolved
    int64_t *rowptr = nullptr;              // (n+1)
    int64_t *extra = nullptr;               // sorted keys i*n+j of guard-band pairs that ARE edges (host verdict)
    int64_t num_extra = 0;
    int64_t total = 0;
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

__global__ void chunk_range_kernel(const double *__restrict__ lat, int64_t n, double *__restrict__ cmin,
                                   double *__restrict__ cmax) {
    const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t nchunks = (n + 31) / 32;
    if (c >= nchunks) return;
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t j = c * 32; j < min(n, c * 32 + 32); ++j) {
        lo = fmin(lo, lat[j]);
        hi = fmax(hi, lat[j]);
    }
    cmin[c] = lo;
    cmax[c] = hi;
}

// 0 = not an edge, 1 = edge, 2 = inside the guard band (host decides)
__device__ __forceinline__ int classify(double d, double thr) {
    const double band = thr * kGuard;
    if (d > thr + band) return 0;
    if (d < thr - band) return 1;
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

// FILL == false: count certain edges per row, append guard-band pairs to `amb` (keys i*n+j).
// FILL == true : write (row, col, weight) in column order; guard-band pairs are edges iff their key is in `extra`.
template <bool FILL>
__global__ void __launch_bounds__(256) edges_kernel(const double *__restrict__ lat, const double *__restrict__ lon, int64_t n,
                                                    double thr, double radius, const double *__restrict__ cmin,
                                                    const double *__restrict__ cmax, int32_t *__restrict__ deg,
                                                    int64_t *__restrict__ amb, unsigned int *__restrict__ amb_count,
                                                    const int64_t *__restrict__ extra, int64_t num_extra,
                                                    const int64_t *__restrict__ rowptr, int64_t *__restrict__ edge_index,
                                                    float *__restrict__ edge_weight, int64_t total) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    if (i >= n) return;
    const double lat_i = lat[i], lon_i = lon[i];
    const double gap = (thr * (1.0 + 2.0 * kGuard)) / radius;  // chunks farther than this in latitude hold no candidates
    const int64_t nchunks = (n + 31) / 32;
    int count = 0;
    int64_t pos = FILL ? rowptr[i] : 0;
    double wi = 0.0;
    if (FILL) wi = deg[i] > 0 ? 1.0 / sqrt(static_cast<double>(deg[i])) : 0.0;
    for (int64_t cb = 0; cb < nchunks; cb += 32) {
        const int64_t c = cb + lane;
        bool cand = false;
        if (c < nchunks) cand = !(cmin[c] - lat_i > gap || lat_i - cmax[c] > gap);
        unsigned todo = __ballot_sync(0xffffffffu, cand);
        while (todo) {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t j = (cb + b) * 32 + lane;
            int cls = 0;
            if (j < n && j != i) cls = classify(hav_km(lat_i, lon_i, lat[j], lon[j], radius), thr);
            if (!FILL) {
                count += (cls == 1);
                if (cls == 2) {
                    const unsigned slot = atomicAdd(amb_count, 1u);  // integer append; order is irrelevant (sorted on host)
                    if (slot < (unsigned)kMaxAmbiguous) amb[slot] = i * n + j;
                }
            } else {
                const bool is_edge = cls == 1 || (cls == 2 && extra_contains(extra, num_extra, i * n + j));
                const unsigned mask = __ballot_sync(0xffffffffu, is_edge);
                if (is_edge) {
                    const int64_t p = pos + __popc(mask & ((1u << lane) - 1u));
                    edge_index[p] = i;
                    edge_index[total + p] = j;
                    const double wj = deg[j] > 0 ? 1.0 / sqrt(static_cast<double>(deg[j])) : 0.0;
                    edge_weight[p] = static_cast<float>((wi * 1.0) * wj);
                }
                pos += __popc(mask);
            }
        }
    }
    if (!FILL) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) count += __shfl_xor_sync(0xffffffffu, count, off);
        if (lane == 0) deg[i] = count;
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
    cudaFree(c->lat);
    cudaFree(c->lon);
    cudaFree(c->cmin);
    cudaFree(c->cmax);
    cudaFree(c->deg);
    cudaFree(c->rowptr);
    cudaFree(c->extra);
    delete c;
    return TECGAT_OK;
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
    const int64_t nchunks = (n + 31) / 32;
    int64_t *amb = nullptr;
    unsigned int *amb_count = nullptr;
    int rc = TECGAT_OK;
    auto fail = [&](int code) {
        cudaFree(amb);
        cudaFree(amb_count);
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
    TG_TRY(cudaMalloc(&c->lat, sizeof(double) * n));
    TG_TRY(cudaMalloc(&c->lon, sizeof(double) * n));
    TG_TRY(cudaMalloc(&c->cmin, sizeof(double) * nchunks));
    TG_TRY(cudaMalloc(&c->cmax, sizeof(double) * nchunks));
    TG_TRY(cudaMalloc(&c->deg, sizeof(int32_t) * n));
    TG_TRY(cudaMalloc(&c->rowptr, sizeof(int64_t) * (n + 1)));
    TG_TRY(cudaMalloc(&amb, sizeof(int64_t) * kMaxAmbiguous));
    TG_TRY(cudaMalloc(&amb_count, sizeof(unsigned int)));
    TG_TRY(cudaMemcpyAsync(c->lat, lat, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    TG_TRY(cudaMemcpyAsync(c->lon, lon, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    TG_TRY(cudaMemsetAsync(amb_count, 0, sizeof(unsigned int), st));
    chunk_range_kernel<<<(unsigned)((nchunks + 255) / 256), 256, 0, st>>>(c->lat, n, c->cmin, c->cmax); tg_count_launch();
    TG_TRY(cudaGetLastError());
    const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    edges_kernel<false><<<blocks, 256, 0, st>>>(c->lat, c->lon, n, thr_km, radius_km, c->cmin, c->cmax, c->deg, amb, amb_count,
                                                 nullptr, 0, nullptr, nullptr, nullptr, 0); tg_count_launch();
    TG_TRY(cudaGetLastError());
    std::vector<int32_t> deg(n);
    unsigned int namb = 0;
    TG_TRY(cudaMemcpyAsync(deg.data(), c->deg, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    TG_TRY(cudaMemcpyAsync(&namb, amb_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    TG_TRY(cudaStreamSynchronize(st));
    if (namb > (unsigned)kMaxAmbiguous) {
        tecgat_set_error("edges_count: %u pairs fall in the guard band (capacity %d)", namb, kMaxAmbiguous);
        return fail(TECGAT_ENOSUP);
    }
    if (ambiguous_host) *ambiguous_host = namb;
    // ---- guard band: the host libm decides, in the reference's operation order ----------------------------------
    std::vector<int64_t> extra;
    if (namb > 0) {
        std::vector<int64_t> keys(namb);
        std::vector<double> hlat(n), hlon(n);
        TG_TRY(cudaMemcpyAsync(keys.data(), amb, sizeof(int64_t) * namb, cudaMemcpyDeviceToHost, st));
        TG_TRY(cudaMemcpyAsync(hlat.data(), c->lat, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        TG_TRY(cudaMemcpyAsync(hlon.data(), c->lon, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
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
        }
        c->num_extra = (int64_t)extra.size();
    }
    std::vector<int64_t> rowptr(n + 1);
    rowptr[0] = 0;
    for (int64_t i = 0; i < n; ++i) rowptr[i + 1] = rowptr[i] + deg[i];
    c->total = rowptr[n];
    TG_TRY(cudaMemcpyAsync(c->rowptr, rowptr.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
    TG_TRY(cudaStreamSynchronize(st));
#undef TG_TRY
    cudaFree(amb);
    cudaFree(amb_count);
    (void)rc;
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
    edges_kernel<true><<<blocks, 256, 0, st>>>(c->lat, c->lon, c->n, c->thr, c->radius, c->cmin, c->cmax, c->deg, nullptr, nullptr,
                                                c->extra, c->num_extra, c->rowptr, edge_index, edge_weight, c->total); tg_count_launch();
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}
