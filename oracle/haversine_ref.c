/* CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py): plain-C restatement of the
 * distance arithmetic that /root/reference/src/graph/graph_constructor.py:56 obtains from
 * scikit-learn:  sklearn/metrics/_dist_metrics.pyx.tp, HaversineDistance.rdist / dist
 *
 *     sin_0 = sin(0.5 * (x1[0] - x2[0]));  sin_1 = sin(0.5 * (x1[1] - x2[1]));
 *     rdist = sin_0*sin_0 + cos(x1[0]) * cos(x2[0]) * sin_1*sin_1;   dist = 2 * asin(sqrt(rdist))
 *
 * followed by `* 6371.0` (graph_constructor.py:53-56) and the inclusive threshold + zero
 * diagonal of construct_binary_adjacency (:75-78).  Pinned in tests/test_oracle_graph.py: the
 * distances are compared bit for bit with sklearn's on the golden grids.
 *
 * Build (oracle/Makefile): gcc -O2 -ffp-contract=off -shared -fPIC -> oracle/_build/libhavref.so
 */
#include <math.h>
#include <stdint.h>

/* One distance, exactly in sklearn's operation order. */
double havref_dist_km(double lat1, double lon1, double lat2, double lon2)
{
    double sin_0 = sin(0.5 * (lat1 - lat2));
    double sin_1 = sin(0.5 * (lon1 - lon2));
    double r = sin_0 * sin_0 + cos(lat1) * cos(lat2) * sin_1 * sin_1;
    return (2.0 * asin(sqrt(r))) * 6371.0;
}

/* Dense row block D[r0:r1, 0:n] (row-major into `out`, (r1-r0)*n doubles). */
void havref_rows(const double *lat, const double *lon, int64_t n, int64_t r0, int64_t r1, double *out)
{
    for (int64_t i = r0; i < r1; ++i)
        for (int64_t j = 0; j < n; ++j)
            out[(i - r0) * n + j] = havref_dist_km(lat[i], lon[i], lat[j], lon[j]);
}

/* Per-row in-threshold neighbour counts (self excluded) for rows [r0, r1). */
void havref_count(const double *lat, const double *lon, int64_t n, int64_t r0, int64_t r1, double thr_km,
                  int64_t *count)
{
    for (int64_t i = r0; i < r1; ++i) {
        int64_t c = 0;
        for (int64_t j = 0; j < n; ++j)
            if (j != i && havref_dist_km(lat[i], lon[i], lat[j], lon[j]) <= thr_km)
                ++c;
        count[i - r0] = c;
    }
}

/* Fill pass: rowptr = exclusive scan of the counts; writes (row, col) pairs row-major. */
void havref_fill(const double *lat, const double *lon, int64_t n, int64_t r0, int64_t r1, double thr_km,
                 const int64_t *rowptr, int64_t *row_out, int64_t *col_out)
{
    for (int64_t i = r0; i < r1; ++i) {
        int64_t p = rowptr[i - r0];
        for (int64_t j = 0; j < n; ++j)
            if (j != i && havref_dist_km(lat[i], lon[i], lat[j], lon[j]) <= thr_km) {
                row_out[p] = i;
                col_out[p] = j;
                ++p;
            }
    }
}
