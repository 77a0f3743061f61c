// reduce.cu -- fixed-order second stage of every parameter-gradient reduction (d att, d bias, d W, d b):
// out[j] = sum_p partials[p, j], one CTA per column, fp64 accumulation, no atomics -> bit-reproducible.
#include "common.cuh"
#include "reduce.cuh"

namespace tg {

struct ReduceJobs {
    ReduceJob job[2];
    int njobs, accumulate;
};

__global__ void __launch_bounds__(256) reduce_columns_kernel(const ReduceJobs jobs) {
    __shared__ double sh[256];
    int j = blockIdx.x, which = 0;
    if (jobs.njobs > 1 && j >= jobs.job[0].width) {
        j -= jobs.job[0].width;
        which = 1;
    }
    const ReduceJob &J = jobs.job[which];
    const float *__restrict__ partials = J.partials;
    const int width = J.width;
    double acc = 0.0;
    for (int64_t p = threadIdx.x; p < J.num; p += 256) acc += static_cast<double>(partials[p * width + j]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (J.segs.out[s] && j >= J.segs.begin[s] && j < J.segs.end[s]) {
                float *o = J.segs.out[s] + (j - J.segs.begin[s]);
                *o = jobs.accumulate ? static_cast<float>(static_cast<double>(*o) + sh[0]) : static_cast<float>(sh[0]);
            }
    }
}

int reduce_columns_multi(const ReduceJob *jobs, int njobs, int accumulate, cudaStream_t st) {
    TG_REQUIRE(njobs >= 1 && njobs <= 2, TECGAT_EINVAL, "reduce: %d jobs", njobs);
    ReduceJobs J;
    int width = 0;
    for (int i = 0; i < 2; ++i) {
        J.job[i] = jobs[i < njobs ? i : 0];
        if (i < njobs) width += jobs[i].width;
    }
    J.njobs = njobs;
    J.accumulate = accumulate;
    if (width <= 0) return TECGAT_OK;
    reduce_columns_kernel<<<width, 256, 0, st>>>(J);
    tg_count_launch();
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

int reduce_columns(const float *partials, int64_t num, int width, const ReduceSegs &segs, cudaStream_t st) {
    if (width <= 0) return TECGAT_OK;
    ReduceJob j{partials, num, width, segs};
    return reduce_columns_multi(&j, 1, 0, st);
}

}  // namespace tg
