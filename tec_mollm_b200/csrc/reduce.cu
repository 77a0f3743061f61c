// reduce.cu -- fixed-order second stage of every parameter-gradient reduction (d att, d bias, d W, d b):
// out[j] = sum_p partials[p, j], one CTA per column, fp64 accumulation, no atomics -> bit-reproducible.
#include "common.cuh"
#include "reduce.cuh"

namespace tg {

__global__ void __launch_bounds__(256) reduce_columns_kernel(const float *__restrict__ partials, int64_t num, int width,
                                                             ReduceSegs segs) {
    __shared__ double sh[256];
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int64_t p = threadIdx.x; p < num; p += 256) acc += static_cast<double>(partials[p * width + j]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (segs.out[s] && j >= segs.begin[s] && j < segs.end[s]) segs.out[s][j - segs.begin[s]] = static_cast<float>(sh[0]);
    }
}

int reduce_columns(const float *partials, int64_t num, int width, const ReduceSegs &segs, cudaStream_t st) {
    if (width <= 0) return TECGAT_OK;
    reduce_columns_kernel<<<width, 256, 0, st>>>(partials, num, width, segs);
    TG_LAUNCH_CHECK();
    return TECGAT_OK;
}

}  // namespace tg
