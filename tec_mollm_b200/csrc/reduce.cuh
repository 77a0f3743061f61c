// reduce.cuh -- host launcher of the fixed-order column reduction (reduce.cu).
#pragma once
#include "common.cuh"

namespace tg {
struct ReduceSegs {  // up to four output segments: columns [begin, end) of the partial matrix go to out[0..end-begin)
    float *out[4];
    int begin[4];
    int end[4];
};
int reduce_columns(const float *partials, int64_t num, int width, const ReduceSegs &segs, cudaStream_t st);

// Several partial matrices finished by ONE launch (one CTA per column of every job); accumulate = 1: out += sum (the
// caller's .grad storage already holds earlier contributions -- replaces autograd's separate accumulation kernels).
struct ReduceJob {
    const float *partials;
    int64_t num;
    int width;
    ReduceSegs segs;
};
int reduce_columns_multi(const ReduceJob *jobs, int njobs, int accumulate, cudaStream_t st);  // njobs <= 2
}  // namespace tg
