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
}  // namespace tg
