// Row softmax by one CTA of 256 threads (shared by softmax_rows_kernel in pointwise.cu and the fused INT8 head in
// fc_stream.cu, so that both produce the same bits): thread t owns elements t, t+256, ...; maximum and sum are folded
// per warp by xor-shuffles and across the 8 warps in warp order.
#pragma once
#include <cuda_runtime.h>

namespace b200 {

constexpr int SOFTMAX_THREADS = 256;

__device__ __forceinline__ float softmax_block_reduce(float v, bool is_max, float* red /* [8] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : v + u;
    }
    __syncthreads();                 // the previous use of red[] is over
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < SOFTMAX_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

// x, y: one row each (global or shared); every thread of the 256-thread CTA must call this
__device__ __forceinline__ void softmax_row_block(const float* x, float* y, int len, float* red) {
    float mx = -3.402823466e+38f;
    for (int i = threadIdx.x; i < len; i += SOFTMAX_THREADS) mx = fmaxf(mx, x[i]);
    mx = softmax_block_reduce(mx, true, red);
    float sum = 0.f;
    for (int i = threadIdx.x; i < len; i += SOFTMAX_THREADS) {
        const float e = expf(x[i] - mx);
        y[i] = e;
        sum += e;
    }
    sum = softmax_block_reduce(sum, false, red);
    for (int i = threadIdx.x; i < len; i += SOFTMAX_THREADS) y[i] = __fdiv_rn(y[i], sum);
}

}  // namespace b200
