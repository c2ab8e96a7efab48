// Deterministic block-level reductions: warp shuffle -> shared memory -> one partial per block.
// The per-block partials are combined by k_reduce_finish in a fixed order, so results do not depend
// on scheduling (the reference's OpenMP reductions are order-nondeterministic, fg:10117, fg:12342).
#pragma once
#include <cuda_runtime.h>

#define FGB_RED_MAXV 32

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// OP: 0 sum, 1 min, 2 max.  vals[NV] per thread -> partials[bid*FGB_RED_MAXV + v] (bid: linear block index);
// sh: NV*32 doubles of shared memory that no thread of the block still reads
template <int NV, int OP>
__device__ __forceinline__ void block_reduce_store_sh(double* vals, double* __restrict__ partials, size_t bid, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int v = 0; v < NV; v++) {
        double x = vals[v];
        x = (OP == 0) ? warp_sum(x) : (OP == 1 ? warp_min(x) : warp_max(x));
        if (lane == 0) sh[v * 32 + wid] = x;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int v = 0; v < NV; v++) {
            double x = (lane < nw) ? sh[v * 32 + lane] : (OP == 0 ? 0.0 : (OP == 1 ? INFINITY : -INFINITY));
            x = (OP == 0) ? warp_sum(x) : (OP == 1 ? warp_min(x) : warp_max(x));
            if (lane == 0) partials[bid * FGB_RED_MAXV + v] = x;
        }
    }
}
template <int NV, int OP>
__device__ __forceinline__ void block_reduce_store(double* vals, double* __restrict__ partials) {
    __shared__ double sh[NV * 32];
    block_reduce_store_sh<NV, OP>(vals, partials, (size_t)blockIdx.x, sh);
}
