// reduce27.cuh — the reference's two-stage float reduction trees for the 27 normal-equation sums,
// restated on registers and warp shuffles (src/cuda/device_functions.h:11-55).  Shared by the ICP and
// SDF tracking kernels.  The association of every addition equals the reference's, so the sums are
// bit-identical; see icp.cu for the mapping.
#pragma once
#include "hkf_internal.cuh"

// the reference's first-stage tree over the 256 values of a tile, given this lane's 8 values
__device__ __forceinline__ float tile_tree(const float* p)
{
    float c = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[4]), __fadd_rn(p[2], p[6])),
                        __fadd_rn(__fadd_rn(p[1], p[5]), __fadd_rn(p[3], p[7])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = __fadd_rn(c, __shfl_down_sync(0xffffffffu, c, o));
    return c;
}

// reduceGbufKernel<<<27,512>>> restated for one row of partials, executed by one warp:
// thread t of the reference sums t, t+512, ...; then 256/128/64/32 pairwise, then 16..1.
__device__ __forceinline__ float second_stage_row(const float* __restrict__ row, int n, int lane)
{
    float s[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        float acc = 0.f;
        // partials were written by other blocks of this launch: read through L2, never the non-coherent path
        for (int i = lane + 32 * m; i < n; i += 512) acc = __fadd_rn(acc, __ldcg(row + i));
        s[m] = acc;
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) s[m] = __fadd_rn(s[m], s[m + 8]);
#pragma unroll
    for (int m = 0; m < 4; ++m) s[m] = __fadd_rn(s[m], s[m + 4]);
    s[0] = __fadd_rn(s[0], s[2]); s[1] = __fadd_rn(s[1], s[3]);
    float c = __fadd_rn(s[0], s[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = __fadd_rn(c, __shfl_down_sync(0xffffffffu, c, o));
    return c;
}

