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

// reduceGbufKernel<<<27,512>>> restated for NR rows of partials at once, executed by one warp:
// thread t of the reference sums t, t+512, ...; then 256/128/64/32 pairwise, then 16..1.
// All loads of the NR rows are issued before any tree so that one L2 round trip covers them.
template <int NR>
__device__ __forceinline__ void second_stage_rows(const float* __restrict__ partials, int n, int lane,
                                                  const int* rows, float* out)
{
    float s[NR][16];
#pragma unroll
    for (int q = 0; q < NR; ++q) {
        const float* row = partials + (size_t)rows[q] * n;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            float acc = 0.f;
            // partials were written by other blocks of this launch: read through L2, never the non-coherent path
            if (rows[q] >= 0)
                for (int i = lane + 32 * m; i < n; i += 512) acc = __fadd_rn(acc, __ldcg(row + i));
            s[q][m] = acc;
        }
    }
#pragma unroll
    for (int q = 0; q < NR; ++q) {
#pragma unroll
        for (int m = 0; m < 8; ++m) s[q][m] = __fadd_rn(s[q][m], s[q][m + 8]);
#pragma unroll
        for (int m = 0; m < 4; ++m) s[q][m] = __fadd_rn(s[q][m], s[q][m + 4]);
        s[q][0] = __fadd_rn(s[q][0], s[q][2]); s[q][1] = __fadd_rn(s[q][1], s[q][3]);
        float c = __fadd_rn(s[q][0], s[q][1]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = __fadd_rn(c, __shfl_down_sync(0xffffffffu, c, o));
        out[q] = c;
    }
}

// fold the [27][n] partials into sums[27] with `nwarps` warps of the calling block (call from all threads)
template <int NWARPS>
__device__ __forceinline__ void second_stage_all(const float* __restrict__ partials, int n, int warp, int lane, float* s_sums)
{
    constexpr int NR = (HKF_SUMS + NWARPS - 1) / NWARPS;
    int rows[NR];
    float out[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) { const int k = warp + q * NWARPS; rows[q] = (k < HKF_SUMS) ? k : -1; }
    second_stage_rows<NR>(partials, n, lane, rows, out);
#pragma unroll
    for (int q = 0; q < NR; ++q) if (rows[q] >= 0 && lane == 0) s_sums[rows[q]] = out[q];
}
