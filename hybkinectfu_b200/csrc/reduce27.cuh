// reduce27.cuh — the reference's two-stage float reduction trees for the 27 normal-equation sums,
// restated on registers, shared memory and warp shuffles (src/cuda/device_functions.h:11-55).  Shared by
// the ICP and SDF tracking kernels.  The association of every addition equals the reference's, so the
// sums are bit-identical; see icp.cu for the mapping.
#pragma once
#include "hkf_internal.cuh"

// the reference's first-stage tree over the 256 values of a tile, given this lane's 8 values
// p[j] = value of thread (lane + 32 j): levels +128, +64, +32 are additions among a lane's own values,
// levels 16..1 a shuffle tree.
__device__ __forceinline__ float tile_tree(const float* p)
{
    float c = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[4]), __fadd_rn(p[2], p[6])),
                        __fadd_rn(__fadd_rn(p[1], p[5]), __fadd_rn(p[3], p[7])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = __fadd_rn(c, __shfl_down_sync(0xffffffffu, c, o));
    return c;
}

// upper-triangular (i, j) of packed sum k: rows of [JtJ | Jtr], 7 + 6 + 5 + 4 + 3 + 2
static __constant__ unsigned char c_sum_i[HKF_SUMS] = { 0,0,0,0,0,0,0, 1,1,1,1,1,1, 2,2,2,2,2, 3,3,3,3, 4,4,4, 5,5 };
static __constant__ unsigned char c_sum_j[HKF_SUMS] = { 0,1,2,3,4,5,6, 1,2,3,4,5,6, 2,3,4,5,6, 3,4,5,6, 4,5,6, 5,6 };

// First stage for tiles whose 256 solver rows r[7] sit in shared memory as s_r[k][256] (k = 0..6, thread-major):
// the warp computes the 256 products r_i * r_j of sum k and folds them with the reference's block tree.
__device__ __forceinline__ float tile_sum_from_rows(const float (*s_r)[256], int k, int lane)
{
    const int i = c_sum_i[k], j = c_sum_j[k];
    float p[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) p[m] = __fmul_rn(s_r[i][lane + 32 * m], s_r[j][lane + 32 * m]);
    return tile_tree(p);
}

// reduceGbufKernel<<<27,512>>> restated for one row of partials, executed by one warp: thread t of the
// reference sums t, t+512, ... in order; then 256/128/64/32 pairwise, then 16..1.  A lane owns the 16
// reference threads t = lane + 32 m.  Loads are predicated, never looped on a data-dependent bound, so that
// all 48 loads of a 1536-partial chunk are in flight together (the looped form serialised them: 43 us at
// 1200 partials, see profiles/r1_icp_*.txt).
template <int G>
__device__ __forceinline__ void second_stage_rows(const float* __restrict__ partials, int n, int lane,
                                                  const int* rows, float* out)
{
    float s[G][16];
#pragma unroll
    for (int q = 0; q < G; ++q)
#pragma unroll
        for (int m = 0; m < 16; ++m) s[q][m] = 0.f;
    for (int base = 0; base < n; base += 1536) {
        float v[G][16][3];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const float* row = partials + (size_t)(rows[q] < 0 ? 0 : rows[q]) * n;
#pragma unroll
            for (int m = 0; m < 16; ++m)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int i = base + lane + 32 * m + 512 * t;
                    // partials were written by other blocks of this launch: read through L2, never the non-coherent path
                    v[q][m][t] = (rows[q] >= 0 && i < n) ? __ldcg(row + i) : 0.f;
                }
        }
#pragma unroll
        for (int q = 0; q < G; ++q)
#pragma unroll
            for (int m = 0; m < 16; ++m)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int i = base + lane + 32 * m + 512 * t;
                    if (i < n) s[q][m] = __fadd_rn(s[q][m], v[q][m][t]);
                }
    }
#pragma unroll
    for (int q = 0; q < G; ++q) {
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

// fold the [27][n] partials into sums[27] with the NWARPS warps of the calling block (call from all threads);
// one row at a time per warp: 48 loads in flight per lane, ~60 registers
template <int NWARPS>
__device__ __forceinline__ void second_stage_all(const float* __restrict__ partials, int n, int warp, int lane, float* s_sums)
{
#pragma unroll 1
    for (int k = warp; k < HKF_SUMS; k += NWARPS) {
        int rows[1] = { k };
        float out[1];
        second_stage_rows<1>(partials, n, lane, rows, out);
        if (lane == 0) s_sums[k] = out[0];
    }
}
