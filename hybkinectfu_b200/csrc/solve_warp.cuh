// solve_warp.cuh — the 6x6 solves of hkf_solver.h spread over the lanes of one warp.
//
// The tracking kernels end every iteration with the host part of the reference's loop (float LLT solve and LU
// determinant through Eigen, src/CameraPoseFinderICP.cpp:113-145) executed by the last block; on one lane that is a
// chain of ~450 dependent-ish instructions per solve.  Here lanes own ROWS: the factorisation advances column by column
// with the row of the pivot broadcast by shuffles, so the five divisions / eliminations of a column run side by side.
// Every entry is still computed by exactly the operations of hkf_llt_solve6 / hkf_det6 in the same order (the
// subtractions of a dot product stay sequential in j), so the results are bit-identical to the single-lane code and to
// the host classes' hkf_solver.h; tests/test_icp_emulated.py compares them on the CPU build of this file.
#pragma once
#include "hkf_solver.h"

// index of A[r][k] (symmetric, r, k in 0..5) or of b[k] (r == 6) in the 27 packed sums (rows of the upper triangle of
// [JtJ | Jtr]: row i starts at 7 i - i (i - 1) / 2)
__device__ __forceinline__ int hkf_packed_idx(int r, int k)
{
    const int i = r < k ? r : k, j = r < k ? k : r;
    return i * 7 - (i * (i - 1)) / 2 + (j - i);
}

// x = A^-1 b through A = L L^T: lanes 0..5 own rows 0..5 of A and L, lane 6 carries b as a seventh row, which makes the
// forward substitution the same recurrence as the factorisation (y[k] = (b[k] - sum_j L[k][j] y[j]) / L[k][k]).  The
// back substitution is inherently serial; every lane runs it on broadcast operands, so every lane returns all six x.
// Call with the whole warp converged.
__device__ __forceinline__ void hkf_llt_solve6_warp(const float* s_sums, int lane, float* x)
{
    const unsigned FULL = 0xffffffffu;
    const int r = lane < 7 ? lane : 6;                   // idle lanes mirror lane 6
    float a[6], L[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { a[k] = s_sums[hkf_packed_idx(r, k)]; L[k] = 0.0f; }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float v = a[k];
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (j < k) {
                const float lkj = __shfl_sync(FULL, L[j], k);
                v = HKF_SUB(v, HKF_MUL(L[j], lkj));
            }
        const float lkk = HKF_SQRT(__shfl_sync(FULL, v, k));
        L[k] = (r == k) ? lkk : HKF_DIV(v, lkk);         // rows above the diagonal hold junk that nobody reads
    }
    float y[6], Lii[6], Lt[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        y[i] = __shfl_sync(FULL, L[i], 6);
        Lii[i] = __shfl_sync(FULL, L[i], i);
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (j > i) Lt[j][i] = __shfl_sync(FULL, L[i], j);
    }
#pragma unroll
    for (int ii = 0; ii < 6; ++ii) {
        const int i = 5 - ii;
        float v = y[i];
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (j > i) v = HKF_SUB(v, HKF_MUL(Lt[j][i], x[j]));
        x[i] = HKF_DIV(v, Lii[i]);
    }
}

// det(A) by LU with partial pivoting (hkf_det6), lanes 0..5 own rows.  Every lane returns the determinant.
__device__ __forceinline__ float hkf_det6_warp(const float* s_sums, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const int r = lane < 6 ? lane : 5;                   // idle lanes mirror lane 5
    float a[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) a[j] = s_sums[hkf_packed_idx(r, j)];
    float sign = 1.0f;
    bool singular = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        float best = fabsf(__shfl_sync(FULL, a[k], k));
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i > k) {
                const float v = fabsf(__shfl_sync(FULL, a[k], i));
                if (v > best) { best = v; piv = i; }
            }
        if (best == 0.0f) singular = true;
        // rows k and piv change places (the reference's sequence of conditional swaps performs exactly this one)
        const int src = (r == k) ? piv : (r == piv) ? k : r;
#pragma unroll
        for (int j = 0; j < 6; ++j) a[j] = __shfl_sync(FULL, a[j], src);
        if (piv != k) sign = -sign;
        const float pk = __shfl_sync(FULL, a[k], k);
        float rowk[6];
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (j > k) rowk[j] = __shfl_sync(FULL, a[j], k);
        if (r > k) {
            const float f = HKF_DIV(a[k], pk);
            a[k] = f;
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (j > k) a[j] = HKF_SUB(a[j], HKF_MUL(f, rowk[j]));
        }
    }
    float p = __shfl_sync(FULL, a[0], 0);
#pragma unroll
    for (int k = 1; k < 6; ++k) p = HKF_MUL(p, __shfl_sync(FULL, a[k], k));
    return singular ? 0.0f : HKF_MUL(sign, p);
}
