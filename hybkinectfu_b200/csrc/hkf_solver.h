// hkf_solver.h — the 6x6 normal-equation solve, the "camera shaking" gate and the pose updates of
// both trackers, written once for host and device.
//
// In the reference this arithmetic runs on the HOST through Eigen, after a blocking 27-float D2H copy
// per iteration (src/CameraPoseFinderICP.cpp:113-145, src/CameraPoseFinderSDF.cpp:44-106,
// src/utils/eigen_utils.cpp:39-127).  Here the last thread block of each tracking kernel runs it on the
// device so the pose never leaves HBM; the host classes' step-by-step mode calls the very same code on
// the CPU.  All float operations go through HKF_MUL/ADD/... (explicit round-to-nearest intrinsics on the
// device, uncontracted operators on the host), so both sides produce the same bits except for the three
// sinf/cosf pairs and the atan2f of the shake gate, where the CUDA and glibc float libms may differ by an
// ulp (double-precision trig was measured to cost ~4 us of single-thread latency per iteration).
//
// Eigen is an un-vendored, un-pinned dependency of the reference (README.md:14): the algorithms below are
// Eigen 3.4's (unblocked LLT, partial-pivot LU determinant, AngleAxis/Quaternion conversions) restated.
#pragma once
#include "hkf_hostdev_math.h"

#if defined(__CUDACC__)
#define HKF_UNROLL _Pragma("unroll")
#else
#define HKF_UNROLL
#endif

#define HKF_SOLVE_OK 0
#define HKF_SOLVE_SINGULAR 1
#define HKF_SOLVE_SHAKE 2

HKF_HD void hkf_unpack27(const float* s, float* A, float* b)
{
    // row i of the packed upper triangle starts at i*7 - i*(i-1)/2
    HKF_UNROLL
    for (int i = 0; i < 6; ++i) {
        const int base = i * 7 - (i * (i - 1)) / 2;
        HKF_UNROLL
        for (int j = 0; j < 7; ++j) {
            if (j >= i) {
                float v = s[base + (j - i)];
                if (j == 6) b[i] = v;
                else { A[i * 6 + j] = v; A[j * 6 + i] = v; }
            }
        }
    }
}

// NOTE on loop shapes: nvcc only keeps these 6x6 work arrays in registers when every loop has constant
// bounds (triangular "for (i = k+1; ...)" nests are left partially rolled and force the arrays into local
// memory, which made the single-thread solve ~50 us).  Hence full-range loops with "if (i > k)" guards,
// which fold away after unrolling.
HKF_HD float hkf_det6(const float* Ain)
{
    float a[36];
    HKF_UNROLL
    for (int i = 0; i < 36; ++i) a[i] = Ain[i];
    float sign = 1.0f;
    bool singular = false;
    HKF_UNROLL
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        float best = fabsf(a[k * 6 + k]);
        HKF_UNROLL
        for (int i = 0; i < 6; ++i) {
            if (i > k) {
                float v = fabsf(a[i * 6 + k]);
                if (v > best) { best = v; piv = i; }
            }
        }
        if (best == 0.0f) singular = true;
        HKF_UNROLL
        for (int i = 0; i < 6; ++i) {
            if (i > k) {
                const bool sw = (piv == i);
                HKF_UNROLL
                for (int j = 0; j < 6; ++j) {
                    const float t0 = a[k * 6 + j], t1 = a[i * 6 + j];
                    a[k * 6 + j] = sw ? t1 : t0;
                    a[i * 6 + j] = sw ? t0 : t1;
                }
                sign = sw ? -sign : sign;
            }
        }
        const float pk = a[k * 6 + k];
        HKF_UNROLL
        for (int i = 0; i < 6; ++i) {
            if (i > k) {
                const float f = HKF_DIV(a[i * 6 + k], pk);
                a[i * 6 + k] = f;
                HKF_UNROLL
                for (int j = 0; j < 6; ++j)
                    if (j > k) a[i * 6 + j] = HKF_SUB(a[i * 6 + j], HKF_MUL(f, a[k * 6 + j]));
            }
        }
    }
    if (singular) return 0.0f;
    float p = a[0];
    HKF_UNROLL
    for (int k = 1; k < 6; ++k) p = HKF_MUL(p, a[k * 6 + k]);
    return HKF_MUL(sign, p);
}

HKF_HD void hkf_llt_solve6(const float* A, const float* b, float* x)
{
    float L[36];
    HKF_UNROLL
    for (int i = 0; i < 36; ++i) L[i] = 0.0f;
    HKF_UNROLL
    for (int k = 0; k < 6; ++k) {
        float d = A[k * 6 + k];
        HKF_UNROLL
        for (int j = 0; j < 6; ++j)
            if (j < k) d = HKF_SUB(d, HKF_MUL(L[k * 6 + j], L[k * 6 + j]));
        const float lkk = HKF_SQRT(d);
        L[k * 6 + k] = lkk;
        HKF_UNROLL
        for (int i = 0; i < 6; ++i) {
            if (i > k) {
                float v = A[i * 6 + k];
                HKF_UNROLL
                for (int j = 0; j < 6; ++j)
                    if (j < k) v = HKF_SUB(v, HKF_MUL(L[i * 6 + j], L[k * 6 + j]));
                L[i * 6 + k] = HKF_DIV(v, lkk);
            }
        }
    }
    float y[6];
    HKF_UNROLL
    for (int i = 0; i < 6; ++i) {
        float v = b[i];
        HKF_UNROLL
        for (int j = 0; j < 6; ++j)
            if (j < i) v = HKF_SUB(v, HKF_MUL(L[i * 6 + j], y[j]));
        y[i] = HKF_DIV(v, L[i * 6 + i]);
    }
    HKF_UNROLL
    for (int ii = 0; ii < 6; ++ii) {
        const int i = 5 - ii;
        float v = y[i];
        // ascending j as in the reference restatement: j = i+1 .. 5
        HKF_UNROLL
        for (int j = 0; j < 6; ++j)
            if (j > i) v = HKF_SUB(v, HKF_MUL(L[j * 6 + i], x[j]));
        x[i] = HKF_DIV(v, L[i * 6 + i]);
    }
}

HKF_HD void hkf_mat3_mul(const float* a, const float* b, float* c)
{
    HKF_UNROLL
    for (int i = 0; i < 3; ++i)
        HKF_UNROLL
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = HKF_ADD(HKF_ADD(HKF_MUL(a[i * 3 + 0], b[0 * 3 + j]), HKF_MUL(a[i * 3 + 1], b[1 * 3 + j])), HKF_MUL(a[i * 3 + 2], b[2 * 3 + j]));
}

HKF_HD void hkf_rot_axis(int axis, float ang, float* R)
{
    // Eigen's AngleAxisf uses the float libm (sinf/cosf); host and device libms may differ by an ulp here
    float s = sinf(ang), c = cosf(ang);
    float one = HKF_ADD(HKF_SUB(1.0f, c), c);
    HKF_UNROLL
    for (int i = 0; i < 9; ++i) R[i] = 0.0f;
    if (axis == 0) { R[0] = one; R[4] = c; R[5] = -s; R[7] = s; R[8] = c; }
    else if (axis == 1) { R[4] = one; R[0] = c; R[2] = s; R[6] = -s; R[8] = c; }
    else { R[8] = one; R[0] = c; R[1] = -s; R[3] = s; R[4] = c; }
}

// angle of Eigen::AngleAxisf(R): matrix -> quaternion -> 2*atan2(|vec|, |w|)
HKF_HD float hkf_rotation_angle(const float* m)
{
    float qw, qx, qy, qz;
    float t = HKF_ADD(HKF_ADD(m[0], m[4]), m[8]);
    if (t > 0.0f) {
        t = HKF_SQRT(HKF_ADD(t, 1.0f));
        qw = HKF_MUL(0.5f, t);
        t = HKF_DIV(0.5f, t);
        qx = HKF_MUL(HKF_SUB(m[7], m[5]), t);
        qy = HKF_MUL(HKF_SUB(m[2], m[6]), t);
        qz = HKF_MUL(HKF_SUB(m[3], m[1]), t);
    } else if (m[0] >= m[4] && m[0] >= m[8]) {          // i = 0, j = 1, k = 2
        t = HKF_SQRT(HKF_ADD(HKF_SUB(HKF_SUB(m[0], m[4]), m[8]), 1.0f));
        qx = HKF_MUL(0.5f, t); t = HKF_DIV(0.5f, t);
        qw = HKF_MUL(HKF_SUB(m[7], m[5]), t); qy = HKF_MUL(HKF_ADD(m[3], m[1]), t); qz = HKF_MUL(HKF_ADD(m[6], m[2]), t);
    } else if (m[4] > m[0] && m[4] >= m[8]) {           // i = 1, j = 2, k = 0
        t = HKF_SQRT(HKF_ADD(HKF_SUB(HKF_SUB(m[4], m[8]), m[0]), 1.0f));
        qy = HKF_MUL(0.5f, t); t = HKF_DIV(0.5f, t);
        qw = HKF_MUL(HKF_SUB(m[2], m[6]), t); qz = HKF_MUL(HKF_ADD(m[7], m[5]), t); qx = HKF_MUL(HKF_ADD(m[1], m[3]), t);
    } else {                                            // i = 2, j = 0, k = 1
        t = HKF_SQRT(HKF_ADD(HKF_SUB(HKF_SUB(m[8], m[0]), m[4]), 1.0f));
        qz = HKF_MUL(0.5f, t); t = HKF_DIV(0.5f, t);
        qw = HKF_MUL(HKF_SUB(m[3], m[1]), t); qx = HKF_MUL(HKF_ADD(m[2], m[6]), t); qy = HKF_MUL(HKF_ADD(m[5], m[7]), t);
    }
    float n = HKF_SQRT(HKF_ADD(HKF_ADD(HKF_MUL(qx, qx), HKF_MUL(qy, qy)), HKF_MUL(qz, qz)));
    if (n == 0.0f) return 0.0f;
    return HKF_MUL(2.0f, atan2f(n, fabsf(qw)));
}

// x -> T = [Rx*Ry*Rz | t] with the shake gate; returns false when gated (a NaN passes, as in the reference)
HKF_HD bool hkf_vector6_to_transform(const float* x, float angle_shake, float dist_shake, float* T)
{
    float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
    hkf_rot_axis(0, x[0], Rx); hkf_rot_axis(1, x[1], Ry); hkf_rot_axis(2, x[2], Rz);
    hkf_mat3_mul(Rx, Ry, Rxy);
    hkf_mat3_mul(Rxy, Rz, R);
    float angle = hkf_rotation_angle(R);
    float d = HKF_SQRT(HKF_ADD(HKF_ADD(HKF_MUL(x[3], x[3]), HKF_MUL(x[4], x[4])), HKF_MUL(x[5], x[5])));
    if (angle > angle_shake || d > dist_shake) return false;
    T[0] = R[0]; T[1] = R[1]; T[2] = R[2];  T[3] = x[3];
    T[4] = R[3]; T[5] = R[4]; T[6] = R[5];  T[7] = x[4];
    T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = x[5];
    T[12] = 0.0f; T[13] = 0.0f; T[14] = 0.0f; T[15] = 1.0f;
    return true;
}

HKF_HD float hkf_norm6(const float* x)
{
    float s = HKF_MUL(x[0], x[0]);
    HKF_UNROLL
    for (int i = 1; i < 6; ++i) s = HKF_ADD(s, HKF_MUL(x[i], x[i]));
    return HKF_SQRT(s);
}

// One ICP iteration's host part (minimizePointToPlaneErrFunc + vector6ToTransformMatrix + left-multiply)
// sums27 -> cur <- T_inc * cur.  Returns HKF_SOLVE_*.  Split in two independent halves so that the fused
// kernel can run the determinant gate on one warp while another one solves (the gate's outcome is only
// combined at the end): hkf_icp_singular() and hkf_icp_step().
HKF_HD bool hkf_icp_singular(const float* sums27)
{
    float A[36], b[6];
    hkf_unpack27(sums27, A, b);
    return (double)hkf_det6(A) < 1E-10;
}

HKF_HD int hkf_icp_step(const float* sums27, float angle_shake, float dist_shake, float* cur)
{
    float A[36], b[6], x[6], T[16];
    hkf_unpack27(sums27, A, b);
    hkf_llt_solve6(A, b, x);
    if (!hkf_vector6_to_transform(x, angle_shake, dist_shake, T)) return HKF_SOLVE_SHAKE;
    hkf_mat44_mul(T, cur, cur);
    return HKF_SOLVE_OK;
}

HKF_HD int hkf_icp_update(const float* sums27, float angle_shake, float dist_shake, float* cur)
{
    if (hkf_icp_singular(sums27)) return HKF_SOLVE_SINGULAR;
    return hkf_icp_step(sums27, angle_shake, dist_shake, cur);
}

// se(3) exponential map in double (src/utils/eigen_utils.cpp:84-127, ViSP-derived)
HKF_HD void hkf_exp_map(const double* v, double* Rd, double* dt)
{
    double u0 = v[0], u1 = v[1], u2 = v[2];
    double theta = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
    double si = sin(theta), co = cos(theta);
    double sinc = (fabs(theta) < 1.0e-8) ? 1.0 : (si / theta);
    double mcosc = (fabs(theta) < 2.5e-4) ? 0.5 : ((1.0 - co) / theta / theta);
    double msinc = (fabs(theta) < 2.5e-4) ? (1. / 6.0) : ((1.0 - si / theta) / theta / theta);
    Rd[0] = co + mcosc * u0 * u0;       Rd[1] = -sinc * u2 + mcosc * u0 * u1; Rd[2] = sinc * u1 + mcosc * u0 * u2;
    Rd[3] = sinc * u2 + mcosc * u1 * u0; Rd[4] = co + mcosc * u1 * u1;        Rd[5] = -sinc * u0 + mcosc * u1 * u2;
    Rd[6] = -sinc * u1 + mcosc * u2 * u0; Rd[7] = sinc * u0 + mcosc * u2 * u1; Rd[8] = co + mcosc * u2 * u2;
    dt[0] = v[3] * (sinc + u0 * u0 * msinc) + v[4] * (u0 * u1 * msinc - u2 * mcosc) + v[5] * (u0 * u2 * msinc + u1 * mcosc);
    dt[1] = v[3] * (u0 * u1 * msinc + u2 * mcosc) + v[4] * (sinc + u1 * u1 * msinc) + v[5] * (u1 * u2 * msinc - u0 * mcosc);
    dt[2] = v[3] * (u0 * u2 * msinc - u1 * mcosc) + v[4] * (u1 * u2 * msinc + u0 * mcosc) + v[5] * (sinc + u2 * u2 * msinc);
}

// One SDF-tracker iteration's host part.  Returns HKF_SOLVE_OK / HKF_SOLVE_SHAKE; *converged is set when
// |x| < 1e-3 (the reference breaks BEFORE composing, src/CameraPoseFinderSDF.cpp:87-90).
HKF_HD int hkf_sdf_update(const float* sums27, float angle_shake, float dist_shake, float* cur, int* converged)
{
    float A[36], b[6], x[6], T[16];
    hkf_unpack27(sums27, A, b);
    hkf_llt_solve6(A, b, x);                       // no determinant guard
    *converged = 0;
    if (!hkf_vector6_to_transform(x, angle_shake, dist_shake, T)) return HKF_SOLVE_SHAKE;
    if (hkf_norm6(x) < 0.001f) { *converged = 1; return HKF_SOLVE_OK; }
    double v[6], Rd[9], dt[3];
    for (int i = 0; i < 6; ++i) v[i] = (double)x[i];
    hkf_exp_map(v, Rd, dt);
    float RtT[9], dtf[3];
    for (int i = 0; i < 3; ++i) {
        dtf[i] = (float)dt[i];
        for (int j = 0; j < 3; ++j) RtT[i * 3 + j] = (float)Rd[j * 3 + i];
    }
    float Rc[9] = { cur[0], cur[1], cur[2], cur[4], cur[5], cur[6], cur[8], cur[9], cur[10] };
    float Rn[9], tn[3];
    hkf_mat3_mul(RtT, Rc, Rn);
    for (int i = 0; i < 3; ++i) {
        float rt = HKF_ADD(HKF_ADD(HKF_MUL(RtT[i * 3 + 0], dtf[0]), HKF_MUL(RtT[i * 3 + 1], dtf[1])), HKF_MUL(RtT[i * 3 + 2], dtf[2]));
        tn[i] = HKF_SUB(cur[i * 4 + 3], rt);
    }
    cur[0] = Rn[0]; cur[1] = Rn[1]; cur[2] = Rn[2];  cur[3] = tn[0];
    cur[4] = Rn[3]; cur[5] = Rn[4]; cur[6] = Rn[5];  cur[7] = tn[1];
    cur[8] = Rn[6]; cur[9] = Rn[7]; cur[10] = Rn[8]; cur[11] = tn[2];
    cur[12] = 0.0f; cur[13] = 0.0f; cur[14] = 0.0f; cur[15] = 1.0f;
    return HKF_SOLVE_OK;
}

// Eigen Quaternion(w,x,y,z).toRotationMatrix(), no normalisation (src/CameraPoseFinderFromFile.cpp:74-75)
HKF_HD void hkf_quat_to_mat3(float w, float x, float y, float z, float* R)
{
    float tx = HKF_MUL(2.0f, x), ty = HKF_MUL(2.0f, y), tz = HKF_MUL(2.0f, z);
    float twx = HKF_MUL(tx, w), twy = HKF_MUL(ty, w), twz = HKF_MUL(tz, w);
    float txx = HKF_MUL(tx, x), txy = HKF_MUL(ty, x), txz = HKF_MUL(tz, x);
    float tyy = HKF_MUL(ty, y), tyz = HKF_MUL(tz, y), tzz = HKF_MUL(tz, z);
    R[0] = HKF_SUB(1.0f, HKF_ADD(tyy, tzz)); R[1] = HKF_SUB(txy, twz); R[2] = HKF_ADD(txz, twy);
    R[3] = HKF_ADD(txy, twz); R[4] = HKF_SUB(1.0f, HKF_ADD(txx, tzz)); R[5] = HKF_SUB(tyz, twx);
    R[6] = HKF_SUB(txz, twy); R[7] = HKF_ADD(tyz, twx); R[8] = HKF_SUB(1.0f, HKF_ADD(txx, tyy));
}
