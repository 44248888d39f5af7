/* TEST INFRASTRUCTURE ONLY — part of the parity oracle, never used by the product path.
 *
 * host_math.h: plain-C restatement of the host-side arithmetic that the reference
 * delegates to Eigen3 (third-party, un-vendored, NO version pinned anywhere in the
 * reference: README.md:14 only says "Eigen3").  PARITY UNPINNED at this boundary: the
 * reference holds no test/golden vector for any of these functions and Eigen is absent
 * from this image, so what follows restates the published algorithms of Eigen 3.4.0
 * (unblocked LLT, PartialPivLU determinant, AngleAxis/Quaternion conversions) and the
 * ViSP-derived exponential map that IS in the reference (src/utils/eigen_utils.cpp:39-127).
 *
 * All single-precision arithmetic here must be compiled with -ffp-contract=off so that
 * the host compiler never fuses a*b+c (the reference's host build has no build flags at
 * all, i.e. x86-64 baseline without FMA).
 *
 * Reference call sites:
 *   CameraPoseFinderICP::minimizePointToPlaneErrFunc   src/CameraPoseFinderICP.cpp:113-145
 *   CameraPoseFinderICP::vector6ToTransformMatrix      src/CameraPoseFinderICP.cpp:95-111
 *   CameraPoseFinderSDF::estimateCameraPose            src/CameraPoseFinderSDF.cpp:44-106
 *   CameraPoseFinderFromFile::estimateCameraPose       src/CameraPoseFinderFromFile.cpp:66-89
 *   Mat44::getInverse / operator*                      src/cuda/Mat.h:230-253,319-444
 */
#ifndef HKF_ORACLE_HOST_MATH_H
#define HKF_ORACLE_HOST_MATH_H
#include <math.h>
#include <string.h>

#ifdef __cplusplus
#define HM_INLINE static inline
#else
#define HM_INLINE static inline
#endif

/* unpack the 27 upper-triangular sums into A (6x6 sym, row-major) and b (6)
 * src/CameraPoseFinderICP.cpp:118-136 */
HM_INLINE void hm_unpack27(const float s[27], float A[36], float b[6])
{
    int shift = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j) {
            float v = s[shift++];
            if (j == 6) b[i] = v;
            else { A[i * 6 + j] = v; A[j * 6 + i] = v; }
        }
}

/* determinant via partial-pivot LU (Eigen: determinant() for n>4 -> PartialPivLU) */
HM_INLINE float hm_det6(const float Ain[36])
{
    float a[36];
    memcpy(a, Ain, sizeof(a));
    float det = 1.0f;
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        float best = fabsf(a[k * 6 + k]);
        for (int i = k + 1; i < 6; ++i) {
            float v = fabsf(a[i * 6 + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (best == 0.0f) return 0.0f;
        if (piv != k) {
            for (int j = 0; j < 6; ++j) { float t = a[k * 6 + j]; a[k * 6 + j] = a[piv * 6 + j]; a[piv * 6 + j] = t; }
            det = -det;
        }
        float pk = a[k * 6 + k];
        for (int i = k + 1; i < 6; ++i) {
            float f = a[i * 6 + k] / pk;
            a[i * 6 + k] = f;
            for (int j = k + 1; j < 6; ++j) a[i * 6 + j] = a[i * 6 + j] - f * a[k * 6 + j];
        }
    }
    /* product of the diagonal, in index order */
    float p = a[0];
    for (int k = 1; k < 6; ++k) p = p * a[k * 6 + k];
    return det * p;
}

/* unblocked Cholesky A = L L^T (lower), then forward/back substitution; float32 throughout.
 * A non-positive pivot is NOT trapped (Eigen's solve() proceeds regardless of info()), so a
 * singular system yields NaN/Inf exactly like SURVEY.md quirk 7 describes. */
HM_INLINE void hm_llt_solve6(const float Ain[36], const float b[6], float x[6])
{
    float L[36];
    memset(L, 0, sizeof(L));
    for (int k = 0; k < 6; ++k) {
        float d = Ain[k * 6 + k];
        for (int j = 0; j < k; ++j) d = d - L[k * 6 + j] * L[k * 6 + j];
        float lkk = sqrtf(d);
        L[k * 6 + k] = lkk;
        for (int i = k + 1; i < 6; ++i) {
            float v = Ain[i * 6 + k];
            for (int j = 0; j < k; ++j) v = v - L[i * 6 + j] * L[k * 6 + j];
            L[i * 6 + k] = v / lkk;
        }
    }
    float y[6];
    for (int i = 0; i < 6; ++i) {
        float v = b[i];
        for (int j = 0; j < i; ++j) v = v - L[i * 6 + j] * y[j];
        y[i] = v / L[i * 6 + i];
    }
    for (int i = 5; i >= 0; --i) {
        float v = y[i];
        for (int j = i + 1; j < 6; ++j) v = v - L[j * 6 + i] * x[j];
        x[i] = v / L[i * 6 + i];
    }
}

/* ICP: det guard then LLT.  returns 0 ok, 1 singular.  src/CameraPoseFinderICP.cpp:138-144 */
HM_INLINE int hm_solve_icp(const float s27[27], float x[6])
{
    float A[36], b[6];
    hm_unpack27(s27, A, b);
    if (hm_det6(A) < 1E-10) return 1;
    hm_llt_solve6(A, b, x);
    return 0;
}
/* SDF: LLT only.  src/CameraPoseFinderSDF.cpp:62-79 */
HM_INLINE void hm_solve_llt(const float s27[27], float x[6])
{
    float A[36], b[6];
    hm_unpack27(s27, A, b);
    hm_llt_solve6(A, b, x);
}
HM_INLINE float hm_norm6(const float x[6])
{
    float s = x[0] * x[0];
    for (int i = 1; i < 6; ++i) s = s + x[i] * x[i];
    return sqrtf(s);
}

HM_INLINE void hm_mat3_mul(const float a[9], const float b[9], float c[9])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = (a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j]) + a[i * 3 + 2] * b[2 * 3 + j];
}

/* Eigen AngleAxis(angle, UnitX/Y/Z).toRotationMatrix(): for a unit coordinate axis the general
 * formula degenerates to the textbook elementary rotations with s=sinf(a), c=cosf(a), and the
 * diagonal entry on the axis is (1-c)+c. */
HM_INLINE void hm_rot_axis(int axis, float ang, float R[9])
{
    float s = sinf(ang), c = cosf(ang);
    float one = (1.0f - c) + c;
    for (int i = 0; i < 9; ++i) R[i] = 0.0f;
    if (axis == 0) { R[0] = one; R[4] = c; R[5] = -s; R[7] = s; R[8] = c; }
    else if (axis == 1) { R[4] = one; R[0] = c; R[2] = s; R[6] = -s; R[8] = c; }
    else { R[8] = one; R[0] = c; R[1] = -s; R[3] = s; R[4] = c; }
}

/* angle of AngleAxisf(R): Eigen goes matrix -> quaternion -> 2*atan2(|vec|, |w|) */
HM_INLINE float hm_rotation_angle(const float m[9])
{
    float qw, qx, qy, qz;
    float t = (m[0] + m[4]) + m[8];
    if (t > 0.0f) {
        t = sqrtf(t + 1.0f);
        qw = 0.5f * t;
        t = 0.5f / t;
        qx = (m[7] - m[5]) * t;
        qy = (m[2] - m[6]) * t;
        qz = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        float q[3];
        t = sqrtf(((m[i * 3 + i] - m[j * 3 + j]) - m[k * 3 + k]) + 1.0f);
        q[i] = 0.5f * t;
        t = 0.5f / t;
        qw = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    float n = sqrtf((qx * qx + qy * qy) + qz * qz);
    if (n == 0.0f) return 0.0f;
    return 2.0f * atan2f(n, fabsf(qw));
}

/* x = (alpha,beta,gamma, tx,ty,tz) -> T = [Rx*Ry*Rz | t], with the "camera shaking" gate.
 * returns 1 on success, 0 when gated.  T is a row-major Mat44.
 * src/CameraPoseFinderICP.cpp:95-111, src/CameraPoseFinderSDF.cpp:24-43 */
HM_INLINE int hm_vector6_to_transform(const float x[6], float angle_shake, float dist_shake, float T[16])
{
    float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
    hm_rot_axis(0, x[0], Rx); hm_rot_axis(1, x[1], Ry); hm_rot_axis(2, x[2], Rz);
    hm_mat3_mul(Rx, Ry, Rxy);
    hm_mat3_mul(Rxy, Rz, R);
    float angle = hm_rotation_angle(R);
    float d = sqrtf((x[3] * x[3] + x[4] * x[4]) + x[5] * x[5]);
    /* NaN compares false on both tests, i.e. a NaN solution PASSES the gate (quirk 7) */
    if (angle > angle_shake || d > dist_shake) return 0;
    T[0] = R[0]; T[1] = R[1]; T[2] = R[2];  T[3] = x[3];
    T[4] = R[3]; T[5] = R[4]; T[6] = R[5];  T[7] = x[4];
    T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = x[5];
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
    return 1;
}

/* Mat44 * Mat44, row-major, plain left-to-right sums (src/cuda/Mat.h:230-253, host build: no FMA) */
HM_INLINE void hm_mat44_mul(const float a[16], const float b[16], float c[16])
{
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r[i * 4 + j] = ((a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j]) + a[i * 4 + 2] * b[2 * 4 + j]) + a[i * 4 + 3] * b[3 * 4 + j];
    memcpy(c, r, sizeof(r));
}

/* general cofactor inverse, float (src/cuda/Mat.h:319-444); term order as in the source */
HM_INLINE void hm_mat44_inverse(const float e[16], float out[16])
{
    float inv[16];
    inv[0] = e[5] * e[10] * e[15] - e[5] * e[11] * e[14] - e[9] * e[6] * e[15] + e[9] * e[7] * e[14] + e[13] * e[6] * e[11] - e[13] * e[7] * e[10];
    inv[4] = -e[4] * e[10] * e[15] + e[4] * e[11] * e[14] + e[8] * e[6] * e[15] - e[8] * e[7] * e[14] - e[12] * e[6] * e[11] + e[12] * e[7] * e[10];
    inv[8] = e[4] * e[9] * e[15] - e[4] * e[11] * e[13] - e[8] * e[5] * e[15] + e[8] * e[7] * e[13] + e[12] * e[5] * e[11] - e[12] * e[7] * e[9];
    inv[12] = -e[4] * e[9] * e[14] + e[4] * e[10] * e[13] + e[8] * e[5] * e[14] - e[8] * e[6] * e[13] - e[12] * e[5] * e[10] + e[12] * e[6] * e[9];
    inv[1] = -e[1] * e[10] * e[15] + e[1] * e[11] * e[14] + e[9] * e[2] * e[15] - e[9] * e[3] * e[14] - e[13] * e[2] * e[11] + e[13] * e[3] * e[10];
    inv[5] = e[0] * e[10] * e[15] - e[0] * e[11] * e[14] - e[8] * e[2] * e[15] + e[8] * e[3] * e[14] + e[12] * e[2] * e[11] - e[12] * e[3] * e[10];
    inv[9] = -e[0] * e[9] * e[15] + e[0] * e[11] * e[13] + e[8] * e[1] * e[15] - e[8] * e[3] * e[13] - e[12] * e[1] * e[11] + e[12] * e[3] * e[9];
    inv[13] = e[0] * e[9] * e[14] - e[0] * e[10] * e[13] - e[8] * e[1] * e[14] + e[8] * e[2] * e[13] + e[12] * e[1] * e[10] - e[12] * e[2] * e[9];
    inv[2] = e[1] * e[6] * e[15] - e[1] * e[7] * e[14] - e[5] * e[2] * e[15] + e[5] * e[3] * e[14] + e[13] * e[2] * e[7] - e[13] * e[3] * e[6];
    inv[6] = -e[0] * e[6] * e[15] + e[0] * e[7] * e[14] + e[4] * e[2] * e[15] - e[4] * e[3] * e[14] - e[12] * e[2] * e[7] + e[12] * e[3] * e[6];
    inv[10] = e[0] * e[5] * e[15] - e[0] * e[7] * e[13] - e[4] * e[1] * e[15] + e[4] * e[3] * e[13] + e[12] * e[1] * e[7] - e[12] * e[3] * e[5];
    inv[14] = -e[0] * e[5] * e[14] + e[0] * e[6] * e[13] + e[4] * e[1] * e[14] - e[4] * e[2] * e[13] - e[12] * e[1] * e[6] + e[12] * e[2] * e[5];
    inv[3] = -e[1] * e[6] * e[11] + e[1] * e[7] * e[10] + e[5] * e[2] * e[11] - e[5] * e[3] * e[10] - e[9] * e[2] * e[7] + e[9] * e[3] * e[6];
    inv[7] = e[0] * e[6] * e[11] - e[0] * e[7] * e[10] - e[4] * e[2] * e[11] + e[4] * e[3] * e[10] + e[8] * e[2] * e[7] - e[8] * e[3] * e[6];
    inv[11] = -e[0] * e[5] * e[11] + e[0] * e[7] * e[9] + e[4] * e[1] * e[11] - e[4] * e[3] * e[9] - e[8] * e[1] * e[7] + e[8] * e[3] * e[5];
    inv[15] = e[0] * e[5] * e[10] - e[0] * e[6] * e[9] - e[4] * e[1] * e[10] + e[4] * e[2] * e[9] + e[8] * e[1] * e[6] - e[8] * e[2] * e[5];
    float det = e[0] * inv[0] + e[1] * inv[4] + e[2] * inv[8] + e[3] * inv[12];
    float detr = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * detr;
}

/* ---- se(3) exponential map, double (src/utils/eigen_utils.cpp:39-127, ViSP-derived) ---- */
HM_INLINE double hm_f_sinc(double sinx, double x) { return (fabs(x) < 1.0e-8) ? 1.0 : (sinx / x); }
HM_INLINE double hm_f_mcosc(double cosx, double x) { return (fabs(x) < 2.5e-4) ? 0.5 : ((1.0 - cosx) / x / x); }
HM_INLINE double hm_f_msinc(double sinx, double x) { return (fabs(x) < 2.5e-4) ? (1. / 6.0) : ((1.0 - sinx / x) / x / x); }

/* v = (u0,u1,u2, v3,v4,v5), delta_t = 1.  Rd row-major 3x3, dt 3 */
HM_INLINE void hm_direct_exponential_map(const double v[6], double Rd[9], double dt[3])
{
    double u0 = v[0], u1 = v[1], u2 = v[2];
    double theta = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
    double si = sin(theta), co = cos(theta);
    double sinc = hm_f_sinc(si, theta), mcosc = hm_f_mcosc(co, theta), msinc = hm_f_msinc(si, theta);
    Rd[0] = co + mcosc * u0 * u0;
    Rd[1] = -sinc * u2 + mcosc * u0 * u1;
    Rd[2] = sinc * u1 + mcosc * u0 * u2;
    Rd[3] = sinc * u2 + mcosc * u1 * u0;
    Rd[4] = co + mcosc * u1 * u1;
    Rd[5] = -sinc * u0 + mcosc * u1 * u2;
    Rd[6] = -sinc * u1 + mcosc * u2 * u0;
    Rd[7] = sinc * u0 + mcosc * u2 * u1;
    Rd[8] = co + mcosc * u2 * u2;
    dt[0] = v[3] * (sinc + u0 * u0 * msinc) + v[4] * (u0 * u1 * msinc - u2 * mcosc) + v[5] * (u0 * u2 * msinc + u1 * mcosc);
    dt[1] = v[3] * (u0 * u1 * msinc + u2 * mcosc) + v[4] * (sinc + u1 * u1 * msinc) + v[5] * (u1 * u2 * msinc - u0 * mcosc);
    dt[2] = v[3] * (u0 * u2 * msinc - u1 * mcosc) + v[4] * (u1 * u2 * msinc + u0 * mcosc) + v[5] * (sinc + u2 * u2 * msinc);
}

/* SDF tracker pose update (src/CameraPoseFinderSDF.cpp:91-100):
 *   R' = Rd^T(float) * R ;  t' = t - Rd^T(float) * dt(float)
 * Affine3d::rotation() in Eigen runs an SVD polar decomposition; for an exact rotation it returns
 * the input to ~1e-16, which vanishes in the cast to float, so Rd is used directly. */
HM_INLINE void hm_sdf_compose(const float cur[16], const float x[6], float next[16])
{
    double v[6], Rd[9], dt[3];
    for (int i = 0; i < 6; ++i) v[i] = (double)x[i];
    hm_direct_exponential_map(v, Rd, dt);
    float RtT[9], dtf[3];
    for (int i = 0; i < 3; ++i) {
        dtf[i] = (float)dt[i];
        for (int j = 0; j < 3; ++j) RtT[i * 3 + j] = (float)Rd[j * 3 + i];
    }
    float Rc[9] = { cur[0], cur[1], cur[2], cur[4], cur[5], cur[6], cur[8], cur[9], cur[10] };
    float Rn[9];
    hm_mat3_mul(RtT, Rc, Rn);
    float tn[3];
    for (int i = 0; i < 3; ++i) {
        float rt = (RtT[i * 3 + 0] * dtf[0] + RtT[i * 3 + 1] * dtf[1]) + RtT[i * 3 + 2] * dtf[2];
        tn[i] = cur[i * 4 + 3] - rt;
    }
    next[0] = Rn[0]; next[1] = Rn[1]; next[2] = Rn[2];  next[3] = tn[0];
    next[4] = Rn[3]; next[5] = Rn[4]; next[6] = Rn[5];  next[7] = tn[1];
    next[8] = Rn[6]; next[9] = Rn[7]; next[10] = Rn[8]; next[11] = tn[2];
    next[12] = 0; next[13] = 0; next[14] = 0; next[15] = 1;
}

/* Eigen Quaternion(w,x,y,z).toRotationMatrix(), no normalisation (src/CameraPoseFinderFromFile.cpp:74-75) */
HM_INLINE void hm_quat_to_mat3(float w, float x, float y, float z, float R[9])
{
    float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
    float twx = tx * w, twy = ty * w, twz = tz * w;
    float txx = tx * x, txy = ty * x, txz = tz * x;
    float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0f - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0f - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0f - (txx + tyy);
}

#endif
