// hkf_hostdev_math.h — small matrix helpers that must give bit-identical results on the host
// (g++, compiled with -ffp-contract=off) and on the device (explicit round-to-nearest intrinsics,
// which nvcc never contracts).  Used for the pose algebra that the reference performs on the host
// (Mat44::getInverse / operator*, src/cuda/Mat.h:230-253,319-444) and that the fused paths perform
// on the device so that the pose never leaves HBM.
#pragma once
#include <math.h>

#if defined(__CUDA_ARCH__)
#define HKF_HD __host__ __device__ __forceinline__
#define HKF_MUL(a, b) __fmul_rn((a), (b))
#define HKF_ADD(a, b) __fadd_rn((a), (b))
#define HKF_SUB(a, b) __fsub_rn((a), (b))
#define HKF_DIV(a, b) __fdiv_rn((a), (b))
#define HKF_SQRT(a) __fsqrt_rn((a))
#elif defined(__CUDACC__)
#define HKF_HD __host__ __device__ __forceinline__
#define HKF_MUL(a, b) ((a) * (b))
#define HKF_ADD(a, b) ((a) + (b))
#define HKF_SUB(a, b) ((a) - (b))
#define HKF_DIV(a, b) ((a) / (b))
#define HKF_SQRT(a) sqrtf((a))
#else
#define HKF_HD static inline
#define HKF_MUL(a, b) ((a) * (b))
#define HKF_ADD(a, b) ((a) + (b))
#define HKF_SUB(a, b) ((a) - (b))
#define HKF_DIV(a, b) ((a) / (b))
#define HKF_SQRT(a) sqrtf((a))
#endif

#if defined(__CUDACC__)
#define HKF_UNROLL4 _Pragma("unroll")
#else
#define HKF_UNROLL4
#endif

// three-factor product (a*b)*c
#define HKF_M3(a, b, c) HKF_MUL(HKF_MUL((a), (b)), (c))

// c = a * b, row-major 4x4, each entry ((a0*b0 + a1*b1) + a2*b2) + a3*b3 without fusion
HKF_HD void hkf_mat44_mul(const float* a, const float* b, float* c)
{
    float r[16];
    HKF_UNROLL4
    for (int i = 0; i < 4; ++i)
        HKF_UNROLL4
        for (int j = 0; j < 4; ++j) {
            float s = HKF_ADD(HKF_MUL(a[i * 4 + 0], b[0 * 4 + j]), HKF_MUL(a[i * 4 + 1], b[1 * 4 + j]));
            s = HKF_ADD(s, HKF_MUL(a[i * 4 + 2], b[2 * 4 + j]));
            s = HKF_ADD(s, HKF_MUL(a[i * 4 + 3], b[3 * 4 + j]));
            r[i * 4 + j] = s;
        }
    HKF_UNROLL4
    for (int i = 0; i < 16; ++i) c[i] = r[i];
}

// one cofactor: s0*(a0*b0*c0) s1 (a1*b1*c1) ... six signed triple products summed left to right,
// exactly the shape of every inv[k] expression in Mat44::getInverse
HKF_HD float hkf_cof6(const float* e, int a0, int b0, int c0, int a1, int b1, int c1, int a2, int b2, int c2,
                      int a3, int b3, int c3, int a4, int b4, int c4, int a5, int b5, int c5, bool lead_neg)
{
    // pattern when lead_neg == false:  + - - + + -   ; when true:  - + + - - +
    float t0 = HKF_M3(lead_neg ? -e[a0] : e[a0], e[b0], e[c0]);
    float t1 = HKF_M3(e[a1], e[b1], e[c1]);
    float t2 = HKF_M3(e[a2], e[b2], e[c2]);
    float t3 = HKF_M3(e[a3], e[b3], e[c3]);
    float t4 = HKF_M3(e[a4], e[b4], e[c4]);
    float t5 = HKF_M3(e[a5], e[b5], e[c5]);
    float s;
    if (!lead_neg) {
        s = HKF_SUB(t0, t1); s = HKF_SUB(s, t2); s = HKF_ADD(s, t3); s = HKF_ADD(s, t4); s = HKF_SUB(s, t5);
    } else {
        s = HKF_ADD(t0, t1); s = HKF_ADD(s, t2); s = HKF_SUB(s, t3); s = HKF_SUB(s, t4); s = HKF_ADD(s, t5);
    }
    return s;
}

// general 4x4 inverse by cofactors in float, term order of Mat44::getInverse (src/cuda/Mat.h:319-444)
HKF_HD void hkf_mat44_inverse(const float* e, float* out)
{
    float inv[16];
    inv[0]  = hkf_cof6(e, 5,10,15, 5,11,14, 9,6,15, 9,7,14, 13,6,11, 13,7,10, false);
    inv[4]  = hkf_cof6(e, 4,10,15, 4,11,14, 8,6,15, 8,7,14, 12,6,11, 12,7,10, true);
    inv[8]  = hkf_cof6(e, 4,9,15, 4,11,13, 8,5,15, 8,7,13, 12,5,11, 12,7,9, false);
    inv[12] = hkf_cof6(e, 4,9,14, 4,10,13, 8,5,14, 8,6,13, 12,5,10, 12,6,9, true);
    inv[1]  = hkf_cof6(e, 1,10,15, 1,11,14, 9,2,15, 9,3,14, 13,2,11, 13,3,10, true);
    inv[5]  = hkf_cof6(e, 0,10,15, 0,11,14, 8,2,15, 8,3,14, 12,2,11, 12,3,10, false);
    inv[9]  = hkf_cof6(e, 0,9,15, 0,11,13, 8,1,15, 8,3,13, 12,1,11, 12,3,9, true);
    inv[13] = hkf_cof6(e, 0,9,14, 0,10,13, 8,1,14, 8,2,13, 12,1,10, 12,2,9, false);
    inv[2]  = hkf_cof6(e, 1,6,15, 1,7,14, 5,2,15, 5,3,14, 13,2,7, 13,3,6, false);
    inv[6]  = hkf_cof6(e, 0,6,15, 0,7,14, 4,2,15, 4,3,14, 12,2,7, 12,3,6, true);
    inv[10] = hkf_cof6(e, 0,5,15, 0,7,13, 4,1,15, 4,3,13, 12,1,7, 12,3,5, false);
    inv[14] = hkf_cof6(e, 0,5,14, 0,6,13, 4,1,14, 4,2,13, 12,1,6, 12,2,5, true);
    inv[3]  = hkf_cof6(e, 1,6,11, 1,7,10, 5,2,11, 5,3,10, 9,2,7, 9,3,6, true);
    inv[7]  = hkf_cof6(e, 0,6,11, 0,7,10, 4,2,11, 4,3,10, 8,2,7, 8,3,6, false);
    inv[11] = hkf_cof6(e, 0,5,11, 0,7,9, 4,1,11, 4,3,9, 8,1,7, 8,3,5, true);
    inv[15] = hkf_cof6(e, 0,5,10, 0,6,9, 4,1,10, 4,2,9, 8,1,6, 8,2,5, false);
    float det = HKF_ADD(HKF_ADD(HKF_ADD(HKF_MUL(e[0], inv[0]), HKF_MUL(e[1], inv[4])), HKF_MUL(e[2], inv[8])), HKF_MUL(e[3], inv[12]));
    float detr = HKF_DIV(1.0f, det);
    HKF_UNROLL4
    for (int i = 0; i < 16; ++i) out[i] = HKF_MUL(inv[i], detr);
}
