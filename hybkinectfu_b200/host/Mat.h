// Mat.h — Mat44 as the host classes use it (src/cuda/Mat.h:160-454): 16 row-major floats, a pose maps camera ->
// world with the translation in m14, m24, m34.  Layout-identical to HkfMat44; the arithmetic (operator*, getInverse)
// is the uncontracted float arithmetic of csrc/hkf_hostdev_math.h, i.e. the reference's host results bit for bit.
#pragma once
#include <cstring>
#include <ostream>
#include "../../include/hkf.h"
#include "../csrc/hkf_hostdev_math.h"

struct float3_ { float x, y, z; };

struct Mat44 {
    float m[16];
    static Mat44 getIdentity() { Mat44 r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f; return r; }
    void setTranslation(float x, float y, float z) { m[3] = x; m[7] = y; m[11] = z; }
    float3_ getTranslation() const { return float3_{ m[3], m[7], m[11] }; }
    void setRotation(const float* r9) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i * 4 + j] = r9[i * 3 + j]; }
    void getRotation(float* r9) const { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r9[i * 3 + j] = m[i * 4 + j]; }
    Mat44 operator*(const Mat44& o) const { Mat44 r; hkf_mat44_mul(m, o.m, r.m); return r; }
    Mat44 getInverse() const { Mat44 r; hkf_mat44_inverse(m, r.m); return r; }
    const HkfMat44* abi() const { return reinterpret_cast<const HkfMat44*>(this); }
    HkfMat44* abi() { return reinterpret_cast<HkfMat44*>(this); }
};
static_assert(sizeof(Mat44) == sizeof(HkfMat44), "Mat44 crosses the C ABI");

inline std::ostream& operator<<(std::ostream& os, const Mat44& a)
{
    for (int i = 0; i < 4; ++i) os << a.m[i * 4] << " " << a.m[i * 4 + 1] << " " << a.m[i * 4 + 2] << " " << a.m[i * 4 + 3] << "\n";
    return os;
}
