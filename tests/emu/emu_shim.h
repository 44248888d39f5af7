// emu_shim.h — compile device kernels of hybkinectfu_b200/csrc for the HOST (g++ -ffp-contract=off), one simulated
// thread at a time, so that kernel SOURCE can be checked against the CPU oracle without a GPU.  Test infrastructure:
// only kernels without warp collectives / shared memory qualify (the ray march).  Every intrinsic used by those kernels
// is an IEEE single operation on both sides: fma -> fmaf, div/sqrt/rcp round-to-nearest, truncating conversions.
#pragma once
#define HKF_EMULATE_DEVICE 1
#include <cuda_runtime.h>      // host API + vector types; the __device__/__global__ annotations become ignored attributes
#include <math.h>
#include <stdint.h>
#include <string.h>

#pragma GCC diagnostic ignored "-Wattributes"

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

// the simulated thread's coordinates; emu_indices.h (included AFTER the csrc headers, whose host code has members called
// gridDim / blockDim) maps threadIdx & co. onto them
struct EmuIdx { unsigned x, y, z; };
static thread_local EmuIdx emu_threadIdx, emu_blockIdx, emu_blockDim, emu_gridDim;

static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
// __expf(x) = ex2.approx(x * log2(e)): exp2f of the same product stands in for the MUFU unit, as in the CPU oracle
#define __expf(x) exp2f((x) * 1.4426950216293334961f)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline int __float2int_rz(float a)
{   // cvt.rzi.s32.f32: saturating, NaN -> 0
    if (a != a) return 0;
    if (a >= 2147483648.0f) return 2147483647;
    if (a <= -2147483648.0f) return (-2147483647 - 1);
    return (int)a;
}
static inline unsigned __float2uint_rz(float a)
{   // cvt.rzi.u32.f32: saturating, NaN -> 0
    if (a != a || a <= 0.0f) return 0u;
    if (a >= 4294967296.0f) return 0xffffffffu;
    return (unsigned)a;
}
static inline int __double2int_rz(double a)
{
    if (a != a) return 0;
    if (a >= 2147483648.0) return 2147483647;
    if (a <= -2147483648.0) return (-2147483647 - 1);
    return (int)a;
}
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }
// atomics: real ones, the block-level emulator runs the threads of a block concurrently
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicMax(unsigned* p, unsigned v)
{
    unsigned o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline unsigned atomicMin(unsigned* p, unsigned v)
{
    unsigned o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
static inline void __threadfence_system() {}
static inline void __threadfence() {}
static inline long long clock64() { return 0; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
