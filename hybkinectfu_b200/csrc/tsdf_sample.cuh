// tsdf_sample.cuh — read access to the packed TSDF slab: nearest voxel and trilinear interpolation.
//
// Replaces tsdfvolume::worldPosToVoxel / getVoxel / interpolateSDF / getInterpolateParams
// (src/cuda/tsdfVolume.h:50-56,81-120,151-174) over the packed {tsdf,weight} layout.
//
// Arithmetic contract (PTX + SASS of the reference's raycast / SDF / marching-cubes kernels):
//   g     = cvt.rzi((p * (float)R) / S)                       mul, div.rn, truncating convert
//   cell  = S / (float)R                                      div.rn
//   ctr   = ((float)g + 0.5f) * cell ;  g -= (p < ctr)        add, mul, compare
//   a     = fma((float)g + 0.5f, -cell, p) / cell             ptxas fuses the mul+sub; div.rn
//   dist  = fma(1-c, (1-b)*((1-a)*v000), c*((1-b)*((1-a)*v001)))
//           then fma(1-c, b*((1-a)*v010), .), fma(c, b*((1-a)*v011), .), fma(1-c, (1-b)*(a*v100), .),
//           fma(c, (1-b)*(a*v101), .), fma(1-c, b*(a*v110), .), fma(c, b*(a*v111), .)
//   each of the eight weights is tested == 0 (unobserved) in the order 000,001,010,011,100,101,110,111
#pragma once
#include "hkf_internal.cuh"

struct VolView {
    const float2* vol;   // plane z_lo is at offset 0
    int R;
    float S, cell, Rf;
    int z_lo, z_hi;      // stored planes
    // peer-memory halo (sharded mode, hkf_shard_peer_halo): planes [kz_lo, z_lo) are read from the lower neighbour's
    // memory (vlo points at its plane kz_lo), planes [z_hi, kz_hi) from the upper neighbour's (vhi points at its plane
    // z_hi) - over NVLink, instead of being stored and integrated twice.  Without it kz_lo = z_lo, kz_hi = z_hi.
    const float2* vlo; const float2* vhi;
    int kz_lo, kz_hi;    // planes this rank can READ
};

static inline VolView make_volview(const hkf_ctx* c)
{
    VolView V;
    V.vol = c->vol; V.R = (int)c->ip.volume_resolution; V.S = c->ip.volume_size;
    V.Rf = (float)V.R; V.cell = V.S / V.Rf;
    V.z_lo = (int)c->z_lo; V.z_hi = (int)c->z_hi;
    V.vlo = c->halo_lo; V.vhi = c->halo_hi;
    V.kz_lo = c->halo_lo ? (int)c->halo_z_lo : V.z_lo; V.kz_hi = c->halo_hi ? (int)c->halo_z_hi : V.z_hi;
    return V;
}

#if defined(__CUDACC__) || defined(HKF_EMULATE_DEVICE)
__device__ __forceinline__ int vol_w2v(const VolView& V, float p)
{
    return __float2int_rz(__fdiv_rn(__fmul_rn(p, V.Rf), V.S));
}

__device__ __forceinline__ size_t vol_index(const VolView& V, int x, int y, int z)
{
    return ((size_t)(z - V.z_lo) * V.R + y) * V.R + x;
}

// voxel (0, 0, z) of a READABLE plane z in [kz_lo, kz_hi): this rank's memory or a neighbour's
__device__ __forceinline__ const float2* vol_plane(const VolView& V, int z)
{
    const size_t pl = (size_t)V.R * V.R;
    if (z < V.z_lo) return V.vlo + (size_t)(z - V.kz_lo) * pl;
    if (z >= V.z_hi) return V.vhi + (size_t)(z - V.z_hi) * pl;
    return V.vol + (size_t)(z - V.z_lo) * pl;
}

// SLAB: planes outside [z_lo, z_hi) are not stored; a read there reports "not available" (-1)
// return: 1 ok, 0 reference semantics "false" (unobserved / out of interpolation range), -1 outside slab
template <bool SLAB>
__device__ __forceinline__ int vol_interp_sdf(const VolView& V, float px, float py, float pz, float& dist)
{
    int gx = vol_w2v(V, px), gy = vol_w2v(V, py), gz = vol_w2v(V, pz);
    const int R = V.R;
    if (gx <= 0 || gx >= R - 1) return 0;
    if (gy <= 0 || gy >= R - 1) return 0;
    if (gz <= 0 || gz >= R - 1) return 0;
    const float cell = V.cell;
    const float vx = __fmul_rn(__fadd_rn((float)gx, 0.5f), cell);
    const float vy = __fmul_rn(__fadd_rn((float)gy, 0.5f), cell);
    const float vz = __fmul_rn(__fadd_rn((float)gz, 0.5f), cell);
    gx = (px < vx) ? gx - 1 : gx;
    gy = (py < vy) ? gy - 1 : gy;
    gz = (pz < vz) ? gz - 1 : gz;
    if (SLAB && (gz < V.kz_lo || gz + 1 >= V.kz_hi)) return -1;
    const float a = __fdiv_rn(__fmaf_rn(__fadd_rn((float)gx, 0.5f), -cell, px), cell);
    const float b = __fdiv_rn(__fmaf_rn(__fadd_rn((float)gy, 0.5f), -cell, py), cell);
    const float c = __fdiv_rn(__fmaf_rn(__fadd_rn((float)gz, 0.5f), -cell, pz), cell);
    const size_t sy = (size_t)R;
    // the two planes of the cell (in sharded mode they may live in different ranks' memory)
    const float2* p0 = SLAB ? vol_plane(V, gz) + ((size_t)gy * R + gx) : V.vol + vol_index(V, gx, gy, gz);
    const float2* p1 = SLAB ? vol_plane(V, gz + 1) + ((size_t)gy * R + gx) : p0 + (size_t)R * R;
    // issue all eight loads before testing: independent requests, same 2 rows x 2 planes
    const float2 v000 = __ldg(p0), v100 = __ldg(p0 + 1);
    const float2 v010 = __ldg(p0 + sy), v110 = __ldg(p0 + sy + 1);
    const float2 v001 = __ldg(p1), v101 = __ldg(p1 + 1);
    const float2 v011 = __ldg(p1 + sy), v111 = __ldg(p1 + sy + 1);
    if (v000.y == 0.f || v001.y == 0.f || v010.y == 0.f || v011.y == 0.f ||
        v100.y == 0.f || v101.y == 0.f || v110.y == 0.f || v111.y == 0.f) return 0;
    const float ia = __fsub_rn(1.0f, a), ib = __fsub_rn(1.0f, b), ic = __fsub_rn(1.0f, c);
    float acc = __fmaf_rn(ic, __fmul_rn(ib, __fmul_rn(ia, v000.x)), __fmul_rn(c, __fmul_rn(ib, __fmul_rn(ia, v001.x))));
    acc = __fmaf_rn(ic, __fmul_rn(b, __fmul_rn(ia, v010.x)), acc);
    acc = __fmaf_rn(c, __fmul_rn(b, __fmul_rn(ia, v011.x)), acc);
    acc = __fmaf_rn(ic, __fmul_rn(ib, __fmul_rn(a, v100.x)), acc);
    acc = __fmaf_rn(c, __fmul_rn(ib, __fmul_rn(a, v101.x)), acc);
    acc = __fmaf_rn(ic, __fmul_rn(b, __fmul_rn(a, v110.x)), acc);
    acc = __fmaf_rn(c, __fmul_rn(b, __fmul_rn(a, v111.x)), acc);
    dist = acc;
    return 1;
}
// tsdfvolume::interpolateColor (src/cuda/tsdfVolume.h:122-148) over the separate colour plane: the same cell, the same
// eight observed-weight tests and the same term order as interpolateSDF, applied to the three colour bytes as floats;
// each channel is truncated (cvt.rzi.u32.f32).  Returns c0 | c1 << 8 | c2 << 16, or HKF_NO_COLOR where the reference
// returns false (whole volumes only).
#define HKF_NO_COLOR 0xffffffffu
__device__ __forceinline__ unsigned vol_interp_color(const VolView& V, const uchar4* __restrict__ color, float px, float py, float pz)
{
    int gx = vol_w2v(V, px), gy = vol_w2v(V, py), gz = vol_w2v(V, pz);
    const int R = V.R;
    if (gx <= 0 || gx >= R - 1) return HKF_NO_COLOR;
    if (gy <= 0 || gy >= R - 1) return HKF_NO_COLOR;
    if (gz <= 0 || gz >= R - 1) return HKF_NO_COLOR;
    const float cell = V.cell;
    const float vx = __fmul_rn(__fadd_rn((float)gx, 0.5f), cell);
    const float vy = __fmul_rn(__fadd_rn((float)gy, 0.5f), cell);
    const float vz = __fmul_rn(__fadd_rn((float)gz, 0.5f), cell);
    gx = (px < vx) ? gx - 1 : gx;
    gy = (py < vy) ? gy - 1 : gy;
    gz = (pz < vz) ? gz - 1 : gz;
    const float a = __fdiv_rn(__fmaf_rn(__fadd_rn((float)gx, 0.5f), -cell, px), cell);
    const float b = __fdiv_rn(__fmaf_rn(__fadd_rn((float)gy, 0.5f), -cell, py), cell);
    const float c = __fdiv_rn(__fmaf_rn(__fadd_rn((float)gz, 0.5f), -cell, pz), cell);
    const size_t i0 = vol_index(V, gx, gy, gz);
    const size_t sy = (size_t)R, sz = (size_t)R * R;
    const size_t off[8] = { 0, sz, sy, sz + sy, 1, sz + 1, sy + 1, sz + sy + 1 };      // 000,001,010,011,100,101,110,111 (x y z)
    uchar4 col[8];
    bool seen = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) { seen = seen && (__ldg(&V.vol[i0 + off[k]].y) != 0.f); col[k] = __ldg(color + i0 + off[k]); }
    if (!seen) return HKF_NO_COLOR;
    const float ia = __fsub_rn(1.0f, a), ib = __fsub_rn(1.0f, b), ic = __fsub_rn(1.0f, c);
    unsigned packed = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (float)(ch == 0 ? col[k].x : ch == 1 ? col[k].y : col[k].z);
        float acc = __fmaf_rn(ic, __fmul_rn(ib, __fmul_rn(ia, v[0])), __fmul_rn(c, __fmul_rn(ib, __fmul_rn(ia, v[1]))));
        acc = __fmaf_rn(ic, __fmul_rn(b, __fmul_rn(ia, v[2])), acc);
        acc = __fmaf_rn(c, __fmul_rn(b, __fmul_rn(ia, v[3])), acc);
        acc = __fmaf_rn(ic, __fmul_rn(ib, __fmul_rn(a, v[4])), acc);
        acc = __fmaf_rn(c, __fmul_rn(ib, __fmul_rn(a, v[5])), acc);
        acc = __fmaf_rn(ic, __fmul_rn(b, __fmul_rn(a, v[6])), acc);
        acc = __fmaf_rn(c, __fmul_rn(b, __fmul_rn(a, v[7])), acc);
        packed |= (__float2uint_rz(acc) & 0xffu) << (8 * ch);
    }
    return packed;
}
#endif
