// integrate.cu — projective TSDF integration as a coalesced streaming pass over z-slabs.
//
// Replaces integrateKernel / cudaIntegrateVolume (src/cuda/integrateVolume.cu:15-97) and
// tsdfvolume::updateVoxel / voxelPosToWorld (src/cuda/tsdfVolume.h:38-74).
//
// Reference shape: one thread per (x,y) column looping serially over z, AoS 12-byte voxels read
// field by field, every voxel of the cube projected every frame, plus a cudaMalloc/cudaFree/D2H
// and an atomicAdd per thread per call.
//
// This design (HBM-bound streaming, no tensor cores):
//  * storage is packed {f32 tsdf, f32 weight} (8 B) in [z][y][x] order, x fastest;
//  * one warp owns one (y,z) row; a lane owns voxel x = chunk + lane, so a warp moves 256 contiguous
//    bytes per load/store instruction and rows of a block are adjacent in y (overlapping depth footprints);
//  * before touching memory the warp intersects its row (a straight line in camera space) with the view
//    frustum and the frame's depth range analytically, so voxels that cannot be updated are neither
//    projected nor read.  The interval is a conservative superset (1 px / 1 voxel margins); the per-voxel
//    tests that decide an update are the reference's, so the update mask is bit-identical;
//  * the first ncu capture (profiles/r1_integrate_a.txt) showed the straightforward kernel to be
//    ISSUE-bound (74 % issue-active, 21 % DRAM), not HBM-bound: ~170 instructions per voxel, dominated by
//    three IEEE divisions, f64 rounding and int<->float conversions on the quarter-rate XU pipe.  The
//    inner loop therefore decides everything it can from cheap approximations and only falls back to the
//    reference's exact sequence where the approximation cannot prove the result:
//      - pixel index: one MUFU.RCP shared by both axes gives the pixel coordinate to < 2e-4 px; unless it
//        lies within 2^-9 px of a rounding boundary the reference's integer is already decided
//        (float->int through the 1.5*2^23 magic constant: FMA-pipe only, no XU conversions);
//      - tsdf = fminf(sdf/trunc, 1) is exactly 1 once sdf >= trunc*(1+2^-19) (free space): no division;
//      - t' = fma(t,w,tsdf)/(w+1) is exactly 1 when t == tsdf == 1: no division;
//      - a voxel whose bits do not change (saturated weight, unchanged tsdf) is not stored.
//    Every voxel still ends with exactly the reference's bits (tests/test_gpu_stages.py).
//  * the depth map (1.2 MB) is read through the read-only path and stays L1/L2-resident;
//  * DRAM traffic ~ 16 B x N_upd.
//
// Arithmetic contract (SASS of the reference built for sm_100, see oracle/Makefile):
//   cell = S / (float)R                     (div.rn, per axis)
//   w?   = ((float)i + 0.5f) * cell         (add, mul)
//   X    = fma(wz, m13, fma(wx, m11, wy*m12)) + m14     (likewise Y, Z with rows 2, 3 of T^-1)
//   skip if Z <= 0
//   u    = (int)trunc((double)((X*fx)/Z + cx) + 0.5)    (mul, div.rn, add, f64 add, cvt.rzi)
//   skip unless 1 <= u < cols-1 and 1 <= v < rows-1
//   d    = trunc_depth[v][u]; skip if d == 0 or !(d < fMaxIntegrateDist)
//   sdf  = d - Z; skip unless sdf > -trunc
//   tsdf = fminf(sdf / trunc, 1)            (div.rn)
//   t'   = fma(t, w, tsdf) / (w + 1);  w' = fminf(w + 1, wmax)
#include "hkf_internal.cuh"
#include "hkf_hostdev_math.h"

struct IntegrateArgs {
    float inv[12];      // rows 1..3 of the inverse pose (world -> camera)
    float cell;
    float fx, fy, cx, cy;
    int cols, rows;
    float trunc, maxdist, wmax;
    int R;
    int z_lo, z_hi;     // stored planes of this slab
    float size;         // S
};

// clip the world-x interval [lo,hi] with a*wx + b >= 0 (approximate division: the caller adds margins)
__device__ __forceinline__ void clip_lin(float& lo, float& hi, float a, float b)
{
    const float eps = 1e-12f;
    const float q = __fdividef(-b, a);
    if (a > eps) lo = fmaxf(lo, q);
    else if (a < -eps) hi = fminf(hi, q);
    else if (b < 0.f) { lo = 1.f; hi = 0.f; }
}

template <bool FROM_STATE, bool COUNT_ONLY>
__global__ void __launch_bounds__(256)
k_integrate(IntegrateArgs A, const float* __restrict__ depth, float2* __restrict__ vol, HkfDeviceState* __restrict__ st)
{
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int R = A.R;
    const int row = blockIdx.x * 8 + warp_in_block;          // row = (z - z_lo) * R + y
    const int nrows = (A.z_hi - A.z_lo) * R;
    if (row >= nrows) return;

    float m11, m12, m13, m14, m21, m22, m23, m24, m31, m32, m33, m34;
    if (FROM_STATE) {
        if (st->track_status != HKF_TRACK_OK) return;        // "camera lost": no integration (HybKinectfu.cpp:123-147)
        const float* p = st->pose_inv;                        // inverse of the tracked pose, written by the tracker
        m11 = p[0]; m12 = p[1]; m13 = p[2]; m14 = p[3];
        m21 = p[4]; m22 = p[5]; m23 = p[6]; m24 = p[7];
        m31 = p[8]; m32 = p[9]; m33 = p[10]; m34 = p[11];
    } else {
        m11 = A.inv[0]; m12 = A.inv[1]; m13 = A.inv[2]; m14 = A.inv[3];
        m21 = A.inv[4]; m22 = A.inv[5]; m23 = A.inv[6]; m24 = A.inv[7];
        m31 = A.inv[8]; m32 = A.inv[9]; m33 = A.inv[10]; m34 = A.inv[11];
    }

    const int zrel = row / R;
    const int y = row - zrel * R;
    const int z = zrel + A.z_lo;
    const float cell = A.cell;
    const float wy = __fmul_rn(__fadd_rn((float)y, 0.5f), cell);
    const float wz = __fmul_rn(__fadd_rn((float)z, 0.5f), cell);
    const float ty1 = __fmul_rn(wy, m12), ty2 = __fmul_rn(wy, m22), ty3 = __fmul_rn(wy, m32);

    // ---- conservative x-interval of this row (analytic, lane-uniform) ----
    const float bx = m12 * wy + m13 * wz + m14;
    const float by = m22 * wy + m23 * wz + m24;
    const float bz = m32 * wy + m33 * wz + m34;
    const float dmax = __uint_as_float(st->dmax_bits);
    const float zmax = fminf(dmax, A.maxdist) + A.trunc + 2e-3f;
    const float z_safe = 0.05f;
    // region A: Zc >= z_safe, pixel within [0.5-1, cols-1.5+1] x [0.5-1, rows-1.5+1], Zc <= zmax
    float loA = 0.f, hiA = A.size;
    clip_lin(loA, hiA, m31, bz - z_safe);
    clip_lin(loA, hiA, -m31, zmax - bz);
    {
        const float pxmin = -0.5f, pxmax = (float)A.cols - 0.5f, pymin = -0.5f, pymax = (float)A.rows - 0.5f;
        clip_lin(loA, hiA, A.fx * m11 + (A.cx - pxmin) * m31, A.fx * bx + (A.cx - pxmin) * bz);
        clip_lin(loA, hiA, -(A.fx * m11 + (A.cx - pxmax) * m31), -(A.fx * bx + (A.cx - pxmax) * bz));
        clip_lin(loA, hiA, A.fy * m21 + (A.cy - pymin) * m31, A.fy * by + (A.cy - pymin) * bz);
        clip_lin(loA, hiA, -(A.fy * m21 + (A.cy - pymax) * m31), -(A.fy * by + (A.cy - pymax) * bz));
    }
    // region B: -0.01 <= Zc <= z_safe (projection is ill-conditioned there: no pixel culling)
    float loB = 0.f, hiB = A.size;
    clip_lin(loB, hiB, m31, bz + 0.01f);
    clip_lin(loB, hiB, -m31, z_safe - bz);
    float lo, hi;
    const bool hasA = loA <= hiA, hasB = loB <= hiB;
    if (hasA && hasB) { lo = fminf(loA, loB); hi = fmaxf(hiA, hiB); }
    else if (hasA) { lo = loA; hi = hiA; }
    else if (hasB) { lo = loB; hi = hiB; }
    else return;
    const float rcell = __fdividef(1.0f, cell);
    int x_lo = (int)floorf(lo * rcell - 0.5f) - 1;
    int x_hi = (int)ceilf(hi * rcell - 0.5f) + 1;
    x_lo = max(x_lo, 0); x_hi = min(x_hi, R - 1);
    if (x_lo > x_hi) return;

    float2* rowp = vol + (size_t)row * R;
    const int cols = A.cols, rows = A.rows;
    const unsigned ulim = (unsigned)(cols - 2), vlim = (unsigned)(rows - 2);
    const float ntrunc = -A.trunc;
    const float thr_one = A.trunc * 1.000002f;
    const float cxh = A.cx + 0.5f, cyh = A.cy + 0.5f;
    const float MAGIC = 12582912.0f;             // 1.5 * 2^23: float add rounds to the nearest integer, kept in the mantissa
    const float eps = 0.001953125f;              // 2^-9 px
    unsigned local_count = 0;

    const int xs = (x_lo & ~31) + lane;
    float xf = (float)xs;
#pragma unroll 2
    for (int x = xs; x <= x_hi; x += 32, xf += 32.0f) {
        const float wx = __fmul_rn(__fadd_rn(xf, 0.5f), cell);
        const float Z = __fadd_rn(__fmaf_rn(wz, m33, __fmaf_rn(wx, m31, ty3)), m34);
        const float X = __fadd_rn(__fmaf_rn(wz, m13, __fmaf_rn(wx, m11, ty1)), m14);
        const float Y = __fadd_rn(__fmaf_rn(wz, m23, __fmaf_rn(wx, m21, ty2)), m24);
        const float ax = __fmul_rn(X, A.fx), ay = __fmul_rn(Y, A.fy);
        bool live = (x < R) && (Z > 0.f);
        // approximate pixel coordinate + 0.5, its nearest integer and floor without leaving the FMA/ALU pipes
        const float rz = hkf_rcp_approx(Z);
        const float fu = __fmaf_rn(ax, rz, cxh), fv = __fmaf_rn(ay, rz, cyh);
        const float tu = __fadd_rn(fu, MAGIC), tv = __fadd_rn(fv, MAGIC);
        const float du = __fsub_rn(fu, __fsub_rn(tu, MAGIC)), dv = __fsub_rn(fv, __fsub_rn(tv, MAGIC));
        int u = __float_as_int(tu) - 0x4B400000 - (du < 0.f ? 1 : 0);
        int v = __float_as_int(tv) - 0x4B400000 - (dv < 0.f ? 1 : 0);
        if (live && (fabsf(du) < eps || fabsf(dv) < eps || !(Z >= 0.01f))) {
            // within 2^-9 px of a rounding boundary (or ill-conditioned): the reference's exact sequence decides
            u = hkf_round_px(hkf_project1(X, A.fx, Z, A.cx));
            v = hkf_round_px(hkf_project1(Y, A.fy, Z, A.cy));
        }
        live = live && ((unsigned)(u - 1) < ulim) && ((unsigned)(v - 1) < vlim);
        float d = 0.f;
        if (live) d = __ldg(depth + v * cols + u);
        const float sdf = __fsub_rn(d, Z);
        live = live && (d != 0.f) && (d < A.maxdist) && (sdf > ntrunc);
        if (COUNT_ONLY) {
            local_count += live ? 1u : 0u;
        } else if (live) {
            const float2 old = rowp[x];
            // fminf(sdf/trunc, 1): the quotient is >= 1 for sure once sdf >= trunc*(1+2^-19): free space, no division
            const float tsdf = (sdf >= thr_one) ? 1.0f : fminf(__fdiv_rn(sdf, A.trunc), 1.0f);
            const float w1 = __fadd_rn(old.y, 1.0f);
            // fma(1,w,1) == w+1 exactly, so the quotient is exactly 1 when t == tsdf == 1
            const float tn = (old.x == 1.0f && tsdf == 1.0f) ? 1.0f : __fdiv_rn(__fmaf_rn(old.x, old.y, tsdf), w1);
            const float wn = fminf(w1, A.wmax);
            // a voxel whose bits do not change (saturated weight, unchanged tsdf) needs no store
            if (tn != old.x || wn != old.y) rowp[x] = make_float2(tn, wn);
        }
    }
    if (COUNT_ONLY) {
        for (int o = 16; o > 0; o >>= 1) local_count += __shfl_down_sync(0xffffffffu, local_count, o);
        if (lane == 0 && local_count) atomicAdd(&st->upd_count, (unsigned long long)local_count);
    }
}

static void fill_args(hkf_ctx* c, const float* inv16, const HkfIntegrateParams* ip, const HkfCameraParams* cam, IntegrateArgs* A)
{
    if (inv16) for (int i = 0; i < 12; ++i) A->inv[i] = inv16[i];
    else for (int i = 0; i < 12; ++i) A->inv[i] = 0.f;
    // voxelPosToWorld: _size.x / _resolution.x  (float / int -> float division)
    A->cell = c->ip.volume_size / (float)(int)c->ip.volume_resolution;
    A->fx = cam->fx; A->fy = cam->fy; A->cx = cam->cx; A->cy = cam->cy;
    A->cols = (int)cam->cols; A->rows = (int)cam->rows;
    A->trunc = ip->fSdfTruncation; A->maxdist = ip->fMaxIntegrateDist; A->wmax = c->ip.weight_max;
    A->R = (int)c->ip.volume_resolution;
    A->z_lo = (int)c->z_lo; A->z_hi = (int)c->z_hi;
    A->size = c->ip.volume_size;
}

extern "C" int hkf_integrate_volume(hkf_ctx* c, int has_color, int use_angle_weight_color, const HkfMat44* transform,
                                    const HkfIntegrateParams* ip, const HkfCameraParams* cam, const HkfCameraParams* rgb_cam)
{
    (void)rgb_cam;
    if (!c || !transform || !ip || !cam) return HKF_ERR_ARG;
    if (has_color || use_angle_weight_color) { c->err = "colour integration is out of scope (use_color=0)"; return HKF_ERR_ARG; }
    if (cam->cols != c->cols[0] || cam->rows != c->rows[0]) { c->err = "camera size mismatch"; return HKF_ERR_ARG; }
    float inv[16];
    hkf_mat44_inverse(transform->m, inv);      // host, float cofactors: cudaIntegrateVolume line 84
    IntegrateArgs A; fill_args(c, inv, ip, cam, &A);
    int nrows = (A.z_hi - A.z_lo) * A.R;
    k_integrate<false, false><<<div_up(nrows, 8), 256, 0, c->stream>>>(A, c->trunc_depth, c->vol, c->st);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

extern "C" int hkf_integrate_tracked(hkf_ctx* c, const HkfFrameParams* fp)
{
    if (!c || !fp) return HKF_ERR_ARG;
    IntegrateArgs A; fill_args(c, nullptr, &fp->integrate, &c->ip.depth_camera, &A);
    int nrows = (A.z_hi - A.z_lo) * A.R;
    k_integrate<true, false><<<div_up(nrows, 8), 256, 0, c->stream>>>(A, c->trunc_depth, c->vol, c->st);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

extern "C" int hkf_count_integrate_updates(hkf_ctx* c, const HkfMat44* transform, const HkfIntegrateParams* ip,
                                           const HkfCameraParams* cam, unsigned long long* n_upd)
{
    if (!c || !transform || !ip || !cam || !n_upd) return HKF_ERR_ARG;
    float inv[16];
    hkf_mat44_inverse(transform->m, inv);
    IntegrateArgs A; fill_args(c, inv, ip, cam, &A);
    // count over OWNED planes only
    A.z_lo = (int)c->own_z0; A.z_hi = (int)c->own_z1;
    HKF_CUDA(c, cudaMemsetAsync(&c->st->upd_count, 0, sizeof(unsigned long long), c->stream));
    int nrows = (A.z_hi - A.z_lo) * A.R;
    k_integrate<false, true><<<div_up(nrows, 8), 256, 0, c->stream>>>(A, c->trunc_depth, c->vol, c->st);
    HKF_LAUNCH_CHECK(c);
    HKF_CUDA(c, cudaMemcpyAsync(n_upd, &c->st->upd_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    HKF_CUDA(c, cudaStreamSynchronize(c->stream));
    return HKF_OK;
}
