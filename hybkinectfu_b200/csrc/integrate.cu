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
    DepthBounds B;      // depth extrema of this frame per pixel tile (hkf_internal.cuh)
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

// Conservative voxel interval [x_lo, x_hi] of row (y, z) that can possibly be updated: the row is a straight
// line in camera space, intersected analytically with the view frustum and the frame's depth range.  A superset
// (1 px / 1 voxel margins); the per-voxel tests that decide an update are the reference's.  Returns false when empty.
__device__ __forceinline__ bool row_interval(const IntegrateArgs& A, const float* m, float wy, float wz, float dmax,
                                             int& x_lo, int& x_hi)
{
    const float bx = m[1] * wy + m[2] * wz + m[3];
    const float by = m[5] * wy + m[6] * wz + m[7];
    const float bz = m[9] * wy + m[10] * wz + m[11];
    const float zmax = fminf(dmax, A.maxdist) + A.trunc + 2e-3f;
    const float z_safe = 0.05f;
    // region A: Zc >= z_safe, pixel within [0.5-1, cols-1.5+1] x [0.5-1, rows-1.5+1], Zc <= zmax
    float loA = 0.f, hiA = A.size;
    clip_lin(loA, hiA, m[8], bz - z_safe);
    clip_lin(loA, hiA, -m[8], zmax - bz);
    {
        const float pxmin = -0.5f, pxmax = (float)A.cols - 0.5f, pymin = -0.5f, pymax = (float)A.rows - 0.5f;
        clip_lin(loA, hiA, A.fx * m[0] + (A.cx - pxmin) * m[8], A.fx * bx + (A.cx - pxmin) * bz);
        clip_lin(loA, hiA, -(A.fx * m[0] + (A.cx - pxmax) * m[8]), -(A.fx * bx + (A.cx - pxmax) * bz));
        clip_lin(loA, hiA, A.fy * m[4] + (A.cy - pymin) * m[8], A.fy * by + (A.cy - pymin) * bz);
        clip_lin(loA, hiA, -(A.fy * m[4] + (A.cy - pymax) * m[8]), -(A.fy * by + (A.cy - pymax) * bz));
    }
    // region B: -0.01 <= Zc <= z_safe (projection is ill-conditioned there: no pixel culling)
    float loB = 0.f, hiB = A.size;
    clip_lin(loB, hiB, m[8], bz + 0.01f);
    clip_lin(loB, hiB, -m[8], z_safe - bz);
    float lo, hi;
    const bool hasA = loA <= hiA, hasB = loB <= hiB;
    if (hasA && hasB) { lo = fminf(loA, loB); hi = fmaxf(hiA, hiB); }
    else if (hasA) { lo = loA; hi = hiA; }
    else if (hasB) { lo = loB; hi = hiB; }
    else return false;
    const float rcell = __fdividef(1.0f, A.cell);
    x_lo = max((int)floorf(lo * rcell - 0.5f) - 1, 0);
    x_hi = min((int)ceilf(hi * rcell - 0.5f) + 1, A.R - 1);
    return x_lo <= x_hi;
}

#define INT_NC 4                       // 32-voxel chunks a warp has in flight: a work item is 128 voxels of one row
#define INT_SEG (32 * INT_NC)
#define INT_THREADS 256

__device__ __forceinline__ void load_pose_inv(const IntegrateArgs& A, const HkfDeviceState* st, bool from_state, float* m)
{
    if (from_state) {
#pragma unroll
        for (int i = 0; i < 12; ++i) m[i] = st->pose_inv[i];   // inverse of the tracked pose, written by the tracker
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) m[i] = A.inv[i];
    }
}

// work item: .x = y | zrel << 12 | segment << 24 ; .y = first | last << 8 (voxel offsets inside the 128-voxel
// segment that can be updated) | mask of chunks to project << 16 | mask of free-space chunks << 20
__device__ __forceinline__ uint2 make_item(int y, int zrel, int seg, int lo_rel, int hi_rel, unsigned gen, unsigned fre)
{
    return make_uint2((unsigned)y | ((unsigned)zrel << 12) | ((unsigned)seg << 24),
                      (unsigned)lo_rel | ((unsigned)hi_rel << 8) | (gen << 16) | (fre << 20));
}

// Pass 1 (thread per voxel row, ~2 M warp instructions in total): the row's conservative interval, cut into
// 128-voxel work items and appended to a compact list (one warp-aggregated atomic per 32 rows), so that pass 2
// is perfectly balanced and never looks at the ~60 % of rows that cannot be touched this frame.
template <bool FROM_STATE>
__global__ void __launch_bounds__(256)
k_integrate_items(IntegrateArgs A, HkfDeviceState* __restrict__ st, uint2* __restrict__ items, int parity)
{
    const int lane = threadIdx.x & 31;
    const int R = A.R;
    const int nrows = (A.z_hi - A.z_lo) * R;                  // row = (z - z_lo) * R + y
    const int row = blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->int_items[parity ^ 1] = 0;   // the counter the NEXT call appends to
    if (FROM_STATE && st->track_status != HKF_TRACK_OK) return;   // "camera lost": no integration (HybKinectfu.cpp:123-147)
    float m[12];
    load_pose_inv(A, st, FROM_STATE, m);
    int x_lo = 1, x_hi = 0, y = 0, zrel = 0;
    if (row < nrows) {
        zrel = row / R;
        y = row - zrel * R;
        const float wy = __fmul_rn(__fadd_rn((float)y, 0.5f), A.cell);
        const float wz = __fmul_rn(__fadd_rn((float)(zrel + A.z_lo), 0.5f), A.cell);
        if (!row_interval(A, m, wy, wz, __uint_as_float(A.B.base[0]), x_lo, x_hi)) { x_lo = 1; x_hi = 0; }
    }
    // Chunk classification (32 voxels).  A chunk's voxel centres project inside the bounding box of its two end
    // voxels' projections (a 3D segment projects to a 2D segment while Z > 0); the box, grown by 1.5 px, touches at
    // most 2 tiles per tile row at the level whose tile width covers it.  From the depth extrema of those tiles:
    //   dead : the nearest voxel is farther than trunc behind the largest valid depth -> "sdf > -trunc"
    //          (integrateVolume.cu:64) fails for every voxel, the chunk is dropped;
    //   free : the box lies in the image interior, every pixel of the tiles is valid and < fMaxIntegrateDist, and
    //          the farthest voxel is more than trunc in front of the smallest depth -> every voxel is updated with
    //          tsdf = fminf(sdf/trunc, 1) = 1 exactly, so pass 2 needs no projection and no depth read for it;
    //   else : general, decided voxel by voxel in pass 2.
    unsigned long long gen_chunks = 0, free_chunks = 0;       // bit k: chunk k of this row (R <= 2048 -> <= 64 chunks)
    if (x_lo <= x_hi) {
        const float wy = __fmul_rn(__fadd_rn((float)y, 0.5f), A.cell);
        const float wz = __fmul_rn(__fadd_rn((float)(zrel + A.z_lo), 0.5f), A.cell);
        const float X0 = m[1] * wy + m[2] * wz + m[3], Y0 = m[5] * wy + m[6] * wz + m[7], Z0 = m[9] * wy + m[10] * wz + m[11];
        const float thr_free = A.trunc * 1.000002f + 2e-4f, thr_dead = A.trunc + 2e-4f;
        auto project = [&](int x, float& u, float& v, float& z) {
            const float wx = ((float)x + 0.5f) * A.cell;
            z = m[8] * wx + Z0;
            const float rz = __fdividef(1.0f, z);
            u = A.fx * (m[0] * wx + X0) * rz + A.cx;
            v = A.fy * (m[4] * wx + Y0) * rz + A.cy;
        };
        float ua, va, za;
        project(x_lo, ua, va, za);
        for (int k = x_lo >> 5; k <= (x_hi >> 5); ++k) {
            float ub, vb, zb;
            project(min(x_hi, 32 * k + 31), ub, vb, zb);
            int cls = 1;                                      // 0 dead, 1 general, 2 free
            if (za >= 0.05f && zb >= 0.05f) {
                const float u0 = fminf(ua, ub) - 1.5f, u1 = fmaxf(ua, ub) + 1.5f;
                const float v0 = fminf(va, vb) - 1.5f, v1 = fmaxf(va, vb) + 1.5f;
                const float w = u1 - u0;
                const int lvl = w <= (float)HKF_TILE_W ? 0 : w <= 2.f * HKF_TILE_W ? 1 : w <= 4.f * HKF_TILE_W ? 2 : w <= 8.f * HKF_TILE_W ? 3 : -1;
                const int ty0 = max((int)floorf(v0 * (1.0f / HKF_TILE_H)), 0);
                const int ty1 = min((int)floorf(v1 * (1.0f / HKF_TILE_H)), A.B.tiles_y - 1);
                if (lvl >= 0 && ty1 - ty0 <= 2) {
                    const float inv_w = 1.0f / (float)(HKF_TILE_W << lvl);
                    const int tx0 = max((int)floorf(u0 * inv_w), 0);
                    const int tx1 = min((int)floorf(u1 * inv_w), A.B.tx[lvl] - 1);
                    unsigned mx = 0, mninv = 0;
                    for (int ty = ty0; ty <= ty1; ++ty)
                        for (int tx = tx0; tx <= tx1; ++tx) {
                            mx = max(mx, __ldg(A.B.base + A.B.off_max[lvl] + ty * A.B.tx[lvl] + tx));
                            mninv = max(mninv, __ldg(A.B.base + A.B.off_min[lvl] + ty * A.B.tx[lvl] + tx));
                        }
                    const float dmx = __uint_as_float(mx), dmn = __uint_as_float(~mninv);   // NaN if no tile was read
                    const bool interior = u0 >= 1.0f && u1 <= (float)(A.cols - 2) && v0 >= 1.0f && v1 <= (float)(A.rows - 2);
                    if (fminf(za, zb) >= dmx + thr_dead) cls = 0;
                    else if (interior && dmn > 0.f && dmx < A.maxdist && dmn - fmaxf(za, zb) >= thr_free) cls = 2;
                }
            }
            if (cls == 1) gen_chunks |= 1ull << k;
            if (cls == 2) free_chunks |= 1ull << k;
            ua = ub; va = vb; za = zb;   // the next chunk starts one voxel after this end: [this end, next end] is a superset
        }
    }
    const unsigned long long live_chunks = gen_chunks | free_chunks;
    const int seg0 = x_lo / INT_SEG, seg1 = x_hi / INT_SEG;
    const unsigned segmask = (1u << INT_NC) - 1;
    int n = 0;
    if (x_lo <= x_hi)
        for (int seg = seg0; seg <= seg1; ++seg) n += ((live_chunks >> (seg * INT_NC)) & segmask) ? 1 : 0;
    int incl = n;                                              // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    unsigned base = 0;
    if (lane == 31) base = atomicAdd(&st->int_items[parity], (unsigned)total);
    base = __shfl_sync(0xffffffffu, base, 31) + (unsigned)(incl - n);
    if (n) {
        for (int seg = seg0; seg <= seg1; ++seg) {
            const unsigned g = (unsigned)(gen_chunks >> (seg * INT_NC)) & segmask, f = (unsigned)(free_chunks >> (seg * INT_NC)) & segmask;
            if (g | f) items[base++] = make_item(y, zrel, seg, max(x_lo - seg * INT_SEG, 0), min(x_hi - seg * INT_SEG, INT_SEG - 1), g, f);
        }
    }
}

// Pass 2: persistent warps stride over the work items; a lane owns voxels x = seg*128 + 32 c + lane, c = 0..3, so
// every warp load/store instruction moves 256 contiguous bytes and a warp keeps 4 of them in flight.
template <bool FROM_STATE, bool COUNT_ONLY>
__global__ void __launch_bounds__(INT_THREADS, 5)
k_integrate(IntegrateArgs A, const float* __restrict__ depth, float2* __restrict__ vol, HkfDeviceState* __restrict__ st,
            const uint2* __restrict__ items, int parity)
{
    const int lane = threadIdx.x & 31;
    const int R = A.R;
    const unsigned n_items = st->int_items[parity];
    const unsigned wid = (blockIdx.x * INT_THREADS + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * INT_THREADS) >> 5;
    if (wid >= n_items) return;
    float m[12];
    load_pose_inv(A, st, FROM_STATE, m);
    const float cell = A.cell;
    const int cols = A.cols;
    const unsigned ulim = (unsigned)(A.cols - 2), vlim = (unsigned)(A.rows - 2);
    const float trunc = A.trunc, ntrunc = -A.trunc;
    const float thr_one = A.trunc * 1.000002f;
    const float MAGIC = 12582912.0f;              // 1.5 * 2^23: adding it leaves the nearest integer in the mantissa
    const float edge = 0.5f - 0.00048828125f;     // within 2^-11 px of a rounding boundary -> the exact sequence decides
    unsigned local_count = 0;

    uint2 it = __ldg(items + wid);
#pragma unroll 1
    for (unsigned item = wid; item < n_items; item += nwarps) {
        const uint2 cur = it;
        if (item + nwarps < n_items) it = __ldg(items + item + nwarps);      // prefetch the next item
        const int y = (int)(cur.x & 0xfffu), zrel = (int)((cur.x >> 12) & 0xfffu);
        const int xb = (int)(cur.x >> 24) * INT_SEG;
        const int rel = lane - (int)(cur.y & 0xffu);                          // voxel offset relative to the first live one
        const unsigned span = ((cur.y >> 8) & 0xffu) - (cur.y & 0xffu);
        const unsigned gen = (cur.y >> 16) & 15u, fre = (cur.y >> 20) & 15u;
        const float wy = __fmul_rn(__fadd_rn((float)y, 0.5f), cell);
        const float wz = __fmul_rn(__fadd_rn((float)(zrel + A.z_lo), 0.5f), cell);
        const float ty1 = __fmul_rn(wy, m[1]), ty2 = __fmul_rn(wy, m[5]), ty3 = __fmul_rn(wy, m[9]);
        float2* __restrict__ vp = vol + ((size_t)(zrel * R + y) * R + xb + lane);
        const float xh = (float)(xb + lane) + 0.5f;             // x + 0.5, exact

        float Z[INT_NC];
        int idx[INT_NC];
#pragma unroll
        for (int c = 0; c < INT_NC; ++c) {
            idx[c] = -1; Z[c] = 0.f;
            if (gen & (1u << c)) {                               // warp-uniform
                const float wx = __fmul_rn(xh + 32.0f * c, cell);
                Z[c] = __fadd_rn(__fmaf_rn(wz, m[10], __fmaf_rn(wx, m[8], ty3)), m[11]);
                const float X = __fadd_rn(__fmaf_rn(wz, m[2], __fmaf_rn(wx, m[0], ty1)), m[3]);
                const float Y = __fadd_rn(__fmaf_rn(wz, m[6], __fmaf_rn(wx, m[4], ty2)), m[7]);
                const float ax = __fmul_rn(X, A.fx), ay = __fmul_rn(Y, A.fy);
                bool in = ((unsigned)(rel + 32 * c) <= span) && (Z[c] > 0.f);
                // approximate pixel coordinate (one MUFU.RCP shared by both axes, < 2e-4 px off) and its nearest
                // integer = floor(p + 0.5), without leaving the FMA/ALU pipes
                const float rz = hkf_rcp_approx(Z[c]);
                const float pu = __fmaf_rn(ax, rz, A.cx), pv = __fmaf_rn(ay, rz, A.cy);
                const float tu = __fadd_rn(pu, MAGIC), tv = __fadd_rn(pv, MAGIC);
                int iu = __float_as_int(tu) - 0x4B400000;
                int iv = __float_as_int(tv) - 0x4B400000;
                const float eu = __fsub_rn(pu, __fsub_rn(tu, MAGIC)), ev = __fsub_rn(pv, __fsub_rn(tv, MAGIC));
                if (in && !(fmaxf(fabsf(eu), fabsf(ev)) < edge)) {     // NaN-safe; ~0.2 % of voxels
                    // the reference's exact sequence: mul, div.rn, add, f64 add, cvt.rzi
                    iu = hkf_round_px(hkf_project1(X, A.fx, Z[c], A.cx));
                    iv = hkf_round_px(hkf_project1(Y, A.fy, Z[c], A.cy));
                }
                in = in && ((unsigned)(iu - 1) < ulim) && ((unsigned)(iv - 1) < vlim);
                if (in) idx[c] = iv * cols + iu;
            }
        }
        float d[INT_NC];
#pragma unroll
        for (int c = 0; c < INT_NC; ++c) d[c] = (idx[c] >= 0) ? __ldg(depth + idx[c]) : 0.f;
        float sdf[INT_NC];
        bool live[INT_NC];
#pragma unroll
        for (int c = 0; c < INT_NC; ++c) {
            sdf[c] = __fsub_rn(d[c], Z[c]);
            live[c] = (d[c] != 0.f) && (d[c] < A.maxdist) && (sdf[c] > ntrunc);
            if (fre & (1u << c)) {                               // warp-uniform: free space for certain (pass 1), tsdf = 1
                sdf[c] = 1e30f;
                live[c] = (unsigned)(rel + 32 * c) <= span;
            }
        }
        if (COUNT_ONLY) {
#pragma unroll
            for (int c = 0; c < INT_NC; ++c) local_count += live[c] ? 1u : 0u;
        } else {
            float2 old[INT_NC];
#pragma unroll
            for (int c = 0; c < INT_NC; ++c) {
                old[c] = make_float2(0.f, 0.f);
                if (live[c]) old[c] = vp[32 * c];
            }
#pragma unroll
            for (int c = 0; c < INT_NC; ++c) {
                const float w1 = __fadd_rn(old[c].y, 1.0f);
                const float wn = fminf(w1, A.wmax);
                // Free space seen as free space before: fminf(sdf/trunc, 1) is exactly 1 once sdf >= trunc*(1+2^-19),
                // and fma(1,w,1)/(w+1) is exactly 1, so t' = 1 without a division.  Everything else (the band around
                // the surface, first observations) takes the reference's two IEEE divisions.
                float tn = 1.0f;
                if (live[c] && !(sdf[c] >= thr_one && old[c].x == 1.0f)) {
                    const float tsdf = (sdf[c] >= thr_one) ? 1.0f : fminf(__fdiv_rn(sdf[c], trunc), 1.0f);
                    tn = __fdiv_rn(__fmaf_rn(old[c].x, old[c].y, tsdf), w1);
                }
                // a voxel whose bits do not change (saturated weight, unchanged tsdf) needs no store
                if (live[c] && (tn != old[c].x || wn != old[c].y)) vp[32 * c] = make_float2(tn, wn);
            }
        }
    }
    if (COUNT_ONLY) {
        for (int o = 16; o > 0; o >>= 1) local_count += __shfl_down_sync(0xffffffffu, local_count, o);
        if (lane == 0 && local_count) atomicAdd(&st->upd_count, (unsigned long long)local_count);
    }
}

// both passes, on the context's stream
template <bool FROM_STATE, bool COUNT_ONLY>
static int launch_integrate(hkf_ctx* c, const IntegrateArgs& A)
{
    const int nrows = (A.z_hi - A.z_lo) * A.R;
    const int parity = c->int_parity;
    c->int_parity ^= 1;
    k_integrate_items<FROM_STATE><<<div_up(nrows, 256), 256, 0, c->stream>>>(A, c->st, c->int_items, parity);
    HKF_LAUNCH_CHECK(c);
    k_integrate<FROM_STATE, COUNT_ONLY><<<c->sm_count * 5, INT_THREADS, 0, c->stream>>>(A, c->trunc_depth, c->vol, c->st, c->int_items, parity);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
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
    A->B = c->bounds;
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
    return launch_integrate<false, false>(c, A);
}

extern "C" int hkf_integrate_tracked(hkf_ctx* c, const HkfFrameParams* fp)
{
    if (!c || !fp) return HKF_ERR_ARG;
    IntegrateArgs A; fill_args(c, nullptr, &fp->integrate, &c->ip.depth_camera, &A);
    return launch_integrate<true, false>(c, A);
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
    int rc = launch_integrate<false, true>(c, A);
    if (rc != HKF_OK) return rc;
    HKF_CUDA(c, cudaMemcpyAsync(n_upd, &c->st->upd_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    HKF_CUDA(c, cudaStreamSynchronize(c->stream));
    return HKF_OK;
}
