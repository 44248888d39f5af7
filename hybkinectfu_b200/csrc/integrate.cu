// integrate.cu — projective TSDF integration as a coalesced streaming pass over z-slabs.
//
// Replaces integrateKernel / cudaIntegrateVolume (src/cuda/integrateVolume.cu:15-97) and
// tsdfvolume::updateVoxel / voxelPosToWorld (src/cuda/tsdfVolume.h:38-74).
//
// Reference shape: one thread per (x,y) column looping serially over z, AoS 12-byte voxels read
// field by field, every voxel of the cube projected every frame, plus a cudaMalloc/cudaFree/D2H
// and an atomicAdd per thread per call.
//
// This design (HBM-bound by construction, no tensor cores; the history of measurements is in DESIGN.md 3.4 and
// profiles/r1_integrate_*, profiles/r2_integrate_*):
//  * storage is packed {f32 tsdf, f32 weight} (8 B) in [z][y][x] order, x fastest; a lane owns voxel x = chunk + lane,
//    so every warp load/store moves 256 contiguous bytes;
//  * pass 1 (k_integrate_items) intersects every group of 4 voxel rows (straight lines in camera space) with the view
//    frustum and the frame's depth range analytically, then classifies each 32-voxel chunk of the interval - one thread
//    per chunk - against per-tile depth extrema published by the preprocess: dead (behind every visible surface), free
//    (in front of every visible surface, all touched pixels integrable: tsdf = 1 exactly) or general, and writes two
//    compact work lists;
//  * pass 2 (k_integrate): persistent warps stride over BOTH lists at once.  The two kinds of work bound differently:
//    free bricks are a pure streaming read-modify-write (no projection, no depth read: HBM-bound), general chunks
//    (4 rows x 32 voxels, every voxel projected) are issue-bound (~60 instructions per 32 voxels).  Round 1 ran them one
//    after the other (52 % DRAM throughput: the general phase starved the memory system, the free phase the issue
//    slots).  Now every warp keeps the rows of its next free bricks in flight as bulk asynchronous copies
//    (cp.async.bulk global -> shared memory, completion on an mbarrier: one elected lane, 4 instructions per 4 KB brick,
//    no registers held) WHILE it projects general chunks, and turns to a brick when its bytes have landed.  The ncu
//    A/B of the two designs is profiles/r2_integrate_*.txt.
//  * the projection decides everything it can from cheap approximations and only falls back to the reference's exact
//    sequence where they cannot prove the result:
//      - pixel index: one MUFU.RCP shared by both axes gives the pixel coordinate to < 2e-4 px; unless it lies within
//        2^-11 px of a rounding boundary the reference's integer is already decided (float->int through the 1.5*2^23
//        magic constant: FMA-pipe only, no XU conversions);
//      - tsdf = fminf(sdf/trunc, 1) is exactly 1 once sdf >= trunc*(1+2^-19) (free space): no division;
//      - t' = fma(t,w,tsdf)/(w+1) is exactly 1 when t == tsdf == 1: no division;
//      - a voxel whose bits do not change (saturated weight, unchanged tsdf) is not stored.
//    The interval and the classification are conservative supersets; the per-voxel tests that decide an update are
//    the reference's, so every voxel ends with exactly the reference's bits (tests/test_gpu_stages.py);
//  * the depth map (1.2 MB) is read through the read-only path and stays L1/L2-resident;
//  * DRAM traffic ~ 16 B x N_upd (measured: profiles/integrate_traffic.json).
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

// colour integration (has_color): the frame's colour image, the new frame's level-0 normals (angle weighting), the
// voxel colour plane
struct IntegrateColor {
    uchar4* color_vol;
    const unsigned char* rgb;       // uchar3 [rgb_rows][rgb_cols]
    const float4* normals;
    int rgb_cols, rgb_rows;
    int angled;
};

#define INT_G 4                        // voxel rows (consecutive y, same z) per work item
#define INT_NC 4                       // 32-voxel chunks per row of a work item: an item is a 4 x 128 voxel brick
#define INT_SEG (32 * INT_NC)
#define INT_THREADS 256

// clip the world-x interval [lo,hi] with a*wx + b(row) >= 0 for the first and last row of a group: b is linear in
// the row index, so the bound of every row in between lies between q0 and q1 (approximate divisions: the caller
// adds margins)
__device__ __forceinline__ void clip_lin2(float& lo, float& hi, float a, float b0, float b1)
{
    const float eps = 1e-12f;
    const float q0 = __fdividef(-b0, a), q1 = __fdividef(-b1, a);
    if (a > eps) lo = fmaxf(lo, fminf(q0, q1));
    else if (a < -eps) hi = fminf(hi, fmaxf(q0, q1));
    else if (b0 < 0.f && b1 < 0.f) { lo = 1.f; hi = 0.f; }
}

// Conservative voxel interval [x_lo, x_hi] that can possibly be updated in ANY of the rows y0..y1 of plane z: a
// row is a straight line in camera space, intersected analytically with the view frustum and the frame's depth
// range.  A superset (1 px / 1 voxel margins); the per-voxel tests that decide an update are the reference's.
__device__ __forceinline__ bool group_interval(const IntegrateArgs& A, const float* m, float wy0, float wy1, float wz, float dmax,
                                               int& x_lo, int& x_hi)
{
    const float bx0 = m[1] * wy0 + m[2] * wz + m[3], bx1 = m[1] * wy1 + m[2] * wz + m[3];
    const float by0 = m[5] * wy0 + m[6] * wz + m[7], by1 = m[5] * wy1 + m[6] * wz + m[7];
    const float bz0 = m[9] * wy0 + m[10] * wz + m[11], bz1 = m[9] * wy1 + m[10] * wz + m[11];
    const float zmax = fminf(dmax, A.maxdist) + A.trunc + 2e-3f;
    const float z_safe = 0.05f;
    // region A: Zc >= z_safe, pixel within [0.5-1, cols-1.5+1] x [0.5-1, rows-1.5+1], Zc <= zmax
    float loA = 0.f, hiA = A.size;
    clip_lin2(loA, hiA, m[8], bz0 - z_safe, bz1 - z_safe);
    clip_lin2(loA, hiA, -m[8], zmax - bz0, zmax - bz1);
    {
        const float pxmin = -0.5f, pxmax = (float)A.cols - 0.5f, pymin = -0.5f, pymax = (float)A.rows - 0.5f;
        const float c0 = A.cx - pxmin, c1 = A.cx - pxmax, c2 = A.cy - pymin, c3 = A.cy - pymax;
        clip_lin2(loA, hiA, A.fx * m[0] + c0 * m[8], A.fx * bx0 + c0 * bz0, A.fx * bx1 + c0 * bz1);
        clip_lin2(loA, hiA, -(A.fx * m[0] + c1 * m[8]), -(A.fx * bx0 + c1 * bz0), -(A.fx * bx1 + c1 * bz1));
        clip_lin2(loA, hiA, A.fy * m[4] + c2 * m[8], A.fy * by0 + c2 * bz0, A.fy * by1 + c2 * bz1);
        clip_lin2(loA, hiA, -(A.fy * m[4] + c3 * m[8]), -(A.fy * by0 + c3 * bz0), -(A.fy * by1 + c3 * bz1));
    }
    // region B: -0.01 <= Zc <= z_safe (projection is ill-conditioned there: no pixel culling)
    float loB = 0.f, hiB = A.size;
    clip_lin2(loB, hiB, m[8], bz0 + 0.01f, bz1 + 0.01f);
    clip_lin2(loB, hiB, -m[8], z_safe - bz0, z_safe - bz1);
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

// Work items (16 bytes; everything pass 2 would otherwise have to derive per item is precomputed once here):
//   .x  voxel index / 4 of the item's first voxel in the stored slab (row y0 of plane zrel, x = the brick's first voxel
//       for a free item, the chunk's first voxel for a general item); rows are R voxels apart
//   .y  y0/4 | zrel << 10 | (first voxel x / 32) << 22           (general items: the voxel coordinates to project)
//   .z  first | last << 8: offsets that can be updated, relative to the item's first voxel (0..127 / 0..31)
//   .w  free items: mask of free chunks | first copied voxel offset << 8 | bytes per row of the bulk copy << 16
__device__ __forceinline__ uint4 make_free_item(const IntegrateArgs& A, int yg, int zrel, int seg, int lo_rel, int hi_rel, unsigned fre)
{
    const unsigned xb = (unsigned)seg * INT_SEG;
    const unsigned cf = (unsigned)__ffs((int)fre) - 1u, cl = (fre & 8u) ? 3u : (fre & 4u) ? 2u : (fre & 2u) ? 1u : 0u;
    const unsigned x1 = min(xb + 32u * cl + 32u, (unsigned)A.R);           // R is a multiple of 4: 32-byte granules
    const unsigned vox = ((unsigned)zrel * (unsigned)A.R + (unsigned)yg * INT_G) * (unsigned)(A.R / 4) + xb / 4;
    return make_uint4(vox, (unsigned)yg | ((unsigned)zrel << 10) | ((xb / 32u) << 22), (unsigned)lo_rel | ((unsigned)hi_rel << 8),
                      fre | ((32u * cf) << 8) | (((x1 - xb - 32u * cf) * 8u) << 16));
}
__device__ __forceinline__ uint4 make_general_item(const IntegrateArgs& A, int yg, int zrel, int seg, int lo_rel, int hi_rel, int c)
{
    const unsigned xc = (unsigned)seg * INT_SEG + 32u * (unsigned)c;
    const unsigned vox = ((unsigned)zrel * (unsigned)A.R + (unsigned)yg * INT_G) * (unsigned)(A.R / 4) + xc / 4;
    const int lo = max(lo_rel - 32 * c, 0), hi = min(hi_rel - 32 * c, 31);
    return make_uint4(vox, (unsigned)yg | ((unsigned)zrel << 10) | ((xc / 32u) << 22), (unsigned)lo | ((unsigned)hi << 8), 0u);
}

// Pass 1 (thread per group of 4 voxel rows, then thread per (group, 128-voxel segment)): the group's conservative
// interval, cut into 32-voxel chunks that are classified against the frame's per-tile depth extrema, and the surviving
// chunks appended to two compact work lists (one warp-aggregated atomic per list per warp): the free-space chunks of a
// 4 x 128 voxel brick form ONE item growing from the front of `items`, every general chunk is an item of its own growing
// from the back.  (A thread per CHUNK was measured too: four times the parallelism, but every thread repeats the
// per-brick set-up - 12.6 M instead of 8.4 M warp instructions, 25.7 instead of 20.3 us, profiles/r2_integrate_a_*.)
//
// Chunk classification.  The voxel centres of a 4-row x 32-voxel chunk form a planar rectangle; while Z > 0 its
// image is the convex hull of its four projected corners, and depth along it is extremal at a corner.  The
// bounding box of the corners, grown by 1.5 px, is looked up in the range-query table of 8 x 8 pixel tiles
// (hkf_internal.cuh: DepthBounds; boxes of more than 256 x 64 pixels stay general).  From the depth extrema of the tiles
// it touches:
//   dead : the nearest corner is farther than trunc behind the largest valid depth -> "sdf > -trunc"
//          (integrateVolume.cu:64) fails for every voxel, the chunk is dropped;
//   free : the box lies in the image interior, every pixel of the tiles is valid and < fMaxIntegrateDist, and the
//          farthest corner is more than trunc in front of the smallest depth -> every voxel is updated with
//          tsdf = fminf(sdf/trunc, 1) = 1 exactly, so pass 2 needs neither a projection nor a depth read for it;
//   else : general, decided voxel by voxel in pass 2 with the reference's arithmetic.
#define INT_GB 64                      // row groups per block of pass 1

// NOFREE (colour integration): every updated voxel needs its colour pixel, so the projection-free class is not used.
// PROFILE (hkf_integrate_profile): per-plane counts of general chunks / free bricks into hist, for slab balancing.
template <bool FROM_STATE, bool NOFREE, bool PROFILE>
__global__ void __launch_bounds__(256, 5)
k_integrate_items(IntegrateArgs A, HkfDeviceState* __restrict__ st, uint4* __restrict__ items, unsigned capacity, int parity, unsigned* __restrict__ hist)
{
    __shared__ int s_lo[INT_GB], s_hi[INT_GB];
    __shared__ int s_any;
    hkf_pdl_enter();
    const int lane = threadIdx.x & 31;
    const int RG = A.R / INT_G;
    const int ngroups = (A.z_hi - A.z_lo) * RG;
    const int nseg = (A.R + INT_SEG - 1) / INT_SEG;
    if (blockIdx.x == 0 && threadIdx.x < 2) st->int_items[(parity ^ 1) * 2 + threadIdx.x] = 0;   // the counters the NEXT call appends to
    if (FROM_STATE && st->track_status != HKF_TRACK_OK) return;   // "camera lost": no integration (HybKinectfu.cpp:123-147)
    float m[12];
    load_pose_inv(A, st, FROM_STATE, m);
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();

    // ---- phase A: one thread per row group computes the group's interval
    if (threadIdx.x < INT_GB) {
        const int g = blockIdx.x * INT_GB + threadIdx.x;
        int x_lo = 1, x_hi = 0;
        if (g < ngroups) {
            const int zrel = g / RG, yg = g - zrel * RG;
            const float wy0 = ((float)(yg * INT_G) + 0.5f) * A.cell;
            const float wy1 = ((float)(yg * INT_G + INT_G - 1) + 0.5f) * A.cell;
            const float wz = ((float)(zrel + A.z_lo) + 0.5f) * A.cell;
            if (!group_interval(A, m, wy0, wy1, wz, __uint_as_float(A.B.base[0]), x_lo, x_hi)) { x_lo = 1; x_hi = 0; }
        }
        s_lo[threadIdx.x] = x_lo; s_hi[threadIdx.x] = x_hi;
        if (x_lo <= x_hi) s_any = 1;
    }
    __syncthreads();
    if (!s_any) return;                                            // nothing of this block's rows can be updated

    // ---- phase B: one thread per (row group, 128-voxel segment) classifies the segment's four chunks
    const float thr_free = A.trunc * 1.000002f + 2e-4f, thr_dead = A.trunc + 2e-4f;
    for (int p = threadIdx.x; p < INT_GB * nseg; p += 256) {       // uniform trip count
        const int gl = p / nseg, seg = p - gl * nseg;
        const int g = blockIdx.x * INT_GB + gl;
        const int x_lo = s_lo[gl], x_hi = s_hi[gl];
        const int xa = max(x_lo, seg * INT_SEG), xe = min(x_hi, seg * INT_SEG + INT_SEG - 1);
        unsigned gm = 0, fm = 0;
        int zrel = 0, yg = 0;
        if (xa <= xe) {
            zrel = g / RG; yg = g - zrel * RG;
            const float wy0 = ((float)(yg * INT_G) + 0.5f) * A.cell;
            const float wy1 = ((float)(yg * INT_G + INT_G - 1) + 0.5f) * A.cell;
            const float wz = ((float)(zrel + A.z_lo) + 0.5f) * A.cell;
            const float Xa = m[1] * wy0 + m[2] * wz + m[3], Ya = m[5] * wy0 + m[6] * wz + m[7], Za = m[9] * wy0 + m[10] * wz + m[11];
            const float Xb = m[1] * wy1 + m[2] * wz + m[3], Yb = m[5] * wy1 + m[6] * wz + m[7], Zb = m[9] * wy1 + m[10] * wz + m[11];
            auto corners = [&](int x, float& u0, float& u1, float& v0, float& v1, float& z0, float& z1) {
                const float wx = ((float)x + 0.5f) * A.cell;
                const float za = m[8] * wx + Za, zb = m[8] * wx + Zb;
                const float ra = __fdividef(1.0f, za), rb = __fdividef(1.0f, zb);
                const float ua = A.fx * (m[0] * wx + Xa) * ra + A.cx, ub = A.fx * (m[0] * wx + Xb) * rb + A.cx;
                const float va = A.fy * (m[4] * wx + Ya) * ra + A.cy, vb = A.fy * (m[4] * wx + Yb) * rb + A.cy;
                u0 = fminf(ua, ub); u1 = fmaxf(ua, ub); v0 = fminf(va, vb); v1 = fmaxf(va, vb); z0 = fminf(za, zb); z1 = fmaxf(za, zb);
            };
            float pu0, pu1, pv0, pv1, pz0, pz1;                    // box of the two corners at the current chunk boundary
            corners(xa, pu0, pu1, pv0, pv1, pz0, pz1);
#pragma unroll 1
            for (int c = (xa >> 5) & (INT_NC - 1); c <= ((xe >> 5) & (INT_NC - 1)); ++c) {
                float qu0, qu1, qv0, qv1, qz0, qz1;
                corners(min(xe, seg * INT_SEG + 32 * c + 31), qu0, qu1, qv0, qv1, qz0, qz1);
                int cls = 1;                                      // 0 dead, 1 general, 2 free
                const float zn = fminf(pz0, qz0), zf = fmaxf(pz1, qz1);
                if (zn >= 0.05f) {
                    const float u0 = fminf(pu0, qu0) - 1.5f, u1 = fmaxf(pu1, qu1) + 1.5f;
                    const float v0 = fminf(pv0, qv0) - 1.5f, v1 = fmaxf(pv1, qv1) + 1.5f;
                    // extrema over the base tiles the box touches: four overlapping power-of-two windows of the table
                    const int tx0 = max((int)floorf(u0 * (1.0f / HKF_BT)), 0), tx1 = min((int)floorf(u1 * (1.0f / HKF_BT)), A.B.tx - 1);
                    const int ty0 = max((int)floorf(v0 * (1.0f / HKF_BT)), 0), ty1 = min((int)floorf(v1 * (1.0f / HKF_BT)), A.B.ty - 1);
                    const int w = tx1 - tx0 + 1, h = ty1 - ty0 + 1;
                    if (w >= 1 && h >= 1 && w <= (1 << (HKF_BKX - 1)) && h <= (1 << (HKF_BKY - 1))) {
                        const int kx = 31 - __clz(w), ky = 31 - __clz(h);
                        const uint2* t = A.B.tab + (size_t)(kx * HKF_BKY + ky) * ((size_t)A.B.tx * A.B.ty);
                        const int xb = tx1 - (1 << kx) + 1, yb = ty1 - (1 << ky) + 1;
                        const uint2 e0 = __ldg(t + (size_t)ty0 * A.B.tx + tx0), e1 = __ldg(t + (size_t)ty0 * A.B.tx + xb);
                        const uint2 e2 = __ldg(t + (size_t)yb * A.B.tx + tx0), e3 = __ldg(t + (size_t)yb * A.B.tx + xb);
                        const unsigned mx = max(max(e0.x, e1.x), max(e2.x, e3.x)), mninv = max(max(e0.y, e1.y), max(e2.y, e3.y));
                        const float dmx = __uint_as_float(mx), dmn = __uint_as_float(~mninv);
                        const bool interior = u0 >= 1.0f && u1 <= (float)(A.cols - 2) && v0 >= 1.0f && v1 <= (float)(A.rows - 2);
                        if (zn >= dmx + thr_dead || dmn >= A.maxdist) cls = 0;     // behind everything / nothing integrable there
                        else if (!NOFREE && interior && dmn > 0.f && dmx < A.maxdist && dmn - zf >= thr_free) cls = 2;
                    }
                }
                if (cls == 1) gm |= 1u << c;
                if (cls == 2) fm |= 1u << c;
                // the next chunk starts one voxel after this end: [this end, next end] is a superset
                pu0 = qu0; pu1 = qu1; pv0 = qv0; pv1 = qv1; pz0 = qz0; pz1 = qz1;
            }
        }
        const unsigned ng = (unsigned)__popc(gm);
        const bool isf = fm != 0;
        if (PROFILE) {                        // slab balancing: where the work of this view is, plane by plane
            if (ng) atomicAdd(&hist[2 * zrel], ng);
            if (isf) atomicAdd(&hist[2 * zrel + 1], 1u);
        }
        unsigned incl = ng;                                        // inclusive warp scan of the general-chunk counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const unsigned totg = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned bf = __ballot_sync(0xffffffffu, isf);
        unsigned baseg = 0, basef = 0;
        if (lane == 0) {
            if (totg) baseg = atomicAdd(&st->int_items[parity * 2 + 1], totg);
            if (bf) basef = atomicAdd(&st->int_items[parity * 2 + 0], (unsigned)__popc(bf));
        }
        baseg = __shfl_sync(0xffffffffu, baseg, 0) + (incl - ng); basef = __shfl_sync(0xffffffffu, basef, 0);
        const int lo_rel = xa - seg * INT_SEG, hi_rel = xe - seg * INT_SEG;
        if (isf) items[basef + __popc(bf & ((1u << lane) - 1))] = make_free_item(A, yg, zrel, seg, lo_rel, hi_rel, fm);
        for (unsigned m2 = gm; m2; m2 &= m2 - 1)
            items[capacity - 1 - baseg++] = make_general_item(A, yg, zrel, seg, lo_rel, hi_rel, __ffs((int)m2) - 1);
    }
}

// t' and w' of one voxel (tsdfvolume::updateVoxel, tsdfVolume.h:57-74, weight 1) with the reference's two IEEE divisions
// skipped where their result is known: fminf(sdf/trunc, 1) is exactly 1 once sdf >= trunc*(1+2^-19), and
// fma(1,w,1)/(w+1) is exactly 1.  Stores only when the voxel's bits change (saturated free space needs no store).
__device__ __forceinline__ void update_voxel(float2* p, const float2 old, float sdf, float trunc, float thr_one, float wmax)
{
    const float w1 = __fadd_rn(old.y, 1.0f);
    const float wn = fminf(w1, wmax);
    float tn = 1.0f;
    if (!(sdf >= thr_one && old.x == 1.0f)) {
        const float tsdf = (sdf >= thr_one) ? 1.0f : fminf(__fdiv_rn(sdf, trunc), 1.0f);
        tn = __fdiv_rn(__fmaf_rn(old.x, old.y, tsdf), w1);
    }
    if (tn != old.x || wn != old.y) *p = make_float2(tn, wn);
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk) global -> shared memory with completion on an mbarrier.  One 4 KB
// stage per WARP: the warp's elected lane arms the barrier with the byte count and issues one copy per voxel row; all
// lanes later wait on the barrier's phase.  One stage is enough: the next brick is requested the moment the current one
// has been read, and the warp then spends ~2 us projecting general chunks - longer than a DRAM round trip.
// tests/emu runs the same code with the copy as a memcpy.
#define INT_STAGE_VOX (INT_G * INT_SEG)              // 4 rows x 128 voxels = 4 KB
#define INT_WARPS (INT_THREADS / 32)
#define INT_SMEM_BYTES (INT_WARPS * INT_STAGE_VOX * 8 + INT_WARPS * 8)

#ifdef HKF_EMULATE_DEVICE
__device__ __forceinline__ void mbar_init(unsigned long long* bar) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long*, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(unsigned long long*, unsigned) {}
__device__ __forceinline__ void mbar_fence_init() {}
#else
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
#endif

// Pass 2: persistent warps stride over both work lists; a lane owns voxels x = first + 32 c + lane of each row, so
// every warp load/store instruction moves 256 contiguous bytes.  Per warp: the rows of its next free brick are always
// in flight as bulk copies; between two bricks the warp projects its share of general chunks (in proportion to the
// lengths of the two lists, so that both run out together).
//
// COLOR (has_color; pass 1 ran with NOFREE, so the free list is empty): a voxel is also dropped when its colour pixel
//   cu = (int)(X*525/Z + 320), cv = (int)(Y*525/Z + 240)     (mul, div.rn, add, cvt.rzi: integrateVolume.cu:56-57, the
//                                                              hard-coded intrinsics are the reference's)
// is outside [1, rgb_cols-2] x [1, rgb_rows-2]; an updated voxel blends its colour (tsdfVolume.h:57-74 as compiled)
//   wc = 2, or 2 * fminf(|n.z| / 0.75, 1) with angle weighting (n = the new frame's normal at the depth pixel)
//   c' = (uchar) trunc(fminf(fma(wc, c_pixel, w * c_old) / (wc + w), 255))        (w = the voxel's OLD weight)
template <bool FROM_STATE, bool COUNT_ONLY, bool COLOR>
__global__ void __launch_bounds__(INT_THREADS, 4)
k_integrate(IntegrateArgs A, const float* __restrict__ depth, float2* __restrict__ vol, HkfDeviceState* __restrict__ st,
            const uint4* __restrict__ items, unsigned capacity, int parity, IntegrateColor CA)
{
    HKF_DYN_SMEM(unsigned long long, dyn);
    hkf_pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned R = (unsigned)A.R;
    const unsigned n_free = st->int_items[parity * 2 + 0], n_gen = st->int_items[parity * 2 + 1];
    const unsigned wid = (blockIdx.x * INT_THREADS + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * INT_THREADS) >> 5;
    const float cell = A.cell;
    const float trunc = A.trunc, thr_one = A.trunc * 1.000002f, wmax = A.wmax;
    unsigned local_count = 0;

    // this warp's stage and barrier
    float2* const stage = reinterpret_cast<float2*>(dyn) + (size_t)warp * INT_STAGE_VOX;
    unsigned long long* const bar = dyn + (size_t)INT_WARPS * INT_STAGE_VOX + warp;
    if (!COUNT_ONLY) {
        if (lane == 0) { mbar_init(bar); mbar_fence_init(); }
        __syncwarp();
    }

    // ---- free bricks.  fi = the brick in flight (next to consume), strided by nwarps; `pend` = its work item.  The item of
    // the brick after it is fetched at the START of a consume and used at its end, so that neither the consume nor the
    // next request waits for an item load (7 % of the kernel's stall samples before: profiles/r2_integrate_e_rmq.txt).
    unsigned fi = wid, phase = 0;
    uint4 pend = make_uint4(0u, 0u, 0u, 0u);
    auto free_issue = [&](const uint4& it) {
        if (lane == 0) {
            const unsigned off = (it.w >> 8) & 0xffu, bytes = it.w >> 16;
            const float2* src = vol + ((size_t)it.x * 4u + off);
            mbar_expect_tx(bar, INT_G * bytes);
#pragma unroll
            for (int r = 0; r < INT_G; ++r) bulk_g2s(stage + r * INT_SEG + off, src + (size_t)r * R, bytes, bar);
        }
#ifdef HKF_EMULATE_DEVICE
        __syncwarp();                         // the host build copies synchronously on lane 0's thread
#endif
    };
    auto free_consume = [&]() {
        const uint4 cur = pend;
        const bool more = fi + nwarps < n_free;
        uint4 nx = cur;
        if (more) nx = __ldg(items + (fi + nwarps));
        const int rel = lane - (int)(cur.z & 0xffu);
        const unsigned span = ((cur.z >> 8) & 0xffu) - (cur.z & 0xffu);
        const unsigned fre = cur.w & 15u;
        bool live[INT_NC];
#pragma unroll
        for (int c = 0; c < INT_NC; ++c) live[c] = (fre & (1u << c)) && ((unsigned)(rel + 32 * c) <= span);
        if (COUNT_ONLY) {
#pragma unroll
            for (int c = 0; c < INT_NC; ++c) local_count += live[c] ? (unsigned)INT_G : 0u;
        } else {
            float2* __restrict__ vp = vol + ((size_t)cur.x * 4u + (unsigned)lane);
            const float2* sp = stage + lane;
            mbar_wait(bar, phase);
            phase ^= 1u;
#pragma unroll
            for (int r = 0; r < INT_G; ++r)
#pragma unroll
                for (int c = 0; c < INT_NC; ++c)
                    if (live[c]) {
                        // tsdf = 1 (free space), weight 1: t' = fma(t, w, 1) / (w + 1); exactly 1 when t == 1 (the common
                        // case: free space stays free) and when the voxel is new (t = w = 0)
                        const float2 old = sp[r * INT_SEG + 32 * c];
                        const float w1 = __fadd_rn(old.y, 1.0f);
                        const float wn = fminf(w1, wmax);
                        if (old.x == 1.0f) {
                            if (wn != old.y) vp[(size_t)r * R + 32 * c] = make_float2(1.0f, wn);
                        } else {
                            const float tn = (old.x == 0.0f && old.y == 0.0f) ? 1.0f : __fdiv_rn(__fmaf_rn(old.x, old.y, 1.0f), w1);
                            vp[(size_t)r * R + 32 * c] = make_float2(tn, wn);      // tn != old.x here or the bits are rewritten as they were
                        }
                    }
            __syncwarp();                     // every lane has read the stage: it may be refilled
        }
        fi += nwarps;
        pend = nx;
        if (!COUNT_ONLY && more) free_issue(nx);
    };
    if (wid < n_free) {
        pend = __ldg(items + wid);
        if (!COUNT_ONLY) free_issue(pend);
    }
    const unsigned my_free = wid < n_free ? (n_free - wid + nwarps - 1) / nwarps : 0u;
    const unsigned my_gen = wid < n_gen ? (n_gen - wid + nwarps - 1) / nwarps : 0u;
    unsigned credit = 0;

    // ---- general chunks: 4 rows x 32 voxels, every voxel is projected
    if (my_gen) {
        float m[12];
        load_pose_inv(A, st, FROM_STATE, m);
        const int cols = A.cols;
        const unsigned ulim = (unsigned)(A.cols - 2), vlim = (unsigned)(A.rows - 2);
        const float ntrunc = -A.trunc;
        const float MAGIC = 12582912.0f;              // 1.5 * 2^23: adding it leaves the nearest integer in the mantissa
        const float edge = 0.5f - 0.00048828125f;     // within 2^-11 px of a rounding boundary -> the exact sequence decides
        uint4 nxt = __ldg(items + (capacity - 1 - wid));
#pragma unroll 1
        for (unsigned item = wid; item < n_gen; item += nwarps) {
            const uint4 cur = nxt;
            if (item + nwarps < n_gen) nxt = __ldg(items + (capacity - 1 - (item + nwarps)));      // prefetch the next item
            const bool inr = (unsigned)(lane - (int)(cur.z & 0xffu)) <= ((cur.z >> 8) & 0xffu) - (cur.z & 0xffu);
            float2* __restrict__ vp = vol + ((size_t)cur.x * 4u + (unsigned)lane);
            // The voxels are requested first: the DRAM round trip then overlaps the projection arithmetic and the depth
            // read instead of following them (a few voxels that turn out not to be updated are read in vain).
            float2 old[INT_G];
#pragma unroll
            for (int r = 0; r < INT_G; ++r) {
                old[r] = make_float2(0.f, 0.f);
                if (!COUNT_ONLY && inr) old[r] = vp[(size_t)r * R];
            }
            const int y0 = (int)(cur.y & 0x3ffu) * INT_G;
            const float wz = __fmul_rn(__fadd_rn((float)((int)((cur.y >> 10) & 0xfffu) + A.z_lo), 0.5f), cell);
            const float wx = __fmul_rn((float)((int)(cur.y >> 22) * 32 + lane) + 0.5f, cell);
            // the x and z parts of the three camera-space sums are shared by the four rows only in exact arithmetic: the
            // reference's association is fma(wz, m_3, fma(wx, m_1, wy*m_2)) + m_4 per voxel, kept as is
            float Z[INT_G];
            int idx[INT_G];
            int cidx[INT_G];                  // COLOR: index of the colour pixel, -1 = outside the rgb image
#pragma unroll
            for (int r = 0; r < INT_G; ++r) {
                const float wy = __fmul_rn(__fadd_rn((float)(y0 + r), 0.5f), cell);
                Z[r] = __fadd_rn(__fmaf_rn(wz, m[10], __fmaf_rn(wx, m[8], __fmul_rn(wy, m[9]))), m[11]);
                const float X = __fadd_rn(__fmaf_rn(wz, m[2], __fmaf_rn(wx, m[0], __fmul_rn(wy, m[1]))), m[3]);
                const float Y = __fadd_rn(__fmaf_rn(wz, m[6], __fmaf_rn(wx, m[4], __fmul_rn(wy, m[5]))), m[7]);
                const float ax = __fmul_rn(X, A.fx), ay = __fmul_rn(Y, A.fy);
                bool in = inr && (Z[r] > 0.f);
                // approximate pixel coordinate (one MUFU.RCP shared by both axes, < 2e-4 px off) and its nearest
                // integer = floor(p + 0.5), without leaving the FMA/ALU pipes
                const float rz = hkf_rcp_approx(Z[r]);
                const float pu = __fmaf_rn(ax, rz, A.cx), pv = __fmaf_rn(ay, rz, A.cy);
                const float tu = __fadd_rn(pu, MAGIC), tv = __fadd_rn(pv, MAGIC);
                int iu = __float_as_int(tu) - 0x4B400000;
                int iv = __float_as_int(tv) - 0x4B400000;
                const float eu = __fsub_rn(pu, __fsub_rn(tu, MAGIC)), ev = __fsub_rn(pv, __fsub_rn(tv, MAGIC));
                if (in && !(fmaxf(fabsf(eu), fabsf(ev)) < edge)) {     // NaN-safe; ~0.2 % of voxels
                    // the reference's exact sequence: mul, div.rn, add, f64 add, cvt.rzi
                    iu = hkf_round_px(hkf_project1(X, A.fx, Z[r], A.cx));
                    iv = hkf_round_px(hkf_project1(Y, A.fy, Z[r], A.cy));
                }
                in = in && ((unsigned)(iu - 1) < ulim) && ((unsigned)(iv - 1) < vlim);
                idx[r] = in ? iv * cols + iu : -1;
                if (COLOR) {
                    cidx[r] = -1;
                    if (in) {
                        const int cu = __float2int_rz(__fadd_rn(__fdiv_rn(__fmul_rn(X, 525.0f), Z[r]), 320.0f));
                        const int cv = __float2int_rz(__fadd_rn(__fdiv_rn(__fmul_rn(Y, 525.0f), Z[r]), 240.0f));
                        if (!(cu >= CA.rgb_cols - 1 || cv >= CA.rgb_rows - 1 || cu < 1 || cv < 1)) cidx[r] = cv * CA.rgb_cols + cu;
                    }
                }
            }
            float d[INT_G];
#pragma unroll
            for (int r = 0; r < INT_G; ++r) d[r] = (idx[r] >= 0) ? __ldg(depth + idx[r]) : 0.f;
#pragma unroll
            for (int r = 0; r < INT_G; ++r) {
                const float sdf = __fsub_rn(d[r], Z[r]);
                bool live = (d[r] != 0.f) && (d[r] < A.maxdist) && (sdf > ntrunc);
                if (COLOR) live = live && cidx[r] >= 0;
                if (COUNT_ONLY) local_count += live ? 1u : 0u;
                else if (live) {
                    update_voxel(vp + (size_t)r * R, old[r], sdf, trunc, thr_one, wmax);
                    if (COLOR) {
                        float wc = 2.0f;
                        if (CA.angled) {
                            const float q = fminf(__fdiv_rn(fabsf(__ldg(&CA.normals[idx[r]].z)), 0.75f), 1.0f);
                            wc = __fadd_rn(q, q);
                        }
                        uchar4* cp = CA.color_vol + ((size_t)cur.x * 4u + (unsigned)lane + (size_t)r * R);
                        const uchar4 oc = *cp;
                        const unsigned char* px = CA.rgb + (size_t)cidx[r] * 3;
                        const float den = __fadd_rn(wc, old[r].y);
                        uchar4 nc;
                        nc.x = (unsigned char)__float2uint_rz(fminf(__fdiv_rn(__fmaf_rn(wc, (float)__ldg(px + 0), __fmul_rn(old[r].y, (float)oc.x)), den), 255.0f));
                        nc.y = (unsigned char)__float2uint_rz(fminf(__fdiv_rn(__fmaf_rn(wc, (float)__ldg(px + 1), __fmul_rn(old[r].y, (float)oc.y)), den), 255.0f));
                        nc.z = (unsigned char)__float2uint_rz(fminf(__fdiv_rn(__fmaf_rn(wc, (float)__ldg(px + 2), __fmul_rn(old[r].y, (float)oc.z)), den), 255.0f));
                        nc.w = 0;
                        *cp = nc;
                    }
                }
            }
            // a free brick whenever the general list has advanced by its share (my_free : my_gen)
            credit += my_free;
            while (credit >= my_gen && fi < n_free) { credit -= my_gen; free_consume(); }
        }
    }
    // ---- the free bricks that are left (all of them when this warp had no general chunk)
    while (fi < n_free) free_consume();

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
    A->B = c->bounds;
}

#ifndef HKF_EMULATE_DEVICE      /* host launch code below (tests/emu stops here) */
static bool fill_color(hkf_ctx* c, const HkfCameraParams* rgb_cam, int angled, IntegrateColor* CA)
{
    if (!c->color || !c->raw_rgb) { c->err = "context was created without use_color"; return false; }
    if (rgb_cam->cols != c->ip.rgb_camera.cols || rgb_cam->rows != c->ip.rgb_camera.rows) { c->err = "rgb camera size mismatch"; return false; }
    CA->color_vol = c->color; CA->rgb = c->raw_rgb; CA->normals = c->new_n[0];
    CA->rgb_cols = (int)rgb_cam->cols; CA->rgb_rows = (int)rgb_cam->rows; CA->angled = angled ? 1 : 0;
    return true;
}

// both passes, on the context's stream
template <bool FROM_STATE, bool COUNT_ONLY, bool COLOR>
static int launch_integrate(hkf_ctx* c, const IntegrateArgs& A, const IntegrateColor& CA = IntegrateColor{})
{
    const int ngroups = (A.z_hi - A.z_lo) * (A.R / INT_G);
    const int parity = c->int_parity;
    c->int_parity ^= 1;
    const dim3 g1(div_up(ngroups, INT_GB)), b1(256);
    if (c->int_hist_on) hkf_launch_pdl(k_integrate_items<FROM_STATE, COLOR, true>, g1, b1, 0, c->stream, A, c->st, c->int_items, c->int_capacity, parity, c->int_hist);
    else hkf_launch_pdl(k_integrate_items<FROM_STATE, COLOR, false>, g1, b1, 0, c->stream, A, c->st, c->int_items, c->int_capacity, parity, (unsigned*)nullptr);
    HKF_LAUNCH_CHECK(c);
    // function attributes are per device: set once per context (dynamic shared memory above the 48 KB default)
    const int variant = COLOR ? 2 : COUNT_ONLY ? 1 : 0;
    if (!c->int_attr_set[variant][FROM_STATE ? 1 : 0]) {
        HKF_CUDA(c, cudaFuncSetAttribute(k_integrate<FROM_STATE, COUNT_ONLY, COLOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INT_SMEM_BYTES));
        c->int_attr_set[variant][FROM_STATE ? 1 : 0] = true;
    }
    hkf_launch_pdl(k_integrate<FROM_STATE, COUNT_ONLY, COLOR>, dim3(c->sm_count * 4), dim3(INT_THREADS), (size_t)INT_SMEM_BYTES, c->stream, A, (const float*)c->trunc_depth, c->vol, c->st,
                   (const uint4*)c->int_items, c->int_capacity, parity, CA);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}


extern "C" int hkf_integrate_volume(hkf_ctx* c, int has_color, int use_angle_weight_color, const HkfMat44* transform,
                                    const HkfIntegrateParams* ip, const HkfCameraParams* cam, const HkfCameraParams* rgb_cam)
{
    if (!c || !transform || !ip || !cam) return HKF_ERR_ARG;
    { const int rc = hkf_shard_flush(c); if (rc != HKF_OK) return rc; }      // see hkf_integrate_tracked
    if (cam->cols != c->cols[0] || cam->rows != c->rows[0]) { c->err = "camera size mismatch"; return HKF_ERR_ARG; }
    float inv[16];
    hkf_mat44_inverse(transform->m, inv);      // host, float cofactors: cudaIntegrateVolume line 84
    IntegrateArgs A; fill_args(c, inv, ip, cam, &A);
    if (has_color) {
        IntegrateColor CA;
        if (!rgb_cam) return HKF_ERR_ARG;
        if (!fill_color(c, rgb_cam, use_angle_weight_color, &CA)) return HKF_ERR_STATE;
        return launch_integrate<false, false, true>(c, A, CA);
    }
    return launch_integrate<false, false, false>(c, A);
}

extern "C" int hkf_integrate_tracked(hkf_ctx* c, const HkfFrameParams* fp)
{
    if (!c || !fp) return HKF_ERR_ARG;
    // sharded mode: a deferred compositing is issued first.  With a peer-memory halo the neighbours' slab raycasts of the
    // previous frame read THIS rank's planes; the compositing handshake is what orders this integration after them
    // (a caller that supplies poses never passes through hkf_icp_track, which would have flushed).
    { const int rc = hkf_shard_flush(c); if (rc != HKF_OK) return rc; }
    IntegrateArgs A; fill_args(c, nullptr, &fp->integrate, &c->ip.depth_camera, &A);
    if (fp->use_color) {
        IntegrateColor CA;
        if (!fill_color(c, &c->ip.rgb_camera, (int)fp->color_angle_weight, &CA)) return HKF_ERR_STATE;
        return launch_integrate<true, false, true>(c, A, CA);
    }
    return launch_integrate<true, false, false>(c, A);
}

// Work profile for slab balancing: while enabled, every integration adds the number of general chunks and free bricks
// of each stored plane to a histogram (two atomics per 4 x 128 voxel brick of the item pass; off by default).
extern "C" int hkf_integrate_profile(hkf_ctx* c, int enable)
{
    if (!c) return HKF_ERR_ARG;
    const size_t n = 2 * (size_t)(c->z_hi - c->z_lo);
    if (enable) {
        if (!c->int_hist) HKF_CUDA(c, cudaMalloc(&c->int_hist, n * sizeof(unsigned)));
        HKF_CUDA(c, cudaMemsetAsync(c->int_hist, 0, n * sizeof(unsigned), c->stream));
    }
    c->int_hist_on = enable != 0;
    return HKF_OK;
}

extern "C" long hkf_integrate_profile_read(hkf_ctx* c, uint32_t* out, size_t cap_entries)
{
    if (!c || !out) return -HKF_ERR_ARG;
    const size_t n = 2 * (size_t)(c->z_hi - c->z_lo);
    if (!c->int_hist) { c->err = "hkf_integrate_profile_read: profiling was never enabled"; return -HKF_ERR_STATE; }
    if (cap_entries < n) { c->err = "hkf_integrate_profile_read: destination too small"; return -HKF_ERR_ARG; }
    cudaError_t e = cudaMemcpyAsync(out, c->int_hist, n * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { c->err = std::string("hkf_integrate_profile_read: ") + cudaGetErrorString(e); return -HKF_ERR_CUDA; }
    return (long)n;
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
    int rc = launch_integrate<false, true, false>(c, A);
    if (rc != HKF_OK) return rc;
    HKF_CUDA(c, cudaMemcpyAsync(n_upd, &c->st->upd_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    HKF_CUDA(c, cudaStreamSynchronize(c->stream));
    return HKF_OK;
}
#endif // HKF_EMULATE_DEVICE
