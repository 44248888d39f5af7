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
// profiles/r1_integrate_*):
//  * storage is packed {f32 tsdf, f32 weight} (8 B) in [z][y][x] order, x fastest; a lane owns voxel x = chunk + lane,
//    so every warp load/store moves 256 contiguous bytes;
//  * pass 1 (k_integrate_items) intersects every group of 4 voxel rows (straight lines in camera space) with the view
//    frustum and the frame's depth range analytically, then classifies each 32-voxel chunk of the interval against
//    per-tile depth extrema published by the preprocess: dead (behind every visible surface), free (in front of every
//    visible surface, all touched pixels integrable: tsdf = 1 exactly) or general, and writes two compact work lists;
//  * pass 2 (k_integrate): persistent warps stride over the lists.  Free bricks are a pure streaming read-modify-write
//    (no projection, no depth read, 16 loads in flight per warp).  General chunks (4 rows x 32 voxels) request their
//    voxels first and then project.  The first ncu capture of a straightforward kernel showed it ISSUE-bound (74 %
//    issue-active, 21 % DRAM: ~170 instructions per voxel, dominated by three IEEE divisions, f64 rounding and
//    int<->float conversions on the quarter-rate XU pipe), so the projection decides everything it can from cheap
//    approximations and only falls back to the reference's exact sequence where they cannot prove the result:
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

// work item: .x = y/4 | zrel << 10 | segment << 22 ; .y = first | last << 8 (voxel offsets inside the 128-voxel
// segment that can be updated) | mask of chunks to project << 16 | mask of free-space chunks << 20
__device__ __forceinline__ uint2 make_item(int yg, int zrel, int seg, int lo_rel, int hi_rel, unsigned gen, unsigned fre)
{
    return make_uint2((unsigned)yg | ((unsigned)zrel << 10) | ((unsigned)seg << 22),
                      (unsigned)lo_rel | ((unsigned)hi_rel << 8) | (gen << 16) | (fre << 20));
}

// Pass 1 (thread per group of 4 voxel rows): the group's conservative interval, cut into 32-voxel chunks that are
// classified against the frame's per-tile depth extrema, and the surviving 4 x 128 voxel bricks appended to two
// compact work lists (one warp-aggregated atomic per list per 32 groups): bricks made of free-space chunks only
// grow from the front of `items`, bricks with chunks that need projecting grow from the back.
//
// Chunk classification.  The voxel centres of a 4-row x 32-voxel chunk form a planar rectangle; while Z > 0 its
// image is the convex hull of its four projected corners, and depth along it is extremal at a corner.  The
// bounding box of the corners, grown by 1.5 px, touches at most 2 tiles per tile row at the level whose tile width
// covers it.  From the depth extrema of those tiles:
//   dead : the nearest corner is farther than trunc behind the largest valid depth -> "sdf > -trunc"
//          (integrateVolume.cu:64) fails for every voxel, the chunk is dropped;
//   free : the box lies in the image interior, every pixel of the tiles is valid and < fMaxIntegrateDist, and the
//          farthest corner is more than trunc in front of the smallest depth -> every voxel is updated with
//          tsdf = fminf(sdf/trunc, 1) = 1 exactly, so pass 2 needs neither a projection nor a depth read for it;
//   else : general, decided voxel by voxel in pass 2 with the reference's arithmetic.
#define INT_GB 64                      // row groups per block of pass 1

// A chunk's pixel box may span up to INT_SPAN tile widths of the level it is tested on (up to INT_SPAN+1 tiles per
// tile row): finer tiles = tighter depth extrema.  Measured on the bench scene: 19.0 M general voxel slots with 3
// instead of 21.7 M with 1, integration 144.8 -> 142.5 us (profiles/r2_experimental.log).
#define INT_SPAN 3
template <bool FROM_STATE>
__global__ void __launch_bounds__(256)
k_integrate_items(IntegrateArgs A, HkfDeviceState* __restrict__ st, uint2* __restrict__ items, unsigned capacity, int parity)
{
    __shared__ int s_lo[INT_GB], s_hi[INT_GB];
    hkf_pdl_enter();
    const int lane = threadIdx.x & 31;
    const int RG = A.R / INT_G;
    const int ngroups = (A.z_hi - A.z_lo) * RG;
    const int nseg = (A.R + INT_SEG - 1) / INT_SEG;
    if (blockIdx.x == 0 && threadIdx.x < 2) st->int_items[(parity ^ 1) * 2 + threadIdx.x] = 0;   // the counters the NEXT call appends to
    if (FROM_STATE && st->track_status != HKF_TRACK_OK) return;   // "camera lost": no integration (HybKinectfu.cpp:123-147)
    float m[12];
    load_pose_inv(A, st, FROM_STATE, m);

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
    }
    __syncthreads();

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
                    const float w = u1 - u0;
                    const int lvl = w <= (float)(INT_SPAN * HKF_TILE_W) ? 0 : w <= (float)(INT_SPAN * 2 * HKF_TILE_W) ? 1 : w <= (float)(INT_SPAN * 4 * HKF_TILE_W) ? 2
                                  : w <= (float)(INT_SPAN * 8 * HKF_TILE_W) ? 3 : -1;
                    const int ty0 = max((int)floorf(v0 * (1.0f / HKF_TILE_H)), 0);
                    const int ty1 = min((int)floorf(v1 * (1.0f / HKF_TILE_H)), A.B.tiles_y - 1);
                    if (lvl >= 0 && ty1 - ty0 <= 3) {
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
                        if (zn >= dmx + thr_dead || dmn >= A.maxdist) cls = 0;     // behind everything / nothing integrable there
                        else if (interior && dmn > 0.f && dmx < A.maxdist && dmn - zf >= thr_free) cls = 2;
                    }
                }
                if (cls == 1) gm |= 1u << c;
                if (cls == 2) fm |= 1u << c;
                // the next chunk starts one voxel after this end: [this end, next end] is a superset
                pu0 = qu0; pu1 = qu1; pv0 = qv0; pv1 = qv1; pz0 = qz0; pz1 = qz1;
            }
        }
        // Work lists: the free-space chunks of a brick form ONE item (pure streaming, 16 loads in flight) that grows
        // from the front of `items`; every general chunk (4 rows x 32 voxels, all four rows need projecting) is an item
        // of its own growing from the back, so pass 2 never spends an instruction on a dead or free chunk slot.
        const unsigned ng = (unsigned)__popc(gm);
        const bool isf = fm != 0;
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
        if (isf) items[basef + __popc(bf & ((1u << lane) - 1))] = make_item(yg, zrel, seg, lo_rel, hi_rel, 0u, fm);
        for (unsigned m2 = gm; m2; m2 &= m2 - 1)
            items[capacity - 1 - baseg++] = make_item(yg, zrel, seg, lo_rel, hi_rel, m2 & (0u - m2), 0u);
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

// Pass 2: persistent warps stride over the work items; a lane owns voxels x = seg*128 + 32 c + lane of each row, so
// every warp load/store instruction moves 256 contiguous bytes.
template <bool FROM_STATE, bool COUNT_ONLY>
__global__ void __launch_bounds__(INT_THREADS, 4)
k_integrate(IntegrateArgs A, const float* __restrict__ depth, float2* __restrict__ vol, HkfDeviceState* __restrict__ st,
            const uint2* __restrict__ items, unsigned capacity, int parity)
{
    hkf_pdl_enter();
    const int lane = threadIdx.x & 31;
    const int R = A.R;
    const unsigned n_free = st->int_items[parity * 2 + 0], n_gen = st->int_items[parity * 2 + 1];
    const unsigned wid = (blockIdx.x * INT_THREADS + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * INT_THREADS) >> 5;
    const float cell = A.cell;
    const float trunc = A.trunc, thr_one = A.trunc * 1.000002f, wmax = A.wmax;
    unsigned local_count = 0;

    // ---- general chunks (the heavier items first): 4 rows x 32 voxels, every voxel is projected
    if (wid < n_gen) {
        float m[12];
        load_pose_inv(A, st, FROM_STATE, m);
        const int cols = A.cols;
        const unsigned ulim = (unsigned)(A.cols - 2), vlim = (unsigned)(A.rows - 2);
        const float ntrunc = -A.trunc;
        const float MAGIC = 12582912.0f;              // 1.5 * 2^23: adding it leaves the nearest integer in the mantissa
        const float edge = 0.5f - 0.00048828125f;     // within 2^-11 px of a rounding boundary -> the exact sequence decides
        uint2 nxt = __ldg(items + (capacity - 1 - wid));
#pragma unroll 1
        for (unsigned item = wid; item < n_gen; item += nwarps) {
            const uint2 cur = nxt;
            if (item + nwarps < n_gen) nxt = __ldg(items + (capacity - 1 - (item + nwarps)));      // prefetch the next item
            const int y0 = (int)(cur.x & 0x3ffu) * INT_G, zrel = (int)((cur.x >> 10) & 0xfffu);
            const int c = __ffs((int)((cur.y >> 16) & 15u)) - 1;                 // the chunk of the segment
            const int xb = (int)(cur.x >> 22) * INT_SEG + 32 * c;
            const bool inr = (unsigned)(lane + 32 * c - (int)(cur.y & 0xffu)) <= ((cur.y >> 8) & 0xffu) - (cur.y & 0xffu);
            float2* __restrict__ vp = vol + ((size_t)(zrel * R + y0) * R + xb + lane);
            // The voxels are requested first: the DRAM round trip then overlaps the projection arithmetic and the depth
            // read instead of following them (a few voxels that turn out not to be updated are read in vain).
            float2 old[INT_G];
#pragma unroll
            for (int r = 0; r < INT_G; ++r) {
                old[r] = make_float2(0.f, 0.f);
                if (!COUNT_ONLY && inr) old[r] = vp[(size_t)r * R];
            }
            const float wz = __fmul_rn(__fadd_rn((float)(zrel + A.z_lo), 0.5f), cell);
            const float wx = __fmul_rn((float)(xb + lane) + 0.5f, cell);
            // the x and z parts of the three camera-space sums are shared by the four rows only in exact arithmetic: the
            // reference's association is fma(wz, m_3, fma(wx, m_1, wy*m_2)) + m_4 per voxel, kept as is
            float Z[INT_G];
            int idx[INT_G];
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
            }
            float d[INT_G];
#pragma unroll
            for (int r = 0; r < INT_G; ++r) d[r] = (idx[r] >= 0) ? __ldg(depth + idx[r]) : 0.f;
#pragma unroll
            for (int r = 0; r < INT_G; ++r) {
                const float sdf = __fsub_rn(d[r], Z[r]);
                const bool live = (d[r] != 0.f) && (d[r] < A.maxdist) && (sdf > ntrunc);
                if (COUNT_ONLY) local_count += live ? 1u : 0u;
                else if (live) update_voxel(vp + (size_t)r * R, old[r], sdf, trunc, thr_one, wmax);
            }
        }
    }

    // ---- free-space bricks: a pure streaming read-modify-write, 16 x 256 B in flight per warp
#pragma unroll 1
    for (unsigned item = wid; item < n_free; item += nwarps) {
        const uint2 cur = __ldg(items + item);
        const int y0 = (int)(cur.x & 0x3ffu) * INT_G, zrel = (int)((cur.x >> 10) & 0xfffu);
        const int xb = (int)(cur.x >> 22) * INT_SEG;
        const int rel = lane - (int)(cur.y & 0xffu);
        const unsigned span = ((cur.y >> 8) & 0xffu) - (cur.y & 0xffu);
        const unsigned fre = (cur.y >> 20) & 15u;
        float2* __restrict__ vp = vol + ((size_t)(zrel * R + y0) * R + xb + lane);
        bool live[INT_NC];
#pragma unroll
        for (int c = 0; c < INT_NC; ++c) live[c] = (fre & (1u << c)) && ((unsigned)(rel + 32 * c) <= span);
        if (COUNT_ONLY) {
#pragma unroll
            for (int c = 0; c < INT_NC; ++c) local_count += live[c] ? (unsigned)INT_G : 0u;
        } else {
            float2 old[INT_G][INT_NC];
#pragma unroll
            for (int r = 0; r < INT_G; ++r)
#pragma unroll
                for (int c = 0; c < INT_NC; ++c) {
                    old[r][c] = make_float2(0.f, 0.f);
                    if (live[c]) old[r][c] = vp[(size_t)r * R + 32 * c];
                }
#pragma unroll
            for (int r = 0; r < INT_G; ++r)
#pragma unroll
                for (int c = 0; c < INT_NC; ++c)
                    if (live[c]) update_voxel(vp + (size_t)r * R + 32 * c, old[r][c], 1e30f, trunc, thr_one, wmax);
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
    A->B = c->bounds;
}

#ifndef HKF_EMULATE_DEVICE      /* host launch code below (tests/emu stops here) */
// both passes, on the context's stream
template <bool FROM_STATE, bool COUNT_ONLY>
static int launch_integrate(hkf_ctx* c, const IntegrateArgs& A)
{
    const int ngroups = (A.z_hi - A.z_lo) * (A.R / INT_G);
    const int parity = c->int_parity;
    c->int_parity ^= 1;
    const dim3 g1(div_up(ngroups, INT_GB)), b1(256);
    hkf_launch_pdl(k_integrate_items<FROM_STATE>, g1, b1, 0, c->stream, A, c->st, c->int_items, c->int_capacity, parity);
    HKF_LAUNCH_CHECK(c);
    hkf_launch_pdl(k_integrate<FROM_STATE, COUNT_ONLY>, dim3(c->sm_count * 4), dim3(INT_THREADS), 0, c->stream, A, (const float*)c->trunc_depth, c->vol, c->st,
                   (const uint2*)c->int_items, c->int_capacity, parity);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
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
#endif // HKF_EMULATE_DEVICE
