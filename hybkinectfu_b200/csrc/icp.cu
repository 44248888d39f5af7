// icp.cu — projective point-to-plane ICP: correspondence, the 27 normal-equation sums, and (fused
// mode) the 6x6 solve, gates and pose update, one launch per iteration with nothing read back.
//
// Replaces computeGbufKernel + findCorrs + buildPointToPlaneSolverRows + reduceGbufKernel
// (src/cuda/CalPointToPlaneErrSolverParams.cu:7-129, src/cuda/device_functions.h:11-55) and, in fused
// mode, the host loop of CameraPoseFinderICP (src/CameraPoseFinderICP.cpp:50-145).
//
// Reference shape: 256-thread blocks run 27 sequential shared-memory tree reductions (108 barriers per
// block), a second kernel reduces the block partials, then the host blocks on a 27-float D2H copy and
// solves with Eigen, 19 times per frame.
//
// This design: one warp owns one 16x16 pixel tile (the reference's block), a lane owns the 8 pixels
// tid = lane + 32 j of that tile.  The reference's block reduction adds tid with tid+128, +64, +32 and
// then folds the remaining 32 values 16/8/4/2/1: in our mapping the first three levels are additions
// among the 8 values a lane already holds in registers, ((p0+p4)+(p2+p6))+((p1+p5)+(p3+p7)), and the
// last five are a __shfl_down tree.  No shared memory, no barrier, no atomics on data, and the float
// sums are BIT-IDENTICAL to the reference's.  The last block to finish (one ticket atomic per block)
// folds the per-tile partials with the reference's second-stage tree (512-strided sums, then
// 256/128/64/32 in registers, then shuffles), and thread 0 runs the Cholesky solve, the determinant and
// shake gates and the pose update from hkf_solver.h.  ICP's working set (<= 20 MB) is L2-resident on
// B200, so this stage is latency-bound, not HBM-bound: what matters is the number of dependent
// launches/barriers, which drops from 19 x (2 launches + sync + D2H + host solve) to 19 launches.
#include "hkf_internal.cuh"
#include "hkf_solver.h"
#include "reduce27.cuh"

struct IcpArgs {
    const float4* new_v; const float4* new_n; const float4* model_v; const float4* model_n;
    int cols, rows;            // level size
    int gx, ntiles;            // tiles per row, total tiles (= the reference's grid)
    int y0, y1;                // pixel-row band handled by this rank (sharded mode); whole image otherwise
    float cx, cy, fx, fy;
    float dist_thres, sin_thres;
    float angle_shake, dist_shake;
    float* partials;           // [27][ntiles]
    int is_last;               // last iteration of the schedule: commit cur -> pose
    float cur[16], last_inv[16];   // only used when !FROM_STATE
};

// Correspondence + solver row for a batch of NB pixels of one lane, in three phases so that the NB
// coalesced loads and then the NB gathers are all in flight together (findCorrs +
// buildPointToPlaneSolverRows; r[.][7] = 0 when there is no correspondence).
template <int NB>
__device__ __forceinline__ void icp_rows(const IcpArgs& A, const float* __restrict__ cur, const float* __restrict__ li,
                                         const int* xs, const int* ys, float (*r)[7])
{
    float4 n[NB], v[NB];
    bool ok[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        ok[b] = !(xs[b] >= A.cols || ys[b] >= A.rows || ys[b] < A.y0 || ys[b] >= A.y1);
        const size_t idx = ok[b] ? (size_t)ys[b] * A.cols + xs[b] : 0;
        n[b] = ld_f4(A.new_n + idx);
        v[b] = ld_f4(A.new_v + idx);
    }
    float qx[NB], qy[NB], qz[NB], nix[NB], niy[NB], niz[NB];
    size_t midx[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        ok[b] = ok[b] && !hkf_is_zero4(n[b]);
        // v_input_g = cur * v (lead-first), n_input_g = cur * n (lead-second): as contracted in the reference PTX
        qx[b] = hkf_row4<true>(cur[0], cur[1], cur[2], cur[3], v[b].x, v[b].y, v[b].z, v[b].w);
        qy[b] = hkf_row4<true>(cur[4], cur[5], cur[6], cur[7], v[b].x, v[b].y, v[b].z, v[b].w);
        qz[b] = hkf_row4<true>(cur[8], cur[9], cur[10], cur[11], v[b].x, v[b].y, v[b].z, v[b].w);
        const float qw = hkf_row4<true>(cur[12], cur[13], cur[14], cur[15], v[b].x, v[b].y, v[b].z, v[b].w);
        nix[b] = hkf_row4<false>(cur[0], cur[1], cur[2], cur[3], n[b].x, n[b].y, n[b].z, n[b].w);
        niy[b] = hkf_row4<false>(cur[4], cur[5], cur[6], cur[7], n[b].x, n[b].y, n[b].z, n[b].w);
        niz[b] = hkf_row4<false>(cur[8], cur[9], cur[10], cur[11], n[b].x, n[b].y, n[b].z, n[b].w);
        // v_cp = last_inv * v_input_g (lead-second); projective association, no z > 0 test
        const float px = hkf_row4<false>(li[0], li[1], li[2], li[3], qx[b], qy[b], qz[b], qw);
        const float py = hkf_row4<false>(li[4], li[5], li[6], li[7], qx[b], qy[b], qz[b], qw);
        const float pz = hkf_row4<false>(li[8], li[9], li[10], li[11], qx[b], qy[b], qz[b], qw);
        const int u = hkf_round_px(hkf_project1(px, A.fx, pz, A.cx));
        const int w = hkf_round_px(hkf_project1(py, A.fy, pz, A.cy));
        ok[b] = ok[b] && !(u < 0 || u >= A.cols || w < 0 || w >= A.rows);
        midx[b] = ok[b] ? (size_t)w * A.cols + u : 0;
    }
    float4 nt[NB], p[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        nt[b] = ld_f4(A.model_n + midx[b]);
        p[b] = ld_f4(A.model_v + midx[b]);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        ok[b] = ok[b] && !hkf_is_zero4(nt[b]);
        const float dx = __fsub_rn(p[b].x, qx[b]), dy = __fsub_rn(p[b].y, qy[b]), dz = __fsub_rn(p[b].z, qz[b]);
        const float d = __fsqrt_rn(hkf_dot3(dx, dy, dz, dx, dy, dz));
        const float sx = hkf_det2(nt[b].y, niz[b], nt[b].z, niy[b]);
        const float sy = hkf_det2(nt[b].z, nix[b], nt[b].x, niz[b]);
        const float sz = hkf_det2(nt[b].x, niy[b], nt[b].y, nix[b]);
        const float s = __fsqrt_rn(hkf_dot3(sx, sy, sz, sx, sy, sz));
        ok[b] = ok[b] && !(d > A.dist_thres || s > A.sin_thres);
        r[b][0] = ok[b] ? hkf_det2(qy[b], nt[b].z, qz[b], nt[b].y) : 0.f;
        r[b][1] = ok[b] ? hkf_det2(qz[b], nt[b].x, qx[b], nt[b].z) : 0.f;
        r[b][2] = ok[b] ? hkf_det2(qx[b], nt[b].y, qy[b], nt[b].x) : 0.f;
        r[b][3] = ok[b] ? nt[b].x : 0.f;
        r[b][4] = ok[b] ? nt[b].y : 0.f;
        r[b][5] = ok[b] ? nt[b].z : 0.f;
        r[b][6] = ok[b] ? hkf_dot3(nt[b].x, nt[b].y, nt[b].z, dx, dy, dz) : 0.f;
    }
}

#define ICP_WARPS 8

template <bool FROM_STATE>
__global__ void __launch_bounds__(ICP_WARPS * 32)
k_icp(IcpArgs A, HkfDeviceState* st)
{
    __shared__ float s_sums[HKF_SUMS];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long t0 = clock64();
    if (FROM_STATE && st->track_status != HKF_TRACK_OK) return;      // a failed iteration ends the frame's tracking

    float cur[16], li[16];
    if (FROM_STATE) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { cur[i] = st->cur[i]; li[i] = st->last_inv[i]; }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { cur[i] = A.cur[i]; li[i] = A.last_inv[i]; }
    }

    const int tile = blockIdx.x * ICP_WARPS + warp;
    if (tile < A.ntiles) {
        const int bx = tile % A.gx, by = tile / A.gx;
        float r[8][7];
#pragma unroll
        for (int jb = 0; jb < 8; jb += 4) {
            int xs[4], ys[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int tid = lane + 32 * (jb + b);
                xs[b] = bx * 16 + (tid & 15); ys[b] = by * 16 + (tid >> 4);
            }
            icp_rows<4>(A, cur, li, xs, ys, &r[jb]);
        }
        int shift = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 7; ++j) {
                float p[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = __fmul_rn(r[k][i], r[k][j]);
                const float c = tile_tree(p);
                if (lane == 0) A.partials[(size_t)shift * A.ntiles + tile] = c;
                ++shift;
            }
    }

    // ---- last block folds the partials and (fused mode) advances the pose
    const long long t1 = clock64();
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&st->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    const long long t2 = clock64();
    __threadfence();
    second_stage_all<ICP_WARPS>(A.partials, A.ntiles, warp, lane, s_sums);
    __syncthreads();
    const long long t3 = clock64();
    if (threadIdx.x < HKF_SUMS) st->sums[threadIdx.x] = s_sums[threadIdx.x];
    if (threadIdx.x == 0) {
        st->ticket = 0;
        if (FROM_STATE) {
            float c2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) c2[i] = cur[i];
            const int rc = hkf_icp_update(s_sums, A.angle_shake, A.dist_shake, c2);
            if (rc != HKF_SOLVE_OK) {
                st->track_status = (rc == HKF_SOLVE_SINGULAR) ? HKF_TRACK_SINGULAR : HKF_TRACK_SHAKE;
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) st->cur[i] = c2[i];
                if (A.is_last) {                      // _pose = cur_transform (CameraPoseFinderICP.cpp:84)
                    float inv[16];
                    hkf_mat44_inverse(c2, inv);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { st->pose[i] = c2[i]; st->pose_inv[i] = inv[i]; }
                }
            }
        }
        st->dbg[0] = t1 - t0; st->dbg[1] = t2 - t1; st->dbg[2] = t3 - t2; st->dbg[3] = clock64() - t3;
    }
}

// start of a frame's tracking: cur = pose, last_inv = pose^-1 (CameraPoseFinderICP.cpp:62-63)
__global__ void k_track_begin(HkfDeviceState* st)
{
    if (threadIdx.x < 16) {
        st->cur[threadIdx.x] = st->pose[threadIdx.x];
        st->last_inv[threadIdx.x] = st->pose_inv[threadIdx.x];
    }
    if (threadIdx.x == 0) { st->track_status = HKF_TRACK_OK; st->ticket = 0; st->sdf_done = 0; }
}

static void fill_icp_args(hkf_ctx* c, int level, const HkfCameraParams* cam, float dist_thres, float sin_thres, IcpArgs* A)
{
    A->new_v = c->new_v[level]; A->new_n = c->new_n[level]; A->model_v = c->model_v[level]; A->model_n = c->model_n[level];
    A->cols = (int)c->cols[level]; A->rows = (int)c->rows[level];
    A->gx = (int)div_up(c->cols[level], 16);
    A->ntiles = A->gx * (int)div_up(c->rows[level], 16);
    A->y0 = 0; A->y1 = A->rows;
    A->cx = cam->cx; A->cy = cam->cy; A->fx = cam->fx; A->fy = cam->fy;
    A->dist_thres = dist_thres; A->sin_thres = sin_thres;
    A->angle_shake = 0.f; A->dist_shake = 0.f;
    A->partials = c->partials;
    A->is_last = 0;
}

extern "C" int hkf_cal_point_to_plane_err_solver_params(hkf_ctx* c, unsigned level, const HkfMat44* cur, const HkfMat44* last_inv,
                                                        const HkfCameraParams* cam, float dist_thres, float norm_sin_thres)
{
    if (!c || !cur || !last_inv || !cam) return HKF_ERR_ARG;
    if ((int)level >= c->levels) { c->err = "pyramid level out of range"; return HKF_ERR_ARG; }
    if (cam->cols != c->cols[level] || cam->rows != c->rows[level]) { c->err = "level camera size mismatch"; return HKF_ERR_ARG; }
    IcpArgs A; fill_icp_args(c, (int)level, cam, dist_thres, norm_sin_thres, &A);
    memcpy(A.cur, cur->m, 64); memcpy(A.last_inv, last_inv->m, 64);
    k_icp<false><<<div_up(A.ntiles, ICP_WARPS), ICP_WARPS * 32, 0, c->stream>>>(A, c->st);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

// level cameras: CameraPoseFinderICP::initPoseFinder halves cols, rows, cx, cy, fx, fy per level
// (src/CameraPoseFinderICP.cpp:36-46); iteration schedule {3} / {10,5} / {10,5,4} (:16-30)
void hkf_level_camera(const HkfCameraParams* base, int level, HkfCameraParams* out)
{
    *out = *base;
    for (int l = 0; l < level; ++l) {
        out->cols /= 2; out->rows /= 2;
        out->cx = out->cx / 2; out->cy = out->cy / 2; out->fx = out->fx / 2; out->fy = out->fy / 2;
    }
}
static const int* icp_schedule(int levels)
{
    static const int s1[3] = { 3, 0, 0 }, s2[3] = { 10, 5, 0 }, s3[3] = { 10, 5, 4 };
    return levels == 1 ? s1 : levels == 2 ? s2 : s3;
}

extern "C" int hkf_icp_track(hkf_ctx* c, const HkfFrameParams* fp)
{
    if (!c || !fp) return HKF_ERR_ARG;
    // cudaDownSampleModelVertices/Normals (the new-frame pyramids were built by hkf_preprocess)
    int rc = hkf_pyramids_fused(c, c->model_v, c->model_n);
    if (rc != HKF_OK) return rc;
    k_track_begin<<<1, 32, 0, c->stream>>>(c->st);
    HKF_LAUNCH_CHECK(c);
    const int* iters = icp_schedule(c->levels);
    for (int l = c->levels - 1; l >= 0; --l) {
        HkfCameraParams cam; hkf_level_camera(&c->ip.depth_camera, l, &cam);
        IcpArgs A; fill_icp_args(c, l, &cam, fp->icp_dist_thres, fp->icp_norm_sin_thres, &A);
        A.angle_shake = fp->icp_angle_shake; A.dist_shake = fp->icp_dist_shake;
        for (int it = 0; it < iters[l]; ++it) {
            A.is_last = (l == 0 && it == iters[l] - 1) ? 1 : 0;
            k_icp<true><<<div_up(A.ntiles, ICP_WARPS), ICP_WARPS * 32, 0, c->stream>>>(A, c->st);
            HKF_LAUNCH_CHECK(c);
        }
    }
    return HKF_OK;
}
