// sdf_track.cu — direct SDF tracking: per depth pixel a numeric SDF Jacobian row from 13 trilinear TSDF
// reads, the 27 sums, and (fused mode) the Gauss-Newton update on the device.
//
// Replaces computeSDFSolverbufKernel / buildSDFSolverRows / cudaCalSDFSolverParams
// (src/cuda/CalSDFErrSolverParams.cu:7-138) and, in fused mode, the host loop of CameraPoseFinderSDF
// (src/CameraPoseFinderSDF.cpp:44-106) including the se(3) exponential map
// (src/utils/eigen_utils.cpp:84-127).
//
// Same thread-per-pixel tiles and bit-identical reduction trees as icp.cu.  The 104 voxel reads per pixel hit
// packed 8-byte voxels (two sectors per trilinear corner pair instead of the reference's 12-byte AoS
// stride) and all eight loads of an evaluation are issued together.  The six rotation-perturbed
// transforms, which the reference builds on the host every iteration (:118-133), are built once per
// block in shared memory with the host's unfused arithmetic.
//
// Arithmetic contract (PTX):  p = (((float)x - cx) * d) / fx ;  T*(p,1) = fma(pz,m3, fma(px,m1, py*m2)) + m4 ;
// row[k] = (s+ - s-) / (h + h) with h = 0.001f (rotations) or S/R (translations); row[6] = sdf0;
// all 13 samples must be valid.
#include "tsdf_sample.cuh"
#include "hkf_solver.h"
#include "reduce27.cuh"

struct SdfArgs {
    VolView V;
    const float* depth;        // truncated depth
    int cols, rows, gx, ntiles;
    float cx, cy, fx, fy;
    float angle_shake, dist_shake;
    float* partials;
    int is_last;
    float cur[16];             // only used when !FROM_STATE
};

__device__ __forceinline__ float row_p1(const float* m, float px, float py, float pz)
{
    return __fadd_rn(__fmaf_rn(pz, m[2], __fmaf_rn(px, m[0], __fmul_rn(py, m[1]))), m[3]);
}

__device__ __forceinline__ void sdf_row(const SdfArgs& A, const float* cur, const float (*T)[16], int x, int y, float* r)
{
#pragma unroll
    for (int k = 0; k < 7; ++k) r[k] = 0.f;
    if (x >= A.cols || y >= A.rows) return;
    const float d = __ldg(A.depth + (size_t)y * A.cols + x);
    if (d == 0.f) return;
    const float px = __fdiv_rn(__fmul_rn(__fsub_rn((float)(unsigned)x, A.cx), d), A.fx);
    const float py = __fdiv_rn(__fmul_rn(__fsub_rn((float)(unsigned)y, A.cy), d), A.fy);
    const float pz = d;
    const VolView& V = A.V;
    const float wx = row_p1(cur + 0, px, py, pz), wy = row_p1(cur + 4, px, py, pz), wz = row_p1(cur + 8, px, py, pz);
    bool ok = true;
    float sdf0, sr[6], sv[6];
    ok &= (vol_interp_sdf<false>(V, wx, wy, wz, sdf0) == 1);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float rx = row_p1(T[k] + 0, px, py, pz), ry = row_p1(T[k] + 4, px, py, pz), rz = row_p1(T[k] + 8, px, py, pz);
        ok &= (vol_interp_sdf<false>(V, rx, ry, rz, sr[k]) == 1);
    }
    const float vh = V.cell;   // v_h = size.x / resolution.x
    ok &= (vol_interp_sdf<false>(V, __fadd_rn(wx, vh), wy, wz, sv[0]) == 1);
    ok &= (vol_interp_sdf<false>(V, __fsub_rn(wx, vh), wy, wz, sv[1]) == 1);
    ok &= (vol_interp_sdf<false>(V, wx, __fadd_rn(wy, vh), wz, sv[2]) == 1);
    ok &= (vol_interp_sdf<false>(V, wx, __fsub_rn(wy, vh), wz, sv[3]) == 1);
    ok &= (vol_interp_sdf<false>(V, wx, wy, __fadd_rn(wz, vh), sv[4]) == 1);
    ok &= (vol_interp_sdf<false>(V, wx, wy, __fsub_rn(wz, vh), sv[5]) == 1);
    if (!ok) return;
    const float wh = 0.001f;
    const float two_wh = __fadd_rn(wh, wh), two_vh = __fadd_rn(vh, vh);
    r[0] = __fdiv_rn(__fsub_rn(sr[0], sr[1]), two_wh);
    r[1] = __fdiv_rn(__fsub_rn(sr[2], sr[3]), two_wh);
    r[2] = __fdiv_rn(__fsub_rn(sr[4], sr[5]), two_wh);
    r[3] = __fdiv_rn(__fsub_rn(sv[0], sv[1]), two_vh);
    r[4] = __fdiv_rn(__fsub_rn(sv[2], sv[3]), two_vh);
    r[5] = __fdiv_rn(__fsub_rn(sv[4], sv[5]), two_vh);
    r[6] = sdf0;
}

#define SDF_TILES 2
#define SDF_THREADS (SDF_TILES * 256)
#define SDF_WARPS (SDF_THREADS / 32)

// One thread per depth pixel, two 16x16 tiles per block (the structure of icp.cu): the 13 trilinear evaluations of
// a pixel are independent gathers, so thread-per-pixel puts 8x more of them in flight than the earlier
// warp-per-tile mapping; the 7-float rows go through shared memory and the 27 sums are folded with the reference's
// trees exactly as in the ICP kernel.  Launches are chained with programmatic dependent launch.
template <bool FROM_STATE>
__global__ void __launch_bounds__(SDF_THREADS, 2)
k_sdf(SdfArgs A, HkfDeviceState* st)
{
    __shared__ float s_r[SDF_TILES][7][256];
    __shared__ float s_T[6][16];
    __shared__ float s_cur[16];
    __shared__ float s_sums[HKF_SUMS];
    __shared__ int s_flag[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tl = threadIdx.x >> 8, t = threadIdx.x & 255;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x < 16) s_cur[threadIdx.x] = FROM_STATE ? __ldcg(&st->cur[threadIdx.x]) : A.cur[threadIdx.x];
    if (threadIdx.x == 32) s_flag[0] = FROM_STATE ? (__ldcg(&st->track_status) == HKF_TRACK_OK && !__ldcg(&st->sdf_done)) : 1;
    __syncthreads();
    if (!s_flag[0]) return;
    if (threadIdx.x < 6) {
        // delta = I with one antisymmetric pair set to -/+ w_h, T = delta * cur (host Mat44 product, no FMA)
        const int neg[6] = { 6, 9, 8, 2, 1, 4 }, pos[6] = { 9, 6, 2, 8, 4, 1 };
        float dlt[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) dlt[i] = (i % 5 == 0) ? 1.f : 0.f;
        dlt[neg[threadIdx.x]] = -0.001f; dlt[pos[threadIdx.x]] = 0.001f;
        float out[16];
        hkf_mat44_mul(dlt, s_cur, out);
#pragma unroll
        for (int i = 0; i < 16; ++i) s_T[threadIdx.x][i] = out[i];
    }
    __syncthreads();

    {
        const int tile = blockIdx.x * SDF_TILES + tl;
        const int bx = tile % A.gx, by = tile / A.gx;
        float rr[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) rr[k] = 0.f;
        // the reference has no image bounds guard (fine for multiples of 16); out-of-image pixels add 0
        if (tile < A.ntiles) sdf_row(A, s_cur, s_T, bx * 16 + (t & 15), by * 16 + (t >> 4), rr);
#pragma unroll
        for (int k = 0; k < 7; ++k) s_r[tl][k][t] = rr[k];
    }
    __syncthreads();
#pragma unroll 1
    for (int job = warp; job < SDF_TILES * HKF_SUMS; job += SDF_WARPS) {
        const int jt = job / HKF_SUMS, k = job - jt * HKF_SUMS;
        const int tile2 = blockIdx.x * SDF_TILES + jt;
        if (tile2 >= A.ntiles) break;
        const float c = tile_sum_from_rows(s_r[jt], k, lane);
        if (lane == 0) __stcg(A.partials + (size_t)k * A.ntiles + tile2, c);
    }

    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tk = atomicAdd(&st->ticket, 1u);
        s_flag[1] = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_flag[1]) return;
    __threadfence();
    second_stage_all<SDF_WARPS>(A.partials, A.ntiles, warp, lane, s_sums);
    __syncthreads();
    if (threadIdx.x < HKF_SUMS) st->sums[threadIdx.x] = s_sums[threadIdx.x];
    if (threadIdx.x == 0) {
        st->ticket = 0;
        if (FROM_STATE) {
            float c2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) c2[i] = s_cur[i];
            int converged = 0;
            const int rc = hkf_sdf_update(s_sums, A.angle_shake, A.dist_shake, c2, &converged);
            if (rc != HKF_SOLVE_OK) {
                st->track_status = HKF_TRACK_SHAKE;
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) st->cur[i] = c2[i];
                if (converged) st->sdf_done = 1;
                if (converged || A.is_last) {              // _pose = cur_transform (CameraPoseFinderSDF.cpp:103)
                    float inv[16];
                    hkf_mat44_inverse(c2, inv);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { st->pose[i] = c2[i]; st->pose_inv[i] = inv[i]; }
                }
            }
        }
    }
}

__global__ void k_track_begin_sdf(HkfDeviceState* st)
{
    if (threadIdx.x < 16) st->cur[threadIdx.x] = st->pose[threadIdx.x];
    if (threadIdx.x == 0) { st->track_status = HKF_TRACK_OK; st->ticket = 0; st->sdf_done = 0; }
}

static int fill_sdf_args(hkf_ctx* c, const HkfCameraParams* cam, SdfArgs* A)
{
    if (!(c->z_lo == 0 && c->z_hi == c->ip.volume_resolution)) { c->err = "SDF tracking needs the whole volume on this context"; return HKF_ERR_STATE; }
    A->V = make_volview(c);
    A->depth = c->trunc_depth;
    A->cols = (int)c->cols[0]; A->rows = (int)c->rows[0];
    A->gx = (int)div_up(c->cols[0], 16);
    A->ntiles = A->gx * (int)div_up(c->rows[0], 16);
    A->cx = cam->cx; A->cy = cam->cy; A->fx = cam->fx; A->fy = cam->fy;
    A->angle_shake = 0.f; A->dist_shake = 0.f;
    A->partials = c->partials;
    A->is_last = 0;
    return HKF_OK;
}

extern "C" int hkf_cal_sdf_solver_params(hkf_ctx* c, const HkfCameraParams* cam, const HkfMat44* cur)
{
    if (!c || !cam || !cur) return HKF_ERR_ARG;
    if (cam->cols != c->cols[0] || cam->rows != c->rows[0]) { c->err = "camera size mismatch"; return HKF_ERR_ARG; }
    SdfArgs A; int rc = fill_sdf_args(c, cam, &A); if (rc) return rc;
    memcpy(A.cur, cur->m, 64);
    k_sdf<false><<<div_up(A.ntiles, SDF_TILES), SDF_THREADS, 0, c->stream>>>(A, c->st);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

extern "C" int hkf_sdf_track(hkf_ctx* c, const HkfFrameParams* fp)
{
    if (!c || !fp) return HKF_ERR_ARG;
    SdfArgs A; int rc = fill_sdf_args(c, &c->ip.depth_camera, &A); if (rc) return rc;
    A.angle_shake = fp->sdf_angle_shake; A.dist_shake = fp->sdf_dist_shake;
    k_track_begin_sdf<<<1, 32, 0, c->stream>>>(c->st);
    HKF_LAUNCH_CHECK(c);
    if (fp->sdf_max_iter == 0) return HKF_OK;          // while (iter < 0) never runs: pose unchanged
    for (unsigned it = 0; it < fp->sdf_max_iter; ++it) {
        A.is_last = (it == fp->sdf_max_iter - 1) ? 1 : 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(div_up(A.ntiles, SDF_TILES)); cfg.blockDim = dim3(SDF_THREADS); cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_sdf<true>, A, c->st);
        HKF_LAUNCH_CHECK(c);
    }
    return HKF_OK;
}
