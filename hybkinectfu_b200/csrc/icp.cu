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
// This design (latency-bound: the <= 20 MB working set is L2-resident on B200, and a frame is a chain of
// 19 dependent iterations, so what counts is the length of the dependent chain inside one iteration):
//  * one thread per pixel, two of the reference's 16x16 pixel tiles at a time per block.  The pixel's new-frame
//    vertex/normal are loaded BEFORE the grid dependency wait (programmatic dependent launch: the blocks of
//    iteration k+1 are already resident and have their coalesced loads in flight while the last block of
//    iteration k is still solving), then the pose is read, the model maps are gathered and the 7-float
//    solver row goes to shared memory (7 KB per tile);
//  * ONE barrier later the 16 warps share the 2 x 27 sums: a lane multiplies r_i r_j for the 8 reference
//    threads tid = lane + 32 m and folds them ((p0+p4)+(p2+p6))+((p1+p5)+(p3+p7)), i.e. the reference's
//    tid <-> tid+128, +64, +32 levels, followed by a __shfl_down tree for its 16/8/4/2/1 levels.  No atomics
//    on data, 1 barrier instead of 108, and the float sums are BIT-IDENTICAL to the reference's;
//  * a block owns the tiles b + 256 k, i.e. exactly the partials that the reference's second stage adds into its
//    threads b and b+256, accumulates them in that order and stores their sum: the last block to finish (one ticket
//    atomic per block) reads [27][<=256] values in one round trip and folds them with the remaining levels of the
//    reference's tree;
//  * warp 0 then runs the Cholesky solve (one lane), the Euler/shake gate and the pose update (trig and the small matrix
//    products spread over its lanes) while warp 1 evaluates the determinant gate concurrently (hkf_solver.h); 16 lanes
//    commit the new working transform.
#include "hkf_internal.cuh"
#include "hkf_solver.h"
#include "reduce27.cuh"
#include <math.h>
#include <float.h>

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
    ShardPeers peers; unsigned seq;   // peer-memory sharded mode (P2P): exchange of the partial sums inside this kernel
};

#ifdef HKF_ICP_STAMPS
__device__ __forceinline__ unsigned long long icp_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define ICP_STAMP(slot) do { if (threadIdx.x == 0) st->stamps[st->stamp_it & 31][slot] = icp_now(); } while (0)
#else
#define ICP_STAMP(slot) do { } while (0)
#endif

#define ICP_TILES 2
#define ICP_THREADS (ICP_TILES * 256)
#define ICP_WARPS (ICP_THREADS / 32)

#ifdef HKF_EMULATE_DEVICE      /* tests/emu: host build of the kernels, no launch machinery */
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// All-reduce of the 27 partial sums over peer memory (call from all threads of the block): this rank's sums go into
// slot [seq & 1][rank] of every rank's mailbox as (value, iteration) words (one 64-bit P2P store each over NVLink),
// every rank polls its own mailbox until all words of the iteration have arrived and adds them in RANK ORDER, so all
// ranks continue with bit-identical sums and poses.  *s_timeout is set when a peer's words did not arrive within ~2 s
// (the caller then ends the frame's tracking with HKF_TRACK_PEER_TIMEOUT instead of solving on garbage).
__device__ __forceinline__ void icp_p2p_allreduce(const ShardPeers& peers, unsigned seq, float* s_sums, int* s_timeout, int warp, int lane)
{
    const int par = (int)(seq & 1u);
    if (warp == 0) {
        if (lane < HKF_SUMS) {
            const unsigned long long word = ((unsigned long long)seq << 32) | __float_as_uint(s_sums[lane]);
            for (int r = 0; r < peers.nranks; ++r)
                *(volatile unsigned long long*)&peers.mail[r]->sums[par][peers.rank][lane] = word;
        }
        const ShardMailbox* mine = peers.mail[peers.rank];
        bool ok = true;
        float acc = 0.f;
        if (lane < HKF_SUMS) {
            const long long tw = clock64();
            for (int r = 0; r < peers.nranks && ok; ++r) {
                unsigned long long word;
                while ((unsigned)((word = *(volatile const unsigned long long*)&mine->sums[par][r][lane]) >> 32) != seq)
                    if (clock64() - tw > 4000000000ll) { ok = false; break; }     // ~2 s: a peer died; do not hang the GPU
                const float v = __uint_as_float((unsigned)word);
                acc = r == 0 ? v : __fadd_rn(acc, v);
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        if (lane < HKF_SUMS) s_sums[lane] = acc;
        if (lane == 0) *s_timeout = ok ? 0 : 1;
    }
    __syncthreads();
}

// findCorrs + buildPointToPlaneSolverRows for one pixel (new-frame vertex v / normal n already loaded); the 7-float
// solver row goes to shared memory, zero when there is no correspondence
template <class ARGS>
__device__ __forceinline__ void icp_pixel_row(const ARGS& A, const float* __restrict__ cur, const float* __restrict__ li,
                                              const float4 n, const float4 v, bool ok, float (*s_row)[256], int t)
{
    // findCorrs + buildPointToPlaneSolverRows; r = 0 when there is no correspondence
    ok = ok && !hkf_is_zero4(n);
    // v_input_g = cur * v (lead-first), n_input_g = cur * n (lead-second): as contracted in the reference PTX
    const float qx = hkf_row4<true>(cur[0], cur[1], cur[2], cur[3], v.x, v.y, v.z, v.w);
    const float qy = hkf_row4<true>(cur[4], cur[5], cur[6], cur[7], v.x, v.y, v.z, v.w);
    const float qz = hkf_row4<true>(cur[8], cur[9], cur[10], cur[11], v.x, v.y, v.z, v.w);
    const float qw = hkf_row4<true>(cur[12], cur[13], cur[14], cur[15], v.x, v.y, v.z, v.w);
    const float nix = hkf_row4<false>(cur[0], cur[1], cur[2], cur[3], n.x, n.y, n.z, n.w);
    const float niy = hkf_row4<false>(cur[4], cur[5], cur[6], cur[7], n.x, n.y, n.z, n.w);
    const float niz = hkf_row4<false>(cur[8], cur[9], cur[10], cur[11], n.x, n.y, n.z, n.w);
    // v_cp = last_inv * v_input_g (lead-second); projective association, no z > 0 test
    const float px = hkf_row4<false>(li[0], li[1], li[2], li[3], qx, qy, qz, qw);
    const float py = hkf_row4<false>(li[4], li[5], li[6], li[7], qx, qy, qz, qw);
    const float pz = hkf_row4<false>(li[8], li[9], li[10], li[11], qx, qy, qz, qw);
    const int u = hkf_round_px(hkf_project1(px, A.fx, pz, A.cx));
    const int w = hkf_round_px(hkf_project1(py, A.fy, pz, A.cy));
    ok = ok && !(u < 0 || u >= A.cols || w < 0 || w >= A.rows);
    const size_t midx = ok ? (size_t)w * A.cols + u : 0;
    const float4 nt = ld_f4(A.model_n + midx);
    const float4 p = ld_f4(A.model_v + midx);
    ok = ok && !hkf_is_zero4(nt);
    const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
    const float d = __fsqrt_rn(hkf_dot3(dx, dy, dz, dx, dy, dz));
    const float sx = hkf_det2(nt.y, niz, nt.z, niy);
    const float sy = hkf_det2(nt.z, nix, nt.x, niz);
    const float sz = hkf_det2(nt.x, niy, nt.y, nix);
    const float sn = __fsqrt_rn(hkf_dot3(sx, sy, sz, sx, sy, sz));
    ok = ok && !(d > A.dist_thres || sn > A.sin_thres);
    s_row[0][t] = ok ? hkf_det2(qy, nt.z, qz, nt.y) : 0.f;
    s_row[1][t] = ok ? hkf_det2(qz, nt.x, qx, nt.z) : 0.f;
    s_row[2][t] = ok ? hkf_det2(qx, nt.y, qy, nt.x) : 0.f;
    s_row[3][t] = ok ? nt.x : 0.f;
    s_row[4][t] = ok ? nt.y : 0.f;
    s_row[5][t] = ok ? nt.z : 0.f;
    s_row[6][t] = ok ? hkf_dot3(nt.x, nt.y, nt.z, dx, dy, dz) : 0.f;
}

// The solve of one iteration (hkf_icp_step, operation for operation) by the calling block: warp 0 solves while warp 1
// evaluates the determinant gate - Cholesky on lane 0, the three sin/cos pairs on lanes 0-2, the 3x3 products on lanes
// 0-8, the shake gate on lane 0 next to the 4x4 product T * cur on lanes 16-31.  Results: s_c2 = T * cur,
// s_flag[2] = singular, s_flag[3] = HKF_SOLVE_*; the caller's next __syncthreads() publishes them.
// The Cholesky solve stays on ONE lane: its divisions and square roots form a dependent chain that the lane's own
// instruction-level parallelism already overlaps column by column.  Spreading the rows over lanes (shuffles for the
// pivot row, bit-identical results) was measured SLOWER on B200: 3.0 instead of 2.5 us per iteration
// (profiles/r2_icp_stamps_*.txt) - the shuffle latency lands on the critical path.
__device__ __forceinline__ void icp_block_solve(const float* s_sums, const float* cur, float angle_shake, float dist_shake,
                                                float* s_c2, int* s_flag, int warp, int lane, unsigned long long* stamp = nullptr)
{
    __shared__ float s_R[3][9], s_Rxy[9], s_Rf[9];
    __shared__ float s_x[6];
    if (warp == 0) {
        if (lane == 0) {
            float Am[36], bv[6], xs[6];
            hkf_unpack27(s_sums, Am, bv);
            hkf_llt_solve6(Am, bv, xs);
#pragma unroll
            for (int i = 0; i < 6; ++i) s_x[i] = xs[i];
#ifdef HKF_ICP_STAMPS
            if (stamp) stamp[6] = icp_now();          // Cholesky solve done
#endif
        }
        __syncwarp();
        float x[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = s_x[i];
        if (lane < 3) {
            float R[9];
            hkf_rot_axis(lane, lane == 0 ? x[0] : lane == 1 ? x[1] : x[2], R);
#pragma unroll
            for (int i = 0; i < 9; ++i) s_R[lane][i] = R[i];
        }
        __syncwarp();
        if (lane < 9) {
            const int i = lane / 3, j = lane - 3 * i;
            s_Rxy[lane] = __fadd_rn(__fadd_rn(__fmul_rn(s_R[0][i * 3 + 0], s_R[1][0 * 3 + j]), __fmul_rn(s_R[0][i * 3 + 1], s_R[1][1 * 3 + j])),
                                    __fmul_rn(s_R[0][i * 3 + 2], s_R[1][2 * 3 + j]));
        }
        __syncwarp();
        if (lane < 9) {
            const int i = lane / 3, j = lane - 3 * i;
            s_Rf[lane] = __fadd_rn(__fadd_rn(__fmul_rn(s_Rxy[i * 3 + 0], s_R[2][0 * 3 + j]), __fmul_rn(s_Rxy[i * 3 + 1], s_R[2][1 * 3 + j])),
                                   __fmul_rn(s_Rxy[i * 3 + 2], s_R[2][2 * 3 + j]));
        }
        __syncwarp();
#ifdef HKF_ICP_STAMPS
        if (stamp && lane == 0) stamp[7] = icp_now();   // rotation assembled
#endif
        if (lane == 0) {
            // the shake gate of hkf_vector6_to_transform (a NaN passes, as in the reference)
            const float angle = hkf_rotation_angle(s_Rf);
            const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x[3], x[3]), __fmul_rn(x[4], x[4])), __fmul_rn(x[5], x[5])));
            s_flag[3] = (angle > angle_shake || d > dist_shake) ? HKF_SOLVE_SHAKE : HKF_SOLVE_OK;
        } else if (lane >= 16) {
            // cur <- T * cur, T = [R | x[3..5]; 0 0 0 1] (hkf_mat44_mul, entry by entry)
            const int e = lane - 16, i = e >> 2, j = e & 3;
            const float t0r = i < 3 ? s_Rf[i * 3 + 0] : 0.f, t1r = i < 3 ? s_Rf[i * 3 + 1] : 0.f, t2r = i < 3 ? s_Rf[i * 3 + 2] : 0.f;
            const float t3r = i == 0 ? x[3] : i == 1 ? x[4] : i == 2 ? x[5] : 1.f;
            float acc = __fadd_rn(__fmul_rn(t0r, cur[0 * 4 + j]), __fmul_rn(t1r, cur[1 * 4 + j]));
            acc = __fadd_rn(acc, __fmul_rn(t2r, cur[2 * 4 + j]));
            acc = __fadd_rn(acc, __fmul_rn(t3r, cur[3 * 4 + j]));
            s_c2[e] = acc;
        }
    } else if (warp == 1 && lane == 0) {
        s_flag[2] = hkf_icp_singular(s_sums) ? 1 : 0;
    }
}

// FROM_STATE: pose/status from device state; SOLVE: the last block also solves and advances the pose (single GPU).
// Sharded ranks without peer memory run <true,false> over their pixel-row band, allreduce st->sums, then k_icp_solve.
//
// Work mapping.  The reference's second stage gives its thread t (of 512) the block partials t, t+512, t+1024, ...
// ("leaf" t), then folds the 512 leaves with a fixed tree whose first level adds leaf t and leaf t+256.  Here block b
// (of min(256, tiles)) owns exactly the tiles of leaves b and b+256: tiles b + 256 k, two per round (thread group 0
// takes the even k = leaf b's terms, group 1 the odd k = leaf b+256's terms), accumulates the two leaf sums in the
// reference's order and adds them, so that only [27][<=256] values reach the second stage: 16 loads per lane in ONE
// round trip for the last block instead of 2 x 48.
template <bool FROM_STATE, bool SOLVE, bool P2P>
__global__ void __launch_bounds__(ICP_THREADS, 2)
k_icp(IcpArgs A, HkfDeviceState* st)
{
    __shared__ float s_r[ICP_TILES][7][256];
    __shared__ float s_m[32];                  // cur | last_inv
    __shared__ float s_acc[2 * HKF_SUMS];      // leaf sums of leaf b | leaf b+256
    __shared__ float s_sums[HKF_SUMS];
    __shared__ int s_flag[5];                  // 0: tracking live, 1: last block, 2: singular, 3: step result, 4: peer timeout
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tl = threadIdx.x >> 8, t = threadIdx.x & 255;
#ifdef HKF_ICP_STAMPS
    const unsigned long long t_entry = icp_now();
#endif
    pdl_launch_dependents();
    const int nblocks = (int)gridDim.x;                                   // = min(256, ntiles)
    const int rounds = (A.ntiles - (int)blockIdx.x + 2 * 256 - 1) / (2 * 256);   // tiles b + 256 k < ntiles, two k per round

    // ---- independent of the previous iteration: this thread's pixel of the new frame (first round)
    int tile = (int)blockIdx.x + 256 * tl;
    float4 n, v;
    bool ok;
    auto load_pixel = [&]() {
        const int bx = tile % A.gx, by = tile / A.gx;
        const int x = bx * 16 + (t & 15), y = by * 16 + (t >> 4);
        ok = tile < A.ntiles && !(x >= A.cols || y >= A.rows || y < A.y0 || y >= A.y1);
        const size_t idx = ok ? (size_t)y * A.cols + x : 0;
        n = ld_f4(A.new_n + idx);
        v = ld_f4(A.new_v + idx);
    };
    load_pixel();

    // ---- everything below depends on the previous launch (pose, status, ticket, partials)
#ifdef HKF_ICP_STAMPS
    const unsigned long long t_loaded = icp_now();
#endif
    pdl_wait();
#ifdef HKF_ICP_STAMPS
    const unsigned long long t_waited = icp_now();
#endif
    if (FROM_STATE) {
        if (threadIdx.x < 32) s_m[threadIdx.x] = __ldcg(threadIdx.x < 16 ? &st->cur[threadIdx.x] : &st->last_inv[threadIdx.x - 16]);
        if (threadIdx.x == 32) s_flag[0] = (__ldcg(&st->track_status) == HKF_TRACK_OK);
    } else {
        if (threadIdx.x < 32) s_m[threadIdx.x] = threadIdx.x < 16 ? A.cur[threadIdx.x] : A.last_inv[threadIdx.x - 16];
        if (threadIdx.x == 32) s_flag[0] = 1;
    }
    __syncthreads();
    if (!s_flag[0]) return;                    // a failed iteration ends the frame's tracking
    const float* cur = s_m; const float* li = s_m + 16;

    // job j = group * 27 + sum, owned by warp j % 16 in accumulator slot j / 16 (lane 0 holds the value)
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
        if (r) { tile = (int)blockIdx.x + 256 * (2 * r + tl); load_pixel(); }
        icp_pixel_row(A, cur, li, n, v, ok, s_r[tl], t);
        __syncthreads();
#pragma unroll
        for (int slot = 0; slot < 4; ++slot) {
            const int job = warp + ICP_WARPS * slot;
            if (job < 2 * HKF_SUMS) {                                     // warp-uniform
                const int g = job / HKF_SUMS, k = job - g * HKF_SUMS;
                if ((int)blockIdx.x + 256 * (2 * r + g) < A.ntiles)       // the reference adds only existing partials
                    acc[slot] = __fadd_rn(acc[slot], tile_sum_from_rows(s_r[g], k, lane));
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int slot = 0; slot < 4; ++slot) {
        const int job = warp + ICP_WARPS * slot;
        if (job < 2 * HKF_SUMS && lane == 0) s_acc[job] = acc[slot];
    }
    __syncthreads();
    // first level of the reference's second-stage tree: leaf b + leaf b+256
    if (threadIdx.x < HKF_SUMS)
        __stcg(A.partials + (size_t)threadIdx.x * 256 + blockIdx.x, __fadd_rn(s_acc[threadIdx.x], s_acc[HKF_SUMS + threadIdx.x]));

    // ---- last block folds the leaf pairs and (fused mode) advances the pose.  The 27 stores above all belong to
    // warp 0: it alone orders them before the ticket (warp barrier, then one fence + atomic by lane 0)
    if (warp == 0) {
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            const unsigned tk = atomicAdd(&st->ticket, 1u);
            s_flag[1] = (tk == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_flag[1]) return;
#ifdef HKF_ICP_STAMPS
    if (threadIdx.x == 0) { st->stamps[st->stamp_it & 31][0] = t_entry; st->stamps[st->stamp_it & 31][1] = t_loaded; st->stamps[st->stamp_it & 31][2] = t_waited; }
#endif
    ICP_STAMP(3);
    __threadfence();
    {
        // rows 2 per warp, all 16 loads in flight; lane owns entries lane + 32 m, the remaining levels 128/64/32 and
        // 16..1 are the tile tree
        float pv[2][8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = warp + ICP_WARPS * q;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int e = lane + 32 * m;
                pv[q][m] = (k < HKF_SUMS && e < nblocks) ? __ldcg(A.partials + (size_t)k * 256 + e) : 0.f;
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = warp + ICP_WARPS * q;
            const float c = tile_tree(pv[q]);
            if (k < HKF_SUMS && lane == 0) s_sums[k] = c;
        }
    }
    __syncthreads();
    if (P2P) {
        icp_p2p_allreduce(A.peers, A.seq, s_sums, &s_flag[4], warp, lane);
        if (s_flag[4]) {                           // a peer never delivered: end the frame's tracking, do not solve on garbage
            if (threadIdx.x == 0) { st->track_status = HKF_TRACK_PEER_TIMEOUT; st->ticket = 0; }
            return;
        }
    }
    ICP_STAMP(4);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + HKF_SUMS) st->sums[threadIdx.x - 64] = s_sums[threadIdx.x - 64];
    if (threadIdx.x == 64) st->ticket = 0;
    if (!SOLVE) return;
    __shared__ float s_c2[16];
    icp_block_solve(s_sums, cur, A.angle_shake, A.dist_shake, s_c2, s_flag, warp, lane);
    ICP_STAMP(5);
    __syncthreads();
    ICP_STAMP(6);
    // det(AtA) < 1e-10 is tested first (CameraPoseFinderICP.cpp:138), then the shake gate
    const int status = s_flag[2] ? HKF_TRACK_SINGULAR : s_flag[3] != HKF_SOLVE_OK ? HKF_TRACK_SHAKE : HKF_TRACK_OK;
    if (status != HKF_TRACK_OK) {
        if (threadIdx.x == 0) st->track_status = status;
        return;
    }
    if (threadIdx.x < 16) {
        st->cur[threadIdx.x] = s_c2[threadIdx.x];
        if (A.is_last) st->pose[threadIdx.x] = s_c2[threadIdx.x];      // _pose = cur_transform (CameraPoseFinderICP.cpp:84)
    }
    if (A.is_last && threadIdx.x == 32) {
        float c2[16], inv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c2[i] = s_c2[i];
        hkf_mat44_inverse(c2, inv);
#pragma unroll
        for (int i = 0; i < 16; ++i) st->pose_inv[i] = inv[i];
    }
#ifdef HKF_ICP_STAMPS
    __syncthreads();
    if (threadIdx.x == 0) { st->stamps[st->stamp_it & 31][7] = icp_now(); st->stamp_it++; }
#endif
}

// start of a frame's tracking: cur = pose, last_inv = pose^-1 (CameraPoseFinderICP.cpp:62-63)
__global__ void k_track_begin(HkfDeviceState* st)
{
    hkf_pdl_enter_ordered();      // k_icp follows and preloads the new frame's maps before its own wait
    if (threadIdx.x < 16) {
        st->cur[threadIdx.x] = st->pose[threadIdx.x];
        st->last_inv[threadIdx.x] = st->pose_inv[threadIdx.x];
    }
    if (threadIdx.x == 0) { st->track_status = HKF_TRACK_OK; st->ticket = 0; st->gen = 0; st->sdf_done = 0; }
}

#ifndef HKF_EMULATE_DEVICE
static inline unsigned icp_grid(int ntiles) { return (unsigned)(ntiles < 256 ? ntiles : 256); }

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
    A->peers = ShardPeers{}; A->seq = 0;
}

extern "C" int hkf_cal_point_to_plane_err_solver_params(hkf_ctx* c, unsigned level, const HkfMat44* cur, const HkfMat44* last_inv,
                                                        const HkfCameraParams* cam, float dist_thres, float norm_sin_thres)
{
    if (!c || !cur || !last_inv || !cam) return HKF_ERR_ARG;
    if ((int)level >= c->levels) { c->err = "pyramid level out of range"; return HKF_ERR_ARG; }
    if (cam->cols != c->cols[level] || cam->rows != c->rows[level]) { c->err = "level camera size mismatch"; return HKF_ERR_ARG; }
    IcpArgs A; fill_icp_args(c, (int)level, cam, dist_thres, norm_sin_thres, &A);
    memcpy(A.cur, cur->m, 64); memcpy(A.last_inv, last_inv->m, 64);
    k_icp<false, false, false><<<icp_grid(A.ntiles), ICP_THREADS, 0, c->stream>>>(A, c->st);
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
// sqrt.rn(x) > t  <=>  x > icp_sq_threshold(t) for every x >= 0 (and for NaN: both false): the largest float whose
// correctly rounded square root does not exceed t.  sqrtf on the host is correctly rounded, as __fsqrt_rn is.
static float icp_sq_threshold(float t)
{
    if (t != t) return INFINITY;                       // x > NaN is false for every x
    if (t < 0.f) return -INFINITY;                     // sqrt(x) >= 0 > t for every x >= 0
    if (t == INFINITY) return INFINITY;
    float x = t * t;
    if (x == INFINITY) x = FLT_MAX;
    while (x < FLT_MAX && sqrtf(nextafterf(x, INFINITY)) <= t) x = nextafterf(x, INFINITY);
    while (x > 0.f && sqrtf(x) > t) x = nextafterf(x, -INFINITY);
    return x;
}
static const int* icp_schedule(int levels)
{
    static const int s1[3] = { 3, 0, 0 }, s2[3] = { 10, 5, 0 }, s3[3] = { 10, 5, 4 };
    return levels == 1 ? s1 : levels == 2 ? s2 : s3;
}

#endif // HKF_EMULATE_DEVICE

// sharded mode: the solve of one iteration on the allreduced sums (every rank computes the same bits)
__global__ void k_icp_solve(HkfDeviceState* st, float angle_shake, float dist_shake, int is_last)
{
    __shared__ float s_sums[HKF_SUMS];
    __shared__ int s_flag[2];
    if (st->track_status != HKF_TRACK_OK) return;
    if (threadIdx.x < HKF_SUMS) s_sums[threadIdx.x] = st->sums[threadIdx.x];
    __syncthreads();
    float c2[16];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c2[i] = st->cur[i];
        s_flag[1] = hkf_icp_step(s_sums, angle_shake, dist_shake, c2);
    } else if (threadIdx.x == 32) {
        s_flag[0] = hkf_icp_singular(s_sums) ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_flag[0]) st->track_status = HKF_TRACK_SINGULAR;
        else if (s_flag[1] != HKF_SOLVE_OK) st->track_status = HKF_TRACK_SHAKE;
        else {
#pragma unroll
            for (int i = 0; i < 16; ++i) st->cur[i] = c2[i];
            if (is_last) {
                float inv[16];
                hkf_mat44_inverse(c2, inv);
#pragma unroll
                for (int i = 0; i < 16; ++i) { st->pose[i] = c2[i]; st->pose_inv[i] = inv[i]; }
            }
        }
    }
}


// ---- the tile phase of k_icp_frame: same results as icp_pixel_row + tile_sum_from_rows, fewer instructions (the phase
// is issue-bound: profiles/r2_icp_frame_a_ncu.txt, 510 warp instructions per warp and round before, ~330 now).
//  * thread T of a tile works on the reference thread tid = ((T & 127) >> 2) + 32 (4 (T >> 7) + (T & 3)) and stores its
//    solver row at index T: the eight values lane + 32 m that a lane of the reduction multiplies are then two
//    conflict-free 128-bit shared-memory loads per row instead of eight 32-bit ones;
//  * a warp reduces FOUR sums at once and shares the five shuffle levels between them (lanes 16-31 take over two of the
//    four sums at level 16, lanes with bit 3 one of the remaining two at level 8): every addition has the operands of
//    the reference's tree (a + b = b + a bit for bit), 6 shuffles instead of 20;
//  * the pixel coordinate comes from one MUFU.RCP and the 1.5 * 2^23 rounding constant; only coordinates that are not
//    clearly positive, not below `lim` or within lim * 2^-20 px of a rounding boundary take the reference's
//    mul / div.rn / add / f64 add / cvt.rzi sequence (|exact - approximate| <= lim * 2^-21 px);
//  * sqrt.rn(x) > t  <=>  x > t2 with t2 = the largest float whose correctly rounded root is <= t (computed on the
//    host): the two IEEE square roots of the distance / normal gates are not evaluated;
//  * the fourth row of an affine transform is (0,0,0,1): q.w = v.w, and n.w = 0 contributes +-0 to sums that are only
//    squared afterwards.
struct IcpPixArgs {
    const float4* model_v; const float4* model_n;
    int cols, rows;
    float cx, cy, fx, fy;
    float dist2, sin2;        // squared-domain gates (icp_sq_threshold)
    float lim, guard;         // approximate projection: trusted for 1 <= coordinate < lim, |fraction - .5| > guard
};

__device__ __forceinline__ void icp_pixel_row_fast(const IcpPixArgs& A, const float* __restrict__ cur, const float* __restrict__ li,
                                                   const float4 n, const float4 v, bool ok, float (*s_row)[256], int t)
{
    ok = ok && !hkf_is_zero4(n);
    const float qx = hkf_row4<true>(cur[0], cur[1], cur[2], cur[3], v.x, v.y, v.z, v.w);
    const float qy = hkf_row4<true>(cur[4], cur[5], cur[6], cur[7], v.x, v.y, v.z, v.w);
    const float qz = hkf_row4<true>(cur[8], cur[9], cur[10], cur[11], v.x, v.y, v.z, v.w);
    const float qw = v.w;
    // n.w = 0: the fourth term is +-0 (the sign of a zero n_input component cannot reach a result)
    const float nix = __fmaf_rn(n.z, cur[2], __fmaf_rn(n.x, cur[0], __fmul_rn(n.y, cur[1])));
    const float niy = __fmaf_rn(n.z, cur[6], __fmaf_rn(n.x, cur[4], __fmul_rn(n.y, cur[5])));
    const float niz = __fmaf_rn(n.z, cur[10], __fmaf_rn(n.x, cur[8], __fmul_rn(n.y, cur[9])));
    const float px = hkf_row4<false>(li[0], li[1], li[2], li[3], qx, qy, qz, qw);
    const float py = hkf_row4<false>(li[4], li[5], li[6], li[7], qx, qy, qz, qw);
    const float pz = hkf_row4<false>(li[8], li[9], li[10], li[11], qx, qy, qz, qw);
    const float MAGIC = 12582912.0f;
    const float mu = __fmul_rn(px, A.fx), mv = __fmul_rn(py, A.fy);
    const float rz = hkf_rcp_approx(pz);
    const float su = __fmaf_rn(mu, rz, A.cx), sv = __fmaf_rn(mv, rz, A.cy);
    const float tu = __fadd_rn(su, MAGIC), tv = __fadd_rn(sv, MAGIC);
    int u = __float_as_int(tu) - 0x4B400000, w = __float_as_int(tv) - 0x4B400000;
    const float eu = __fsub_rn(su, __fsub_rn(tu, MAGIC)), ev = __fsub_rn(sv, __fsub_rn(tv, MAGIC));
    if (!(fminf(su, sv) >= 1.0f && fmaxf(su, sv) < A.lim && fmaxf(fabsf(eu), fabsf(ev)) < 0.5f - A.guard)) {   // NaN-safe
        u = hkf_round_px(__fadd_rn(__fdiv_rn(mu, pz), A.cx));
        w = hkf_round_px(__fadd_rn(__fdiv_rn(mv, pz), A.cy));
    }
    ok = ok && !(u < 0 || u >= A.cols || w < 0 || w >= A.rows);
    const unsigned midx = ok ? (unsigned)w * (unsigned)A.cols + (unsigned)u : 0u;
    const float4 nt = ld_f4(A.model_n + midx);
    const float4 p = ld_f4(A.model_v + midx);
    ok = ok && !hkf_is_zero4(nt);
    const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
    const float dd = hkf_dot3(dx, dy, dz, dx, dy, dz);
    const float sx = hkf_det2(nt.y, niz, nt.z, niy);
    const float sy = hkf_det2(nt.z, nix, nt.x, niz);
    const float sz = hkf_det2(nt.x, niy, nt.y, nix);
    const float ss = hkf_dot3(sx, sy, sz, sx, sy, sz);
    ok = ok && !(dd > A.dist2 || ss > A.sin2);
    s_row[0][t] = ok ? hkf_det2(qy, nt.z, qz, nt.y) : 0.f;
    s_row[1][t] = ok ? hkf_det2(qz, nt.x, qx, nt.z) : 0.f;
    s_row[2][t] = ok ? hkf_det2(qx, nt.y, qy, nt.x) : 0.f;
    s_row[3][t] = ok ? nt.x : 0.f;
    s_row[4][t] = ok ? nt.y : 0.f;
    s_row[5][t] = ok ? nt.z : 0.f;
    s_row[6][t] = ok ? hkf_dot3(nt.x, nt.y, nt.z, dx, dy, dz) : 0.f;
}

// products + the tile tree for the sums k0 .. k0+3 of one tile (rows stored by icp_pixel_row_fast); the total of sum
// k0 + q ends up in lanes 8 q .. 8 q + 7.  Sums >= 27 are computed on row pair (6,6) and ignored by the caller.
__device__ __forceinline__ float icp_tile_sums4(const float (*s_r)[256], int k0, int lane)
{
    float c[4];
    int ip = -1;
    float4 a0, a1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = min(k0 + q, HKF_SUMS - 1);
        const int i = c_sum_i[k], j = c_sum_j[k];
        if (i != ip) {                                    // warp-uniform; consecutive sums mostly share row i
            a0 = reinterpret_cast<const float4*>(s_r[i])[lane]; a1 = reinterpret_cast<const float4*>(s_r[i])[32 + lane];
            ip = i;
        }
        const float4 b0 = reinterpret_cast<const float4*>(s_r[j])[lane], b1 = reinterpret_cast<const float4*>(s_r[j])[32 + lane];
        const float p0 = __fmul_rn(a0.x, b0.x), p1 = __fmul_rn(a0.y, b0.y), p2 = __fmul_rn(a0.z, b0.z), p3 = __fmul_rn(a0.w, b0.w);
        const float p4 = __fmul_rn(a1.x, b1.x), p5 = __fmul_rn(a1.y, b1.y), p6 = __fmul_rn(a1.z, b1.z), p7 = __fmul_rn(a1.w, b1.w);
        c[q] = __fadd_rn(__fadd_rn(__fadd_rn(p0, p4), __fadd_rn(p2, p6)), __fadd_rn(__fadd_rn(p1, p5), __fadd_rn(p3, p7)));
    }
    const bool up16 = lane & 16, up8 = lane & 8;
    const float ra = __shfl_xor_sync(0xffffffffu, up16 ? c[0] : c[2], 16);
    const float rb = __shfl_xor_sync(0xffffffffu, up16 ? c[1] : c[3], 16);
    const float a = __fadd_rn(up16 ? c[2] : c[0], ra), b = __fadd_rn(up16 ? c[3] : c[1], rb);
    const float rc = __shfl_xor_sync(0xffffffffu, up8 ? a : b, 8);
    float s = __fadd_rn(up8 ? b : a, rc);
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
    return s;
}

// ---------------------------------------------------------------- the whole frame's ICP in ONE launch
// All 19 iterations of the {10,5,4} schedule (CameraPoseFinderICP.cpp:16-30,66-83) run inside one grid of
// min(256, tiles of level 0) co-resident blocks.  Per iteration: the blocks that own tiles of the level (block b owns the
// tiles b + 256 k, as in k_icp) store their [27] leaf-pair sums and arrive on a monotonically increasing counter; EVERY
// block then waits for the arrivals, folds the [27][<=256] partials and runs the solve itself - all blocks compute the
// same bits, so the new working transform never travels: it stays in each block's shared memory.  One grid-wide
// synchronisation point per iteration instead of a kernel boundary (completion, memory flush, dependent release) plus
// the block tail of the launch chain; partials are double-buffered by iteration parity (a block can be at most one
// iteration ahead of the slowest one).  Measured: profiles/r2_icp_*.txt.
struct IcpLevel {
    const float4* new_v; const float4* new_n; const float4* model_v; const float4* model_n;
    int cols, rows, gx, ntiles;
    int y0, y1;
    float cx, cy, fx, fy;
    int iters;
    unsigned gx_magic;         // ceil(2^24 / gx) when (ntiles + 512) * gx < 2^24: tile / gx = (tile * gx_magic) >> 24; else 0
    float lim, guard;          // approximate projection (icp_pixel_row_fast)
};
struct IcpFrameArgs {
    IcpLevel lv[HKF_MAX_LEVELS];
    int levels;
    float dist2, sin2, angle_shake, dist_shake;   // dist2 / sin2: squared-domain gates (icp_sq_threshold)
    float* partials;           // [2][27][256]
    ShardPeers peers; unsigned seq0;
};

#ifndef HKF_EMULATE_DEVICE
// Arrival / wait of the per-iteration grid synchronisation.  The writer releases (membar + red, no L1 invalidation);
// the reader polls with a relaxed L2 load and takes NO acquire fence: every acquire on sm_100 carries CCTL.IVALL, which
// throws away the L1 copies of the frame's (read-only) vertex / normal maps 19 times per frame - the L1 hit rate of the
// gathers was 20 % with it.  Everything that other blocks write during the kernel (the partial sums) is read with
// ld.global.cg, i.e. from L2, the point of coherence, after the polled value has been seen there.
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void group_barrier(int g)       // the 256 threads of tile group g
{
    asm volatile("bar.sync %0, 256;" :: "r"(1 + g) : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

#ifdef HKF_ICP_STAMPS
#define ICPF_STAMP(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) st->stamps[itg & 31][slot] = icp_now(); } while (0)
#else
#define ICPF_STAMP(slot) do { } while (0)
#endif

template <bool P2P>
__global__ void __launch_bounds__(ICP_THREADS, 2)
k_icp_frame(IcpFrameArgs F, HkfDeviceState* st)
{
    __shared__ float s_r[ICP_TILES][7][256];
    __shared__ float s_m[32];                  // cur | last_inv
    __shared__ float s_acc[2 * HKF_SUMS];
    __shared__ float s_sums[HKF_SUMS];
    __shared__ float s_c2[16];
    __shared__ int s_flag[5];                  // 0: tracking live, 2: singular, 3: step result, 4: timeout
    __shared__ IcpPixArgs s_P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tl = threadIdx.x >> 8, t = threadIdx.x & 255;

    // ---- independent of the previous launch: this thread's pixel of the new frame (coarsest level, first round)
    int lvl = F.levels - 1;
    int have_tile = -1, have_lvl = -1;
    float4 n, v;
    bool ok;
    // thread t of a tile stands for the reference's thread tref (see icp_pixel_row_fast)
    const int tref = ((t & 127) >> 2) + 32 * (4 * (t >> 7) + (t & 3));
    auto load_pixel = [&](int l, int tile) {
        if (tile == have_tile && l == have_lvl) return;               // still in registers
        const IcpLevel& L = F.lv[l];
        const int by = L.gx_magic ? (int)(((unsigned long long)(unsigned)tile * L.gx_magic) >> 24) : tile / L.gx;
        const int bx = tile - by * L.gx;
        const int x = bx * 16 + (tref & 15), y = by * 16 + (tref >> 4);
        ok = tile < L.ntiles && !(x >= L.cols || y >= L.rows || y < L.y0 || y >= L.y1);
        const unsigned idx = ok ? (unsigned)y * (unsigned)L.cols + (unsigned)x : 0u;
        n = ld_f4(L.new_n + idx);
        v = ld_f4(L.new_v + idx);
        have_tile = tile; have_lvl = l;
    };
    load_pixel(lvl, (int)blockIdx.x + 256 * tl);
    pdl_wait();
    if (threadIdx.x < 32) s_m[threadIdx.x] = __ldcg(threadIdx.x < 16 ? &st->cur[threadIdx.x] : &st->last_inv[threadIdx.x - 16]);
    if (threadIdx.x == 32) { s_flag[0] = (__ldcg(&st->track_status) == HKF_TRACK_OK); s_flag[4] = 0; }
    __syncthreads();
    if (!s_flag[0]) return;
    const float* cur = s_m; const float* li = s_m + 16;

    unsigned target = 0;
    int itg = 0;                                // iteration index over the whole schedule
    for (; lvl >= 0; --lvl) {
        const IcpLevel& L = F.lv[lvl];
        if (threadIdx.x == 0) {                 // the level's pixel parameters: shared memory, not 12 registers per thread
            s_P.model_v = L.model_v; s_P.model_n = L.model_n; s_P.cols = L.cols; s_P.rows = L.rows;
            s_P.cx = L.cx; s_P.cy = L.cy; s_P.fx = L.fx; s_P.fy = L.fy; s_P.dist2 = F.dist2; s_P.sin2 = F.sin2;
            s_P.lim = L.lim; s_P.guard = L.guard;
        }
        __syncthreads();
        const IcpPixArgs& P = s_P;
        const int nb = L.ntiles < 256 ? L.ntiles : 256;              // blocks that own tiles of this level
        const bool mine = (int)blockIdx.x < nb;
        const int rounds = mine ? (L.ntiles - (int)blockIdx.x + 2 * 256 - 1) / (2 * 256) : 0;
#pragma unroll 1
        for (int it = 0; it < L.iters; ++it, ++itg) {
            const bool last = (lvl == 0 && it == L.iters - 1);
            if (last) pdl_launch_dependents();                        // the integration's item pass may become resident
            ICPF_STAMP(0);
            float* const part = F.partials + (size_t)(itg & 1) * (HKF_SUMS * 256);
            if (mine) {
                // warps 8 g .. 8 g + 6 reduce tile g (the tile their own threads wrote): sums 4 (warp & 7) .. + 3
                const int g = warp >> 3, k0 = 4 * (warp & 7);
                float acc = 0.f;
                // the 8 warps of tile group tl write and reduce their own tile: the two groups of a block only meet at the
                // block barrier after the loop (named barriers 1 + tl, 256 threads), so that one group's shared-memory
                // bound reduction can overlap the other's loads.  The next round's pixel is requested before the
                // reduction of this one.
                load_pixel(lvl, (int)blockIdx.x + 256 * tl);
#pragma unroll 1
                for (int r = 0; r < rounds; ++r) {
                    icp_pixel_row_fast(P, cur, li, n, v, ok, s_r[tl], t);
                    if (r + 1 < rounds) load_pixel(lvl, (int)blockIdx.x + 256 * (2 * (r + 1) + tl));
                    group_barrier(tl);
                    if (k0 < HKF_SUMS && (int)blockIdx.x + 256 * (2 * r + g) < L.ntiles)     // the reference adds only existing partials
                        acc = __fadd_rn(acc, icp_tile_sums4(s_r[g], k0, lane));
                    group_barrier(tl);
                }
                if ((lane & 7) == 0 && k0 + (lane >> 3) < HKF_SUMS) s_acc[g * HKF_SUMS + k0 + (lane >> 3)] = acc;
                // the first pixel of the NEXT iteration (same level, or the next finer one) travels during the wait and the solve
                if (it + 1 < L.iters) load_pixel(lvl, (int)blockIdx.x + 256 * tl);
                else if (lvl > 0) load_pixel(lvl - 1, (int)blockIdx.x + 256 * tl);
                __syncthreads();
                if (warp == 0) {
                    if (lane < HKF_SUMS) __stcg(part + (size_t)lane * 256 + blockIdx.x, __fadd_rn(s_acc[lane], s_acc[HKF_SUMS + lane]));
                    __syncwarp();
                    if (lane == 0) red_release_add(&st->gen, 1u);
                }
            }
            target += (unsigned)nb;
            ICPF_STAMP(1);
            // ---- every block waits for the level's blocks (bounded: a lost block must not hang the GPU)
            if (threadIdx.x == 0) {
                const long long t0 = clock64();
                while (ld_relaxed_u32(&st->gen) < target)
                    if (clock64() - t0 > 4000000000ll) { s_flag[4] = 1; break; }
            }
            __syncthreads();
            if (s_flag[4]) { if (blockIdx.x == 0 && threadIdx.x == 0) st->track_status = HKF_TRACK_PEER_TIMEOUT; return; }
            ICPF_STAMP(2);
            {
                float pv[2][8];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int k = warp + ICP_WARPS * q;
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const int e = lane + 32 * m;
                        pv[q][m] = (k < HKF_SUMS && e < nb) ? __ldcg(part + (size_t)k * 256 + e) : 0.f;
                    }
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int k = warp + ICP_WARPS * q;
                    const float c = tile_tree(pv[q]);
                    if (k < HKF_SUMS && lane == 0) s_sums[k] = c;
                }
            }
            __syncthreads();
            if (P2P) {
                // this rank's sums go out once (block 0); every block reads the mailbox of its own rank
                const unsigned seq = F.seq0 + (unsigned)itg;
                const int par = (int)(seq & 1u);
                if (warp == 0) {
                    if (blockIdx.x == 0 && lane < HKF_SUMS) {
                        const unsigned long long word = ((unsigned long long)seq << 32) | __float_as_uint(s_sums[lane]);
                        for (int r = 0; r < F.peers.nranks; ++r)
                            *(volatile unsigned long long*)&F.peers.mail[r]->sums[par][F.peers.rank][lane] = word;
                    }
                    const ShardMailbox* mine_box = F.peers.mail[F.peers.rank];
                    bool okp = true;
                    float a = 0.f;
                    if (lane < HKF_SUMS) {
                        const long long tw = clock64();
                        for (int r = 0; r < F.peers.nranks && okp; ++r) {
                            unsigned long long word;
                            while ((unsigned)((word = *(volatile const unsigned long long*)&mine_box->sums[par][r][lane]) >> 32) != seq)
                                if (clock64() - tw > 4000000000ll) { okp = false; break; }
                            const float w = __uint_as_float((unsigned)word);
                            a = r == 0 ? w : __fadd_rn(a, w);
                        }
                    }
                    okp = __all_sync(0xffffffffu, okp);
                    if (lane < HKF_SUMS) s_sums[lane] = a;
                    if (lane == 0 && !okp) s_flag[4] = 1;
                }
                __syncthreads();
                if (s_flag[4]) { if (blockIdx.x == 0 && threadIdx.x == 0) st->track_status = HKF_TRACK_PEER_TIMEOUT; return; }
            }
            ICPF_STAMP(3);
            if (blockIdx.x == 0 && threadIdx.x >= 64 && threadIdx.x < 64 + HKF_SUMS) st->sums[threadIdx.x - 64] = s_sums[threadIdx.x - 64];
#ifdef HKF_ICP_STAMPS
            icp_block_solve(s_sums, cur, F.angle_shake, F.dist_shake, s_c2, s_flag, warp, lane, blockIdx.x == 0 ? st->stamps[itg & 31] : nullptr);
#else
            icp_block_solve(s_sums, cur, F.angle_shake, F.dist_shake, s_c2, s_flag, warp, lane);
#endif
            __syncthreads();
            ICPF_STAMP(4);
            const int status = s_flag[2] ? HKF_TRACK_SINGULAR : s_flag[3] != HKF_SOLVE_OK ? HKF_TRACK_SHAKE : HKF_TRACK_OK;
            if (status != HKF_TRACK_OK) {                              // every block takes the same decision
                if (blockIdx.x == 0 && threadIdx.x == 0) st->track_status = status;
                return;
            }
            if (threadIdx.x < 16) s_m[threadIdx.x] = s_c2[threadIdx.x];
            if (last && blockIdx.x == 0) {
                if (threadIdx.x < 16) { st->cur[threadIdx.x] = s_c2[threadIdx.x]; st->pose[threadIdx.x] = s_c2[threadIdx.x]; }
                if (threadIdx.x == 32) {
                    float c2[16], inv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) c2[i] = s_c2[i];
                    hkf_mat44_inverse(c2, inv);
#pragma unroll
                    for (int i = 0; i < 16; ++i) st->pose_inv[i] = inv[i];
                }
            }
            __syncthreads();                                           // s_m is read by the next iteration's pixels
            ICPF_STAMP(5);
        }
    }
}
#endif

#ifndef HKF_EMULATE_DEVICE      /* host launch code below (tests/emu stops here) */
extern "C" int hkf_icp_track(hkf_ctx* c, const HkfFrameParams* fp)
{
    if (!c || !fp) return HKF_ERR_ARG;
    int rc = hkf_shard_flush(c);             // sharded mode: the model maps of the last raycast, composited now
    if (rc != HKF_OK) return rc;
    // cudaDownSampleModelVertices/Normals (the new-frame pyramids were built by hkf_preprocess)
    if (c->levels >= 2) {
        rc = hkf_pyramids_fused(c, c->model_v, c->model_n, nullptr, true);      // also starts the frame's tracking state
        if (rc != HKF_OK) return rc;
    } else {
        hkf_launch_pdl(k_track_begin, dim3(1), dim3(32), 0, c->stream, c->st);
        HKF_LAUNCH_CHECK(c);
    }
    const int* iters = icp_schedule(c->levels);
    // hkf_shard_set_icp_mode(ctx, 1) / HKF_SHARD_ICP=replicated: every rank evaluates the whole image (no exchange at all; the poses are bit-identical
    // because inputs and arithmetic are).  Default in sharded mode: banded rows + all-reduce of the 27 sums.
    const bool sharded = c->nranks > 1 && !c->icp_replicated;
    if (!sharded || c->p2p) {
        // the whole schedule in one launch of co-resident blocks (k_icp_frame).  The launch is cooperative where the
        // driver accepts that together with programmatic stream serialisation: co-residency is then guaranteed rather
        // than implied by 256 blocks x 512 threads fitting the 148 SMs twice over.
        IcpFrameArgs F{};
        F.levels = c->levels;
        F.dist2 = icp_sq_threshold(fp->icp_dist_thres); F.sin2 = icp_sq_threshold(fp->icp_norm_sin_thres);
        F.angle_shake = fp->icp_angle_shake; F.dist_shake = fp->icp_dist_shake;
        F.partials = c->partials;
        int total = 0;
        for (int l = 0; l < c->levels; ++l) {
            HkfCameraParams cam; hkf_level_camera(&c->ip.depth_camera, l, &cam);
            IcpLevel& L = F.lv[l];
            L.new_v = c->new_v[l]; L.new_n = c->new_n[l]; L.model_v = c->model_v[l]; L.model_n = c->model_n[l];
            L.cols = (int)c->cols[l]; L.rows = (int)c->rows[l];
            L.gx = (int)div_up(c->cols[l], 16); L.ntiles = L.gx * (int)div_up(c->rows[l], 16);
            L.y0 = 0; L.y1 = L.rows;
            if (sharded) {      // this rank's band of pixel rows
                L.y0 = (int)((long long)L.rows * c->rank / c->nranks);
                L.y1 = (int)((long long)L.rows * (c->rank + 1) / c->nranks);
            }
            L.cx = cam.cx; L.cy = cam.cy; L.fx = cam.fx; L.fy = cam.fy;
            L.iters = iters[l];
            L.gx_magic = ((unsigned long long)(L.ntiles + 512) * (unsigned)L.gx < (1ull << 24)) ? (unsigned)(((1u << 24) + L.gx - 1) / L.gx) : 0u;
            const int dim = L.cols > L.rows ? L.cols : L.rows;
            L.lim = dim <= 1024 ? 1024.f : dim <= 2048 ? 2048.f : 0.f;       // larger images: always the exact sequence
            L.guard = L.lim * (1.0f / 1048576.0f);
            total += iters[l];
        }
        if (sharded) { F.peers = c->peers; F.seq0 = c->shard_seq + 1; c->shard_seq += (unsigned)total; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(icp_grid(F.lv[0].ntiles)); cfg.blockDim = dim3(ICP_THREADS); cfg.stream = c->stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.attrs = attr;
        for (;;) {
            cfg.numAttrs = c->icp_coop ? 2 : 1;
            const cudaError_t e = sharded ? cudaLaunchKernelEx(&cfg, k_icp_frame<true>, F, c->st) : cudaLaunchKernelEx(&cfg, k_icp_frame<false>, F, c->st);
            if (e != cudaSuccess && c->icp_coop) { cudaGetLastError(); c->icp_coop = false; continue; }   // not accepted together: plain launch
            break;
        }
        HKF_LAUNCH_CHECK(c);
        return HKF_OK;
    }
    // sharded without peer memory: per-iteration launches with an NCCL allreduce in between
    for (int l = c->levels - 1; l >= 0; --l) {
        HkfCameraParams cam; hkf_level_camera(&c->ip.depth_camera, l, &cam);
        IcpArgs A; fill_icp_args(c, l, &cam, fp->icp_dist_thres, fp->icp_norm_sin_thres, &A);
        A.angle_shake = fp->icp_angle_shake; A.dist_shake = fp->icp_dist_shake;
        A.y0 = (int)((long long)A.rows * c->rank / c->nranks);
        A.y1 = (int)((long long)A.rows * (c->rank + 1) / c->nranks);
        for (int it = 0; it < iters[l]; ++it) {
            A.is_last = (l == 0 && it == iters[l] - 1) ? 1 : 0;
            k_icp<true, false, false><<<icp_grid(A.ntiles), ICP_THREADS, 0, c->stream>>>(A, c->st);
            HKF_LAUNCH_CHECK(c);
            rc = hkf_shard_allreduce_sums(c);
            if (rc != HKF_OK) return rc;
            k_icp_solve<<<1, 64, 0, c->stream>>>(c->st, A.angle_shake, A.dist_shake, A.is_last);
            HKF_LAUNCH_CHECK(c);
        }
    }
    return HKF_OK;
}
#endif // HKF_EMULATE_DEVICE (host launch code)
