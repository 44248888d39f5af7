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
__device__ __forceinline__ void icp_pixel_row(const IcpArgs& A, const float* __restrict__ cur, const float* __restrict__ li,
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
    // ---- the solve (hkf_icp_step, operation for operation) on warp 0 while warp 1 evaluates the determinant gate: Cholesky
    // on lane 0, the three sin/cos pairs on lanes 0-2, the 3x3 products on lanes 0-8, the shake gate on lane 0
    // next to the 4x4 product T * cur on lanes 16-31
    __shared__ float s_R[3][9], s_Rxy[9], s_Rf[9], s_c2[16];
    if (warp == 0) {
        // The Cholesky solve stays on ONE lane: its divisions and square roots form a dependent chain that the lane's own
        // instruction-level parallelism already overlaps column by column.  Spreading the rows over lanes (shuffles for the
        // pivot row, bit-identical results) was measured SLOWER on B200: 3.0 instead of 2.5 us per iteration
        // (profiles/r2_icp_stamps_*.txt) - the shuffle latency lands on the critical path.
        __shared__ float s_x[6];
        if (lane == 0) {
            float Am[36], bv[6], xs[6];
            hkf_unpack27(s_sums, Am, bv);
            hkf_llt_solve6(Am, bv, xs);
#pragma unroll
            for (int i = 0; i < 6; ++i) s_x[i] = xs[i];
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
        if (lane == 0) {
            // the shake gate of hkf_vector6_to_transform (a NaN passes, as in the reference)
            const float angle = hkf_rotation_angle(s_Rf);
            const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x[3], x[3]), __fmul_rn(x[4], x[4])), __fmul_rn(x[5], x[5])));
            s_flag[3] = (angle > A.angle_shake || d > A.dist_shake) ? HKF_SOLVE_SHAKE : HKF_SOLVE_OK;
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
    } else if (threadIdx.x == 32) {
        s_flag[2] = hkf_icp_singular(s_sums) ? 1 : 0;
    }
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


#ifndef HKF_EMULATE_DEVICE      /* host launch code below (tests/emu stops here) */
extern "C" int hkf_icp_track(hkf_ctx* c, const HkfFrameParams* fp)
{
    if (!c || !fp) return HKF_ERR_ARG;
    // cudaDownSampleModelVertices/Normals (the new-frame pyramids were built by hkf_preprocess)
    int rc;
    if (c->levels >= 2) {
        rc = hkf_pyramids_fused(c, c->model_v, c->model_n, nullptr, true);      // also starts the frame's tracking state
        if (rc != HKF_OK) return rc;
    } else {
        hkf_launch_pdl(k_track_begin, dim3(1), dim3(32), 0, c->stream, c->st);
        HKF_LAUNCH_CHECK(c);
    }
    const int* iters = icp_schedule(c->levels);
    // HKF_SHARD_ICP=replicated: every rank evaluates the whole image (no exchange at all; the poses are bit-identical
    // because inputs and arithmetic are).  Default in sharded mode: banded rows + all-reduce of the 27 sums.
    static const bool replicated = getenv("HKF_SHARD_ICP") && !strcmp(getenv("HKF_SHARD_ICP"), "replicated");
    const bool sharded = c->nranks > 1 && !replicated;
    if (!sharded || c->p2p) {
        // One launch per iteration, chained with programmatic dependent launch: the next iteration's blocks become
        // resident and issue their pose-independent loads while this iteration's last block is still reducing and
        // solving.  (A single persistent cooperative kernel with a ticket/generation grid barrier was measured slower:
        // 17.4 vs 13.7 us per iteration.)
        for (int l = c->levels - 1; l >= 0; --l) {
            HkfCameraParams cam; hkf_level_camera(&c->ip.depth_camera, l, &cam);
            IcpArgs A; fill_icp_args(c, l, &cam, fp->icp_dist_thres, fp->icp_norm_sin_thres, &A);
            A.angle_shake = fp->icp_angle_shake; A.dist_shake = fp->icp_dist_shake;
            if (sharded) {      // this rank's band of pixel rows
                A.y0 = (int)((long long)A.rows * c->rank / c->nranks);
                A.y1 = (int)((long long)A.rows * (c->rank + 1) / c->nranks);
                A.peers = c->peers;
            }
            for (int it = 0; it < iters[l]; ++it) {
                A.is_last = (l == 0 && it == iters[l] - 1) ? 1 : 0;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(icp_grid(A.ntiles)); cfg.blockDim = dim3(ICP_THREADS); cfg.stream = c->stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                if (sharded) { A.seq = ++c->shard_seq; cudaLaunchKernelEx(&cfg, k_icp<true, true, true>, A, c->st); }
                else cudaLaunchKernelEx(&cfg, k_icp<true, true, false>, A, c->st);
                HKF_LAUNCH_CHECK(c);
            }
        }
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
