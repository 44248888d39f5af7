// raycast.cu — fixed-step ray marching of the TSDF into world-frame vertex / normal maps.
//
// Replaces raycastKernel / raySample / gradientForPoint / getMinTime / getMaxTime and
// cudaRaycastingVolume (src/cuda/raycastingVolume.cu:15-176).
//
// Parity is defined by the sample positions, not just the surface: t starts at t_min and is advanced by
// repeated float addition of the increment; each sample reads the NEAREST voxel's tsdf; the first
// (+ -> -) pair ends the ray whether or not the two trilinear refinements and the six gradient
// evaluations succeed (a failure leaves the pixel zero).  None of that is altered here.
//
// What changes: packed 8-byte voxels (a sample touches one 4-byte word instead of a 12-byte AoS record),
// warps mapped to 8x4 pixel patches so that the rays of a warp traverse neighbouring voxels and share
// sectors, all eight corner loads of a trilinear evaluation issued before any is tested, the pose taken
// from device memory (no host round trip), and a z-slab mode for sharded volumes in which each rank
// reports at most the crossings whose later sample lies in its OWNED planes, as a 64-bit
// (t | ok | rank) key for nearest-hit compositing.
//
// Arithmetic contract (PTX of the reference built for sm_100):
//   cam_dir = (vx, vy, 1) * rcp.rn(sqrt.rn(fma(vx,vx, vy*vy) + 1)),  vx = ((float)x - cx) / fx
//   dir     = fma(m_4, 0, fma(cd.z, m_3, fma(cd.x, m_1, cd.y*m_2)))   per pose row; 0 -> 1e-15
//   t_min   = max(max over axes of ((dir>0 ? 0 : S) - o)/dir, near/cd.z), t_max likewise with min
//   pos     = fma(dir, t, o);  t += inc
//   alpha   = t - (inc*ftdt)/(ftdt - ft);  vertex = fma(dir, alpha, o)
//   normal  = (fx1-fx2, fy1-fy2, fz1-fz2) * rcp.rn(len), len = sqrt.rn(dot3), fails if (double)len < 1e-8
#include "tsdf_sample.cuh"

#define RAY_NB 8          // independent samples of a ray in flight

struct RaycastArgs {
    VolView V;
    int cols, rows;
    float cx, cy, fx, fy;
    float inc, near_plane, far_plane;
    float pose[16];
    int own_z0, own_z1, rank;
    int nranks; unsigned epoch;        // peer-memory mode: epoch of this raycast (its parity selects the inbox)
    int halo_wait;                     // peer-memory halo in use: wait for the neighbours' integration (see the kernel)
    ShardPeers peers;                  // peer-memory mode: where hits are pushed (peers.nranks == 0: no push)
    // has_color (raycastingVolume.cu:101-106): the voxel colour plane and raycast_rgb; nullptr = depth only.  The colour
    // of a hit is interpolated at the refined vertex BEFORE the gradient is evaluated and is written whenever that
    // interpolation succeeds, even if the gradient then fails and the pixel's vertex / normal stay zero.
    const uchar4* color_vol; unsigned char* color_out;
};

// FAST: R is a power of two and the stored planes hold at most 2^32 voxels (the host checks both) - the cheaper
// sample-index arithmetic below applies; the results are the same bits either way.
//
// Empty-space skipping (brick bitmaps "may hold a negative voxel" written by the integration item pass: unflagged 16^3
// bricks not read, whole batches skipped over unflagged 64^3 bricks) was built, shown exact on the CPU and on the GPU,
// and measured SLOWER on B200 - 120 us against 78 us for this kernel (profiles/r2_experimental.log): the flag lookups sit
// in front of every sample load and lengthen the dependent chain of a march that is latency-bound, not traffic-bound.
// It has been removed.
// COLOR: has_color (whole volumes only); a separate instantiation so that the depth-only kernels keep their registers.
template <bool FROM_STATE, bool SLAB, bool FAST, bool COLOR = false>
__global__ void __launch_bounds__(256, SLAB ? 3 : 0)
k_raycast(RaycastArgs A, float4* __restrict__ out_v, float4* __restrict__ out_n,
          unsigned long long* __restrict__ keys, const HkfDeviceState* __restrict__ st)
{
    hkf_pdl_enter();
#ifndef HKF_EMULATE_DEVICE             // (the host emulation of this kernel runs one thread at a time, one rank)
    if (SLAB && A.halo_wait) {
        // peer-memory halo: the neighbours' planes are read in place, so their integration of this frame must be complete.
        // Block 0 announces this rank's (stream order: its integration launches have finished; system fence), every block
        // waits for the two neighbours' announcements of the same epoch.  (The reverse hazard - a neighbour integrating the
        // NEXT frame while this march still reads its planes - is excluded by the compositing handshake: its integration
        // follows its compositing, which waits for this rank's raycast.)
        __shared__ int s_go;
        if (threadIdx.x == 0) {
            const ShardPeers& P = A.peers;
            if (blockIdx.x == 0 && blockIdx.y == 0) {
                __threadfence_system();
                if (P.rank > 0) *(volatile unsigned*)&P.mail[P.rank - 1]->integ[P.rank] = A.epoch;
                if (P.rank + 1 < P.nranks) *(volatile unsigned*)&P.mail[P.rank + 1]->integ[P.rank] = A.epoch;
            }
            bool ok = true;
            const long long t0 = clock64();
            for (int d = -1; d <= 1 && ok; d += 2) {
                const int r = P.rank + d;
                if (r < 0 || r >= P.nranks) continue;
                while ((int)(*(volatile const unsigned*)&P.mail[P.rank]->integ[r] - A.epoch) < 0)
                    if (clock64() - t0 > 4000000000ll) { ok = false; break; }          // ~2 s: a peer died
            }
            __threadfence_system();
            s_go = ok;
        }
        __syncthreads();
        if (!s_go) return;                 // the compositing times out on the same peer and sets HKF_TRACK_PEER_TIMEOUT
    }
#endif
    // 16x16 pixel block, warps as 8x4 patches
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= A.cols || y >= A.rows) return;
    const size_t pix = (size_t)y * A.cols + x;
    float4 ov = make_float4(0.f, 0.f, 0.f, 0.f), on = ov;
    unsigned long long key = ~0ull;
    if (COLOR) { A.color_out[pix * 3 + 0] = 0; A.color_out[pix * 3 + 1] = 0; A.color_out[pix * 3 + 2] = 0; }

    const float* P = FROM_STATE ? st->pose : A.pose;
    const float m11 = P[0], m12 = P[1], m13 = P[2], ox = P[3];
    const float m21 = P[4], m22 = P[5], m23 = P[6], oy = P[7];
    const float m31 = P[8], m32 = P[9], m33 = P[10], oz = P[11];
    const VolView& V = A.V;
    const float S = V.S;

    const float vx = __fdiv_rn(__fsub_rn((float)(unsigned)x, A.cx), A.fx);
    const float vy = __fdiv_rn(__fsub_rn((float)(unsigned)y, A.cy), A.fy);
    const float len = __fsqrt_rn(__fadd_rn(__fmaf_rn(vx, vx, __fmul_rn(vy, vy)), 1.0f));
    float cdx = 0.f, cdy = 0.f, cdz = 0.f;
    if (!((double)len < 1e-8)) { const float r = __frcp_rn(len); cdx = __fmul_rn(vx, r); cdy = __fmul_rn(vy, r); cdz = r; }
    float dx = __fmaf_rn(ox, 0.f, __fmaf_rn(cdz, m13, __fmaf_rn(cdx, m11, __fmul_rn(cdy, m12))));
    float dy = __fmaf_rn(oy, 0.f, __fmaf_rn(cdz, m23, __fmaf_rn(cdx, m21, __fmul_rn(cdy, m22))));
    float dz = __fmaf_rn(oz, 0.f, __fmaf_rn(cdz, m33, __fmaf_rn(cdx, m31, __fmul_rn(cdy, m32))));
    if (dx == 0.f) dx = 1e-15f;
    if (dy == 0.f) dy = 1e-15f;
    if (dz == 0.f) dz = 1e-15f;
    const float txmin = __fdiv_rn(__fsub_rn(dx > 0.f ? 0.f : S, ox), dx);
    const float tymin = __fdiv_rn(__fsub_rn(dy > 0.f ? 0.f : S, oy), dy);
    const float tzmin = __fdiv_rn(__fsub_rn(dz > 0.f ? 0.f : S, oz), dz);
    const float txmax = __fdiv_rn(__fsub_rn(dx > 0.f ? S : 0.f, ox), dx);
    const float tymax = __fdiv_rn(__fsub_rn(dy > 0.f ? S : 0.f, oy), dy);
    const float tzmax = __fdiv_rn(__fsub_rn(dz > 0.f ? S : 0.f, oz), dz);
    float tmin = fmaxf(fmaxf(txmin, tymin), tzmin);
    float tmax = fminf(fminf(txmax, tymax), tzmax);
    tmin = fmaxf(tmin, __fdiv_rn(A.near_plane, cdz));
    tmax = fminf(tmax, __fdiv_rn(A.far_plane, cdz));

    if (tmin < tmax) {
        const int R = V.R;
        float t = tmin, last_sdf = 0.f, last_t = tmin;
        bool last_known = true;
        float t_leave = 3.0e38f;
        int nlo = 0, nhi = 0;              // SLAB: planes whose samples this ray needs
        if (SLAB) {
            // z is monotonic along a ray: the one sample before the owned planes lies below them when the ray climbs and
            // above them when it descends; on the far side nothing is needed (a crossing there is another rank's).  The
            // batches below read speculatively - eight samples span more planes than a thin slab has - and a sample
            // outside [nlo, nhi) is not read where it lies (with a peer-memory halo that would be a neighbour's memory,
            // over NVLink, for a value nobody uses) but replaced by "unknown".
            nlo = dz >= 0.f ? V.kz_lo : A.own_z0;
            nhi = dz >= 0.f ? A.own_z1 : V.kz_hi;
            // A rank reports the crossings whose LATER sample lies in a plane it owns (any earlier crossing of the ray
            // belongs to another rank, whose smaller key wins the compositing).  So only the samples inside the owned
            // planes (one cell of slack for the rounding of the voxel index) and the one sample before them matter:
            // t advances through everything earlier with the same repeated float addition (the sample sequence is part
            // of parity) without the voxel arithmetic, and the march stops where the ray leaves the owned planes.  The
            // halo planes are only touched by that one earlier sample and by the refinement of a crossing.
            const float zl = A.own_z0 == 0 ? -1.0e30f : ((float)A.own_z0 - 1.0f) * V.cell;
            const float zh = A.own_z1 == R ? 1.0e30f : ((float)A.own_z1 + 1.0f) * V.cell;
            const float ta = (zl - oz) / dz, tb = (zh - oz) / dz;
            const float t_enter = fminf(ta, tb) - A.inc;
            t_leave = fmaxf(ta, tb);
            if (t < t_enter) {
                while (t < t_enter && t < tmax) t = __fadd_rn(A.inc, t);
                last_known = false;
            }
        }
        // Nearest-voxel index of a sample, (int)((p*R)/S) clamped: the quotient is first taken with a multiplication by
        // 1/S (within 2 ulp of the IEEE quotient); only when that lands within 1e-3 of an integer does the reference's
        // div.rn decide.  The samples of a ray do not depend on each other, so they are taken RAY_NB at a time:
        // independent loads in flight, then the crossing test in order.
        const float rS = __frcp_rn(S);
        const unsigned Rm1 = (unsigned)(R - 1);
        auto voxel = [&](float p) -> int {
            const float q = __fmul_rn(p, V.Rf);
            const float qa = __fmul_rn(q, rS);
            int g = __float2int_rz(qa);
            if (fabsf(qa - rintf(qa)) < 1.0e-3f) g = __float2int_rz(__fdiv_rn(q, S));
            return (int)min((unsigned)max(g, 0), Rm1);
        };
        // FAST sample index.  With R = 2^k the product p*R is exact, so (p*R)*(1/S) == p*(R*(1/S)) bit for bit: one
        // multiplication gives the same approximate quotient qa as `voxel` (within 2 ulp of the IEEE quotient, both
        // below R).  All three axes share ONE test: the distance of qa to its nearest integer (1.5*2^23 rounding, FMA
        // pipe) must exceed 4 ulp(R) on every axis and the truncated indices must lie in [0, R) - otherwise (a few
        // percent of the warp-samples) `voxel` decides as before, clamping included.
        const float rRS = __fmul_rn(V.Rf, rS);
        const float guard = V.Rf * 2.384185791015625e-7f;       // R * 2^-22
        const float MAGIC = 12582912.0f;
        const bool mono = A.inc > 0.f;                          // t never decreases: ts[b] <= ts[NB-1]
        constexpr int NB = SLAB ? 4 : RAY_NB;                   // a slab ray has a handful of samples to take
        // nearest-voxel indices of the sample at parameter tt (the reference's integers, whichever path computes them)
        auto voxel3 = [&](float tt, int& gx, int& gy, int& gz) {
            const float px = __fmaf_rn(dx, tt, ox), py = __fmaf_rn(dy, tt, oy), pz = __fmaf_rn(dz, tt, oz);
            if (FAST) {
                const float qx = __fmul_rn(px, rRS), qy = __fmul_rn(py, rRS), qz = __fmul_rn(pz, rRS);
                const float ex = fabsf(__fsub_rn(qx, __fsub_rn(__fadd_rn(qx, MAGIC), MAGIC)));
                const float ey = fabsf(__fsub_rn(qy, __fsub_rn(__fadd_rn(qy, MAGIC), MAGIC)));
                const float ez = fabsf(__fsub_rn(qz, __fsub_rn(__fadd_rn(qz, MAGIC), MAGIC)));
                gx = __float2int_rz(qx); gy = __float2int_rz(qy); gz = __float2int_rz(qz);
                if (!(fminf(fminf(ex, ey), ez) >= guard) || (unsigned)(gx | gy | gz) > Rm1) { gx = voxel(px); gy = voxel(py); gz = voxel(pz); }
            } else {
                gx = voxel(px); gy = voxel(py); gz = voxel(pz);
            }
        };
        bool done = false;
        while (!done && t < tmax && t < t_leave) {
            float ts[NB], sd[NB];
            int gzs[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                ts[b] = t;
                int gx, gy, gz;
                voxel3(t, gx, gy, gz);
                gzs[b] = gz;
                if (SLAB) {
                    // keep the speculative read inside the readable planes; a plane of the halo comes from a neighbour's memory
                    const int gzc = min(max(gz, nlo), nhi - 1);
                    sd[b] = __ldg(&(vol_plane(V, gzc) + ((unsigned)gy * (unsigned)R + (unsigned)gx))->x);
                } else if (FAST) sd[b] = __ldg(&V.vol[((unsigned)gz * (unsigned)R + (unsigned)gy) * (unsigned)R + (unsigned)gx].x);
                else sd[b] = __ldg(&V.vol[vol_index(V, gx, gy, gz)].x);
                t = __fadd_rn(A.inc, t);
            }
            // Nothing can happen in this batch (the common case) unless a (+ -> -) pair shows up, the ray reaches t_max,
            // or a sample lies outside the stored planes: one predicate instead of the in-order scan below.
            {
                bool event = !(mono && ts[NB - 1] < tmax) || !last_known;
                float prev = last_sdf;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    event |= (prev > 0.0f) & (sd[b] < 0.0f);
                    if (SLAB) event |= (unsigned)(gzs[b] - nlo) >= (unsigned)(nhi - nlo);
                    prev = sd[b];
                }
                if (!event) {
                    last_sdf = sd[NB - 1]; last_t = ts[NB - 1];
                    continue;
                }
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (done) break;
                if (!(ts[b] < tmax)) { done = true; break; }
                const int gz = gzs[b];
                const bool known = !SLAB || (gz >= nlo && gz < nhi);
                const float sdf = known ? sd[b] : 0.f;
                if (known && last_known && last_sdf > 0.0f && sdf < 0.0f) {
                    // first zero crossing: the ray ends here whatever happens next
                    done = true;
                    const bool mine = !SLAB || (gz >= A.own_z0 && gz < A.own_z1);
                    if (!mine) break;          // another rank owns this crossing and reports it; nothing later can win
                    const float tc = ts[b];
                    const float px = __fmaf_rn(dx, tc, ox), py = __fmaf_rn(dy, tc, oy), pz = __fmaf_rn(dz, tc, oz);
                    const float lx = __fmaf_rn(dx, last_t, ox), ly = __fmaf_rn(dy, last_t, oy), lz = __fmaf_rn(dz, last_t, oz);
                    bool ok = false;
                    float ftdt, ft;
                    float wx = 0.f, wy = 0.f, wz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
                    do {
                        if (vol_interp_sdf<SLAB>(V, px, py, pz, ftdt) != 1) break;
                        if (vol_interp_sdf<SLAB>(V, lx, ly, lz, ft) != 1) break;
                        const float alpha = __fsub_rn(tc, __fdiv_rn(__fmul_rn(A.inc, ftdt), __fsub_rn(ftdt, ft)));
                        wx = __fmaf_rn(dx, alpha, ox); wy = __fmaf_rn(dy, alpha, oy); wz = __fmaf_rn(dz, alpha, oz);
                        if (COLOR) {
                            const unsigned c3 = vol_interp_color(V, A.color_vol, wx, wy, wz);
                            if (c3 != HKF_NO_COLOR) {
                                A.color_out[pix * 3 + 0] = (unsigned char)c3; A.color_out[pix * 3 + 1] = (unsigned char)(c3 >> 8); A.color_out[pix * 3 + 2] = (unsigned char)(c3 >> 16);
                            }
                        }
                        // gradientForPoint(volume, last_world_pos, vertex)
                        const int hx = vol_w2v(V, lx), hy = vol_w2v(V, ly), hz = vol_w2v(V, lz);
                        if (hx <= 1 || hx >= R - 2) break;
                        if (hy <= 1 || hy >= R - 2) break;
                        if (hz <= 1 || hz >= R - 2) break;
                        const float cell = V.cell;
                        float f1, f2;
                        if (vol_interp_sdf<SLAB>(V, __fadd_rn(wx, cell), wy, wz, f1) != 1) break;
                        if (vol_interp_sdf<SLAB>(V, __fsub_rn(wx, cell), wy, wz, f2) != 1) break;
                        nx = __fsub_rn(f1, f2);
                        if (vol_interp_sdf<SLAB>(V, wx, __fadd_rn(wy, cell), wz, f1) != 1) break;
                        if (vol_interp_sdf<SLAB>(V, wx, __fsub_rn(wy, cell), wz, f2) != 1) break;
                        ny = __fsub_rn(f1, f2);
                        if (vol_interp_sdf<SLAB>(V, wx, wy, __fadd_rn(wz, cell), f1) != 1) break;
                        if (vol_interp_sdf<SLAB>(V, wx, wy, __fsub_rn(wz, cell), f2) != 1) break;
                        nz = __fsub_rn(f1, f2);
                        const float l = __fsqrt_rn(hkf_dot3(nx, ny, nz, nx, ny, nz));
                        if ((double)l < 1e-8) break;
                        const float r = __frcp_rn(l);
                        nx = __fmul_rn(nx, r); ny = __fmul_rn(ny, r); nz = __fmul_rn(nz, r);
                        ok = true;
                    } while (0);
                    if (ok) { ov = make_float4(wx, wy, wz, 1.0f); on = make_float4(nx, ny, nz, 0.f); }
                    key = ((unsigned long long)__float_as_uint(tc) << 32) | ((unsigned long long)(ok ? 1u : 0u) << 8) | (unsigned)A.rank;
                    break;
                }
                last_sdf = sdf; last_known = known; last_t = ts[b];
            }
        }
    }
    out_v[pix] = ov;
    out_n[pix] = on;
    if (SLAB) {
        keys[pix] = key;
        if (A.peers.nranks > 0 && key != ~0ull) {
            // a hit: its key (and, if the refinement succeeded, its vertex / normal) goes into the inbox of every rank,
            // this one included.  Posted stores: nothing waits for them here.
            const size_t npix = (size_t)A.cols * A.rows;
            const size_t slot = ((size_t)(A.epoch & 1u) * HKF_MAX_RANKS + (unsigned)A.rank) * npix + pix;
            for (int r = 0; r < A.peers.nranks; ++r) {
                A.peers.keys_in[r][slot] = key;
                if ((key >> 8) & 1ull) { A.peers.pay_in[r][2 * slot] = ov; A.peers.pay_in[r][2 * slot + 1] = on; }
            }
        }
    }
}

#ifndef HKF_EMULATE_DEVICE      /* host launch code below (tests/emu stops here) */
static void fill_ray_args(hkf_ctx* c, const float* pose, const HkfRayCasterParams* rp, const HkfCameraParams* cam,
                          float near_plane, float far_plane, RaycastArgs* A)
{
    A->V = make_volview(c);
    A->cols = (int)cam->cols; A->rows = (int)cam->rows;
    A->cx = cam->cx; A->cy = cam->cy; A->fx = cam->fx; A->fy = cam->fy;
    A->inc = rp->fRayIncrement; A->near_plane = near_plane; A->far_plane = far_plane;
    if (pose) memcpy(A->pose, pose, 64); else memset(A->pose, 0, 64);
    A->own_z0 = (int)c->own_z0; A->own_z1 = (int)c->own_z1; A->rank = c->rank;
    A->nranks = c->nranks; A->epoch = 0;
    A->peers = ShardPeers{};
    A->color_vol = nullptr; A->color_out = nullptr;
    A->halo_wait = 0;
    if (c->p2p) { A->peers = c->peers; A->epoch = ++c->shard_epoch; A->halo_wait = (c->halo_lo || c->halo_hi) ? 1 : 0; }
}

static inline bool is_slab(const hkf_ctx* c) { return !(c->z_lo == 0 && c->z_hi == c->ip.volume_resolution); }
static inline bool is_fast(const hkf_ctx* c)
{
    const unsigned long long R = c->ip.volume_resolution;
    return (R & (R - 1)) == 0 && (unsigned long long)(c->z_hi - c->z_lo) * R * R <= 0x100000000ull;
}

extern "C" int hkf_raycasting_volume(hkf_ctx* c, int has_color, const HkfMat44* transform, const HkfRayCasterParams* rp,
                                     const HkfCameraParams* cam, float near_plane, float far_plane)
{
    if (!c || !transform || !rp || !cam) return HKF_ERR_ARG;
    if (has_color && !c->color) { c->err = "context was created without use_color"; return HKF_ERR_STATE; }
    if (cam->cols != c->cols[0] || cam->rows != c->rows[0]) { c->err = "camera size mismatch"; return HKF_ERR_ARG; }
    RaycastArgs A; fill_ray_args(c, transform->m, rp, cam, near_plane, far_plane, &A);
    if (has_color) { A.color_vol = c->color; A.color_out = c->raycast_rgb; }
    dim3 b(256), g(div_up(A.cols, 16), div_up(A.rows, 16));
    const bool fast = is_fast(c);
    if (is_slab(c)) {
        if (fast) k_raycast<false, true, true><<<g, b, 0, c->stream>>>(A, c->model_v[0], c->model_n[0], c->ray_keys, c->st);
        else k_raycast<false, true, false><<<g, b, 0, c->stream>>>(A, c->model_v[0], c->model_n[0], c->ray_keys, c->st);
    } else if (has_color) {
        if (fast) k_raycast<false, false, true, true><<<g, b, 0, c->stream>>>(A, c->model_v[0], c->model_n[0], c->ray_keys, c->st);
        else k_raycast<false, false, false, true><<<g, b, 0, c->stream>>>(A, c->model_v[0], c->model_n[0], c->ray_keys, c->st);
    } else {
        if (fast) k_raycast<false, false, true><<<g, b, 0, c->stream>>>(A, c->model_v[0], c->model_n[0], c->ray_keys, c->st);
        else k_raycast<false, false, false><<<g, b, 0, c->stream>>>(A, c->model_v[0], c->model_n[0], c->ray_keys, c->st);
    }
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

extern "C" int hkf_raycast_tracked(hkf_ctx* c, const HkfFrameParams* fp)
{
    if (!c || !fp) return HKF_ERR_ARG;
    { const int rc = hkf_shard_flush(c); if (rc != HKF_OK) return rc; }      // one raycast at a time in the inboxes
    RaycastArgs A; fill_ray_args(c, nullptr, &fp->raycast, &c->ip.depth_camera, fp->trunc_min, fp->trunc_max, &A);
    if (fp->use_color) {
        if (!c->color) { c->err = "context was created without use_color"; return HKF_ERR_STATE; }
        A.color_vol = c->color; A.color_out = c->raycast_rgb;
    }
    dim3 b(256), g(div_up(A.cols, 16), div_up(A.rows, 16));
    const HkfDeviceState* st = c->st;
    const bool fast = is_fast(c);
    if (is_slab(c)) {
        if (fast) hkf_launch_pdl(k_raycast<true, true, true>, g, b, 0, c->stream, A, c->model_v[0], c->model_n[0], c->ray_keys, st);
        else hkf_launch_pdl(k_raycast<true, true, false>, g, b, 0, c->stream, A, c->model_v[0], c->model_n[0], c->ray_keys, st);
    } else if (fp->use_color) {
        if (fast) hkf_launch_pdl(k_raycast<true, false, true, true>, g, b, 0, c->stream, A, c->model_v[0], c->model_n[0], c->ray_keys, st);
        else hkf_launch_pdl(k_raycast<true, false, false, true>, g, b, 0, c->stream, A, c->model_v[0], c->model_n[0], c->ray_keys, st);
    } else {
        if (fast) hkf_launch_pdl(k_raycast<true, false, true>, g, b, 0, c->stream, A, c->model_v[0], c->model_n[0], c->ray_keys, st);
        else hkf_launch_pdl(k_raycast<true, false, false>, g, b, 0, c->stream, A, c->model_v[0], c->model_n[0], c->ray_keys, st);
    }
    HKF_LAUNCH_CHECK(c);
    // sharded volume: nearest-hit compositing of the per-slab partial raycasts -> full model maps on every rank
    // (HKF_SHARD_MANUAL_COMPOSITE=1: the caller invokes hkf_shard_composite_raycast itself, e.g. to time it separately)
    // The compositing is deferred to the next call that needs the model maps (hkf_shard_flush): the next frame's preprocess
    // does not, and runs while the other ranks' hits are still arriving.
    static const bool manual = getenv("HKF_SHARD_MANUAL_COMPOSITE") && getenv("HKF_SHARD_MANUAL_COMPOSITE")[0] == '1';
    if (c->nranks > 1 && !manual) c->composite_pending = true;
    return HKF_OK;
}
#endif // HKF_EMULATE_DEVICE
