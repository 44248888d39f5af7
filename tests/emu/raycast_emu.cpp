// raycast_emu.cpp — the ray-march kernel of hybkinectfu_b200/csrc/raycast.cu compiled for the host and run thread by
// thread (tests/test_raycast_emulated.py compares it with the CPU oracle).  Test infrastructure, not a CPU fallback: the
// product library never links this.
#include "emu_shim.h"
#include "../../hybkinectfu_b200/csrc/tsdf_sample.cuh"
#include "emu_indices.h"
#include "../../hybkinectfu_b200/csrc/raycast.cu"

extern "C" int emu_raycast(const float* vol, int R, float S, unsigned cols, unsigned rows, float cx, float cy, float fx, float fy,
                           const float* pose16, float inc, float near_plane, float far_plane,
                           int fast, float* out_v4, float* out_n4)
{
    RaycastArgs A;
    memset(&A, 0, sizeof(A));
    A.V.vol = reinterpret_cast<const float2*>(vol);
    A.V.R = R; A.V.S = S; A.V.Rf = (float)R; A.V.cell = S / (float)R; A.V.z_lo = 0; A.V.z_hi = R; A.V.kz_lo = 0; A.V.kz_hi = R;
    A.cols = (int)cols; A.rows = (int)rows; A.cx = cx; A.cy = cy; A.fx = fx; A.fy = fy;
    A.inc = inc; A.near_plane = near_plane; A.far_plane = far_plane;
    memcpy(A.pose, pose16, 64);
    A.own_z0 = 0; A.own_z1 = R; A.rank = 0; A.nranks = 1; A.epoch = 0;
    float4* ov = reinterpret_cast<float4*>(out_v4);
    float4* on = reinterpret_cast<float4*>(out_n4);
    const unsigned gx = (cols + 15) / 16, gy = (rows + 15) / 16;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) {
            emu_blockIdx = { bx, by, 0 }; emu_blockDim = { 256, 1, 1 }; emu_gridDim = { gx, gy, 1 };
            for (unsigned t = 0; t < 256; ++t) {
                emu_threadIdx = { t, 0, 0 };
                if (fast) k_raycast<false, false, true>(A, ov, on, nullptr, nullptr);
                else k_raycast<false, false, false>(A, ov, on, nullptr, nullptr);
            }
        }
    return 0;
}

// One z-slab of the volume (csrc/shard.cu): `vol` holds planes [z_lo, z_hi); with a peer-memory halo (hkf_shard_peer_halo)
// vlo / vhi point at the neighbours' planes [kz_lo, z_lo) and [z_hi, kz_hi) - on the host a "peer" is another array.
// Writes this slab's partial maps and its 64-bit (t | ok | rank) keys; no pushes (peers.nranks = 0).
extern "C" int emu_raycast_slab(const float* vol, const float* vlo, const float* vhi, int R, float S, int z_lo, int z_hi, int kz_lo, int kz_hi,
                                int own_z0, int own_z1, int rank, unsigned cols, unsigned rows, float cx, float cy, float fx, float fy,
                                const float* pose16, float inc, float near_plane, float far_plane,
                                int fast, float* out_v4, float* out_n4, unsigned long long* keys)
{
    RaycastArgs A;
    memset(&A, 0, sizeof(A));
    A.V.vol = reinterpret_cast<const float2*>(vol);
    A.V.vlo = reinterpret_cast<const float2*>(vlo); A.V.vhi = reinterpret_cast<const float2*>(vhi);
    A.V.R = R; A.V.S = S; A.V.Rf = (float)R; A.V.cell = S / (float)R;
    A.V.z_lo = z_lo; A.V.z_hi = z_hi; A.V.kz_lo = kz_lo; A.V.kz_hi = kz_hi;
    A.cols = (int)cols; A.rows = (int)rows; A.cx = cx; A.cy = cy; A.fx = fx; A.fy = fy;
    A.inc = inc; A.near_plane = near_plane; A.far_plane = far_plane;
    memcpy(A.pose, pose16, 64);
    A.own_z0 = own_z0; A.own_z1 = own_z1; A.rank = rank; A.nranks = 0; A.epoch = 0;
    float4* ov = reinterpret_cast<float4*>(out_v4);
    float4* on = reinterpret_cast<float4*>(out_n4);
    const unsigned gx = (cols + 15) / 16, gy = (rows + 15) / 16;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) {
            emu_blockIdx = { bx, by, 0 }; emu_blockDim = { 256, 1, 1 }; emu_gridDim = { gx, gy, 1 };
            for (unsigned t = 0; t < 256; ++t) {
                emu_threadIdx = { t, 0, 0 };
                if (fast) k_raycast<false, true, true>(A, ov, on, keys, nullptr);
                else k_raycast<false, true, false>(A, ov, on, keys, nullptr);
            }
        }
    return 0;
}
