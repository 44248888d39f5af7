// shard_emu.cpp — the N > 1 raycast data path on the host: the SLAB instantiation of the ray-march kernel
// (hybkinectfu_b200/csrc/raycast.cu) PUSHING its hits into every rank's inbox, then k_composite_p2p
// (hybkinectfu_b200/csrc/shard.cu) compositing from a rank's own inboxes - the kernel sources, compiled for the host.
// "Ranks" are arrays of one process; the cross-rank handshake words are preset (every peer has announced), everything
// else - slot layout, double buffering by epoch parity, the reset of used slots, the winner's payload - is the code
// that runs on the GPUs.  Test infrastructure, not a CPU fallback: the product library never links this.
#include "emu_shim.h"
#include "emu_block.h"
#include "../../hybkinectfu_b200/csrc/tsdf_sample.cuh"
#include "emu_indices.h"
#include "../../hybkinectfu_b200/csrc/raycast.cu"
#include "../../hybkinectfu_b200/csrc/shard.cu"

extern "C" int emu_shard_max_ranks() { return HKF_MAX_RANKS; }
extern "C" int emu_shard_can_run() { return emu_probe_threads(256) ? 1 : 0; }

// whole: the whole volume [R][R][R][2]; rank r owns planes [cuts[r], cuts[r+1]).  peer = 0: a rank's slab holds its halo
// planes; peer = 1: it holds the owned planes and reads the halo in the neighbours' planes (here: the same array).
// keys_in / pay_in: the inboxes of all ranks, [nranks][2][HKF_MAX_RANKS][npix] u64 / [..][npix][8] floats, kept by the
// caller from one epoch to the next (keys initialised to ~0).  out_v / out_n: [nranks][npix][4], every rank's composited maps.
extern "C" int emu_sharded_raycast(const float* whole, int R, float S, int nranks, const int* cuts, int halo, int peer,
                                   unsigned cols, unsigned rows, float cx, float cy, float fx, float fy, const float* pose16,
                                   float inc, float near_plane, float far_plane, int fast, unsigned epoch,
                                   unsigned long long* keys_in, float* pay_in, float* out_v, float* out_n)
{
    if (nranks < 1 || nranks > HKF_MAX_RANKS) return 1;
    const size_t npix = (size_t)cols * rows, pl = (size_t)R * R;
    static ShardMailbox mail[HKF_MAX_RANKS];
    ShardPeers P{};
    P.nranks = nranks;
    for (int r = 0; r < nranks; ++r) {
        P.mail[r] = &mail[r];
        P.keys_in[r] = keys_in + (size_t)r * 2 * HKF_MAX_RANKS * npix;
        P.pay_in[r] = reinterpret_cast<float4*>(pay_in) + (size_t)r * 2 * HKF_MAX_RANKS * npix * 2;
    }
    const float2* vol = reinterpret_cast<const float2*>(whole);
    std::vector<unsigned long long> keys(npix);
    const unsigned gx = (cols + 15) / 16, gy = (rows + 15) / 16;
    for (int r = 0; r < nranks; ++r) {
        const int z0 = cuts[r], z1 = cuts[r + 1];
        const int k0 = z0 - halo > 0 ? z0 - halo : 0, k1 = z1 + halo < R ? z1 + halo : R;
        RaycastArgs A;
        memset(&A, 0, sizeof(A));
        A.V.R = R; A.V.S = S; A.V.Rf = (float)R; A.V.cell = S / (float)R;
        if (peer) {
            A.V.vol = vol + (size_t)z0 * pl; A.V.z_lo = z0; A.V.z_hi = z1;
            A.V.vlo = vol + (size_t)k0 * pl; A.V.vhi = vol + (size_t)z1 * pl; A.V.kz_lo = k0; A.V.kz_hi = k1;
        } else {
            A.V.vol = vol + (size_t)k0 * pl; A.V.z_lo = k0; A.V.z_hi = k1; A.V.kz_lo = k0; A.V.kz_hi = k1;
        }
        A.cols = (int)cols; A.rows = (int)rows; A.cx = cx; A.cy = cy; A.fx = fx; A.fy = fy;
        A.inc = inc; A.near_plane = near_plane; A.far_plane = far_plane;
        memcpy(A.pose, pose16, 64);
        A.own_z0 = z0; A.own_z1 = z1; A.rank = r; A.nranks = nranks; A.epoch = epoch;
        A.peers = P; A.peers.rank = r;
        float4* ov = reinterpret_cast<float4*>(out_v) + (size_t)r * npix;
        float4* on = reinterpret_cast<float4*>(out_n) + (size_t)r * npix;
        for (unsigned by = 0; by < gy; ++by)
            for (unsigned bx = 0; bx < gx; ++bx) {
                emu_blockIdx = { bx, by, 0 }; emu_blockDim = { 256, 1, 1 }; emu_gridDim = { gx, gy, 1 };
                for (unsigned t = 0; t < 256; ++t) {
                    emu_threadIdx = { t, 0, 0 };
                    if (fast) k_raycast<false, true, true>(A, ov, on, keys.data(), nullptr);
                    else k_raycast<false, true, false>(A, ov, on, keys.data(), nullptr);
                }
            }
    }
    // every rank has finished its raycast and announced it: the handshake of k_composite_p2p finds all words in place
    for (int r = 0; r < nranks; ++r)
        for (int q = 0; q < nranks; ++q) mail[r].ready[q] = epoch;
    for (int r = 0; r < nranks; ++r) {
        ShardPeers Pr = P;
        Pr.rank = r;
        HkfDeviceState st;
        memset(&st, 0, sizeof(st));
        float4* ov = reinterpret_cast<float4*>(out_v) + (size_t)r * npix;
        float4* on = reinterpret_cast<float4*>(out_n) + (size_t)r * npix;
        emu_launch(dim3((unsigned)((npix + 255) / 256)), dim3(256), [&] { k_composite_p2p(Pr, ov, on, (unsigned)npix, epoch, &st); });
        if (st.track_status != 0) return 2;
    }
    return 0;
}
