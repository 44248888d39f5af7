// icp_emu.cpp — host emulation of the ICP kernel of hybkinectfu_b200/csrc/icp.cu: k_icp in its wrapper-parity form (the
// 27 sums of one iteration against the oracle) and in its fused form (sums + the last block's 6x6 solve + gates + pose update, iteration after iteration, against the loop `oracle sums -> host solve`).
// Test infrastructure; the product library never links this.
#include "emu_shim.h"
#include "emu_block.h"
#include "../../hybkinectfu_b200/csrc/hkf_internal.cuh"
#include "emu_indices.h"
#include "../../hybkinectfu_b200/csrc/icp.cu"

static void fill(IcpArgs* A, const float* nv, const float* nn, const float* mv, const float* mn, int cols, int rows,
                 float cx, float cy, float fx, float fy, float dist_thres, float sin_thres, float angle_shake, float dist_shake, float* partials)
{
    memset(A, 0, sizeof(*A));
    A->new_v = reinterpret_cast<const float4*>(nv); A->new_n = reinterpret_cast<const float4*>(nn);
    A->model_v = reinterpret_cast<const float4*>(mv); A->model_n = reinterpret_cast<const float4*>(mn);
    A->cols = cols; A->rows = rows;
    A->gx = (cols + 15) / 16; A->ntiles = A->gx * ((rows + 15) / 16);
    A->y0 = 0; A->y1 = rows;
    A->cx = cx; A->cy = cy; A->fx = fx; A->fy = fy;
    A->dist_thres = dist_thres; A->sin_thres = sin_thres; A->angle_shake = angle_shake; A->dist_shake = dist_shake;
    A->partials = partials;
}

// one iteration of the default kernel with the transforms passed in (the wrapper-parity form): the 27 sums
extern "C" void emu_icp_sums(const float* nv, const float* nn, const float* mv, const float* mn, int cols, int rows,
                             float cx, float cy, float fx, float fy, const float* cur16, const float* last_inv16,
                             float dist_thres, float sin_thres, float* out27)
{
    std::vector<float> partials(HKF_SUMS * 256, 0.f);
    IcpArgs A;
    fill(&A, nv, nn, mv, mn, cols, rows, cx, cy, fx, fy, dist_thres, sin_thres, 0.f, 0.f, partials.data());
    memcpy(A.cur, cur16, 64); memcpy(A.last_inv, last_inv16, 64);
    HkfDeviceState st;
    memset(&st, 0, sizeof(st));
    const unsigned grid = (unsigned)(A.ntiles < 256 ? A.ntiles : 256);
    emu_launch(dim3(grid), dim3(ICP_THREADS), [&]() { k_icp<false, false, false>(A, &st); });
    memcpy(out27, st.sums, HKF_SUMS * sizeof(float));
}

// `iters` launches of the fused kernel (device-resident working transform, last-block solve) on the state `st`
extern "C" void emu_icp_fused(const float* nv, const float* nn, const float* mv, const float* mn, int cols, int rows,
                              float cx, float cy, float fx, float fy, float dist_thres, float sin_thres, float angle_shake,
                              float dist_shake, int iters, void* st)
{
    std::vector<float> partials(HKF_SUMS * 256, 0.f);
    IcpArgs A;
    fill(&A, nv, nn, mv, mn, cols, rows, cx, cy, fx, fy, dist_thres, sin_thres, angle_shake, dist_shake, partials.data());
    const unsigned grid = (unsigned)(A.ntiles < 256 ? A.ntiles : 256);
    for (int it = 0; it < iters; ++it)
        emu_launch(dim3(grid), dim3(ICP_THREADS), [&]() { k_icp<true, true, false>(A, static_cast<HkfDeviceState*>(st)); });
}

extern "C" unsigned emu_state_bytes() { return (unsigned)sizeof(HkfDeviceState); }
extern "C" void emu_state_init(void* st, const float* cur16, const float* last_inv16)
{
    HkfDeviceState* s = static_cast<HkfDeviceState*>(st);
    memset(s, 0, sizeof(*s));
    memcpy(s->cur, cur16, 64); memcpy(s->last_inv, last_inv16, 64);
    s->track_status = HKF_TRACK_OK;
}
extern "C" void emu_state_read(const void* st, float* cur16, float* sums27, int* status)
{
    const HkfDeviceState* s = static_cast<const HkfDeviceState*>(st);
    memcpy(cur16, s->cur, 64); memcpy(sums27, s->sums, HKF_SUMS * sizeof(float)); *status = s->track_status;
}

extern "C" int emu_can_spawn(int n) { return emu_probe_threads(n) ? 1 : 0; }
