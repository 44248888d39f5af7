#!/usr/bin/env python
"""Per-stage CUDA-event timings of the fused path on the bench workload (development aid).
Each stage is repeated back to back on the state reached after a few tracked frames."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import bench_cfg, make_sequence  # noqa: E402
from hybkinectfu_b200 import capi, synth  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = 30
    cfg = bench_cfg(R)
    frames, gt = make_sequence(12)
    ctx = capi.Context(cfg)
    stream = torch.cuda.ExternalStream(ctx.lib.hkf_stream(ctx._h))
    dev = torch.from_numpy(frames.view(np.int16)).cuda()
    fb = frames[0].nbytes
    ctx.set_pose(synth.pose0().astype(np.float32))
    for i in range(10):
        ctx.preprocess_device(dev.data_ptr() + i * fb)
        if i:
            ctx.icp_track()
        ctx.integrate_tracked(); ctx.raycast_tracked()
    pose, st = ctx.frame_result()
    assert st == 0

    def timed(name, fn, per=1):
        fn(); ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        ctx.sync()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f"{name:34s} {us:9.1f} us" + (f"   ({us/per:7.1f} us each of {per})" if per > 1 else ""))
        return us

    t = {}
    t["pre"] = timed("preprocess (fused, 2 launches)", lambda: ctx.preprocess_device(dev.data_ptr() + 10 * fb))

    def track():
        ctx.set_pose  # noqa
        ctx.icp_track()
    # keep the start pose fixed: icp_track commits the pose, so restore it each time (host call, async memcpy)
    def track_fixed():
        ctx.lib.hkf_set_pose(ctx._h, capi.Mat44.from_np(pose))
        ctx.icp_track()
    t["icp"] = timed("icp_track (pyr + begin + 19 iters)", lambda: ctx.icp_track(), 21)
    t["int"] = timed("integrate_tracked", lambda: ctx.integrate_tracked())
    t["ray"] = timed("raycast_tracked", lambda: ctx.raycast_tracked())
    ctx.lib.hkf_set_pose(ctx._h, capi.Mat44.from_np(pose))
    for l in range(3):
        timed(f"icp sums only level {l}", lambda l=l: ctx.lib.hkf_cal_point_to_plane_err_solver_params(
            ctx._h, l, capi.Mat44.from_np(pose), capi.Mat44.from_np(np.linalg.inv(pose)), cfg.level_camera(l), 0.1, 0.2))
    timed("sdf sums only", lambda: ctx.lib.hkf_cal_sdf_solver_params(ctx._h, cfg.camera(), capi.Mat44.from_np(pose)))
    nupd = ctx.count_updates(pose)
    print(f"N_upd {nupd} ({100.0*nupd/R**3:.1f}%)  integrate algorithmic GB/s {(16*nupd+4*640*480)/t['int']/1e3:.0f}")
    print(f"sum of stages {t['pre']+t['icp']+t['int']+t['ray']:.1f} us -> {1e6/(t['pre']+t['icp']+t['int']+t['ray']):.0f} fps")
    ctx.close()


if __name__ == "__main__":
    main()
