#!/usr/bin/env python
"""per-frame wall times of the pipelined host-frame loop (development aid)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import bench_cfg, make_sequence
from hybkinectfu_b200 import capi, synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 120
pipe = (sys.argv[2] if len(sys.argv) > 2 else "1") == "1"
frames, gt = make_sequence(K + 1)
host = torch.from_numpy(frames.view(np.int16)).pin_memory()
fb = frames[0].nbytes
ctx = capi.Context(bench_cfg(512))
ctx.set_pose(synth.pose0().astype(np.float32))
ts = []
for k in range(K):
    t0 = time.perf_counter()
    if pipe:
        if k == 0:
            ctx._ck(ctx.lib.hkf_preprocess(ctx._h, host.data_ptr(), 0, ctx.fp), "pre")
        else:
            ctx.swap_frame()
        ctx.preprocess_next(host.data_ptr() + (k + 1) * fb, False)
    else:
        ctx._ck(ctx.lib.hkf_preprocess(ctx._h, host.data_ptr() + k * fb, 0, ctx.fp), "pre")
    t1 = time.perf_counter()
    if k:
        ctx.icp_track()
    ctx.integrate_tracked(); ctx.raycast_tracked()
    t2 = time.perf_counter()
    p, st = ctx.frame_result()
    t3 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1, t3 - t2, st))
a = np.array([t[:3] for t in ts]) * 1e6
for k in list(range(0, K, max(K // 12, 1))):
    print(k, a[k].round(0), ts[k][3])
print("mean us per frame", a.sum(1)[5:].mean())
