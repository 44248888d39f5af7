#!/usr/bin/env python
"""Run a dozen fused frames of the bench workload (for ncu: select kernels with -k / -s / -c)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import bench_cfg, make_sequence  # noqa: E402
from hybkinectfu_b200 import capi, synth  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
mode = sys.argv[3] if len(sys.argv) > 3 else "icp"
cfg = bench_cfg(R)
frames, gt = make_sequence(n)
ctx = capi.Context(cfg)
ctx.set_pose(synth.pose0().astype(np.float32))
for i in range(n):
    ctx.preprocess(frames[i])
    if i:
        ctx.icp_track() if mode == "icp" else ctx.sdf_track()
    ctx.integrate_tracked()
    ctx.raycast_tracked()
pose, st = ctx.frame_result()
print("status", st, "launches", ctx.launches())
ctx.close()
