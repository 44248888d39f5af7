#!/usr/bin/env python
"""Stage times of ONE z-slab context (no peers) of a 1024^3 volume cut 8 ways: integrate and slab raycast alone."""
import dataclasses
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import bench_cfg, make_sequence  # noqa: E402
from hybkinectfu_b200 import capi, shard, synth  # noqa: E402

R = 1024
cfg0 = bench_cfg(R)
halo = shard.halo_planes(cfg0.ray_increment, cfg0.volume_size, R)
ranges = shard.balanced_slab_ranges(cfg0, 8, halo)
if len(sys.argv) > 1:                    # owned plane counts of a measured layout, e.g. 372,101,86,79,73,67,45,201
    owned = [int(v) for v in sys.argv[1].split(",")]
    cuts = np.concatenate([[0], np.cumsum(owned)])
    assert cuts[-1] == R
    ranges = [(int(cuts[i]), int(cuts[i + 1]), max(int(cuts[i]) - halo, 0), min(int(cuts[i + 1]) + halo, R)) for i in range(len(owned))]
frames, gt = make_sequence(8)
for r in range(len(ranges)):
    z0, z1, s0, s1 = ranges[r]
    cfg = dataclasses.replace(cfg0, slab_z0=z0, slab_z1=z1, slab_halo=halo, max_triangles=1)
    ctx = capi.Context(cfg)
    stream = torch.cuda.ExternalStream(ctx.lib.hkf_stream(ctx._h))
    for i in range(6):
        ctx.preprocess(frames[i]); ctx.integrate(gt[i].astype(np.float32))
    ctx.sync()

    def timed(fn, reps=20):
        fn(); ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream); ctx.sync()
        return e0.elapsed_time(e1) * 1e3 / reps
    p = gt[5].astype(np.float32)
    print(f"rank {r}: owned [{z0},{z1}) stored {s1 - s0} planes: integrate {timed(lambda: ctx.integrate(p)):7.1f} us   slab raycast {timed(lambda: ctx.raycast(p)):7.1f} us")
    ctx.close()
