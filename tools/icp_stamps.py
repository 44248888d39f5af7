#!/usr/bin/env python
"""%globaltimer stamps of the 19 tracking iterations of one frame inside k_icp_frame (profiling build:
`make -C hybkinectfu_b200 stamps`, run with HKF_LIB=hybkinectfu_b200/libhkf_b200_stamps.so).  Per iteration, block 0:
  tiles+arrive    correspondence, tile sums, partial store, fence + arrival
  wait for all    until every block of the level has arrived
  fold partials   second stage over the [27][<=256] partials
  solve           6x6 solve / gates on warps 0 and 1
  commit          new working transform into shared memory"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import bench_cfg, make_sequence  # noqa: E402
from hybkinectfu_b200 import capi, synth  # noqa: E402


def main():
    cfg = bench_cfg(512)
    frames, gt = make_sequence(12)
    ctx = capi.Context(cfg)
    ctx.set_pose(synth.pose0().astype(np.float32))
    for i in range(10):
        ctx.preprocess(frames[i])
        if i:
            ctx.icp_track()
        ctx.integrate_tracked(); ctx.raycast_tracked()
    ctx.sync()
    ctx.preprocess(frames[10])
    for rep in range(3):
        ctx.icp_track(); ctx.sync()
    st = np.zeros((32, 8), np.uint64)
    got = ctx.lib.hkf_download(ctx._h, 100, 0, st.ctypes.data, st.nbytes)
    assert got == st.nbytes, "not a stamps build (HKF_LIB=hybkinectfu_b200/libhkf_b200_stamps.so)"
    # k_icp_frame: block 0 stamps iteration k of the schedule into slot k
    rows = [st[k].astype(np.int64) for k in range(19)]
    names = ["tiles+arrive", "wait for all", "fold partials", "solve", "(cholesky", "rotation", "gate+T*cur)", "commit", "total"]
    print("iter level " + " ".join(f"{n:>14s}" for n in names))
    lv = [2] * 4 + [1] * 5 + [0] * 10          # coarse to fine: 4, 5 and 10 iterations (CameraPoseFinderICP.cpp:16-30)
    tot = np.zeros(len(names))
    for k, r in enumerate(rows):
        d = [r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[6] - r[3], r[7] - r[6], r[4] - r[7], r[5] - r[4], r[5] - r[0]]
        tot += np.array(d)
        print(f"{k:4d} {lv[k]:5d} " + " ".join(f"{v:14d}" for v in d))
    print("sum (ns)   " + " ".join(f"{v:14.0f}" for v in tot))
    print(f"frame: first iteration start -> last iteration end {rows[18][5] - rows[0][0]} ns")
    ctx.close()


if __name__ == "__main__":
    main()
