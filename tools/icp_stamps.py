#!/usr/bin/env python
"""%globaltimer stamps of the 19 tracking launches of one frame (profiling build: `make -C hybkinectfu_b200 stamps`,
run with HKF_LIB=hybkinectfu_b200/libhkf_b200_stamps.so).  Per iteration, for the LAST block of the launch:
  resident->loaded  new-frame pixel loads issued before the grid-dependency wait
  loaded->waited    waiting for the previous launch to complete (griddepcontrol.wait)
  waited->ticket    pose read, correspondence, tile sums, partial store, fence + ticket
  ticket->sums      second stage over the partials
  sums->solved      6x6 solve / gates on warps 0 and 1
  solved->end       block barrier + commit
  end->next waited  from this launch's last instruction to the next launch's blocks passing their wait"""
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
    # launch j of the context sits in slot j % 32: 9 tracked frames + 3 repetitions = 12 x 19 launches so far
    total = 19 * 12
    rows = [st[(total - 19 + k) % 32].astype(np.int64) for k in range(19)]
    names = ["res->loaded", "loaded->waited", "waited->ticket", "ticket->sums", "sums->solved", "solved->bar", "bar->end", "end->next waited", "total"]
    print("iter level " + " ".join(f"{n:>16s}" for n in names))
    lv = [2] * 4 + [1] * 5 + [0] * 10          # coarse to fine: 4, 5 and 10 iterations (CameraPoseFinderICP.cpp:16-30)
    tot = np.zeros(len(names))
    for k, r in enumerate(rows):
        d = [r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[7] - r[6],
             (rows[k + 1][2] - r[7]) if k + 1 < 19 else 0, (rows[k + 1][2] - r[2]) if k + 1 < 19 else 0]
        tot += np.array(d)
        print(f"{k:4d} {lv[k]:5d} " + " ".join(f"{v:16d}" for v in d))
    print("sum (ns)   " + " ".join(f"{v:16.0f}" for v in tot))
    ctx.close()


if __name__ == "__main__":
    main()
