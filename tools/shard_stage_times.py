#!/usr/bin/env python
"""Per-rank stage times of the z-slab sharded mode with the ranks held together by host barriers, so that no stage's time
contains waiting for another rank: integrate | slab raycast (with its pushes) | compositing.  Run under torchrun:

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/shard_stage_times.py

HKF_SHARD_PEER_HALO=0/1 selects the halo layout (stored / read in the neighbours' memory)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bench_cfg, make_sequence  # noqa: E402
from hybkinectfu_b200 import shard, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
os.environ["HKF_SHARD_MANUAL_COMPOSITE"] = "1"
cfg = bench_cfg(1024)
n = 40
frames, gt = make_sequence(n)
ctx = shard.sharded_context(cfg, rank, world, local)
ctx._ck(ctx.lib.hkf_shard_set_icp_mode(ctx._h, 1), "icp mode")
st = torch.cuda.ExternalStream(ctx.lib.hkf_stream(ctx._h), device=local)
ctx.set_pose(synth.pose0().astype(np.float32))
acc = np.zeros(4)
cnt = 0
for i in range(n):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    ctx.preprocess(frames[i])
    ev[0].record(st)
    if i:
        ctx.icp_track()
    ev[1].record(st)
    ctx.sync(); dist.barrier()
    ev[2].record(st); ctx.integrate_tracked(); ev[3].record(st)
    ctx.sync(); dist.barrier()
    ev[4].record(st); ctx.raycast_tracked(); ev[5].record(st)
    ctx.sync(); dist.barrier()
    ev[6].record(st); ctx._ck(ctx.lib.hkf_shard_composite_raycast(ctx._h), "composite"); ev[7].record(st)
    ctx.sync(); dist.barrier()
    if i >= 8:
        acc += [ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]), ev[4].elapsed_time(ev[5]), ev[6].elapsed_time(ev[7])]
        cnt += 1
p, status = ctx.frame_result()
acc *= 1e3 / cnt
t = torch.tensor([ctx.cfg.slab_z0, ctx.cfg.slab_z1, *acc], dtype=torch.float64, device=f"cuda:{local}")
out = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(out, t)
if rank == 0:
    print(f"N={world} halo {'peer memory' if ctx.peer_halo else 'stored'}  (us per frame, ranks held together by barriers; track status {status})")
    for r, x in enumerate(out):
        x = x.cpu().numpy()
        print(f"  rank {r}: owned [{int(x[0])},{int(x[1])})  icp {x[2]:7.1f}  integrate {x[3]:7.1f}  slab raycast+push {x[4]:7.1f}  composite {x[5]:7.1f}")
ctx.close()
dist.barrier()
dist.destroy_process_group()
