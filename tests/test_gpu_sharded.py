"""z-slab sharded mode on real GPUs (needs >= 2 visible devices: run with `gpurun --gpus 2`): one process per GPU,
NCCL communicator owned by libhkf_b200.so.

 * poses given: every rank integrates only its stored planes; the composited raycast (nearest-hit over ranks) must
   equal the single-GPU raycast of the whole volume BIT FOR BIT, and each slab must equal the same planes of the
   single-GPU volume;
 * ICP tracking: poses bitwise identical on all ranks, within the north_star tolerance of the single-GPU run (the
   27 sums are added in a different order), volume planes equal wherever the pose bits are equal.
"""
import os
import socket

import numpy as np
import pytest

from conftest import bits_equal, canon_tris, small_config

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _frames(n=5):
    from hybkinectfu_b200 import synth
    cam = synth.Camera().scaled(192, 144)
    poses = synth.trajectory(200)[::4][:n].astype(np.float32)
    return np.stack([synth.render_depth_mm(p.astype(np.float64), cam) for p in poses]), poses


def _worker(rank, world, port, q, res, peer_halo):
    os.environ["HKF_SHARD_PEER_HALO"] = "1" if peer_halo else "0"
    import torch
    import torch.distributed as dist
    from hybkinectfu_b200 import capi, shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        cfg = small_config(res)
        frames, gt = _frames()
        ctx = shard.sharded_context(cfg, rank, world, rank)
        out = {"slab": (ctx.cfg.slab_z0, ctx.cfg.slab_z1, ctx.cfg.slab_halo), "peer_halo": ctx.peer_halo}
        # ---- poses given
        ctx.set_pose(gt[0])
        for i, mm in enumerate(frames):
            ctx.preprocess(mm)
            ctx.set_pose(gt[i])
            ctx.integrate_tracked()
            ctx.raycast_tracked()
        p, st = ctx.frame_result()
        out["given_v"] = ctx.download(capi.BUF_MODEL_VERTICES)
        out["given_n"] = ctx.download(capi.BUF_MODEL_NORMALS)
        out["given_vol"] = ctx.download(capi.BUF_VOLUME)
        dist.barrier()                    # marching cubes reads neighbour planes (peer halo): everybody has finished
        n = ctx.marching_cubes()
        out["given_tris"] = ctx.download(capi.BUF_TRIANGLES)[:n].copy()
        dist.barrier()
        # ---- the same, pipelined: no host synchronisation between the calls of a frame or between frames (the deferred
        # compositing and, with a peer-memory halo, the ordering of a rank's integration after the neighbours' marches
        # rest on the device-side handshakes alone)
        import ctypes as C
        ctx.reset_volume()
        ctx.set_pose(gt[0])
        dev = torch.from_numpy(frames.view(np.int16)).cuda()
        for rep in range(3):
            for i in range(len(frames)):
                ctx._ck(ctx.lib.hkf_preprocess(ctx._h, dev.data_ptr() + i * frames[0].nbytes, 1, C.byref(ctx.fp)), "hkf_preprocess")
                ctx.set_pose(gt[i])
                ctx.integrate_tracked()
                ctx.raycast_tracked()
        ctx.frame_result()
        out["async_v"] = ctx.download(capi.BUF_MODEL_VERTICES)
        out["async_vol"] = ctx.download(capi.BUF_VOLUME)
        dist.barrier()
        # ---- ICP tracking
        ctx.reset_volume()
        ctx.set_pose(synth.pose0().astype(np.float32))
        poses, sts = [], []
        for i, mm in enumerate(frames):
            ctx.preprocess(mm)
            if i:
                ctx.icp_track()
            ctx.integrate_tracked()
            ctx.raycast_tracked()
            p, st = ctx.frame_result()
            poses.append(p); sts.append(st)
        out["icp_poses"] = np.stack(poses); out["icp_sts"] = np.array(sts)
        out["icp_v"] = ctx.download(capi.BUF_MODEL_VERTICES)
        ctx.close()
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("peer_halo", [1, 0])
@pytest.mark.parametrize("world,res", [(2, 128), (max(2, min(_ngpu(), 8)), 256)])
def test_sharded_equals_single_gpu(world, res, peer_halo):
    """peer_halo=1: the halo planes are read in the neighbours' memory (hkf_shard_peer_halo), every rank stores exactly
    the planes it owns; peer_halo=0: every rank stores and integrates its halo itself"""
    import torch.multiprocessing as mp
    from hybkinectfu_b200 import capi, shard, synth
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q, res, peer_halo)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # single-GPU reference of the same code
    cfg = small_config(res)
    frames, gt = _frames()
    c = capi.Context(cfg)
    c.set_pose(gt[0])
    for i, mm in enumerate(frames):
        c.preprocess(mm); c.set_pose(gt[i]); c.integrate_tracked(); c.raycast_tracked()
    c.frame_result()
    v1, n1, vol1 = c.download(capi.BUF_MODEL_VERTICES), c.download(capi.BUF_MODEL_NORMALS), c.download(capi.BUF_VOLUME)
    ntri = c.marching_cubes()
    tris1 = c.download(capi.BUF_TRIANGLES)[:ntri].copy()
    assert (n1[..., :3] != 0).any(-1).sum() > 1000 and ntri > 1000
    halo = shard.halo_planes(cfg.ray_increment, cfg.volume_size, cfg.resolution)
    os.environ["HKF_SHARD_PEER_HALO"] = "1" if peer_halo else "0"
    ranges, peer = shard.plan_slabs(cfg, world, halo, True)
    assert peer == bool(peer_halo)
    assert np.array_equal(canon_tris(np.concatenate([got[r]["given_tris"] for r in range(world)])).view(np.uint32),
                          canon_tris(tris1).view(np.uint32)), "per-slab marching cubes do not add up to the whole mesh"
    for r in range(world):
        assert bits_equal(got[r]["given_v"], v1), f"composited vertices differ on rank {r}"
        assert bits_equal(got[r]["given_n"], n1), f"composited normals differ on rank {r}"
        z0, z1, s0, s1 = ranges[r]
        assert got[r]["peer_halo"] == peer and (s0, s1) == ((z0, z1) if peer else (max(z0 - halo, 0), min(z1 + halo, cfg.resolution)))
        assert got[r]["slab"] == (z0, z1, 0 if peer else halo)
        assert bits_equal(got[r]["given_vol"], vol1[s0:s1]), f"slab {r} differs from planes [{s0},{s1}) of the whole volume"
    # pipelined, three passes over the frames
    c.reset_volume(); c.set_pose(gt[0])
    for rep in range(3):
        for i, mm in enumerate(frames):
            c.preprocess(mm); c.set_pose(gt[i]); c.integrate_tracked(); c.raycast_tracked()
    c.frame_result()
    v3, vol3 = c.download(capi.BUF_MODEL_VERTICES), c.download(capi.BUF_VOLUME)
    for r in range(world):
        z0, z1, s0, s1 = ranges[r]
        assert bits_equal(got[r]["async_v"], v3), f"pipelined run: composited vertices differ on rank {r}"
        assert bits_equal(got[r]["async_vol"], vol3[s0:s1]), f"pipelined run: slab {r} differs"
    c.reset_volume(); c.set_pose(synth.pose0().astype(np.float32))
    poses = []
    for i, mm in enumerate(frames):
        c.preprocess(mm)
        if i:
            c.icp_track()
        c.integrate_tracked(); c.raycast_tracked()
        p, st = c.frame_result()
        assert st == 0
        poses.append(p)
    poses = np.stack(poses)
    c.close()
    for r in range(world):
        assert (got[r]["icp_sts"] == 0).all()
        assert bits_equal(got[r]["icp_poses"], got[0]["icp_poses"]), "ranks disagree on the pose"
        assert bits_equal(got[r]["icp_v"], got[0]["icp_v"])
    d = np.abs(got[0]["icp_poses"] - poses)
    assert d[:, :3, :3].max() <= 1e-5 and d[:, :3, 3].max() <= 1e-4, (d[:, :3, :3].max(), d[:, :3, 3].max())
