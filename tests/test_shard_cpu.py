"""N > 1 host-side logic of the z-slab sharded mode on CPU: slab / halo / band arithmetic, the compositing rule, and
two-process gloo runs (world_size 2) of (a) the collective form of nearest-hit compositing, (b) the banded ICP sums
+ allreduce + identical solve on both ranks, (c) slab-wise integration by the oracle with no volume exchange."""
import os
import socket

import numpy as np
import pytest

import _oracle as O
from conftest import small_config, bits_equal
from hybkinectfu_b200 import shard, synth


def test_slabs_partition_the_volume_with_halos():
    for R, G in ((1024, 8), (512, 4), (128, 3), (64, 2)):
        halo = shard.halo_planes(0.032, 2.0, R)
        assert halo == int(np.ceil(0.032 / (2.0 / R))) + 3
        rs = shard.slab_ranges(R, G, halo)
        assert rs[0][0] == 0 and rs[-1][1] == R
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))              # owned planes tile [0, R)
        assert all(s0 == max(z0 - halo, 0) and s1 == min(z1 + halo, R) for z0, z1, s0, s1 in rs)
    assert shard.halo_planes(0.032, 2.0, 1024) == 20
    # work-balanced slabs: still a partition, thinner where the frustum is wide
    cfg = small_config(128)
    for G in (2, 4, 8):
        rs = shard.balanced_slab_ranges(cfg, G, 7)
        assert rs[0][0] == 0 and rs[-1][1] == 128 and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all(z1 - z0 >= 8 for z0, z1, _, _ in rs)
        w = shard.plane_work(cfg)
        stored = np.array([w[s0:s1].sum() for _, _, s0, s1 in rs])        # what a rank really processes: owned + halo
        assert stored.max() <= 1.25 * stored.mean()
    bands = [shard.icp_band(120, r, 8) for r in range(8)]
    assert bands[0][0] == 0 and bands[-1][1] == 120 and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))


def test_composite_rule_nearest_crossing_wins_even_if_it_failed():
    H, W = 4, 5
    t = np.full((3, H, W), np.inf, np.float32)
    ok = np.zeros((3, H, W), np.uint64)
    keys = np.full((3, H, W), shard.NO_HIT)
    v = np.zeros((3, H, W, 4), np.float32)
    n = np.zeros((3, H, W, 4), np.float32)
    # pixel (0,0): rank 1 hits at 0.5 (ok), rank 2 at 0.9 (ok) -> rank 1
    # pixel (0,1): rank 0 crossing at 0.4 FAILED refinement, rank 2 at 0.8 ok -> empty pixel
    # pixel (0,2): nobody
    for r, (y, x, tt, good) in ((1, (0, 0, 0.5, 1)), (2, (0, 0, 0.9, 1)), (0, (0, 1, 0.4, 0)), (2, (0, 1, 0.8, 1))):
        keys[r, y, x] = shard.pack_key(np.float32(tt), good, r)
        if good:
            v[r, y, x] = (r + 1, tt, 0, 1); n[r, y, x] = (0, 0, -1, 0)
    best, cv, cn = shard.composite(keys, v, n)
    assert tuple(cv[0, 0]) == (2, 0.5, 0, 1) and tuple(cn[0, 0]) == (0, 0, -1, 0)
    assert not cv[0, 1].any() and (int(best[0, 1]) & 0xFF) == 0 and ((int(best[0, 1]) >> 8) & 1) == 0
    assert best[0, 2] == shard.NO_HIT and not cv[0, 2].any()
    # key order == t order for positive floats
    ts = np.array([0.3, 0.30000004, 1.0, 3.9], np.float32)
    ks = shard.pack_key(ts, np.ones(4), 7)
    assert (np.diff(ks.astype(np.float64)) > 0).all()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = small_config(64)
        cam = synth.Camera().scaled(192, 144)
        poses = synth.trajectory(200)[::4][:3].astype(np.float32)
        frames = [synth.render_depth_mm(p.astype(np.float64), cam) for p in poses]
        out = {}
        # (0) the unique-id plumbing: rank 0's 128 bytes reach everyone
        uid = shard.broadcast_unique_id(lambda: bytes(range(128)))
        out["uid_ok"] = uid == bytes(range(128))
        # (c) slab-wise integration, no exchange: every rank integrates its stored planes of the same frames
        halo = shard.halo_planes(cfg.ray_increment, cfg.volume_size, cfg.resolution)
        z0, z1, s0, s1 = shard.slab_ranges(cfg.resolution, world, halo)[rank]
        vol = np.zeros((64, 64, 64, 2), np.float32)
        for i in range(2):
            tr = O.trunc(O.decode_mm(frames[i]), cfg.trunc_min, cfg.trunc_max)
            O.integrate(vol, cfg, tr, poses[i], s0, s1)
        out["slab"] = (s0, s1, vol[s0:s1].copy())
        # (b) banded ICP sums: pixels outside the rank's band contribute nothing; allreduce; same solve everywhere
        full = np.zeros((64, 64, 64, 2), np.float32)
        for i in range(2):
            O.integrate(full, cfg, O.trunc(O.decode_mm(frames[i]), cfg.trunc_min, cfg.trunc_max), poses[i])
        mv, mn = O.raycast(full, cfg, poses[1])
        tr = O.trunc(O.decode_mm(frames[2]), cfg.trunc_min, cfg.trunc_max)
        nv = O.vertices(O.bilateral(tr, cfg.sigma_pixel, cfg.sigma_depth), O.cam_of(cfg))
        nn = O.normals(nv)
        y0, y1 = shard.icp_band(cfg.rows, rank, world)
        nn_band = np.zeros_like(nn); nn_band[y0:y1] = nn[y0:y1]            # a zero normal = "no correspondence"
        inv = O.mat44_inverse(poses[1])
        part = O.icp_sums(nv, nn_band, mv, mn, O.cam_of(cfg), poses[1], inv, cfg.icp_dist_thres, cfg.icp_norm_sin_thres)
        import torch
        tsum = torch.from_numpy(part.copy())
        dist.all_reduce(tsum)
        out["sums"] = tsum.numpy()
        out["sums_full"] = O.icp_sums(nv, nn, mv, mn, O.cam_of(cfg), poses[1], inv, cfg.icp_dist_thres, cfg.icp_norm_sin_thres)
        ok, x = O.solve_icp(out["sums"])
        out["x"] = x
        # (a) compositing with collectives == the single-process rule
        rng = np.random.default_rng(5)
        t_all = rng.uniform(0.3, 3.0, (world, 6, 7)).astype(np.float32)
        hit = rng.random((world, 6, 7)) < 0.6
        okk = rng.random((world, 6, 7)) < 0.8
        keys = np.where(hit, np.stack([shard.pack_key(t_all[r], okk[r], r) for r in range(world)]), shard.NO_HIT)
        v = rng.normal(size=(world, 6, 7, 4)).astype(np.float32) * (hit & okk)[..., None]
        n = rng.normal(size=(world, 6, 7, 4)).astype(np.float32) * (hit & okk)[..., None]
        best, cv, cn = shard.composite_distributed(keys[rank], v[rank], n[rank])
        rb, rv, rn = shard.composite(keys, v, n)
        out["composite_ok"] = bool((best == rb).all() and np.array_equal(cv, rv) and np.array_equal(cn, rn))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_host_path():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in range(2):
        assert res[r]["uid_ok"] and res[r]["composite_ok"]
    # allreduced sums identical on both ranks, equal to the one-rank sums up to float summation order
    assert np.array_equal(res[0]["sums"], res[1]["sums"]) and np.array_equal(res[0]["x"], res[1]["x"])
    full = res[0]["sums_full"]
    assert np.abs(res[0]["sums"] - full).max() <= 2e-5 * np.abs(full).max()
    assert np.abs(full).max() > 0
    # slab-wise integration reproduces the whole-volume integration plane for plane, overlaps included
    cfg = small_config(64)
    cam = synth.Camera().scaled(192, 144)
    poses = synth.trajectory(200)[::4][:3].astype(np.float32)
    whole = np.zeros((64, 64, 64, 2), np.float32)
    for i in range(2):
        O.integrate(whole, cfg, O.trunc(O.decode_mm(synth.render_depth_mm(poses[i].astype(np.float64), cam)), cfg.trunc_min, cfg.trunc_max), poses[i])
    for r in range(2):
        s0, s1, planes = res[r]["slab"]
        assert bits_equal(planes, whole[s0:s1])


def test_plan_slabs_chooses_the_halo_layout(monkeypatch):
    """shard.plan_slabs: with peer memory the halo is READ in the neighbours' volumes (hkf_shard_peer_halo) - the cut
    tiles [0, R), balances the owned planes only and keeps every slab at least `halo` planes thick; without peer memory,
    when switched off, or when the volume is too small for that, every rank stores its halo"""
    from conftest import small_config
    cfg = small_config(256)
    halo = shard.halo_planes(cfg.ray_increment, cfg.volume_size, cfg.resolution)
    monkeypatch.delenv("HKF_SHARD_PEER_HALO", raising=False)
    for n in (2, 4, 8):
        rs, peer = shard.plan_slabs(cfg, n, halo, True)
        assert peer
        assert rs[0][0] == 0 and rs[-1][1] == 256 and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all(z1 - z0 >= halo and (s0, s1) == (z0, z1) for z0, z1, s0, s1 in rs)
        rs2, peer2 = shard.plan_slabs(cfg, n, halo, False)                 # NCCL fallback: no peer memory to read
        assert not peer2 and all(s0 == max(z0 - halo, 0) and s1 == min(z1 + halo, 256) for z0, z1, s0, s1 in rs2)
    monkeypatch.setenv("HKF_SHARD_PEER_HALO", "0")
    assert not shard.plan_slabs(cfg, 4, halo, True)[1]
    monkeypatch.delenv("HKF_SHARD_PEER_HALO")
    tiny = small_config(32)
    h = shard.halo_planes(tiny.ray_increment, tiny.volume_size, 32)
    assert not shard.plan_slabs(tiny, 8, h, True)[1]                       # 4 planes per rank < halo: stored halos
    assert shard.plan_slabs(cfg, 1, halo, True) == ([(0, 256, 0, 256)], False)
    # every rank computes the same plan from a measured cost profile
    cost = np.linspace(1.0, 3.0, 256)
    a = shard.plan_slabs(cfg, 4, halo, True, plane_cost=cost)
    assert a == shard.plan_slabs(cfg, 4, halo, True, plane_cost=cost.copy()) and a[1]
    owned = [cost[z0:z1].sum() for z0, z1, _, _ in a[0]]
    assert max(owned) <= 1.15 * (sum(owned) / 4)
