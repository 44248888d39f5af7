"""z-slab sharding of one TSDF volume across the GPUs of a box (BASELINE.json configs[3]).

Host-side logic only: how the volume is cut into slabs with halos, which pixel rows of an ICP level a rank
evaluates, the nearest-hit compositing rule, and the torch.distributed plumbing that distributes the NCCL unique
id of the C library's own communicator.  The device work is csrc/shard.cu, icp.cu and raycast.cu.

The functions here are backend-agnostic (they are exercised with gloo on CPU in tests/test_shard_cpu.py); the
only NCCL calls the data path makes are the ones inside libhkf_b200.so.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import os

import numpy as np

NO_HIT = np.uint64(0xFFFFFFFFFFFFFFFF)
# ICP in sharded mode (env HKF_SHARD_ICP overrides): "banded" = every rank evaluates a band of pixel rows and the 27 sums
# are exchanged; "replicated" = every rank evaluates the whole image (no exchange; ICP is latency-bound, not work-bound)
DEFAULT_ICP_MODE = "banded"


def halo_planes(ray_increment: float, volume_size: float, resolution: int) -> int:
    """Planes a slab must replicate per interior side: a crossing is owned by the slab of its LATER sample, its
    earlier sample lies up to one ray increment before it, the two trilinear refinements reach one more voxel and
    the central-difference gradient one cell plus one voxel beyond that (src/cuda/raycastingVolume.cu:15-43,65-119)."""
    cell = volume_size / resolution
    return int(math.ceil(ray_increment / cell)) + 3


def slab_ranges(resolution: int, nranks: int, halo: int):
    """[(own_z0, own_z1, stored_z0, stored_z1)] per rank: owned planes partition [0, R), stored = owned + halo"""
    out = []
    for r in range(nranks):
        z0 = resolution * r // nranks
        z1 = resolution * (r + 1) // nranks
        out.append((z0, z1, max(z0 - halo, 0), min(z1 + halo, resolution)))
    return out


def plane_work(cfg, pose=None, samples: int = 48) -> np.ndarray:
    """Relative integration + raycast work of each z plane for a camera at `pose` (default: HybKinectfu's initial
    pose): the fraction of the plane inside the view frustum and the integration depth range, plus a constant."""
    from . import synth
    pose = synth.pose0(cfg.volume_size, cfg.trunc_min) if pose is None else np.asarray(pose, np.float64)
    inv = np.linalg.inv(pose)
    R, S = cfg.resolution, cfg.volume_size
    g = (np.arange(samples) + 0.5) * (S / samples)
    xx, yy = np.meshgrid(g, g)
    w = np.zeros(R)
    for z in range(R):
        wz = (z + 0.5) * S / R
        p = np.stack([xx, yy, np.full_like(xx, wz), np.ones_like(xx)], -1) @ inv.T
        Z = p[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = p[..., 0] * cfg.fx / Z + cfg.cx
            v = p[..., 1] * cfg.fy / Z + cfg.cy
        inside = (Z > 0) & (Z < cfg.max_integrate_dist + cfg.sdf_trunc) & (u >= 0) & (u < cfg.cols) & (v >= 0) & (v < cfg.rows)
        w[z] = inside.mean()
    # measured on B200 (1024^3, integrate + raycast per plane): cost grows like the square root of the covered
    # fraction (free-space bricks are cheap, the per-row and near-surface work is not) plus a small constant
    return np.sqrt(w) + 0.05


def plane_cost_from_profile(profile: np.ndarray) -> np.ndarray:
    """relative cost of every plane from a measured work profile [R][general chunks, free bricks] (Context.profile_read,
    summed over a representative run): a general chunk (4 x 32 voxels, every voxel projected) costs about four times a
    free brick (4 x 128 voxels streamed) on B200 (profiles/r2_integrate_e_rmq.txt), plus the per-plane share of the
    classification pass"""
    c = profile[:, 0].astype(np.float64) + 0.23 * profile[:, 1].astype(np.float64)
    return c + 0.02 * max(c.mean(), 1e-9)


def balanced_slab_ranges(cfg, nranks: int, halo: int, pose=None, min_planes: int = 8, plane_cost=None):
    """slab_ranges with boundaries chosen so that every rank gets the same share of plane_work over the planes it
    STORES (owned + halo: the halo is integrated and marched redundantly, so thin slabs cost more than they own).
    The view frustum widens with depth: equal-thickness slabs are not equal work.
    halo = 0 with min_planes >= the raycast halo: the peer-memory-halo cut (nothing stored or integrated twice)."""
    R = cfg.resolution
    if nranks == 1:
        return [(0, R, 0, R)]
    if R < nranks:
        raise ValueError(f"{nranks} slabs do not fit into {R} planes")
    min_planes = max(1, min(min_planes, R // nranks))      # a small volume over many ranks: thinner slabs, never an empty one
    w = plane_work(cfg, pose) if plane_cost is None else np.asarray(plane_cost, np.float64)
    assert len(w) == R
    cum = np.concatenate([[0.0], np.cumsum(w)])

    def stored_cost(z0, z1):
        return cum[min(z1 + halo, R)] - cum[max(z0 - halo, 0)]

    def greedy(limit):
        cuts, z0 = [0], 0
        for r in range(nranks - 1):
            z1 = z0 + min_planes
            while z1 < R - min_planes * (nranks - 1 - r) and stored_cost(z0, z1 + 1) <= limit:
                z1 += 1
            cuts.append(z1)
            z0 = z1
        cuts.append(R)
        return cuts

    lo, hi = 0.0, float(cum[-1])
    for _ in range(40):                      # smallest per-rank budget for which the last slab fits too
        mid = 0.5 * (lo + hi)
        c = greedy(mid)
        if stored_cost(c[-2], R) <= mid:
            hi = mid
        else:
            lo = mid
    cuts = greedy(hi)
    return [(cuts[r], cuts[r + 1], max(cuts[r] - halo, 0), min(cuts[r + 1] + halo, R)) for r in range(nranks)]


def plan_slabs(cfg, nranks: int, halo: int, p2p: bool, balanced: bool = True, plane_cost=None):
    """(ranges, peer_halo): the slab cut of sharded_context.  With peer memory (and env HKF_SHARD_PEER_HALO != 0) the
    halo planes are READ in the neighbours' memory by the slab raycast (hkf_shard_peer_halo) instead of being stored
    and integrated a second time: the cut then balances the owned planes only and keeps every slab at least `halo`
    planes thick (a ray's halo reads must not reach past the adjacent rank); if the volume is too small for that the
    stored-halo layout is used.  Every rank computes the same answer from the same arguments."""
    R = cfg.resolution
    peer = bool(p2p) and nranks > 1 and os.environ.get("HKF_SHARD_PEER_HALO", "1") != "0" and R // nranks >= max(halo, 8)
    if peer:
        if balanced:
            rs = balanced_slab_ranges(cfg, nranks, 0, min_planes=max(halo, 8), plane_cost=plane_cost)
        else:
            rs = slab_ranges(R, nranks, 0)
        if min(z1 - z0 for z0, z1, _, _ in rs) >= halo:
            return rs, True
    rs = balanced_slab_ranges(cfg, nranks, halo, plane_cost=plane_cost) if balanced else slab_ranges(R, nranks, halo)
    return rs, False


def icp_band(rows: int, rank: int, nranks: int):
    """pixel rows [y0, y1) of an ICP level evaluated by `rank` (csrc/icp.cu, hkf_icp_track)"""
    return rows * rank // nranks, rows * (rank + 1) // nranks


def pack_key(t: np.ndarray, ok: np.ndarray, rank: int) -> np.ndarray:
    """(t | ok | rank) compositing key of csrc/raycast.cu: float bits of the crossing's sample parameter in the high
    word (positive floats order like their bit patterns), refinement success in bit 8, rank in the low byte"""
    bits = np.asarray(t, np.float32).view(np.uint32).astype(np.uint64)
    return (bits << np.uint64(32)) | (np.asarray(ok, np.uint64) << np.uint64(8)) | np.uint64(rank)


def composite(keys: np.ndarray, vertices: np.ndarray, normals: np.ndarray):
    """Single-process statement of the compositing rule: keys [G][H][W] (NO_HIT where a rank reports nothing),
    payloads [G][H][W][4].  The minimum key wins, also when its refinement failed (the reference's ray ends at its
    first crossing, src/cuda/raycastingVolume.cu:96-108); a failed winner's payload is zero."""
    best = keys.min(axis=0)
    win = (keys == best[None]) & (keys != NO_HIT)
    assert (win.sum(axis=0) <= 1).all(), "two ranks claim the same crossing"
    v = (vertices * win[..., None]).sum(axis=0, dtype=np.float32)
    n = (normals * win[..., None]).sum(axis=0, dtype=np.float32)
    return best, v, n


def composite_distributed(keys: np.ndarray, vertices: np.ndarray, normals: np.ndarray):
    """The same rule with collectives, as csrc/shard.cu does it: all_reduce(MIN) on the keys, losers zero their
    payload, all_reduce(SUM) on the payload.  Works on any torch.distributed backend (gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    k = torch.from_numpy(keys.view(np.int64).copy())          # keys < 2^63 except NO_HIT, which maps to -1
    k[k < 0] = torch.iinfo(torch.int64).max
    best = k.clone()
    dist.all_reduce(best, op=dist.ReduceOp.MIN)
    win = (k == best) & (k != torch.iinfo(torch.int64).max)
    pay = torch.from_numpy(np.concatenate([vertices, normals], axis=-1).astype(np.float32)) * win[..., None]
    dist.all_reduce(pay, op=dist.ReduceOp.SUM)
    out = best.numpy().copy()
    out[out == np.iinfo(np.int64).max] = -1
    return out.view(np.uint64), pay.numpy()[..., :4], pay.numpy()[..., 4:]


IPC_BYTES = 256          # HKF_SHARD_IPC_BYTES


def all_gather_bytes(mine: bytes) -> bytes:
    """concatenation, in rank order, of every rank's equally sized byte string"""
    import torch
    import torch.distributed as dist
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)


def broadcast_unique_id(make_id, src: int = 0) -> bytes:
    """rank `src` creates the 128-byte NCCL unique id (hkf_shard_unique_id), everybody receives it"""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank() == src:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def sharded_context(cfg, rank: int, nranks: int, device: int, p2p=None, balanced: bool = True, plane_cost=None):
    """hkf context of `rank`'s slab, joined to the library's own NCCL communicator.  Needs an initialised
    torch.distributed process group (any backend) to hand the unique id round."""
    from . import capi
    halo = halo_planes(cfg.ray_increment, cfg.volume_size, cfg.resolution)
    if p2p is None:
        p2p = os.environ.get("HKF_SHARD_P2P", "1") != "0" and nranks <= 8
    ranges, peer_halo = plan_slabs(cfg, nranks, halo, p2p, balanced, plane_cost)
    z0, z1, _, _ = ranges[rank]
    c = dataclasses.replace(cfg, slab_z0=z0, slab_z1=z1, slab_halo=0 if peer_halo else halo)
    ctx = capi.Context(c, device=device)

    def make_id():
        b = (C.c_char * 128)()
        ctx._ck(ctx.lib.hkf_shard_unique_id(b), "hkf_shard_unique_id")
        return bytes(b)

    uid = broadcast_unique_id(make_id)
    ctx._ck(ctx.lib.hkf_shard_init(ctx._h, rank, nranks, uid), "hkf_shard_init")
    if p2p:
        h = (C.c_char * IPC_BYTES)()
        ctx._ck(ctx.lib.hkf_shard_ipc_handles(ctx._h, h), "hkf_shard_ipc_handles")
        allh = all_gather_bytes(bytes(h))
        ctx._ck(ctx.lib.hkf_shard_connect(ctx._h, allh, nranks), "hkf_shard_connect")
        if peer_halo:
            own = (C.c_uint * (2 * nranks))(*[v for r_ in ranges for v in r_[:2]])
            ctx._ck(ctx.lib.hkf_shard_peer_halo(ctx._h, halo, own, nranks), "hkf_shard_peer_halo")
    ctx.peer_halo = bool(peer_halo)
    return ctx


def gather_plane_profile(ctx, cfg, rank: int, nranks: int, device: int) -> np.ndarray:
    """[R][2] work profile of the whole volume: every rank contributes the counters of the planes it OWNS (halo planes
    are profiled by their owners), summed over the ranks"""
    import torch
    import torch.distributed as dist
    R = cfg.resolution
    mine = ctx.profile_read().astype(np.int64)
    z_lo = max(ctx.cfg.slab_z0 - ctx.cfg.slab_halo, 0)
    full = np.zeros((R, 2), np.int64)
    full[ctx.cfg.slab_z0:ctx.cfg.slab_z1] = mine[ctx.cfg.slab_z0 - z_lo:ctx.cfg.slab_z1 - z_lo]
    t = torch.from_numpy(full)
    if dist.get_backend() == "nccl":
        t = t.to(f"cuda:{device}")
    dist.all_reduce(t)
    return t.cpu().numpy()
