"""z-slab mode without a second GPU: the slabs of one volume are separate contexts on cuda:0, composited on the host
with hybkinectfu_b200.shard.composite (the rule csrc/shard.cu applies over NVLink).  Covers the SLAB variants of the
integration, raycast and marching-cubes kernels in a single-GPU run; tests/test_gpu_sharded.py is the multi-process
version with the device-side exchange."""
import dataclasses

import numpy as np
import pytest

from conftest import bits_equal, canon_tris, small_config

pytestmark = pytest.mark.gpu


def _frames(n=5):
    from hybkinectfu_b200 import synth
    cam = synth.Camera().scaled(192, 144)
    poses = synth.trajectory(200)[::4][:n].astype(np.float32)
    return np.stack([synth.render_depth_mm(p.astype(np.float64), cam) for p in poses]), poses


def _run(ctx, frames, gt):
    from hybkinectfu_b200 import capi
    ctx.set_pose(gt[0])
    for i, mm in enumerate(frames):
        ctx.preprocess(mm)
        ctx.set_pose(gt[i])
        ctx.integrate_tracked()
        ctx.raycast_tracked()
    ctx.frame_result()
    out = {"v": ctx.download(capi.BUF_MODEL_VERTICES), "n": ctx.download(capi.BUF_MODEL_NORMALS),
           "vol": ctx.download(capi.BUF_VOLUME)}
    n = ctx.marching_cubes()
    out["tris"] = ctx.download(capi.BUF_TRIANGLES)[:n]
    return out


@pytest.mark.parametrize("res,nslabs,balanced", [(128, 2, False), (256, 3, True)])
def test_slab_contexts_composite_to_the_whole_volume(res, nslabs, balanced):
    from hybkinectfu_b200 import capi, shard
    cfg = small_config(res)
    frames, gt = _frames()
    with capi.Context(cfg) as c:
        whole = _run(c, frames, gt)
    assert (whole["n"][..., :3] != 0).any(-1).sum() > 1000 and len(whole["tris"]) > 1000
    halo = shard.halo_planes(cfg.ray_increment, cfg.volume_size, cfg.resolution)
    ranges = shard.balanced_slab_ranges(cfg, nslabs, halo) if balanced else shard.slab_ranges(res, nslabs, halo)
    keys, vs, ns, tris = [], [], [], []
    for r, (z0, z1, s0, s1) in enumerate(ranges):
        with capi.Context(dataclasses.replace(cfg, slab_z0=z0, slab_z1=z1, slab_halo=halo)) as c:
            part = _run(c, frames, gt)
            keys.append(c.download(capi.BUF_RAYCAST_KEYS))
        assert part["vol"].shape[0] == s1 - s0
        assert bits_equal(part["vol"], whole["vol"][s0:s1]), f"slab {r} differs from planes [{s0},{s1}) of the whole volume"
        vs.append(part["v"]); ns.append(part["n"]); tris.append(part["tris"])
    best, v, n = shard.composite(np.stack(keys), np.stack(vs), np.stack(ns))
    assert bits_equal(v, whole["v"]) and bits_equal(n, whole["n"])
    # every cell is extracted by exactly one slab (the owner of its z plane)
    assert sum(len(t) for t in tris) == len(whole["tris"])
    assert np.array_equal(canon_tris(np.concatenate(tris)).view(np.uint32), canon_tris(whole["tris"]).view(np.uint32))
