"""Exactness of the empty-space-skipping ray march planned for csrc/raycast.cu, checked on the CPU: the model
(oracle/hkf_oracle.c: orc_raycast_skip) must reproduce orc_raycast bit for bit, and its instrumentation must count zero
violations of the two claims the skipping rests on (a sample in an unflagged brick has tsdf >= 0; the samples of a batch
lie in the voxel box of its first and last sample)."""
import numpy as np
import pytest

import _oracle as O
from conftest import bits_equal, small_config


def _volume(res, nframes=6):
    from hybkinectfu_b200 import synth
    cfg = small_config(res)
    cam = synth.Camera().scaled(cfg.cols, cfg.rows)
    poses = synth.trajectory(200)[::5][:nframes]
    vol = np.zeros((res, res, res, 2), np.float32)
    for p in poses:
        mm = synth.render_depth_mm(p, cam)
        d = O.trunc(O.decode_mm(mm), cfg.trunc_min, cfg.trunc_max)
        O.integrate(vol, cfg, d, p.astype(np.float32))
    return cfg, vol, poses


@pytest.mark.parametrize("res,fine,coarse,batch", [(64, 8, 16, 4), (128, 16, 32, 8), (128, 8, 64, 8)])
def test_skip_model_equals_reference_march(res, fine, coarse, batch):
    cfg, vol, poses = _volume(res)
    ff, fc = O.ray_flags(vol, res, fine), O.ray_flags(vol, res, coarse)
    assert ff.any() and not ff.all()
    views = [poses[0], poses[-1]]
    # an awkward view: from outside the volume, rotated about two axes
    a, b = 0.4, -0.3
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    T = np.eye(4); T[:3, :3] = Ry @ Rx; T[:3, 3] = [0.2, 1.4, -0.5]
    views.append(T)
    skipped = 0
    for pose in views:
        pose = np.asarray(pose, np.float32)
        v0, n0 = O.raycast(vol, cfg, pose)
        v1, n1, st = O.raycast_skip(vol, cfg, pose, ff, fine, fc, coarse, batch)
        assert st[3] == 0, f"claims violated {st[3]} times"
        assert bits_equal(v0, v1) and bits_equal(n0, n1)
        assert st[1] + st[2] <= st[0]
        skipped += st[1] + st[2]
    assert skipped > 0


def test_all_flags_set_is_the_plain_march():
    cfg, vol, poses = _volume(64, 3)
    ones_f, ones_c = np.ones((8, 8, 8), np.uint8), np.ones((4, 4, 4), np.uint8)
    pose = np.asarray(poses[1], np.float32)
    v0, n0 = O.raycast(vol, cfg, pose)
    v1, n1, st = O.raycast_skip(vol, cfg, pose, ones_f, 8, ones_c, 16, 8)
    assert bits_equal(v0, v1) and bits_equal(n0, n1) and st[1] == 0 and st[2] == 0 and st[3] == 0
