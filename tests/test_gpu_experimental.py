"""Device paths that are OFF by default and have not been measured yet (written without GPU time left in the round).
They only run with HKF_TEST_EXPERIMENTAL=1, each in a subprocess that sets the environment switch of the path, and
must reproduce the default path bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("HKF_TEST_EXPERIMENTAL") != "1", reason="experimental paths: set HKF_TEST_EXPERIMENTAL=1")]

_TRACK = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from conftest import small_config
from hybkinectfu_b200 import capi, synth
cfg = small_config(128)
cam = synth.Camera().scaled(192, 144)
gt = synth.trajectory(200)[::4][:8]
frames = [synth.render_depth_mm(p, cam) for p in gt]
ctx = capi.Context(cfg)
ctx.set_pose(synth.pose0().astype(np.float32))
poses, maps = [], []
for i, mm in enumerate(frames):
    ctx.preprocess(mm)
    if i:
        ctx.icp_track()
    ctx.integrate_tracked(); ctx.raycast_tracked()
    p, st = ctx.frame_result()
    assert st == 0, st
    poses.append(p)
    maps.append(np.concatenate([ctx.download(capi.BUF_MODEL_VERTICES).ravel(), ctx.download(capi.BUF_MODEL_NORMALS).ravel()]))
np.save(sys.argv[1], np.stack(poses))
np.save(sys.argv[1] + ".maps.npy", np.stack(maps))
np.save(sys.argv[1] + ".vol.npy", ctx.download(capi.BUF_VOLUME))
"""


def _poses(tmp_path, name, env):
    out = str(tmp_path / (name + ".npy"))
    r = subprocess.run([sys.executable, "-c", _TRACK.format(root=ROOT), out], env=dict(os.environ, **env),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out), np.load(out + ".maps.npy"), np.load(out + ".vol.npy")


@pytest.mark.parametrize("cs", ["8", "16"])
def test_cluster_icp_equals_per_iteration_launches(tmp_path, cs):
    ref, _, _ = _poses(tmp_path, "default", {"HKF_ICP_CLUSTER": "0"})
    got, _, _ = _poses(tmp_path, "cluster" + cs, {"HKF_ICP_CLUSTER": cs})
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))


def test_raycast_empty_space_skipping_is_exact(tmp_path):
    """HKF_RAY_SKIP=1: brick flags from the integration item pass + skipped samples / batches in the ray march; the model
    maps of every frame (and therefore the poses) must not change by a bit (CPU model: tests/test_raycast_skip_model.py)"""
    ref, ref_maps, ref_vol = _poses(tmp_path, "default", {"HKF_RAY_SKIP": "0"})
    got, got_maps, got_vol = _poses(tmp_path, "skip", {"HKF_RAY_SKIP": "1"})
    assert np.array_equal(ref_maps.view(np.uint32), got_maps.view(np.uint32))
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
    assert np.array_equal(ref_vol.view(np.uint32), got_vol.view(np.uint32))


def test_wider_tile_span_classifier_is_exact(tmp_path):
    """HKF_INT_SPAN=3: the integration item pass tests a chunk against up to four finer depth tiles per tile row; the
    classification stays conservative, so the volume (and everything downstream) must not change by a bit"""
    ref, ref_maps, ref_vol = _poses(tmp_path, "default", {"HKF_INT_SPAN": "1"})
    got, got_maps, got_vol = _poses(tmp_path, "span3", {"HKF_INT_SPAN": "3"})
    assert np.array_equal(ref_vol.view(np.uint32), got_vol.view(np.uint32))
    assert np.array_equal(ref_maps.view(np.uint32), got_maps.view(np.uint32))
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))


def test_split_preprocess_is_exact(tmp_path):
    """HKF_PRE_V2=1: tile kernel without the normals halo + one launch for vertices, normals and both pyramid levels
    (warp-shuffle 2x2 / 4x4 blocks); every downstream bit must stay the same"""
    ref, ref_maps, ref_vol = _poses(tmp_path, "default", {"HKF_PRE_V2": "0"})
    got, got_maps, got_vol = _poses(tmp_path, "prev2", {"HKF_PRE_V2": "1"})
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
    assert np.array_equal(ref_vol.view(np.uint32), got_vol.view(np.uint32))
    assert np.array_equal(ref_maps.view(np.uint32), got_maps.view(np.uint32))


def test_split_preprocess_maps_equal_stage_by_stage():
    """the new-frame maps of every pyramid level, default vs HKF_PRE_V2=1, in two subprocesses (the switch is read once)"""
    script = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from conftest import small_config
from hybkinectfu_b200 import capi, synth
cfg = small_config(64)
cam = synth.Camera().scaled(192, 144)
mm = synth.render_depth_mm(synth.trajectory(200)[7], cam, noise_sigma_mm=1.0, rng=np.random.default_rng(3))
ctx = capi.Context(cfg)
ctx.preprocess(mm)
out = [ctx.download(b).ravel() for b in (capi.BUF_RAW_DEPTH, capi.BUF_TRUNC_DEPTH, capi.BUF_FILTERED_DEPTH)]
for l in range(3):
    out += [ctx.download(capi.BUF_NEW_VERTICES, l).ravel(), ctx.download(capi.BUF_NEW_NORMALS, l).ravel()]
np.save(sys.argv[1], np.concatenate(out))
"""
    import tempfile
    res = []
    for v in ("0", "1"):
        with tempfile.TemporaryDirectory() as d:
            out = os.path.join(d, "o.npy")
            r = subprocess.run([sys.executable, "-c", script.format(root=ROOT), out], env=dict(os.environ, HKF_PRE_V2=v),
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            res.append(np.load(out))
    assert np.array_equal(res[0].view(np.uint32), res[1].view(np.uint32))
