"""Parity at the sizes BASELINE.json's configs are quoted on (the other GPU tests use 192x144 images and <= 256^3 volumes so
that the CPU oracle finishes in seconds): every comparison here is against the reference's own UNMODIFIED kernels
(oracle/_ref/libhkf_ref.so, driven by oracle/ref_harness.cu) running on the same GPU.

  C2  the whole 500-frame periodic trajectory 640x480 -> 512^3, fused ICP pipeline vs the reference pipeline frame by frame (pose tolerance of
      north_star: 1e-5 rotation, 1e-4 m translation; status equal), final volumes compared voxel by voxel
  C3  SDF tracking at 512^3 frame by frame, then marching cubes on the SAME volume bits: count + sorted multiset exact
  C5  every stage at 1280x960 (where the reference's own pyramid wrappers cannot launch: the fenced kernels define it)
  C4  1024^3 on one GPU: integration planes, raycast maps and the ICP sums bit for bit
"""
import numpy as np
import pytest

import _oracle as O
from conftest import bits_equal, canon_tris, n_mismatch
from hybkinectfu_b200 import capi, synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libhkf_ref.so not present")]

ROT_TOL, TRANS_TOL = 1e-5, 1e-4          # north_star: per-frame pose tolerance


def _sequence(n, cols=640, rows=480, stride=1):
    cam = synth.Camera() if (cols, rows) == (640, 480) else synth.Camera().scaled(cols, rows)
    poses = synth.trajectory(500)[::stride][:n]
    return np.stack([synth.render_depth_mm(p, cam) for p in poses]), poses.astype(np.float32), cam


def _pose_delta(p, q):
    return float(np.abs(p[:3, :3] - q[:3, :3]).max()), float(np.abs(p[:3, 3] - q[:3, 3]).max())


def _volume_agreement(ours, ref):
    """two volumes built from poses that agree to ~1e-7: identical weights almost everywhere, tsdf within 1e-4"""
    same_w = ours[..., 1] == ref[..., 1]
    observed = int((ref[..., 1] > 0).sum())
    frac_w = float((~same_w).sum()) / max(observed, 1)
    d = np.abs(ours[..., 0] - ref[..., 0])[same_w]
    return observed, frac_w, float(d.max()), float((d > 1e-4).sum()) / max(observed, 1)


def test_c2_500_frames_512_fused_pipeline_vs_reference_pipeline():
    """the pose tolerance measured over the full 500-frame sequence, against the reference's poses"""
    cfg = capi.Config(resolution=512, pyramid_levels=3, max_triangles=1)
    frames, gt, _ = _sequence(500)
    ref = O.Ref(cfg)
    ref.reset_volume()
    worst = [0.0, 0.0]
    with capi.Context(cfg) as c:
        c.set_pose(synth.pose0().astype(np.float32))
        for i in range(len(frames)):
            c.preprocess(frames[i])
            if i:
                c.icp_track()
            c.integrate_tracked(); c.raycast_tracked()
            p, st = c.frame_result()
            q, st2 = ref.process(frames[i], i, 0)
            assert st == st2 == 0, (i, st, st2)
            dr, dt = _pose_delta(p, q)
            worst = [max(worst[0], dr), max(worst[1], dt)]
            assert dr <= ROT_TOL and dt <= TRANS_TOL, (i, dr, dt)
        # both track the analytic trajectory
        assert np.abs(p[:3, 3] - gt[-1][:3, 3]).max() < 5e-3
        ours, theirs = c.download(capi.BUF_VOLUME), ref.download(O.REF_VOLUME)
    ref.check()
    observed, frac_w, dmax, frac_d = _volume_agreement(ours, theirs)
    print(f"C2 500 frames: worst pose delta rot {worst[0]:.2e} trans {worst[1]:.2e}; observed voxels {observed}, "
          f"weight differs on {frac_w:.2e} of them, max |dtsdf| {dmax:.2e}, |dtsdf| > 1e-4 on {frac_d:.2e}")
    assert observed > 20_000_000
    assert frac_w < 1e-3 and frac_d < 1e-3


def test_c3_sdf_tracking_512_and_marching_cubes():
    cfg = capi.Config(resolution=512, pyramid_levels=3, max_triangles=3_000_000)
    frames, gt, _ = _sequence(24)
    ref = O.Ref(cfg)
    ref.reset_volume()
    with capi.Context(cfg) as c:
        c.set_pose(synth.pose0().astype(np.float32))
        for i in range(len(frames)):
            c.preprocess(frames[i])
            if i:
                c.sdf_track()
            c.integrate_tracked(); c.raycast_tracked()
            p, st = c.frame_result()
            q, st2 = ref.process(frames[i], i, 1)
            assert st == st2 == 0, (i, st, st2)
            dr, dt = _pose_delta(p, q)
            assert dr <= ROT_TOL and dt <= TRANS_TOL, (i, dr, dt)
        # marching cubes on the same bits: hand the reference's volume to our context
        vol = ref.download(O.REF_VOLUME)
        c.upload(capi.BUF_VOLUME, vol)
        n = c.marching_cubes()
        n_ref = ref.marching_cubes()
        assert n == n_ref and n > 300_000, (n, n_ref)
        assert bits_equal(canon_tris(c.download(capi.BUF_TRIANGLES)), canon_tris(ref.download(O.REF_TRIANGLES)))
    ref.check()


def test_c5_stages_at_1280x960_vs_fenced_reference_kernels():
    cam = synth.Camera().scaled(1280, 960)
    cfg = capi.Config(cols=1280, rows=960, cx=cam.cx, cy=cam.cy, fx=cam.fx, fy=cam.fy, resolution=512, pyramid_levels=3,
                      max_triangles=1)
    frames, poses, _ = _sequence(4, 1280, 960, stride=3)
    ref = O.Ref(cfg)
    ref.reset_volume()
    with capi.Context(cfg) as c:
        for i in range(3):
            c.preprocess(frames[i]); ref.preprocess(frames[i])
            if i == 0:
                for b, rb in ((capi.BUF_RAW_DEPTH, O.REF_RAW_DEPTH), (capi.BUF_TRUNC_DEPTH, O.REF_TRUNC_DEPTH), (capi.BUF_FILTERED_DEPTH, O.REF_FILTERED_DEPTH)):
                    assert bits_equal(c.download(b), ref.download(rb)), b
                for l in range(3):
                    assert bits_equal(c.download(capi.BUF_NEW_VERTICES, l), ref.download(O.REF_NEW_VERTICES, l)), l
                    assert bits_equal(c.download(capi.BUF_NEW_NORMALS, l), ref.download(O.REF_NEW_NORMALS, l)), l
            c.integrate(poses[i]); ref.integrate(poses[i])
        gv, rv = c.download(capi.BUF_VOLUME), ref.download(O.REF_VOLUME)
        assert (rv[..., 1] > 0).sum() > 20_000_000
        assert bits_equal(gv[..., 1], rv[..., 1]), "update mask differs from the reference kernels"
        assert bits_equal(gv[..., 0], rv[..., 0])
        del gv, rv
        c.raycast(poses[2]); ref.raycast(poses[2])
        c.downsample_model(); ref.downsample_model()
        for l in range(3):
            mv = ref.download(O.REF_MODEL_VERTICES, l)
            assert bits_equal(c.download(capi.BUF_MODEL_VERTICES, l), mv), l
            assert bits_equal(c.download(capi.BUF_MODEL_NORMALS, l), ref.download(O.REF_MODEL_NORMALS, l)), l
        assert (mv[..., 3] != 0).sum() > 1000
        c.preprocess(frames[3]); ref.preprocess(frames[3])
        last_inv = O.mat44_inverse(poses[2])
        for l in range(3):
            a, b = c.icp_params(l, poses[2], last_inv), ref.icp_params(l, poses[2], last_inv)
            assert np.abs(b).max() > 1 and n_mismatch(a, b) == 0, l
        assert n_mismatch(c.sdf_params(poses[3]), ref.sdf_params(poses[3])) == 0
    ref.check()


def test_c4_1024_single_gpu_planes_raycast_and_sums():
    cfg = capi.Config(resolution=1024, pyramid_levels=3, max_triangles=1)
    frames, poses, _ = _sequence(4, stride=3)
    ref = O.Ref(cfg)
    ref.reset_volume()
    with capi.Context(cfg) as c:
        for i in range(3):
            c.preprocess(frames[i]); ref.preprocess(frames[i])
            c.integrate(poses[i]); ref.integrate(poses[i])
        R = cfg.resolution
        got = c.download(capi.BUF_VOLUME)
        checked = 0
        for z0, z1 in ((0, 32), (300, 332), (690, 754), (800, 864), (992, 1024)):
            want = ref.download_volume_planes(z0, z1)
            assert bits_equal(got[z0:z1], want), (z0, z1)
            checked += int((want[..., 1] > 0).sum())
        assert checked > 10_000_000
        del got
        c.raycast(poses[2]); ref.raycast(poses[2])
        c.downsample_model(); ref.downsample_model()
        for l in range(3):
            mv = ref.download(O.REF_MODEL_VERTICES, l)
            assert bits_equal(c.download(capi.BUF_MODEL_VERTICES, l), mv), l
            assert bits_equal(c.download(capi.BUF_MODEL_NORMALS, l), ref.download(O.REF_MODEL_NORMALS, l)), l
        assert (mv[..., 3] != 0).sum() > 1000
        c.preprocess(frames[3]); ref.preprocess(frames[3])
        last_inv = O.mat44_inverse(poses[2])
        for l in range(3):
            assert n_mismatch(c.icp_params(l, poses[2], last_inv), ref.icp_params(l, poses[2], last_inv)) == 0, l
        assert n_mismatch(c.sdf_params(poses[3]), ref.sdf_params(poses[3])) == 0
    ref.check()
