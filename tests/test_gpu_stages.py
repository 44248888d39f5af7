"""GPU parity tests, stage by stage, through the C ABI (include/hkf.h).

Every stage of the CUDA path is fed the golden INPUT of that stage (outputs of the reference's own
kernels, tests/golden/stages_small.npz) and must reproduce the golden OUTPUT bit for bit — including the
bilateral filter, since __expf is the same hardware instruction on both sides.  Then the same stages
are compared with the CPU oracle on other seeded inputs, and with the unmodified reference kernels
(oracle/_ref/libhkf_ref.so) at BASELINE sizes when that library travelled to the box.
"""
import numpy as np
import pytest

import _oracle as O
from conftest import bits_equal, n_mismatch, small_config, canon_tris
from hybkinectfu_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx64():
    c = capi.Context(small_config(64))
    yield c
    c.close()


def test_preprocess_wrappers_bit_exact(ctx64, stages):
    c = ctx64
    c.upload_depth_mm(stages["frames"][0])
    c.trunc_depth(); c.bilateral(); c.vertices(); c.normals(); c.downsample_new()
    assert bits_equal(c.download(capi.BUF_RAW_DEPTH), stages["raw"])
    assert bits_equal(c.download(capi.BUF_TRUNC_DEPTH), stages["trunc"])
    assert bits_equal(c.download(capi.BUF_FILTERED_DEPTH), stages["filt"])
    for l in range(3):
        assert bits_equal(c.download(capi.BUF_NEW_VERTICES, l), stages[f"new_v{l}"]), l
        assert bits_equal(c.download(capi.BUF_NEW_NORMALS, l), stages[f"new_n{l}"]), l


def test_preprocess_fused_bit_exact(ctx64, stages):
    c = ctx64
    c.preprocess(stages["frames"][0])
    assert bits_equal(c.download(capi.BUF_RAW_DEPTH), stages["raw"])
    assert bits_equal(c.download(capi.BUF_TRUNC_DEPTH), stages["trunc"])
    assert bits_equal(c.download(capi.BUF_FILTERED_DEPTH), stages["filt"])
    for l in range(3):
        assert bits_equal(c.download(capi.BUF_NEW_VERTICES, l), stages[f"new_v{l}"]), l
        assert bits_equal(c.download(capi.BUF_NEW_NORMALS, l), stages[f"new_n{l}"]), l


def test_integration_bit_exact(ctx64, stages):
    c = ctx64
    c.reset_volume()
    for i in range(3):
        c.preprocess(stages["frames"][i])
        n = c.count_updates(stages["poses"][i])
        before = c.download(capi.BUF_VOLUME)
        c.integrate(stages["poses"][i])
        vol = c.download(capi.BUF_VOLUME)
        g = stages[f"vol_after{i}"]
        assert bits_equal(vol[..., 1], g[..., 1]), f"update mask / weights differ after frame {i}"
        assert bits_equal(vol[..., 0], g[..., 0]), f"tsdf differs after frame {i}"
        # N_upd reported for the roofline equals the number of voxels whose weight or tsdf changed or
        # was rewritten: every updated voxel gains weight until the cap (128), so count weight changes
        assert n == int((vol[..., 1] != before[..., 1]).sum())


def test_raycast_bit_exact(ctx64, stages):
    c = ctx64
    c.upload(capi.BUF_VOLUME, stages["vol_after2"])
    c.raycast(stages["poses"][2])
    c.downsample_model()
    for l in range(3):
        assert bits_equal(c.download(capi.BUF_MODEL_VERTICES, l), stages[f"model_v{l}"]), l
        assert bits_equal(c.download(capi.BUF_MODEL_NORMALS, l), stages[f"model_n{l}"]), l


def test_icp_sums_bit_exact(ctx64, stages):
    c = ctx64
    for l in range(3):
        c.upload(capi.BUF_NEW_VERTICES, stages[f"icp_new_v{l}"], l)
        c.upload(capi.BUF_NEW_NORMALS, stages[f"icp_new_n{l}"], l)
        c.upload(capi.BUF_MODEL_VERTICES, stages[f"model_v{l}"], l)
        c.upload(capi.BUF_MODEL_NORMALS, stages[f"model_n{l}"], l)
    for l in range(3):
        s = c.icp_params(l, stages["icp_cur"], stages["icp_last_inv"])
        assert n_mismatch(s, stages[f"icp_sums{l}"]) == 0, (l, s, stages[f"icp_sums{l}"])


def test_sdf_sums_bit_exact(ctx64, stages):
    c = ctx64
    c.upload(capi.BUF_VOLUME, stages["vol_after2"])
    c.upload(capi.BUF_TRUNC_DEPTH, stages["icp_trunc"])
    s = c.sdf_params(stages["icp_cur"])
    assert n_mismatch(s, stages["sdf_sums"]) == 0, (s, stages["sdf_sums"])


def test_marching_cubes_multiset_exact(ctx64, stages):
    c = ctx64
    c.upload(capi.BUF_VOLUME, stages["vol_after2"])
    n = c.marching_cubes()
    assert n == int(stages["mc_count"][0])
    tris = c.download(capi.BUF_TRIANGLES)
    assert len(tris) == n
    assert bits_equal(canon_tris(tris), stages["mc_tris_sorted"])
    # the counter is cleared per call (the reference only clears at allocation)
    assert c.marching_cubes() == n


# ---------------------------------------------------------------- against the CPU oracle, other inputs
def test_against_cpu_oracle_other_frames():
    cfg = small_config(128)
    cam = synth.Camera().scaled(192, 144)
    poses = synth.trajectory(120, phase=0.7)[::5][:4].astype(np.float32)
    frames = np.stack([synth.render_depth_mm(p.astype(np.float64), cam, noise_sigma_mm=1.0) for p in poses])
    with capi.Context(cfg) as c:
        vol = np.zeros((128, 128, 128, 2), np.float32)
        for i in range(3):
            c.preprocess(frames[i])
            tr = O.trunc(O.decode_mm(frames[i]), cfg.trunc_min, cfg.trunc_max)
            assert bits_equal(c.download(capi.BUF_TRUNC_DEPTH), tr)
            c.integrate(poses[i])
            O.integrate(vol, cfg, tr, poses[i])
        gv = c.download(capi.BUF_VOLUME)
        assert bits_equal(gv, vol)
        c.raycast(poses[2])
        v, n = O.raycast(vol, cfg, poses[2])
        assert bits_equal(c.download(capi.BUF_MODEL_VERTICES), v)
        assert bits_equal(c.download(capi.BUF_MODEL_NORMALS), n)
        c.downsample_model()
        # ICP sums of frame 3 against that model, all levels, bit-exact (same float trees)
        c.preprocess(frames[3])
        last_inv = O.mat44_inverse(poses[2])
        for l in range(3):
            s_gpu = c.icp_params(l, poses[2], last_inv)
            s_cpu = O.icp_sums(c.download(capi.BUF_NEW_VERTICES, l), c.download(capi.BUF_NEW_NORMALS, l),
                               c.download(capi.BUF_MODEL_VERTICES, l), c.download(capi.BUF_MODEL_NORMALS, l),
                               O.cam_of(cfg, l), poses[2], last_inv, cfg.icp_dist_thres, cfg.icp_norm_sin_thres)
            assert n_mismatch(s_gpu, s_cpu) == 0, l
        s_gpu = c.sdf_params(poses[3])
        s_cpu = O.sdf_sums(vol, cfg, c.download(capi.BUF_TRUNC_DEPTH), poses[3])
        assert n_mismatch(s_gpu, s_cpu) == 0
        n_gpu = c.marching_cubes()
        n_cpu, t_cpu = O.marching_cubes(vol, cfg)
        assert n_gpu == n_cpu
        assert bits_equal(canon_tris(c.download(capi.BUF_TRIANGLES)), canon_tris(t_cpu))


def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    m[i, i], m[i, j], m[j, i], m[j, j] = c, -s, s, c
    return m


@pytest.mark.parametrize("name,resolution,size,pose_fn", [
    # camera INSIDE the volume, pitched and rolled: voxels behind the camera (Z <= 0), rows grazing the image border,
    # the ill-conditioned Z ~ 0 region of the row intervals
    ("inside", 96, 2.0, lambda: np.array([[1, 0, 0, 1.0], [0, 1, 0, 0.9], [0, 0, 1, 0.45], [0, 0, 0, 1.0]]) @ _rot(0, 0.35) @ _rot(2, 0.6) @ _rot(1, -0.4)),
    # the reference's default volume: 3 m (R/S is not a power of two -> the raycast's exact-division fallback), odd-ish R
    ("size3", 100, 3.0, lambda: synth.pose0(3.0, 0.3) @ _rot(1, 0.2)),
    # looking back out of the volume: almost nothing is in view
    ("away", 64, 2.0, lambda: synth.pose0() @ _rot(1, 2.6)),
])
def test_integration_raycast_exact_for_awkward_views(name, resolution, size, pose_fn):
    """bit-exact TSDF / raycast against the CPU oracle where the culling logic of integrate.cu is stressed"""
    import dataclasses
    cfg = dataclasses.replace(small_config(resolution), volume_size=size)
    cam = synth.Camera().scaled(192, 144)
    base = pose_fn()
    poses = [(base @ _rot(1, 0.02 * i)).astype(np.float32) for i in range(3)]
    # the scene is rendered from a pose inside the 2 m scene cube regardless of the volume size
    frames = [synth.render_depth_mm(synth.trajectory(60)[5 * i], cam, noise_sigma_mm=0.5) for i in range(3)]
    with capi.Context(cfg) as c:
        vol = np.zeros((resolution, resolution, resolution, 2), np.float32)
        total = 0
        for i in range(3):
            c.preprocess(frames[i])
            tr = O.trunc(O.decode_mm(frames[i]), cfg.trunc_min, cfg.trunc_max)
            n_gpu = c.count_updates(poses[i])
            c.integrate(poses[i])
            n_cpu = O.integrate(vol, cfg, tr, poses[i])
            assert n_gpu == n_cpu, (name, i, n_gpu, n_cpu)
            total += n_cpu
            assert bits_equal(c.download(capi.BUF_VOLUME), vol), (name, i)
        if name != "away":
            assert total > 1000
        c.raycast(poses[2])
        v, n = O.raycast(vol, cfg, poses[2])
        assert bits_equal(c.download(capi.BUF_MODEL_VERTICES), v), name
        assert bits_equal(c.download(capi.BUF_MODEL_NORMALS), n), name


# ---------------------------------------------------------------- edge cases
def test_empty_and_degenerate_frames():
    cfg = small_config(64)
    with capi.Context(cfg) as c:
        # all-invalid depth: nothing integrates, raycast empty, ICP sums zero, MC empty
        z = np.zeros((cfg.rows, cfg.cols), np.uint16)
        c.preprocess(z)
        p0 = synth.pose0().astype(np.float32)
        assert c.count_updates(p0) == 0
        c.integrate(p0)
        assert not c.download(capi.BUF_VOLUME).any()
        c.raycast(p0)
        assert not c.download(capi.BUF_MODEL_VERTICES).any()
        assert c.marching_cubes() == 0
        s = c.icp_params(0, p0, O.mat44_inverse(p0))
        assert not s.any()
        # depth beyond both truncations (65535 mm) and below the minimum (100 mm)
        far = np.full((cfg.rows, cfg.cols), 65535, np.uint16)
        far[:, ::2] = 100
        c.preprocess(far)
        assert not c.download(capi.BUF_TRUNC_DEPTH).any()
        # a camera looking away from the volume integrates nothing
        away = p0.copy(); away[0, 0] = -1; away[2, 2] = -1
        frames = synth.render_depth_mm(synth.pose0(), synth.Camera().scaled(192, 144))
        c.preprocess(frames)
        assert c.count_updates(away) == 0


def test_weight_saturates_at_max():
    cfg = small_config(64)
    cfg.weight_max = 3.0
    cam = synth.Camera().scaled(192, 144)
    f = synth.render_depth_mm(synth.pose0(), cam)
    p0 = synth.pose0().astype(np.float32)
    with capi.Context(cfg) as c:
        vol = np.zeros((64, 64, 64, 2), np.float32)
        tr = O.trunc(O.decode_mm(f), cfg.trunc_min, cfg.trunc_max)
        c.preprocess(f)
        for _ in range(5):
            c.integrate(p0)
            O.integrate(vol, cfg, tr, p0)
        g = c.download(capi.BUF_VOLUME)
        assert g[..., 1].max() == 3.0
        assert bits_equal(g, vol)


def test_abi_argument_errors():
    cfg = small_config(64)
    with capi.Context(cfg) as c:
        lib, h = c.lib, c._h
        import ctypes as C
        m = capi.Mat44.from_np(np.eye(4))
        ip = capi.IntegrateParams(0.04, 2.0)
        cam = cfg.camera()
        assert lib.hkf_integrate_volume(h, 1, 0, C.byref(m), C.byref(ip), C.byref(cam), C.byref(cam)) == 4   # colour on a context created without use_color: HKF_ERR_STATE
        bad = capi.CameraParams(10, 10, 1, 1, 1, 1)
        assert lib.hkf_integrate_volume(h, 0, 0, C.byref(m), C.byref(ip), C.byref(bad), C.byref(bad)) == 1
        assert lib.hkf_cal_point_to_plane_err_solver_params(h, 7, C.byref(m), C.byref(m), C.byref(cam), 0.1, 0.2) == 1
        assert lib.hkf_download(h, 99, 0, None, 0) < 0
    bad_cfg = small_config(64); bad_cfg.pyramid_levels = 4
    with pytest.raises(capi.HkfError):
        capi.Context(bad_cfg)


# ---------------------------------------------------------------- against the unmodified reference kernels at BASELINE size
@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libhkf_ref.so not present")
def test_against_reference_kernels_640x480_256():
    cfg = capi.Config(resolution=256, pyramid_levels=3, max_triangles=3000000)
    cam = synth.Camera()
    poses = synth.trajectory(100)[:4].astype(np.float32)
    frames = np.stack([synth.render_depth_mm(p.astype(np.float64), cam) for p in poses])
    ref = O.Ref(cfg)
    ref.reset_volume()
    with capi.Context(cfg) as c:
        for i in range(3):
            c.preprocess(frames[i]); ref.preprocess(frames[i])
            if i == 0:
                assert bits_equal(c.download(capi.BUF_FILTERED_DEPTH), ref.download(O.REF_FILTERED_DEPTH))
                for l in range(3):
                    assert bits_equal(c.download(capi.BUF_NEW_VERTICES, l), ref.download(O.REF_NEW_VERTICES, l)), l
                    assert bits_equal(c.download(capi.BUF_NEW_NORMALS, l), ref.download(O.REF_NEW_NORMALS, l)), l
            c.integrate(poses[i]); ref.integrate(poses[i])
        gv, rv = c.download(capi.BUF_VOLUME), ref.download(O.REF_VOLUME)
        assert bits_equal(gv[..., 1], rv[..., 1]), "update mask differs from the reference kernels"
        assert bits_equal(gv[..., 0], rv[..., 0])
        c.raycast(poses[2]); ref.raycast(poses[2])
        c.downsample_model(); ref.downsample_model()
        for l in range(3):
            assert bits_equal(c.download(capi.BUF_MODEL_VERTICES, l), ref.download(O.REF_MODEL_VERTICES, l)), l
            assert bits_equal(c.download(capi.BUF_MODEL_NORMALS, l), ref.download(O.REF_MODEL_NORMALS, l)), l
        c.preprocess(frames[3]); ref.preprocess(frames[3])
        last_inv = O.mat44_inverse(poses[2])
        for l in range(3):
            assert n_mismatch(c.icp_params(l, poses[2], last_inv), ref.icp_params(l, poses[2], last_inv)) == 0, l
        assert n_mismatch(c.sdf_params(poses[3]), ref.sdf_params(poses[3])) == 0
        n = c.marching_cubes()
        assert n == ref.marching_cubes()
        assert bits_equal(canon_tris(c.download(capi.BUF_TRIANGLES)), canon_tris(ref.download(O.REF_TRIANGLES)))
    ref.check()
