"""The retained host classes (hybkinectfu_b200/host: HybKinectfu, CameraPoseFinder{ICP,SDF,FromFile},
MeshGeneratorMarchingcube, DataSourceProducerSynthetic, TrajectoryRecorder, AppParams / INI) driven through the small
extern "C" facade of libhkf_host.so.

CPU part: parameter parsing (config.ini keys of the reference), the synthetic producer against synth.py, the
TUM trajectory write -> nearest-timestamp replay round trip, mesh welding + OBJ.  GPU part: the whole
HybKinectfu::processNewFrame loop in the three tracker modes against the Python/C-ABI path and the golden
reference poses, fused and stepwise (reference control flow) modes.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, bits_equal
from hybkinectfu_b200 import capi, synth

LIB = os.path.join(ROOT, "hybkinectfu_b200", "libhkf_host.so")


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(LIB)
    lib.hkfhost_last_error.restype = C.c_char_p
    lib.hkfhost_missing_keys.restype = C.c_char_p
    lib.hkfhost_ctx.restype = C.c_void_p
    lib.hkfhost_synth_pose.argtypes = [C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.hkfhost_synth_render.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.hkfhost_write_tum.argtypes = [C.c_char_p, C.c_uint, C.c_double, C.c_double, C.c_double]
    lib.hkfhost_process.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_void_p, C.c_void_p]
    return lib


def write_ini(path, cols=192, rows=144, res=128, levels=3, sdf=0, fromfile=0, stepwise=0, record=0, traj_in="", traj_out="", mesh="", frames=200,
              synthetic=1, dataset=0, source="", color=0, angle=0):
    cam = synth.Camera().scaled(cols, rows)
    open(path, "w").write(f"""[switch]
use_color={color}
record_rgbd=0
read_rgbd_from_rgbddataset={dataset}
record_trajectory={record}
read_trajactory_from_file={fromfile}
color_angle_weight={angle}
use_sdf_tracker={sdf}
use_synthetic_data={synthetic}
stepwise_tracking={stepwise}
[camera]
rgb_cols={cols}
rgb_rows={rows}
rgb_cx={cam.cx}
rgb_cy={cam.cy}
rgb_fx={cam.fx}
rgb_fy={cam.fy}
depth_cols={cols}
depth_rows={rows}
depth_cx={cam.cx!r}
depth_cy={cam.cy!r}
depth_fx={cam.fx!r}
depth_fy={cam.fy!r}
[kinectfusion]
depth_trunc_max=4
depth_trunc_min=0.3
filter_sigma_pixel=2
filter_sigma_depth=0.02
volume_resolution={res}
volume_size_meter=2
volume_max_weight=128
integrate_sdf_trunc=0.04
integrate_depth_trunc=2
raycast_increment_factor=0.8
[icp]
icp_pyramid_level={levels}
icp_thre_dist=0.1
icp_thre_sin_angle=0.2
camera_shake_dist=0.4
camera_shake_angle=0.4
[sdfTracker]
max_iter_nums=6
camera_shake_dist=0.4
camera_shake_angle=0.4
[mesh]
max_triangle_num=400000
[IO]
rgbd_record_filename=
rgbd_source_filename={source}
trajectory_record_filename={traj_out}
trajectory_source_filename={traj_in}
mesh_filename={mesh}
[synthetic]
frames={frames}
phase=0
""")


# ------------------------------------------------------------------------------------------------ CPU
def test_params_from_reference_style_ini(host, tmp_path):
    ini = tmp_path / "config.ini"
    write_ini(str(ini), cols=640, rows=480, res=512)
    assert host.hkfhost_read_params(str(ini).encode()) == 0
    fp, ip = capi.FrameParams(), capi.InitParams()
    host.hkfhost_get_frame_params(C.byref(fp), C.byref(ip))
    assert (ip.depth_camera.cols, ip.depth_camera.rows, ip.pyramid_levels, ip.volume_resolution) == (640, 480, 3, 512)
    assert ip.depth_camera.fx == 525.0 and ip.depth_camera.cx == 319.5 and ip.volume_size == 2.0 and ip.weight_max == 128.0
    # fRayIncrement = raycast_increment_factor * integrate_sdf_trunc as a FLOAT product (src/AppParamsProducer.cpp:113-117)
    assert np.float32(fp.raycast.fRayIncrement) == np.float32(0.8) * np.float32(0.04)
    assert (fp.trunc_min, fp.trunc_max, fp.sdf_max_iter) == (np.float32(0.3), 4.0, 6)
    assert fp.icp_dist_thres == np.float32(0.1) and fp.icp_norm_sin_thres == np.float32(0.2)
    assert host.hkfhost_switches() & 16 and not host.hkfhost_switches() & 1


def test_reference_root_config_ini_quirks(host, tmp_path):
    """the reference's root config.ini lacks max_iter_nums and would crash its parser; sections must
    not inherit keys from earlier sections"""
    ini = tmp_path / "c.ini"
    ini.write_text("[switch]\nuse_color=0\n[icp]\ncamera_shake_dist=0.25\n[sdfTracker]\ncamera_shake_angle=0.5\n; comment\n# comment\n")
    assert host.hkfhost_read_params(str(ini).encode()) == 0
    fp = capi.FrameParams()
    host.hkfhost_get_frame_params(C.byref(fp), None)
    assert fp.sdf_max_iter == 6                                  # default kept
    assert fp.icp_dist_shake == np.float32(0.25)
    assert fp.sdf_dist_shake != np.float32(0.25)                 # [sdfTracker] did not inherit [icp]'s key
    assert b"max_iter_nums" in host.hkfhost_missing_keys()
    assert host.hkfhost_read_params(b"/nonexistent.ini") == 1


def test_synthetic_producer_equals_python_synth(host):
    cam = synth.Camera().scaled(96, 72)
    for i, phase in ((0, 0.0), (7, 0.0), (33, 0.7)):
        p = np.zeros(16)
        host.hkfhost_synth_pose(i, 50, phase, 2.0, 0.3, p.ctypes.data)
        want = synth.trajectory(50, phase)[i]
        assert np.abs(p.reshape(4, 4) - want).max() < 1e-12
        out = np.zeros((72, 96), np.uint16)
        host.hkfhost_synth_render(want.ctypes.data, 96, 72, cam.cx, cam.cy, cam.fx, cam.fy, out.ctypes.data)
        ref = synth.render_depth_mm(want, cam)
        # camera intrinsics cross the facade as float32: allow the last millimetre to differ on a few pixels
        d = np.abs(out.astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() < 0.01
    assert np.array_equal(synth.trajectory(50)[0], synth.pose0())
    # the colour image of the same frames (use_color=1)
    host.hkfhost_synth_render_color.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    want = synth.trajectory(50, 0.7)[33]
    img = np.zeros((72, 96, 3), np.uint8)
    host.hkfhost_synth_render_color(want.ctypes.data, 96, 72, cam.cx, cam.cy, cam.fx, cam.fy, img.ctypes.data)
    ref = synth.render_color(want, cam)
    d = np.abs(img.astype(np.int32) - ref.astype(np.int32)).max(-1)
    assert (d != 0).mean() < 0.02 and img.any()          # float32 intrinsics: a few pixels on checker / surface boundaries


def test_trajectory_record_and_replay_round_trip(host, tmp_path):
    """TrajectoryRecorder -> file -> CameraPoseFinderFromFile: pose_i = (pose0 * file0^-1) * file_i, nearest timestamp"""
    n = 12
    poses = synth.trajectory(60)[::5].astype(np.float32)
    ts = np.arange(n) / 30.0
    path = str(tmp_path / "traj.txt").encode()
    assert host.hkfhost_record_trajectory(path, poses.ctypes.data, ts.ctypes.data, n) == 0
    lines = open(path).read().splitlines()
    assert lines[0].startswith("#") and lines[2] == "# timestamp tx ty tz qx qy qz qw" and len(lines) == 3 + n
    assert len(lines[3].split()) == 8
    out = np.zeros((n, 4, 4), np.float32)
    query = ts + 0.004 * np.sin(np.arange(n))                   # jittered timestamps still pick the nearest record
    assert host.hkfhost_replay_trajectory(path, poses[0].ctypes.data, query.ctypes.data, n, out.ctypes.data) == 0
    assert bits_equal(out[0], poses[0])                          # frame 0 keeps the initial pose
    assert np.abs(out - poses).max() < 5e-6                      # 6 significant digits in the file
    # the synthetic ground-truth writer uses 9 decimals
    gt = str(tmp_path / "gt.txt").encode()
    assert host.hkfhost_write_tum(gt, 60, 0.0, 2.0, 0.3) == 0
    ts60 = np.arange(60) / 30.0
    out60 = np.zeros((60, 4, 4), np.float32)
    assert host.hkfhost_replay_trajectory(gt, synth.pose0().astype(np.float32).ctypes.data, ts60.ctypes.data, 60, out60.ctypes.data) == 0
    assert np.abs(out60 - synth.trajectory(60)).max() < 2e-6


def test_weld_and_obj(host, tmp_path):
    # two triangles sharing an edge + one exact duplicate + one degenerate
    tri = np.zeros((4, 3, 6), np.float32)
    tri[0, :, :3] = [[0, 0, 0], [1, 0, 0], [0, 1, 0]]
    tri[1, :, :3] = [[1, 0, 0], [1, 1, 0], [0, 1, 0]]
    tri[2] = tri[0]
    tri[3, :, :3] = [[2, 2, 2], [2, 2, 2], [3, 3, 3]]
    nv, nf = C.c_uint(), C.c_uint()
    obj = str(tmp_path / "m.obj")
    assert host.hkfhost_weld(tri.ctypes.data, 4, obj.encode(), C.byref(nv), C.byref(nf)) == 0
    assert (nv.value, nf.value) == (6, 2)                        # 4 mesh vertices + the 2 of the degenerate triangle
    txt = open(obj).read().splitlines()
    assert sum(l.startswith("v ") for l in txt) == 6 and sum(l.startswith("f ") for l in txt) == 2
    # vertex normals (ml::MeshData::computeVertexNormals): both faces lie in z = 0 and are counter-clockwise -> +z
    vn = np.array([[float(x) for x in l.split()[1:]] for l in txt if l.startswith("vn ")])
    assert len(vn) == 6 and np.allclose(vn[np.abs(vn).sum(1) > 0], [0, 0, 1])
    assert all("//" in l for l in txt if l.startswith("f "))
    # the same mesh as PLY and OFF
    for ext, head in (("ply", "ply"), ("off", "OFF")):
        path = str(tmp_path / f"m.{ext}")
        assert host.hkfhost_weld(tri.ctypes.data, 4, path.encode(), None, None) == 0
        lines = open(path).read().splitlines()
        assert lines[0] == head
        if ext == "ply":
            assert "element vertex 6" in lines and "element face 2" in lines and "property float nx" in lines
            body = lines[lines.index("end_header") + 1:]
            assert len(body) == 8 and body[-1].split()[0] == "3" and len(body[0].split()) == 6
        else:
            assert lines[1] == "6 2 0" and len(lines) == 2 + 8
    # order independence: shuffling the soup gives the same file
    obj2 = str(tmp_path / "m2.obj")
    assert host.hkfhost_weld(np.ascontiguousarray(tri[::-1]).ctypes.data, 4, obj2.encode(), None, None) == 0
    assert open(obj).read() == open(obj2).read()


def test_tum_dataset_reader_matches_opencv(host, tmp_path):
    """DataSourceProducerRGBDDataset: depth.txt parsing, 16-bit PNG decoding, /5 to millimetres and cv::pyrDown, checked
    against Python OpenCV (the reference uses the C++ API of the same library, src/DataSourceProducerRGBDDataset.cpp:100-125)"""
    cv2 = pytest.importorskip("cv2")
    from hybkinectfu_b200 import synth
    d = tmp_path / "ds"
    (d / "depth").mkdir(parents=True)
    cam = synth.Camera()
    poses = synth.trajectory(30)[::10]
    imgs = []
    lines = ["# depth maps", "# file: synthetic", "# timestamp filename"]
    for i, p in enumerate(poses):
        mm = synth.render_depth_mm(p, cam).astype(np.uint32)
        raw = np.minimum(mm * 5 + (i % 5), 65535).astype(np.uint16)           # TUM convention: 5000 units per metre
        assert cv2.imwrite(str(d / "depth" / f"{i}.png"), raw)
        imgs.append(raw)
        lines.append(f"{1000.5 + i / 30:.6f} depth/{i}.png")
    (d / "depth.txt").write_text("\n".join(lines) + "\n")
    host.hkfhost_read_dataset.argtypes = [C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
    # (a) full resolution
    out = np.zeros((3, 480, 640), np.uint16); ts = np.zeros(3)
    assert host.hkfhost_read_dataset((str(d) + "/").encode(), 640, 480, 3, out.ctypes.data, ts.ctypes.data) == 3, host.hkfhost_last_error()
    for i in range(3):
        want = cv2.divide(imgs[i], 5)                                        # cv::Mat /= 5 on CV_16U
        assert np.array_equal(out[i], want)
    assert np.allclose(ts, [1000.5, 1000.5 + 1 / 30, 1000.5 + 2 / 30], atol=1e-6)
    # (b) half-size depth camera -> pyrDown, as with the reference's root config.ini (depth_cols=320)
    out2 = np.zeros((3, 240, 320), np.uint16)
    assert host.hkfhost_read_dataset((str(d) + "/").encode(), 320, 240, 3, out2.ctypes.data, ts.ctypes.data) == 3
    for i in range(3):
        assert np.array_equal(out2[i], cv2.pyrDown(cv2.divide(imgs[i], 5)))
    # (c) the list ends: fewer frames than asked for, no crash; a wrong camera size is an error
    out4 = np.zeros((5, 480, 640), np.uint16); ts4 = np.zeros(5)
    assert host.hkfhost_read_dataset((str(d) + "/").encode(), 640, 480, 5, out4.ctypes.data, ts4.ctypes.data) == 3
    assert host.hkfhost_read_dataset((str(d) + "/").encode(), 100, 100, 1, out4.ctypes.data, ts4.ctypes.data) == 0
    assert host.hkfhost_read_dataset(b"/nonexistent/", 640, 480, 1, out4.ctypes.data, ts4.ctypes.data) == -1
    # (d) the PNG reader on its own: 8-bit RGB with every filter type OpenCV may pick
    rgb = (np.random.default_rng(1).integers(0, 255, (37, 53, 3))).astype(np.uint8)
    cv2.imwrite(str(tmp_path / "c.png"), rgb)
    host.hkfhost_read_png.argtypes = [C.c_char_p] + [C.c_void_p] * 5 + [C.c_size_t]
    w, h, ch, dp = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
    buf = np.zeros(37 * 53 * 3, np.uint16)
    assert host.hkfhost_read_png(str(tmp_path / "c.png").encode(), C.byref(w), C.byref(h), C.byref(ch), C.byref(dp), buf.ctypes.data, buf.size) == 0
    assert (w.value, h.value, ch.value, dp.value) == (53, 37, 3, 8)
    assert np.array_equal(buf.reshape(37, 53, 3), rgb[..., ::-1])               # OpenCV writes BGR as RGB


def _random_sums(rng, n_rows=400):
    """27 packed sums of a random, well-conditioned point-to-plane system whose solution is a small motion"""
    J = rng.normal(size=(n_rows, 6)).astype(np.float32)
    x_true = (rng.normal(size=6) * np.array([0.01, 0.01, 0.01, 0.005, 0.005, 0.005])).astype(np.float32)
    r = (J @ x_true + rng.normal(size=n_rows).astype(np.float32) * 1e-4).astype(np.float32)
    rows = np.concatenate([J, r[:, None]], axis=1).astype(np.float64)
    s = [np.sum(rows[:, i] * rows[:, j]) for i in range(6) for j in range(i, 7)]
    return np.array(s, np.float32), x_true


def test_host_solves_equal_oracle_and_numpy(host):
    """csrc/hkf_solver.h (what the device and the stepwise host classes run) against the oracle's independent
    restatement (bit for bit: same algorithms, same libm) and against float64 NumPy (the Eigen boundary is unpinned in
    the reference, so this is the external sanity check SURVEY.md 8c asks for)"""
    import _oracle as O
    rng = np.random.default_rng(7)
    host.hkfhost_det6.restype = C.c_float
    host.hkfhost_icp_update.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    host.hkfhost_sdf_update.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    cur0 = synth.trajectory(50)[7].astype(np.float32)
    for k in range(25):
        s27, x_true = _random_sums(rng)
        # LLT solve: oracle bit-exact, NumPy within float32 conditioning
        x = np.zeros(6, np.float32)
        host.hkfhost_llt_solve6(s27.ctypes.data, x.ctypes.data)
        ok, xo = O.solve_icp(s27)
        assert ok == 0 and bits_equal(x, xo)
        A = np.zeros((6, 6)); b = np.zeros(6); it = iter(s27.astype(np.float64))
        for i in range(6):
            for j in range(i, 7):
                v = next(it)
                if j == 6:
                    b[i] = v
                else:
                    A[i, j] = A[j, i] = v
        xn = np.linalg.solve(A, b)
        assert np.abs(x - xn).max() <= 2e-5 * max(1.0, np.abs(xn).max())
        det = host.hkfhost_det6(s27.ctypes.data)
        assert abs(det - np.linalg.det(A)) <= 1e-3 * abs(np.linalg.det(A))
        # whole ICP update: T(x) * cur with Rx*Ry*Rz and the shake gate
        cur = cur0.copy()
        rc = host.hkfhost_icp_update(s27.ctypes.data, 0.4, 0.4, cur.ctypes.data)
        T = np.zeros((4, 4), np.float32)
        assert O.cpu().orc_vector6_to_transform(O._fp(xo), C.c_float(0.4), C.c_float(0.4), O._fp(T)) == 1 and rc == 0
        assert bits_equal(cur, O.mat44_mul(T, cur0))
        from scipy.spatial.transform import Rotation
        Rn = (Rotation.from_euler("x", xn[0]) * Rotation.from_euler("y", xn[1]) * Rotation.from_euler("z", xn[2])).as_matrix()
        assert np.abs(T[:3, :3] - Rn).max() < 5e-6 and np.abs(T[:3, 3] - xn[3:]).max() < 2e-5
        # SDF update: se(3) exponential map and the reference's odd composition R <- R_d^T R, t <- t - R_d^T t_d
        cur = cur0.copy()
        conv = C.c_int()
        assert host.hkfhost_sdf_update(s27.ctypes.data, 0.4, 0.4, cur.ctypes.data, C.byref(conv)) == 0
        nxt = np.zeros((4, 4), np.float32)
        O.cpu().orc_sdf_compose(O._fp(np.ascontiguousarray(cur0).reshape(16)), O._fp(xo), O._fp(nxt))
        if not conv.value:
            assert bits_equal(cur, nxt)
            from scipy.linalg import expm
            w, v = xn[:3], xn[3:]
            xi = np.zeros((4, 4)); xi[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]; xi[:3, 3] = v
            E = expm(xi)
            want = np.eye(4); want[:3, :3] = E[:3, :3].T @ cur0[:3, :3]; want[:3, 3] = cur0[:3, 3] - E[:3, :3].T @ E[:3, 3]
            assert np.abs(cur - want).max() < 5e-6
    # gates: a singular system and a "shaking" step are reported, the transform is left alone
    cur = cur0.copy()
    assert host.hkfhost_icp_update(np.zeros(27, np.float32).ctypes.data, 0.4, 0.4, cur.ctypes.data) == 1 and bits_equal(cur, cur0)
    s27, _ = _random_sums(rng)
    assert host.hkfhost_icp_update(s27.ctypes.data, 1e-6, 1e-6, cur.ctypes.data) == 2 and bits_equal(cur, cur0)


def test_init_without_gpu_reports_no_fallback(host, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ini = tmp_path / "config.ini"
    write_ini(str(ini))
    assert host.hkfhost_init(str(ini).encode(), 0) == 1
    assert b"no CPU fallback" in host.hkfhost_last_error()


# ------------------------------------------------------------------------------------------------ GPU
def _run_host(host, ini, frames, ts=None):
    assert host.hkfhost_init(str(ini).encode(), 0) == 0, host.hkfhost_last_error()
    poses, sts = [], []
    for i, mm in enumerate(frames):
        p = np.zeros((4, 4), np.float32)
        st = C.c_int()
        mm = np.ascontiguousarray(mm)
        rc = host.hkfhost_process(mm.ctypes.data, mm.shape[1], mm.shape[0], i, float(i / 30.0 if ts is None else ts[i]), p.ctypes.data, C.byref(st))
        assert rc == 0, host.hkfhost_last_error()
        poses.append(p); sts.append(st.value)
    return np.stack(poses), np.array(sts)


@pytest.mark.gpu
@pytest.mark.parametrize("sdf", [0, 1])
def test_hybkinectfu_fused_equals_c_abi_path_and_reference(host, pipeline, tmp_path, sdf):
    from test_gpu_pipeline import run_fused, pose_err, ROT_TOL, TRANS_TOL
    from conftest import small_config
    ini = tmp_path / "config.ini"
    write_ini(str(ini), sdf=sdf, mesh=str(tmp_path / "mesh.obj"))
    poses, sts = _run_host(host, ini, pipeline["frames"])
    want, wsts, c = run_fused(small_config(128), pipeline["frames"], sdf)
    assert (sts == wsts).all() and bits_equal(poses, want)       # the C++ classes add nothing to the arithmetic
    gold = pipeline["sdf_poses" if sdf else "icp_poses"]
    rot, trans = pose_err(poses, gold)
    assert rot <= ROT_TOL and trans <= TRANS_TOL, (rot, trans)
    # same volume as the python-driven context, and the mesh export works on it
    h = capi.Context.__new__(capi.Context)
    vol_host = np.zeros((128, 128, 128, 2), np.float32)
    lib = capi.load()
    assert lib.hkf_download(C.c_void_p(host.hkfhost_ctx()), capi.BUF_VOLUME, 0, vol_host.ctypes.data, vol_host.nbytes) == vol_host.nbytes
    assert bits_equal(vol_host, c.download(capi.BUF_VOLUME))
    ntri, nv, nf = C.c_uint(), C.c_uint(), C.c_uint()
    assert host.hkfhost_mesh(str(tmp_path / "mesh.obj").encode(), C.byref(ntri), C.byref(nv), C.byref(nf)) == 0
    assert ntri.value == c.marching_cubes() and 0 < nv.value < 3 * ntri.value and nf.value <= ntri.value
    c.close()
    host.hkfhost_release()


@pytest.mark.gpu
def test_hybkinectfu_stepwise_reference_control_flow(host, pipeline, tmp_path):
    """per-iteration wrapper calls + host solves (the reference's loop) give the fused path's poses"""
    from test_gpu_pipeline import pose_err, ROT_TOL, TRANS_TOL
    ini = tmp_path / "config.ini"
    write_ini(str(ini), stepwise=1)
    poses, sts = _run_host(host, ini, pipeline["frames"])
    host.hkfhost_release()
    assert (sts == pipeline["icp_status"]).all()
    rot, trans = pose_err(poses, pipeline["icp_poses"])
    assert rot <= ROT_TOL and trans <= TRANS_TOL, (rot, trans)
    write_ini(str(ini), stepwise=1, sdf=1)
    poses, sts = _run_host(host, ini, pipeline["frames"])
    host.hkfhost_release()
    rot, trans = pose_err(poses, pipeline["sdf_poses"])
    assert (sts == pipeline["sdf_status"]).all() and rot <= ROT_TOL and trans <= TRANS_TOL, (rot, trans)


@pytest.mark.gpu
def test_hybkinectfu_poses_from_file(host, pipeline, tmp_path):
    """C1 shape: CameraPoseFinderFromFile replaying the ground-truth trajectory; the volume must equal integration
    with the very poses the finder produced"""
    n = len(pipeline["frames"])
    traj = tmp_path / "gt.txt"
    synth.write_tum_trajectory(str(traj), pipeline["gt_poses"].astype(np.float64))
    ini = tmp_path / "config.ini"
    write_ini(str(ini), fromfile=1, traj_in=str(traj), record=1, traj_out=str(tmp_path / "rec.txt"))
    # the file holds 6-decimal timestamps; a frame stamped later than the last record is "camera lost" in the reference too
    ts = np.round(np.arange(n) / 30.0, 6)
    poses, sts = _run_host(host, ini, pipeline["frames"], ts)
    assert (sts == 0).all()
    assert np.abs(poses - pipeline["gt_poses"]).max() < 2e-6
    vol_host = np.zeros((128, 128, 128, 2), np.float32)
    lib = capi.load()
    assert lib.hkf_download(C.c_void_p(host.hkfhost_ctx()), capi.BUF_VOLUME, 0, vol_host.ctypes.data, vol_host.nbytes) == vol_host.nbytes
    host.hkfhost_release()
    from conftest import small_config
    from test_gpu_pipeline import run_fused
    _, _, c = run_fused(small_config(128), pipeline["frames"], 2, poses)
    assert bits_equal(vol_host, c.download(capi.BUF_VOLUME))
    c.close()
    rec = open(tmp_path / "rec.txt").read().splitlines()
    assert len(rec) == 3 + n


@pytest.mark.gpu
def test_main_controller_loop_with_synthetic_source(host, tmp_path):
    ini = tmp_path / "config.ini"
    write_ini(str(ini), frames=40, mesh=str(tmp_path / "m.obj"))
    assert host.hkfhost_init(str(ini).encode(), 0) == 0, host.hkfhost_last_error()
    assert host.hkfhost_run(6, 1) == 6
    host.hkfhost_release()
    assert os.path.getsize(tmp_path / "m.obj") > 1000


@pytest.mark.gpu
def test_main_controller_on_a_tum_format_dataset(host, pipeline, tmp_path):
    """the caller-side data formats end to end: depth PNGs + depth.txt (TUM RGB-D layout) read by
    DataSourceProducerRGBDDataset, poses from groundtruth.txt through CameraPoseFinderFromFile, MainController loop"""
    cv2 = pytest.importorskip("cv2")
    d = tmp_path / "ds"
    (d / "depth").mkdir(parents=True)
    frames = pipeline["frames"]
    n = len(frames)
    lines = ["# depth maps", "# file: synthetic", "# timestamp filename"]
    for i, mm in enumerate(frames):
        assert cv2.imwrite(str(d / "depth" / f"{i:04d}.png"), (mm.astype(np.uint32) * 5).astype(np.uint16))
        lines.append(f"{i / 30:.6f} depth/{i:04d}.png")
    (d / "depth.txt").write_text("\n".join(lines) + "\n")
    synth.write_tum_trajectory(str(d / "groundtruth.txt"), pipeline["gt_poses"].astype(np.float64))
    ini = tmp_path / "config.ini"
    write_ini(str(ini), synthetic=0, dataset=1, source=str(d) + "/", fromfile=1, traj_in=str(d / "groundtruth.txt"))
    assert host.hkfhost_init(str(ini).encode(), 0) == 0, host.hkfhost_last_error()
    assert host.hkfhost_run(n + 3, 0) == n                       # the loop ends when the dataset does
    vol = np.zeros((128, 128, 128, 2), np.float32)
    lib = capi.load()
    assert lib.hkf_download(C.c_void_p(host.hkfhost_ctx()), capi.BUF_VOLUME, 0, vol.ctypes.data, vol.nbytes) == vol.nbytes
    host.hkfhost_release()
    # frames round-trip exactly (mm*5 -> /5), poses are the file's: same volume as the in-memory path with those poses
    from conftest import small_config
    from test_gpu_pipeline import run_fused
    ts = np.round(np.arange(n) / 30.0, 6)
    out = np.zeros((n, 4, 4), np.float32)
    assert host.hkfhost_replay_trajectory(str(d / "groundtruth.txt").encode(), synth.pose0().astype(np.float32).ctypes.data, ts.ctypes.data, n, out.ctypes.data) == 0
    _, _, c = run_fused(small_config(128), frames, 2, out)
    assert bits_equal(vol, c.download(capi.BUF_VOLUME))
    c.close()


@pytest.mark.gpu
def test_hybkinectfu_colour_mode_equals_c_abi_path(host, pipeline, tmp_path):
    """use_color=1 (the reference's default, config.ini:2) through HybKinectfu / MeshGeneratorMarchingcube: the volume,
    its colours and raycast_rgb equal the C-ABI path fed the same frames; the PLY carries vertex colours"""
    import dataclasses
    from conftest import small_config
    # 640x480: the reference's colour pixel uses hard-coded 525 / 320 / 240 intrinsics (integrateVolume.cu:56-57), so a
    # scaled-down camera would colour - and, with use_color, integrate - only a strip of the volume
    cam = synth.Camera()
    gt = synth.trajectory(200)[::4][:5]
    frames = np.stack([synth.render_depth_mm(p, cam) for p in gt])
    bgr = np.stack([synth.render_color(p, cam) for p in gt])
    ini = tmp_path / "config.ini"
    write_ini(str(ini), cols=640, rows=480, color=1, angle=1, mesh=str(tmp_path / "mesh.ply"))
    assert host.hkfhost_init(str(ini).encode(), 0) == 0, host.hkfhost_last_error()
    host.hkfhost_process_color.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_void_p, C.c_void_p]
    poses = []
    for i, mm in enumerate(frames):
        p = np.zeros((4, 4), np.float32); st = C.c_int()
        mm = np.ascontiguousarray(mm); im = np.ascontiguousarray(bgr[i])
        assert host.hkfhost_process_color(mm.ctypes.data, im.ctypes.data, 640, 480, i, i / 30.0, p.ctypes.data, C.byref(st)) == 0, host.hkfhost_last_error()
        assert st.value == 0
        poses.append(p)
    lib = capi.load()
    h = C.c_void_p(host.hkfhost_ctx())
    vol = np.zeros((128, 128, 128, 2), np.float32); col = np.zeros((128, 128, 128, 4), np.uint8); rgb = np.zeros((480, 640, 3), np.uint8)
    assert lib.hkf_download(h, capi.BUF_VOLUME, 0, vol.ctypes.data, vol.nbytes) == vol.nbytes
    assert lib.hkf_download(h, capi.BUF_VOLUME_COLOR, 0, col.ctypes.data, col.nbytes) == col.nbytes
    assert lib.hkf_download(h, capi.BUF_RAYCAST_RGB, 0, rgb.ctypes.data, rgb.nbytes) == rgb.nbytes
    ntri, nv, nf = C.c_uint(), C.c_uint(), C.c_uint()
    assert host.hkfhost_mesh(str(tmp_path / "mesh.ply").encode(), C.byref(ntri), C.byref(nv), C.byref(nf)) == 0
    host.hkfhost_release()
    cfg = dataclasses.replace(capi.Config(resolution=128, max_triangles=400000), use_color=1, color_angle_weight=1)
    with capi.Context(cfg) as c:
        c.set_pose(synth.pose0().astype(np.float32))
        for i, mm in enumerate(frames):
            c.preprocess(mm); c.upload_color(bgr[i])
            if i:
                c.icp_track()
            c.integrate_tracked(); c.raycast_tracked()
            p, st = c.frame_result()
            assert st == 0 and bits_equal(p, poses[i])
        assert bits_equal(c.download(capi.BUF_VOLUME), vol)
        assert (c.download(capi.BUF_VOLUME_COLOR) == col).all() and col[..., :3].any()
        assert (c.download(capi.BUF_RAYCAST_RGB) == rgb).all() and rgb.any()
        assert c.marching_cubes() == ntri.value
    lines = open(tmp_path / "mesh.ply").read().splitlines()
    assert "property uchar red" in lines and f"element vertex {nv.value}" in lines
    body = lines[lines.index("end_header") + 1:]
    assert len(body[0].split()) == 9 and any(int(x) > 0 for x in body[0].split()[6:])
