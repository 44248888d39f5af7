"""Host emulation of the ICP kernel source (tests/emu): k_icp against the oracle's 27 sums at every pyramid level, and the
FUSED form of the same kernel - sums, the last block's 6x6 Cholesky solve and LU determinant, the gates and the pose
update spread over two warps, one launch per iteration - against the reference loop `oracle sums -> host solve`
(hkf_solver.h on the CPU), bit for bit, failure paths included."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
from conftest import ROOT, small_config

EMU_DIR = os.path.join(ROOT, "tests", "emu")
HOST_LIB = os.path.join(ROOT, "hybkinectfu_b200", "libhkf_host.so")


def _f(x):
    return C.c_float(float(x))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def scene():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    r = subprocess.run(["make", "-C", EMU_DIR, "libicp_emu.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    if not C.CDLL(os.path.join(EMU_DIR, "libicp_emu.so")).emu_can_spawn(512 + 64):
        pytest.skip("this container does not allow the 512 threads of an emulated block")
    from hybkinectfu_b200 import synth
    cfg = small_config(64)
    cam = synth.Camera().scaled(cfg.cols, cfg.rows)
    poses = synth.trajectory(200)[::5][:4]
    R = cfg.resolution
    vol = np.zeros((R, R, R, 2), np.float32)
    for p in poses[:3]:
        O.integrate(vol, cfg, O.trunc(O.decode_mm(synth.render_depth_mm(p, cam)), cfg.trunc_min, cfg.trunc_max), p.astype(np.float32))
    prev = poses[2].astype(np.float32)
    mv0, mn0 = O.raycast(vol, cfg, prev)
    fl = O.bilateral(O.trunc(O.decode_mm(synth.render_depth_mm(poses[3], cam)), cfg.trunc_min, cfg.trunc_max), cfg.sigma_pixel, cfg.sigma_depth)
    nv0 = O.vertices(fl, O.cam_of(cfg))
    nn0 = O.normals(nv0)

    def pyr(a, is_normal):
        out = [a]
        for _ in range(2):
            out.append(O.down(out[-1], is_normal))
        return out

    return {"cfg": cfg, "prev": np.ascontiguousarray(prev).reshape(16), "last_inv": np.ascontiguousarray(O.mat44_inverse(prev), np.float32).reshape(16),
            "NV": pyr(nv0, False), "NN": pyr(nn0, True), "MV": pyr(mv0, False), "MN": pyr(mn0, True),
            "emu": os.path.join(EMU_DIR, "libicp_emu.so")}


@pytest.mark.parametrize("level", [2, 1, 0])
def test_default_kernel_sums_equal_oracle(scene, level):
    cfg = scene["cfg"]
    lib = C.CDLL(scene["emu"])
    oc = O.cam_of(cfg, level)
    maps = [scene[k][level] for k in ("NV", "NN", "MV", "MN")]
    want = O.icp_sums(*maps, oc, scene["prev"].reshape(4, 4), scene["last_inv"].reshape(4, 4), cfg.icp_dist_thres, cfg.icp_norm_sin_thres)
    got = np.zeros(27, np.float32)
    lib.emu_icp_sums(*[_p(m) for m in maps], int(oc.cols), int(oc.rows), _f(oc.cx), _f(oc.cy), _f(oc.fx), _f(oc.fy),
                     _p(scene["prev"]), _p(scene["last_inv"]), _f(cfg.icp_dist_thres), _f(cfg.icp_norm_sin_thres), _p(got))
    assert np.abs(want).max() > 1 and np.array_equal(want.view(np.uint32), got.view(np.uint32))


def _reference_loop(scene, level, iters, angle_shake, maps):
    cfg = scene["cfg"]
    host = C.CDLL(HOST_LIB)
    host.hkfhost_icp_update.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    oc = O.cam_of(cfg, level)
    cur = scene["prev"].copy()
    sums, status = None, 0
    for _ in range(iters):
        sums = O.icp_sums(*maps, oc, cur.reshape(4, 4), scene["last_inv"].reshape(4, 4), cfg.icp_dist_thres, cfg.icp_norm_sin_thres)
        rc = host.hkfhost_icp_update(_p(sums), angle_shake, cfg.icp_dist_shake, _p(cur))
        if rc != 0:
            status = rc          # HKF_SOLVE_SINGULAR = 1 / HKF_SOLVE_SHAKE = 2 == HKF_TRACK_SINGULAR / HKF_TRACK_SHAKE
            break
    return cur, sums, status


def _fused(scene, level, iters, angle_shake, maps):
    cfg = scene["cfg"]
    oc = O.cam_of(cfg, level)
    lib = C.CDLL(scene["emu"])
    st = np.zeros(lib.emu_state_bytes() + 64, np.uint8)
    lib.emu_state_init(_p(st), _p(scene["prev"]), _p(scene["last_inv"]))
    lib.emu_icp_fused(*[_p(m) for m in maps], int(oc.cols), int(oc.rows), _f(oc.cx), _f(oc.cy), _f(oc.fx), _f(oc.fy),
                      _f(cfg.icp_dist_thres), _f(cfg.icp_norm_sin_thres), _f(angle_shake), _f(cfg.icp_dist_shake), iters, _p(st))
    cur, sums, status = np.zeros(16, np.float32), np.zeros(27, np.float32), C.c_int(0)
    lib.emu_state_read(_p(st), _p(cur), _p(sums), C.byref(status))
    return cur, sums, status.value


@pytest.mark.parametrize("level,iters", [(2, 10), (1, 5), (0, 2)])
def test_fused_kernel_equals_the_iteration_loop(scene, level, iters):
    """the last block's solve / gates / commit spread over its warps give the bits of the host-side hkf_solver.h loop"""
    cfg = scene["cfg"]
    maps = [scene[k][level] for k in ("NV", "NN", "MV", "MN")]
    c0, s0, st0 = _reference_loop(scene, level, iters, cfg.icp_angle_shake, maps)
    c1, s1, st1 = _fused(scene, level, iters, cfg.icp_angle_shake, maps)
    assert st0 == 0 and st1 == 0
    assert not np.array_equal(c0, scene["prev"])                       # the pose did move
    assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32)) and np.array_equal(s0.view(np.uint32), s1.view(np.uint32))


def test_fused_kernel_failure_paths(scene):
    cfg = scene["cfg"]
    level = 2
    maps = [scene[k][level] for k in ("NV", "NN", "MV", "MN")]
    # shake gate: an absurdly small angle limit trips on the first iteration; the working transform stays untouched
    c0, _, st0 = _reference_loop(scene, level, 10, 1e-9, maps)
    c1, _, st1 = _fused(scene, level, 10, 1e-9, maps)
    assert st0 == 2 and st1 == 2 and np.array_equal(c0.view(np.uint32), c1.view(np.uint32)) and np.array_equal(c1, scene["prev"])
    # no correspondences at all: singular normal equations
    empty = [maps[0], maps[1], np.zeros_like(maps[2]), np.zeros_like(maps[3])]
    c0, _, st0 = _reference_loop(scene, level, 10, cfg.icp_angle_shake, empty)
    c1, _, st1 = _fused(scene, level, 10, cfg.icp_angle_shake, empty)
    assert st0 == 1 and st1 == 1 and np.array_equal(c0.view(np.uint32), c1.view(np.uint32))
