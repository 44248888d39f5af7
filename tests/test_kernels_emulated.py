"""Block-level host emulation (tests/emu/emu_block.h: one std::thread per simulated CUDA thread, barriers for
__syncthreads and the warp collectives) of the preprocess and integration KERNEL SOURCES against the CPU oracle, bit for
bit.  The kernels are pinned on the GPU by tests/test_gpu_*.py; here the same sources run on this box, also at sizes the
GPU tests do not use (odd image sizes, a volume that is not a power of two)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
from conftest import ROOT, bits_equal, n_mismatch, small_config
from test_raycast_emulated import _run as _raycast_emulated

EMU_DIR = os.path.join(ROOT, "tests", "emu")


def _build(target):
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    r = subprocess.run(["make", "-C", EMU_DIR, target], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return C.CDLL(os.path.join(EMU_DIR, target))


@pytest.fixture(scope="module")
def frame():
    lib = _build("libframe_emu.so")
    if not lib.emu_can_spawn(512):
        pytest.skip("this container does not allow the threads of an emulated block")
    lib.emu_frame_create.restype = C.c_void_p
    lib.emu_frame_buffer.restype = C.c_void_p
    lib.emu_integrate.restype = C.c_uint64
    return lib


@pytest.fixture(scope="module")
def ray():
    return _build("libraycast_emu.so")


def _f(x):
    return C.c_float(float(x))


def _preprocess(lib, h, cfg, mm):
    mm = np.ascontiguousarray(mm, np.uint16)
    assert lib.emu_preprocess(C.c_void_p(h), mm.ctypes.data_as(C.c_void_p), _f(cfg.trunc_min), _f(cfg.trunc_max), _f(cfg.sigma_pixel),
                              _f(cfg.sigma_depth), _f(cfg.cx), _f(cfg.cy), _f(cfg.fx), _f(cfg.fy)) == 0


def _buffer(lib, h, which, shape):
    p = lib.emu_frame_buffer(C.c_void_p(h), which)
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=shape).copy()


def test_preprocess_kernels_equal_oracle(frame):
    from hybkinectfu_b200 import synth
    cfg = small_config(64)
    cam = synth.Camera().scaled(cfg.cols, cfg.rows)
    mm = synth.render_depth_mm(synth.trajectory(200)[7], cam, noise_sigma_mm=1.0, rng=np.random.default_rng(3))
    raw = O.decode_mm(mm)
    tr = O.trunc(raw, cfg.trunc_min, cfg.trunc_max)
    fl = O.bilateral(tr, cfg.sigma_pixel, cfg.sigma_depth)
    v0 = O.vertices(fl, O.cam_of(cfg)); n0 = O.normals(v0)
    v1, n1 = O.down(v0, False), O.down(n0, True)
    v2, n2 = O.down(v1, False), O.down(n1, True)
    h = frame.emu_frame_create(cfg.cols, cfg.rows)
    try:
        _preprocess(frame, h, cfg, mm)
        for i, (name, want) in enumerate(zip(["raw", "trunc", "filtered", "v0", "n0", "v1", "n1", "v2", "n2"], [raw, tr, fl, v0, n0, v1, n1, v2, n2])):
            assert n_mismatch(want, _buffer(frame, h, i, want.shape)) == 0, name
    finally:
        frame.emu_frame_destroy(C.c_void_p(h))


def _integrate_sequence(lib, cfg, poses):
    from hybkinectfu_b200 import synth
    R = cfg.resolution
    cam = synth.Camera().scaled(cfg.cols, cfg.rows)
    vol = np.zeros((R, R, R, 2), np.float32)
    want = np.zeros_like(vol)
    items = [0, 0]
    h = lib.emu_frame_create(cfg.cols, cfg.rows)
    try:
        for p in poses:
            mm = synth.render_depth_mm(p, cam)
            _preprocess(lib, h, cfg, mm)
            inv = np.ascontiguousarray(O.mat44_inverse(p.astype(np.float32)), np.float32).reshape(16)
            cnt = lib.emu_integrate(C.c_void_p(h), vol.ctypes.data_as(C.c_void_p), R, _f(cfg.volume_size), _f(cfg.weight_max),
                                    inv.ctypes.data_as(C.c_void_p), _f(cfg.sdf_trunc), _f(cfg.max_integrate_dist), _f(cfg.cx), _f(cfg.cy),
                                    _f(cfg.fx), _f(cfg.fy))
            items[0] += cnt & 0xffffffff; items[1] += cnt >> 32
            O.integrate(want, cfg, O.trunc(O.decode_mm(mm), cfg.trunc_min, cfg.trunc_max), p.astype(np.float32))
    finally:
        lib.emu_frame_destroy(C.c_void_p(h))
    return vol, want, items


def test_integration_kernels_equal_oracle(frame):
    from hybkinectfu_b200 import synth
    cfg = small_config(64)
    vol, want, items = _integrate_sequence(frame, cfg, synth.trajectory(200)[::7][:2])
    assert (want[..., 1] > 0).sum() > 10000 and items[1] > 0
    assert n_mismatch(vol, want) == 0


def test_odd_image_size_and_non_power_of_two_volume(frame, ray):
    """sizes the GPU tests do not use: 193 x 145 pixels (odd: the pyramid levels floor, tiles are cut at both edges) into a
    68^3 volume (not a multiple of the 32-voxel chunk nor of the bricks; the ray march takes its generic index path)"""
    import dataclasses
    from hybkinectfu_b200 import synth
    cols, rows, R = 193, 145, 68
    cam = synth.Camera().scaled(cols, rows)
    cfg = dataclasses.replace(small_config(64), cols=cols, rows=rows, cx=cam.cx, cy=cam.cy, fx=cam.fx, fy=cam.fy, resolution=R)
    traj = synth.trajectory(200)
    mm = synth.render_depth_mm(traj[7], cam, noise_sigma_mm=1.0, rng=np.random.default_rng(3))
    raw = O.decode_mm(mm)
    tr = O.trunc(raw, cfg.trunc_min, cfg.trunc_max)
    fl = O.bilateral(tr, cfg.sigma_pixel, cfg.sigma_depth)
    v0 = O.vertices(fl, O.cam_of(cfg)); n0 = O.normals(v0)
    v1, n1 = O.down(v0, False), O.down(n0, True)
    v2, n2 = O.down(v1, False), O.down(n1, True)
    h = frame.emu_frame_create(cols, rows)
    try:
        _preprocess(frame, h, cfg, mm)
        for i, (name, want) in enumerate(zip(["raw", "trunc", "filtered", "v0", "n0", "v1", "n1", "v2", "n2"], [raw, tr, fl, v0, n0, v1, n1, v2, n2])):
            assert n_mismatch(want, _buffer(frame, h, i, want.shape)) == 0, name
    finally:
        frame.emu_frame_destroy(C.c_void_p(h))
    vol, want, items = _integrate_sequence(frame, cfg, traj[::9][:2])
    assert items[0] > 0 and items[1] > 0 and n_mismatch(vol, want) == 0
    pose = traj[20].astype(np.float32)
    a, b = O.raycast(want, cfg, pose)
    c, d = _raycast_emulated(ray, want, cfg, pose, 0)
    assert (b[..., :3] != 0).any(-1).sum() > 500 and bits_equal(a, c) and bits_equal(b, d)
