"""The ray-march KERNEL SOURCE (hybkinectfu_b200/csrc/raycast.cu), compiled for the host by tests/emu and run one simulated
thread at a time, against the CPU oracle - bit for bit.  The kernel has no warp collectives and no shared memory, and
every intrinsic it uses is a single IEEE operation on both sides, so the host build executes the same arithmetic: the
generic and the FAST (power-of-two resolution) index paths, from views the GPU tests do not use."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
from conftest import ROOT, bits_equal, small_config

EMU_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    r = subprocess.run(["make", "-C", EMU_DIR, "libraycast_emu.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return C.CDLL(os.path.join(EMU_DIR, "libraycast_emu.so"))


def _run(lib, vol, cfg, pose, fast):
    v = np.zeros((cfg.rows, cfg.cols, 4), np.float32)
    n = np.zeros_like(v)
    p = np.ascontiguousarray(pose, np.float32).reshape(16)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = lib.emu_raycast(ptr(vol), cfg.resolution, C.c_float(cfg.volume_size), cfg.cols, cfg.rows, C.c_float(cfg.cx), C.c_float(cfg.cy),
                         C.c_float(cfg.fx), C.c_float(cfg.fy), ptr(p), C.c_float(cfg.ray_increment), C.c_float(cfg.trunc_min),
                         C.c_float(cfg.trunc_max), int(fast), ptr(v), ptr(n))
    assert rc == 0
    return v, n


def _scene(res, nframes=6):
    from hybkinectfu_b200 import synth
    cfg = small_config(res)
    cam = synth.Camera().scaled(cfg.cols, cfg.rows)
    poses = synth.trajectory(200)[::5][:nframes]
    vol = np.zeros((res, res, res, 2), np.float32)
    for p in poses:
        d = O.trunc(O.decode_mm(synth.render_depth_mm(p, cam)), cfg.trunc_min, cfg.trunc_max)
        O.integrate(vol, cfg, d, p.astype(np.float32))
    a, b = 0.4, -0.3
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    T = np.eye(4); T[:3, :3] = Ry @ Rx; T[:3, 3] = [0.2, 1.4, -0.5]        # from outside the volume, rotated about two axes
    return cfg, vol, [poses[0], poses[-1], T]


@pytest.mark.parametrize("res", [64, 128])
def test_kernel_source_equals_oracle_in_every_variant(emu, res):
    cfg, vol, views = _scene(res)
    for pose in views:
        pose = np.asarray(pose, np.float32)
        v0, n0 = O.raycast(vol, cfg, pose)
        assert (n0[..., :3] != 0).any(-1).sum() > 500
        for fast in (0, 1):
            v1, n1 = _run(emu, vol, cfg, pose, fast)
            assert bits_equal(v0, v1) and bits_equal(n0, n1), f"variant FAST={fast}"
