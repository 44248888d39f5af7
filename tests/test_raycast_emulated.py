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


def _run_slab(lib, vol, cfg, pose, fast, rng, halo, peer, rank):
    """slab `rng` = (own_z0, own_z1) of the whole volume `vol`; peer=False: the slab array holds owned + halo planes;
    peer=True: it holds the owned planes only and the halo planes are handed over as the neighbours' arrays"""
    R = cfg.resolution
    z0, z1 = rng
    k0, k1 = max(z0 - halo, 0), min(z1 + halo, R)
    if peer:
        own = np.ascontiguousarray(vol[z0:z1]); lo = np.ascontiguousarray(vol[k0:z0]); hi = np.ascontiguousarray(vol[z1:k1])
        s0, s1 = z0, z1
    else:
        own = np.ascontiguousarray(vol[k0:k1]); lo = hi = None
        s0, s1 = k0, k1
    v = np.zeros((cfg.rows, cfg.cols, 4), np.float32)
    n = np.zeros_like(v)
    keys = np.zeros((cfg.rows, cfg.cols), np.uint64)
    p = np.ascontiguousarray(pose, np.float32).reshape(16)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None
    rc = lib.emu_raycast_slab(ptr(own), ptr(lo), ptr(hi), R, C.c_float(cfg.volume_size), s0, s1, k0 if peer else s0, k1 if peer else s1,
                              z0, z1, rank, cfg.cols, cfg.rows, C.c_float(cfg.cx), C.c_float(cfg.cy), C.c_float(cfg.fx), C.c_float(cfg.fy),
                              ptr(p), C.c_float(cfg.ray_increment), C.c_float(cfg.trunc_min), C.c_float(cfg.trunc_max), int(fast),
                              ptr(v), ptr(n), ptr(keys))
    assert rc == 0
    return v, n, keys


@pytest.mark.parametrize("peer", [False, True])
def test_slab_kernel_source_composites_to_the_whole_raycast(emu, peer):
    """N > 1 raycast on the CPU: the SLAB instantiation of the kernel source over three unequal slabs - with the halo
    planes stored in the slab, and with the halo read in the neighbours' arrays (hkf_shard_peer_halo) - composited with
    the rule of k_composite_p2p (shard.composite) equals the oracle's raycast of the whole volume bit for bit, from
    views that climb and descend through the planes"""
    from hybkinectfu_b200 import shard
    cfg, vol, views = _scene(64)
    halo = shard.halo_planes(cfg.ray_increment, cfg.volume_size, cfg.resolution)
    R = cfg.resolution
    cuts = [0, max(halo, 20), max(halo, 20) + max(halo, 22), R]
    assert all(b - a >= halo for a, b in zip(cuts, cuts[1:]))
    # from above the floor, looking down and back towards -z: the rays DESCEND through the planes (the previous sample of
    # a crossing then lies in the slab above)
    f = np.array([0.0, 1.5, -0.3]); f /= np.linalg.norm(f)
    back = np.eye(4); back[:3, 0] = [1.0, 0.0, 0.0]; back[:3, 1] = [0.0, f[2], -f[1]]; back[:3, 2] = f; back[:3, 3] = [1.0, 0.2, 1.5]
    for pose in views + [back]:
        pose = np.asarray(pose, np.float32)
        v0, n0 = O.raycast(vol, cfg, pose)
        assert (n0[..., :3] != 0).any(-1).sum() > 300
        for fast in (0, 1):
            parts = [_run_slab(emu, vol, cfg, pose, fast, (cuts[r], cuts[r + 1]), halo, peer, r) for r in range(3)]
            best, v, n = shard.composite(np.stack([p[2] for p in parts]), np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
            assert bits_equal(v, v0) and bits_equal(n, n0), f"peer={peer} FAST={fast}"


@pytest.fixture(scope="module")
def shard_emu():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    r = subprocess.run(["make", "-C", EMU_DIR, "libshard_emu.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    lib = C.CDLL(os.path.join(EMU_DIR, "libshard_emu.so"))
    if not lib.emu_shard_can_run():
        pytest.skip("cannot create 256 threads in this container")
    return lib


@pytest.mark.parametrize("peer", [1, 0])
def test_hit_pushes_and_compositing_kernel_sources_over_three_epochs(shard_emu, peer):
    """The N > 1 raycast data path without GPUs: three "ranks" (arrays of one process) run the SLAB ray-march kernel
    source, which pushes every hit into the inbox of every rank, and the compositing kernel source (k_composite_p2p)
    reads each rank's own inboxes.  Three raycasts in a row from different poses: the inboxes are double-buffered by
    epoch parity and the used slots are reset, so the third epoch lands on the slots of the first.  Every rank's
    composited maps must equal the oracle's raycast of the whole volume bit for bit."""
    import dataclasses
    from hybkinectfu_b200 import shard, synth
    lib = shard_emu
    cfg, vol, views = _scene(64)
    cam = synth.Camera().scaled(96, 72)
    cfg = dataclasses.replace(cfg, cols=96, rows=72, cx=cam.cx, cy=cam.cy, fx=cam.fx, fy=cam.fy)
    R, npix, nranks = cfg.resolution, 96 * 72, 3
    halo = shard.halo_planes(cfg.ray_increment, cfg.volume_size, R)
    cuts = np.array([0, max(halo, 20), max(halo, 20) + max(halo, 22), R], np.int32)
    maxr = lib.emu_shard_max_ranks()
    keys_in = np.full((nranks, 2, maxr, npix), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    pay_in = np.zeros((nranks, 2, maxr, npix, 8), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    volc = np.ascontiguousarray(vol, np.float32)
    for epoch, pose in enumerate(views, start=1):
        pose = np.ascontiguousarray(pose, np.float32)
        v0, n0 = O.raycast(vol, cfg, pose)
        assert (n0[..., :3] != 0).any(-1).sum() > 200
        out_v = np.full((nranks, cfg.rows, cfg.cols, 4), 7.0, np.float32)
        out_n = np.full_like(out_v, 7.0)
        rc = lib.emu_sharded_raycast(ptr(volc), R, C.c_float(cfg.volume_size), nranks, ptr(cuts), halo, peer, cfg.cols, cfg.rows,
                                     C.c_float(cfg.cx), C.c_float(cfg.cy), C.c_float(cfg.fx), C.c_float(cfg.fy), ptr(pose.reshape(16)),
                                     C.c_float(cfg.ray_increment), C.c_float(cfg.trunc_min), C.c_float(cfg.trunc_max), epoch & 1, epoch,
                                     ptr(keys_in), ptr(pay_in), ptr(out_v), ptr(out_n))
        assert rc == 0
        for r in range(nranks):
            assert bits_equal(out_v[r], v0) and bits_equal(out_n[r], n0), f"epoch {epoch}, rank {r}"
        # every slot that was used has been reset to "no hit" by the rank that owns the inbox
        assert (keys_in[:, epoch & 1] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
