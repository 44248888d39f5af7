"""Mesh post-processing against the reference's OWN code.

oracle/_ref/libmesh_ref.so is the reference's ml::MeshData (src/utils/mesh/meshData.{h,cpp}) compiled from where it lies
and driven as MeshGeneratorMarchingcube::saveMesh drives it (src/MeshGeneratorMarchingcube.cpp:61-97): approximate 1e-4
vertex merge, degenerate + duplicate face removal, vertex normals.  The host class of this repo
(hybkinectfu_b200/host/MeshGeneratorMarchingcube.cpp) restates that and must produce the SAME indexed mesh - vertex
list, face list, normals, colours, element for element - when both see the soup in the same (canonical) order.
"""
import ctypes as C
import os

import numpy as np
import pytest

import _oracle as O
from conftest import GOLDEN, ROOT, small_config

HOST_LIB = os.path.join(ROOT, "hybkinectfu_b200", "libhkf_host.so")
MESH_REF = os.path.join(ROOT, "oracle", "_ref", "libmesh_ref.so")

pytestmark = pytest.mark.skipif(not os.path.exists(MESH_REF), reason="oracle/_ref/libmesh_ref.so not built (needs /root/reference at build time)")


def canonical(tris):
    p = np.ascontiguousarray(tris[:, :, :3]).reshape(len(tris), 9)
    return np.ascontiguousarray(tris[np.lexsort(p.T[::-1])])


def run_ref(tris, use_color):
    lib = C.CDLL(MESH_REF)
    n = len(tris)
    nv, nf = C.c_uint(), C.c_uint()
    v = np.zeros((3 * n, 3), np.float32); nr = np.zeros((3 * n, 3), np.float32); col = np.zeros((3 * n, 4), np.float32)
    f = np.zeros((3 * n, 3), np.uint32)
    t = np.ascontiguousarray(tris, np.float32)
    assert lib.ref_mesh_process(t.ctypes.data_as(C.c_void_p), C.c_uint(n), int(use_color), C.byref(nv), C.byref(nf), v.ctypes.data_as(C.c_void_p),
                                nr.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p)) == 0
    return v[:nv.value], nr[:nv.value], col[:nv.value], f[:nf.value]


def run_host(tris, use_color):
    lib = C.CDLL(HOST_LIB)
    n = len(tris)
    nv, nf = C.c_uint(), C.c_uint()
    v = np.zeros((3 * n, 3), np.float32); nr = np.zeros((3 * n, 3), np.float32); col = np.zeros((3 * n, 4), np.float32)
    f = np.zeros((n, 3), np.uint32)
    t = np.ascontiguousarray(tris, np.float32)
    assert lib.hkfhost_weld_arrays(t.ctypes.data_as(C.c_void_p), C.c_uint(n), int(use_color), C.byref(nv), C.byref(nf), v.ctypes.data_as(C.c_void_p),
                                   nr.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p)) == 0
    return v[:nv.value], nr[:nv.value], col[:nv.value], f[:nf.value]


def same_mesh(a, b, use_color):
    (va, na, ca, fa), (vb, nb, cb, fb) = a, b
    assert va.shape == vb.shape and fa.shape == fb.shape, (va.shape, vb.shape, fa.shape, fb.shape)
    assert np.array_equal(va.view(np.uint32), vb.view(np.uint32)), "welded vertex lists differ"
    assert np.array_equal(fa, fb), "face lists differ"
    assert np.array_equal(na.view(np.uint32), nb.view(np.uint32)) or np.array_equal(na, nb), "vertex normals differ"
    if use_color:
        assert np.array_equal(ca, cb), "vertex colours differ"


def test_weld_equals_reference_meshdata_on_oracle_soup():
    """a 128^3 volume integrated and extracted by the CPU oracle (a few thousand triangles), shuffled"""
    cfg = small_config(128)
    from hybkinectfu_b200 import synth
    cam = synth.Camera().scaled(192, 144)
    poses = synth.trajectory(200)[::4][:3]
    vol = np.zeros((128, 128, 128, 2), np.float32)
    col = np.zeros((128, 128, 128, 4), np.uint8)
    for p in poses:
        tr = O.trunc(O.decode_mm(synth.render_depth_mm(p, cam)), cfg.trunc_min, cfg.trunc_max)
        O.integrate_color(vol, col, cfg, tr, np.zeros((cfg.rows, cfg.cols, 4), np.float32), synth.render_color(p, cam), p.astype(np.float32), 0)
    n, tris = O.marching_cubes_color(vol, col, cfg)
    assert n > 2000
    rng = np.random.default_rng(3)
    shuffled = tris[rng.permutation(n)]
    for use_color in (0, 1):
        ref = run_ref(canonical(tris), use_color)
        got = run_host(shuffled, use_color)                 # the host class sorts by itself: order independent
        same_mesh(got, ref, use_color)
        assert len(ref[0]) < 0.6 * 3 * n and len(ref[3]) <= n
    if os.path.exists(os.path.join(GOLDEN, "color_small.npz")):
        g = np.load(os.path.join(GOLDEN, "color_small.npz"))
        t = g["mc_tris_sorted_1"].reshape(-1, 3, 6)
        same_mesh(run_host(t, 1), run_ref(canonical(t), 1), 1)


def test_weld_merges_near_duplicates_like_the_reference():
    """vertices one float ulp apart (the two interpolation directions of a shared edge) and slivers shorter than the
    1e-4 m cell: merged / collapsed exactly where ml::MeshData does"""
    base = np.array([[0.30001, 0.5, 0.7], [0.35, 0.5, 0.7], [0.30001, 0.56, 0.7]], np.float32)
    t = np.zeros((4, 3, 6), np.float32)
    t[0, :, :3] = base
    t[1, :, :3] = base[[1, 0, 2]] + np.array([0, 0, 0], np.float32)
    t[1, 1, 0] = np.nextafter(t[1, 1, 0], np.float32(1))          # 1 ulp off: same vertex after the merge
    t[1, 2, :3] = [0.35, 0.56, 0.7]
    t[2, :, :3] = [[0.9, 0.9, 0.9], [0.90004, 0.9, 0.9], [0.9, 0.95, 0.9]]     # 4e-5 m edge: collapses -> face removed
    t[3, :, :3] = [[0.1, 0.1, 0.1], [0.1004, 0.1, 0.1], [0.1, 0.15, 0.1]]      # 4e-4 m edge: stays
    same_mesh(run_host(t, 0), run_ref(canonical(t), 0), 0)
    v, _, _, f = run_host(t, 0)
    assert len(f) == 3 and len(v) == 4 + 2 + 3


@pytest.mark.gpu
def test_weld_equals_reference_meshdata_512_soup():
    """the C3-sized soup: 640x480 frames into 512^3, marching cubes on the GPU, both post-processings on the host"""
    import time
    from hybkinectfu_b200 import capi, synth
    cfg = capi.Config(resolution=512, max_triangles=3000000, use_color=1)
    cam = synth.Camera()
    poses = synth.trajectory(500)[:24:4]
    with capi.Context(cfg) as c:
        for p in poses:
            c.preprocess(synth.render_depth_mm(p, cam)); c.upload_color(synth.render_color(p, cam)); c.integrate(p.astype(np.float32))
        n = c.marching_cubes()
        tris = c.download(capi.BUF_TRIANGLES)
    assert n > 300000 and len(tris) == n
    t0 = time.time(); got = run_host(tris, 1); t1 = time.time(); ref = run_ref(canonical(tris), 1); t2 = time.time()
    print(f"{n} triangles -> {len(ref[0])} vertices, {len(ref[3])} faces; host class {t1 - t0:.2f} s, reference ml::MeshData {t2 - t1:.2f} s")
    same_mesh(got, ref, 1)


def test_weld_tables_grow_on_a_soup_that_does_not_weld():
    """20 000 triangles scattered over 2 m: hardly any vertex is shared, so the weld registers three vertices per triangle
    (its tables are sized for a marching-cubes soup - half a vertex per triangle - and must double twice), plus the
    whole soup a second time (every face a duplicate) and a block of triangles on a 5e-5 m lattice, where vertices of
    neighbouring cells merge by the first-found rule"""
    rng = np.random.default_rng(11)
    n = 20000
    base = rng.uniform(0.0, 2.0, size=(n, 1, 3)).astype(np.float32)
    tris = np.zeros((n, 3, 6), np.float32)
    tris[:, :, :3] = base + rng.uniform(-0.01, 0.01, size=(n, 3, 3)).astype(np.float32)
    tris[:, :, 3:] = rng.integers(0, 256, size=(n, 3, 3)).astype(np.float32)
    lattice = np.zeros((3000, 3, 6), np.float32)
    lattice[:, :, :3] = (rng.integers(0, 60, size=(3000, 3, 3)) * np.float32(5e-5) + np.float32(0.5)).astype(np.float32)
    lattice[:, :, 3:] = 200.0
    soup = np.concatenate([tris, tris, lattice])
    for use_color in (0, 1):
        ref = run_ref(canonical(soup), use_color)
        got = run_host(soup[rng.permutation(len(soup))], use_color)
        assert len(ref[0]) > 2.9 * n and len(ref[3]) < len(soup) - n + 1
        same_mesh(ref, got, use_color)
