#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference kernels.

Run on a B200 box (the reference's device code needs a GPU):
    python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/.  The reference holds no golden vectors of its
own, so these outputs of its own kernels (oracle/_ref/libhkf_ref.so, built by
oracle/Makefile from /root/reference/src/cuda/*.cu with nvcc defaults) are what pins the CPU oracle
and, through it and directly, the CUDA path.

Fixtures (small on purpose: 192x144 frames, 64^3 / 128^3 volumes over 2 m):
  stages_small.npz    every per-stage buffer for one frame, volume after 3 integrations, raycast maps,
                      model pyramids, ICP sums per level, SDF sums, marching-cubes triangles
  pipeline_small.npz  per-frame poses + status of the full ICP and SDF pipelines, final volume digest
  color_small.npz     `python tests/golden/make_golden.py <outdir> color`: the colour path (use_color=1, with and
                      without angle weighting): volume + voxel colours after 3 integrations, raycast maps + raycast_rgb,
                      marching-cubes triangles with vertex colours
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _oracle as O  # noqa: E402
from hybkinectfu_b200 import synth  # noqa: E402
from hybkinectfu_b200.capi import Config  # noqa: E402


def small_config(resolution=64, levels=3):
    cam = synth.Camera().scaled(192, 144)
    return Config(cols=192, rows=144, cx=cam.cx, cy=cam.cy, fx=cam.fx, fy=cam.fy, resolution=resolution,
                  pyramid_levels=levels, max_triangles=400000), cam


def small_sequence(n_frames, stride=4):
    cfg, cam = small_config()
    poses = synth.trajectory(200)[::stride][:n_frames]
    frames = np.stack([synth.render_depth_mm(p, cam) for p in poses])
    return frames, poses.astype(np.float32)


def canon_tris(t):
    """order-independent form of a triangle soup: sort triangles by their 9 position floats"""
    if len(t) == 0:
        return np.zeros((0, 9), np.float32)
    p = np.ascontiguousarray(t[:, :, :3]).reshape(len(t), 9)
    key = np.lexsort(p.T[::-1])
    return p[key]


def canon_tris_color(t):
    """order-independent form of a coloured triangle soup: rows of 18 floats (pos, colour per vertex), sorted"""
    if len(t) == 0:
        return np.zeros((0, 18), np.float32)
    p = np.ascontiguousarray(t).reshape(len(t), 18)
    key = np.lexsort(p.T[::-1])
    return p[key]


def color_sequence(n_frames, stride=4):
    cfg, cam = small_config()
    poses = synth.trajectory(200)[::stride][:n_frames]
    return np.stack([synth.render_color(p, cam) for p in poses])


def main_color(outdir):
    os.makedirs(outdir, exist_ok=True)
    frames, poses = small_sequence(4)
    bgr = color_sequence(4)
    cfg, cam = small_config(64)
    g = {"frames": frames, "poses": poses, "bgr": bgr}
    for angled in (0, 1):
        ref = O.Ref(cfg)
        ref.set_color(1, angled)
        ref.reset_volume()
        for i in range(3):
            ref.preprocess(frames[i])
            ref.upload_rgb(bgr[i])
            ref.integrate(poses[i])
        g[f"new_n0_{angled}"] = ref.download(O.REF_NEW_NORMALS, 0)
        g[f"trunc_{angled}"] = ref.download(O.REF_TRUNC_DEPTH)
        g[f"vol_{angled}"] = ref.download(O.REF_VOLUME)
        g[f"vol_color_{angled}"] = ref.download(O.REF_VOLUME_COLOR)
        ref.raycast(poses[2])
        g[f"model_v0_{angled}"] = ref.download(O.REF_MODEL_VERTICES, 0)
        g[f"model_n0_{angled}"] = ref.download(O.REF_MODEL_NORMALS, 0)
        g[f"raycast_rgb_{angled}"] = ref.download(O.REF_RAYCAST_RGB)
        n = ref.marching_cubes()
        tris = ref.download(O.REF_TRIANGLES)
        assert len(tris) == n
        g[f"mc_count_{angled}"] = np.array([n], np.int64)
        g[f"mc_tris_sorted_{angled}"] = canon_tris_color(tris)
        ref.set_color(0, 0)
        ref.check()
        print(f"color (angled={angled}): coloured voxels", int((g[f"vol_color_{angled}"][..., :3] != 0).any(-1).sum()),
              "observed", int((g[f"vol_{angled}"][..., 1] > 0).sum()), "raycast rgb px", int((g[f"raycast_rgb_{angled}"] != 0).any(-1).sum()),
              "triangles", n)
    np.savez_compressed(os.path.join(outdir, "color_small.npz"), **g)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    frames, poses = small_sequence(8)
    # ---------------- per-stage fixture
    cfg, cam = small_config(64)
    ref = O.Ref(cfg)
    g = {"frames": frames, "poses": poses}
    ref.preprocess(frames[0])
    g["raw"] = ref.download(O.REF_RAW_DEPTH)
    g["trunc"] = ref.download(O.REF_TRUNC_DEPTH)
    g["filt"] = ref.download(O.REF_FILTERED_DEPTH)
    for l in range(3):
        g[f"new_v{l}"] = ref.download(O.REF_NEW_VERTICES, l)
        g[f"new_n{l}"] = ref.download(O.REF_NEW_NORMALS, l)
    ref.reset_volume()
    for i in range(3):
        ref.preprocess(frames[i])
        ref.integrate(poses[i])
        g[f"vol_after{i}"] = ref.download(O.REF_VOLUME)
    ref.raycast(poses[2])
    ref.downsample_model()
    for l in range(3):
        g[f"model_v{l}"] = ref.download(O.REF_MODEL_VERTICES, l)
        g[f"model_n{l}"] = ref.download(O.REF_MODEL_NORMALS, l)
    # ICP sums: new maps of frame 3 against the model raycast from pose 2, starting at pose 2
    ref.preprocess(frames[3])
    for l in range(3):
        g[f"icp_new_v{l}"] = ref.download(O.REF_NEW_VERTICES, l)
        g[f"icp_new_n{l}"] = ref.download(O.REF_NEW_NORMALS, l)
    g["icp_trunc"] = ref.download(O.REF_TRUNC_DEPTH)
    last_inv = O.mat44_inverse(poses[2])
    g["icp_cur"] = poses[2]
    g["icp_last_inv"] = last_inv
    for l in range(3):
        g[f"icp_sums{l}"] = ref.icp_params(l, poses[2], last_inv)
    g["sdf_sums"] = ref.sdf_params(poses[2])
    n = ref.marching_cubes()
    tris = ref.download(O.REF_TRIANGLES)
    assert len(tris) == n
    g["mc_count"] = np.array([n], np.int64)
    g["mc_tris_sorted"] = canon_tris(tris)
    ref.check()
    np.savez_compressed(os.path.join(outdir, "stages_small.npz"), **g)
    print("stages_small: mc triangles", n, "valid raycast px", int((g["model_n0"][..., :3] != 0).any(-1).sum()))

    # ---------------- full pipelines (R = 128)
    cfg, cam = small_config(128)
    h = {"frames": frames, "gt_poses": poses}
    for mode, name in ((0, "icp"), (1, "sdf"), (2, "file")):
        ref = O.Ref(cfg)
        ref.reset_volume()
        ps, sts = [], []
        for i in range(len(frames)):
            p, st = ref.process(frames[i], i, mode, poses[i] if mode == 2 else None)
            ps.append(p)
            sts.append(st)
        vol = ref.download(O.REF_VOLUME)
        h[f"{name}_poses"] = np.stack(ps)
        h[f"{name}_status"] = np.array(sts, np.int32)
        h[f"{name}_vol_sha256"] = np.array([digest(vol)])
        h[f"{name}_vol_observed"] = np.array([int((vol[..., 1] > 0).sum())], np.int64)
        h[f"{name}_vol_tsdf_sum"] = np.array([float(vol[..., 0].astype(np.float64).sum())])
        h[f"{name}_vol_plane64"] = vol[64].copy()
        n = ref.marching_cubes()
        h[f"{name}_mc_count"] = np.array([n], np.int64)
        ref.check()
        print(name, "status", sts, "mc", n, "terr(mm)", [round(float(np.linalg.norm(ps[i][:3, 3] - poses[i][:3, 3]) * 1e3), 2) for i in range(len(frames))])
    np.savez_compressed(os.path.join(outdir, "pipeline_small.npz"), **h)


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "color":
    main_color(sys.argv[1])
    sys.exit(0)

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
