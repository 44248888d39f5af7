"""Synthetic world for parity tests and benchmarks.

An analytic scene inside the S = 2 m volume cube, rendered to 16-bit millimetre depth frames (the
DepthFrameData contract of the reference, src/FrameData.h:57,70) along a smooth trajectory that
starts exactly at the reference's initial pose (src/HybKinectfu.cpp:51-54).  Everything is
deterministic double-precision NumPy; the u16 frames are the single source fed to the reference
oracle, the CPU oracle and the CUDA path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# scene: planes as (axis, offset) and spheres as (centre, radius)
PLANES = ((1, 1.7), (2, 1.6), (0, 0.2))          # floor y=1.7, back wall z=1.6, side wall x=0.2
SPHERES = (((1.0, 1.1, 1.2), 0.35), ((1.45, 1.45, 0.9), 0.2))
SCENE_CENTRE = np.array([1.0, 1.0, 1.2])


@dataclass
class Camera:
    cols: int = 640
    rows: int = 480
    cx: float = 319.5
    cy: float = 239.5
    fx: float = 525.0
    fy: float = 525.0

    def scaled(self, cols: int, rows: int) -> "Camera":
        """Same field of view at another resolution (pixel centres at integer coordinates)."""
        sx, sy = cols / self.cols, rows / self.rows
        return Camera(cols, rows, (self.cx + 0.5) * sx - 0.5, (self.cy + 0.5) * sy - 0.5, self.fx * sx, self.fy * sy)


def pose0(volume_size: float = 2.0, trunc_min: float = 0.3) -> np.ndarray:
    p = np.eye(4)
    p[:3, 3] = (volume_size / 2.0, volume_size / 2.0, -trunc_min)
    return p


def trajectory(n: int, phase: float = 0.0, volume_size: float = 2.0, trunc_min: float = 0.3) -> np.ndarray:
    """n camera->world poses (float64 4x4). Frame 0 is exactly pose0 when phase == 0."""
    poses = np.zeros((n, 4, 4))
    p0 = pose0(volume_size, trunc_min)
    base = None
    for i in range(n):
        a = 2.0 * math.pi * i / n + phase
        th = 0.25 * math.sin(a)
        d = np.array([0.10 * math.sin(a), 0.05 * math.sin(2 * a), 0.05 * (1.0 - math.cos(a))])
        ry = np.eye(4)
        c, s = math.cos(th), math.sin(th)
        ry[0, 0], ry[0, 2], ry[2, 0], ry[2, 2] = c, s, -s, c
        tc, tmc, td = np.eye(4), np.eye(4), np.eye(4)
        tc[:3, 3] = SCENE_CENTRE
        tmc[:3, 3] = -SCENE_CENTRE
        td[:3, 3] = d
        m = td @ tc @ ry @ tmc
        if base is None:
            base = np.linalg.inv(m)          # with a phase offset, re-anchor so that frame 0 == pose0
        poses[i] = m @ base @ p0
    return poses


def render_depth_mm(pose: np.ndarray, cam: Camera, noise_sigma_mm: float = 0.0, rng=None) -> np.ndarray:
    """uint16 [rows][cols] millimetre depth (camera z of the nearest hit), 0 = no hit."""
    u = np.arange(cam.cols, dtype=np.float64)
    v = np.arange(cam.rows, dtype=np.float64)
    uu, vv = np.meshgrid(u, v)
    dc = np.stack([(uu - cam.cx) / cam.fx, (vv - cam.cy) / cam.fy, np.ones_like(uu)], axis=-1)
    R, o = pose[:3, :3], pose[:3, 3]
    d = dc @ R.T                                   # world direction per unit camera z
    best = np.full(uu.shape, np.inf)
    for axis, off in PLANES:
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (off - o[axis]) / d[..., axis]
        s = np.where((s > 1e-9) & np.isfinite(s), s, np.inf)
        best = np.minimum(best, s)
    for centre, rad in SPHERES:
        oc = o - np.asarray(centre)
        a = np.sum(d * d, axis=-1)
        b = 2.0 * (d @ oc)
        c = float(oc @ oc) - rad * rad
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore"):
            sq = np.sqrt(np.where(disc >= 0, disc, np.nan))
            s1 = (-b - sq) / (2 * a)
            s2 = (-b + sq) / (2 * a)
        s = np.where(s1 > 1e-9, s1, s2)
        s = np.where((disc >= 0) & (s > 1e-9), s, np.inf)
        best = np.minimum(best, s)
    mm = best * 1000.0
    if noise_sigma_mm > 0:
        rng = rng or np.random.default_rng(0)
        mm = mm + rng.normal(0.0, noise_sigma_mm, size=mm.shape)
    mm = np.where(np.isfinite(mm) & (mm < 65535.0) & (mm > 0), np.rint(mm), 0.0)
    return mm.astype(np.uint16)


# surface colours (B, G, R like the OpenCV frames of the reference, src/HybKinectfu.cpp:90-91): planes, then spheres
SURFACE_BGR = ((90, 180, 200), (220, 160, 90), (120, 220, 120), (80, 80, 230), (60, 240, 240))


def render_color(pose: np.ndarray, cam: Camera) -> np.ndarray:
    """uint8 [rows][cols][3] colour image of the same analytic scene from the same camera (the reference registers
    colour to depth): every surface has a base colour, darkened on a 0.1 m checkerboard of the hit point and shaded
    with a smooth ramp along world x; pixels without a hit are black.  Deterministic double-precision NumPy."""
    u = np.arange(cam.cols, dtype=np.float64)
    v = np.arange(cam.rows, dtype=np.float64)
    uu, vv = np.meshgrid(u, v)
    dc = np.stack([(uu - cam.cx) / cam.fx, (vv - cam.cy) / cam.fy, np.ones_like(uu)], axis=-1)
    R, o = pose[:3, :3], pose[:3, 3]
    d = dc @ R.T
    best = np.full(uu.shape, np.inf)
    surf = np.full(uu.shape, -1, dtype=np.int64)
    k = 0
    for axis, off in PLANES:
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (off - o[axis]) / d[..., axis]
        s = np.where((s > 1e-9) & np.isfinite(s), s, np.inf)
        surf = np.where(s < best, k, surf)
        best = np.minimum(best, s)
        k += 1
    for centre, rad in SPHERES:
        oc = o - np.asarray(centre)
        a = np.sum(d * d, axis=-1)
        b = 2.0 * (d @ oc)
        c = float(oc @ oc) - rad * rad
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore"):
            sq = np.sqrt(np.where(disc >= 0, disc, np.nan))
            s1 = (-b - sq) / (2 * a)
            s2 = (-b + sq) / (2 * a)
        s = np.where(s1 > 1e-9, s1, s2)
        s = np.where((disc >= 0) & (s > 1e-9), s, np.inf)
        surf = np.where(s < best, k, surf)
        best = np.minimum(best, s)
        k += 1
    hit = np.isfinite(best)
    p = o + d * np.where(hit, best, 0.0)[..., None]
    cell = np.floor(p / 0.1 + 1e-6).astype(np.int64)
    check = ((cell[..., 0] + cell[..., 1] + cell[..., 2]) & 1).astype(np.float64)
    shade = (1.0 - 0.4 * check) * (0.75 + 0.25 * np.clip(p[..., 0] / 2.0, 0.0, 1.0))
    base = np.asarray(SURFACE_BGR, dtype=np.float64)[np.clip(surf, 0, len(SURFACE_BGR) - 1)]
    img = np.where(hit[..., None], np.rint(base * shade[..., None]), 0.0)
    return np.clip(img, 0, 255).astype(np.uint8)


def render_sequence(n: int, cam: Camera, phase: float = 0.0, noise_sigma_mm: float = 0.0):
    poses = trajectory(n, phase)
    rng = np.random.default_rng(0)
    frames = np.stack([render_depth_mm(poses[i], cam, noise_sigma_mm, rng) for i in range(n)])
    return frames, poses


def mat_to_quat(Rm: np.ndarray):
    """rotation matrix -> (qx,qy,qz,qw)"""
    t = np.trace(Rm)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        return ((Rm[2, 1] - Rm[1, 2]) / s, (Rm[0, 2] - Rm[2, 0]) / s, (Rm[1, 0] - Rm[0, 1]) / s, 0.25 * s)
    i = int(np.argmax(np.diag(Rm)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(Rm[i, i] - Rm[j, j] - Rm[k, k] + 1.0) * 2
    q = [0.0, 0.0, 0.0, 0.0]
    q[i] = 0.25 * s
    q[j] = (Rm[j, i] + Rm[i, j]) / s
    q[k] = (Rm[k, i] + Rm[i, k]) / s
    q[3] = (Rm[k, j] - Rm[j, k]) / s
    return tuple(q)


def write_tum_trajectory(path: str, poses: np.ndarray, fps: float = 30.0) -> None:
    """TUM format with the 3 header lines CameraPoseFinderFromFile skips
    (src/CameraPoseFinderFromFile.cpp:26-29): timestamp tx ty tz qx qy qz qw"""
    with open(path, "w") as f:
        f.write("# ground truth trajectory\n# file: synthetic\n# timestamp tx ty tz qx qy qz qw\n")
        for i, p in enumerate(poses):
            qx, qy, qz, qw = mat_to_quat(p[:3, :3])
            t = p[:3, 3]
            f.write(f"{i / fps:.6f} {t[0]:.9f} {t[1]:.9f} {t[2]:.9f} {qx:.9f} {qy:.9f} {qz:.9f} {qw:.9f}\n")
