"""ctypes binding of include/hkf.h (libhkf_b200.so).

This is plumbing for tests and the benchmark driver: every call goes straight through the C ABI
that the reference's host classes would bind.  There is no fallback of any kind: if the shared
library is missing the import fails, and if no sm_100 GPU is usable hkf_create() fails.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HKF_LIB") or os.path.join(_HERE, "libhkf_b200.so")      # HKF_LIB: a profiling build of the same library


class HkfError(RuntimeError):
    pass


class CameraParams(C.Structure):
    _fields_ = [("cols", C.c_uint32), ("rows", C.c_uint32), ("cx", C.c_float), ("cy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float)]


class Mat44(C.Structure):
    _fields_ = [("m", C.c_float * 16)]

    @staticmethod
    def from_np(a) -> "Mat44":
        m = Mat44()
        flat = np.asarray(a, dtype=np.float32).reshape(16)
        for i in range(16):
            m.m[i] = float(flat[i])
        return m

    def to_np(self) -> np.ndarray:
        return np.array(list(self.m), dtype=np.float32).reshape(4, 4)


class IntegrateParams(C.Structure):
    _fields_ = [("fSdfTruncation", C.c_float), ("fMaxIntegrateDist", C.c_float)]


class RayCasterParams(C.Structure):
    _fields_ = [("fRayIncrement", C.c_float)]


class InitParams(C.Structure):
    _fields_ = [("depth_camera", CameraParams), ("pyramid_levels", C.c_uint32), ("volume_resolution", C.c_uint32),
                ("volume_size", C.c_float), ("weight_max", C.c_float), ("max_triangles", C.c_uint32),
                ("slab_z0", C.c_uint32), ("slab_z1", C.c_uint32), ("slab_halo", C.c_uint32),
                ("use_color", C.c_uint32), ("rgb_camera", CameraParams)]


class FrameParams(C.Structure):
    _fields_ = [("trunc_min", C.c_float), ("trunc_max", C.c_float), ("sigma_pixel", C.c_float), ("sigma_depth", C.c_float),
                ("integrate", IntegrateParams), ("raycast", RayCasterParams),
                ("icp_dist_thres", C.c_float), ("icp_norm_sin_thres", C.c_float), ("icp_dist_shake", C.c_float),
                ("icp_angle_shake", C.c_float), ("sdf_max_iter", C.c_uint32), ("sdf_dist_shake", C.c_float),
                ("sdf_angle_shake", C.c_float), ("use_color", C.c_uint32), ("color_angle_weight", C.c_uint32)]


# enum HkfBuffer
BUF_RAW_DEPTH, BUF_TRUNC_DEPTH, BUF_FILTERED_DEPTH, BUF_NEW_VERTICES, BUF_NEW_NORMALS, BUF_MODEL_VERTICES, \
    BUF_MODEL_NORMALS, BUF_VOLUME, BUF_SOLVER_SUMS, BUF_TRIANGLES, BUF_POSE, BUF_RAYCAST_KEYS, BUF_RAW_RGB, \
    BUF_RAYCAST_RGB, BUF_VOLUME_COLOR = range(15)
TRACK_OK, TRACK_SINGULAR, TRACK_SHAKE = 0, 1, 2

# every symbol include/hkf.h declares (checked by tests/test_abi.py against the header text)
EXPORTS = [
    "hkf_create", "hkf_destroy", "hkf_last_error", "hkf_sync", "hkf_stream", "hkf_reset_volume", "hkf_download",
    "hkf_upload", "hkf_buffer_bytes", "hkf_launch_count", "hkf_upload_depth_mm", "hkf_trunc_depth",
    "hkf_bilateral_filter_depth", "hkf_calculate_new_vertices", "hkf_calculate_new_normals",
    "hkf_downsample_new_vertices", "hkf_downsample_new_normals", "hkf_downsample_model_vertices",
    "hkf_downsample_model_normals", "hkf_cal_point_to_plane_err_solver_params", "hkf_cal_sdf_solver_params",
    "hkf_integrate_volume", "hkf_raycasting_volume", "hkf_marchingcube", "hkf_triangle_count", "hkf_preprocess",
    "hkf_set_pose", "hkf_icp_track", "hkf_sdf_track", "hkf_integrate_tracked", "hkf_raycast_tracked",
    "hkf_frame_result", "hkf_count_integrate_updates", "hkf_shard_unique_id", "hkf_shard_init",
    "hkf_shard_composite_raycast", "hkf_shard_ipc_handles", "hkf_shard_connect", "hkf_shard_peer_halo", "hkf_preprocess_next", "hkf_swap_frame",
    "hkf_upload_color", "hkf_integrate_profile", "hkf_integrate_profile_read", "hkf_shard_set_icp_mode",
]

_lib = None


def load() -> C.CDLL:
    """dlopen libhkf_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HkfError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C hybkinectfu_b200` — there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.hkf_create.argtypes = [C.POINTER(InitParams), C.c_int, C.POINTER(P)]
    lib.hkf_destroy.argtypes = [P]
    lib.hkf_destroy.restype = None
    lib.hkf_last_error.argtypes = [P]
    lib.hkf_last_error.restype = C.c_char_p
    lib.hkf_sync.argtypes = [P]
    lib.hkf_stream.argtypes = [P]
    lib.hkf_stream.restype = C.c_void_p
    lib.hkf_reset_volume.argtypes = [P]
    lib.hkf_download.argtypes = [P, C.c_int, C.c_uint, C.c_void_p, C.c_size_t]
    lib.hkf_download.restype = C.c_long
    lib.hkf_upload.argtypes = [P, C.c_int, C.c_uint, C.c_void_p, C.c_size_t]
    lib.hkf_upload.restype = C.c_long
    lib.hkf_buffer_bytes.argtypes = [P, C.c_int, C.c_uint]
    lib.hkf_buffer_bytes.restype = C.c_size_t
    lib.hkf_launch_count.argtypes = [P]
    lib.hkf_launch_count.restype = C.c_uint64
    lib.hkf_upload_depth_mm.argtypes = [P, C.c_void_p, C.c_int]
    lib.hkf_upload_color.argtypes = [P, C.c_void_p, C.c_int]
    lib.hkf_trunc_depth.argtypes = [P, C.c_float, C.c_float]
    lib.hkf_bilateral_filter_depth.argtypes = [P, C.c_float, C.c_float]
    lib.hkf_calculate_new_vertices.argtypes = [P, C.POINTER(CameraParams)]
    for n in ("hkf_calculate_new_normals", "hkf_downsample_new_vertices", "hkf_downsample_new_normals",
              "hkf_downsample_model_vertices", "hkf_downsample_model_normals", "hkf_shard_composite_raycast"):
        getattr(lib, n).argtypes = [P]
    lib.hkf_cal_point_to_plane_err_solver_params.argtypes = [P, C.c_uint, C.POINTER(Mat44), C.POINTER(Mat44),
                                                             C.POINTER(CameraParams), C.c_float, C.c_float]
    lib.hkf_cal_sdf_solver_params.argtypes = [P, C.POINTER(CameraParams), C.POINTER(Mat44)]
    lib.hkf_integrate_volume.argtypes = [P, C.c_int, C.c_int, C.POINTER(Mat44), C.POINTER(IntegrateParams),
                                         C.POINTER(CameraParams), C.POINTER(CameraParams)]
    lib.hkf_raycasting_volume.argtypes = [P, C.c_int, C.POINTER(Mat44), C.POINTER(RayCasterParams),
                                          C.POINTER(CameraParams), C.c_float, C.c_float]
    lib.hkf_marchingcube.argtypes = [P, C.c_int, C.c_float]
    lib.hkf_triangle_count.argtypes = [P, C.POINTER(C.c_uint)]
    lib.hkf_preprocess.argtypes = [P, C.c_void_p, C.c_int, C.POINTER(FrameParams)]
    lib.hkf_preprocess_next.argtypes = [P, C.c_void_p, C.c_int, C.POINTER(FrameParams)]
    lib.hkf_swap_frame.argtypes = [P]
    lib.hkf_set_pose.argtypes = [P, C.POINTER(Mat44)]
    for n in ("hkf_icp_track", "hkf_sdf_track", "hkf_integrate_tracked", "hkf_raycast_tracked"):
        getattr(lib, n).argtypes = [P, C.POINTER(FrameParams)]
    lib.hkf_frame_result.argtypes = [P, C.POINTER(Mat44), C.POINTER(C.c_int)]
    lib.hkf_count_integrate_updates.argtypes = [P, C.POINTER(Mat44), C.POINTER(IntegrateParams),
                                                C.POINTER(CameraParams), C.POINTER(C.c_ulonglong)]
    lib.hkf_integrate_profile.argtypes = [P, C.c_int]
    lib.hkf_integrate_profile_read.argtypes = [P, C.c_void_p, C.c_size_t]
    lib.hkf_integrate_profile_read.restype = C.c_long
    lib.hkf_shard_set_icp_mode.argtypes = [P, C.c_int]
    lib.hkf_shard_unique_id.argtypes = [C.c_void_p]
    lib.hkf_shard_init.argtypes = [P, C.c_int, C.c_int, C.c_void_p]
    lib.hkf_shard_ipc_handles.argtypes = [P, C.c_void_p]
    lib.hkf_shard_connect.argtypes = [P, C.c_void_p, C.c_int]
    lib.hkf_shard_peer_halo.argtypes = [P, C.c_uint, C.c_void_p, C.c_int]
    _lib = lib
    return lib


@dataclass
class Config:
    """The AppParams values the hot path reads (src/AppParams.h), defaults = root config.ini:10-48
    with use_color=0 and max_iter_nums from src/config.ini:44."""
    cols: int = 640
    rows: int = 480
    cx: float = 319.5
    cy: float = 239.5
    fx: float = 525.0
    fy: float = 525.0
    trunc_min: float = 0.3
    trunc_max: float = 4.0
    sigma_pixel: float = 2.0
    sigma_depth: float = 0.02
    resolution: int = 256
    volume_size: float = 2.0
    weight_max: float = 128.0
    sdf_trunc: float = 0.04
    max_integrate_dist: float = 2.0
    raycast_increment_factor: float = 0.8
    pyramid_levels: int = 3
    icp_dist_thres: float = 0.1
    icp_norm_sin_thres: float = 0.2
    icp_dist_shake: float = 0.4
    icp_angle_shake: float = 0.4
    sdf_max_iter: int = 6
    sdf_dist_shake: float = 0.4
    sdf_angle_shake: float = 0.4
    max_triangles: int = 1500000
    slab_z0: int = 0
    slab_z1: int = 0
    slab_halo: int = 0
    use_color: int = 0              # _switch_params.useRGBData (the rgb camera = the depth camera, as in config.ini)
    color_angle_weight: int = 0     # _switch_params.colorAngleWeight

    @property
    def ray_increment(self) -> float:
        # AppParamsProducer: fRayIncrement = factor * fSdfTruncation in float (src/AppParamsProducer.cpp:113-117)
        return float(np.float32(self.raycast_increment_factor) * np.float32(self.sdf_trunc))

    def camera(self) -> CameraParams:
        return CameraParams(self.cols, self.rows, self.cx, self.cy, self.fx, self.fy)

    def init_params(self) -> InitParams:
        return InitParams(self.camera(), self.pyramid_levels, self.resolution, self.volume_size, self.weight_max,
                          self.max_triangles, self.slab_z0, self.slab_z1, self.slab_halo, self.use_color, self.camera())

    def frame_params(self) -> FrameParams:
        return FrameParams(self.trunc_min, self.trunc_max, self.sigma_pixel, self.sigma_depth,
                           IntegrateParams(self.sdf_trunc, self.max_integrate_dist), RayCasterParams(self.ray_increment),
                           self.icp_dist_thres, self.icp_norm_sin_thres, self.icp_dist_shake, self.icp_angle_shake,
                           self.sdf_max_iter, self.sdf_dist_shake, self.sdf_angle_shake, self.use_color, self.color_angle_weight)

    def level_shape(self, level: int):
        return self.rows >> level, self.cols >> level

    def level_camera(self, level: int) -> CameraParams:
        # CameraPoseFinderICP::initPoseFinder halves per level (src/CameraPoseFinderICP.cpp:36-46)
        c = self.camera()
        cols, rows = c.cols, c.rows
        cx, cy, fx, fy = np.float32(c.cx), np.float32(c.cy), np.float32(c.fx), np.float32(c.fy)
        for _ in range(level):
            cols //= 2
            rows //= 2
            cx, cy, fx, fy = cx / np.float32(2), cy / np.float32(2), fx / np.float32(2), fy / np.float32(2)
        return CameraParams(cols, rows, float(cx), float(cy), float(fx), float(fy))


class Context:
    """RAII wrapper of hkf_ctx."""

    def __init__(self, cfg: Config, device: int = 0):
        self.lib = load()
        self.cfg = cfg
        self._h = C.c_void_p()
        ip = cfg.init_params()
        rc = self.lib.hkf_create(C.byref(ip), device, C.byref(self._h))
        if rc != 0:
            raise HkfError(f"hkf_create failed ({rc}): {self.lib.hkf_last_error(None).decode()}")
        self.fp = cfg.frame_params()

    def close(self):
        if self._h:
            self.lib.hkf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise HkfError(f"{what} failed ({rc}): {self.lib.hkf_last_error(self._h).decode()}")

    # ---- buffers
    def download(self, buf: int, level: int = 0) -> np.ndarray:
        nb = self.lib.hkf_buffer_bytes(self._h, buf, level)
        raw = np.empty(nb, dtype=np.uint8)
        got = self.lib.hkf_download(self._h, buf, level, raw.ctypes.data, nb)
        if got < 0:
            raise HkfError(f"hkf_download({buf}) failed ({got}): {self.lib.hkf_last_error(self._h).decode()}")
        raw = raw[:got]
        rows, cols = self.cfg.level_shape(level)
        if buf in (BUF_RAW_DEPTH, BUF_TRUNC_DEPTH, BUF_FILTERED_DEPTH):
            return raw.view(np.float32).reshape(self.cfg.rows, self.cfg.cols)
        if buf in (BUF_NEW_VERTICES, BUF_NEW_NORMALS, BUF_MODEL_VERTICES, BUF_MODEL_NORMALS):
            return raw.view(np.float32).reshape(rows, cols, 4)
        if buf == BUF_VOLUME:
            R = self.cfg.resolution
            return raw.view(np.float32).reshape(-1, R, R, 2)
        if buf == BUF_TRIANGLES:
            return raw.view(np.float32).reshape(-1, 3, 6)
        if buf == BUF_RAYCAST_KEYS:
            return raw.view(np.uint64).reshape(self.cfg.rows, self.cfg.cols)
        if buf in (BUF_RAW_RGB, BUF_RAYCAST_RGB):
            return raw.reshape(self.cfg.rows, self.cfg.cols, 3)
        if buf == BUF_VOLUME_COLOR:
            R = self.cfg.resolution
            return raw.reshape(-1, R, R, 4)
        return raw.view(np.float32)

    def upload(self, buf: int, arr: np.ndarray, level: int = 0):
        a = np.ascontiguousarray(arr)
        got = self.lib.hkf_upload(self._h, buf, level, a.ctypes.data, a.nbytes)
        if got < 0:
            raise HkfError(f"hkf_upload({buf}) failed ({got}): {self.lib.hkf_last_error(self._h).decode()}")

    def sync(self):
        self._ck(self.lib.hkf_sync(self._h), "hkf_sync")

    def launches(self) -> int:
        return int(self.lib.hkf_launch_count(self._h))

    def reset_volume(self):
        self._ck(self.lib.hkf_reset_volume(self._h), "hkf_reset_volume")

    # ---- wrapper-parity layer
    def upload_depth_mm(self, mm: np.ndarray):
        a = np.ascontiguousarray(mm, dtype=np.uint16)
        self._ck(self.lib.hkf_upload_depth_mm(self._h, a.ctypes.data, 0), "hkf_upload_depth_mm")
        self.sync()      # `a` may be a temporary

    def upload_color(self, bgr: np.ndarray):
        a = np.ascontiguousarray(bgr, dtype=np.uint8)
        assert a.shape == (self.cfg.rows, self.cfg.cols, 3)
        self._ck(self.lib.hkf_upload_color(self._h, a.ctypes.data, 0), "hkf_upload_color")
        self.sync()      # `a` may be a temporary

    def trunc_depth(self):
        self._ck(self.lib.hkf_trunc_depth(self._h, self.cfg.trunc_min, self.cfg.trunc_max), "hkf_trunc_depth")

    def bilateral(self):
        self._ck(self.lib.hkf_bilateral_filter_depth(self._h, self.cfg.sigma_pixel, self.cfg.sigma_depth), "hkf_bilateral_filter_depth")

    def vertices(self):
        cam = self.cfg.camera()
        self._ck(self.lib.hkf_calculate_new_vertices(self._h, C.byref(cam)), "hkf_calculate_new_vertices")

    def normals(self):
        self._ck(self.lib.hkf_calculate_new_normals(self._h), "hkf_calculate_new_normals")

    def downsample_new(self):
        self._ck(self.lib.hkf_downsample_new_vertices(self._h), "hkf_downsample_new_vertices")
        self._ck(self.lib.hkf_downsample_new_normals(self._h), "hkf_downsample_new_normals")

    def downsample_model(self):
        self._ck(self.lib.hkf_downsample_model_vertices(self._h), "hkf_downsample_model_vertices")
        self._ck(self.lib.hkf_downsample_model_normals(self._h), "hkf_downsample_model_normals")

    def icp_params(self, level: int, cur, last_inv) -> np.ndarray:
        cam = self.cfg.level_camera(level)
        a, b = Mat44.from_np(cur), Mat44.from_np(last_inv)
        self._ck(self.lib.hkf_cal_point_to_plane_err_solver_params(self._h, level, C.byref(a), C.byref(b), C.byref(cam),
                                                                   self.cfg.icp_dist_thres, self.cfg.icp_norm_sin_thres),
                 "hkf_cal_point_to_plane_err_solver_params")
        return self.download(BUF_SOLVER_SUMS)[:27].copy()

    def sdf_params(self, cur) -> np.ndarray:
        cam = self.cfg.camera()
        a = Mat44.from_np(cur)
        self._ck(self.lib.hkf_cal_sdf_solver_params(self._h, C.byref(cam), C.byref(a)), "hkf_cal_sdf_solver_params")
        return self.download(BUF_SOLVER_SUMS)[:27].copy()

    def integrate(self, pose):
        cam = self.cfg.camera()
        ip = IntegrateParams(self.cfg.sdf_trunc, self.cfg.max_integrate_dist)
        m = Mat44.from_np(pose)
        self._ck(self.lib.hkf_integrate_volume(self._h, self.cfg.use_color, self.cfg.color_angle_weight if self.cfg.use_color else 0,
                                               C.byref(m), C.byref(ip), C.byref(cam), C.byref(cam)), "hkf_integrate_volume")

    def profile(self, enable: bool):
        self._ck(self.lib.hkf_integrate_profile(self._h, 1 if enable else 0), "hkf_integrate_profile")

    def profile_read(self) -> np.ndarray:
        """[stored planes][general chunks, free bricks] accumulated since profile(True)"""
        n = 2 * ((min(self.cfg.slab_z1 + self.cfg.slab_halo, self.cfg.resolution) - max(self.cfg.slab_z0 - self.cfg.slab_halo, 0))
                 if (self.cfg.slab_z0 or self.cfg.slab_z1) else self.cfg.resolution)
        out = np.zeros(n, np.uint32)
        got = self.lib.hkf_integrate_profile_read(self._h, out.ctypes.data, n)
        if got < 0:
            raise HkfError(f"hkf_integrate_profile_read failed ({got}): {self.lib.hkf_last_error(self._h).decode()}")
        return out[:got].reshape(-1, 2)

    def count_updates(self, pose) -> int:
        cam = self.cfg.camera()
        ip = IntegrateParams(self.cfg.sdf_trunc, self.cfg.max_integrate_dist)
        m = Mat44.from_np(pose)
        n = C.c_ulonglong(0)
        self._ck(self.lib.hkf_count_integrate_updates(self._h, C.byref(m), C.byref(ip), C.byref(cam), C.byref(n)), "hkf_count_integrate_updates")
        return int(n.value)

    def raycast(self, pose):
        cam = self.cfg.camera()
        rp = RayCasterParams(self.cfg.ray_increment)
        m = Mat44.from_np(pose)
        self._ck(self.lib.hkf_raycasting_volume(self._h, self.cfg.use_color, C.byref(m), C.byref(rp), C.byref(cam), self.cfg.trunc_min, self.cfg.trunc_max), "hkf_raycasting_volume")

    def marching_cubes(self) -> int:
        # MeshGeneratorMarchingcube::generateMesh threshold 300*S/R (src/MeshGeneratorMarchingcube.cpp:26-27)
        thr = float(np.float32(300) * np.float32(self.cfg.volume_size) / np.float32(self.cfg.resolution))
        self._ck(self.lib.hkf_marchingcube(self._h, self.cfg.use_color, thr), "hkf_marchingcube")
        n = C.c_uint(0)
        self._ck(self.lib.hkf_triangle_count(self._h, C.byref(n)), "hkf_triangle_count")
        return int(n.value)

    # ---- fused layer
    def preprocess(self, mm: np.ndarray):
        a = np.ascontiguousarray(mm, dtype=np.uint16)
        self._ck(self.lib.hkf_preprocess(self._h, a.ctypes.data, 0, C.byref(self.fp)), "hkf_preprocess")
        self.sync()

    def preprocess_device(self, dev_ptr: int):
        self._ck(self.lib.hkf_preprocess(self._h, C.c_void_p(dev_ptr), 1, C.byref(self.fp)), "hkf_preprocess")

    def preprocess_host_async(self, host_ptr: int):
        self._ck(self.lib.hkf_preprocess(self._h, C.c_void_p(host_ptr), 0, C.byref(self.fp)), "hkf_preprocess")

    def preprocess_next(self, ptr: int, is_device: bool):
        """upload + preprocess the NEXT frame on the side stream (hkf_preprocess_next); swap_frame() makes it current"""
        self._ck(self.lib.hkf_preprocess_next(self._h, C.c_void_p(ptr), 1 if is_device else 0, C.byref(self.fp)), "hkf_preprocess_next")

    def swap_frame(self):
        self._ck(self.lib.hkf_swap_frame(self._h), "hkf_swap_frame")

    def set_pose(self, pose):
        m = Mat44.from_np(pose)
        self._ck(self.lib.hkf_set_pose(self._h, C.byref(m)), "hkf_set_pose")

    def icp_track(self):
        self._ck(self.lib.hkf_icp_track(self._h, C.byref(self.fp)), "hkf_icp_track")

    def sdf_track(self):
        self._ck(self.lib.hkf_sdf_track(self._h, C.byref(self.fp)), "hkf_sdf_track")

    def integrate_tracked(self):
        self._ck(self.lib.hkf_integrate_tracked(self._h, C.byref(self.fp)), "hkf_integrate_tracked")

    def raycast_tracked(self):
        self._ck(self.lib.hkf_raycast_tracked(self._h, C.byref(self.fp)), "hkf_raycast_tracked")

    def frame_result(self):
        m = Mat44()
        st = C.c_int(0)
        self._ck(self.lib.hkf_frame_result(self._h, C.byref(m), C.byref(st)), "hkf_frame_result")
        return m.to_np(), int(st.value)
