"""ctypes bindings of the parity oracles (TEST INFRASTRUCTURE):
  * oracle/liboracle.so        — the plain-C CPU restatement (oracle/hkf_oracle.c)
  * oracle/_ref/libhkf_ref.so  — the UNMODIFIED reference CUDA kernels behind oracle/ref_harness.cu
                                 (GPU box only; built in the dev container from /root/reference)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
CPU_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libhkf_ref.so")

F32P = C.POINTER(C.c_float)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(F32P)


class OrcCamera(C.Structure):
    _fields_ = [("cols", C.c_uint32), ("rows", C.c_uint32), ("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float)]


class OrcParams(C.Structure):
    _fields_ = [("cam", OrcCamera), ("trunc_min", C.c_float), ("trunc_max", C.c_float), ("sigma_pixel", C.c_float),
                ("sigma_depth", C.c_float), ("resolution", C.c_int), ("volume_size", C.c_float), ("weight_max", C.c_float),
                ("sdf_trunc", C.c_float), ("max_integrate_dist", C.c_float), ("ray_increment", C.c_float),
                ("pyramid_levels", C.c_int), ("icp_dist_thres", C.c_float), ("icp_norm_sin_thres", C.c_float),
                ("icp_dist_shake", C.c_float), ("icp_angle_shake", C.c_float), ("sdf_max_iter", C.c_int),
                ("sdf_dist_shake", C.c_float), ("sdf_angle_shake", C.c_float)]


class RefParams(C.Structure):
    _fields_ = [("cols", C.c_uint32), ("rows", C.c_uint32), ("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("trunc_min", C.c_float), ("trunc_max", C.c_float), ("sigma_pixel", C.c_float), ("sigma_depth", C.c_float),
                ("resolution", C.c_uint32), ("volume_size", C.c_float), ("weight_max", C.c_float),
                ("sdf_trunc", C.c_float), ("max_integrate_dist", C.c_float), ("ray_increment", C.c_float),
                ("pyramid_levels", C.c_uint32), ("icp_dist_thres", C.c_float), ("icp_norm_sin_thres", C.c_float),
                ("icp_dist_shake", C.c_float), ("icp_angle_shake", C.c_float), ("sdf_max_iter", C.c_uint32),
                ("sdf_dist_shake", C.c_float), ("sdf_angle_shake", C.c_float), ("max_triangles", C.c_uint32)]


def cam_of(cfg, level: int = 0) -> OrcCamera:
    c = cfg.level_camera(level)
    return OrcCamera(c.cols, c.rows, c.cx, c.cy, c.fx, c.fy)


def orc_params(cfg) -> OrcParams:
    return OrcParams(cam_of(cfg), cfg.trunc_min, cfg.trunc_max, cfg.sigma_pixel, cfg.sigma_depth, cfg.resolution,
                     cfg.volume_size, cfg.weight_max, cfg.sdf_trunc, cfg.max_integrate_dist, cfg.ray_increment,
                     cfg.pyramid_levels, cfg.icp_dist_thres, cfg.icp_norm_sin_thres, cfg.icp_dist_shake,
                     cfg.icp_angle_shake, cfg.sdf_max_iter, cfg.sdf_dist_shake, cfg.sdf_angle_shake)


def ref_params(cfg) -> RefParams:
    return RefParams(cfg.cols, cfg.rows, cfg.cx, cfg.cy, cfg.fx, cfg.fy, cfg.trunc_min, cfg.trunc_max, cfg.sigma_pixel,
                     cfg.sigma_depth, cfg.resolution, cfg.volume_size, cfg.weight_max, cfg.sdf_trunc,
                     cfg.max_integrate_dist, cfg.ray_increment, cfg.pyramid_levels, cfg.icp_dist_thres,
                     cfg.icp_norm_sin_thres, cfg.icp_dist_shake, cfg.icp_angle_shake, cfg.sdf_max_iter,
                     cfg.sdf_dist_shake, cfg.sdf_angle_shake, cfg.max_triangles)


_cpu = None


def cpu() -> C.CDLL:
    """the C restatement; (re)built on demand with the system gcc"""
    global _cpu
    if _cpu is not None:
        return _cpu
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("hkf_oracle.c", "hkf_oracle_track.inc", "hkf_oracle.h", "host_math.h", "mc_tables_packed.h")]
    if (not os.path.exists(CPU_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(CPU_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(CPU_LIB)
    lib.orc_integrate.restype = C.c_uint64
    lib.orc_marching_cubes.restype = C.c_uint32
    lib.orc_tracker_create.restype = C.c_void_p
    lib.orc_tracker_create.argtypes = [C.POINTER(OrcParams)]
    lib.orc_tracker_destroy.argtypes = [C.c_void_p]
    lib.orc_tracker_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_void_p, F32P]
    lib.orc_tracker_volume.restype = F32P
    lib.orc_tracker_volume.argtypes = [C.c_void_p]
    lib.orc_tracker_map.restype = F32P
    lib.orc_tracker_map.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.orc_tracker_depth.restype = F32P
    lib.orc_tracker_depth.argtypes = [C.c_void_p, C.c_int]
    lib.orc_tracker_times.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    _cpu = lib
    return lib


# ---------------------------------------------------------------- numpy-level helpers (CPU oracle)
def decode_mm(mm: np.ndarray) -> np.ndarray:
    mm = np.ascontiguousarray(mm, dtype=np.uint16)
    out = np.empty(mm.shape, np.float32)
    cpu().orc_decode_mm(mm.ctypes.data_as(C.c_void_p), _fp(out), C.c_size_t(mm.size))
    return out


def trunc(raw, tmin, tmax):
    out = np.empty_like(raw)
    cpu().orc_trunc(_fp(raw), _fp(out), C.c_size_t(raw.size), C.c_float(tmin), C.c_float(tmax))
    return out


def bilateral(d, sigma_pixel, sigma_depth):
    out = np.empty_like(d)
    cpu().orc_bilateral(_fp(d), _fp(out), d.shape[1], d.shape[0], C.c_float(sigma_pixel), C.c_float(sigma_depth))
    return out


def vertices(d, cam: OrcCamera):
    out = np.empty(d.shape + (4,), np.float32)
    cpu().orc_vertices(_fp(d), _fp(out), C.byref(cam))
    return out


def normals(v):
    out = np.empty_like(v)
    cpu().orc_normals(_fp(v), _fp(out), v.shape[1], v.shape[0])
    return out


def down(v, is_normal: bool):
    out = np.empty((v.shape[0] >> 1, v.shape[1] >> 1, 4), np.float32)
    fn = cpu().orc_down_normals if is_normal else cpu().orc_down_vertices
    fn(_fp(v), _fp(out), v.shape[1], v.shape[0])
    return out


def integrate(vol, cfg, depth, pose, z0=0, z1=None) -> int:
    R = cfg.resolution
    z1 = R if z1 is None else z1
    cam = cam_of(cfg)
    p = np.ascontiguousarray(pose, dtype=np.float32).reshape(16)
    return int(cpu().orc_integrate(_fp(vol), R, C.c_float(cfg.volume_size), C.c_float(cfg.weight_max), _fp(depth), C.byref(cam),
                                   _fp(p), C.c_float(cfg.sdf_trunc), C.c_float(cfg.max_integrate_dist), z0, z1))


def integrate_color(vol, color, cfg, depth, normals, bgr, pose, angled) -> int:
    """has_color integration: vol f32 [R][R][R][2] and color u8 [R][R][R][4] are updated in place"""
    R = cfg.resolution
    cam = cam_of(cfg)
    p = np.ascontiguousarray(pose, dtype=np.float32).reshape(16)
    assert color.dtype == np.uint8 and color.flags["C_CONTIGUOUS"] and bgr.dtype == np.uint8 and bgr.flags["C_CONTIGUOUS"]
    lib = cpu()
    lib.orc_integrate_color.restype = C.c_uint64
    return int(lib.orc_integrate_color(_fp(vol), color.ctypes.data_as(C.c_void_p), R, C.c_float(cfg.volume_size), C.c_float(cfg.weight_max),
                                       _fp(depth), _fp(normals), bgr.ctypes.data_as(C.c_void_p), cfg.cols, cfg.rows, int(angled),
                                       C.byref(cam), _fp(p), C.c_float(cfg.sdf_trunc), C.c_float(cfg.max_integrate_dist), 0, R))


def raycast_color(vol, color, cfg, pose):
    cam = cam_of(cfg)
    v = np.empty((cfg.rows, cfg.cols, 4), np.float32)
    n = np.empty_like(v)
    rgb = np.empty((cfg.rows, cfg.cols, 3), np.uint8)
    p = np.ascontiguousarray(pose, np.float32).reshape(16)
    cpu().orc_raycast_color(_fp(vol), color.ctypes.data_as(C.c_void_p), cfg.resolution, C.c_float(cfg.volume_size), C.byref(cam), _fp(p),
                            C.c_float(cfg.ray_increment), C.c_float(cfg.trunc_min), C.c_float(cfg.trunc_max), _fp(v), _fp(n),
                            rgb.ctypes.data_as(C.c_void_p))
    return v, n, rgb


def marching_cubes_color(vol, color, cfg, max_tris=None):
    max_tris = max_tris or cfg.max_triangles
    tris = np.zeros((max_tris, 3, 6), np.float32)
    thr = np.float32(300) * np.float32(cfg.volume_size) / np.float32(cfg.resolution)
    lib = cpu()
    lib.orc_marching_cubes_color.restype = C.c_uint32
    n = int(lib.orc_marching_cubes_color(_fp(vol), color.ctypes.data_as(C.c_void_p), cfg.resolution, C.c_float(cfg.volume_size),
                                         C.c_float(float(thr)), _fp(tris), C.c_uint32(max_tris)))
    return n, tris[:min(n, max_tris)]


def icp_sums(nv, nn, mv, mn, cam: OrcCamera, cur, last_inv, dist_thres, sin_thres):
    out = np.zeros(27, np.float32)
    a = np.ascontiguousarray(cur, np.float32).reshape(16)
    b = np.ascontiguousarray(last_inv, np.float32).reshape(16)
    cpu().orc_icp_sums(_fp(nv), _fp(nn), _fp(mv), _fp(mn), C.byref(cam), _fp(a), _fp(b), C.c_float(dist_thres), C.c_float(sin_thres), _fp(out))
    return out


def raycast(vol, cfg, pose):
    cam = cam_of(cfg)
    v = np.empty((cfg.rows, cfg.cols, 4), np.float32)
    n = np.empty_like(v)
    p = np.ascontiguousarray(pose, np.float32).reshape(16)
    cpu().orc_raycast(_fp(vol), cfg.resolution, C.c_float(cfg.volume_size), C.byref(cam), _fp(p), C.c_float(cfg.ray_increment),
                      C.c_float(cfg.trunc_min), C.c_float(cfg.trunc_max), _fp(v), _fp(n))
    return v, n


def sdf_sums(vol, cfg, depth, cur):
    cam = cam_of(cfg)
    out = np.zeros(27, np.float32)
    a = np.ascontiguousarray(cur, np.float32).reshape(16)
    cpu().orc_sdf_sums(_fp(vol), cfg.resolution, C.c_float(cfg.volume_size), _fp(depth), C.byref(cam), _fp(a), _fp(out))
    return out


def marching_cubes(vol, cfg, max_tris=None):
    max_tris = max_tris or cfg.max_triangles
    tris = np.zeros((max_tris, 3, 6), np.float32)
    thr = np.float32(300) * np.float32(cfg.volume_size) / np.float32(cfg.resolution)
    n = int(cpu().orc_marching_cubes(_fp(vol), cfg.resolution, C.c_float(cfg.volume_size), C.c_float(float(thr)), _fp(tris), C.c_uint32(max_tris)))
    return n, tris[:min(n, max_tris)]


def solve_icp(s27):
    """det gate + LLT solve of the restated host part (oracle/host_math.h); returns (ok, x[6])"""
    a = np.ascontiguousarray(s27, np.float32).reshape(27)
    x = np.zeros(6, np.float32)
    ok = int(cpu().orc_solve_icp(_fp(a), _fp(x)))
    return ok, x


def mat44_inverse(m):
    a = np.ascontiguousarray(m, np.float32).reshape(16)
    out = np.empty(16, np.float32)
    cpu().orc_mat44_inverse(_fp(a), _fp(out))
    return out.reshape(4, 4)


def mat44_mul(a, b):
    x = np.ascontiguousarray(a, np.float32).reshape(16)
    y = np.ascontiguousarray(b, np.float32).reshape(16)
    out = np.empty(16, np.float32)
    cpu().orc_mat44_mul(_fp(x), _fp(y), _fp(out))
    return out.reshape(4, 4)


class CpuTracker:
    """oracle whole-frame pipeline (HybKinectfu::processNewFrame restated)"""

    def __init__(self, cfg):
        self.cfg = cfg
        self.lib = cpu()
        p = orc_params(cfg)
        self.h = self.lib.orc_tracker_create(C.byref(p))
        assert self.h

    def process(self, mm, frame_id, mode=0, given_pose=None):
        mm = np.ascontiguousarray(mm, np.uint16)
        out = np.empty(16, np.float32)
        gp = None
        if given_pose is not None:
            gp = np.ascontiguousarray(given_pose, np.float32).reshape(16)
        st = self.lib.orc_tracker_process(self.h, mm.ctypes.data_as(C.c_void_p), frame_id, mode,
                                          gp.ctypes.data_as(C.c_void_p) if gp is not None else None, _fp(out))
        return out.reshape(4, 4), int(st)

    def volume(self):
        R = self.cfg.resolution
        return np.ctypeslib.as_array(self.lib.orc_tracker_volume(self.h), shape=(R, R, R, 2))

    def map(self, which, level=0):
        rows, cols = self.cfg.level_shape(level)
        return np.ctypeslib.as_array(self.lib.orc_tracker_map(self.h, which, level), shape=(rows, cols, 4))

    def depth(self, which):
        return np.ctypeslib.as_array(self.lib.orc_tracker_depth(self.h, which), shape=(self.cfg.rows, self.cfg.cols))

    def times(self):
        t = (C.c_double * 4)()
        self.lib.orc_tracker_times(self.h, t)
        return list(t)

    def close(self):
        if self.h:
            self.lib.orc_tracker_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


# ---------------------------------------------------------------- reference CUDA build (GPU box only)
REF_RAW_DEPTH, REF_TRUNC_DEPTH, REF_FILTERED_DEPTH, REF_NEW_VERTICES, REF_NEW_NORMALS, REF_MODEL_VERTICES, \
    REF_MODEL_NORMALS, REF_VOLUME, REF_ICP_REDUCED, REF_TRIANGLES, REF_VOLUME_COLOR, REF_RAYCAST_RGB, REF_RAW_RGB = range(13)


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


class Ref:
    """drives the unmodified reference kernels through oracle/ref_harness.cu"""
    _lib = None

    def __init__(self, cfg):
        if Ref._lib is None:
            lib = C.CDLL(REF_LIB)
            lib.ref_download.restype = C.c_long
            lib.ref_download.argtypes = [C.c_int, C.c_uint, C.c_void_p, C.c_size_t]
            lib.ref_upload.restype = C.c_long
            lib.ref_upload.argtypes = [C.c_int, C.c_uint, C.c_void_p, C.c_size_t]
            lib.ref_marchingcube.restype = C.c_uint
            lib.ref_process_frame.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_void_p, F32P]
            lib.ref_download_volume_planes.restype = C.c_long
            lib.ref_download_volume_planes.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_size_t]
            Ref._lib = lib
        self.lib = Ref._lib
        self.cfg = cfg
        p = ref_params(cfg)
        rc = self.lib.ref_init(C.byref(p))
        if rc != 0:
            raise RuntimeError(f"ref_init failed: {rc}")
        self.lib.ref_set_color(0, 0)

    def _shape(self, buf, level):
        rows, cols = self.cfg.level_shape(level)
        if buf in (REF_RAW_DEPTH, REF_TRUNC_DEPTH, REF_FILTERED_DEPTH):
            return (self.cfg.rows, self.cfg.cols), np.float32
        if buf in (REF_NEW_VERTICES, REF_NEW_NORMALS, REF_MODEL_VERTICES, REF_MODEL_NORMALS):
            return (rows, cols, 4), np.float32
        if buf == REF_VOLUME:
            R = self.cfg.resolution
            return (R, R, R, 2), np.float32
        if buf == REF_ICP_REDUCED:
            return (27,), np.float32
        if buf == REF_VOLUME_COLOR:
            R = self.cfg.resolution
            return (R, R, R, 4), np.uint8
        if buf in (REF_RAYCAST_RGB, REF_RAW_RGB):
            return (self.cfg.rows, self.cfg.cols, 3), np.uint8
        raise ValueError(buf)

    def download(self, buf, level=0):
        if buf == REF_TRIANGLES:
            out = np.empty((self.cfg.max_triangles, 3, 6), np.float32)
            got = self.lib.ref_download(buf, 0, out.ctypes.data, out.nbytes)
            assert got >= 0
            return out[: got // 72].copy()
        shape, dt = self._shape(buf, level)
        out = np.empty(shape, dt)
        got = self.lib.ref_download(buf, level, out.ctypes.data, out.nbytes)
        assert got == out.nbytes, (buf, got, out.nbytes)
        return out

    def download_volume_planes(self, z0, z1):
        """planes [z0, z1) as packed {tsdf, weight} (for volumes too large to download whole)"""
        R = self.cfg.resolution
        out = np.empty((z1 - z0, R, R, 2), np.float32)
        got = self.lib.ref_download_volume_planes(z0, z1, out.ctypes.data, out.nbytes)
        assert got == out.nbytes, (got, out.nbytes)
        return out

    def upload(self, buf, arr, level=0):
        a = np.ascontiguousarray(arr)
        got = self.lib.ref_upload(buf, level, a.ctypes.data, a.nbytes)
        assert got == a.nbytes, (buf, got, a.nbytes)

    def upload_depth_mm(self, mm):
        a = np.ascontiguousarray(mm, np.uint16)
        self.lib.ref_upload_depth_mm(a.ctypes.data_as(C.c_void_p))

    def set_color(self, use_color, angle_weight=0):
        self.lib.ref_set_color(int(use_color), int(angle_weight))

    def upload_rgb(self, bgr):
        a = np.ascontiguousarray(bgr, np.uint8)
        assert a.shape == (self.cfg.rows, self.cfg.cols, 3)
        self.lib.ref_upload_rgb(a.ctypes.data_as(C.c_void_p))

    def preprocess(self, mm):
        self.upload_depth_mm(mm)
        self.lib.ref_trunc(); self.lib.ref_bilateral(); self.lib.ref_vertices(); self.lib.ref_normals()
        self.lib.ref_downsample_new()

    def reset_volume(self):
        self.lib.ref_reset_volume()

    def integrate(self, pose):
        p = np.ascontiguousarray(pose, np.float32).reshape(16)
        self.lib.ref_integrate(_fp(p))

    def raycast(self, pose):
        p = np.ascontiguousarray(pose, np.float32).reshape(16)
        self.lib.ref_raycast(_fp(p))

    def downsample_model(self):
        self.lib.ref_downsample_model()

    def icp_params(self, level, cur, last_inv):
        a = np.ascontiguousarray(cur, np.float32).reshape(16)
        b = np.ascontiguousarray(last_inv, np.float32).reshape(16)
        out = np.empty(27, np.float32)
        self.lib.ref_icp_params(C.c_uint(level), _fp(a), _fp(b), _fp(out))
        return out

    def sdf_params(self, cur):
        a = np.ascontiguousarray(cur, np.float32).reshape(16)
        out = np.empty(27, np.float32)
        self.lib.ref_sdf_params(_fp(a), _fp(out))
        return out

    def marching_cubes(self) -> int:
        return int(self.lib.ref_marchingcube())

    def process(self, mm, frame_id, mode=0, given_pose=None):
        mm = np.ascontiguousarray(mm, np.uint16)
        out = np.empty(16, np.float32)
        gp = None
        if given_pose is not None:
            gp = np.ascontiguousarray(given_pose, np.float32).reshape(16)
        st = self.lib.ref_process_frame(mm.ctypes.data_as(C.c_void_p), frame_id, mode,
                                        gp.ctypes.data_as(C.c_void_p) if gp is not None else None, _fp(out))
        return out.reshape(4, 4), int(st)

    def check(self):
        e = self.lib.ref_last_cuda_error()
        assert e == 0, f"reference CUDA error {e}"
