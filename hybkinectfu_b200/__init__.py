"""hybkinectfu_b200 — B200-native per-frame hot path of hybKinectfu behind the reference's API.

csrc/   hand-written sm_100a CUDA kernels + the C ABI declared in include/hkf.h  (libhkf_b200.so)
host/   C++ mirror of the reference's host classes (HybKinectfu, CameraPoseFinder*, ...) on top of that ABI
capi.py ctypes plumbing used by tests/ and bench.py
synth.py synthetic analytic scene + trajectory
"""
from . import capi, synth  # noqa: F401
from .capi import Config, Context, HkfError  # noqa: F401
