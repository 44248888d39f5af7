#!/usr/bin/env python
"""Development report (GPU box): stage-by-stage mismatch counts of the CUDA path vs the unmodified
reference kernels and vs the CPU oracle, plus coarse timings.  Not a test: it never asserts, it prints."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as O  # noqa: E402
from hybkinectfu_b200 import capi, synth  # noqa: E402
from hybkinectfu_b200.capi import Config  # noqa: E402


def cmp(name, a, b, tol=0.0):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"  {name:28s} SHAPE {a.shape} vs {b.shape}")
        return
    bits = (a.view(np.uint32) != b.view(np.uint32)) if a.dtype == np.float32 else (a != b)
    # +0 / -0 are different bits but equal values
    nb = int(bits.sum())
    with np.errstate(invalid="ignore"):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    d = np.where(np.isnan(d), 0, d)
    print(f"  {name:28s} bit-mismatch {nb:9d} / {a.size:9d}   max|d| {d.max() if d.size else 0:.3e}   zero-pattern-mismatch {int(((a == 0) != (b == 0)).sum())}")


def run(cfg, cam, frames, poses, label, do_ref=True, do_cpu=True):
    print(f"== {label}: {cfg.cols}x{cfg.rows}, R={cfg.resolution}")
    ctx = capi.Context(cfg)
    ref = O.Ref(cfg) if (do_ref and O.ref_available()) else None
    # --- per-wrapper path
    ctx.upload_depth_mm(frames[0]); ctx.trunc_depth(); ctx.bilateral(); ctx.vertices(); ctx.normals(); ctx.downsample_new()
    ctx.sync()
    mine = {"raw": ctx.download(capi.BUF_RAW_DEPTH), "trunc": ctx.download(capi.BUF_TRUNC_DEPTH), "filt": ctx.download(capi.BUF_FILTERED_DEPTH)}
    for l in range(cfg.pyramid_levels):
        mine[f"v{l}"] = ctx.download(capi.BUF_NEW_VERTICES, l); mine[f"n{l}"] = ctx.download(capi.BUF_NEW_NORMALS, l)
    # --- fused path
    ctx.preprocess(frames[0])
    fused = {"raw": ctx.download(capi.BUF_RAW_DEPTH), "trunc": ctx.download(capi.BUF_TRUNC_DEPTH), "filt": ctx.download(capi.BUF_FILTERED_DEPTH)}
    for l in range(cfg.pyramid_levels):
        fused[f"v{l}"] = ctx.download(capi.BUF_NEW_VERTICES, l); fused[f"n{l}"] = ctx.download(capi.BUF_NEW_NORMALS, l)
    print(" fused vs per-wrapper")
    for k in mine:
        cmp(k, fused[k], mine[k])
    if ref:
        ref.preprocess(frames[0])
        r = {"raw": ref.download(O.REF_RAW_DEPTH), "trunc": ref.download(O.REF_TRUNC_DEPTH), "filt": ref.download(O.REF_FILTERED_DEPTH)}
        for l in range(cfg.pyramid_levels):
            r[f"v{l}"] = ref.download(O.REF_NEW_VERTICES, l); r[f"n{l}"] = ref.download(O.REF_NEW_NORMALS, l)
        print(" CUDA path vs reference kernels")
        for k in mine:
            cmp(k, mine[k], r[k])
    if do_cpu:
        raw = O.decode_mm(frames[0]); tr = O.trunc(raw, cfg.trunc_min, cfg.trunc_max); fl = O.bilateral(tr, cfg.sigma_pixel, cfg.sigma_depth)
        v0 = O.vertices(fl, O.cam_of(cfg)); n0 = O.normals(v0)
        print(" CPU oracle vs", "reference kernels" if ref else "CUDA path")
        base = r if ref else mine
        cmp("raw", raw, base["raw"]); cmp("trunc", tr, base["trunc"]); cmp("filt", fl, base["filt"])
        cmp("v0", v0, base["v0"]); cmp("n0", n0, base["n0"])
        # feed the oracle the GPU's filtered map to isolate the __expf difference
        v0b = O.vertices(base["filt"], O.cam_of(cfg)); n0b = O.normals(v0b)
        cmp("v0 (from GPU filt)", v0b, base["v0"]); cmp("n0 (from GPU filt)", n0b, base["n0"])
        v1 = O.down(base["v0"], False); n1 = O.down(base["n0"], True)
        cmp("v1 (from GPU v0)", v1, base["v1"]); cmp("n1 (from GPU n0)", n1, base["n1"])
    # --- integration
    nint = min(3, len(frames))
    ctx.reset_volume()
    if ref:
        ref.reset_volume()
    cpu_vol = np.zeros((cfg.resolution,) * 3 + (2,), np.float32) if (do_cpu and cfg.resolution <= 256) else None
    for i in range(nint):
        ctx.preprocess(frames[i])
        nupd = ctx.count_updates(poses[i])
        ctx.integrate(poses[i])
        ctx.sync()
        if ref:
            ref.preprocess(frames[i]); ref.integrate(poses[i])
        if cpu_vol is not None:
            t = O.trunc(O.decode_mm(frames[i]), cfg.trunc_min, cfg.trunc_max)
            t0 = time.time(); ncpu = O.integrate(cpu_vol, cfg, t, poses[i]); dt = time.time() - t0
            print(f"  frame {i}: N_upd CUDA {nupd}  CPU {ncpu}  (cpu {dt*1e3:.1f} ms, {O.cpu().orc_num_threads()} threads)")
        else:
            print(f"  frame {i}: N_upd CUDA {nupd} ({100.0*nupd/cfg.resolution**3:.1f}% of R^3)")
    vol = ctx.download(capi.BUF_VOLUME)
    if ref:
        rv = ref.download(O.REF_VOLUME)
        print(" volume: CUDA path vs reference kernels")
        cmp("tsdf", vol[..., 0], rv[..., 0]); cmp("weight", vol[..., 1], rv[..., 1])
        ref.check()
    if cpu_vol is not None:
        print(" volume: CPU oracle vs", "reference kernels" if ref else "CUDA path")
        b = rv if ref else vol
        cmp("tsdf", cpu_vol[..., 0], b[..., 0]); cmp("weight", cpu_vol[..., 1], b[..., 1])
    # --- timing of integration (coarse, wall clock around a batch; CUDA-event timing lives in bench.py)
    ctx.preprocess(frames[nint - 1])
    nupd = ctx.count_updates(poses[nint - 1])
    for _ in range(3):
        ctx.integrate(poses[nint - 1])
    ctx.sync()
    reps = 20
    t0 = time.time()
    for _ in range(reps):
        ctx.integrate(poses[nint - 1])
    ctx.sync()
    dt = (time.time() - t0) / reps
    bytes_alg = 16 * nupd + 4 * cfg.cols * cfg.rows
    print(f" integrate: {dt*1e6:.1f} us/frame, N_upd {nupd}, algorithmic {bytes_alg/1e6:.1f} MB -> {bytes_alg/dt/1e9:.0f} GB/s; full-sweep equiv {16*cfg.resolution**3/dt/1e9:.0f} GB/s")
    t0 = time.time()
    for _ in range(reps):
        ctx.preprocess_device(0) if False else None
    print(f" launches so far: {ctx.launches()}")
    ctx.close()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "small"):
        cam = synth.Camera().scaled(192, 144)
        cfg = Config(cols=192, rows=144, cx=cam.cx, cy=cam.cy, fx=cam.fx, fy=cam.fy, resolution=64, pyramid_levels=3, max_triangles=400000)
        poses = synth.trajectory(200)[::4][:4]
        frames = np.stack([synth.render_depth_mm(p, cam) for p in poses])
        run(cfg, cam, frames, poses.astype(np.float32), "small")
    if which in ("all", "c1"):
        cam = synth.Camera()
        cfg = Config(resolution=256, pyramid_levels=3)
        poses = synth.trajectory(100)[:3]
        frames = np.stack([synth.render_depth_mm(p, cam) for p in poses])
        run(cfg, cam, frames, poses.astype(np.float32), "C1 640x480 / 256^3")
    if which in ("all", "c2"):
        cam = synth.Camera()
        cfg = Config(resolution=512, pyramid_levels=3)
        poses = synth.trajectory(500)[:3]
        frames = np.stack([synth.render_depth_mm(p, cam) for p in poses])
        run(cfg, cam, frames, poses.astype(np.float32), "C2 640x480 / 512^3", do_cpu=False)


if __name__ == "__main__":
    main()
