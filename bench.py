#!/usr/bin/env python
"""bench.py — fused frames/s of the hybKinectfu hot path on B200 (BASELINE.json metric).

A step = one depth frame through preprocess -> 3-level projective ICP -> TSDF integrate -> raycast.
Workload at N=1 = BASELINE.json configs[1]: 640x480 synthetic sequence into a 512^3 / 2 m volume.
N>1 (torchrun, one rank per GPU): N independent replicas of that workload with phase-shifted
trajectories ("scaling": "weak", aggregate frames/s; the path shards by sequence with no data-path
collective).  `--workload c4` instead runs BASELINE configs[3]: ONE sequence into a 1024^3 volume that is
z-slab-sharded over the N GPUs ("scaling": "strong"), with nearest-hit raycast compositing and the NCCL
allreduce of the ICP sums.

  value      frames/s with every input frame already resident in HBM, CUDA-event timed on the
             context's stream
  e2e        the same through the reference-facing call sequence with HOST frames: per frame a 614 KB
             H2D of the 16-bit frame from pinned memory and a D2H of pose + tracking status
  roofline   TSDF integration kernel: (16 B x N_upd + 4 B x W*H) / mean CUDA-event duration of that
             kernel inside the timed region, against the measured HBM copy bandwidth
  cpu_baseline  the C oracle port of the same pipeline on the host cores, bounded sample (N=1, rank 0)

--impl reference runs the reference arm: the reference's own unmodified CUDA kernels through its own
wrappers (oracle/_ref/libhkf_ref.so, built from /root/reference by oracle/Makefile) with the restated
host solves, or the C oracle port when that library is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fused frames/s at 640x480 into 512^3 TSDF (ICP + integrate + raycast)"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every 10 ms from a thread (the timed
    region of the default run is ~0.1 s, too short for `nvidia-smi -lms`, which is the fallback when NVML is missing)"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.samples = []          # (sm_mhz, sm_max_mhz, reasons bitmask)
        self.proc = None
        self.t = None
        self.stop_flag = False
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def sample_now(self):
        """one NVML sample from the calling thread (used while the GPU is still working through the timed region, so
        that there is at least one in-region sample even if the polling thread was starved of the GIL)"""
        n = self.nvml
        if n is None:
            return
        try:
            mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
            try:
                r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            self.samples.append((mhz, self.max_mhz, r))
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag:
            self.sample_now()
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 7:
                continue
            try:
                bits = 0
                for bit, val in zip((0x8, 0x40, 0x20, 0x4), f[3:7]):     # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
                    if val.lower().startswith("active"):
                        bits |= bit
                self.samples.append((float(f[0]), float(f[1]), bits))
            except ValueError:
                continue

    def stop(self):
        self.stop_flag = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        if self.t:
            self.t.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"], "samples": 0}
        sm = [s[0] for s in self.samples]
        bits = 0
        for s in self.samples:
            bits |= s[2]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(s[1] for s in self.samples),
                "reasons": sorted(v for k, v in names.items() if bits & k), "samples": len(sm),
                "source": "NVML, 10 ms period" if self.nvml else "nvidia-smi -lms 20"}


def make_sequence(n, phase=0.0, cols=640, rows=480):
    """n synthetic u16 frames (+ ground-truth poses) of the periodic 500-frame trajectory; cached in /tmp"""
    from hybkinectfu_b200 import synth
    cache = f"/tmp/hkf_seq_{cols}x{rows}_{n}_{phase:.4f}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        return z["frames"], z["poses"]
    cam = synth.Camera().scaled(cols, rows) if (cols, rows) != (640, 480) else synth.Camera()
    traj = synth.trajectory(500, phase)
    poses = np.stack([traj[i % 500] for i in range(n)])
    frames = np.stack([synth.render_depth_mm(p, cam) for p in poses])
    try:
        np.savez(cache, frames=frames, poses=poses)
    except Exception:
        pass
    return frames, poses


def bench_cfg(resolution=512):
    from hybkinectfu_b200.capi import Config
    return Config(resolution=resolution, pyramid_levels=3, max_triangles=1)


def ate_rmse(est, gt):
    d = est[:, :3, 3] - gt[:, :3, 3]
    return float(np.sqrt((d * d).sum(1).mean()))


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    from hybkinectfu_b200 import capi, synth
    torch.cuda.set_device(local_rank)
    cfg = bench_cfg(512)
    W, K = args.warmup, args.steps
    n = W + K
    frames, gt = make_sequence(n, phase=2 * np.pi * rank / max(world, 1) if world > 1 else 0.0)
    ctx = capi.Context(cfg, device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.lib.hkf_stream(ctx._h), device=local_rank)
    p0 = synth.pose0().astype(np.float32)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def frame(i, src_ptr, is_dev):
        ctx._ck(ctx.lib.hkf_preprocess(ctx._h, src_ptr, is_dev, ctx.fp), "hkf_preprocess")
        if i != 0:
            ctx.icp_track()
        ctx.integrate_tracked()
        ctx.raycast_tracked()

    # ---------------- value: inputs resident in HBM
    dev_frames = torch.from_numpy(frames.view(np.int16)).to(f"cuda:{local_rank}")
    fbytes = frames[0].nbytes
    ctx.reset_volume(); ctx.set_pose(p0)
    for i in range(W):
        frame(i, dev_frames.data_ptr() + i * fbytes, 1)
    ctx.sync()
    barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local_rank); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ie = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = ctx.launches()
    ev0.record(stream)
    for k in range(K):
        i = W + k
        # frames are already resident in HBM: nothing to overlap, the plain per-frame call sequence
        ctx._ck(ctx.lib.hkf_preprocess(ctx._h, dev_frames.data_ptr() + i * fbytes, 1, ctx.fp), "hkf_preprocess")
        ctx.icp_track()
        ie[k][0].record(stream)
        ctx.integrate_tracked()
        ie[k][1].record(stream)
        ctx.raycast_tracked()
        if k % 50 == 25:
            sampler.sample_now()             # the host runs ahead of the GPU: these fall inside the timed region
    ev1.record(stream)
    sampler.sample_now()
    ctx.sync(); torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches() - l0
    ms_total = ev0.elapsed_time(ev1)
    int_ms = np.array([a.elapsed_time(b) for a, b in ie])
    pose_end, st_end = ctx.frame_result()

    # ---------------- N_upd of the timed frames (untimed re-run, deterministic => identical poses)
    ctx.reset_volume(); ctx.set_pose(p0)
    n_upd, est = [], []
    for i in range(n):
        ctx._ck(ctx.lib.hkf_preprocess(ctx._h, dev_frames.data_ptr() + i * fbytes, 1, ctx.fp), "hkf_preprocess")
        if i != 0:
            ctx.icp_track()
        p, st = ctx.frame_result()
        est.append(p)
        if i >= W:
            n_upd.append(ctx.count_updates(p))
        ctx.integrate_tracked(); ctx.raycast_tracked()
    ctx.sync()
    est = np.stack(est)
    n_upd = np.array(n_upd, dtype=np.float64)
    alg_bytes = 16.0 * n_upd + 4.0 * cfg.cols * cfg.rows
    achieved = float((alg_bytes / (int_ms * 1e-3)).mean() / 1e9)

    # ---------------- e2e: host frames, H2D + pose D2H every step, per-frame sync (the API a user calls).
    # The loop synchronises with the host every frame, so it is sensitive to host scheduling noise on a shared box:
    # it is run twice and the faster run is reported (both are listed in e2e.runs_fps).
    host_frames = torch.from_numpy(frames.view(np.int16)).pin_memory()
    e2e_runs = []
    for rep in range(2):
        ctx.reset_volume(); ctx.set_pose(p0)

        def e2e_frame(i, first, has_next):
            if first:
                ctx._ck(ctx.lib.hkf_preprocess(ctx._h, host_frames.data_ptr() + i * fbytes, 0, ctx.fp), "hkf_preprocess")
            else:
                ctx.swap_frame()
            if has_next:
                ctx.preprocess_next(host_frames.data_ptr() + (i + 1) * fbytes, False)     # the next frame's H2D overlaps this frame
            if i:
                ctx.icp_track()
            ctx.integrate_tracked()
            ctx.raycast_tracked()
            return ctx.frame_result()                                                    # D2H pose + status, sync

        for i in range(W):                       # warm-up with the very same call sequence (frame W is prepared by frame W-1)
            e2e_frame(i, i == 0, True)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(K):
            e2e_frame(W + k, False, k + 1 < K)
        torch.cuda.synchronize()
        e2e_runs.append(time.perf_counter() - t0)
        barrier()
    e2e_s = min(e2e_runs)

    # ---------------- aggregate over ranks (max time)
    t_val, t_e2e = ms_total * 1e-3, e2e_s
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([t_val, t_e2e], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_val, t_e2e = float(t[0]), float(t[1])
        lt = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(lt)
        launches = int(lt[0])
    peak, peak_src = measured_hbm_peak()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "integrate_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    out = {
        "metric": METRIC, "value": world * K / t_val, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_val * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic analytic scene (floor, two walls, two spheres), 16-bit mm depth, periodic 500-frame trajectory",
        "config": {"workload": "C2 full pipeline: 640x480, 3-level pyramid projective ICP {10,5,4} + TSDF integrate + raycast, 512^3 volume over 2 m"
                   + ("" if world == 1 else f", {world} independent replicas (one per GPU, phase-shifted trajectories)"),
                   "volume": "512^3 packed {f32 tsdf,f32 weight} = 1.07 GB", "parallelism": "1 GPU" if world == 1 else f"replicas x{world}",
                   "l2": "per-frame working set (1.07 GB volume) exceeds the 126 MB L2; no explicit flush",
                   "track_status_last": int(st_end), "ate_rmse_m": ate_rmse(est[W:], gt[W:])},
        "e2e": {"value": world * K / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(fbytes), "d2h_bytes_per_step": 264,
                "runs_fps": [K / x for x in e2e_runs],
                "note": "per frame: hkf_swap_frame (preprocess) + hkf_preprocess_next(host u16 frame i+1: its H2D overlaps this frame) + icp_track + integrate + raycast + hkf_frame_result (pose D2H, sync); best of two runs"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_integrate_items + k_integrate (TSDF integration, both launches timed together)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": float(alg_bytes.mean()), "n_upd_mean": float(n_upd.mean()),
                     "kernel_ms_mean": float(int_ms.mean()), "kernel_share_of_step": float(int_ms.sum() / ms_total),
                     "full_sweep_equiv_GBps": float(16.0 * cfg.resolution ** 3 / (int_ms.mean() * 1e-3) / 1e9)},
    }
    ctx.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(cfg, frames, args.cpu_frames)
    return out


def cpu_baseline(cfg, frames, nframes):
    """the C oracle port (OpenMP) of the same pipeline on the host cores, bounded sample"""
    import _oracle as O
    tr = O.CpuTracker(cfg)
    tr.process(frames[0], 0, 0)                       # frame 0: integrate + raycast only (warm-up, untimed)
    t0 = time.perf_counter()
    for i in range(1, nframes + 1):
        _, st = tr.process(frames[i], i, 0)
    dt = time.perf_counter() - t0
    times = tr.times()
    tr.close()
    return {"value": nframes / dt, "unit": "frames/s", "cores": int(O.cpu().orc_num_threads()), "kind": "port",
            "sample": f"frames 1..{nframes} of the same 640x480 / 512^3 sequence (frame 0 untimed), gcc -O2 -fopenmp C port of the reference loops",
            "stage_seconds": {"preprocess": times[0], "track": times[1], "integrate": times[2], "raycast": times[3]}}


# ---------------------------------------------------------------------------------------------- C1 / C3 / C5
def run_config(args, rank, world, local_rank):
    """The other BASELINE.json configs on one GPU (per rank), device-resident frames, CUDA-event timed:
    c1  integration-only, 100 frames 640x480, poses from file (ground truth), 256^3
    c3  SDF direct tracking (CameraPoseFinderSDF) + integrate + raycast on 512^3, then marching-cubes export
    c5  1280x960 sequences into 512^3, ICP pipeline, one independent sequence per GPU"""
    import dataclasses
    import torch
    from hybkinectfu_b200 import capi, synth
    torch.cuda.set_device(local_rank)
    wl = args.workload
    W, K = args.warmup, args.steps
    if wl == "c1":
        cfg = dataclasses.replace(bench_cfg(256), max_triangles=1)
        W, K = 0, min(K, 100) if K != 200 else 100
        frames, gt = make_sequence(100)
    elif wl == "c3":
        cfg = dataclasses.replace(bench_cfg(512), max_triangles=4000000)
        frames, gt = make_sequence(W + K)
    else:
        cam = synth.Camera().scaled(1280, 960)
        cfg = dataclasses.replace(bench_cfg(512), cols=1280, rows=960, cx=cam.cx, cy=cam.cy, fx=cam.fx, fy=cam.fy, max_triangles=1)
        frames, gt = make_sequence(W + K, phase=2 * np.pi * rank / 8.0, cols=1280, rows=960)
    ctx = capi.Context(cfg, device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.lib.hkf_stream(ctx._h), device=local_rank)
    dev = torch.from_numpy(frames.view(np.int16)).to(f"cuda:{local_rank}")
    fb = frames[0].nbytes
    ctx.set_pose(synth.pose0().astype(np.float32))
    gt32 = gt.astype(np.float32)

    def frame(i):
        ctx._ck(ctx.lib.hkf_preprocess(ctx._h, dev.data_ptr() + i * fb, 1, ctx.fp), "hkf_preprocess")
        if wl == "c1":
            ctx.set_pose(gt32[i])
        elif i:
            ctx.sdf_track() if wl == "c3" else ctx.icp_track()
        ctx.integrate_tracked()
        if wl != "c1":
            ctx.raycast_tracked()

    for i in range(W):
        frame(i)
    ctx.sync(); torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches()
    n_upd = 0
    if wl == "c1":
        # hkf_set_pose synchronises (host-provided poses): time the integration launches themselves
        tot = 0.0
        for i in range(K):
            ctx._ck(ctx.lib.hkf_preprocess(ctx._h, dev.data_ptr() + i * fb, 1, ctx.fp), "hkf_preprocess")
            ctx.set_pose(gt32[i])
            n_upd += ctx.count_updates(gt32[i])
            ev0.record(stream); ctx.integrate_tracked(); ev1.record(stream); ctx.sync()
            tot += ev0.elapsed_time(ev1) * 1e-3
        t = tot
    else:
        ev0.record(stream)
        for k in range(K):
            frame(W + k)
        ev1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        t = ev0.elapsed_time(ev1) * 1e-3
    launches = ctx.launches() - l0
    pose, st = ctx.frame_result()
    extra = {"track_status_last": int(st), "pose_err_vs_gt_m": float(np.abs(pose[:3, 3] - gt[(W + K - 1) % len(gt)][:3, 3]).max())}
    if wl == "c1":
        peak, _ = measured_hbm_peak()
        gbs = (16.0 * n_upd + 4.0 * cfg.cols * cfg.rows * K) / t / 1e9
        extra.update({"integration_GBps_algorithmic": gbs, "frac_of_hbm_peak": gbs / peak, "n_upd_mean": n_upd / K})
    if wl == "c3":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.marching_cubes()                                  # first call uploads the case table (one-off)
        e0.record(stream); ntri = ctx.marching_cubes(); e1.record(stream); ctx.sync()
        extra.update({"marching_cubes_triangles": int(ntri), "marching_cubes_ms": e0.elapsed_time(e1)})
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([t], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
    names = {"c1": "C1 integration-only: 100 frames 640x480, poses from file, 256^3 (metric = integrations/s)",
             "c3": "C3 SDF direct tracking (<= 6 Gauss-Newton iterations) + integrate + raycast, 512^3, marching-cubes export at the end",
             "c5": f"C5 1280x960 sequences into 512^3, ICP pipeline, {world} independent replica(s)"}
    out = {"metric": "fused frames/s" if wl != "c1" else "TSDF integrations/s", "value": world * K / t, "unit": "frames/s", "n_gpus": world, "steps": K,
           "warmup": W, "ms_per_step": t * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic analytic scene, 16-bit mm depth", "config": dict({"workload": names[wl]}, **extra), "gpu_launches": int(launches)}
    ctx.close()
    return out


# ---------------------------------------------------------------------------------------------- C4: z-slab sharded
def run_c4(args, rank, world, local_rank):
    """BASELINE.json configs[3]: ONE 640x480 sequence into a 1024^3 volume, z-slab-sharded over `world` GPUs
    (strong scaling): banded ICP + NCCL allreduce of the 27 sums, slab-local integration (halo planes integrated
    redundantly, no volume exchange), nearest-hit compositing of the per-slab raycasts."""
    import dataclasses
    import torch
    from hybkinectfu_b200 import capi, shard, synth
    torch.cuda.set_device(local_rank)
    cfg = dataclasses.replace(bench_cfg(args.c4_resolution), max_triangles=1)
    W, K = args.warmup, args.steps
    frames, gt = make_sequence(W + K)
    if world > 1:
        ctx = shard.sharded_context(cfg, rank, world, local_rank)
    else:
        ctx = capi.Context(cfg, device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.lib.hkf_stream(ctx._h), device=local_rank)
    dev_frames = torch.from_numpy(frames.view(np.int16)).to(f"cuda:{local_rank}")
    fbytes = frames[0].nbytes
    ctx.set_pose(synth.pose0().astype(np.float32))

    def frame(i):
        ctx._ck(ctx.lib.hkf_preprocess(ctx._h, dev_frames.data_ptr() + i * fbytes, 1, ctx.fp), "hkf_preprocess")
        if i:
            ctx.icp_track()
        ctx.integrate_tracked()
        ctx.raycast_tracked()
        if world > 1 and os.environ.get("HKF_SHARD_MANUAL_COMPOSITE") == "1":
            ctx._ck(ctx.lib.hkf_shard_composite_raycast(ctx._h), "composite")

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    for i in range(W):
        frame(i)
    ctx.sync(); barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches()
    ev0.record(stream)
    est = []
    for k in range(K):
        frame(W + k)
    ev1.record(stream)
    ctx.sync(); torch.cuda.synchronize(); barrier()
    t = ev0.elapsed_time(ev1) * 1e-3
    launches = ctx.launches() - l0
    pose, st = ctx.frame_result()
    # per-stage CUDA-event times of this rank (a few extra frames of the periodic sequence, untimed above)
    stage = np.zeros(4)
    comp_us = 0.0
    manual_composite = world > 1 and os.environ.get("HKF_SHARD_MANUAL_COMPOSITE") == "1"
    nst = min(10, W + K - 1)
    for j in range(nst):
        i = 1 + j
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record(stream)
        ctx._ck(ctx.lib.hkf_preprocess(ctx._h, dev_frames.data_ptr() + i * fbytes, 1, ctx.fp), "hkf_preprocess")
        ev[1].record(stream)
        ctx.icp_track(); ev[2].record(stream)
        ctx.integrate_tracked(); ev[3].record(stream)
        ctx.raycast_tracked()
        if manual_composite:
            evc = torch.cuda.Event(enable_timing=True); evc.record(stream)
            ctx._ck(ctx.lib.hkf_shard_composite_raycast(ctx._h), "composite")
        ev[4].record(stream)
        ctx.sync()
        stage += np.array([ev[a].elapsed_time(ev[a + 1]) * 1e3 for a in range(4)])
        if manual_composite:
            comp_us += evc.elapsed_time(ev[4]) * 1e3
    stage /= max(nst, 1)
    all_stage = [stage.tolist()]
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([t], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
        g = [torch.zeros(4, dtype=torch.float64, device=f"cuda:{local_rank}") for _ in range(world)]
        dist.all_gather(g, torch.tensor(stage, dtype=torch.float64, device=f"cuda:{local_rank}"))
        all_stage = [x.cpu().tolist() for x in g]
    halo = shard.halo_planes(cfg.ray_increment, cfg.volume_size, cfg.resolution)
    slabs = shard.balanced_slab_ranges(cfg, world, halo) if world > 1 else shard.slab_ranges(cfg.resolution, 1, 0)
    out = {"metric": f"fused frames/s at 640x480 into {cfg.resolution}^3 TSDF, z-slab-sharded (ICP + integrate + raycast)", "value": K / t,
           "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t * 1e3 / K, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic analytic scene, 16-bit mm depth",
           "config": {"workload": f"C4: one 640x480 sequence, {cfg.resolution}^3 volume over 2 m, z-slabs over {world} GPU(s), halo {halo if world > 1 else 0} planes",
                      "owned_planes_per_rank": [s[1] - s[0] for s in slabs], "stored_planes_per_rank": [s[3] - s[2] for s in slabs], "track_status_last": int(st),
                      "rank0_stage_us": {"preprocess": float(stage[0]), "icp": float(stage[1]), "integrate": float(stage[2]), "raycast+composite": float(stage[3])},
                      "integrate_us_per_rank": [round(x[2], 1) for x in all_stage], "rank0_composite_us": comp_us / max(nst, 1) if manual_composite else None,
                      "note": "per-stage times include waiting for the slowest rank at the next exchange (ICP sums, raycast compositing)",
                      "pose_err_vs_gt_m": float(np.abs(pose[:3, 3] - gt[W + K - 1][:3, 3]).max())},
           "gpu_launches": int(launches)}
    ctx.close()
    return out


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return None
    import _oracle as O
    cfg = bench_cfg(512)
    cfg.max_triangles = 1
    W, K = args.warmup, args.steps
    if O.ref_available():
        try:
            import torch
            has_gpu = torch.cuda.is_available()
        except Exception:
            has_gpu = False
    else:
        has_gpu = False
    if O.ref_available() and has_gpu:
        frames, gt = make_sequence(W + K)
        ref = O.Ref(cfg)
        ref.reset_volume()
        for i in range(W):
            ref.process(frames[i], i, 0)
        ref.check()
        t0 = time.perf_counter()
        for k in range(K):
            _, st = ref.process(frames[W + k], W + k, 0)
        ref.check()
        dt = time.perf_counter() - t0
        fps = K / dt
        kind, cores = "reference", 1
        sample = (f"{K} frames after {W} warm-up: the reference's unmodified CUDA kernels (oracle/_ref, nvcc defaults, sm_100) through its own "
                  "wrappers incl. their cudaDeviceSynchronize and per-iteration 27-float D2H, host solves restated without Eigen, 1 host thread")
        steps = K
    else:
        # bounded CPU sample of the oracle port, all host threads
        nfr = min(K, args.cpu_frames)
        frames, gt = make_sequence(nfr + 1)
        tr = O.CpuTracker(cfg)
        tr.process(frames[0], 0, 0)
        t0 = time.perf_counter()
        for i in range(1, nfr + 1):
            tr.process(frames[i], i, 0)
        dt = time.perf_counter() - t0
        tr.close()
        fps = nfr / dt
        kind, cores, steps = "port", int(O.cpu().orc_num_threads()), nfr
        sample = f"{nfr} frames of the same workload on the host cores (C oracle port, OpenMP); reference CUDA build or GPU not available"
    return {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": W,
        "ms_per_step": 1e3 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic analytic scene, 16-bit mm depth, periodic 500-frame trajectory",
        "config": {"workload": "C2 full pipeline: 640x480, 3-level pyramid projective ICP {10,5,4} + TSDF integrate + raycast, 512^3 volume over 2 m"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0 if kind == "port" else 1228800, "d2h_bytes_per_step": 0 if kind == "port" else 19 * 108},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=12, help="frames of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="c2 (default): BASELINE configs[1], replicas at N>1; c4: z-slab-sharded single sequence; c1/c3/c5: the other BASELINE configs")
    ap.add_argument("--c4-resolution", type=int, default=1024)
    args = ap.parse_args()
    if args.warmup < 1:
        args.warmup = 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("HKF_NCCL_DEBUG", "WARN")     # keep stdout to the one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if args.workload == "c4":
        out = run_c4(args, rank, world, local_rank)
    elif args.workload in ("c1", "c3", "c5"):
        out = run_config(args, rank, world, local_rank)
    else:
        out = run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
