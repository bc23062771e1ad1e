#!/usr/bin/env python
"""bench.py -- Mvoxels/s of the OF-driven separable Gaussian denoise (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path on the host cores

A step = one full three-pass filter (Z, Y, X; src/flowdenoising.py:285-290) of the workload volume.
Workload at every N: BASELINE.json configs[1] -- 512x1024x1024 float32 FIB-SEM-like synthetic volume, sigma=2 per
axis, Farneback defaults (levels=3, winsize=5, iterations=3, poly 5/1.2). At N>1 the same volume is slab-sharded
(strong scaling; halo exchange + all-to-all re-slab, flowdenoising_b200/dist.py).

JSON keys follow the bench contract: `value` = device-resident whole-job throughput, `e2e` = the same through the
reference-facing Python classes with host buffers (H2D/D2H inside the timed region), `roofline` for the dominant
kernel (k_flow_iter) from CUDA events inside the timed region, `cpu_baseline` = the reference's CPU path (oracle
driver calling cv2 exactly like src/flowdenoising.py:306-327) timed on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mvoxels/s OF-Gaussian denoise"
UNIT = "Mvoxel/s"
# SURVEY.md §8d algorithmic-bytes model, cfg 2: 3 * 3811 B/voxel
MODEL_BYTES_PER_VOXEL_CFG2 = 11434.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", type=int, nargs=3, default=[512, 1024, 1024], help="Z Y X (default: cfg 2)")
    ap.add_argument("--sigma", type=float, default=2.0)
    ap.add_argument("--no-of", action="store_true", help="cfg 3: OF disabled (plain separable Gaussian)")
    ap.add_argument("--fast-noof", action="store_true", help="no-OF path with float32 FMA arithmetic")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--cpu-slices", type=int, default=0, help="slices per axis in the CPU sample (0: one per core)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
def synthetic_volume_torch(shape, device, seed=1, noise_sigma=10.0):
    """FIB-SEM-like 8-bit-amplitude volume generated on the device (structure drifting ~1 px/slice
    + membranes + Gaussian noise, clipped to [0, 255])."""
    import torch
    Z, Y, X = shape
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(shape, dtype=torch.float32, device=device)
    y = torch.arange(Y, device=device, dtype=torch.float32)[None, :, None]
    x = torch.arange(X, device=device, dtype=torch.float32)[None, None, :]
    step = 16
    for z0 in range(0, Z, step):
        z1 = min(Z, z0 + step)
        z = torch.arange(z0, z1, device=device, dtype=torch.float32)[:, None, None]
        v = 128.0 + 60.0 * torch.sin(x / 9.0 + z / 13.0) * torch.cos(y / 7.0 - z / 17.0) \
            + 30.0 * torch.sin((x + y) / 23.0)
        rr = torch.sqrt((x - X / 2 - 0.7 * z) ** 2 + (y - Y / 2 + 0.4 * z) ** 2)
        v = v + 35.0 * torch.tanh(4.0 * torch.sin(rr / 37.0))           # membrane-like rings drifting with z
        v = v + noise_sigma * torch.randn(v.shape, device=device, generator=g)
        out[z0:z1] = torch.clamp(v, 0.0, 255.0)
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.err = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get_reasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.2)
        except Exception as e:  # NVML missing: report, do not fail the bench
            self.err = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        d = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
             "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            d["error"] = self.err
        return d


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(vol_host, kernels, use_of, slices_per_axis, cores):
    """The reference's CPU path (oracle driver -> cv2.calcOpticalFlowFarneback / cv2.remap with a thread pool of
    `cores` workers, src/flowdenoising.py:187-206, :306-327) on `slices_per_axis` slices of each pass of the full-size
    volume; returns (extrapolated Mvoxel/s, seconds spent, description)."""
    from oracle import fd_oracle as O
    Z, Y, X = vol_host.shape
    o = O.OracleDenoiser(cores, vol_host, use_OF=use_of, backend="cv2")
    total = 0.0
    spent = 0.0
    for axis, n in enumerate((Z, Y, X)):
        ns = min(n, slices_per_axis)
        idx = [int(i) for i in np.linspace(0, n - 1, ns).round()]
        t0 = time.perf_counter()
        o.filter_along_axis(axis, kernels[axis], indices=idx)
        dt = time.perf_counter() - t0
        spent += dt
        total += dt * n / ns
    mvox = Z * Y * X / total / 1e6
    try:
        import cv2
        ver = cv2.__version__
    except Exception:
        ver = "?"
    desc = (f"{slices_per_axis} slices of each of the Z/Y/X passes on the full {Z}x{Y}x{X} volume, {cores} threads, "
            f"cv2 {ver}; extrapolated linearly in slices (per-slice cost is uniform)")
    return mvox, spent, desc


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import fd_oracle as O
    shape = tuple(args.shape)
    cores = os.cpu_count() or 1
    spa = args.cpu_slices or cores
    vol = O.synthetic_volume(shape, seed=1, noise_sigma=10.0) if np.prod(shape) <= (1 << 24) else None
    if vol is None:
        # full-size host volume: cheap separable synthetic structure + noise (same family as the GPU generator)
        rng = np.random.default_rng(1)
        Z, Y, X = shape
        vol = np.empty(shape, np.float32)
        y = np.arange(Y, dtype=np.float32)[:, None]
        x = np.arange(X, dtype=np.float32)[None, :]
        for z in range(Z):
            v = 128.0 + 60.0 * np.sin(x / 9.0 + z / 13.0) * np.cos(y / 7.0 - z / 17.0) + 30.0 * np.sin((x + y) / 23.0)
            rr = np.sqrt((x - X / 2 - 0.7 * z) ** 2 + (y - Y / 2 + 0.4 * z) ** 2)
            v = v + 35.0 * np.tanh(4.0 * np.sin(rr / 37.0)) + rng.standard_normal((Y, X), dtype=np.float32) * 10.0
            vol[z] = np.clip(v, 0, 255)
    kernels = [O.get_gaussian_kernel(args.sigma)] * 3
    vals = []
    desc = ""
    for i in range(args.warmup + args.steps):
        mv, spent, desc = cpu_reference_sample(vol, kernels, not args.no_of, spa, cores)
        if i >= args.warmup:
            vals.append(mv)
    v = float(np.mean(vals))
    nvox = float(np.prod(shape))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": nvox / v / 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    Z, Y, X = args.shape
    return {"workload": f"{Z}x{Y}x{X} float32 FIB-SEM-like synthetic volume, sigma={args.sigma:g} per axis, "
                        + ("OF disabled (plain separable Gaussian)" if args.no_of else
                           "Farneback levels=3 winsize=5 iterations=3 poly_n=5 poly_sigma=1.2"),
            "baseline_config": "BASELINE.json configs[2]" if args.no_of else "BASELINE.json configs[1]",
            "shape": [Z, Y, X], "sigma": args.sigma, "of": not args.no_of,
            "l2_policy": "inputs_exceed_l2 (volume and cached expansions are >> 126 MB)",
            "model_bytes_per_voxel": 24.0 if args.no_of else MODEL_BYTES_PER_VOXEL_CFG2}


# ----------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    from flowdenoising_b200 import _lib
    from flowdenoising_b200 import flowdenoising as fd
    from flowdenoising_b200.engine import DeviceEngine, FlowParams, gaussian_kernel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    shape = tuple(args.shape)
    Z, Y, X = shape
    nvox = float(Z * Y * X)
    kernels = [gaussian_kernel(args.sigma)] * 3
    flow = None if args.no_of else FlowParams()
    exact = not args.fast_noof
    eng = DeviceEngine(device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- workload (device resident before the timed region) ----
    if world == 1:
        d_vol = synthetic_volume_torch(shape, device, seed=1)
        runner = lambda: eng.filter(d_vol, kernels, flow, exact=exact)
    else:
        from flowdenoising_b200.dist import DistributedDenoiser
        dd = DistributedDenoiser(eng, shape, flow, exact=exact)
        z0, z1 = dd.z_range
        full_seeded = synthetic_volume_torch(shape, device, seed=1) if Z * Y * X <= (1 << 31) else None
        d_slab = full_seeded[z0:z1].contiguous() if full_seeded is not None else \
            synthetic_volume_torch((z1 - z0, Y, X), device, seed=1 + rank)
        del full_seeded
        runner = lambda: dd.filter(d_slab, kernels)

    # ---- device-resident throughput (`value`) ----
    for _ in range(args.warmup):
        res = runner()
    barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    lib.fdn_reset_launch_count()
    lib.fdn_profile_reset()
    lib.fdn_profile_enable(1)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res = runner()
    ev1.record()
    barrier()
    lib.fdn_profile_enable(0)
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = int(lib.fdn_launch_count())
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = nvox / (ms_step * 1e-3) / 1e6

    # ---- per-kernel event times -> roofline of the dominant kernel ----
    kern = {}
    for i in range(lib.fdn_profile_kernel_count()):
        ms = C.c_double(); n = C.c_int64(); by = C.c_double()
        _lib.check(lib.fdn_profile_read(i, C.byref(ms), C.byref(n), C.byref(by)))
        if n.value:
            kern[lib.fdn_profile_kernel_name(i).decode()] = {"ms": ms.value, "launches": int(n.value), "bytes": by.value}
    # the dominant kernel per pyramid-level image size (launch records carry n, h, w)
    by_level = {}
    flow_id = [i for i in range(lib.fdn_profile_kernel_count()) if lib.fdn_profile_kernel_name(i) == b"k_flow_iter"][0]
    for i in range(lib.fdn_profile_record_count()):
        kid = C.c_int(); rn = C.c_int(); rh = C.c_int(); rw = C.c_int(); ms = C.c_double(); by = C.c_double()
        _lib.check(lib.fdn_profile_record(i, C.byref(kid), C.byref(rn), C.byref(rh), C.byref(rw), C.byref(ms), C.byref(by)))
        if kid.value == flow_id:
            d = by_level.setdefault(f"{rh.value}x{rw.value}", {"ms": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += ms.value; d["bytes"] += by.value; d["launches"] += 1
    lib.fdn_profile_reset()
    peak, peak_src = measured_peak()
    dom = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
    roofline = None
    if dom:
        kd = kern[dom]
        achieved = kd["bytes"] / (kd["ms"] * 1e-3) / 1e9
        # DRAM traffic per launch: ratio measured once with `ncu --set full` (profiles/r1_traffic.json) x the
        # algorithmic bytes of this run's average launch; null if no capture is committed for the kernel
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tr = json.load(f).get(dom)
            if tr:
                traffic = tr["dram_bytes_per_algorithmic_byte"] * kd["bytes"] / kd["launches"]
                traffic_src = tr["source"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src,
                    "avg_launch_ms": kd["ms"] / kd["launches"], "launches": kd["launches"],
                    "algorithmic_bytes_per_launch": kd["bytes"] / kd["launches"],
                    "share_of_step": kd["ms"] / (ms_total if ms_total > 0 else 1.0),
                    "kernel_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in sorted(kern.items())},
                    # every kernel of the path against the same peak (algorithmic bytes / CUDA-event time)
                    "kernel_frac_of_peak": {k: round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 3)
                                            for k, v in sorted(kern.items()) if v["ms"] > 0},
                    "flow_iter_by_level": {k: {"ms_per_step": round(v["ms"] / args.steps, 2), "launches": v["launches"],
                                               "frac_of_peak": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 3)}
                                           for k, v in by_level.items() if v["ms"] > 0}}
    model_b = 24.0 if args.no_of else MODEL_BYTES_PER_VOXEL_CFG2 * (1.0 if shape == (512, 1024, 1024) else float("nan"))
    whole_job_frac = (model_b * nvox / (ms_step * 1e-3) / 1e9) / (world * peak) if model_b == model_b else None

    # ---- end to end through the reference-facing classes with host buffers ----
    e2e = None
    if not args.skip_e2e and world == 1:
        host = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        host.copy_(d_vol)
        pristine = host.clone()
        h_vol = host.numpy()
        n_e2e = max(1, min(args.steps, 2))
        times = []
        for i in range(1 + n_e2e):          # first one is warm-up
            h_vol[...] = pristine.numpy()   # filter() overwrites vol with the Z+Y intermediate (reference :289)
            obj = fd.GaussianDenoising(os.cpu_count(), h_vol) if args.no_of else \
                fd.FlowDenoising(os.cpu_count(), h_vol, 3, 5, fd.get_flow_with_prev_flow, fd.warp_slice)
            obj.exact = exact
            obj.filtered_vol = torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            obj.filter(kernels)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i > 0:
                times.append(dt)
        e2e_s = float(np.mean(times))
        e2e = {"value": nvox / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(nvox * 4),
               "d2h_bytes_per_step": int(nvox * 8), "ms_per_step": e2e_s * 1e3, "steps": n_e2e,
               "api": "flowdenoising_b200.flowdenoising.FlowDenoising(P, vol_numpy, ...).filter(kernels)"}
        del host, pristine, obj
    elif world > 1 and not args.skip_e2e:
        # host slab -> device -> distributed filter -> host slab, per rank (pinned); max over ranks
        host = torch.empty(d_slab.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(d_slab)
        out_host = torch.empty(d_slab.shape, dtype=torch.float32, pin_memory=True)
        times = []
        for i in range(2):
            barrier()
            t0 = time.perf_counter()
            d_tmp = host.to(device, non_blocking=True)
            zy, zyx = dd.filter(d_tmp, kernels)
            out_host.copy_(zyx, non_blocking=True)
            barrier()
            times.append(time.perf_counter() - t0)
        tt = torch.tensor([times[-1]], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
        e2e = {"value": nvox / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(nvox * 4),
               "d2h_bytes_per_step": int(nvox * 4), "ms_per_step": e2e_s * 1e3, "steps": 1,
               "api": "flowdenoising_b200.dist.DistributedDenoiser.filter on pinned host slabs"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cores = os.cpu_count() or 1
        vol_host = d_vol.cpu().numpy()
        mv, spent, desc = cpu_reference_sample(vol_host, kernels, not args.no_of, args.cpu_slices or cores, cores)
        cpu = {"value": mv, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc, "seconds_spent": spent}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks, "gpu_launches": launches,
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
            "whole_job_model_frac_of_hbm_peak": whole_job_frac,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
