"""Development tool: where does the end-to-end time of the plugin's in-core filter() go?

Times, on cfg 2 (512x1024x1024, sigma 2): pinned H2D / D2H rates (contiguous and pitched column ranges), the three
passes device-resident as one call per pass and as the windowed sequence filter() uses, and filter() itself with its
phases stamped by CUDA events. Prints one JSON object. Not part of the product or the tests.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from flowdenoising_b200 import _lib                                     # noqa: E402
from flowdenoising_b200 import flowdenoising as fd                      # noqa: E402
from flowdenoising_b200.engine import DeviceEngine, FlowParams, gaussian_kernel  # noqa: E402
from flowdenoising_b200._lib import View                                # noqa: E402


def ev_ms(fn, reps=1):
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    Z, Y, X = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (512, 1024, 1024)))
    lib = _lib.load()
    eng = DeviceEngine()
    dev = eng.device
    out = {}
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    d_vol = (torch.rand((Z, Y, X), device=dev, generator=g) * 200 + torch.randn((Z, Y, X), device=dev, generator=g) * 10)
    host = torch.empty((Z, Y, X), dtype=torch.float32, pin_memory=True)
    host.copy_(d_vol)
    host2 = torch.empty((Z, Y, X), dtype=torch.float32, pin_memory=True)
    st = torch.cuda.current_stream().cuda_stream
    nbytes = 4.0 * Z * Y * X

    # ---- transfer rates ----
    t = ev_ms(lambda: d_vol.copy_(host, non_blocking=True)); out["h2d_GBps"] = nbytes / t / 1e6
    t = ev_ms(lambda: host2.copy_(d_vol, non_blocking=True)); out["d2h_GBps"] = nbytes / t / 1e6
    for w in (X - 128, 128, 64):
        t = ev_ms(lambda: _lib.check(lib.fdn_copy2d_async(host2.data_ptr(), 4 * X, d_vol.data_ptr(), 4 * X, 4 * w,
                                                          Z * Y, 1, st)))
        out[f"d2h_cols{w}_GBps"] = 4.0 * w * Z * Y / t / 1e6
        out[f"d2h_cols{w}_ms"] = t
    # a column range staged contiguous on the device first, then one linear copy
    stage = torch.empty((Z, Y, 128), dtype=torch.float32, device=dev)
    t = ev_ms(lambda: _lib.check(lib.fdn_copy2d_async(stage.data_ptr(), 4 * 128, d_vol.data_ptr(), 4 * X, 4 * 128,
                                                      Z * Y, 2, st)))
    out["d2d_cols128_ms"] = t

    # ---- compute: one call per pass vs the windowed sequence ----
    k = gaussian_kernel(2.0)
    ks = [k, k, k]
    flow = FlowParams()
    for _ in range(2):
        res = eng.filter(d_vol, ks, flow); del res
    out["filter_plain_ms"] = ev_ms(lambda: eng.filter(d_vol, ks, flow))
    a = torch.empty_like(d_vol); b = torch.empty_like(d_vol)
    head, tail = 64, 128
    vz = lambda n, f: View(Z, n, f, 1, Y, X, Y * X, X, Y * X, X)
    out["z_full_ms"] = ev_ms(lambda: eng.filter_view(d_vol, a, vz(Z, 0), k, flow))
    out["z_head_ms"] = ev_ms(lambda: eng.filter_view(d_vol, a, vz(head, 0), k, flow))
    out["z_tail_ms"] = ev_ms(lambda: eng.filter_view(d_vol, a[head:], vz(Z - head, head), k, flow))
    vt = eng.transpose_yx(d_vol)
    ot = torch.empty_like(vt)
    vx = lambda n, f: View(X, n, f, 1, Z, Y, Y, X * Y, Y, X * Y)
    out["x_full_ms"] = ev_ms(lambda: eng.filter_view(vt, ot, vx(X, 0), k, flow))
    out["x_body_ms"] = ev_ms(lambda: eng.filter_view(vt, ot, vx(X - tail, 0), k, flow))
    out["x_tail_ms"] = ev_ms(lambda: eng.filter_view(vt, ot.view(-1)[(X - tail) * Y:], vx(tail, X - tail), k, flow))
    del a, b, vt, ot, stage
    eng.release_workspace()
    torch.cuda.empty_cache()

    # ---- filter() of the plugin, wall clock, plain and overlapped ----
    for name, thr in (("plain", 1 << 62), ("overlapped", 0)):
        fd._OVERLAP_MIN_BYTES = thr
        ts = []
        for i in range(3):
            host.copy_(d_vol); torch.cuda.synchronize()
            obj = fd.FlowDenoising(1, host.numpy())
            obj.filtered_vol = host2.numpy()
            fd._TRACE = [] if thr == 0 else None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            obj.filter(ks)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
            if fd._TRACE:
                tr = fd._TRACE
                out["trace_device_ms"] = {lab: round(tr[0][1].elapsed_time(e), 2) for lab, e, _ in tr[1:]}
                out["trace_host_ms"] = {lab: round((w - t0) * 1e3, 2) for lab, _e, w in tr}
            fd._TRACE = None
        out[f"e2e_{name}_ms"] = ts
    print(json.dumps(out))


if __name__ == "__main__":
    main()
