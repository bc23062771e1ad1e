"""Development harness for the flow-iteration kernel (stage 3): times the iterations of one pyramid level in isolation
on realistic inputs and checks that the warp-specialised kernel (k_flow_iter_ws) and the strip kernel (k_flow_iter,
pinned against the oracle) write identical bits.

    [FDN_LIB_PATH=labso/variant.so] python tools/flow_iter_lab.py [--n 128] [--h 1024] [--w 1024] [--iters 3] [--reps 5]

Inputs: n+delta slices of the bench's synthetic volume -> polynomial expansions (fdn_polyexp); pair b = (slice b,
b+delta). The timed call runs `iters` iterations from a zero flow (what the level-0 launches of a pass see when the
coarser levels found nothing); --flow-scale adds a smooth synthetic displacement field to stress gathers far from
the identity position. Reports ms per iteration and the algorithmic-bytes roofline fraction (56 B per pixel and
iteration against MEASURED_PEAKS.json).
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--h", type=int, default=1024)
    ap.add_argument("--w", type=int, default=1024)
    ap.add_argument("--win", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--flow-scale", type=float, default=0.0)
    ap.add_argument("--delta", type=int, default=1, help="slice distance of a pair")
    ap.add_argument("--check", action="store_true", help="compare with the strip kernel bit for bit")
    a = ap.parse_args()

    import torch
    from bench import synthetic_volume_torch
    from flowdenoising_b200.engine import DeviceEngine

    eng = DeviceEngine()
    lib = eng.lib
    n, h, w = a.n, a.h, a.w
    dev = torch.device("cuda:0")
    vol = synthetic_volume_torch((n + a.delta, h, w), dev)
    rf = lib.fdn_polyexp_floats(h, w)
    R = torch.empty((n + a.delta, rf), dtype=torch.float32, device=dev)
    rc = lib.fdn_polyexp(vol.data_ptr(), n + a.delta, h, w, 5, 1.2, R.data_ptr(), None)
    assert rc == 0, lib.fdn_last_error()
    nscr = lib.fdn_flow_iteration_scratch_bytes(n, h, w)
    scr = torch.zeros(nscr, dtype=torch.uint8, device=dev)
    merged = hasattr(lib, "fdn_flow_iterations") and getattr(lib.fdn_flow_iterations, "argtypes", None) is not None
    bufs = [torch.zeros((n, h, w, 2), dtype=torch.float32, device=dev) for _ in range(3)]
    f0 = torch.zeros_like(bufs[0])
    if a.flow_scale:
        yy = torch.arange(h, device=dev, dtype=torch.float32)[None, :, None]
        xx = torch.arange(w, device=dev, dtype=torch.float32)[None, None, :]
        f0[..., 0] = a.flow_scale * torch.sin(xx / 97.0 + yy / 61.0)
        f0[..., 1] = a.flow_scale * torch.cos(xx / 83.0 - yy / 71.0)

    def run():
        bufs[0].copy_(f0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if merged:
            res = C.c_void_p()
            rc = lib.fdn_flow_iterations(R[0].data_ptr(), R[a.delta].data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(),
                                         bufs[2].data_ptr(), n, h, w, a.win, a.iters, scr.data_ptr(), nscr, None,
                                         C.byref(res))
            assert rc == 0, lib.fdn_last_error()
            out = [b for b in bufs if b.data_ptr() == res.value][0]
        else:
            for i in range(a.iters):
                rc = lib.fdn_flow_iteration(R[0].data_ptr(), R[a.delta].data_ptr(), bufs[i % 3].data_ptr(),
                                            bufs[(i + 1) % 3].data_ptr(), n, h, w, a.win, scr.data_ptr(), nscr, None)
                assert rc == 0, lib.fdn_last_error()
            out = bufs[a.iters % 3]
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out

    run(); run()
    ts = sorted(run()[0] for _ in range(a.reps))
    med = ts[len(ts) // 2] / a.iters
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6545.9
    gb = 56.0 * n * h * w / 1e9
    out = run()[1]
    mag = out.abs()
    print(f"{os.environ.get('FDN_LIB_PATH', 'product')}: {n} pairs {h}x{w} x{a.iters} it, merged={merged}: "
          f"{med:.3f} ms/iteration (min {ts[0] / a.iters:.3f}) -> {gb / med * 1e3:.0f} GB/s algorithmic = "
          f"{gb / med * 1e3 / peak:.3f} of the HBM peak; result mean|flow| {mag.mean().item():.3f}")
    if a.check and hasattr(lib, "fdn_set_flow_iter_variant"):
        got = out.clone()
        lib.fdn_set_flow_iter_variant(0)
        ref = run()[1]
        lib.fdn_set_flow_iter_variant(1)
        same = torch.equal(ref.view(torch.int32), got.view(torch.int32))
        print("bit-identical to the strip kernel:", same)
        if not same:
            sys.exit(1)


if __name__ == "__main__":
    main()
