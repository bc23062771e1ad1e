"""Development harness for the flow-iteration kernel (stage 3): times one launch shape in isolation on realistic
inputs and checks that the warp-specialised kernel (k_flow_iter_ws) and the strip kernel (k_flow_iter, already pinned
against the oracle) write identical bits.

    python tools/flow_iter_lab.py [--n 128] [--h 1024] [--w 1024] [--win 5] [--reps 10] [--iters 3]

Inputs: n+1 slices of the bench's synthetic volume -> polynomial expansions (fdn_polyexp); pair b = (slice b, b+1).
The flow fed to the timed launch is the result of `iters - 1` earlier iterations from a zero flow (what the
level-0 launches of a pass see when the coarser levels found nothing), optionally scaled (--flow-scale) to
stress gathers far from the identity position.

Phase-removal experiments (which phase is the critical path?): build with
    FDN_NVCC_EXTRA=-DFDN_WS_EXPERIMENTS python -m flowdenoising_b200._build
and run with FDN_EXP=<bit mask> (1: no scan chain, 2: no packet wait, 4: no column-sum update, 8: no solve). Results are
wrong with any bit set; the product build ignores FDN_EXP.
"""
import argparse
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--h", type=int, default=1024)
    ap.add_argument("--w", type=int, default=1024)
    ap.add_argument("--win", type=int, default=5)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--flow-scale", type=float, default=1.0)
    ap.add_argument("--delta", type=int, default=1, help="slice distance of a pair")
    ap.add_argument("--only", default=None, help="old | ws: skip the comparison")
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
    scr = torch.empty(nscr, dtype=torch.uint8, device=dev)

    def run(variant, fin, fout):
        os.environ["FDN_FLOW_ITER"] = variant
        rc = lib.fdn_flow_iteration(R[0].data_ptr(), R[a.delta].data_ptr(), fin.data_ptr(), fout.data_ptr(), n, h, w,
                                    a.win, scr.data_ptr(), nscr, None)
        assert rc == 0, lib.fdn_last_error()

    f0 = torch.zeros((n, h, w, 2), dtype=torch.float32, device=dev)
    f1 = torch.empty_like(f0)
    for _ in range(a.iters - 1):
        run("old", f0, f1)
        f0, f1 = f1, f0
    if a.flow_scale != 1.0:
        f0 *= a.flow_scale
    torch.cuda.synchronize()
    mag = f0.abs()
    print(f"input flow: mean|d| {mag.mean().item():.3f}  p99 {mag.flatten()[::97].quantile(0.99).item():.3f}  "
          f"max {mag.max().item():.2f}")

    outs = {}
    for variant in (["old", "ws"] if a.only is None else [a.only]):
        out = torch.empty_like(f0)
        run(variant, f0, out)
        run(variant, f0, out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
        ev[0].record()
        for i in range(a.reps):
            run(variant, f0, out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))
        med = ts[len(ts) // 2]
        gb = 56.0 * n * h * w / 1e9
        print(f"{variant}: median {med:.3f} ms  min {ts[0]:.3f}  -> {gb / med * 1e3:.0f} GB/s algorithmic "
              f"({gb / med * 1e3 / 6545.9:.3f} of the measured HBM peak)")
        outs[variant] = out
    if len(outs) == 2:
        same = torch.equal(outs["old"].view(torch.int32), outs["ws"].view(torch.int32))
        nd = (outs["old"].view(torch.int32) != outs["ws"].view(torch.int32)).sum().item()
        print("bit-identical:", same, "differing words:", nd)
        if not same:
            d = (outs["old"] - outs["ws"]).abs()
            idx = torch.nonzero(outs["old"].view(torch.int32) != outs["ws"].view(torch.int32))[:5]
            print("max|d|", d.max().item(), "first diffs at", idx.tolist())
            sys.exit(1)


if __name__ == "__main__":
    main()
