#!/usr/bin/env python
"""Golden vectors at the BENCHMARKED slice sizes, from the UNMODIFIED reference (/root/reference/src/flowdenoising.py).

    python oracle/gen_golden_big.py [--only cfg2|cfg4]

Run in the build container only (the GPU box has no /root/reference). Same import recipe as oracle/gen_golden.py
(SURVEY.md App. D). The inputs are seeded synthetic volumes (oracle.fd_oracle.synthetic_volume: integer-valued,
reproducible on any host) that the tests regenerate; the fixtures hold only the input's SHA-256, the SHA-256 of
every output slice and a small crop of each slice for diagnostics -- a few hundred KB in total.

cfg2_crops.npz -- BASELINE.json configs[1] slice geometry (sigma = 2 -> 17 taps, levels=3, winsize=5):
  z:  volume (20, 1024, 1024), Z pass, output slices 0 (periodic wrap) and 9   (1024 x 1024 slices, 4 pyramid levels)
  y:  volume (512, 20, 1024),  Y pass, output slice 3                          (512 x 1024 slices, row-strided views)
  x:  volume (512, 1024, 20),  X pass, output slice 11                         (512 x 1024 slices, element-strided)
cfg4_crops.npz -- BASELINE.json configs[3] parameters (sigma = (4, 2, 2), levels=5, winsize=9, uint8 input):
  z:  volume (34, 1024, 2048) uint8, Z pass with sigma 4 (33 taps), output slice 17   (6 pyramid levels, 79-tap blur)
  y:  volume (256, 20, 2048)  uint8, Y pass with sigma 2, output slice 5              (256 x 2048 slices, 4 levels)
Reference code driven: FlowDenoising.filter_along_{Z,Y,X}_slice (:306-373) with get_flow_with_prev_flow (:65-87)
and warp_slice (:55-63); TIFF input is cast to float32 before filtering (:475).
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import fd_oracle as O          # noqa: E402  (only for synthetic_volume)
from oracle.gen_golden import load_reference, versions, sha, GOLDEN   # noqa: E402

# name -> (shape, seed, noise, axis, sigma, levels, winsize, output slices, dtype of the stored input)
CASES = {
    "cfg2": {
        "z": ((20, 1024, 1024), 101, 10.0, 0, 2.0, 3, 5, (0, 9)),
        "y": ((512, 20, 1024), 102, 10.0, 1, 2.0, 3, 5, (3,)),
        "x": ((512, 1024, 20), 103, 10.0, 2, 2.0, 3, 5, (11,)),
    },
    "cfg4": {
        "z": ((34, 1024, 2048), 104, 10.0, 0, 4.0, 5, 9, (17,)),
        "y": ((256, 20, 2048), 105, 10.0, 1, 2.0, 5, 9, (5,)),
    },
}
CROP = 96   # stored corner crop of every output slice


def run_case(ref, shape, seed, noise, axis, sigma, l, w, slices, as_uint8):
    vol = O.synthetic_volume(shape, seed=seed, noise_sigma=noise)
    if as_uint8:
        vol = vol.astype(np.uint8).astype(np.float32)    # TIFF stack -> float32 at load (src/flowdenoising.py:475)
    ref.vol, ref.l, ref.w = vol, l, w
    ref.get_flow = ref.get_flow_with_prev_flow
    kernel = ref.get_gaussian_kernel(sigma)
    obj = ref.FlowDenoising(1, vol, l, w, ref.get_flow_with_prev_flow, ref.warp_slice)
    f = [obj.filter_along_Z_slice, obj.filter_along_Y_slice, obj.filter_along_X_slice][axis]
    outs = []
    for s in slices:
        t0 = time.perf_counter()
        f(s, kernel)
        o = np.ascontiguousarray(np.take(obj.filtered_vol, s, axis=axis))
        outs.append(o)
        print(f"  axis {axis} slice {s}: {o.shape} in {time.perf_counter() - t0:.1f} s", flush=True)
    return vol, outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    ref = load_reference()
    ver = versions()
    for cfg, cases in CASES.items():
        if a.only and a.only != cfg:
            continue
        d = {"versions": str(ver)}
        for name, (shape, seed, noise, axis, sigma, l, w, slices) in cases.items():
            print(cfg, name, shape, flush=True)
            vol, outs = run_case(ref, shape, seed, noise, axis, sigma, l, w, slices, as_uint8=(cfg == "cfg4"))
            d[f"{name}_shape"] = np.array(shape)
            d[f"{name}_params"] = np.array([seed, axis, l, w], dtype=np.int64)
            d[f"{name}_noise_sigma"] = np.array([noise, sigma])
            d[f"{name}_slices"] = np.array(slices)
            d[f"{name}_input_sha256"] = sha(vol)
            for s, o in zip(slices, outs):
                d[f"{name}_sha_{s}"] = sha(o)
                d[f"{name}_crop_{s}"] = o[:CROP, :CROP].copy()
                d[f"{name}_mean_{s}"] = float(o.astype(np.float64).mean())
        path = os.path.join(GOLDEN, f"{cfg}_crops.npz")
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path), flush=True)


if __name__ == "__main__":
    main()
