#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src/flowdenoising.py).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

The reference is imported with empty stub modules for its file-I/O imports (mrcfile, skimage, tifffile,
imageio -- not installed here, not on the hot path) exactly as described in SURVEY.md App. D, and its own
classes / functions are driven on seeded synthetic volumes. Versions that produced the committed fixtures
are stored in every file (cv2 4.13.0, scipy 1.18.1, numpy 2.3.5).

Fixtures (all small):
  kernels.npz        get_gaussian_kernel(sigma) taps (src/flowdenoising.py:34-45)
  flows.npz          get_flow_with_prev_flow / without (:65-114) on slice pairs, zero-init and chained;
                     warp_slice (:55-63) outputs
  toy_of.npz         FlowDenoising(...).filter per pass (Z, ZY, ZYX) on a 12x64x72 volume (:285-290, :306-373)
  toy_noof.npz       GaussianDenoising(...).filter per pass on the same volume (:133-158)
  toy_of_recompute.npz  same with get_flow_without_prev_flow (--recompute_flow, :89-114, :442-447), Z pass
  cfg1_slices.npz    BASELINE.json configs[0] (64x256x256 f32, sigma=2, Farneback defaults): a few slices of
                     the volume after each pass + SHA-256 of the input
"""
import argparse
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import fd_oracle as O  # noqa: E402  (only for synthetic_volume)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_PATH = "/root/reference/src/flowdenoising.py"


def load_reference():
    for name in ["imageio", "tifffile", "skimage", "skimage.io", "mrcfile"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("fd_ref", REF_PATH)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref.args = argparse.Namespace(input="synthetic")
    return ref


def versions():
    import cv2, scipy
    return dict(cv2=cv2.__version__, scipy=scipy.__version__, numpy=np.__version__)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_filter(ref, vol, sigmas, use_of, l=3, w=5, recompute=False, P=5, passes=3):
    """Drive the reference classes; returns [after Z, after ZY, after ZYX]."""
    v = vol.copy()
    ref.vol, ref.l, ref.w = v, l, w
    gf = ref.get_flow_without_prev_flow if recompute else ref.get_flow_with_prev_flow
    ref.get_flow = gf
    kernels = [ref.get_gaussian_kernel(s) for s in sigmas]
    obj = ref.FlowDenoising(P, v, l, w, gf, ref.warp_slice) if use_of else ref.GaussianDenoising(P, v)
    outs = []
    obj.filter_along_Z(kernels[0]); outs.append(obj.filtered_vol.copy())
    if passes > 1:
        obj.vol[...] = obj.filtered_vol[...]
        obj.filter_along_Y(kernels[1]); outs.append(obj.filtered_vol.copy())
    if passes > 2:
        obj.vol[...] = obj.filtered_vol[...]
        obj.filter_along_X(kernels[2]); outs.append(obj.filtered_vol.copy())
    return outs


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = load_reference()
    ver = versions()
    print("reference loaded; versions", ver)

    # --- kernels ---
    sig = [0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 4.0]
    np.savez(os.path.join(GOLDEN, "kernels.npz"), sigmas=np.array(sig),
             **{f"k_{s}": ref.get_gaussian_kernel(s) for s in sig}, versions=str(ver))

    # --- flows / warps on slice pairs ---
    d = {}
    cases = [("a", (128, 160), 3, 5, 11), ("b", (96, 130), 3, 5, 12), ("d", (128, 128), 5, 9, 14)]
    for name, shape, l, w, seed in cases:
        v = O.synthetic_volume((4,) + shape, seed=seed, noise_sigma=10.0)
        d[f"{name}_vol"] = v.astype(np.uint8)
        d[f"{name}_lw"] = np.array([l, w])
        centre = v[0]
        prev = np.zeros(shape + (2,), np.float32)
        for j in (1, 2, 3):   # chained like the reference's forward chain (:319-324)
            flow = ref.get_flow_with_prev_flow(v[j], centre, l, w, prev)
            prev = flow
            if j != 2:
                d[f"{name}_flow_chain{j}"] = flow.copy()
            if j == 3:
                d[f"{name}_warp_chain{j}"] = ref.warp_slice(v[j], flow)
        d[f"{name}_flow_noprev2"] = ref.get_flow_without_prev_flow(v[2], centre, l, w, None)
    np.savez_compressed(os.path.join(GOLDEN, "flows.npz"), versions=str(ver), **d)

    # --- toy volumes through the reference classes ---
    toy = O.synthetic_volume((12, 64, 72), seed=21, noise_sigma=8.0)
    sig3 = (1.0, 1.5, 0.75)
    of = run_filter(ref, toy, sig3, True)
    np.savez_compressed(os.path.join(GOLDEN, "toy_of.npz"), vol=toy.astype(np.uint8), sigmas=np.array(sig3),
                        l=3, w=5, Z=of[0], ZY=of[1], ZYX=of[2], versions=str(ver))
    # P-invariance of the reference itself
    of1 = run_filter(ref, toy, sig3, True, P=1)
    assert all(np.array_equal(a, b) for a, b in zip(of, of1)), "reference not P-invariant?"
    noof = run_filter(ref, toy, sig3, False)
    np.savez_compressed(os.path.join(GOLDEN, "toy_noof.npz"), vol=toy.astype(np.uint8), sigmas=np.array(sig3),
                        Z=noof[0], ZY=noof[1], ZYX=noof[2], versions=str(ver))
    # float-valued (non-integer) toy for the no-OF path incl. sigma=2 (17 taps > Z: multiple wraps)
    toyf = O.synthetic_volume((10, 24, 40), seed=22, noise_sigma=8.0, quantise=False) - 100.0
    noof2 = run_filter(ref, toyf.astype(np.float32), (2.0, 2.0, 2.0), False)
    np.savez_compressed(os.path.join(GOLDEN, "toy_noof_float.npz"), vol=toyf.astype(np.float32),
                        sigmas=np.array((2.0, 2.0, 2.0)), Z=noof2[0], ZY=noof2[1], ZYX=noof2[2], versions=str(ver))
    rec = run_filter(ref, toy, sig3, True, recompute=True, passes=1)
    np.savez_compressed(os.path.join(GOLDEN, "toy_of_recompute.npz"), vol=toy.astype(np.uint8),
                        sigmas=np.array(sig3), l=3, w=5, Z=rec[0], versions=str(ver))
    print("toy fixtures done")

    # --- cfg 1 (BASELINE.json configs[0]) ---
    vol = O.synthetic_volume((64, 256, 256), seed=0, noise_sigma=20.0)
    outs = run_filter(ref, vol, (2.0, 2.0, 2.0), True, P=os.cpu_count())
    zs = np.array([0, 37])
    np.savez_compressed(os.path.join(GOLDEN, "cfg1_slices.npz"), input_sha256=sha(vol), zs=zs,
                        Z=outs[0][zs], ZY=outs[1][zs], ZYX=outs[2][zs],
                        sha_Z=sha(outs[0]), sha_ZY=sha(outs[1]), sha_ZYX=sha(outs[2]),
                        mean_ZYX=float(outs[2].astype(np.float64).mean()), versions=str(ver))
    print("cfg1 done")
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
