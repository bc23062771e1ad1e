"""Times the no-OF passes (exact / fast) on a 512x1024x1024 volume: [FDN_LIB_PATH=variant.so] python tools/noof_lab.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synthetic_volume_torch
from flowdenoising_b200.engine import DeviceEngine, gaussian_kernel
eng = DeviceEngine()
dev = torch.device("cuda:0")
vol = synthetic_volume_torch((512, 1024, 1024), dev)
out = torch.empty_like(vol)
k = gaussian_kernel(2.0)
res = []
for exact in (True, False):
    for axis in (0, 1, 2):
        for _ in range(2):
            eng.filter_along_axis(vol, out, axis, k, None, exact=exact)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.filter_along_axis(vol, out, axis, k, None, exact=exact)
        e1.record(); torch.cuda.synchronize()
        res.append(f"{'exact' if exact else 'fast'} axis{axis} {e0.elapsed_time(e1) / 5:.3f} ms")
print(os.environ.get("FDN_LIB_PATH", "product"), "|", " | ".join(res))
