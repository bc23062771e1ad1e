"""Development tool: interleaved A/B/D samples of the plugin's filter() on cfg 2 with pinned arrays --
A = upload, passes, download; B = hidden transfers (head 64 / tail 128); D = only the download hidden (head 0).
Prints wall-clock ms per sample and, for B and D, the CUDA-event stamps of the phases. One JSON object."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from flowdenoising_b200 import flowdenoising as fd                      # noqa: E402
from flowdenoising_b200.engine import gaussian_kernel                   # noqa: E402
from bench import synthetic_volume_torch                                # noqa: E402

Z, Y, X = 512, 1024, 1024
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
d_vol = synthetic_volume_torch((Z, Y, X), dev, seed=1)
host = torch.empty((Z, Y, X), dtype=torch.float32, pin_memory=True)
host2 = torch.empty((Z, Y, X), dtype=torch.float32, pin_memory=True)
k = gaussian_kernel(2.0)
ks = [k, k, k]
plans = {"A": None, "B": (64, 128), "D": (0, 128)}
out = {v: [] for v in plans}
traces = {v: [] for v in plans}
cur = {"plan": None}
fd.GaussianDenoising._overlap_plan = lambda self, torch_, ks_: cur["plan"]


def once(variant, record=True):
    host.copy_(d_vol); torch.cuda.synchronize()
    obj = fd.FlowDenoising(1, host.numpy())
    obj.filtered_vol = host2.numpy()
    cur["plan"] = plans[variant]
    fd._TRACE = [] if plans[variant] else None
    time.sleep(0.5)                       # the device idles before a user's call
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    obj.filter(ks)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    if record:
        out[variant].append(round(dt, 1))
        if fd._TRACE:
            tr = fd._TRACE
            traces[variant].append({lab: round(tr[0][1].elapsed_time(e), 1) for lab, e, _ in tr[1:]})
    fd._TRACE = None


once("A", False); once("B", False)
for _ in range(rounds):
    for v in ("A", "B", "D"):
        once(v)
print(json.dumps({"ms": out, "traces": traces}))
