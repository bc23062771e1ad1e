"""Development tool (companion of overlap_lab.py): why do the first slices of the Z pass run slower inside filter()?
Times the 64-slice head window of cfg 2 (a) back to back, (b) after the device idled for a second, (c) while a 2 GiB
host-to-device copy runs on another stream, (d) while a device-to-host copy runs. One JSON object."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from flowdenoising_b200.engine import DeviceEngine, FlowParams, gaussian_kernel  # noqa: E402
from flowdenoising_b200._lib import View                                        # noqa: E402

Z, Y, X = 512, 1024, 1024
eng = DeviceEngine()
g = torch.Generator(device="cuda"); g.manual_seed(1)
d_vol = torch.rand((Z, Y, X), device="cuda", generator=g) * 200 + torch.randn((Z, Y, X), device="cuda", generator=g) * 10
other = torch.empty_like(d_vol)
host = torch.empty((Z, Y, X), dtype=torch.float32, pin_memory=True)
a = torch.empty_like(d_vol)
k = gaussian_kernel(2.0)
flow = FlowParams()
head = 64
v = View(Z, head, 0, 1, Y, X, Y * X, X, Y * X, X)
side = torch.cuda.Stream()


def run(pre=None):
    torch.cuda.synchronize()
    if pre:
        pre()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.filter_view(d_vol, a, v, k, flow)
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1), 2)


def h2d():
    with torch.cuda.stream(side):
        other.copy_(host, non_blocking=True)


def d2h():
    with torch.cuda.stream(side):
        host.copy_(other, non_blocking=True)


out = {}
run(); run()
out["back_to_back"] = [run() for _ in range(3)]
out["after_1s_idle"] = [run(lambda: time.sleep(1.0)) for _ in range(3)]
out["after_0.2s_idle"] = [run(lambda: time.sleep(0.2)) for _ in range(2)]
out["with_h2d"] = [run(h2d) for _ in range(3)]
out["with_d2h"] = [run(d2h) for _ in range(3)]
out["idle_then_h2d"] = [run(lambda: (time.sleep(1.0), h2d())) for _ in range(2)]
print(json.dumps(out))
