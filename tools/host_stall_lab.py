"""Development tool: which host call of filter()'s prologue can stall while a large pinned upload is in flight?
Times each call (ms) right after issuing a 1 GiB host-to-device copy on a side stream, 8 repetitions."""
import ctypes as C, json, os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from flowdenoising_b200 import _lib                                   # noqa: E402
from flowdenoising_b200._lib import View, OfParams                    # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
n = 256 << 20
host = torch.empty(n, dtype=torch.float32, pin_memory=True)
d = torch.empty(n, dtype=torch.float32, device=dev)
side = torch.cuda.Stream()
keep = [torch.empty(512 << 20, dtype=torch.float32, device=dev) for _ in range(3)]
del keep
v = View(512, 512, 0, 1, 1024, 1024, 1 << 20, 1024, 1 << 20, 1024)
ofp = OfParams(3, 5, 3, 5, 1.2, 1)
res = {}


def t(name, fn):
    t0 = time.perf_counter()
    r = fn()
    res.setdefault(name, []).append(round((time.perf_counter() - t0) * 1e3, 3))
    return r


for rep in range(8):
    torch.cuda.synchronize()
    time.sleep(0.2)
    with torch.cuda.stream(side):
        t("issue_h2d", lambda: d.copy_(host, non_blocking=True))
    t("mem_get_info", lambda: torch.cuda.mem_get_info(dev))
    a = t("empty_2GiB", lambda: torch.empty(512 << 20, dtype=torch.float32, device=dev))
    b = t("empty_2GiB_b", lambda: torch.empty(512 << 20, dtype=torch.float32, device=dev))
    t("workspace_bytes", lambda: lib.fdn_workspace_bytes(C.byref(v), 17, C.byref(ofp), 512))
    t("new_stream", lambda: torch.cuda.Stream())
    t("is_pinned", lambda: host.is_pinned())
    e = torch.cuda.Event()
    t("event_record_wait", lambda: (e.record(side), torch.cuda.current_stream().wait_event(e)))
    t("mem_get_info_2", lambda: torch.cuda.mem_get_info(dev))
    t("sync", torch.cuda.synchronize)
    del a, b
print(json.dumps(res))
