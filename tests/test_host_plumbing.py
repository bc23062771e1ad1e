"""CPU tests of the host-side plumbing of the plugin's filter() that needs no device: when the transfer-hiding
sequence is chosen and with which windows, and the page-mapping helper for fresh result arrays."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from flowdenoising_b200 import flowdenoising as fd  # noqa: E402
from flowdenoising_b200.engine import FlowParams  # noqa: E402


class _Obj(fd.GaussianDenoising):
    """The decision logic only: no engine, no device."""

    def __init__(self, shape, dtype=np.float32, of=True):
        self.vol = np.empty(shape, dtype)
        self.filtered_vol = np.empty(shape, dtype)
        self._flow_params = FlowParams() if of else None


def test_overlap_plan_decisions(monkeypatch):
    ks = [np.ones(17)] * 3
    # ordinary (pageable) arrays: a quarter of the axis on both ends
    assert _Obj((512, 1024, 1024))._overlap_plan(torch, ks) == (128, 256)
    assert _Obj((256, 2048, 2048), np.uint8)._overlap_plan(torch, ks) == (64, 512)
    # small volumes, the no-OF path and planes larger than a staging buffer keep the plain sequence
    assert _Obj((64, 256, 256))._overlap_plan(torch, ks) is None
    assert _Obj((512, 1024, 1024), of=False)._overlap_plan(torch, ks) is None
    assert _Obj((8, 4100, 4100))._overlap_plan(torch, ks) is None
    # page-locked float32 arrays move at PCIe speed: about 256 MB of slices / 128 columns are enough
    monkeypatch.setattr(fd, "_is_pinned_f32", lambda a, t: True)
    assert _Obj((512, 1024, 1024))._overlap_plan(torch, ks) == (64, 128)
    assert _Obj((256, 2048, 2048))._overlap_plan(torch, ks) == (16, 128)      # 16 MB slices: 256 MB of them
    assert _Obj((40, 4000, 4000))._overlap_plan(torch, ks) == (8, 128)         # never fewer than 8 ...
    assert _Obj((12, 4000, 4000))._overlap_plan(torch, ks) == (6, 128)         # ... or more than half the axis
    monkeypatch.setattr(fd, "_OVERLAP_MIN_BYTES", 0)
    assert _Obj((1, 64, 64))._overlap_plan(torch, ks) is None          # nothing to split
    assert _Obj((2, 64, 2))._overlap_plan(torch, ks) == (1, 1)


@pytest.mark.parametrize("dtype", [np.float32, np.uint8, np.int16])
def test_prefault_leaves_the_values_alone(dtype):
    rng = np.random.default_rng(0)
    a = (rng.random((7, 33, 1031)) * 200).astype(dtype)
    if dtype == np.float32:
        a[0, 0, :4] = [-0.0, np.inf, -np.inf, np.nan]
    ref = a.copy()
    th = fd._prefault(a)
    assert th is not None
    th.join()
    assert a.tobytes() == ref.tobytes()
    assert fd._prefault(a[:, :, ::2]) is None          # not contiguous: left alone
    ro = a.copy(); ro.setflags(write=False)
    assert fd._prefault(ro) is None


@pytest.mark.parametrize("shape,src_t,dst_t", [((16, 512, 1024), np.float32, np.float32), ((16, 512, 1024), np.uint8, np.float32),
                                               ((9, 700, 900), np.float32, np.uint8), ((1, 4096, 1024), np.float32, np.float32),
                                               ((3, 7, 5), np.float32, np.int16)])
def test_threaded_staging_copy_equals_copyto(shape, src_t, dst_t):
    """_pcopy = np.copyto(..., casting="unsafe"), whatever the split over the host threads (pieces along axis 0)."""
    rng = np.random.default_rng(5)
    src = (rng.random(shape) * 250).astype(src_t)
    got = np.zeros(shape, dst_t)
    ref = np.zeros(shape, dst_t)
    fd._pcopy(got, src)
    np.copyto(ref, src, casting="unsafe")
    assert got.tobytes() == ref.tobytes()
    # a range of columns of a larger array (the column downloads), neighbours untouched
    big = np.full(shape[:2] + (shape[2] + 9,), 7, dst_t)
    fd._pcopy(big[:, :, 4:4 + shape[2]], src)
    assert np.array_equal(big[:, :, 4:4 + shape[2]], ref) and (big[:, :, :4] == 7).all() and (big[:, :, -5:] == 7).all()


def test_pick_chunk_asks_the_driver_only_when_the_held_workspace_is_too_small():
    """DeviceEngine._pick_chunk: whole view when the workspace already held is large enough (no cudaMemGetInfo: it
    blocks while a large copy is in flight), otherwise sized from the free memory, by bisection when that is short."""
    from types import SimpleNamespace
    from flowdenoising_b200.engine import DeviceEngine
    from flowdenoising_b200._lib import View, OfParams
    calls = []

    def mem_get_info(_dev):
        calls.append(1)
        return free[0], 1 << 40
    free = [10_000]
    eng = object.__new__(DeviceEngine)
    eng.torch = SimpleNamespace(cuda=SimpleNamespace(mem_get_info=mem_get_info))
    eng.lib = SimpleNamespace(fdn_workspace_bytes=lambda v, klen, ofp, c: 100 * (c + 16))   # slices + 2r of cache
    eng.device = None
    eng.workspace_limit_bytes = None
    eng._ws = SimpleNamespace(numel=lambda: 5_000)
    v = View(64, 32, 0, 1, 8, 8, 64, 8, 64, 8)
    ofp = OfParams(3, 5, 3, 5, 1.2, 1)
    assert eng._pick_chunk(v, 17, ofp) == 32 and not calls             # 4800 bytes fit the 5000 held
    eng._ws = SimpleNamespace(numel=lambda: 1_000)
    assert eng._pick_chunk(v, 17, ofp) == 32 and len(calls) == 1       # 4800 <= 0.85 * (10000 + 1000)
    free[0] = 3_000
    assert eng._pick_chunk(v, 17, ofp) == 18 and len(calls) == 2       # 100 * (c + 16) <= 3400 -> c = 18
    eng._ws = None
    free[0] = 1_000
    with pytest.raises(Exception):
        eng._pick_chunk(v, 17, ofp)                                    # not even one slice
    eng.workspace_limit_bytes = 2_000                                  # an explicit limit never asks the driver
    n = len(calls)
    assert eng._pick_chunk(v, 17, ofp) == 4 and len(calls) == n


class _FakeCuda:
    """Stream / event stand-ins: the host-side sequencing of _Upload on CPU tensors (copies are synchronous here)."""

    class Event:
        def __init__(self, enable_timing=False):
            self.recorded = False

        def record(self, stream=None):
            self.recorded = True

        def synchronize(self):
            assert self.recorded

    class Stream:
        cuda_stream = 0

        def __init__(self, device=None):
            self.waited = []

        def wait_stream(self, other):
            self.waited.append(other)

        def wait_event(self, e):
            assert e.recorded, "the compute stream was told to wait for an event that was never recorded"

        def synchronize(self):
            pass

    _current = Stream()

    @classmethod
    def current_stream(cls):
        return cls._current

    class _Ctx:
        def __init__(self, *_a):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False
    stream = _Ctx
    device = _Ctx


class _FakeTorch:
    """torch with a stand-in `cuda` namespace (everything else is the real module)."""
    cuda = _FakeCuda

    def __getattr__(self, name):
        return getattr(torch, name)


@pytest.mark.parametrize("dtype", [np.float32, np.uint8])
def test_upload_pieces_arrive_in_order_and_only_after_start(monkeypatch, dtype):
    """_Upload: nothing moves before start(); ready(k) returns once pieces 0..k are in place; the pieces of the
    transfer-hiding order {far end, head, rest} together are the whole volume (cast to float32 on the way)."""
    ft = _FakeTorch()
    monkeypatch.setattr(fd, "_pinned_pair", lambda t: [torch.empty(fd._STAGE_BYTES // 4, dtype=torch.float32)
                                                       for _ in range(2)])
    rng = np.random.default_rng(2)
    Z, Y, X = 40, 64, 96
    vol = (rng.random((Z, Y, X)) * 255).astype(dtype)
    r, head = 4, 10
    ranges = [(Z - r, Z), (0, head + r), (head + r, Z - r)]
    up = fd._Upload(vol, ft, "cpu", ranges)
    up.d.fill_(-1.0)
    assert up.thread is None and not any(f.is_set() for f in up.flags)
    assert float(up.d.max()) == -1.0                                   # allocated, not started
    up.start()
    up.ready(1)                                                         # far end + head
    got = up.d.numpy()
    assert np.array_equal(got[Z - r:], vol[Z - r:].astype(np.float32))
    assert np.array_equal(got[:head + r], vol[:head + r].astype(np.float32))
    up.ready(2)
    up.close()
    assert np.array_equal(up.d.numpy(), vol.astype(np.float32))
    # a failing source surfaces on the calling thread instead of hanging ready()
    class Bad:
        shape = (Z, Y, X)

        def __getitem__(self, k):
            raise IOError("unreadable")
    up2 = fd._Upload(Bad(), ft, "cpu", [(0, Z)])
    up2.start()
    with pytest.raises(IOError):
        up2.ready(0)


def test_upload_of_a_pinned_array_is_issued_by_start(monkeypatch):
    """The direct (page-locked float32) route of _Upload: same contract, no host thread."""
    ft = _FakeTorch()
    monkeypatch.setattr(fd, "_is_pinned_f32", lambda a, t: True)
    Z, Y, X = 12, 8, 16
    vol = np.random.default_rng(4).random((Z, Y, X)).astype(np.float32)
    up = fd._Upload(vol, ft, "cpu", [(10, 12), (0, 5), (5, 10)])
    up.d.fill_(-1.0)
    assert not any(e.recorded for e in up.events)
    up.start()
    assert up.thread is None and all(e.recorded for e in up.events) and all(f.is_set() for f in up.flags)
    assert up.stream.waited == [ft.cuda.current_stream()]              # ordered after earlier users of the memory
    up.ready(2)
    up.close()
    assert np.array_equal(up.d.numpy(), vol)


# ---- filter() with hidden transfers, host logic only: an oracle-backed stand-in for the device engine (same method
# contracts as DeviceEngine, CPU tensors, include/fdn_b200.h view semantics incl. windows of a periodic view) ----
class _CpuEngine:
    def __init__(self):
        self.torch = _FakeTorch()
        self.device = torch.device("cpu")
        self.views = []

    @staticmethod
    def _flat(t):
        assert t.is_contiguous()
        return t.view(-1).numpy()

    def reserve_workspace(self, view, klen, flow):
        pass

    def filter_view(self, d_in, d_out, v, kernel, flow, chunk=None, exact=True):
        from oracle import fd_oracle as O
        self.views.append((v.n_in, v.n_out, v.halo, v.periodic))
        src = np.lib.stride_tricks.as_strided(self._flat(d_in), (v.n_in, v.H, v.W),
                                              (4 * v.in_slice_stride, 4 * v.in_row_stride, 4))
        dst = np.lib.stride_tricks.as_strided(self._flat(d_out), (v.n_out, v.H, v.W),
                                              (4 * v.out_slice_stride, 4 * v.out_row_stride, 4))
        assert v.periodic and 0 <= v.halo < v.n_in and v.n_out <= v.n_in
        o = O.OracleDenoiser(1, np.ascontiguousarray(src), use_OF=flow is not None, backend="c",
                             **({} if flow is None else dict(l=flow.levels, w=flow.winsize)))
        for s in range(v.n_out):
            idx = (s + v.halo) % v.n_in
            o.filter_slice(0, idx, kernel)          # the oracle wraps like the reference (% shape)
            dst[s] = o.filtered_vol[idx]

    def filter_along_axis(self, vol, out, axis, kernel, flow, chunk=None, exact=True, scratch=None):
        from flowdenoising_b200._lib import View
        Z, Y, X = vol.shape
        assert axis == 1
        self.filter_view(vol, out, View(Y, Y, 0, 1, Z, X, X, Y * X, X, Y * X), kernel, flow)

    def transpose_yx(self, src, dst=None):
        n, A, B = src.shape
        if dst is None:
            dst = torch.empty((n, B, A), dtype=torch.float32)
        dst.copy_(src.transpose(1, 2))
        return dst

    def transpose_strided(self, src, src_off, in_sn, in_sa, dst, dst_off, out_sn, out_sb, n, A, B):
        s, d = self._flat(src), self._flat(dst)
        i = np.arange(n)[:, None, None]
        a = np.arange(A)[None, :, None]
        b = np.arange(B)[None, None, :]
        d[dst_off + i * out_sn + b * out_sb + a] = s[src_off + i * in_sn + a * in_sa + b]


class _CpuDownload:
    """Stand-in for _Download / _DownloadCols (device -> caller's array): inline copies."""

    def __init__(self, t, dst, *cols_and_torch):
        if len(cols_and_torch) == 3:
            x0, x1, _t = cols_and_torch
            np.copyto(dst[:, :, x0:x1], t.numpy()[:, :, x0:x1], casting="unsafe")
        else:
            np.copyto(dst, t.numpy(), casting="unsafe")

    def wait(self):
        pass


@pytest.mark.parametrize("shape,head,tail,dtype", [((12, 24, 32), 3, 5, np.float32), ((12, 24, 32), 0, 7, np.float32),
                                                   ((7, 24, 32), 3, 1, np.float32), ((12, 24, 32), 4, 8, np.uint8)])
def test_filter_with_hidden_transfers_host_logic_vs_oracle(monkeypatch, shape, head, tail, dtype):
    """_filter_overlapped on the stand-in engine == the oracle's filter(): order of the upload pieces, the windows of
    the Z and the X pass (offsets into the output, wrap-around neighbours), the buffer roles (vol <- Z+Y,
    filtered_vol <- Z+Y+X, src/flowdenoising.py:285-290), progress."""
    from oracle import fd_oracle as O
    monkeypatch.setattr(fd, "_pinned_pair", lambda t: [torch.empty(fd._STAGE_BYTES // 4, dtype=torch.float32)
                                                       for _ in range(2)])
    monkeypatch.setattr(fd, "_Download", _CpuDownload)
    monkeypatch.setattr(fd, "_DownloadCols", _CpuDownload)
    monkeypatch.setattr(fd.GaussianDenoising, "_begin_device_call", lambda self: None)
    monkeypatch.setattr(fd.GaussianDenoising, "_end_device_call",
                        lambda self, n: setattr(self, "_progress_done", self._progress_done + n))
    base = O.synthetic_volume(shape, seed=9, noise_sigma=6.0)
    vol = np.clip(base, 0, 255).astype(dtype)
    kernels = [O.get_gaussian_kernel(0.5), O.get_gaussian_kernel(0.5), O.get_gaussian_kernel(0.75)]   # r = 2, 2, 3
    ref = O.OracleDenoiser(1, vol.astype(np.float32), use_OF=True, l=2, w=5, backend="c")
    ref_zyx = ref.filter(kernels)
    obj = fd.FlowDenoising(1, vol.copy(), 2, 5)
    eng = _CpuEngine()
    res = obj._filter_overlapped(eng, [np.asarray(k, np.float64) for k in kernels], obj._flow(), head, tail)
    Z, Y, X = shape
    assert res is obj.filtered_vol and obj.progress == Z + Y + X
    assert np.array_equal(res, ref_zyx.astype(dtype)) and np.array_equal(obj.vol, ref.vol.astype(dtype))
    split_z = head > 0 and head + 4 < Z
    want = ([(Z, head, 0, 1), (Z, Z - head, head, 1)] if split_z else [(Z, Z, 0, 1)]) + \
           [(Y, Y, 0, 1), (X, X - tail, 0, 1), (X, tail, X - tail, 1)]
    assert eng.views == want
