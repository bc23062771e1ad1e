"""GPU tests of the transfer-hiding plumbing of the plugin's in-core filter(): windows of a periodic view (a pass run
as two device calls), the pitched copy of the C ABI, and filter() itself with the first Z slices / last X columns
split off -- every variant must give the bits of the plain upload, passes, download sequence, which the other GPU
tests pin against the reference (src/flowdenoising.py:285-290, :306-373)."""
import ctypes as C

import numpy as np
import pytest

from oracle import fd_oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    from flowdenoising_b200.engine import DeviceEngine
    return DeviceEngine()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


@pytest.mark.parametrize("use_of", [True, False])
@pytest.mark.parametrize("shape,first,count", [((20, 64, 96), 0, 6), ((20, 64, 96), 6, 14), ((20, 64, 96), 17, 3),
                                               ((5, 48, 64), 3, 2)])
def test_window_of_a_periodic_view_equals_the_whole_pass(eng, use_of, shape, first, count):
    """View(n_in, n_out < n_in, halo = first output slice, periodic): the same slices as the whole periodic pass,
    also when the window's neighbours wrap (first - r < 0, first + count + r > n_in, n_in < r)."""
    from flowdenoising_b200.engine import FlowParams
    from flowdenoising_b200._lib import View
    vol = O.synthetic_volume(shape, seed=7, noise_sigma=8.0)
    k = O.get_gaussian_kernel(1.5)     # r = 6
    p = FlowParams() if use_of else None
    Z, Y, X = shape
    d_in = dev(vol)
    full = torch.empty_like(d_in)
    eng.filter_along_axis(d_in, full, 0, k, p)
    out = torch.full_like(d_in, -7.0)
    eng.filter_view(d_in, out[first:], View(Z, count, first, 1, Y, X, Y * X, X, Y * X, X), k, p)
    assert torch.equal(out[first:first + count], full[first:first + count])
    assert bool((out[:first] == -7.0).all()) and bool((out[first + count:] == -7.0).all())   # nothing else written
    if use_of:      # the window with a bounded workspace (chunks inside the window)
        out2 = torch.empty_like(d_in)
        eng.filter_view(d_in, out2[first:], View(Z, count, first, 1, Y, X, Y * X, X, Y * X, X), k, p, chunk=2)
        assert torch.equal(out2[first:first + count], full[first:first + count])
    # a window along the middle axis of a [Z][X][Y] volume: the X pass of filter()
    vt = d_in.transpose(1, 2).contiguous()            # [Z][X][Y], slices along X
    A = X
    f = min(first, A - 1)
    fullx = torch.empty_like(vt)
    eng.filter_view(vt, fullx, View(A, A, 0, 1, Z, Y, Y, A * Y, Y, A * Y), k, p)
    outx = torch.full_like(vt, -7.0)
    eng.filter_view(vt, outx.view(-1)[f * Y:], View(A, count, f, 1, Z, Y, Y, A * Y, Y, A * Y), k, p)
    assert torch.equal(outx[:, f:f + count], fullx[:, f:f + count])
    assert bool((outx[:, :f] == -7.0).all()) and bool((outx[:, f + count:] == -7.0).all())


def test_periodic_view_arguments_are_checked(eng):
    from flowdenoising_b200.engine import FlowParams
    from flowdenoising_b200._lib import View
    d = torch.zeros((8, 32, 32), dtype=torch.float32, device="cuda")
    o = torch.zeros_like(d)
    k = O.get_gaussian_kernel(0.5)
    for bad in (View(8, 9, 0, 1, 32, 32, 1024, 32, 1024, 32), View(8, 4, 8, 1, 32, 32, 1024, 32, 1024, 32),
                View(8, 4, -1, 1, 32, 32, 1024, 32, 1024, 32)):
        for p in (FlowParams(), None):
            with pytest.raises(ValueError):
                eng.filter_view(d, o, bad, k, p)


def test_copy2d_async_moves_a_range_of_columns():
    from flowdenoising_b200 import _lib
    lib = _lib.load()
    Z, Y, X = 3, 5, 40
    t = torch.arange(Z * Y * X, dtype=torch.float32, device="cuda").view(Z, Y, X)
    host = torch.full((Z, Y, X), -1.0, dtype=torch.float32, pin_memory=True)
    st = torch.cuda.current_stream().cuda_stream
    x0, x1 = 7, 29
    _lib.check(lib.fdn_copy2d_async(host.data_ptr() + 4 * x0, 4 * X, t.data_ptr() + 4 * x0, 4 * X, 4 * (x1 - x0),
                                    Z * Y, 1, st))
    torch.cuda.synchronize()
    ref = torch.full((Z, Y, X), -1.0)
    ref[:, :, x0:x1] = t[:, :, x0:x1].cpu()
    assert torch.equal(host, ref)
    back = torch.zeros_like(t)
    _lib.check(lib.fdn_copy2d_async(back.data_ptr(), 4 * X, host.data_ptr(), 4 * X, 4 * X, Z * Y, 0, st))   # H2D
    torch.cuda.synchronize()
    assert torch.equal(back.cpu(), host)
    with pytest.raises(ValueError):
        _lib.check(lib.fdn_copy2d_async(back.data_ptr(), 4, host.data_ptr(), 4 * X, 4 * X, Z * Y, 0, st))   # pitch < width


def _run_filter(fd, vol, kernels, lw, pinned_out=False):
    obj = fd.FlowDenoising(2, vol, lw[0], lw[1], fd.get_flow_with_prev_flow, fd.warp_slice)
    if pinned_out:
        obj.filtered_vol = torch.empty(vol.shape, dtype=torch.float32, pin_memory=True).numpy()
    res = obj.filter(kernels)
    return np.array(res), np.array(obj.vol), obj.progress


@pytest.mark.parametrize("host", ["pageable", "pinned", "uint8"])
@pytest.mark.parametrize("shape", [(24, 72, 80), (10, 40, 52)])
def test_filter_with_hidden_transfers_equals_plain_filter(monkeypatch, host, shape):
    """FlowDenoising.filter() with the Z pass started on its first slices during the upload and the last X columns
    computed during the download (forced on a toy volume) == the plain sequence, for pinned float32 arrays (direct
    DMA), ordinary arrays (staging threads) and an integer volume (cast on the way, quirk Q3)."""
    from flowdenoising_b200 import flowdenoising as fd
    Z, Y, X = shape
    base = O.synthetic_volume(shape, seed=3, noise_sigma=10.0)
    kernels = [O.get_gaussian_kernel(s) for s in (1.0, 0.75, 0.5)]      # r = 4, 3, 2
    lw = (2, 5)

    def make():
        if host == "uint8":
            return np.clip(base, 0, 255).astype(np.uint8)
        if host == "pinned":
            t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            t.copy_(torch.from_numpy(base))
            return t.numpy()
        return base.copy()

    monkeypatch.setattr(fd, "_OVERLAP_MIN_BYTES", 1 << 62)
    ref, ref_zy, _ = _run_filter(fd, make(), kernels, lw, pinned_out=(host == "pinned"))
    monkeypatch.setattr(fd, "_OVERLAP_MIN_BYTES", 0)
    plans = []
    orig = fd.GaussianDenoising._filter_overlapped

    def spy(self, eng, ks, flow, head, tail):
        plans.append((head, tail))
        return orig(self, eng, ks, flow, head, tail)
    monkeypatch.setattr(fd.GaussianDenoising, "_filter_overlapped", spy)
    got, got_zy, progress = _run_filter(fd, make(), kernels, lw, pinned_out=(host == "pinned"))
    assert plans and 1 <= plans[0][0] < Z and 1 <= plans[0][1] < X          # the split path did run
    assert got.dtype == ref.dtype and np.array_equal(got, ref)
    assert np.array_equal(got_zy, ref_zy)
    assert progress == Z + Y + X
