"""Bounds and race checks of our own (compute-sanitizer is closed on this GPU pool, profiles/r2_sanitizer.txt):
guard bands around every buffer the kernels write, and bit-repeatability of the inter-block protocols."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle import fd_oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

GUARD = 1 << 20
PATTERN = 0xA5


@pytest.fixture(scope="module")
def eng():
    from flowdenoising_b200.engine import DeviceEngine
    return DeviceEngine()


class Arena:
    """One device allocation; sub-buffers separated by guard bands filled with a byte pattern."""

    def __init__(self, sizes):
        self.offsets = []
        off = GUARD
        for s in sizes:
            s = (int(s) + 255) // 256 * 256
            self.offsets.append((off, s))
            off += s + GUARD
        self.buf = torch.full((off,), PATTERN, dtype=torch.uint8, device="cuda")

    def view(self, i, dtype=torch.float32):
        off, s = self.offsets[i]
        return self.buf[off:off + s].view(dtype)

    def check(self):
        pos = 0
        for off, s in self.offsets:
            assert bool((self.buf[pos:off] == PATTERN).all()), f"guard band before offset {off} was written"
            pos = off + s
        assert bool((self.buf[pos:] == PATTERN).all()), "guard band at the end was written"


def images(shape, n, seed):
    v = O.synthetic_volume((n,) + tuple(shape), seed=seed, noise_sigma=6.0)
    return v.astype(np.float32)


@pytest.mark.parametrize("shape,n,cap", [((64, 1024), 6, 6), ((40, 300), 3, 8), ((33, 128), 2, 2), ((96, 500), 3, 5),
                                         ((17, 96), 4, 4)])
@pytest.mark.parametrize("win", [5, 9])
def test_flow_iterations_stay_in_bounds_and_repeat(eng, shape, n, cap, win):
    lib = eng.lib
    h, w = shape
    imgs = images(shape, n + 1, 3)
    rf = lib.fdn_polyexp_floats(h, w)
    nscr = lib.fdn_flow_iteration_scratch_bytes(cap, h, w)           # scratch sized for `cap` >= n pairs
    fl_bytes = 8 * n * h * w
    ar = Arena([4 * rf * (n + 1), fl_bytes, fl_bytes, fl_bytes, nscr, 4 * (n + 1) * h * w])
    R, f0, f1, f2, scr, d_img = ar.view(0), ar.view(1), ar.view(2), ar.view(3), ar.view(4, torch.uint8), ar.view(5)
    d_img[:(n + 1) * h * w].copy_(torch.from_numpy(imgs).cuda().view(-1))
    assert lib.fdn_polyexp(d_img.data_ptr(), n + 1, h, w, 5, 1.2, R.data_ptr(), None) == 0, lib.fdn_last_error()
    scr.zero_()
    rng = np.random.default_rng(8)
    flow = torch.from_numpy((rng.standard_normal((n, h, w, 2)) * 3).astype(np.float32)).cuda()
    digests = set()
    for rep in range(12):
        f0[:2 * n * h * w].copy_(flow.view(-1))
        res = C.c_void_p()
        rc = lib.fdn_flow_iterations(R.data_ptr(), R.data_ptr() + 4 * rf, f0.data_ptr(), f1.data_ptr(), f2.data_ptr(), n,
                                     h, w, win, 3, scr.data_ptr(), nscr, None, C.byref(res))
        assert rc == 0, lib.fdn_last_error()
        out = {f0.data_ptr(): f0, f1.data_ptr(): f1, f2.data_ptr(): f2}[res.value]
        digests.add(hashlib.sha256(out[:2 * n * h * w].cpu().numpy().tobytes()).hexdigest())
    ar.check()
    assert len(digests) == 1, "the flow iteration is not repeatable (a race between blocks?)"
    if win == 5:      # ... and equal to the strip kernel's result
        lib.fdn_set_flow_iter_variant(0)
        try:
            f0[:2 * n * h * w].copy_(flow.view(-1))
            res = C.c_void_p()
            assert lib.fdn_flow_iterations(R.data_ptr(), R.data_ptr() + 4 * rf, f0.data_ptr(), f1.data_ptr(),
                                           f2.data_ptr(), n, h, w, win, 3, scr.data_ptr(), nscr, None, C.byref(res)) == 0
            out = {f0.data_ptr(): f0, f1.data_ptr(): f1, f2.data_ptr(): f2}[res.value]
            assert hashlib.sha256(out[:2 * n * h * w].cpu().numpy().tobytes()).hexdigest() in digests
        finally:
            lib.fdn_set_flow_iter_variant(1)
        ar.check()


@pytest.mark.parametrize("shape,chunk", [((11, 64, 96), 4), ((9, 40, 500), 9), ((7, 70, 244), 3)])
def test_pass_stays_in_bounds(eng, shape, chunk):
    """A whole OF pass with an exactly sized workspace between guard bands (chunked, last chunk shorter), the output
    volume guarded as well; repeated passes give the same bits."""
    from flowdenoising_b200._lib import OfParams, View
    lib = eng.lib
    vol = O.synthetic_volume(shape, seed=6, noise_sigma=8.0)
    Z, Y, X = shape
    k = np.ascontiguousarray(O.get_gaussian_kernel(1.0))
    kp = k.ctypes.data_as(C.POINTER(C.c_double))
    of = OfParams(3, 5, 3, 5, 1.2, 1)
    for axis in (0, 1):
        v = View(Z, Z, 0, 1, Y, X, Y * X, X, Y * X, X) if axis == 0 else View(Y, Y, 0, 1, Z, X, X, Y * X, X, Y * X)
        need = lib.fdn_workspace_bytes(C.byref(v), k.size, C.byref(of), chunk)
        assert need > 0
        ar = Arena([4 * vol.size, 4 * vol.size, need])
        d_in, d_out, ws = ar.view(0), ar.view(1), ar.view(2, torch.uint8)
        d_in[:vol.size].copy_(torch.from_numpy(vol).cuda().view(-1))
        digests = set()
        for rep in range(3):
            rc = lib.fdn_filter_axis(d_in.data_ptr(), d_out.data_ptr(), C.byref(v), kp, k.size, C.byref(of), chunk,
                                     ws.data_ptr(), need, None)
            assert rc == 0, lib.fdn_last_error()
            digests.add(hashlib.sha256(d_out[:vol.size].cpu().numpy().tobytes()).hexdigest())
        ar.check()
        assert len(digests) == 1
        ref = O.flow_axis_c(vol, axis, k)
        assert np.array_equal(d_out[:vol.size].cpu().numpy().reshape(shape), ref)
