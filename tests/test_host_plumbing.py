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
