"""Out-of-core streaming (flowdenoising_b200/streaming.py): small forced slabs must reproduce the in-core passes bit
for bit; the mean-padding border must equal the periodic filter of a mean-padded volume (what the reference's
sequential variant computes, src/flowdenoising_sequential.py:88-89)."""
import numpy as np
import pytest

from oracle import fd_oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    from flowdenoising_b200.engine import DeviceEngine
    return DeviceEngine()


def in_core(eng, vol, kernels, flow):
    zy, zyx = eng.filter(torch.from_numpy(vol).cuda(), kernels, flow)
    return zy.cpu().numpy(), zyx.cpu().numpy()


@pytest.mark.parametrize("use_of", [False, True])
def test_streaming_equals_in_core(eng, tmp_path, use_of):
    from flowdenoising_b200.engine import FlowParams
    from flowdenoising_b200.streaming import StreamingDenoiser
    vol = O.synthetic_volume((21, 40, 48), seed=11, noise_sigma=8.0)
    kernels = [O.get_gaussian_kernel(s) for s in (1.0, 0.5, 1.0)]
    flow = FlowParams() if use_of else None
    zy_ref, zyx_ref = in_core(eng, vol, kernels, flow)
    # the volume as a read-only memory map, slabs of 5 output slices (ragged last slab, halo wraps around both ends)
    path = tmp_path / "vol.f32"
    vol.tofile(path)
    mm = np.memmap(path, dtype=np.float32, mode="r", shape=vol.shape)
    sd = StreamingDenoiser(flow, slab_slices=5)
    zy, zyx = sd.filter(mm, kernels)
    assert np.array_equal(zy, zy_ref) and np.array_equal(zyx, zyx_ref)
    # in-place semantics of the reference on a writable array; slabs larger than the axis
    v2 = vol.copy()
    out = np.empty_like(v2)
    sd2 = StreamingDenoiser(flow, slab_slices=64)
    zy2, zyx2 = sd2.filter(v2, kernels, filtered_vol=out)
    assert zy2 is v2 and zyx2 is out
    assert np.array_equal(v2, zy_ref) and np.array_equal(out, zyx_ref)
    # uint8 source (cast while gathering)
    v8 = vol.astype(np.uint8)
    zy8, zyx8 = StreamingDenoiser(flow, slab_slices=7).filter(v8, kernels)
    a8, b8 = in_core(eng, v8.astype(np.float32), kernels, flow)
    assert np.array_equal(zy8, a8) and np.array_equal(zyx8, b8)


def test_streaming_mean_border(eng):
    from flowdenoising_b200.engine import FlowParams
    from flowdenoising_b200.streaming import StreamingDenoiser
    vol = O.synthetic_volume((14, 36, 40), seed=4, noise_sigma=6.0)
    k = O.get_gaussian_kernel(1.0)
    r = k.size // 2
    mean = np.float32(vol.mean())
    for flow in (None, FlowParams()):
        sd = StreamingDenoiser(flow, border="mean", slab_slices=4)
        for axis in range(3):
            out = np.empty_like(vol)
            sd.filter_axis(vol, out, axis, k, mean=float(mean))
            pad = [(0, 0)] * 3
            pad[axis] = (r, r)
            padded = np.pad(vol, pad, constant_values=mean)
            ref = torch.empty(padded.shape, dtype=torch.float32, device="cuda")
            eng.filter_along_axis(torch.from_numpy(padded).cuda(), ref, axis, k, flow)
            sl = [slice(None)] * 3
            sl[axis] = slice(r, r + vol.shape[axis])
            assert np.array_equal(out, ref.cpu().numpy()[tuple(sl)]), (axis, flow is not None)


def test_streaming_two_lanes_same_device(eng):
    """Two lanes (here: the same device twice) deal the slabs round-robin; results do not depend on the dealing."""
    from flowdenoising_b200.engine import FlowParams
    from flowdenoising_b200.streaming import StreamingDenoiser
    vol = O.synthetic_volume((16, 32, 40), seed=2, noise_sigma=6.0)
    kernels = [O.get_gaussian_kernel(1.0)] * 3
    zy_ref, zyx_ref = in_core(eng, vol, kernels, FlowParams())
    sd = StreamingDenoiser(FlowParams(), slab_slices=3, devices=[0, 0])
    zy, zyx = sd.filter(vol.copy(), kernels)
    assert np.array_equal(zy, zy_ref) and np.array_equal(zyx, zyx_ref)
