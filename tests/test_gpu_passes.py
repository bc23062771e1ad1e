"""GPU parity tests of whole passes (filter_along_Z/Y/X, filter) against the reference's golden outputs
(tests/golden, produced by the unmodified reference) and the CPU oracle.

North-star tolerances (BASELINE.json): OF path PSNR >= 50 dB and max|d| <= 1e-3 * data range; no-OF path <= 1-ulp
scale. What is asserted here is tighter: both paths are BIT-IDENTICAL to the reference's output (fixtures produced by
the unmodified reference with cv2 4.13.0 on an AVX2 x86 host). Anything less would not be stable: a last-ulp flow
difference that crosses a 1/32-px bin of cv2.remap's map quantiser moves a voxel by up to contrast/32 * tap weight
(SURVEY.md §7 "hard parts"), which is how a merely "close" Farneback breaks the 1e-3 bound on noisy data.
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle import fd_oracle as O
from conftest import psnr

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    from flowdenoising_b200.engine import DeviceEngine
    return DeviceEngine()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def report(name, got, ref, data_range=255.0):
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    p = psnr(got, ref, data_range)
    frac = float(np.mean(got == ref))
    print(f"{name}: PSNR {p:.1f} dB  max|d| {d.max():.3e} ({d.max() / data_range:.2e} of range)  "
          f"bit-equal {frac:.6f}  n(|d|>1e-3*range) {int((d > 1e-3 * data_range).sum())}")
    return p, d.max() / data_range, frac


def check_of(name, got, ref, data_range=255.0):
    p, dmax, frac = report(name, got, ref, data_range)
    assert p >= 50.0                       # north star: PSNR >= 50 dB
    assert dmax <= 1e-3                    # north star: max|d| <= 1e-3 of the data range
    assert frac == 1.0                     # this implementation: bit-identical to the reference


@pytest.mark.parametrize("name", ["toy_noof.npz", "toy_noof_float.npz"])
def test_noof_passes_bit_exact_vs_reference(eng, golden, name):
    g = golden(name)
    vol = g["vol"].astype(np.float32)
    kernels = [O.get_gaussian_kernel(float(s)) for s in g["sigmas"]]
    a = torch.empty_like(dev(vol)); b = torch.empty_like(a); c = torch.empty_like(a)
    eng.filter_along_axis(dev(vol), a, 0, kernels[0], None)
    assert np.array_equal(a.cpu().numpy(), g["Z"])
    eng.filter_along_axis(a, b, 1, kernels[1], None)
    assert np.array_equal(b.cpu().numpy(), g["ZY"])
    eng.filter_along_axis(b, c, 2, kernels[2], None)
    assert np.array_equal(c.cpu().numpy(), g["ZYX"])
    zy, zyx = eng.filter(dev(vol), kernels, None)
    assert np.array_equal(zy.cpu().numpy(), g["ZY"]) and np.array_equal(zyx.cpu().numpy(), g["ZYX"])
    # fast arithmetic (float32 FMA): <= 1-ulp-scale relative error (north star), here <= 4 ulp of the data range
    zy2, zyx2 = eng.filter(dev(vol), kernels, None, exact=False)
    rng_ = float(np.abs(g["ZYX"]).max())
    assert np.abs(zyx2.cpu().numpy() - g["ZYX"]).max() <= 4 * np.finfo(np.float32).eps * rng_


def test_noof_large_random(eng):
    rng = np.random.default_rng(0)
    vol = (rng.standard_normal((37, 50, 130)) * 100).astype(np.float32)
    for axis, sigma in [(0, 2.0), (1, 1.0), (2, 3.0)]:
        k = O.get_gaussian_kernel(sigma)
        out = torch.empty_like(dev(vol))
        eng.filter_along_axis(dev(vol), out, axis, k, None)
        assert np.array_equal(out.cpu().numpy(), O.gauss_axis_c(vol, axis, k)), axis


def numpy_axis_filter(vol, axis, k):
    """The reference's arithmetic (src/flowdenoising.py:133-158 under NumPy >= 2): float64 product and sum, rounded to
    float32 after every tap; periodic along the axis."""
    acc = np.zeros(vol.shape, np.float32)
    r = k.size // 2
    with np.errstate(all="ignore"):
        for i in range(k.size):
            acc = (acc.astype(np.float64) + np.roll(vol, r - i, axis=axis).astype(np.float64) * k[i]).astype(np.float32)
    return acc


def same_bits_or_nan(a, b):
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(nan | (a.view(np.uint32) == b.view(np.uint32))))


@pytest.mark.parametrize("shape", [(40, 24, 136), (19, 8, 512), (140, 6, 20)])
def test_noof_exact_special_values(eng, shape):
    """The exact mode rounds most taps with integer instructions, which is valid for zero / normal float32 sums only;
    anything else (subnormals, huge values, infinities, NaN, tiny values that cancel) must fall back to conversions.
    Covers aligned and unaligned widths (vector widths 1 / 2 / 4) and segments longer than one window period."""
    rng = np.random.default_rng(3)
    vol = (rng.standard_normal(shape) * 50).astype(np.float32)
    flat = vol.reshape(-1)
    n = flat.size
    pick = lambda m: rng.choice(n, m, replace=False)
    flat[pick(n // 7)] = 0.0
    flat[pick(n // 50)] = -0.0
    flat[pick(n // 40)] = np.float32(1e-42)          # subnormal
    flat[pick(n // 40)] = np.float32(-3e-39)
    flat[pick(n // 40)] = np.float32(2.5e-12)        # normal, below the integer path's range
    flat[pick(n // 40)] = np.float32(3e33)           # above it
    flat[pick(n // 60)] = np.float32(3e38)           # sums overflow to +-inf
    flat[pick(n // 60)] = np.float32(-3e38)
    flat[pick(n // 200)] = np.inf
    flat[pick(n // 200)] = np.nan
    for axis, sigma in [(0, 2.0), (1, 0.5), (2, 2.0), (0, 4.0), (2, 1.0)]:
        k = O.get_gaussian_kernel(sigma)
        out = torch.empty_like(dev(vol))
        eng.filter_along_axis(dev(vol), out, axis, k, None)
        assert same_bits_or_nan(out.cpu().numpy(), numpy_axis_filter(vol, axis, k)), (axis, sigma)
    # cancellation: values of similar magnitude and opposite sign, tiny and ordinary
    vol2 = (rng.standard_normal(shape) * 1e-8).astype(np.float32)
    vol2[::2] *= -1
    for axis in range(3):
        k = O.get_gaussian_kernel(1.5)
        out = torch.empty_like(dev(vol2))
        eng.filter_along_axis(dev(vol2), out, axis, k, None)
        assert same_bits_or_nan(out.cpu().numpy(), numpy_axis_filter(vol2, axis, k)), axis


def test_of_passes_vs_reference_toy(eng, golden):
    from flowdenoising_b200.engine import FlowParams
    g = golden("toy_of.npz")
    vol = g["vol"].astype(np.float32)
    kernels = [O.get_gaussian_kernel(float(s)) for s in g["sigmas"]]
    p = FlowParams(int(g["l"]), int(g["w"]))
    d_in = dev(vol)
    a = torch.empty_like(d_in); b = torch.empty_like(d_in); c = torch.empty_like(d_in)
    eng.filter_along_axis(d_in, a, 0, kernels[0], p)
    check_of("toy Z", a.cpu().numpy(), g["Z"])
    eng.filter_along_axis(dev(g["Z"]), b, 1, kernels[1], p)      # each pass checked from the reference's own input
    check_of("toy Y", b.cpu().numpy(), g["ZY"])
    eng.filter_along_axis(dev(g["ZY"]), c, 2, kernels[2], p)
    check_of("toy X", c.cpu().numpy(), g["ZYX"])
    zy, zyx = eng.filter(d_in, kernels, p)
    check_of("toy ZY (chained)", zy.cpu().numpy(), g["ZY"])
    check_of("toy ZYX (chained)", zyx.cpu().numpy(), g["ZYX"])


def test_of_recompute_flow_vs_reference(eng, golden):
    from flowdenoising_b200.engine import FlowParams
    g = golden("toy_of_recompute.npz")
    vol = g["vol"].astype(np.float32)
    k = O.get_gaussian_kernel(float(g["sigmas"][0]))
    out = torch.empty_like(dev(vol))
    eng.filter_along_axis(dev(vol), out, 0, k, FlowParams(int(g["l"]), int(g["w"]), use_prev_flow=False))
    check_of("toy Z recompute_flow", out.cpu().numpy(), g["Z"])


def test_chunking_and_halo_views_do_not_change_results(eng):
    """Chunked processing (bounded workspace) and non-periodic slabs with an explicit periodic halo (the multi-GPU
    building block) must give bit-identical results to the single periodic pass."""
    from flowdenoising_b200.engine import FlowParams
    from flowdenoising_b200._lib import View
    vol = O.synthetic_volume((20, 64, 96), seed=31, noise_sigma=8.0)
    k = O.get_gaussian_kernel(1.0)     # r = 4
    r = k.size // 2
    p = FlowParams()
    d_in = dev(vol)
    full = torch.empty_like(d_in)
    eng.filter_along_axis(d_in, full, 0, k, p)
    for chunk in (3, 7, 12):
        out = torch.empty_like(d_in)
        eng.filter_along_axis(d_in, out, 0, k, p, chunk=chunk)
        assert torch.equal(out, full), f"chunk={chunk}"
    # slab [5, 15) with halo, taken from the periodically padded volume
    Z, Y, X = vol.shape
    s0, s1 = 5, 15
    idx = [(z % Z) for z in range(s0 - r, s1 + r)]
    slab = dev(vol[idx])
    out = torch.empty((s1 - s0, Y, X), dtype=torch.float32, device="cuda")
    v = View(len(idx), s1 - s0, r, 0, Y, X, Y * X, X, Y * X, X)
    eng.filter_view(slab, out, v, k, p)
    assert torch.equal(out, full[s0:s1])
    out2 = torch.empty_like(out)
    eng.filter_view(slab, out2, v, k, None)
    ref = torch.empty_like(d_in)
    eng.filter_along_axis(d_in, ref, 0, k, None)
    assert torch.equal(out2, ref[s0:s1])


def test_cfg1_vs_reference_slices(eng, golden):
    """BASELINE.json configs[0]: 64x256x256 float32, sigma=2, Farneback defaults; the reference's own output slices."""
    from flowdenoising_b200.engine import FlowParams
    g = golden("cfg1_slices.npz")
    vol = O.synthetic_volume((64, 256, 256), seed=0, noise_sigma=20.0)
    assert hashlib.sha256(vol.tobytes()).hexdigest() == str(g["input_sha256"])
    k = O.get_gaussian_kernel(2.0)
    d_in = dev(vol)
    p = FlowParams()
    a = torch.empty_like(d_in); b = torch.empty_like(d_in)
    eng.filter_along_axis(d_in, a, 0, k, p)
    zs = g["zs"]
    check_of("cfg1 Z", a.cpu().numpy()[zs], g["Z"])
    eng.filter_along_axis(a, b, 1, k, p)
    check_of("cfg1 ZY", b.cpu().numpy()[zs], g["ZY"])
    zy, zyx = eng.filter(d_in, [k, k, k], p)
    got = zyx.cpu().numpy()
    check_of("cfg1 ZYX", got[zs], g["ZYX"])
    assert hashlib.sha256(got.tobytes()).hexdigest() == str(g["sha_ZYX"])      # whole volume, not just the slices
    assert hashlib.sha256(zy.cpu().numpy().tobytes()).hexdigest() == str(g["sha_ZY"])
    assert abs(float(got.astype(np.float64).mean()) - float(g["mean_ZYX"])) < 1e-4


def test_module_surface_drop_in(golden):
    """The reference-facing Python classes (host numpy in/out) give the reference's results."""
    from flowdenoising_b200 import flowdenoising as fd
    g = golden("toy_of.npz")
    vol = g["vol"].astype(np.float32)
    kernels = [fd.get_gaussian_kernel(float(s)) for s in g["sigmas"]]
    obj = fd.FlowDenoising(4, vol.copy(), int(g["l"]), int(g["w"]), fd.get_flow_with_prev_flow, fd.warp_slice)
    obj.filter_along_Z(kernels[0])
    check_of("module Z", obj.filtered_vol, g["Z"])
    obj2 = fd.FlowDenoising(4, vol.copy(), int(g["l"]), int(g["w"]), fd.get_flow_with_prev_flow, fd.warp_slice)
    res = obj2.filter(kernels)
    check_of("module ZYX", res, g["ZYX"])
    check_of("module vol==ZY", obj2.vol, g["ZY"])
    # single-slice entry point (reference :306-327)
    obj3 = fd.FlowDenoising(1, vol.copy(), int(g["l"]), int(g["w"]))
    obj3.filter_along_Z_slice(3, kernels[0])
    check_of("module Z slice 3", obj3.filtered_vol[3], g["Z"][3])
    # per-chunk entry points (reference :160-173): one device call per chunk, chunks tile the axis like the reference's
    # pool.starmap arguments (chunk_index, chunk_size, chunk_offset)
    obj4 = fd.FlowDenoising(2, vol.copy(), int(g["l"]), int(g["w"]))
    Z, Y, X = vol.shape
    cs = Z // 2
    assert obj4.filter_along_Z_chunk(0, cs, 0, kernels[0]) == 0
    assert obj4.filter_along_Z_chunk(1, cs, 0, kernels[0]) == 1
    obj4.filter_along_Z_chunk(0, Z - 2 * cs, 2 * cs, kernels[0])      # the remainder chunk (:190-192)
    check_of("module Z chunks", obj4.filtered_vol, g["Z"])
    assert obj4.progress == Z
    obj5 = fd.FlowDenoising(2, vol.copy(), int(g["l"]), int(g["w"]))
    obj5.filter_along_Y_chunk(0, Y, 0, kernels[1])
    obj6 = fd.FlowDenoising(2, vol.copy(), int(g["l"]), int(g["w"]))
    obj6.filter_along_Y(kernels[1])
    assert np.array_equal(obj5.filtered_vol, obj6.filtered_vol)
    obj5.filter_along_X_chunk(1, 3, 2, kernels[2])
    obj6.filter_along_X(kernels[2])
    assert np.array_equal(obj5.filtered_vol[:, :, 5:8], obj6.filtered_vol[:, :, 5:8])
    n = golden("toy_noof.npz")
    gd = fd.GaussianDenoising(2, n["vol"].astype(np.float32))
    assert np.array_equal(gd.filter(kernels), n["ZYX"])
    # free functions
    f = golden("flows.npz")
    v = f["a_vol"].astype(np.float32)
    flow = fd.get_flow_with_prev_flow(v[1], v[0], 3, 5, np.zeros(v[0].shape + (2,), np.float32))
    epe = np.sqrt(((flow - f["a_flow_chain1"]) ** 2).sum(-1))
    assert epe.mean() < 1e-6
    w = fd.warp_slice(v[3], f["a_flow_chain3"])
    assert np.array_equal(w, f["a_warp_chain3"])
    with pytest.raises(AssertionError):
        obj.filter_along_Z(np.ones(4) / 4)      # even kernel: same assert as the reference (:309)


@pytest.mark.parametrize("shape,sigmas,lw", [((7, 45, 67), (0.5, 1.0, 0.5), (3, 5)), ((9, 70, 53), (1.0, 0.5, 0.5), (2, 7)),
                                             ((5, 45, 67), (1.5, 0.5, 0.5), (3, 5)), ((3, 64, 64), (1.0, 0.5, 0.5), (3, 9))])
def test_of_odd_shapes_vs_c_oracle(eng, shape, sigmas, lw):
    """Odd sizes (non-power-of-two pyramid levels, partial strips/tiles, scalar tails of the blur): every pass equals
    the C oracle (itself bit-exact vs cv2) bit for bit. The last two volumes are SHORTER than the kernel along Z
    (13 taps on 5 slices, 9 taps on 3): the reference's `% shape` (:312, :320) then wraps more than once and a slice
    meets itself as its own neighbour."""
    from flowdenoising_b200.engine import FlowParams
    vol = O.synthetic_volume(shape, seed=61, noise_sigma=8.0)
    p = FlowParams(lw[0], lw[1])
    cur = vol
    for axis in range(3):
        k = O.get_gaussian_kernel(sigmas[axis])
        out = torch.empty_like(dev(cur))
        eng.filter_along_axis(dev(cur), out, axis, k, p)
        ref = O.flow_axis_c(cur, axis, k, levels=lw[0], winsize=lw[1])
        check_of(f"odd {shape} axis {axis}", out.cpu().numpy(), ref)
        cur = ref


def test_chunk_remainder_with_mixed_kernels(eng):
    """the last chunk of a pass is smaller than the others; every scratch offset of the flow
    iteration must depend on the chunk CAPACITY only. Width 384 mixes both flow-iteration kernels over the pyramid
    levels (384 / 192 / 96 / 48 px: warp-specialised, then the strip kernel with 2 strips and with 1), chunk = 65 leaves
    remainders of 1 and 2 output slices."""
    from flowdenoising_b200.engine import FlowParams
    p = FlowParams()
    k = O.get_gaussian_kernel(0.5)      # r = 2: 4 flows per slice keep the test short
    for Z in (66, 67):
        vol = O.synthetic_volume((Z, 256, 384), seed=71, noise_sigma=8.0)
        d_in = dev(vol)
        full = torch.empty_like(d_in)
        eng.filter_along_axis(d_in, full, 0, k, p)
        out = torch.empty_like(d_in)
        eng.filter_along_axis(d_in, out, 0, k, p, chunk=65)
        assert torch.equal(out, full), f"Z={Z}"
    # spot-check the unchunked result against the oracle
    ref = O.flow_axis_c(vol, 0, k, s0=5, s1=6)
    assert np.array_equal(full.cpu().numpy()[5], ref[5])


def _crop_case(golden, fname, name):
    g = golden(fname)
    shape = tuple(int(s) for s in g[f"{name}_shape"])
    seed, axis, l, w = (int(x) for x in g[f"{name}_params"])
    noise, sigma = (float(x) for x in g[f"{name}_noise_sigma"])
    vol = O.synthetic_volume(shape, seed=seed, noise_sigma=noise)
    assert hashlib.sha256(vol.tobytes()).hexdigest() == str(g[f"{name}_input_sha256"]), "synthetic input differs"
    return g, vol, axis, sigma, l, w, [int(s) for s in g[f"{name}_slices"]]


def _check_crop(g, name, got, axis, slices):
    for s in slices:
        o = np.ascontiguousarray(np.take(got, s, axis=axis))
        crop = g[f"{name}_crop_{s}"]
        check_of(f"{name} slice {s} (crop)", o[:crop.shape[0], :crop.shape[1]], crop)
        assert hashlib.sha256(o.tobytes()).hexdigest() == str(g[f"{name}_sha_{s}"]), f"{name} slice {s}: SHA-256 differs"


@pytest.mark.parametrize("name", ["z", "y", "x"])
def test_cfg2_crops_vs_reference(eng, golden, name):
    """BASELINE.json configs[1] at the BENCHMARKED slice size: 1024 x 1024 (Z pass) and 512 x 1024 (Y / X passes)
    slices, sigma = 2, Farneback defaults -- output slices of the unmodified reference (oracle/gen_golden_big.py),
    SHA-256 of the whole slice. k_flow_iter_ws runs 9 strips per image here."""
    from flowdenoising_b200.engine import FlowParams
    g, vol, axis, sigma, l, w, slices = _crop_case(golden, "cfg2_crops.npz", name)
    k = O.get_gaussian_kernel(sigma)
    d_in = dev(vol)
    out = torch.empty_like(d_in)
    eng.filter_along_axis(d_in, out, axis, k, FlowParams(l, w))
    _check_crop(g, name, out.cpu().numpy(), axis, slices)
    del out, d_in
    eng.release_workspace()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["z", "y"])
def test_cfg4_crops_vs_reference(eng, golden, name):
    """BASELINE.json configs[3] parameters: sigma = (4, 2, 2), levels = 5, winsize = 9, uint8 TIFF stack (cast to
    float32 at load, src/flowdenoising.py:475), 2048-wide slices: 1024 x 2048 with 6 pyramid levels and a 33-tap Z
    kernel, 256 x 2048 with 4 levels."""
    from flowdenoising_b200.engine import FlowParams
    g, vol, axis, sigma, l, w, slices = _crop_case(golden, "cfg4_crops.npz", name)
    vol8 = vol.astype(np.uint8)
    assert np.array_equal(vol8.astype(np.float32), vol)
    k = O.get_gaussian_kernel(sigma)
    d_in = torch.from_numpy(vol8).cuda().to(torch.float32)      # the TIFF -> float32 cast, on the device
    out = torch.empty_like(d_in)
    # the Z-pass case only needs slice 17: a slab view with an explicit halo keeps the test short
    if name == "z":
        from flowdenoising_b200._lib import View
        r = k.size // 2
        Z, Y, X = vol.shape
        s = slices[0]
        idx = [(z % Z) for z in range(s - r, s + r + 1)]
        slab = d_in[idx].contiguous()
        o1 = torch.empty((1, Y, X), dtype=torch.float32, device="cuda")
        eng.filter_view(slab, o1, View(len(idx), 1, r, 0, Y, X, Y * X, X, Y * X, X), k, FlowParams(l, w))
        out[s] = o1[0]
    else:
        eng.filter_along_axis(d_in, out, axis, k, FlowParams(l, w))
    _check_crop(g, name, out.cpu().numpy(), axis, slices)
    del out, d_in
    eng.release_workspace()
    torch.cuda.empty_cache()
