"""GPU parity tests, stage by stage, through the C ABI (ctypes) against the CPU oracle.

The oracle's C restatement is bit-exact against cv2 4.13 / the reference (tests/test_oracle_golden.py), and the
kernels follow the same arithmetic, so most comparisons demand bit-exactness; where the GPU legitimately differs
(float64 summation order of the box window, ~1e-16 relative) the tolerance is written in the test.
"""
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


def R_floats(h, w):
    p = h * w
    return 4 * p + (p + 3) // 4 * 4


def R_to_oracle_layout(R, h, w):   # library layout (n, R_floats): [h*w] float4 + [h*w] float  ->  (n, h, w, 5)
    n = R.shape[0]
    out = np.empty((n, h, w, 5), np.float32)
    out[..., :4] = R[:, :4 * h * w].reshape(n, h, w, 4)
    out[..., 4] = R[:, 4 * h * w:5 * h * w].reshape(n, h, w)
    return out


def R_from_oracle_layout(R):  # (n, h, w, 5) -> (n, R_floats)
    n, h, w, _ = R.shape
    out = np.zeros((n, R_floats(h, w)), np.float32)
    out[:, :4 * h * w] = R[..., :4].reshape(n, -1)
    out[:, 4 * h * w:5 * h * w] = R[..., 4].reshape(n, -1)
    return out


def images(shape, n, seed):
    return O.synthetic_volume((n,) + shape, seed=seed, noise_sigma=10.0)


SHAPES = [(128, 160), (96, 130), (100, 77), (33, 35)]


@pytest.mark.parametrize("shape", SHAPES + [(256, 320)])
def test_pyramid_levels_bit_exact(eng, shape):
    """Stage 1: GaussianBlur + resize per level == cv2 (via the bit-exact C oracle)."""
    H, W = shape
    n = 3
    imgs = images(shape, n, 1) + np.float32(0.37)     # non-integer values
    d = dev(imgs)
    tmp = torch.empty(2 * n * H * W, dtype=torch.float32, device="cuda")
    for k, (h, w, ksz, sigma) in enumerate(O.level_geometry(H, W, 5)):
        out = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
        rc = eng.lib.fdn_pyramid_level(d.data_ptr(), n, H, W, H * W, W, ksz, sigma, h, w, tmp.data_ptr(),
                                       out.data_ptr(), None)
        assert rc == 0, eng.lib.fdn_last_error()
        got = out.cpu().numpy()
        for i in range(n):
            ref = O.pyramid_level(imgs[i], ksz, sigma, h, w)
            assert np.array_equal(got[i], ref), f"level {k} {shape}: max|d|={np.abs(got[i] - ref).max()}"


def test_pyramid_strided_view(eng):
    """Slices taken along Y of a [Z,Y,X] volume (slice_stride = X, row_stride = Y*X) read correctly."""
    vol = images((40, 48), 36, 2)          # Z=36, Y=40, X=48
    d = dev(vol)
    Z, Y, X = vol.shape
    tmp = torch.empty(2 * Y * Z * X, dtype=torch.float32, device="cuda")
    out = torch.empty((Y, Z, X), dtype=torch.float32, device="cuda")
    rc = eng.lib.fdn_pyramid_level(d.data_ptr(), Y, Z, X, X, Y * X, 3, 0.0, Z, X, tmp.data_ptr(), out.data_ptr(), None)
    assert rc == 0, eng.lib.fdn_last_error()
    got = out.cpu().numpy()
    for y in (0, 7, 39):
        assert np.array_equal(got[y], O.pyramid_level(vol[:, y, :], 3, 0.0, Z, X))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("poly", [(5, 1.2), (7, 1.5)])
def test_polyexp_bit_exact(eng, shape, poly):
    """Stage 2: FarnebackPolyExp."""
    n = 2
    imgs = images(shape, n, 3) * np.float32(0.731)
    h, w = shape
    assert eng.lib.fdn_polyexp_floats(h, w) == R_floats(h, w)
    R = torch.zeros((n, R_floats(h, w)), dtype=torch.float32, device="cuda")
    rc = eng.lib.fdn_polyexp(dev(imgs).data_ptr(), n, h, w, poly[0], poly[1], R.data_ptr(), None)
    assert rc == 0, eng.lib.fdn_last_error()
    got = R_to_oracle_layout(R.cpu().numpy(), h, w)
    for i in range(n):
        ref = O.polyexp(imgs[i], poly[0], poly[1])
        assert np.array_equal(got[i], ref), f"max|d|={np.abs(got[i] - ref).max()}"


@pytest.mark.parametrize("shape", SHAPES + [(40, 300), (9, 520), (64, 64), (33, 128), (17, 96), (70, 244)])
@pytest.mark.parametrize("win", [5, 9, 15])
def test_flow_iteration(eng, shape, win):
    """Stage 3: UpdateMatrices + box blur + solve, both running sums reproduced exactly -> bit-exact."""
    n = 2
    h, w = shape
    imgs = images(shape, 2 * n, 4)
    R = np.stack([O.polyexp(imgs[i]) for i in range(2 * n)])           # (2n, h, w, 5)
    rng = np.random.default_rng(5)
    flow = (rng.standard_normal((n, h, w, 2)) * 1.5).astype(np.float32)
    flow[0, :4, :4] = 50.0      # out-of-image lookups take the border branch
    dR = dev(R_from_oracle_layout(R))
    dflow = dev(flow)
    out = torch.empty_like(dflow)
    nscr = eng.lib.fdn_flow_iteration_scratch_bytes(n, h, w)
    scr = torch.empty(nscr, dtype=torch.uint8, device="cuda")
    rc = eng.lib.fdn_flow_iteration(dR[:n].data_ptr(), dR[n:].data_ptr(), dflow.data_ptr(), out.data_ptr(), n, h, w,
                                    win, scr.data_ptr(), nscr, None)
    assert rc == 0, eng.lib.fdn_last_error()
    got = out.cpu().numpy()
    for i in range(n):
        M = O.update_matrices(R[i], R[n + i], flow[i])
        ref = O.blur_solve(M, win)
        d = np.abs(got[i] - ref)
        assert np.array_equal(got[i], ref), f"max|d|={d.max()} bit-equal fraction {np.mean(got[i] == ref)}"


def _run_iterations(eng, dR, n, flow, h, w, win, iters, variant, scr, merged):
    """`iters` chained iterations of one level through the C ABI; merged: fdn_flow_iterations (one launch runs up to
    three iterations), else one fdn_flow_iteration call per iteration. The scratch is reused without clearing."""
    lib = eng.lib
    lib.fdn_set_flow_iter_variant(variant)
    try:
        cur = dev(flow)
        if merged:
            t1, t2 = torch.empty_like(cur), torch.empty_like(cur)
            res = C.c_void_p()
            rc = lib.fdn_flow_iterations(dR[:n].data_ptr(), dR[1:].data_ptr(), cur.data_ptr(), t1.data_ptr(), t2.data_ptr(),
                                         n, h, w, win, iters, scr.data_ptr(), scr.numel(), None, C.byref(res))
            assert rc == 0, lib.fdn_last_error()
            out = {cur.data_ptr(): cur, t1.data_ptr(): t1, t2.data_ptr(): t2}[res.value]
            return out.cpu().numpy()
        for _ in range(iters):
            out = torch.empty_like(cur)
            rc = lib.fdn_flow_iteration(dR[:n].data_ptr(), dR[1:].data_ptr(), cur.data_ptr(), out.data_ptr(), n, h, w,
                                        win, scr.data_ptr(), scr.numel(), None)
            assert rc == 0, lib.fdn_last_error()
            cur = out
        return cur.cpu().numpy()
    finally:
        lib.fdn_set_flow_iter_variant(1)


# (64, 1024) and (32, 2048): 9 and 18 strips of the warp-specialised kernel with h >= 16 (the benchmarked widths)
@pytest.mark.parametrize("win", [5, 9])
@pytest.mark.parametrize("shape,scale,n", [((96, 256), 1.0, 5), ((130, 372), 6.0, 5), ((64, 64), 3.0, 5),
                                           ((64, 1024), 2.0, 40), ((32, 2048), 4.0, 3), ((50, 1000), 1.0, 2),
                                           ((16, 464), 2.0, 3), ((23, 480), 2.0, 3)])
def test_flow_iteration_kernels_agree(eng, shape, scale, n, win):
    """The warp-specialised kernel (k_flow_iter_ws, default for winsize 5 and 9) and the strip kernel (k_flow_iter, pinned
    against the oracle above) write the same bits, also for flows of many pixels (gathers far from the identity
    position, out-of-image lookups), with three iterations merged into one launch (ticketed work items, per-pair
    dependencies between the iterations) and over chained launches that reuse the scratch without clearing it."""
    h, w = shape
    imgs = images(shape, 2, 21)
    rng = np.random.default_rng(22)
    imgs = np.concatenate([imgs] + [np.clip(imgs[:1] + rng.normal(0, 6, (1,) + shape), 0, 255).astype(np.float32)
                                    for _ in range(n - 1)])
    R = np.stack([O.polyexp(imgs[i]) for i in range(n + 1)])
    flow = (rng.standard_normal((n, h, w, 2)) * scale).astype(np.float32)
    flow[1, :, : w // 2] += 7.5        # a coherent drift on half an image
    flow[2 % n, 5:9, 10:40] = 1e4      # far outside the image
    dR = dev(R_from_oracle_layout(R))
    nscr = eng.lib.fdn_flow_iteration_scratch_bytes(n, h, w)
    scr = torch.zeros(nscr, dtype=torch.uint8, device="cuda")
    ref = _run_iterations(eng, dR, n, flow, h, w, win, 3, 0, scr, merged=False)
    for iters, merged in ((3, True), (3, False), (3, True)):
        got = _run_iterations(eng, dR, n, flow, h, w, win, iters, 1, scr, merged)
        assert np.array_equal(ref.view(np.int32), got.view(np.int32)), (iters, merged)
    # other iteration counts through the merged entry point (4 = 3 + 1 launches, 2, 1) against the strip kernel
    for iters in (1, 2, 4):
        a = _run_iterations(eng, dR, n, flow, h, w, win, iters, 0, scr, merged=True)
        b = _run_iterations(eng, dR, n, flow, h, w, win, iters, 1, scr, merged=True)
        assert np.array_equal(a.view(np.int32), b.view(np.int32)), iters


@pytest.mark.parametrize("win", [5, 9])
def test_flow_iteration_ws_vs_oracle_wide(eng, win):
    """k_flow_iter_ws on 9 strips (64 x 1024) against the oracle itself, not only against the strip kernel."""
    n, h, w = 2, 64, 1024
    imgs = images((h, w), 2 * n, 31)
    R = np.stack([O.polyexp(imgs[i]) for i in range(2 * n)])
    rng = np.random.default_rng(32)
    flow = (rng.standard_normal((n, h, w, 2)) * 1.5).astype(np.float32)
    dR = dev(R_from_oracle_layout(R))
    out = torch.empty((n, h, w, 2), dtype=torch.float32, device="cuda")
    nscr = eng.lib.fdn_flow_iteration_scratch_bytes(n, h, w)
    scr = torch.empty(nscr, dtype=torch.uint8, device="cuda")
    rc = eng.lib.fdn_flow_iteration(dR[:n].data_ptr(), dR[n:].data_ptr(), dev(flow).data_ptr(), out.data_ptr(), n, h, w,
                                    win, scr.data_ptr(), nscr, None)
    assert rc == 0, eng.lib.fdn_last_error()
    got = out.cpu().numpy()
    for i in range(n):
        ref = O.blur_solve(O.update_matrices(R[i], R[n + i], flow[i]), win)
        assert np.array_equal(got[i], ref), f"bit-equal fraction {np.mean(got[i] == ref)}"


def test_flow_resampling_bit_exact(eng):
    rng = np.random.default_rng(6)
    n = 2
    for (H, W, h, w) in [(128, 160, 32, 40), (128, 160, 64, 80), (96, 130, 48, 65), (100, 77, 50, 38), (260, 300, 32, 38)]:
        flow = (rng.standard_normal((n, H, W, 2)) * 2).astype(np.float32)
        out = torch.empty((n, h, w, 2), dtype=torch.float32, device="cuda")
        scale = 0.25
        rc = eng.lib.fdn_flow_area_down(dev(flow).data_ptr(), n, H, W, out.data_ptr(), h, w, scale, None)
        assert rc == 0, eng.lib.fdn_last_error()
        for i in range(n):
            ref = O.resize_area(flow[i], h, w) * np.float32(scale)
            assert np.array_equal(out[i].cpu().numpy(), ref), (H, W, h, w)
    for (hin, win, h, w) in [(32, 40, 64, 80), (48, 65, 96, 130), (50, 38, 100, 77), (16, 19, 33, 37)]:
        flow = (rng.standard_normal((n, hin, win, 2)) * 2).astype(np.float32)
        out = torch.empty((n, h, w, 2), dtype=torch.float32, device="cuda")
        rc = eng.lib.fdn_flow_upsample(dev(flow).data_ptr(), n, hin, win, out.data_ptr(), h, w, None)
        assert rc == 0, eng.lib.fdn_last_error()
        for i in range(n):
            ref = O.resize_linear(flow[i], h, w, ipp=False) * np.float32(2)
            assert np.array_equal(out[i].cpu().numpy(), ref), (hin, win, h, w)


def test_warp_accumulate_bit_exact(eng):
    """Stage 4: cv2.remap (1/32-px quantiser, replicate border) + float64 accumulate rounded per tap."""
    rng = np.random.default_rng(7)
    n, H, W = 3, 70, 90
    imgs = images((H, W), n, 8) + np.float32(0.25)
    flow = (rng.standard_normal((n, H, W, 2)) * 3).astype(np.float32)
    flow[1] *= 30   # far out-of-image samples
    flow[2, ::2] = np.float32(0.015625)   # exact ties of the 1/32 quantiser
    acc0 = rng.standard_normal((n, H, W)).astype(np.float32) * 10
    from flowdenoising_b200.engine import DeviceEngine  # noqa: F401
    acc = dev(acc0)
    k = 0.19947114020071635
    eng.warp_accumulate(dev(imgs), dev(flow), k, acc)
    got = acc.cpu().numpy()
    for i in range(n):
        ref = acc0[i].copy()
        O.lib().fdo_accumulate(ref.ctypes.data_as(C.POINTER(C.c_float)),
                               O.warp_slice_c(imgs[i], flow[i]).ctypes.data_as(C.POINTER(C.c_float)),
                               C.c_double(k), C.c_size_t(ref.size))
        assert np.array_equal(got[i], ref)
    # identity warp = centre tap
    acc = dev(acc0)
    eng.warp_accumulate(dev(imgs), None, k, acc)
    ref = (acc0.astype(np.float64) + imgs.astype(np.float64) * k).astype(np.float32)
    assert np.array_equal(acc.cpu().numpy(), ref)


@pytest.mark.parametrize("case", ["a", "b", "d"])
def test_farneback_vs_reference_golden(eng, golden, case):
    """Whole Farneback vs the reference's get_flow outputs (cv2.calcOpticalFlowFarneback): north-star tolerance is
    mean EPE <= 0.05 px; this implementation is expected to be bit-exact up to isolated last-ulp differences."""
    from flowdenoising_b200.engine import FlowParams
    g = golden("flows.npz")
    v = g[f"{case}_vol"].astype(np.float32)
    l, w = (int(x) for x in g[f"{case}_lw"])
    centre = dev(v[0:1])
    flow = torch.zeros((1,) + v[0].shape + (2,), dtype=torch.float32, device="cuda")
    p = FlowParams(l, w, 3, 5, 1.2, True)
    for j in (1, 2, 3):
        eng.farneback(centre, dev(v[j:j + 1]), flow, p)
        if j != 2:
            ref = g[f"{case}_flow_chain{j}"]
            got = flow[0].cpu().numpy()
            epe = np.sqrt(((got - ref) ** 2).sum(-1))
            print(f"case {case} chain{j}: EPE mean {epe.mean():.3e} max {epe.max():.3e} exact {np.mean(got == ref):.5f}")
            assert epe.mean() <= 0.05                      # north-star bound
            assert np.array_equal(got, ref)                # what this implementation delivers
    f2 = torch.zeros_like(flow)
    eng.farneback(centre, dev(v[2:3]), f2, FlowParams(l, w, 3, 5, 1.2, False))
    ref = g[f"{case}_flow_noprev2"]
    assert np.array_equal(f2[0].cpu().numpy(), ref)


def test_farneback_vs_live_cv2(eng):
    cv2 = pytest.importorskip("cv2")
    from flowdenoising_b200.engine import FlowParams
    for shape, l, w, it in [((256, 256), 3, 5, 3), ((64, 256), 3, 5, 3), ((200, 333), 3, 7, 2), ((512, 384), 5, 9, 3),
                            ((40, 50), 3, 5, 1)]:
        v = images(shape, 2, 9)
        ref = cv2.calcOpticalFlowFarneback(v[0], v[1], None, 0.5, l, w, it, 5, 1.2, 0)
        flow = torch.zeros((1,) + shape + (2,), dtype=torch.float32, device="cuda")
        eng.farneback(dev(v[0:1]), dev(v[1:2]), flow, FlowParams(l, w, it, 5, 1.2, False))
        got = flow[0].cpu().numpy()
        epe = np.sqrt(((got - ref) ** 2).sum(-1))
        print(f"{shape} l{l} w{w}: EPE mean {epe.mean():.3e} max {epe.max():.3e} exact {np.mean(got == ref):.5f}")
        assert epe.mean() <= 0.05          # north-star bound
        assert np.array_equal(got, ref)    # what this implementation actually delivers (AVX2 host, see oracle)


def test_transpose(eng):
    rng = np.random.default_rng(10)
    a = rng.standard_normal((5, 37, 70)).astype(np.float32)
    out = eng.transpose_yx(dev(a))
    assert np.array_equal(out.cpu().numpy(), np.transpose(a, (0, 2, 1)))


def _launch_names(lib):
    return [lib.fdn_launch_log_name(i).decode() for i in range(lib.fdn_launch_log_count())]


def test_levels_launch_their_specialised_kernels(eng):
    """Which kernel a pyramid level / window size really launches (round 1 shipped register-window blurs for 7 and 17
    taps while the pyramid uses 9 and 19: dead code). The benchmarked slice sizes must take the specialised paths:
    4-outputs-per-thread blurs for every level's tap count, the TMA polynomial expansion, the exact-x2 upsample and
    the warp-specialised flow iteration for winsize 5 and 9; other window sizes take the strip kernel."""
    from flowdenoising_b200.engine import FlowParams, level_geometry
    lib = eng.lib
    for (H, W, levels, win) in [(256, 1024, 3, 5), (512, 2048, 5, 9), (96, 512, 3, 7)]:
        geo = level_geometry(H, W, levels)
        assert [g[2] for g in geo] == [3, 3, 9, 19, 39, 79][:len(geo)]      # smoothing taps per level
        assert len(geo) == {3: 4, 5: 5}[levels] if H >= 256 else len(geo) == 2
        v = images((H, W), 2, 41)
        flow = torch.zeros((1, H, W, 2), dtype=torch.float32, device="cuda")
        lib.fdn_launch_log_enable(1)
        eng.farneback(dev(v[0:1]), dev(v[1:2]), flow, FlowParams(levels, win, 3, 5, 1.2, True))
        torch.cuda.synchronize()
        lib.fdn_launch_log_enable(0)
        names = _launch_names(lib)
        nl = len(geo)
        assert names.count("k_blur_rows4") == 2 * nl and names.count("k_blur_cols4") == 2 * nl, names
        assert "k_blur_rows" not in names and "k_blur_cols" not in names
        assert names.count("k_polyexp_tma") == 2 * nl and "k_polyexp" not in names
        assert names.count("k_flow_upsample_x2") == nl - 1 and "k_flow_upsample" not in names
        if win in (5, 9):
            ws_levels = sum(1 for (h, w, _k, _s) in geo if w >= 64 and h >= 16 and w % 4 == 0)
            assert names.count("k_flow_iter_ws") == ws_levels, names           # one launch = the three iterations
            assert names.count("k_flow_iter") == 3 * (nl - ws_levels)
        else:
            assert names.count("k_flow_iter") == 3 * nl and "k_flow_iter_ws" not in names
    # a whole OF pass: both chain directions of a chain step in ONE warp kernel launch, one finishing launch per chunk
    vol = dev(images((64, 256), 12, 43))
    out = torch.empty_like(vol)
    lib.fdn_launch_log_enable(1)
    eng.filter_along_axis(vol, out, 0, O.get_gaussian_kernel(1.0), FlowParams())
    torch.cuda.synchronize()
    lib.fdn_launch_log_enable(0)
    names = _launch_names(lib)
    r = 4   # sigma 1 -> 9 taps
    assert names.count("k_warp_pair") == r and names.count("k_acc_finish") == 1 and "k_warp_acc" not in names
    assert lib.fdn_launch_log_count() == len(names) and lib.fdn_launch_log_name(10 ** 6) == b""
