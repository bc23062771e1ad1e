"""Multi-rank host logic of flowdenoising_b200/dist.py (slab ranges, periodic halo exchange, all-to-all re-slab with
packing / transposing unpack) exercised on CPU with the gloo backend, world sizes 2 and 3.

The compute object is a CPU stand-in built on the ORACLE (test infrastructure): it implements the four device
operations DistributedDenoiser calls (filter_view, copy3d, transpose_strided, empty) with NumPy semantics identical
to the C-ABI contracts in include/fdn_b200.h. The distributed result must equal the single-process oracle bit for
bit, for both the no-OF and the OF path, through filter() (device slabs) and filter_host() (host slabs, passes split
into windows). The GPU twin of this test is tests/test_gpu_dist.py.
"""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from oracle import fd_oracle as O  # noqa: E402


class CpuOps:
    """Oracle-backed stand-in for DeviceEngine (same method contracts, CPU tensors)."""
    torch = torch

    def empty(self, shape):
        return torch.full(tuple(int(s) for s in shape), float("nan"), dtype=torch.float32)

    @staticmethod
    def _flat(t):
        assert t.is_contiguous()
        return t.view(-1).numpy()

    def copy3d(self, src, src_off, in_sa, in_sb, b0, bw, c0, cw, dst, dst_off, out_sa, out_sb, A, B, C):
        s, d = self._flat(src), self._flat(dst)
        a = np.arange(A)[:, None, None]
        b = np.arange(B)[None, :, None]
        c = np.arange(C)[None, None, :]
        d[dst_off + a * out_sa + b * out_sb + c] = s[src_off + a * in_sa + ((b0 + b) % bw) * in_sb + (c0 + c) % cw]

    def transpose_strided(self, src, src_off, in_sn, in_sa, dst, dst_off, out_sn, out_sb, n, A, B):
        s, d = self._flat(src), self._flat(dst)
        i = np.arange(n)[:, None, None]
        a = np.arange(A)[None, :, None]
        b = np.arange(B)[None, None, :]
        d[dst_off + i * out_sn + b * out_sb + a] = s[src_off + i * in_sn + a * in_sa + b]

    def filter_view(self, d_in, d_out, v, kernel, flow, chunk=None, exact=True):
        src = np.lib.stride_tricks.as_strided(self._flat(d_in), (v.n_in, v.H, v.W),
                                              (4 * v.in_slice_stride, 4 * v.in_row_stride, 4))
        dst = np.lib.stride_tricks.as_strided(self._flat(d_out), (v.n_out, v.H, v.W),
                                              (4 * v.out_slice_stride, 4 * v.out_row_stride, 4))
        assert not v.periodic and v.halo >= kernel.size // 2
        slab = np.ascontiguousarray(src)
        o = O.OracleDenoiser(1, slab, use_OF=flow is not None, backend="c",
                             **({} if flow is None else dict(l=flow.levels, w=flow.winsize)))
        for s in range(v.n_out):
            o.filter_slice(0, s + v.halo, kernel)       # never wraps: halo >= r on both sides
            dst[s] = o.filtered_vol[s + v.halo]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, sigmas, use_of, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flowdenoising_b200.dist import DistributedDenoiser, split_range
        from flowdenoising_b200.engine import FlowParams
        vol = O.synthetic_volume(shape, seed=41, noise_sigma=6.0)
        kernels = [O.get_gaussian_kernel(s) for s in sigmas]
        dd = DistributedDenoiser(CpuOps(), shape, FlowParams() if use_of else None)
        zs, ze = split_range(shape[0], world, rank)
        assert dd.z_range == (zs, ze)
        zy, zyx = dd.filter(torch.from_numpy(vol[zs:ze].copy()), kernels, want_zy=True)
        np.save(os.path.join(result_dir, f"zy_{rank}.npy"), zy.numpy())
        np.save(os.path.join(result_dir, f"zyx_{rank}.npy"), zyx.numpy())
        # host slabs in / out with the Z pass and the X pass split into windows (the transfer-hiding sequence)
        out_host = torch.full((ze - zs,) + tuple(shape[1:]), float("nan"), dtype=torch.float32)
        res = dd.filter_host(torch.from_numpy(vol[zs:ze].copy()), out_host, kernels)
        assert torch.equal(res, out_host)
        np.save(os.path.join(result_dir, f"zyxh_{rank}.npy"), out_host.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,sigmas,use_of", [
    (2, (10, 12, 14), (1.0, 0.5, 1.0), False),      # r = 4 > slab/2: halo comes from both neighbours
    (3, (11, 13, 10), (0.5, 1.0, 0.5), False),      # uneven splits, X smaller than 2r+... wraps
    (2, (6, 40, 48), (0.5, 0.5, 0.5), True),        # OF path (r = 2), level 0 only for the thin passes
    (2, (24, 12, 14), (0.5, 0.5, 0.5), False),      # filter_host splits the Z pass (head 3 of 12) and the X pass (6 + 1)
    (3, (20, 13, 11), (0.5, 1.0, 0.5), False),      # ... with uneven slabs: Zl = 7, 7, 6 and Xl = 4, 4, 3 (tail 1)
    (2, (16, 40, 48), (0.5, 0.5, 0.5), True),       # ... on the OF path
])
def test_distributed_equals_single_process_oracle(tmp_path, world, shape, sigmas, use_of):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, shape, sigmas, use_of, str(tmp_path)), nprocs=world, join=True)
    vol = O.synthetic_volume(shape, seed=41, noise_sigma=6.0)
    kernels = [O.get_gaussian_kernel(s) for s in sigmas]
    o = O.OracleDenoiser(1, vol.copy(), use_OF=use_of, backend="c")
    ref_zyx = o.filter(kernels)
    ref_zy = o.vol
    zy = np.concatenate([np.load(tmp_path / f"zy_{r}.npy") for r in range(world)])
    zyx = np.concatenate([np.load(tmp_path / f"zyx_{r}.npy") for r in range(world)])
    assert zy.shape == ref_zy.shape and not np.isnan(zy).any() and not np.isnan(zyx).any()
    assert np.array_equal(zy, ref_zy)
    assert np.array_equal(zyx, ref_zyx)
    zyxh = np.concatenate([np.load(tmp_path / f"zyxh_{r}.npy") for r in range(world)])
    assert np.array_equal(zyxh, ref_zyx)


def test_split_range_covers_axis():
    from flowdenoising_b200.dist import split_range
    for n in (1, 7, 64, 1000):
        for parts in (1, 2, 3, 8):
            if parts > n:
                continue
            r = [split_range(n, parts, i) for i in range(parts)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(parts - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
