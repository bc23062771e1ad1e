"""GPU tests of the slab-sharded path (flowdenoising_b200/dist.py): packing / transposing kernels, non-periodic halo
views and the NCCL exchanges. With one GPU the world has a single rank (every exchange degenerates to a local copy
but all kernels and views are exercised); with >= 2 GPUs two ranks run over NCCL. Results must be bit-identical to
the single-GPU periodic pass (SURVEY.md §8e "Determinism")."""
import os
import socket

import numpy as np
import pytest

from oracle import fd_oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, sigmas, use_of, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from flowdenoising_b200.dist import DistributedDenoiser
        from flowdenoising_b200.engine import DeviceEngine, FlowParams
        eng = DeviceEngine(torch.device("cuda", rank))
        vol = O.synthetic_volume(shape, seed=51, noise_sigma=8.0)
        kernels = [O.get_gaussian_kernel(s) for s in sigmas]
        flow = FlowParams() if use_of else None
        dd = DistributedDenoiser(eng, shape, flow)
        zs, ze = dd.z_range
        zy, zyx = dd.filter(torch.from_numpy(vol[zs:ze].copy()).cuda(), kernels, want_zy=True)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"zy_{rank}.npy"), zy.cpu().numpy())
        np.save(os.path.join(out_dir, f"zyx_{rank}.npy"), zyx.cpu().numpy())
        # host slabs in / out: copy stream, the Z pass and the X pass split into windows, pitched downloads
        host = torch.from_numpy(vol[zs:ze].copy()).pin_memory()
        out_host = torch.full(host.shape, float("nan"), dtype=torch.float32).pin_memory()
        for _ in range(2):      # the second call re-uses cached buffers while copies of the first may still be queued
            res = dd.filter_host(host, out_host, kernels)
        torch.cuda.synchronize()
        assert torch.equal(res.cpu(), out_host)
        np.save(os.path.join(out_dir, f"zyxh_{rank}.npy"), out_host.numpy())
    finally:
        dist.destroy_process_group()


def _single_gpu_reference(shape, sigmas, use_of):
    from flowdenoising_b200.engine import DeviceEngine, FlowParams
    eng = DeviceEngine()
    vol = O.synthetic_volume(shape, seed=51, noise_sigma=8.0)
    kernels = [O.get_gaussian_kernel(s) for s in sigmas]
    zy, zyx = eng.filter(torch.from_numpy(vol).cuda(), kernels, FlowParams() if use_of else None)
    return zy.cpu().numpy(), zyx.cpu().numpy()


@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("shape,sigmas,use_of", [((24, 40, 72), (1.0, 1.0, 0.5), False),
                                                 ((18, 64, 80), (1.0, 0.5, 0.75), True)])
def test_distributed_equals_single_gpu(tmp_path, world, shape, sigmas, use_of):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ref_zy, ref_zyx = _single_gpu_reference(shape, sigmas, use_of)
    mp.spawn(_worker, args=(world, _free_port(), shape, sigmas, use_of, str(tmp_path)), nprocs=world, join=True)
    zy = np.concatenate([np.load(tmp_path / f"zy_{r}.npy") for r in range(world)])
    zyx = np.concatenate([np.load(tmp_path / f"zyx_{r}.npy") for r in range(world)])
    assert np.array_equal(zy, ref_zy)
    assert np.array_equal(zyx, ref_zyx)
    zyxh = np.concatenate([np.load(tmp_path / f"zyxh_{r}.npy") for r in range(world)])
    assert np.array_equal(zyxh, ref_zyx)
