"""CPU, world_size 2, gloo: the batch sharding and the one collective of the multi-GPU path
(contiguous slices per rank, no data-path traffic, all_gather of per-member results)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpsim_b200.distributed import shard_range, gather_slices
    lo, hi = shard_range(total, rank, world)
    # per-member "norms" and complex "amplitudes" that encode the global member index
    norms = torch.arange(lo, hi, dtype=torch.float32) * 0.5
    amps = (torch.arange(lo, hi, dtype=torch.float32)[:, None] + 1j * torch.arange(3, dtype=torch.float32)[None, :]).to(torch.complex64)
    full_n = gather_slices(norms, total)
    full_a = gather_slices(amps, total)
    if rank == 0:
        np.save(os.path.join(out_dir, "norms.npy"), full_n.numpy())
        np.save(os.path.join(out_dir, "amps.npy"), full_a.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_slices_world2(tmp_path, total):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    norms = np.load(tmp_path / "norms.npy")
    amps = np.load(tmp_path / "amps.npy")
    np.testing.assert_allclose(norms, np.arange(total) * 0.5)
    assert amps.shape == (total, 3)
    np.testing.assert_allclose(amps.real, np.arange(total)[:, None] * np.ones((1, 3)))
    np.testing.assert_allclose(amps.imag, np.ones((total, 1)) * np.arange(3)[None, :])
