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


# ---- batched parameter sweeps (mpsim_cirq/simulator.py:67-87) sharded over the process group -----------
class _OracleBatch:
    """Stand-in for MPSBatch on a box without a GPU: the same results interface, computed by the CPU
    oracle.  Only the HOST logic of MPSimulator.simulate_sweep_batched is under test here (sharding,
    gathers, result order); the device path has its own -m gpu test."""

    class _Chain:
        device = "cpu"

    def __init__(self, op_lists, nqubits, options):
        from oracle.mps_oracle import OracleMPS
        from mpsim_b200.node import tensor_of
        self._chain = self._Chain()
        self.members = []
        for ops in op_lists:
            ora = OracleMPS(nqubits, dtype=np.complex128)
            for op in ops:
                t = np.asarray(tensor_of(op.node(copy=False)))
                if len(op.qudit_indices) == 1:
                    ora.apply_one_qudit_gate(t, *op.qudit_indices)
                else:
                    ora.apply_two_qudit_gate(t, *op.qudit_indices, **options)
            self.members.append(ora)

    def norms_device(self):
        import torch
        return torch.tensor([m.norm() for m in self.members], dtype=torch.float32)

    def amplitudes_device(self, bits):
        import torch
        n = bits.shape[1]
        idx = (bits.astype(np.int64) << np.arange(n - 1, -1, -1)[None, :]).sum(axis=1)
        return torch.tensor(np.stack([m.wavefunction()[idx] for m in self.members]), dtype=torch.complex64)


def _sweep_circuit(n):
    from tests import _fake_cirq as fc
    ops = [fc.H(0)] + [fc.CNOT(i, i + 1) for i in range(n - 1)] + [fc.Rx("t", q) for q in range(n)]
    ops += [fc.CZPow(0.25, i, i + 1) for i in range(0, n - 1, 2)] + [fc.Rx("t", 0)]
    return fc.Circuit(ops)


def _sweep_worker(rank, world, port, total, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpsim_b200.mpsim_cirq.simulator import MPSimulator
    n = 5
    sim = MPSimulator({"maxsvals": 4})
    sim._run_batch = lambda op_lists, nq: _OracleBatch(op_lists, nq, {"maxsvals": 4})
    params = [{"t": 0.1 * (i + 1)} for i in range(total)]
    bits = np.array([[0] * n, [1] * n, [1, 0, 1, 0, 1]], dtype=np.uint8)
    res = sim.simulate_sweep_batched(_sweep_circuit(n), params, amplitudes=bits)
    np.save(os.path.join(out_dir, f"range{rank}.npy"), np.array(res.local_range))
    if rank == 0:
        np.save(os.path.join(out_dir, "norms.npy"), res.norms)
        np.save(os.path.join(out_dir, "amps.npy"), res.amplitudes)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 1])
def test_sweep_batched_world2(tmp_path, total):
    """Resolvers are sharded contiguously over the ranks (a rank may hold none) and the gathered norms /
    amplitudes come back in resolver order, equal to a one-process run of every resolver."""
    import torch.multiprocessing as mp
    from mpsim_b200.distributed import shard_range
    port = _free_port()
    mp.spawn(_sweep_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert tuple(np.load(tmp_path / f"range{r}.npy")) == shard_range(total, r, 2)
    from mpsim_b200.mpsim_cirq.simulator import MPSimulator
    n = 5
    sim = MPSimulator({"maxsvals": 4})
    params = [{"t": 0.1 * (i + 1)} for i in range(total)]
    op_lists, nq = sim._translate(_sweep_circuit(n), params, None)
    ref = _OracleBatch(op_lists, nq, {"maxsvals": 4})
    bits = np.array([[0] * n, [1] * n, [1, 0, 1, 0, 1]], dtype=np.uint8)
    np.testing.assert_allclose(np.load(tmp_path / "norms.npy"), ref.norms_device().numpy(), rtol=1e-6)
    np.testing.assert_allclose(np.load(tmp_path / "amps.npy"), ref.amplitudes_device(bits).numpy(), atol=1e-6)
