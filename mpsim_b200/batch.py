"""Batches of independent circuits / trajectories on one GPU (BASELINE.json config 4).

All members share one circuit STRUCTURE (which sites each gate touches, canonical form,
maxsvals) and differ in their gate matrices.  Because kept singular-value counts are data
independent, every member has identical tensor shapes at every step, so each step of the
plan is one uniform batched launch over (applications of the layer) x (batch members).
Across GPUs the batch is sharded by contiguous slices with no traffic during the simulation
(``mpsim_b200.distributed``).
"""
from typing import Any, Dict, Optional, Sequence

import numpy as np

from mpsim_b200.planner import SWAP_TENSOR, plan_operations
from mpsim_b200.store import CompiledPlan, DeviceChain


class MPSBatch:
    def __init__(self, nbatch: int, nqudits: int, qudit_dimension: int = 2, device: Any = None) -> None:
        if nqudits < 2:
            raise ValueError(f"Number of qudits must be greater than 2 but is {nqudits}.")
        if nbatch < 1:
            raise ValueError("nbatch must be positive.")
        self.nbatch, self.nqudits, self.qudit_dimension = int(nbatch), int(nqudits), int(qudit_dimension)
        self._chain = DeviceChain(nqudits, qudit_dimension, nbatch, device)

    # ------------------------------------------------------------------ planning
    def compile(self, ops: Sequence[Any], record_svals: bool = False, **kwargs: Any) -> CompiledPlan:
        """``ops``: circuit structure as ``circuits.Op`` (tensor, indices, keep_left_canonical);
        the tensors only fix shapes here (and are staged as shared default gates).  ``kwargs``
        are the reference's truncation options (``maxsvals`` / ``fraction``)."""
        triples = []
        for op in ops:
            kw = dict(kwargs)
            if len(op.indices) == 2 and not op.keep_left_canonical:
                kw["keep_left_canonical"] = False
            triples.append((np.asarray(op.tensor), tuple(op.indices), kw))
        plan = plan_operations(self.nqudits, self.qudit_dimension, self._chain.bonds, triples)
        return self._chain.compile(plan, per_batch_gates=True, record_svals=record_svals)

    def stage_gates(self, cp: CompiledPlan, gates: np.ndarray) -> None:
        """Fill the pinned staging buffer from per-member gates ``[nops][nbatch][d**(2 nq)]``
        (ragged rows zero padded to ``d**4``), inserting SWAPs and control/target flips where
        the plan did (``mpsim/core.py:1031-1033``)."""
        import torch
        d = self.qudit_dimension
        width = d ** 4
        gates = np.asarray(gates)
        assert gates.shape[1] == self.nbatch and gates.shape[2] == width
        host = cp.gates_host.numpy()                   # [ngates][B][width], pinned
        src = np.array([s for s, _ in cp.plan.gate_src], dtype=np.int64)
        flip = np.array([f for _, f in cp.plan.gate_src], dtype=bool)
        user = src >= 0
        if user.any():
            host[user] = gates[src[user]]
        if (~user).any():
            host[~user] = SWAP_TENSOR.reshape(1, 1, -1).astype(np.complex64)
        for g in np.nonzero(flip)[0]:
            host[g] = host[g].reshape(-1, d, d, d, d).transpose(0, 2, 1, 4, 3).reshape(-1, width)
        del torch

    def run(self, cp: CompiledPlan, upload: bool = True) -> None:
        self._chain.run(cp, upload=upload)

    def reset(self) -> None:
        self._chain.reset()

    # ------------------------------------------------------------------ results
    def bond_dimensions(self):
        return list(self._chain.bonds[1:-1])

    def norms_device(self):
        return self._chain.norms()

    def norms(self) -> np.ndarray:
        self._chain.check_status()
        return self._chain.norms().cpu().numpy()

    def amplitudes_device(self, bitstrings):
        return self._chain.amplitudes(bitstrings)

    def amplitudes(self, bitstrings) -> np.ndarray:
        self._chain.check_status()
        return self._chain.amplitudes(bitstrings).cpu().numpy()

    def wavefunction(self, member: int) -> np.ndarray:
        self._chain.check_status()
        return self._chain.wavefunction(int(member)).cpu().numpy()

    def member(self, b: int):
        """Member ``b`` as an ``MPS`` of its own (a copy of its site tensors)."""
        from mpsim_b200.core import MPS
        self._chain.check_status()
        return MPS._from_chain(self._chain.member(b))

    def renormalize(self, to_norm: float = 1.0) -> None:
        """Per-member ``MPS.renormalize`` (``mpsim/core.py:567-594``); norms stay on the device."""
        import torch
        norms = self._chain.norms()
        if bool((norms < 1e-15).any().item()):
            raise ValueError("Norm of MPS is numerically zero, cannot renormalize.")
        self._chain.scale(torch.pow(float(to_norm) / norms, 1.0 / self.nqudits))

    def singular_values(self, cp: CompiledPlan) -> Optional[np.ndarray]:
        """[napplications (program order)][nbatch][width] if the plan was compiled with record_svals."""
        if cp.svals is None:
            return None
        sv = cp.svals.cpu().numpy()
        out = np.empty_like(sv)
        for p, idx in enumerate(cp.order2):
            out[idx] = sv[p]
        return out

    def status(self, cp: CompiledPlan) -> np.ndarray:
        """int32 [napplications (program order)][nbatch][2] = (status, sweeps)."""
        n2 = len(cp.plan.apps2)
        info = cp.info.cpu().numpy()[: n2 * self.nbatch].reshape(n2, self.nbatch, 2)
        out = np.empty_like(info)
        for p, idx in enumerate(cp.order2):
            out[idx] = info[p]
        return out
