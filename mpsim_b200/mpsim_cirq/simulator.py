"""``MPSimulator`` (``mpsim/mpsim_cirq/simulator.py:14-88``) on the GPU path.

With Cirq installed it subclasses ``cirq.sim.SimulatesFinalState`` like the reference; without
it (this image) it is a plain class with the same ``simulate`` / ``simulate_sweep`` surface that
accepts any circuit object implementing ``all_qubits()`` / ``all_operations()`` (and,
optionally, ``_resolve_parameters_``)."""
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from mpsim_b200.core import MPS, MPSOperation
from mpsim_b200 import gates as _gates
from mpsim_b200.node import tensor_of
from mpsim_b200.mpsim_cirq.circuits import MPSimCircuit, mps_operation_from_gate_operation

try:  # pragma: no cover - cirq is absent in this image
    import cirq as _cirq
    _Base = _cirq.sim.SimulatesFinalState
except Exception:  # noqa: BLE001
    _cirq = None
    _Base = object


def _is_circuit(program: Any) -> bool:
    if isinstance(program, MPSimCircuit):
        return True
    if _cirq is not None and isinstance(program, _cirq.Circuit):
        return True
    return _cirq is None and hasattr(program, "all_qubits") and hasattr(program, "all_operations")


def _resolvers(params: Any) -> List[Any]:
    if _cirq is not None:
        return list(_cirq.study.to_resolvers(params))
    if params is None:
        return [None]
    if isinstance(params, dict):
        return [params]
    return list(params)


def _resolve(program: Any, resolver: Any) -> Any:
    if _cirq is not None:
        return _cirq.protocols.resolve_parameters(program, resolver)
    if resolver is None or not hasattr(program, "_resolve_parameters_"):
        return program
    return program._resolve_parameters_(resolver)


class SweepResult:
    """What ``MPSimulator.simulate_sweep_batched`` returns: the rank's slice of the sweep as one
    ``MPSBatch`` plus the per-resolver results gathered over the process group (if any)."""

    def __init__(self, batch: Any, local_range: Any, total: int, norms: np.ndarray,
                 amplitudes: Optional[np.ndarray]) -> None:
        self.batch = batch                  #: MPSBatch holding resolvers local_range[0]:local_range[1]
        self.local_range = local_range
        self.total = total
        self.norms = norms                  #: float32 [total]
        self.amplitudes = amplitudes        #: complex64 [total][nbits] or None

    def mps(self, index: int) -> MPS:
        """Resolver ``index`` (must be local to this rank) as an ``MPS`` of its own."""
        lo, hi = self.local_range
        if not lo <= index < hi:
            raise IndexError(f"resolver {index} lives on another rank (this rank holds [{lo}, {hi}))")
        return self.batch.member(index - lo)


def _sweep_structure(op_lists: Sequence[Sequence[MPSOperation]]) -> Optional[List[Any]]:
    """The common structure [(indices, nqudits-of-gate)] of the resolved circuits, or None when they
    differ, contain a non-unitary one-qudit gate (sequential path: orthonormalise + renormalise,
    ``mpsim/core.py:816-845``) or a gate on three or more qudits."""
    first = [op.qudit_indices for op in op_lists[0]]
    for ops in op_lists:
        if [op.qudit_indices for op in ops] != first:
            return None
        for op in ops:
            if not op.is_valid() or len(op.qudit_indices) not in (1, 2):
                return None
            if op.is_single_qudit_operation() and not _gates.is_unitary(tensor_of(op.node(copy=False))):
                return None
    return first


class MPSimulator(_Base):
    def __init__(self, options: dict = {}):   # noqa: B006  (signature of simulator.py:16)
        """``options``: ``maxsvals`` (int) or ``fraction`` (float), forwarded to every two-qudit
        gate (``simulator.py:16-30, 86``).  ``batch_sweeps`` (default True; not forwarded): run the
        resolvers of a sweep as ONE batched simulation when their circuits share a structure."""
        self._options = options

    def _gate_options(self) -> Dict[str, Any]:
        return {k: v for k, v in self._options.items() if k != "batch_sweeps"}

    def simulate(self, program: Any, param_resolver: Any = None, qubit_order: Any = None,
                 initial_state: Any = None) -> MPS:
        return self.simulate_sweep(program, param_resolver, qubit_order, initial_state)[0]

    def simulate_sweep(self, program: Any, params: Any = None, qubit_order: Any = None,
                       initial_state: Any = None) -> List[MPS]:
        """``simulator.py:32-88``: one ``MPS`` per parameter resolver."""
        if not _is_circuit(program):
            raise ValueError(f"Program is of type {type(program)} but should be either a cirq.Circuit or "
                             "mpsim.mpsim_cirq.MPSimCircuit.")
        op_lists, nqubits = self._translate(program, params, qubit_order)
        if len(op_lists) > 1 and self._options.get("batch_sweeps", True) and _sweep_structure(op_lists) is not None:
            batch = self._run_batch(op_lists, nqubits)
            return [batch.member(b) for b in range(len(op_lists))]
        trial_results = []
        for operations in op_lists:
            mps = MPS(nqudits=nqubits)
            mps.apply(operations, **self._gate_options())
            trial_results.append(mps)
        return trial_results

    def simulate_sweep_batched(self, program: Any, params: Any = None, qubit_order: Any = None,
                               amplitudes: Any = None, group: Any = None) -> SweepResult:
        """The resolvers of ``simulator.py:67-87`` as one batched simulation, sharded over the
        ``torch.distributed`` process group when one is initialised (contiguous slice per rank,
        ``distributed.shard_range``; no traffic while simulating).  Norms and, if ``amplitudes``
        (bitstrings ``[nbits][nqubits]``) is given, those amplitudes are all-gathered at the end
        (SURVEY.md 8(e)).  The circuits must share a structure (same gates on the same qubits --
        what a parameter sweep is); anything else raises ``ValueError``."""
        import torch
        import torch.distributed as dist
        from mpsim_b200.distributed import gather_slices, shard_range
        if not _is_circuit(program):
            raise ValueError(f"Program is of type {type(program)} but should be either a cirq.Circuit or "
                             "mpsim.mpsim_cirq.MPSimCircuit.")
        op_lists, nqubits = self._translate(program, params, qubit_order)
        if _sweep_structure(op_lists) is None:
            raise ValueError("simulate_sweep_batched needs resolved circuits of one structure with unitary gates; "
                             "use simulate_sweep for anything else.")
        total = len(op_lists)
        rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
        lo, hi = shard_range(total, rank, world)
        batch = self._run_batch(op_lists[lo:hi], nqubits) if hi > lo else None
        if batch is not None:
            dev = batch._chain.device
        else:       # a rank without members still takes part in the gathers
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        norms = batch.norms_device() if batch is not None else torch.zeros(0, dtype=torch.float32, device=dev)
        norms = gather_slices(norms, total, group)
        amps = None
        if amplitudes is not None:
            bits = np.asarray(amplitudes, dtype=np.uint8).reshape(-1, nqubits)
            local = batch.amplitudes_device(bits) if batch is not None else \
                torch.zeros((0, bits.shape[0]), dtype=torch.complex64, device=dev)
            amps = gather_slices(local, total, group).cpu().numpy()
        return SweepResult(batch, (lo, hi), total, norms.cpu().numpy(), amps)

    # ------------------------------------------------------------------ helpers
    def _translate(self, program: Any, params: Any, qubit_order: Any):
        """Resolve and translate every resolver's circuit (``simulator.py:67-84``)."""
        op_lists, nqubits = [], 0
        for prs in _resolvers(params):
            solved = _resolve(program, prs)
            qubits = solved.all_qubits()
            if _cirq is not None and qubit_order is not None:
                ordered = _cirq.ops.QubitOrder.as_qubit_order(qubit_order).order_for(qubits)
            elif qubit_order is not None and not callable(qubit_order):
                ordered = list(qubit_order)
            else:
                ordered = sorted(qubits)
            qubit_to_index_map = {q: i for i, q in enumerate(ordered)}
            nqubits = len(qubits)
            op_lists.append([mps_operation_from_gate_operation(op, qubit_to_index_map)
                             for op in solved.all_operations()])
        return op_lists, nqubits

    def _run_batch(self, op_lists: Sequence[Sequence[MPSOperation]], nqubits: int):
        """Compile the shared structure once, stage every resolver's gates, run."""
        from mpsim_b200.batch import MPSBatch
        from mpsim_b200.circuits import Op
        nb, nops = len(op_lists), len(op_lists[0])
        batch = MPSBatch(nb, nqubits)
        structure = [Op(tensor_of(op.node(copy=False)), op.qudit_indices, True) for op in op_lists[0]]
        cp = batch.compile(structure, **self._gate_options())
        gates = np.zeros((nops, nb, 16), dtype=np.complex64)
        for b, ops in enumerate(op_lists):
            for t, op in enumerate(ops):
                flat = np.asarray(tensor_of(op.node(copy=False))).reshape(-1)
                gates[t, b, :flat.size] = flat
        batch.stage_gates(cp, gates)
        batch.run(cp)
        return batch
