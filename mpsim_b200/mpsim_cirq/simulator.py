"""``MPSimulator`` (``mpsim/mpsim_cirq/simulator.py:14-88``) on the GPU path.

With Cirq installed it subclasses ``cirq.sim.SimulatesFinalState`` like the reference; without
it (this image) it is a plain class with the same ``simulate`` / ``simulate_sweep`` surface that
accepts any circuit object implementing ``all_qubits()`` / ``all_operations()`` (and,
optionally, ``_resolve_parameters_``)."""
from typing import Any, List

from mpsim_b200.core import MPS
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


class MPSimulator(_Base):
    def __init__(self, options: dict = {}):   # noqa: B006  (signature of simulator.py:16)
        """``options``: ``maxsvals`` (int) or ``fraction`` (float), forwarded to every two-qudit
        gate (``simulator.py:16-30, 86``)."""
        self._options = options

    def simulate(self, program: Any, param_resolver: Any = None, qubit_order: Any = None,
                 initial_state: Any = None) -> MPS:
        return self.simulate_sweep(program, param_resolver, qubit_order, initial_state)[0]

    def simulate_sweep(self, program: Any, params: Any = None, qubit_order: Any = None,
                       initial_state: Any = None) -> List[MPS]:
        """``simulator.py:32-88``: one ``MPS`` per parameter resolver."""
        if not _is_circuit(program):
            raise ValueError(f"Program is of type {type(program)} but should be either a cirq.Circuit or "
                             "mpsim.mpsim_cirq.MPSimCircuit.")
        trial_results = []
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
            mps = MPS(nqudits=len(qubits))
            operations = [mps_operation_from_gate_operation(op, qubit_to_index_map)
                          for op in solved.all_operations()]
            mps.apply(operations, **self._options)
            trial_results.append(mps)
        return trial_results
