"""Cirq operation -> ``MPSOperation`` translation (``mpsim/mpsim_cirq/circuits.py:12-46``)."""
from typing import Any, Dict, List

import numpy as np

from mpsim_b200.core import MPSOperation, CannotConvertToMPSOperation
from mpsim_b200.node import Node


def mps_operation_from_gate_operation(gate_operation: Any, qudit_to_index_map: Dict[Any, int]) -> MPSOperation:
    """``circuits.py:12-46``: ``op._unitary_()`` reshaped to ``[2]*2k``; operations without a
    unitary raise ``CannotConvertToMPSOperation``."""
    num_qudits = len(gate_operation.qubits)
    qudit_dimension = 2
    qudit_indices = tuple(qudit_to_index_map[q] for q in gate_operation.qubits)
    if not gate_operation._has_unitary_():
        raise CannotConvertToMPSOperation(
            f"Cannot convert operation {gate_operation} into an MPS Operation because the operation "
            "does not have a unitary.")
    tensor = np.reshape(gate_operation._unitary_(), [qudit_dimension] * 2 * num_qudits)
    return MPSOperation(Node(tensor), qudit_indices, qudit_dimension)


MPSOperation.from_gate_operation = staticmethod(mps_operation_from_gate_operation)


class MPSimCircuit:
    """Pre-translated circuit (``circuits.py:49-93``).  Wraps any circuit object of the protocol;
    with real Cirq installed it still behaves as a view of the wrapped ``cirq.Circuit``."""

    def __init__(self, cirq_circuit: Any, device: Any = None, qubit_order: Any = None) -> None:
        self._circuit = cirq_circuit
        self.device = device
        self._qudit_to_index_map = {q: i for i, q in enumerate(sorted(cirq_circuit.all_qubits()))}
        self._mps_operations = self._translate_to_mps_operations()

    def all_qubits(self):
        return self._circuit.all_qubits()

    def all_operations(self):
        return self._circuit.all_operations()

    def __getattr__(self, name: str) -> Any:
        return getattr(self._circuit, name)

    def _translate_to_mps_operations(self) -> List[MPSOperation]:
        return [mps_operation_from_gate_operation(op, self._qudit_to_index_map)
                for op in self._circuit.all_operations()]
