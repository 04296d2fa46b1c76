"""Cirq adapter (``mpsim/mpsim_cirq``).  Cirq itself is optional: circuits are accepted through
a duck-typed protocol (``all_qubits()``, ``all_operations()``, ``op.qubits``,
``op._has_unitary_()``, ``op._unitary_()``)."""
from mpsim_b200.mpsim_cirq.circuits import mps_operation_from_gate_operation, MPSimCircuit
from mpsim_b200.mpsim_cirq.simulator import MPSimulator

__all__ = ["MPSimulator", "MPSimCircuit", "mps_operation_from_gate_operation"]
