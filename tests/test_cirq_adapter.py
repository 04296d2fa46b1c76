"""The host-only cases of the reference's mpsim_cirq/circuits_test.py, ported against the duck-typed
circuit protocol of tests/_fake_cirq.py (cirq~=0.8 is not installable here; the adapter only uses
op.qubits / op._has_unitary_() / op._unitary_() / circuit.all_qubits() / all_operations()).
Qubits are plain integers (sortable, hashable -- what the adapter needs of cirq.LineQubit)."""
import numpy as np
import pytest

from tests._fake_cirq import Circuit, CNOT, H, Op

_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Z = np.diag([1, -1]).astype(complex)
_ZZ = np.diag([1, -1, -1, 1]).astype(complex)


def _mp():
    import mpsim_b200
    import mpsim_b200.mpsim_cirq     # attaches MPSOperation.from_gate_operation (circuits.py:46)
    return mpsim_b200


def test_from_gate_operation():                                  # circuits_test.py:12-53
    MPSOperation = _mp().MPSOperation
    op = Op((0,), _X)
    mps_op = MPSOperation.from_gate_operation(op, {0: 0})
    assert mps_op.is_valid() and mps_op.is_single_qudit_operation()
    assert mps_op.qudit_indices == (0,)
    assert np.allclose(mps_op.tensor(), op._unitary_())

    op = Op((0, 1), _ZZ)
    mps_op = MPSOperation.from_gate_operation(op, {0: 0, 1: 1})
    assert mps_op.is_valid() and mps_op.is_two_qudit_operation()
    assert mps_op.qudit_indices == (0, 1)
    assert np.allclose(mps_op.tensor(), op._unitary_())

    op = CNOT(0, 2)
    mps_op = MPSOperation.from_gate_operation(op, {i: i for i in range(3)})
    assert mps_op.is_valid()
    assert mps_op.qudit_indices == (0, 2)
    assert np.allclose(mps_op.tensor(), op._unitary_())


def test_operation_without_unitary_cannot_be_converted():        # circuits.py:32-36
    from mpsim_b200.core import CannotConvertToMPSOperation
    with pytest.raises(CannotConvertToMPSOperation):
        _mp().MPSOperation.from_gate_operation(Op((0,), None), {0: 0})


def test_mpsim_circuit_translation():                            # circuits_test.py:56-120
    from mpsim_b200.mpsim_cirq import MPSimCircuit
    empty = MPSimCircuit(Circuit([]))
    assert len(list(empty.all_qubits())) == 0 and len(list(empty.all_operations())) == 0

    ops = [H(0), Op((0,), _Z), H(0)]
    mps_ops = MPSimCircuit(Circuit(ops))._mps_operations
    assert len(mps_ops) == len(ops)
    for gate_op, mps_op in zip(ops, mps_ops):
        assert np.allclose(gate_op._unitary_(), mps_op.tensor())
        assert mps_op.qudit_indices == (0,) and mps_op.qudit_dimension == 2

    for _ in range(20):
        circuit = MPSimCircuit(Circuit([H(0), CNOT(0, 1)]))
        assert circuit._qudit_to_index_map == {0: 0, 1: 1}
    ops = [H(0), CNOT(0, 1)]
    mps_ops = MPSimCircuit(Circuit(ops))._mps_operations
    assert len(mps_ops) == len(ops)
    for gate_op, mps_op in zip(ops, mps_ops):
        assert np.allclose(gate_op._unitary_(), mps_op.tensor())
        assert mps_op.qudit_dimension == 2
    assert mps_ops[0].qudit_indices == (0,) and mps_ops[1].qudit_indices == (0, 1)
