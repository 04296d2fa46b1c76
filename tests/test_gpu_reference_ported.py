"""The reference's own tests for the path (mpsim/core_test.py), ported one to one and run on the
GPU implementation through its reference-shaped API.  Every test cites the lines it follows;
assertions are the reference's, with np.allclose's absolute tolerance at 1e-6 (complex64 on the
device; the reference computes in complex128 and uses the 1e-8 default)."""
from copy import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _close(a, b):
    return np.allclose(a, b, atol=1e-6)


def _mp():
    import mpsim_b200
    return mpsim_b200


def test_mps_one_qudit_and_valid_product_states():              # core_test.py:71-83
    MPS = _mp().MPS
    for d in (2, 3, 10, 20, 100):
        with pytest.raises(ValueError):
            MPS(nqudits=1, qudit_dimension=d)
    for n in (2, 3, 7, 19):
        for d in (2, 3, 9):
            assert MPS(nqudits=n, qudit_dimension=d).is_valid()


def test_max_bond_dimensions():                                  # core_test.py:86-154
    MPS = _mp().MPS
    assert MPS(nqudits=5)._max_bond_dimensions == [2, 4, 4, 2]
    assert MPS(nqudits=7)._max_bond_dimensions == [2, 4, 8, 8, 4, 2]
    assert MPS(nqudits=6)._max_bond_dimensions == [2, 4, 8, 4, 2]
    assert MPS(nqudits=8)._max_bond_dimensions == [2, 4, 8, 16, 8, 4, 2]
    d = 4
    mps = MPS(nqudits=5, qudit_dimension=d)
    assert mps._max_bond_dimensions == [4, 16, 16, 4]
    assert mps.wavefunction().shape == (d ** 5,)
    mps = MPS(nqudits=7, qudit_dimension=d)
    assert mps._max_bond_dimensions == [4, 16, 64, 64, 16, 4]
    assert mps.wavefunction().shape == (d ** 7,)
    d = 10
    mps = MPS(nqudits=4, qudit_dimension=d)
    assert mps._max_bond_dimensions == [10, 100, 10]
    assert mps.wavefunction().shape == (d ** 4,)
    mps = MPS(nqudits=6, qudit_dimension=d)
    assert mps._max_bond_dimensions == [10, 100, 1000, 100, 10]
    assert mps.wavefunction().shape == (d ** 6,)
    mps = MPS(nqudits=10)
    assert [mps.max_bond_dimension_of(i) for i in (0, -1, 3, 4, 5)] == [2, 2, 16, 32, 16]
    mps = MPS(nqudits=6, qudit_dimension=d)
    assert [mps.max_bond_dimension_of(i) for i in (0, 1, 2, 3, -1)] == [d, d ** 2, d ** 3, d ** 2, d]


def test_bond_dimensions_product_state_and_qutrit_wavefunction():    # core_test.py:157-165, 299-313
    MPS = _mp().MPS
    n = 5
    for d in range(3, 10):
        mps = MPS(nqudits=n, qudit_dimension=d)
        assert mps.bond_dimensions() == [1] * (n - 1)
    mps = MPS(nqudits=3)
    assert isinstance(mps.wavefunction(), np.ndarray)
    assert mps.wavefunction().shape == (8,)
    assert _close(mps.wavefunction(), np.array([1.0] + [0.0] * 7, dtype=np.complex64))
    mps = MPS(nqudits=3, qudit_dimension=3)
    assert mps.wavefunction().shape == (27,)
    assert _close(mps.wavefunction(), [1] + [0] * 26)
    assert mps.is_valid()


@pytest.mark.parametrize("left", [True, False])
def test_apply_twoq_cnot_and_swap_five_qubits(left):             # core_test.py:523-617
    MPS = _mp().MPS
    n = 5
    for a in range(n - 1):
        b = a + 1
        mps = MPS(n)
        mps.x(a)
        mps.cnot(a, b, keep_left_canonical=left)
        correct = np.zeros((2 ** n,))
        bits = ["0"] * n
        bits[a] = bits[b] = "1"
        correct[int("".join(bits), 2)] = 1.0
        assert _close(mps.wavefunction(), correct)
    mps = MPS(nqudits=2)
    mps.x(0)
    mps.swap(0, 1, keep_left_canonical=left)
    assert _close(mps.wavefunction(), [0.0, 1.0, 0.0, 0.0])
    mps = MPS(nqudits=2)
    mps.swap(0, 1, keep_left_canonical=left)
    assert _close(mps.wavefunction(), [1.0, 0.0, 0.0, 0.0])
    for i in range(n - 1):
        mps = MPS(n)
        mps.x(i)
        mps.swap(i, i + 1, keep_left_canonical=left)
        correct = np.zeros((2 ** n,))
        bits = ["0"] * n
        bits[i + 1] = "1"
        correct[int("".join(bits), 2)] = 1.0
        assert _close(mps.wavefunction(), correct)


def test_move_node_three_qubits():                               # core_test.py:631-664
    MPS = _mp().MPS
    mps = MPS(nqudits=3, qudit_dimension=2)
    mps.x(0)
    mps.move_node_from_left_to_right(0, 1)
    assert _close(mps.wavefunction(), [0., 0., 1., 0., 0., 0., 0., 0.])
    mps = MPS(nqudits=3, qudit_dimension=2)
    mps.x(2)
    mps.move_node_from_right_to_left(2, 0)
    assert _close(mps.wavefunction(), [0., 0., 0., 0., 1., 0., 0., 0.])
    mps = MPS(nqudits=3, qudit_dimension=2)
    mps.h(0)
    mps.move_node_from_left_to_right(0, 1)
    assert _close(mps.wavefunction(), np.array([1., 0., 1., 0., 0., 0., 0., 0.]) / np.sqrt(2))
    mps = MPS(nqudits=3, qudit_dimension=2)
    mps.h(2)
    mps.move_node_from_right_to_left(2, 1)
    assert _close(mps.wavefunction(), np.array([1., 0., 1., 0., 0., 0., 0., 0.]) / np.sqrt(2))


def test_move_node_ten_qubits_and_errors():                      # core_test.py:667-712
    MPS = _mp().MPS
    n = 10
    mps = MPS(nqudits=n, qudit_dimension=2)
    mps.x(0)
    mps.move_node_from_left_to_right(0, 4)
    correct = np.zeros((2 ** n,)); correct[2 ** 5] = 1.
    assert _close(mps.wavefunction(), correct)
    mps.move_node_from_left_to_right(4, 9)
    correct = np.zeros((2 ** n,)); correct[1] = 1.
    assert _close(mps.wavefunction(), correct)
    mps = MPS(nqudits=n, qudit_dimension=2)
    mps.x(9)
    mps.move_node_from_right_to_left(9, 5)
    correct = np.zeros((2 ** n,)); correct[2 ** 4] = 1.
    assert _close(mps.wavefunction(), correct)
    mps.move_node_from_right_to_left(5, 0)
    correct = np.zeros((2 ** n,)); correct[2 ** (n - 1)] = 1.
    assert _close(mps.wavefunction(), correct)
    mps = MPS(nqudits=5)
    with pytest.raises(ValueError):
        mps.move_node_from_left_to_right(current_node_index=4, final_node_index=0)
    with pytest.raises(ValueError):
        mps.move_node_from_right_to_left(current_node_index=0, final_node_index=4)


def test_move_node_then_apply_two_qubit_gate():                  # core_test.py:715-770
    MPS = _mp().MPS
    n = 5
    mps = MPS(nqudits=n)
    mps.x(0)
    correct = np.zeros(shape=(2 ** n,)); correct[16] = 1.
    assert _close(mps.wavefunction(), correct)
    mps.swap(3, 4)
    assert _close(mps.wavefunction(), correct)
    mps.move_node_from_left_to_right(0, 3)
    mps.swap(3, 4)
    correct = np.zeros(shape=(2 ** n,)); correct[1] = 1.
    assert _close(mps.wavefunction(), correct)
    mps = MPS(nqudits=n)
    mps.x(n - 1)
    correct = np.zeros(shape=(2 ** n,)); correct[1] = 1.
    assert _close(mps.wavefunction(), correct)
    mps.swap(0, 1)
    assert _close(mps.wavefunction(), correct)
    mps.move_node_from_right_to_left(4, 1)
    mps.swap(0, 1)
    correct = np.zeros(shape=(2 ** n,)); correct[2 ** (n - 1)] = 1.
    assert _close(mps.wavefunction(), correct)
    for n in range(3, 10 + 1):
        mps = MPS(nqudits=n)
        mps.x(0)
        mps.move_node_from_left_to_right(0, n - 2)
        mps.cnot(n - 2, n - 1)
        mps.move_node_from_right_to_left(n - 2, 0)
        correct = np.zeros(shape=(2 ** n,)); correct[2 ** (n - 1) + 1] = 1.
        assert _close(mps.wavefunction(), correct)


@pytest.mark.parametrize("left", [True, False])
def test_twoq_gates_in_succession_and_validity(left):           # core_test.py:784-887
    MPS = _mp().MPS
    mps = MPS(2)
    mps.x(0)
    mps.h(-1)
    mps.cnot(0, 1, keep_left_canonical=left)
    mps.h(-1)
    mps.cnot(0, 1, keep_left_canonical=left)
    mps.x(0)
    assert _close(mps.wavefunction(), [0.0, 1.0, 0.0, 0.0])
    mps = MPS(2)
    mps.x(1)
    mps.h(-1)
    mps.cnot(0, 1, keep_left_canonical=False)
    mps.h(-1)
    mps.h(-1)
    mps.cnot(0, 1, keep_left_canonical=False)
    mps.h(-1)
    mps.cnot(0, 1, keep_left_canonical=False)
    mps.cnot(0, 1, keep_left_canonical=False)
    assert mps.is_valid()
    mps = MPS(3)
    mps.x(0)
    mps.cnot(0, 1, keep_left_canonical=left)
    assert mps.is_valid()
    mps.h(0); mps.h(1)
    mps.cnot(0, 1, keep_left_canonical=left)
    mps.h(0); mps.h(1)
    assert mps.is_valid()
    mps = MPS(3)
    mps.x(2)
    mps.cnot(1, 2, keep_left_canonical=True)
    mps.cnot(0, 1, keep_left_canonical=True)
    assert mps.is_valid()


def test_keep_half_bond_dimension_singular_values():             # core_test.py:947-973
    mp = _mp()
    mps = mp.MPS(nqudits=4)
    assert mps.bond_dimensions() == [1, 1, 1]
    assert mps.max_bond_dimensions() == [2, 4, 2]
    mps.r(-1)
    mps.apply_two_qudit_gate(mp.cnot(), 0, 1, fraction=1)
    assert mps.bond_dimensions() == [2, 1, 1]
    mps = mp.MPS(nqudits=4)
    mps.r(-1)
    mps.apply_two_qudit_gate(mp.cnot(), 0, 1, fraction=0.5)
    assert mps.bond_dimensions() == [1, 1, 1]


@pytest.mark.parametrize("chi", [1, 2, 4, 8, 16])
def test_max_bond_dimension_not_surpassed(chi):                  # core_test.py:1471-1497
    mp = _mp()
    np.random.seed(chi)
    nqubits = depth = 10
    mps = mp.MPS(nqudits=nqubits, qudit_dimension=2)
    singles = (mp.hgate(), mp.xgate(), mp.zgate())
    czgate = mp.cphase(exp=0.5)
    for _ in range(depth):
        for i in range(nqubits):
            mps.apply(mp.MPSOperation(singles[np.random.randint(3)], (i,)))
        for i in range(nqubits):
            j = int(np.random.choice(list(set(range(nqubits)) - {i})))
            mps.apply(mp.MPSOperation(czgate, (i, j)), maxsvals=chi)
        assert all(bond_dimension <= chi for bond_dimension in mps.bond_dimensions())


def test_equal_and_copy():                                       # core_test.py:1500-1542
    mp = _mp()
    for n in (2, 3, 5, 10):
        for d in (2, 3, 5, 10):
            mps1 = mp.MPS(nqudits=n, qudit_dimension=d)
            mps2 = mp.MPS(nqudits=n, qudit_dimension=d)
            assert mps1 == mps1 and mps2 == mps2 and mps1 == mps2
            if d == 2:
                mps1.apply(mp.MPSOperation(mp.xgate(), 0))
                assert mps1 != mps2
                mps2.apply(mp.MPSOperation(mp.xgate(), 0))
                assert mps1 == mps2
            for mps_copy in (copy(mps1), mps1.copy()):
                assert mps_copy is not mps1
                assert mps_copy == mps1
    assert mp.MPS(nqudits=10, qudit_dimension=2, tensor_prefix="mps1_") == \
        mp.MPS(nqudits=10, qudit_dimension=2, tensor_prefix="mps2_")


def test_get_free_edge_of():                                     # core_test.py:165-173
    MPS = _mp().MPS
    for n in range(2, 10):
        for d in (2, 3, 4):
            mps = MPS(nqudits=n, qudit_dimension=d)
            for i in range(n):
                free_edge = mps.get_free_edge_of(i, copy=False)
                assert free_edge.is_dangling()
                assert free_edge.node1.name == f"q{i}"


def test_get_left_and_right_connected_edges():                   # core_test.py:176-227
    MPS = _mp().MPS
    for d in (2, 3, 4):
        mps = MPS(nqudits=3, qudit_dimension=d)
        assert mps.get_left_connected_edge_of(0) is None
        edge = mps.get_left_connected_edge_of(1)
        assert not edge.is_dangling()
        assert (edge.node1.name, edge.node2.name) == ("q0", "q1")
        edge = mps.get_left_connected_edge_of(2)
        assert not edge.is_dangling()
        assert (edge.node1.name, edge.node2.name) == ("q2", "q1")
        edge = mps.get_right_connected_edge_of(0)
        assert not edge.is_dangling()
        assert (edge.node1.name, edge.node2.name) == ("q0", "q1")
        edge = mps.get_right_connected_edge_of(1)
        assert not edge.is_dangling()
        assert (edge.node1.name, edge.node2.name) == ("q2", "q1")
        assert mps.get_right_connected_edge_of(2) is None
    n = 10
    for d in (2, 3, 4):
        mps = MPS(nqudits=n, qudit_dimension=d)
        for i in range(1, n - 1):
            assert mps.get_right_connected_edge_of(i - 1) == mps.get_left_connected_edge_of(i)


def _gates_mod():
    """mpsim.gates of whichever package _mp() returns (the reference keeps them in mpsim.gates)."""
    import importlib
    return importlib.import_module(_mp().__name__ + ".gates")


def test_mps_operation_properties():                             # core_test.py:25-67
    mp, g = _mp(), _gates_mod()
    op = mp.MPSOperation(g.igate(), qudit_indices=0, qudit_dimension=2)
    assert op.qudit_indices == (0,) and op.qudit_dimension == 2
    assert op.is_valid() and op.is_unitary()
    assert op.is_single_qudit_operation() and not op.is_two_qudit_operation()

    np.random.seed(1)
    tensor = np.random.randn(2, 2)
    node = type(g.igate())(tensor)
    op = mp.MPSOperation(node, qudit_indices=(0,), qudit_dimension=2)
    assert len(node.edges) == len(op.node(copy=True).edges)
    assert _close(tensor, op.tensor())

    op = mp.MPSOperation(g.cnot(), qudit_indices=(0, 1), qudit_dimension=2)
    assert op.qudit_indices == (0, 1) and op.qudit_dimension == 2
    assert not op.is_single_qudit_operation() and op.is_two_qudit_operation()

    op = mp.MPSOperation(g.cnot(), qudit_indices=(0, 2), qudit_dimension=2)
    assert op.qudit_indices == (0, 2) and op.is_valid()
    assert not op.is_single_qudit_operation() and op.is_two_qudit_operation()


@pytest.mark.parametrize("left", [True, False])
def test_apply_twoq_cnot_four_qubits(left):                      # core_test.py:488-520
    MPS = _mp().MPS
    for prepare, cnot, index in (((1,), (1, 2), 6), ((), (1, 2), 0), ((2,), (2, 3), 3), ((0,), (0, 1), 12)):
        mps = MPS(nqudits=4)
        for q in prepare:
            mps.x(q)
        mps.cnot(*cnot, keep_left_canonical=left)
        correct = np.zeros(16)
        correct[index] = 1.0
        assert _close(mps.wavefunction(), correct)


@pytest.mark.parametrize("left", [True, False])
def test_qubit_hopping_left_to_right(left):                      # core_test.py:618-628
    n = 8
    mps = _mp().MPS(n)
    mps.h(0)
    for i in range(1, n - 1):
        mps.swap(i, i + 1, keep_left_canonical=left)
    correct = np.zeros(2 ** n)
    correct[0] = correct[2 ** (n - 1)] = 1.0 / np.sqrt(2)
    assert _close(mps.wavefunction(), correct)


def test_valid_after_orthonormalize_right_edges():               # core_test.py:1261-1285
    mp, g = _mp(), _gates_mod()
    n = 3
    mps = mp.MPS(nqudits=n)
    mps.apply([mp.MPSOperation(g.hgate(), (i,)) for i in range(n)])
    before = mps.wavefunction()
    assert [mps.bond_dimension_of(0), mps.bond_dimension_of(1)] == [1, 1]
    for site in (0, 1):
        mps.orthonormalize_right_edge_of(site)
        assert mps.is_valid()
        assert [mps.bond_dimension_of(0), mps.bond_dimension_of(1)] == [1, 1]
        assert _close(mps.wavefunction(), before)


def test_apply_povm_product_state():                             # core_test.py:1288-1338
    mp, g = _mp(), _gates_mod()
    pi0 = g.computational_basis_projector(state=0)
    n = 3
    mps = mp.MPS(nqudits=n)
    mps.apply([mp.MPSOperation(g.hgate(), i) for i in range(n)])          # |+++>
    assert np.isclose(mps.norm(), 1.0, atol=1e-6)
    assert mps.bond_dimensions() == [1, 1]
    for site in range(n):                   # |0><0| on one more qubit each time, nothing else
        mps.apply_one_qudit_gate(pi0, site, ortho_after_non_unitary=False, renormalize_after_non_unitary=False)
        assert mps.is_valid()
        assert np.isclose(mps.norm(), np.sqrt(0.5) ** (site + 1), atol=1e-6)
        assert mps.bond_dimensions() == [1, 1]
        ones = 2 ** (n - 1 - site)
        correct = np.sqrt(0.5) ** 3 * np.array([1] * ones + [0] * (8 - ones))
        assert _close(mps.wavefunction(), correct)


def test_orthonormalize_all_tensors_edge_cases():                # core_test.py:1431-1448
    MPS = _mp().MPS
    for n in range(2, 8):
        for d in (2, 3, 4):
            mps = MPS(nqudits=n, qudit_dimension=d)
            correct = mps.wavefunction()
            for site in range(n - 1):
                mps.orthonormalize_right_edge_of(site)
                assert mps.is_valid()
                assert _close(mps.wavefunction(), correct)
                assert np.isclose(mps.norm(), 1.0, atol=1e-6)
            for site in range(1, n):
                mps.orthonormalize_left_edge_of(site)
                assert mps.is_valid()
                assert _close(mps.wavefunction(), correct)
                assert np.isclose(mps.norm(), 1.0, atol=1e-6)
