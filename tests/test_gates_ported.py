"""The reference's gates_test.py, ported (host only: mpsim_b200.gates is numpy).  `_g()` is the
module under test; oracle/run_ported_tests_on_reference.py points it at the reference's mpsim.gates
to check the port itself.  Every test cites the lines it follows."""
import numpy as np
import pytest


def _g():
    import mpsim_b200.gates
    return mpsim_b200.gates


def test_common_gates_are_unitary():                             # gates_test.py:21-27
    g = _g()
    for gate in (g.igate(), g.hgate(), g.xgate(), g.ygate(), g.zgate(), g.cnot()):
        assert g.is_unitary(gate)
    for exp in np.linspace(start=0, stop=2 * np.pi, num=100):
        assert g.is_unitary(g.cphase(exp))


@pytest.mark.parametrize("dim", [2, 3])
def test_computational_basis_projectors(dim):                    # gates_test.py:30-47, 65-75
    g = _g()
    for state in range(dim):
        projector = g.computational_basis_projector(state, dim)
        correct = np.zeros((dim, dim))
        correct[state, state] = 1.0
        assert np.array_equal(projector.tensor, correct)
        assert g.is_projector(projector)
        assert not g.is_unitary(projector)
        assert str(projector) == f"|{state}><{state}|"


def test_invalid_projectors():                                   # gates_test.py:50-62
    g = _g()
    with pytest.raises(ValueError):
        g.computational_basis_projector(state=-1)
    with pytest.raises(ValueError):
        g.computational_basis_projector(state=2, dim=-1)
    with pytest.raises(ValueError):
        g.computational_basis_projector(state=10, dim=8)


def test_haar_random_unitary():                                  # gates_test.py:78-83
    g = _g()
    for n in (2, 3, 4):
        for d in (2, 3, 5):
            assert g.is_unitary(g.haar_random_unitary(nqudits=n, qudit_dimension=d, seed=1))
