"""Pins the oracle's restatement of SURVEY.md 8(f) rows 2-3 (from_wavefunction, reduced density
matrices, sampling, expectation values, inner products) against vectors from the unmodified
reference, and against the reference's own known-answer tests (ported).  CPU only."""
import numpy as np
import pytest

from oracle.mps_oracle import OracleMPS, HGATE, XGATE
from tests import _observables as obs


class OracleAdapter:
    @staticmethod
    def new(n, d):
        return OracleMPS(n, d, dtype=np.complex128)

    @staticmethod
    def from_wavefunction(wf, n, d):
        return OracleMPS.from_wavefunction(wf, n, d, dtype=np.complex128)

    @staticmethod
    def apply1(mps, tensor, i):
        mps.apply_one_qudit_gate(tensor, i)

    @staticmethod
    def apply2(mps, tensor, i, j, **kw):
        mps.apply_two_qudit_gate(tensor, i, j, **kw)

    @staticmethod
    def expectation(mps, tensor, indices):
        return mps.expectation(tensor, indices)


@pytest.mark.parametrize("name", obs.STATES)
def test_oracle_observables_match_reference(name):
    obs.check_state(OracleAdapter, name, atol=1e-10)


def test_oracle_inner_products_match_reference():
    obs.check_inner_products(OracleAdapter, atol=1e-10)


def test_oracle_from_wavefunction_matches_reference():
    obs.check_from_wavefunction(OracleAdapter, atol=1e-10)
    # the reference hands sqrt(S) to both sides of every cut (tn.split_node, core.py:311-316):
    # the left-most node is U sqrt(S) up to the per-column phase LAPACK picks
    for t in range(int(obs.Z["n_fw"])):
        n, d = (int(v) for v in obs.Z[f"fw{t}_nd"])
        mps = OracleMPS.from_wavefunction(obs.Z[f"fw{t}_input"], n, d, dtype=np.complex128)
        ref0 = obs.Z[f"fw{t}_node0"]
        np.testing.assert_allclose(np.abs(mps.sites[0][0]), np.abs(ref0), atol=1e-10)


def test_from_wavefunction_invalid_args():                     # core_test.py:283-296
    with pytest.raises(TypeError):
        OracleMPS.from_wavefunction({1, 2, 3, 4}, 2, 2)
    with pytest.raises(ValueError):
        OracleMPS.from_wavefunction([1., 0., 0., 0.], 3, 2)
    with pytest.raises(ValueError):
        OracleMPS.from_wavefunction([1., 0.], 1, 2)
    with pytest.raises(ValueError):
        OracleMPS.from_wavefunction(np.array([[1., 0.], [0., 1.]]), 2, 2)


def test_expectation_two_qubit_known_answers():                # core_test.py:1545-1563
    mps = OracleMPS(2)
    assert np.isclose(mps.expectation(HGATE, (0,)), 1 / np.sqrt(2))
    assert np.isclose(mps.expectation(XGATE, (0,)), 0.0)
    mps.x(0)
    assert np.isclose(mps.expectation(HGATE, (0,)), -1 / np.sqrt(2))
    with pytest.raises(ValueError):
        mps.expectation(np.array([[0, 1], [0, 0]]), (0,))


def test_rdm_invalid_indices_and_sampling_known_answers():     # core_test.py:1621-1632, 1714-1737
    mps = OracleMPS(2)
    with pytest.raises(IndexError):
        mps.reduced_density_matrix(-1)
    with pytest.raises(IndexError):
        mps.reduced_density_matrix(22)
    with pytest.raises(ValueError):
        mps.reduced_density_matrix([0, 0])
    for d in (2, 3, 5):
        samples = OracleMPS(2, d).sample(nsamples=20)
        assert len(samples) == 20 and all(set(s) == {0} for s in samples)
    np.random.seed(1)
    mps = OracleMPS(3)
    mps.h(-1)
    hist = mps.sample(nsamples=100, as_hist=True, as_string=True)
    assert all(abs(v / 100 - 1 / 8) < 0.1 for v in hist.values())
    with pytest.raises(ValueError):
        mps.sample(nsamples=0)
    with pytest.raises(ValueError):
        mps.sample(nsamples=1.5)
