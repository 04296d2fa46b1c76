"""GPU parity for SURVEY.md 8(f) rows 2-3 through the reference-shaped API (which calls the
C-ABI): from_wavefunction, reduced_density_matrix, sample, expectation, inner_product.

Checked against (1) vectors from the UNMODIFIED reference (tests/golden/observables.npz) with
the same assertions the CPU oracle passes (tests/_observables.py), (2) the reference's own tests,
ported (file:line cited), (3) dense-vector arithmetic on seeded random states.
Tolerance: 1e-5 absolute on amplitudes / matrix entries of complex64 states (north_star: 1e-4)."""
import numpy as np
import pytest

from tests import _observables as obs

pytestmark = pytest.mark.gpu
ATOL = 1e-5


def _mp():
    import mpsim_b200
    return mpsim_b200


class DeviceAdapter:
    @staticmethod
    def new(n, d):
        return _mp().MPS(n, d)

    @staticmethod
    def from_wavefunction(wf, n, d):
        return _mp().MPS.from_wavefunction(wf, nqudits=n, qudit_dimension=d)

    @staticmethod
    def apply1(mps, tensor, i):
        mps.apply_one_qudit_gate(_mp().Node(np.array(tensor, copy=True)), i)

    @staticmethod
    def apply2(mps, tensor, i, j, **kw):
        mps.apply_two_qudit_gate(_mp().Node(np.array(tensor, copy=True)), i, j, **kw)

    @staticmethod
    def expectation(mps, tensor, indices):
        return mps.expectation(_mp().MPSOperation(_mp().Node(np.array(tensor, copy=True)), indices))


@pytest.mark.parametrize("name", obs.STATES)
def test_observables_match_reference(name):
    obs.check_state(DeviceAdapter, name, atol=ATOL)


def test_inner_products_match_reference():
    obs.check_inner_products(DeviceAdapter, atol=ATOL)


def test_from_wavefunction_matches_reference():
    obs.check_from_wavefunction(DeviceAdapter, atol=ATOL)


def density_matrix_from_state_vector(state, indices, d=2):
    """Dense ground truth (stand-in for cirq.density_matrix_from_state_vector, core_test.py:13)."""
    n = int(round(np.log(state.size) / np.log(d)))
    psi = state.reshape([d] * n)
    rest = [i for i in range(n) if i not in indices]
    m = np.transpose(psi, list(indices) + rest).reshape(d ** len(indices), -1)
    return m @ m.conj().T


def test_from_wavefunction_known_answers_and_errors():          # core_test.py:230-296
    MPS = _mp().MPS
    for n in (2, 3):
        wf = np.zeros(2 ** n)
        wf[0] = 1
        mps = MPS.from_wavefunction(wf, nqudits=n, qudit_dimension=2)
        assert isinstance(mps, MPS) and mps.nqudits == n and mps.qudit_dimension == 2
        assert np.allclose(mps.wavefunction(), wf, atol=ATOL)
        assert mps.is_valid()
        assert np.isclose(mps.norm(), 1.0, atol=ATOL)
    np.random.seed(1)
    for n in range(2, 8):
        for _ in range(3):
            wf = np.random.rand(2 ** n)
            wf /= np.linalg.norm(wf, ord=2)
            assert np.allclose(MPS.from_wavefunction(wf, nqudits=n).wavefunction(), wf, atol=ATOL)
    np.random.seed(11)
    for n in range(2, 5):
        for d in (2, 3, 4):
            wf = np.random.rand(d ** n)
            wf /= np.linalg.norm(wf, ord=2)
            assert np.allclose(MPS.from_wavefunction(wf, nqudits=n, qudit_dimension=d).wavefunction(), wf, atol=ATOL)
    with pytest.raises(TypeError):
        MPS.from_wavefunction({1, 2, 3, 4}, nqudits=2, qudit_dimension=2)
    with pytest.raises(ValueError):
        MPS.from_wavefunction([1., 0., 0., 0.], nqudits=3, qudit_dimension=2)
    with pytest.raises(ValueError):
        MPS.from_wavefunction([1., 0.], nqudits=1, qudit_dimension=2)
    with pytest.raises(ValueError):
        MPS.from_wavefunction(np.array([[1., 0.], [0., 1.]]), nqudits=2, qudit_dimension=2)


def test_from_wavefunction_ten_qubits_uses_the_large_svd_path():   # cuts of 2 x 512 ... 32 x 32
    rng = np.random.RandomState(5)
    wf = rng.randn(1024) + 1j * rng.randn(1024)
    wf /= np.linalg.norm(wf)
    mps = _mp().MPS.from_wavefunction(wf, nqudits=10)
    assert mps.bond_dimensions() == [2, 4, 8, 16, 32, 16, 8, 4, 2]
    assert np.abs(mps.wavefunction() - wf).max() < ATOL
    rdm = mps.reduced_density_matrix([7, 2])
    assert np.abs(rdm - density_matrix_from_state_vector(wf, [7, 2])).max() < ATOL


def test_dagger_random_wavefunctions():                          # core_test.py:1566-1585
    np.random.seed(10)
    for n in (2, 3, 5, 10):
        wf = np.random.randn(2 ** n) + np.random.randn(2 ** n) * 1j
        wf /= np.linalg.norm(wf, ord=2)
        mps = _mp().MPS.from_wavefunction(wf, nqudits=n)
        assert np.allclose(mps.wavefunction(), wf, atol=ATOL)
        mps.dagger()
        assert np.allclose(mps.wavefunction(), wf.conj(), atol=ATOL)


def test_expectation_two_qubit_mps():                            # core_test.py:1545-1563
    mp = _mp()
    mps = mp.MPS(nqudits=2)
    mps_copy = mps.copy()
    h0 = mp.MPSOperation(mp.hgate(), 0)
    assert np.isclose(mps.expectation(h0), 1. / np.sqrt(2), atol=ATOL)
    assert mps == mps_copy
    x0 = mp.MPSOperation(mp.xgate(), 0)
    assert np.isclose(mps.expectation(x0), 0., atol=ATOL)
    assert mps == mps_copy
    mps.apply(mp.MPSOperation(mp.xgate(), 0))
    assert np.isclose(mps.expectation(h0), -1. / np.sqrt(2), atol=ATOL)
    with pytest.raises(ValueError):                              # not Hermitian (core.py:736-737)
        mps.expectation(mp.MPSOperation(mp.Node(np.array([[0., 1.], [0., 0.]])), 0))
    with pytest.raises(ValueError):                              # dimension mismatch (core.py:739-746)
        mps.expectation(mp.MPSOperation(mp.Node(np.eye(3)), 0, qudit_dimension=3))


def test_reduced_density_matrix_simple_and_invalid():            # core_test.py:1588-1632
    mp = _mp()
    mps = mp.MPS(nqudits=2, qudit_dimension=2)
    for i in (0, 1):
        assert np.allclose(mps.reduced_density_matrix(node_indices=i), [[1., 0.], [0., 0.]])
        assert mps == mp.MPS(nqudits=2, qudit_dimension=2)
    mps.apply(mp.MPSOperation(mp.xgate(), 0))
    assert np.allclose(mps.reduced_density_matrix(node_indices=0), [[0., 0.], [0., 1.]])
    assert np.allclose(mps.reduced_density_matrix(node_indices=1), [[1., 0.], [0., 0.]])
    mps.apply(mp.MPSOperation(mp.xgate(), 1))
    assert np.allclose(mps.reduced_density_matrix(node_indices=0), [[0., 0.], [0., 1.]])
    assert np.allclose(mps.reduced_density_matrix(node_indices=1), [[0., 0.], [0., 1.]])
    mps = mp.MPS(nqudits=2)
    with pytest.raises(IndexError):
        mps.reduced_density_matrix(node_indices=-1)
    with pytest.raises(IndexError):
        mps.reduced_density_matrix(node_indices=22)
    with pytest.raises(ValueError):
        mps.reduced_density_matrix(node_indices=[0, 0])


def test_density_matrices_random_states():                       # core_test.py:1635-1711
    np.random.seed(5)
    for _ in range(5):
        wf = np.random.randn(8) + np.random.randn(8) * 1j
        wf /= np.linalg.norm(wf)
        mps = _mp().MPS.from_wavefunction(wf, nqudits=3)
        for i in [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)]:
            for idx in (i, tuple(reversed(i))):
                rdm = mps.reduced_density_matrix(node_indices=idx)
                assert np.allclose(rdm, density_matrix_from_state_vector(wf, idx), atol=ATOL)
        assert np.allclose(mps.wavefunction(), wf, atol=ATOL)
    np.random.seed(1)
    for n in (3, 5, 8):
        for _ in range(4):
            wf = np.random.randn(2 ** n) + np.random.randn(2 ** n) * 1j
            wf /= np.linalg.norm(wf)
            mps = _mp().MPS.from_wavefunction(wf, nqudits=n)
            size = np.random.randint(low=1, high=n)
            sites = [int(q) for q in np.random.choice(range(n), size=size, replace=False)]
            rdm = mps.reduced_density_matrix(node_indices=sites)
            assert np.allclose(rdm, density_matrix_from_state_vector(wf, sites), atol=ATOL)


def test_qutrit_density_matrix_and_expectation():
    rng = np.random.RandomState(9)
    n, d = 4, 3
    wf = rng.randn(d ** n) + 1j * rng.randn(d ** n)
    wf /= np.linalg.norm(wf)
    mp = _mp()
    mps = mp.MPS.from_wavefunction(wf, nqudits=n, qudit_dimension=d)
    rdm = mps.reduced_density_matrix([2, 0])
    assert np.abs(rdm - density_matrix_from_state_vector(wf, [2, 0], d)).max() < ATOL
    h = rng.randn(d, d) + 1j * rng.randn(d, d)
    u, _ = np.linalg.qr(h)
    obs_m = u @ np.diag([1., -1., 1.]) @ u.conj().T              # Hermitian and unitary
    val = mps.expectation(mp.MPSOperation(mp.Node(obs_m), 1, qudit_dimension=d))
    ref = np.trace(density_matrix_from_state_vector(wf, [1], d) @ obs_m).real
    assert abs(val - ref) < 4 * ATOL


def test_sample_zero_state_and_uniform():                        # core_test.py:1714-1737
    mp = _mp()
    for d in (2, 3, 5):
        samples = mp.MPS(nqudits=2, qudit_dimension=d).sample(nsamples=100)
        assert len(samples) == 100
        for sample in samples:
            assert set(sample) == {0}
    np.random.seed(1)
    n, nsamples = 3, 100
    mps = mp.MPS(nqudits=n)
    mps.apply([mp.MPSOperation(mp.hgate(), i) for i in range(n)])
    hist = mps.sample(nsamples=nsamples, as_hist=True, as_string=True)
    for freq in np.array(list(hist.values())) / nsamples:
        assert np.abs(freq - 1. / 2 ** n) < 1. / np.sqrt(nsamples)
    assert all(isinstance(k, str) and len(k) == n for k in hist)
    with pytest.raises(ValueError):
        mps.sample(nsamples=0)
    with pytest.raises(ValueError):
        mps.sample(nsamples=2.0)


def test_marginals_of_a_chi64_chain_sum_to_one():
    """Size-independent property at the bench workload's shape: every single-site marginal of a
    40-qubit chi=64 brickwork state sums to the squared norm."""
    mp = _mp()
    from mpsim_b200 import circuits, observables
    n = 40
    ops = circuits.brickwork(n, 12, seed=4)
    mps = mp.MPS(n)
    mps._execute([(op.tensor, op.indices, {"maxsvals": 64, "keep_left_canonical": op.keep_left_canonical})
                  for op in ops])
    nrm2 = mps.norm() ** 2
    marg = observables.site_marginals(mps)
    assert marg.shape == (n, 2)
    assert np.abs(marg.sum(axis=1) - nrm2).max() < 1e-4 * nrm2
    assert (marg > -1e-6).all()
