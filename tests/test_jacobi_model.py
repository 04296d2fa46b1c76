"""Validates the algorithm the CUDA small-SVD kernel implements (schedule, Householder
preconditioning, rotation formulas, stable sort / tie-breaking, extraction) on the CPU model."""
import numpy as np
import pytest

from tests import _jacobi_model as jm


def _graded(rng, m, n, decay):
    a = rng.randn(m, n) + 1j * rng.randn(m, n)
    u, s, vh = np.linalg.svd(a, full_matrices=False)
    s = s * np.exp(-np.arange(len(s)) / len(s) * decay)
    return ((u * s) @ vh).astype(np.complex64)


def test_schedule_covers_every_pair_once():
    for nb in (2, 4, 8, 16, 32):
        seen = set()
        for r, pairs in enumerate(jm.block_rounds(nb)):
            blocks = [b for p in pairs for b in p]
            assert sorted(blocks) == list(range(nb))          # disjoint, complete
            for (I, J) in pairs:
                rows = [4 * I + i for i in range(4)] + [4 * J + i for i in range(4)]
                for sub in ((jm.INTRA if r == 0 else []) + jm.CROSS):
                    assert len({x for p in sub for x in p}) == 8   # disjoint within a sub-round
                    for (x, y) in sub:
                        key = (min(rows[x], rows[y]), max(rows[x], rows[y]))
                        assert key not in seen
                        seen.add(key)
        n = 4 * nb
        assert len(seen) == n * (n - 1) // 2


@pytest.mark.parametrize("shape", [(2, 2), (4, 4), (8, 16), (16, 8), (32, 32), (64, 32), (48, 64)])
@pytest.mark.parametrize("decay", [0.0, 12.0])
@pytest.mark.parametrize("left", [True, False])
def test_split_matches_lapack(shape, decay, left):
    rng = np.random.RandomState(hash((shape, decay, left)) % 2 ** 31)
    m, n = shape
    M = _graded(rng, m, n, decay)
    sref = np.linalg.svd(M.astype(np.complex128), compute_uv=False)
    k = min(m, n)
    lft, rgt, sig, sweeps = jm.split(M, k, left)
    assert sweeps <= 12
    assert np.abs(sig - sref).max() <= 1e-5 * sref[0]
    np.testing.assert_allclose(lft @ rgt, M, atol=2e-5 * sref[0])
    iso = lft if left else rgt.conj().T
    np.testing.assert_allclose(iso.conj().T @ iso, np.eye(k), atol=5e-5)
    # truncation: best rank-k' approximation
    kk = max(1, k // 2)
    lft, rgt, sig, _ = jm.split(M, kk, left)
    u, s, vh = np.linalg.svd(M.astype(np.complex128), full_matrices=False)
    best = (u[:, :kk] * s[:kk]) @ vh[:kk]
    np.testing.assert_allclose(lft @ rgt, best, atol=5e-5 * sref[0] + 20 * s[kk] * 1e-3)


def test_bell_tie_keeps_first():                 # README.md:48-53, core_test.py:915-922
    M = np.diag([2 ** -0.5, 2 ** -0.5]).astype(np.complex64)
    for left in (True, False):
        lft, rgt, sig, _ = jm.split(M, 1, left)
        np.testing.assert_allclose(lft @ rgt, np.diag([2 ** -0.5, 0]), atol=1e-7)


def test_rank_deficient_keeps_orthonormal_isometry():      # core_test.py:932-944
    M = np.zeros((4, 4), np.complex64)
    M[0, 0] = 1
    M[1, 2] = 0.5
    for left in (True, False):
        lft, rgt, sig, _ = jm.split(M, 4, left)
        np.testing.assert_allclose(sig, [1, 0.5, 0, 0], atol=1e-7)
        iso = lft if left else rgt.conj().T
        np.testing.assert_allclose(iso.conj().T @ iso, np.eye(4), atol=1e-6)
        np.testing.assert_allclose(lft @ rgt, M, atol=1e-7)


def test_denormal_diagonal_entry_regression():
    """theta of a GHZ circuit (16x8, four singular values ~1, four ~1e-8) on which a
    denormal |x0|^2 made one Householder reflector non-unitary by 2e-4."""
    import os
    M = np.load(os.path.join(os.path.dirname(__file__), "golden", "ghz_theta_16x8.npy"))
    for left in (True, False):
        lft, rgt, sig, _ = jm.split(M, 8, left)
        np.testing.assert_allclose(lft @ rgt, M, atol=1e-6)
        iso = lft if left else rgt.conj().T
        np.testing.assert_allclose(iso.conj().T @ iso, np.eye(8), atol=2e-6)
