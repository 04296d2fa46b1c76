"""CPU model (numpy, complex64) of the CUDA small-SVD kernel (mpsim_b200/csrc/svd_small.cu):
Householder QR preconditioning (R only) -> one-sided Jacobi on the rows of R with the block
tournament schedule -> stable descending sort -> W = X V_k -> isometry Q from a Householder QR of
W -> weighted factor Q^H X.

It exists to validate schedule, formulas, tie-breaking and extraction on the CPU (tests/
test_jacobi_model.py) before anything runs on a GPU; it is NOT used by the product.

Invariant: Q has orthonormal columns by construction, so Q (Q^H X) is an exact orthogonal
projection of X, converged or not.
"""
import numpy as np

F = np.float32
C = np.complex64


def block_rounds(nb):
    """Circle-method tournament on nb (even) blocks: list of rounds, each a list of (I, J)."""
    assert nb % 2 == 0
    if nb == 2:
        return [[(0, 1)]]
    m = nb - 1
    rounds = []
    for r in range(m):
        pairs = [(m, r)]
        for i in range(1, nb // 2):
            pairs.append(((r + i) % m, (r - i) % m))
        rounds.append(pairs)
    return rounds


INTRA = [[(0, 1), (2, 3), (4, 5), (6, 7)], [(0, 2), (1, 3), (4, 6), (5, 7)], [(0, 3), (1, 2), (4, 7), (5, 6)]]
CROSS = [[(i, 4 + (i + s) % 4) for i in range(4)] for s in range(4)]


def householder_qr(Y, Z, m):
    """In place: Y[:m] <- R (upper triangular), Z <- H_k ... H_1 Z.  Reflectors are skipped
    when the column is already zero below the diagonal (keeps diagonal inputs untouched) or so
    small that 1/|x|^2 would overflow; a diagonal entry with |x0|^2 < 1e-30 is treated as 0
    (its square is a denormal with few significant bits -- the phase x0/|x0| computed from it
    made one reflector of a GHZ circuit non-unitary by 2e-4 on the GPU)."""
    L = Y.shape[1]
    for j in range(min(m - 1, L)):
        x = Y[j:m, j].copy()
        tail2 = F((np.abs(x[1:]) ** 2).sum())
        x0 = x[0]
        ax0sq = F(F(x0.real) * F(x0.real) + F(x0.imag) * F(x0.imag))
        if ax0sq < F(1e-30):
            x0, ax0sq = C(0), F(0)
        if not (tail2 > 0 and F(tail2 + ax0sq) > F(1e-30)):
            continue
        ax0 = F(np.sqrt(ax0sq))
        normx = F(np.sqrt(F(tail2 + ax0sq)))
        phase = C(x0 * F(F(1) / ax0)) if ax0 > 0 else C(1)
        alpha = C(-phase * normx)
        v = x.copy()
        v[0] = x0 - alpha
        tau = F(1.0) / F(normx * F(normx + ax0))      # 2 / ||v||^2
        for A in ((Y,) if Z is None else (Y, Z)):
            w = np.conj(v) @ A[j:m, :]
            A[j:m, :] -= np.outer(tau * v, w).astype(C)
        Y[j, j] = alpha
        Y[j + 1:m, j] = 0


def _fma(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def rotation_params(a, b, g):
    """(c, s) of the 2x2 unitary [[c, s], [-conj(s), c]] that diagonalises the Gram matrix
    [[a, g], [conj(g), b]].  s is computed first; c is then DERIVED from s as
    sqrt(1 - |s|^2) so that the applied matrix is unitary to rounding WITHOUT BIAS (a direct
    c = 1/sqrt(1+t^2) has E[c^2+|s|^2-1] ~ +8e-9 per rotation, which drifts every row norm --
    i.e. every singular value -- by ~1e-5 over the ~800 rotations a row sees in a 128x128
    solve; random rounding of the same size only gives ~1e-6)."""
    g2 = F(g.real * g.real + g.imag * g.imag)
    rg = F(1) / F(np.sqrt(g2))
    zeta = F(F(a - b) * F(F(0.5) * rg))
    t = F(np.copysign(F(1), zeta)) / F(abs(zeta) + F(np.sqrt(_fma(zeta, zeta, F(1)))))
    ct = F(F(t / F(np.sqrt(_fma(t, t, F(1))))) * rg)
    sr = F(ct * g.real)
    si = F(ct * g.imag)
    h = _fma(sr, sr, F(si * si))
    if h < F(0.0625):
        # sqrt(1-h) by its series with ONE final rounding: sqrt() of the pre-rounded 1-h is
        # biased low by ~1.5e-8 near 1 (odd grid points are exact ties that always round down)
        poly = _fma(h, _fma(h, _fma(h, _fma(h, F(0.02734375), F(0.0390625)), F(0.0625)), F(0.125)), F(0.5))
        c = _fma(-h, poly, F(1))
    else:
        c = F(np.sqrt(_fma(-sr, sr, _fma(-si, si, F(1)))))
    return c, C(complex(sr, si))


def _r(x):
    return x.astype(F).astype(np.float64)


def apply_rotation(c, s, yp, yq):
    """[yp'; yq'] = [[c, s], [-conj(s), c]] [yp; yq] with the FMA chains of the CUDA code:
    re = fma(c, p.re, fma(s.re, q.re, -(s.im * q.im))) etc. (each fma rounds once)."""
    c = np.float64(c)
    sr, si = np.float64(s.real), np.float64(s.imag)
    pr, pi = yp.real.astype(np.float64), yp.imag.astype(np.float64)
    qr, qi = yq.real.astype(np.float64), yq.imag.astype(np.float64)
    npr = _r(c * pr + _r(sr * qr - _r(si * qi)))
    npi = _r(c * pi + _r(sr * qi + _r(si * qr)))
    # -conj(s) * p = (-sr p.re - si p.im) + i (si p.re - sr p.im)
    nqr = _r(c * qr - _r(sr * pr + _r(si * pi)))
    nqi = _r(c * qi + _r(si * pr - _r(sr * pi)))
    return (npr + 1j * npi).astype(C), (nqr + 1j * nqi).astype(C)


def rotate(Y, Z, p, q, tol2):
    yp, yq = Y[p], Y[q]
    a = F(np.vdot(yp, yp).real)
    b = F(np.vdot(yq, yq).real)
    g = C(np.vdot(yq, yp))                     # sum yp * conj(yq)
    g2 = F(g.real * g.real + g.imag * g.imag)
    if not (g2 > tol2 * a * b) or g2 == 0:
        return 0
    c, s = rotation_params(a, b, g)
    Y[p], Y[q] = apply_rotation(c, s, yp, yq)
    if Z is not None:
        Z[p], Z[q] = apply_rotation(c, s, Z[p].copy(), Z[q].copy())
    return 1


def orthogonalize_rows(M, max_sweeps=30, tol=3e-6, qr=True):
    """Returns Y, sweeps, rotations: rows of Y orthogonal, Y = (some unitary) @ M."""
    nv, L = M.shape
    nvp = max(8, (nv + 7) // 8 * 8)
    Y = np.zeros((nvp, L), C)
    Y[:nv] = M
    if qr:
        householder_qr(Y, None, nv)
    nact = min(nv, L) if qr else nv            # rows >= L of R are exactly zero
    nb = max(2, (nact + 7) // 8 * 2)
    rounds = block_rounds(nb)
    tol2 = F(tol * tol)
    total = 0
    for sweep in range(max_sweeps):
        nrot = 0
        for r, pairs in enumerate(rounds):
            for (I, J) in pairs:
                rows = [4 * I + i for i in range(4)] + [4 * J + i for i in range(4)]
                for sub in ((INTRA if r == 0 else []) + CROSS):
                    for (x, y) in sub:
                        nrot += rotate(Y, None, rows[x], rows[y], tol2)
        total += nrot
        if nrot == 0:
            break
    return Y, sweep + 1, total


def householder_q(W):
    """First k columns of the unitary of a Householder QR of W (nv x k): H_0 ... H_{k-1} [I_k; 0],
    with the kernel's skip rules (no reflector for an already reduced or negligible column)."""
    nv, k = W.shape
    A = W.astype(C).copy()
    refl = []
    for j in range(k):
        x = A[j:, j].copy()
        tail2 = F((np.abs(x[1:]) ** 2).sum())
        x0 = x[0]
        ax0sq = F(F(x0.real) * F(x0.real) + F(x0.imag) * F(x0.imag))
        if ax0sq < F(1e-30):
            x0, ax0sq = C(0), F(0)
        if not (tail2 > 0 and F(tail2 + ax0sq) > F(1e-30)):
            refl.append(None)
            continue
        ax0 = F(np.sqrt(ax0sq))
        normx = F(np.sqrt(F(tail2 + ax0sq)))
        phase = C(x0 * F(F(1) / ax0)) if ax0 > 0 else C(1)
        alpha = C(-phase * normx)
        v = x.copy()
        v[0] = x0 - alpha
        tau = F(1.0) / F(normx * F(normx + ax0))
        w = np.conj(v) @ A[j:, j + 1:]
        A[j:, j + 1:] -= np.outer(tau * v, w).astype(C)
        refl.append((v, tau))
    Q = np.zeros((nv, k), C)
    Q[:k, :k] = np.eye(k, dtype=C)
    for j in range(k - 1, -1, -1):
        if refl[j] is None:
            continue
        v, tau = refl[j]
        w = np.conj(v) @ Q[j:, j:]
        Q[j:, j:] -= np.outer(tau * v, w).astype(C)
    return Q


def split(M, k, left_canonical=True, **kw):
    """The kernel's contract.  Returns (left (m x k), right (k x n), sigma_sorted, sweeps).
    left_canonical:      left = U,    right = S Vh   (vectors = rows of M)
    not left_canonical:  left = U S,  right = Vh     (vectors = columns of M)."""
    m, n = M.shape
    X = (M if left_canonical else np.ascontiguousarray(M.T)).astype(C)
    nv = X.shape[0]
    mx = float(np.abs(np.concatenate([X.real.ravel(), X.imag.ravel()])).max()) if X.size else 0.0
    scale = F(2.0 ** (1 - np.frexp(mx)[1])) if mx > 0 else F(1)       # max|x| -> [1, 2)
    Xs = (X * scale).astype(C)
    Y, sweeps, _ = orthogonalize_rows(Xs, **kw)
    norms = np.sqrt((np.abs(Y) ** 2).sum(axis=1, dtype=F)).astype(F)
    norms[nv:] = -1                           # padding rows sort last
    perm = np.argsort(-norms, kind="stable")
    sk = norms[perm[:k]]
    rinv = np.where(sk > F(1e-12), F(1) / np.maximum(sk, F(1e-30)), F(0)).astype(F)
    W = ((Xs @ np.conj(Y[perm[:k]]).T).astype(C) * rinv * rinv).astype(C)      # nv x k, columns ~ u_j
    Q = householder_q(W)
    wgt = (np.conj(Q).T @ X).astype(C)        # k x L
    sig = (norms[perm][:min(m, n)] / scale).astype(F)
    if left_canonical:
        return Q, wgt, sig, sweeps                  # U = Q ; S Vh = Q^H X
    return wgt.T.copy(), Q.T.copy(), sig, sweeps    # U S = (Q^H X)^T ; Vh = Q^T
