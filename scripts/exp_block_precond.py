"""Experiment (build container, CPU): sweep counts of the block one-sided Jacobi of svd_large.cu
(16-row blocks, circle method, ONE cyclic two-sided pass on the 32 x 32 Gram per pair) with and
without a QR preconditioner, float64 emulation (counts sweeps, not rounding).
python scripts/exp_block_precond.py [n] [kind]     kind: flat | graded | lowrank"""
import sys
import numpy as np
import scipy.linalg as sl

BLK = 16


def rounds(nb):
    m = nb - 1
    return [[(m, r)] + [((r + i) % m, (r - i) % m) for i in range(1, nb // 2)] for r in range(m)]


def intra_sets():
    m = BLK - 1
    out = []
    for t in range(m):
        prs = []
        for blk in range(2):
            for i in range(8):
                a, b = (m, t) if i == 0 else ((t + i) % m, (t - i) % m)
                prs.append((blk * BLK + min(a, b), blk * BLK + max(a, b)))
        out.append(prs)
    return out


CROSS = [[(i, BLK + (i + s) % BLK) for i in range(BLK)] for s in range(BLK)]
INTRA = intra_sets()


def pass_on_gram(G, first, tol2, eta2g):
    """one cyclic pass of two-sided rotations on the Hermitian G; returns Q (32 x 32), nrot"""
    n = G.shape[0]
    Q = np.eye(n, dtype=complex)
    nrot = 0
    for prs in ((INTRA if first else []) + CROSS):
        for (p, q) in prs:
            a, b, g = G[p, p].real, G[q, q].real, G[p, q]
            g2 = abs(g) ** 2
            if not (g2 > tol2 * a * b and g2 > eta2g * max(a, b) and g2 > 1e-300):
                continue
            zeta = (a - b) / (2 * abs(g))
            t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1 + zeta * zeta))
            c = 1 / np.sqrt(1 + t * t)
            s = c * t * g / abs(g)
            J = np.array([[c, s], [-np.conj(s), c]])
            G[[p, q], :] = J @ G[[p, q], :]
            G[:, [p, q]] = G[:, [p, q]] @ J.conj().T
            Q[[p, q], :] = J @ Q[[p, q], :]
            nrot += 1
    return Q, nrot


def block_jacobi_sweeps(X, tol=3e-6, eta=3e-7, max_sweeps=60):
    X = X.astype(complex).copy()
    nv = X.shape[0]
    nb = (nv + BLK - 1) // BLK
    nb += nb % 2
    Xp = np.zeros((nb * BLK, X.shape[1]), complex)
    Xp[:nv] = X
    R = rounds(nb)
    gmax = 0.0
    for sweep in range(max_sweeps):
        tot = 0
        gnext = 0.0
        for r, prs in enumerate(R):
            for (I, J) in prs:
                rows = list(range(I * BLK, (I + 1) * BLK)) + list(range(J * BLK, (J + 1) * BLK))
                Y = Xp[rows]
                G = Y @ Y.conj().T
                gnext = max(gnext, G.diagonal().real.max())
                Q, nrot = pass_on_gram(G, r == 0, tol * tol, eta * eta * gmax)
                if nrot:
                    Xp[rows] = Q @ Y
                    tot += nrot
        gmax = gnext
        if tot == 0:
            return sweep + 1
    return max_sweeps


def make(n, kind, rng):
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    if kind == "flat":
        return a / np.sqrt(2 * n)
    u, s, vh = np.linalg.svd(a)
    if kind == "graded":
        s = np.exp(-np.arange(n) / n * 8.0)
    elif kind == "lowrank":
        s = np.where(np.arange(n) < n // 8, 1.0, 0.0) * np.exp(-np.arange(n) / n * 4.0)
    return (u * s) @ vh


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    kinds = sys.argv[2:] or ["flat", "graded", "lowrank"]
    rng = np.random.default_rng(0)
    for kind in kinds:
        if kind.endswith(".npy"):
            M = np.load(kind)
        else:
            M = make(n, kind, rng)
        M = M.astype(np.complex64).astype(complex)           # fp32 input noise, like the device
        plain = block_jacobi_sweeps(M)
        Rq = np.linalg.qr(M, mode="r")
        withqr = block_jacobi_sweeps(Rq)
        Rp = sl.qr(M, mode="r", pivoting=True)[0]
        withpqr = block_jacobi_sweeps(Rp)
        print(f"{kind} {M.shape}: plain {plain} sweeps, rows of R (QR) {withqr}, rows of R (pivoted QR) {withpqr}", flush=True)
