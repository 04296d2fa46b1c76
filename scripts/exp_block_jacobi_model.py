"""numpy model of the block one-sided Jacobi of svd_large.cu (16-row blocks, circle-method rounds, ONE cyclic
pass on the 32 x 32 pair Gram, the two rotation criteria) used to count sweeps under preconditioners:
    python scripts/exp_block_jacobi_model.py thetas.npy plain rowsort qr qrsort
thetas.npy: complex64 [n][512][512], e.g. captured from the oracle on the chi = 256 workload by hooking
oracle.mps_oracle._svd_trunc.  Measured (round 2): plain 12-16 sweeps, rowsort the same, qr 9-11, qrsort 8-10.
"""
import numpy as np, sys, time
C = np.complex64; F = np.float32
mats = np.load(sys.argv.pop(1)) if len(sys.argv) > 1 and sys.argv[1].endswith(".npy") else None
B = 16

def circle(nb):
    m = nb - 1
    R = []
    for r in range(m):
        p = [(m, r)] + [((r + g) % m, (r - g + m) % m) for g in range(1, nb // 2)]
        R.append(p)
    return R

def intra_sets():
    # circle method inside each 16-block (15 sets of 8 pairs per block -> 16 rotations per set over both blocks)
    sets = []
    for r, pairs in enumerate(circle(16)):
        s = [(a, b) for a, b in pairs] + [(16 + a, 16 + b) for a, b in pairs]
        sets.append(s)
    return sets
CROSS = [[(i, 16 + (i + s) % 16) for i in range(16)] for s in range(16)]
INTRA = intra_sets()

def evd_pass(G, first, tol2, eta2g, inner=1):
    """G [np,32,32] Hermitian; returns Q [np,32,32] with Q G Q^H more diagonal; count of rotations"""
    npairs = G.shape[0]
    Q = np.tile(np.eye(32, dtype=np.complex128), (npairs, 1, 1))
    G = G.astype(np.complex128).copy()
    nrot = 0
    ar = np.arange(npairs)[:, None]
    for it in range(inner):
        for sset in ((INTRA if first else []) + CROSS):
            p = np.array([a for a, b in sset]); q = np.array([b for a, b in sset])
            a = G[:, p, p].real; b = G[:, q, q].real; g = G[:, p, q]
            g2 = np.abs(g) ** 2
            act = (g2 > tol2 * a * b) & (g2 > eta2g[:, None] * np.maximum(a, b)) & (g2 > 1e-30)
            nrot += int(act.sum())
            if not act.any():
                continue
            d = a - b
            h = np.sqrt(d * d + 4 * g2); w = h + np.abs(d)
            with np.errstate(divide="ignore", invalid="ignore"):
                s2 = np.where(act, 2 * g2 / (h * w), 0.0)
                s = np.where(act, np.where(d >= 0, 1, -1) * np.sqrt(s2) * g / np.sqrt(np.where(g2 > 0, g2, 1)), 0)
            c = np.sqrt(1 - s2)
            # J = [[c, s], [-conj(s), c]] on rows (p,q): rows' = J rows ; G <- J G J^H
            for M in (G, Q):
                rp = M[:, p, :].copy(); rq = M[:, q, :].copy()
                M[:, p, :] = c[:, :, None] * rp + s[:, :, None] * rq
                M[:, q, :] = c[:, :, None] * rq - np.conj(s)[:, :, None] * rp
            cp_ = G[:, :, p].copy(); cq = G[:, :, q].copy()
            G[:, :, p] = c[:, None, :] * cp_ + np.conj(s)[:, None, :] * cq
            G[:, :, q] = c[:, None, :] * cq - s[:, None, :] * cp_
    return Q, nrot

def block_jacobi(X, max_sweeps=40, inner=1, eta=3e-7, verbose=False):
    X = X.astype(C).copy()
    nv, L = X.shape
    nb = nv // B
    rounds = circle(nb)
    tol = max(3e-6, 4 * 5.96e-8 * np.sqrt(L)); tol2 = tol * tol
    gmax = float((np.abs(X) ** 2).sum(axis=1).max())
    for sweep in range(max_sweeps):
        tot = 0
        for r, pairs in enumerate(rounds):
            idx = np.array([[I * B + t for t in range(B)] + [J * B + t for t in range(B)] for I, J in pairs])
            Xp = X[idx]                                  # [np,32,L]
            G = np.einsum("pil,pjl->pij", Xp, np.conj(Xp)).astype(C)
            Q, nrot = evd_pass(G, r == 0, tol2, np.full(len(pairs), eta * eta * gmax), inner)
            tot += nrot
            X[idx] = np.einsum("pij,pjl->pil", Q.astype(C), Xp).astype(C)
        gmax = float((np.abs(X) ** 2).sum(axis=1).max())
        if verbose: print("  sweep", sweep + 1, "rotations", tot)
        if tot == 0:
            return X, sweep + 1
    return X, max_sweeps

if __name__ == "__main__":
    which = sys.argv[1:] or ["plain", "rowsort"]
    for mi in range(0, len(mats), 2):
        X0 = mats[mi]
        sv = np.linalg.svd(X0.astype(np.complex128), compute_uv=False)
        for name in which:
            X = X0
            if name == "rowsort":
                X = X0[np.argsort(-np.linalg.norm(X0, axis=1), kind="stable")]
            if name == "qr":
                X = np.linalg.qr(X0.astype(np.complex128))[1].astype(C)
            if name == "qrsort":
                pi = np.argsort(-np.linalg.norm(X0, axis=0), kind="stable")
                X = np.linalg.qr(X0[:, pi].astype(np.complex128))[1].astype(C)
            if name == "transpose":
                X = X0.T.copy()
            t0 = time.time()
            Y, sw = block_jacobi(X, inner=(3 if name == "inner3" else 1))
            s = np.sort(np.linalg.norm(Y.astype(np.complex128), axis=1))[::-1]
            print("matrix %d %-10s sweeps %2d  sigma err %.1e  (%.0fs)  cond %.1e" % (mi, name, sw, np.abs(s - sv).max() / sv.max(), time.time() - t0, sv[0] / max(sv[-1], 1e-300)), flush=True)
