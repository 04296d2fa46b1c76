"""Experiment (build container only): Jacobi sweep counts under different preconditioners,
on theta matrices captured from a chi=64 brickwork circuit run through the oracle."""
import sys, numpy as np
sys.path.insert(0, '.')
import oracle.mps_oracle as mo
import scipy.linalg as sl

caps = []
orig = mo._svd_trunc
def cap(mat, k, e):
    if mat.shape == (128, 128) and len(caps) < 400: caps.append(mat.copy())
    return orig(mat, k, e)
mo._svd_trunc = cap
rs = np.random.RandomState(7)
n, depth = 16, 24
m = mo.OracleMPS(n)
for layer in range(depth):
    for i in range(layer % 2, n - 1, 2):
        g = mo.haar_random_unitary(2, 2, rs).reshape(2, 2, 2, 2) if 'rs' in mo.haar_random_unitary.__code__.co_varnames else mo.haar_random_unitary(2,2).reshape(2,2,2,2)
        m.apply_two_qudit_gate(g, i, i + 1, maxsvals=64)
print(len(caps), "thetas")

def rounds(nb):
    mm = nb - 1
    out = []
    for r in range(mm):
        pr = [(mm, r)] + [((r + i) % mm, (r - i) % mm) for i in range(1, nb // 2)]
        out.append(pr)
    return out

def jacobi_sweeps(Y, tol=3e-6, maxs=30, order='tournament'):
    Y = Y.astype(np.complex128).copy()
    nv = Y.shape[0]
    R = rounds(nv)
    for sweep in range(maxs):
        nrot = 0
        for pr in R:
            for (p, q) in pr:
                if p > q: p, q = q, p
                yp, yq = Y[p], Y[q]
                a = np.vdot(yp, yp).real; b = np.vdot(yq, yq).real
                g = np.vdot(yq, yp)
                g2 = abs(g) ** 2
                if not g2 > tol * tol * a * b or g2 == 0: continue
                zeta = (a - b) / (2 * abs(g))
                t = np.sign(zeta) / (abs(zeta) + np.sqrt(1 + zeta * zeta)) if zeta != 0 else 1.0
                c = 1 / np.sqrt(1 + t * t); s = c * t * g / abs(g)
                Y[p], Y[q] = c * yp + s * yq, -np.conj(s) * yp + c * yq
                nrot += 1
        if nrot == 0: return sweep + 1
    return maxs

res = {}
for idx in list(range(0, len(caps), max(1, len(caps) // 12)))[:12]:
    M = caps[idx]
    q, r = np.linalg.qr(M)
    q2, r2, piv = sl.qr(M, pivoting=True)
    # rows sorted by norm (descending) before QR == column pivot only at start
    nrm = np.linalg.norm(M, axis=0); o = np.argsort(-nrm)
    q3, r3 = np.linalg.qr(M[:, o])
    # second QR: R^H = Q' R' ; Jacobi on rows of R'^... (use R'^H rows? compare both)
    q4, r4 = np.linalg.qr(r2.conj().T)
    for name, Y in [('plain', M), ('qr', r), ('qr_piv', r2), ('qr_sort', r3), ('qr_piv_lq_rows', r4), ('qr_piv_lq_cols', r4.conj().T), ('qr_lq', np.linalg.qr(r.conj().T)[1]), ('qr_lq_T', np.linalg.qr(r.conj().T)[1].conj().T)]:
        res.setdefault(name, []).append(jacobi_sweeps(Y))
    print(idx, {k: v[-1] for k, v in res.items()}, flush=True)
for k, v in res.items(): print(k, np.mean(v))
