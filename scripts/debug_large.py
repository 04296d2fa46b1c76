import numpy as np, sys
sys.path.insert(0, ".")
from tests.test_gpu_kernels import _svd, _graded
for (m, n) in [(256, 256), (200, 136), (512, 512)]:
    for dec in (0.0, 12.0):
        rng = np.random.RandomState(1)
        M = _graded(rng, m, n, dec)
        k = min(m, n)
        for lc in (1,):
            left, right, sv, info = _svd(M[None], k, lc)
            sref = np.linalg.svd(M.astype(np.complex128), compute_uv=False)
            iso = left[0] if lc else right[0].conj().T
            Y = right[0] if lc else left[0].T
            G = Y.astype(np.complex128) @ Y.astype(np.complex128).conj().T
            d = np.sqrt(np.abs(np.diag(G))); off = np.abs(G - np.diag(np.diag(G)))
            ab = (off / (d.max() * np.maximum.outer(d, d) + 1e-300)).max()
            print((m, n, dec, lc), "info", info[0], "sv err/smax %.2e" % (np.abs(sv[0] - sref).max() / sref[0]),
                  "recon %.2e" % (np.abs(left[0] @ right[0] - M).max() / sref[0]),
                  "iso %.2e" % np.abs(iso.conj().T @ iso - np.eye(k)).max(), "abs-offdiag %.2e" % ab)
