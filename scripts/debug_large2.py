import numpy as np, sys
sys.path.insert(0, ".")
from tests.test_gpu_kernels import _svd, _graded
rng = np.random.RandomState(1)
M = _graded(rng, 256, 256, 12.0)
left, right, sv, info = _svd(M[None], 256, 1)
sref = np.linalg.svd(M.astype(np.complex128), compute_uv=False)
err = np.abs(sv[0] - sref) / sref[0]
i = np.argsort(-err)[:8]
print("info", info[0]); print("worst idx", i); print("err", err[i]); print("sv", sv[0][i]); print("ref", sref[i])
Y = right[0].astype(np.complex128)
nr = np.sqrt((np.abs(Y) ** 2).sum(1))
print("row norms vs sv", np.abs(nr - sv[0]).max())
G = Y @ Y.conj().T
d = np.sqrt(np.abs(np.diag(G))); off = np.abs(G - np.diag(np.diag(G)))
cosm = off / (np.outer(d, d) + 1e-300)
j = i[0]
k = np.argsort(-cosm[j])[:5]
print("row", j, "largest cos with rows", k, cosm[j][k], "their sigma", d[k])
