import numpy as np, sys, os, subprocess
sys.path.insert(0, ".")
M = np.load("scripts/bad_theta.npy")
code = r'''
import numpy as np, sys
sys.path.insert(0, ".")
from tests.test_gpu_kernels import _svd
M = np.load("scripts/bad_theta.npy")
for lc in (1,):
    left, right, sv, info = _svd(M[None], 8, lc)
    print("  sv", sv[0][:5], "info", info[0], "recon %.2e iso %.2e" % (np.abs(left[0] @ right[0] - M).max(), np.abs(left[0].conj().T @ left[0] - np.eye(8)).max()))
left, right, sv, info = _svd(M[None], 8, 1)
'''
for env in ({}, {"MPSB_SVD_MAX_SWEEPS": "0"}, {"MPSB_SVD_NO_QR": "1"}, {"MPSB_SVD_MAX_SWEEPS": "1"}):
    print(env)
    e = dict(os.environ); e.update(env)
    print(subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True).stdout)
