"""Times mpsb_svd on a batch of large matrices (block Jacobi path): python scripts/prof_large.py njobs m"""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from mpsim_b200 import _lib
njobs = int(sys.argv[1]) if len(sys.argv) > 1 else 50
m = int(sys.argv[2]) if len(sys.argv) > 2 else 512
lib = _lib.load(require_device=True)
rng = np.random.default_rng(0)
a = rng.standard_normal((njobs, m, m)) + 1j * rng.standard_normal((njobs, m, m))
u, s, vh = np.linalg.svd(a[:4])
s = s * np.exp(-np.arange(m) / m * 6.0)[None, :]
base = ((u * s[:, None, :]) @ vh).astype(np.complex64)
mats = np.concatenate([base] * ((njobs + 3) // 4))[:njobs]
x = torch.from_numpy(mats).cuda()
k = m // 2
left = torch.empty((njobs, m, k), dtype=torch.complex64, device="cuda")
right = torch.empty((njobs, k, m), dtype=torch.complex64, device="cuda")
sv = torch.empty((njobs, m), dtype=torch.float32, device="cuda")
info = torch.zeros((njobs, 2), dtype=torch.int32, device="cuda")
ws = torch.empty(lib.mpsb_svd_workspace_bytes(njobs, m, m), dtype=torch.uint8, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.mpsb_svd(x.data_ptr(), njobs, m, m, k, 1, left.data_ptr(), right.data_ptr(), sv.data_ptr(),
                            info.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
sref = np.linalg.svd(mats[0].astype(np.complex128), compute_uv=False)
err = np.abs(sv[0].cpu().numpy() - sref).max() / sref[0]
print(f"njobs={njobs} m={m}: {e0.elapsed_time(e1):.2f} ms  sweeps {info[:, 1].float().mean().item():.1f} status {int(info[:, 0].sum())} sigma err {err:.2e}")
