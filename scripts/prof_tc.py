"""Small driver for ncu: theta through the tcgen05 kernel at chi (default 1024), a few launches."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from mpsim_b200 import _lib
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
jobs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = _lib.load(require_device=True)
d = 2
A = torch.randn((jobs, chi, d, chi), dtype=torch.complex64, device="cuda")
B = torch.randn((jobs, chi, d, chi), dtype=torch.complex64, device="cuda")
G = torch.randn((jobs, 16), dtype=torch.complex64, device="cuda")
desc = np.zeros(1, dtype=_lib.GATE2_DESC)
desc[0] = (A.data_ptr(), B.data_ptr(), 0, 0, G.data_ptr(), 0, chi * d * chi, chi * d * chi, 0, 0, 16, 0)
ddesc = _lib.to_device_bytes(desc, "cuda")
theta = torch.empty((jobs, d * chi, d * chi), dtype=torch.complex64, device="cuda")
ws = torch.empty(lib.mpsb_theta_workspace_bytes(1, jobs, d, chi, chi, chi) + 256, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.mpsb_theta(ddesc.data_ptr(), 1, jobs, d, chi, chi, chi, theta.data_ptr(), ws.data_ptr(), ws.numel(),
                              _lib.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"chi={chi} jobs={jobs}: {ms:.3f} ms, {(8.0 * 4 * chi ** 3 + 8 * 16 * chi * chi) * jobs / ms / 1e9:.1f} TFLOP/s complex-equivalent")
