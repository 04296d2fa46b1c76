"""Small driver for ncu: a batch of 128x128 complex64 SVDs through mpsb_svd (k = 64)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from mpsim_b200 import _lib

njobs = int(sys.argv[1]) if len(sys.argv) > 1 else 148
m = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lib = _lib.load(require_device=True)
import os
rng = np.random.default_rng(0)
a = rng.standard_normal((njobs, m, m)) + 1j * rng.standard_normal((njobs, m, m))
u, s, vh = np.linalg.svd(a)
s = s * np.exp(-np.arange(m) / m * 8.0)[None, :]
mats = ((u * s[:, None, :]) @ vh).astype(np.complex64)
x = torch.from_numpy(mats).cuda()
k = m // 2
left = torch.empty((njobs, m, k), dtype=torch.complex64, device="cuda")
right = torch.empty((njobs, k, m), dtype=torch.complex64, device="cuda")
sv = torch.empty((njobs, m), dtype=torch.float32, device="cuda")
info = torch.zeros((njobs, 2), dtype=torch.int32, device="cuda")
ws = torch.empty(lib.mpsb_svd_workspace_bytes(njobs, m, m), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.mpsb_svd(x.data_ptr(), njobs, m, m, k, 1, left.data_ptr(), right.data_ptr(), sv.data_ptr(),
                            info.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    print("ms", e0.elapsed_time(e1), "sweeps", info[:, 1].float().mean().item(), "status", int(info[:, 0].sum()))

import ctypes
clk = (ctypes.c_longlong * 32)()
if hasattr(lib, "mpsb_debug_phase_clocks") and lib.mpsb_debug_phase_clocks(clk) == 0:
    names = ["load", "qr1", "jacobi", "sort", "W=XV", "qr2+formQ", "P=QhX", "write"]
    c = list(clk)
    print("phase cycles (CTA 0):", {n: c[i + 1] - c[i] for i, n in enumerate(names)}, "total", c[8] - c[0])
