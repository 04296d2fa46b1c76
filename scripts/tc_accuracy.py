"""Error of the tcgen05 3xTF32 complex GEMM and of the FFMA cgemm kernel against complex128, vs K;
and timing of the tcgen05 GEMM at theta-like shapes."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from mpsim_b200 import _lib

lib = _lib.load(require_device=True)
rng = np.random.RandomState(0)


def run_tc(a, b):
    nb, M, K = a.shape; N = b.shape[2]
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dc = torch.zeros((nb, M, N), dtype=torch.complex64, device="cuda")
    ws = torch.empty(lib.mpsb_cgemm_tc_workspace_bytes(M, N, K, nb) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mpsb_cgemm_tc(da.data_ptr(), M * K, db.data_ptr(), K * N, dc.data_ptr(), N, M * N, M, N, K, nb,
                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    return dc.cpu().numpy()


def run_ffma(a, b):
    nb, M, K = a.shape; N = b.shape[2]
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dc = torch.zeros((nb, M, N), dtype=torch.complex64, device="cuda")
    _lib.check(lib.mpsb_cgemm(da.data_ptr(), K, 1, 0, M * K, db.data_ptr(), N, 1, 0, K * N, dc.data_ptr(), N, M * N,
                              M, N, K, nb, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return dc.cpu().numpy()


for K in (16, 64, 256, 1024, 2048):
    M = N = 256
    a = (rng.randn(1, M, K) + 1j * rng.randn(1, M, K)).astype(np.complex64)
    b = (rng.randn(1, K, N) + 1j * rng.randn(1, K, N)).astype(np.complex64)
    ref = a.astype(np.complex128) @ b.astype(np.complex128)
    scale = np.abs(ref).max()
    for name, fn in (("tc", run_tc), ("ffma", run_ffma), ("torch", lambda x, y: (torch.from_numpy(x).cuda() @ torch.from_numpy(y).cuda()).cpu().numpy())):
        c = fn(a, b)
        e = c - ref
        print(f"K={K:5d} {name:5s} max|err|/max|C| = {np.abs(e).max() / scale:.2e}  rms = {np.sqrt((np.abs(e) ** 2).mean()) / scale:.2e}  mean(err.re) = {e.real.mean() / scale:+.2e}")

# timing at theta-like shapes (M = N = 2 chi, K = chi), njobs jobs
for chi, njobs in ((64, 512), (256, 50), (1024, 8)):
    M = N = 2 * chi; K = chi
    a = torch.randn((njobs, M, K), dtype=torch.complex64, device="cuda")
    b = torch.randn((njobs, K, N), dtype=torch.complex64, device="cuda")
    c = torch.zeros((njobs, M, N), dtype=torch.complex64, device="cuda")
    ws = torch.empty(lib.mpsb_cgemm_tc_workspace_bytes(M, N, K, njobs) + 256, dtype=torch.uint8, device="cuda")
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.mpsb_cgemm_tc(a.data_ptr(), M * K, b.data_ptr(), K * N, c.data_ptr(), N, M * N, M, N, K, njobs,
                                     ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 8.0 * M * N * K * njobs
    e0.record(); cc = a @ b; e1.record(); torch.cuda.synchronize()
    e0.record(); cc = a @ b; e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1)
    print(f"chi={chi} jobs={njobs}: tc (prep + gemm) {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s complex-equivalent; torch.matmul complex64 {ms_t:.3f} ms = {fl / ms_t / 1e9:.1f}")
