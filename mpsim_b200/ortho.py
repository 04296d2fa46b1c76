"""Edge orthonormalisation after non-unitary one-qudit gates.

``mpsim/core.py:847-939``: SVD one site with a DATA-DEPENDENT cut (drop the largest tail of
singular values whose 2-norm is <= threshold * norm) and push ``S.Vh`` (or ``U.S``) into the
neighbour.  The SVD and the absorb GEMM run on the device (``mpsb_svd`` / ``mpsb_cgemm``); only
the singular values come back to the host to decide the cut, as the reference's
``max_truncation_err`` semantics require.
"""
from typing import Tuple

import numpy as np

from mpsim_b200 import _lib


def _device_svd(mat, left_canonical: bool):
    """mat: torch complex64 [m][n] on the device -> (left [m][k], right [k][n], svals [k]), k = min(m, n)."""
    import torch
    lib = _lib.load(require_device=True)
    m, n = mat.shape
    k = min(m, n)
    mat = mat.contiguous()
    left = torch.empty((m, k), dtype=torch.complex64, device=mat.device)
    right = torch.empty((k, n), dtype=torch.complex64, device=mat.device)
    sv = torch.empty(k, dtype=torch.float32, device=mat.device)
    info = torch.zeros(2, dtype=torch.int32, device=mat.device)
    need = lib.mpsb_svd_workspace_bytes(1, m, n)
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=mat.device)
    _lib.check(lib.mpsb_svd(mat.data_ptr(), 1, m, n, k, 1 if left_canonical else 0, left.data_ptr(),
                            right.data_ptr(), sv.data_ptr(), info.data_ptr(), ws.data_ptr(), ws.numel(),
                            _lib.stream_ptr()), "mpsb_svd")
    status = int(info[0].item())
    if status != 0:
        import warnings
        from mpsim_b200.store import SVDNotConverged
        warnings.warn(f"mpsb_svd of a {m} x {n} site reached its sweep limit while still rotating",
                      SVDNotConverged, stacklevel=3)
    return left, right, sv


def _device_matmul(a, b):
    """Complex64 [M][K] @ [K][N] through the library's own GEMM kernel."""
    import torch
    lib = _lib.load(require_device=True)
    a, b = a.contiguous(), b.contiguous()
    M, K = a.shape
    K2, N = b.shape
    assert K == K2
    c = torch.empty((M, N), dtype=torch.complex64, device=a.device)
    _lib.check(lib.mpsb_cgemm(a.data_ptr(), K, 1, 0, 0, b.data_ptr(), N, 1, 0, 0, c.data_ptr(), N, 0,
                              M, N, K, 1, _lib.stream_ptr()), "mpsb_cgemm")
    return c


#: numerically zero singular values of a rank-deficient site come out of the fp32 Jacobi SVD at
#: ~eps32 * sigma_max * sqrt(min(m, n)); the reference's complex128 SVD returns them at ~1e-16 and
#: drops them with its 1e-8 * norm cut.  The cut is clamped to this floor so that the bond dimension
#: after a non-unitary gate matches the reference's.
_FP32_ZERO_FLOOR = 8.0 * float(np.finfo(np.float32).eps)


def _keep_count(svals: np.ndarray, max_truncation_err: float) -> int:
    """tensornetwork 0.2.1 svd_decomposition: number of values whose tail norm exceeds the bound
    (the bound clamped to the fp32 noise floor of the device SVD, see ``_FP32_ZERO_FLOOR``)."""
    sv = svals.astype(np.float64)
    if sv.size:
        max_truncation_err = max(max_truncation_err, _FP32_ZERO_FLOOR * sv.max() * np.sqrt(sv.size))
    trunc_errs = np.sqrt(np.cumsum(np.square(sv[::-1])))
    return int(np.count_nonzero(trunc_errs > max_truncation_err))


def orthonormalize_right_edge_of(mps, node_index: int, threshold: float = 1e-8) -> None:
    """``mpsim/core.py:847-892``: site <- U, right neighbour <- S.Vh.neighbour."""
    if not 0 <= node_index < mps._nqudits - 1:
        raise ValueError("Invalid edge index.")
    chain = mps._chain
    with mps._device_guard():
        _right(mps, chain, node_index, threshold)


def _right(mps, chain, node_index: int, threshold: float) -> None:
    a = chain.site_view(node_index).clone()
    cl, d, cr = a.shape
    err = threshold * mps.norm()
    u, svh, sv = _device_svd(a.reshape(cl * d, cr), True)
    keep = _keep_count(sv.cpu().numpy(), err)
    nxt = chain.site_view(node_index + 1).clone()
    new_next = _device_matmul(svh[:keep].contiguous(), nxt.reshape(cr, -1)) if keep else svh[:0] @ nxt.reshape(cr, -1)
    chain.set_site(node_index, u[:, :keep].contiguous().reshape(cl, d, keep), 0)
    chain.set_site(node_index + 1, new_next.reshape(keep, d, nxt.shape[2]), 0)


def orthonormalize_left_edge_of(mps, node_index: int, threshold: float = 1e-8) -> None:
    """``mpsim/core.py:894-939``: site <- Vh, left neighbour <- neighbour.U.S."""
    if not 0 < node_index <= mps._nqudits - 1:
        raise ValueError("Invalid edge index.")
    chain = mps._chain
    with mps._device_guard():
        _left(mps, chain, node_index, threshold)


def _left(mps, chain, node_index: int, threshold: float) -> None:
    a = chain.site_view(node_index).clone()
    cl, d, cr = a.shape
    err = threshold * mps.norm()
    us, vh, sv = _device_svd(a.reshape(cl, d * cr), False)
    keep = _keep_count(sv.cpu().numpy(), err)
    prv = chain.site_view(node_index - 1).clone()
    pl = prv.shape[0]
    new_prev = _device_matmul(prv.reshape(-1, cl), us[:, :keep].contiguous()) if keep else prv.reshape(-1, cl) @ us[:, :0]
    chain.set_site(node_index, vh[:keep].contiguous().reshape(keep, d, cr), 0)
    chain.set_site(node_index - 1, new_prev.reshape(pl, d, keep), 0)
