"""What users do with a simulated state, on the device store:

* ``reduced_density_matrix``  -- ``mpsim/core.py:596-652``
* ``sample``                  -- ``mpsim/core.py:654-721``
* ``expectation``             -- ``mpsim/core.py:723-751``
* ``from_wavefunction``       -- ``mpsim/core.py:245-328``

Every contraction and factorisation goes through the C-ABI (``mpsb_cgemm`` with explicit
strides, ``mpsb_svd``); torch only allocates, reshapes and permutes device buffers.  There is no
CPU path: the host sees the final d^k x d^k matrix, the n x d table of marginals, or nothing.
"""
from typing import Any, Dict, List, Sequence, Tuple, Union

import numpy as np

from mpsim_b200 import _lib

MAX_GEMM_BATCH = 65535


def _cgemm(a_ptr: int, a_rs: int, a_cs: int, conj_a: bool, a_bs: int,
           b_ptr: int, b_rs: int, b_cs: int, conj_b: bool, b_bs: int,
           out, M: int, N: int, K: int, nbatch: int) -> None:
    """out[nbatch][M][N] (dense) = op(A) . op(B) with element strides, ``cgemm_kernel``."""
    lib = _lib.load(require_device=True)
    _lib.check(lib.mpsb_cgemm(a_ptr, a_rs, a_cs, 1 if conj_a else 0, a_bs, b_ptr, b_rs, b_cs, 1 if conj_b else 0, b_bs,
                              out.data_ptr(), N, M * N, M, N, K, nbatch, _lib.stream_ptr()), "mpsb_cgemm")


def _indices(mps, node_indices: Union[int, Sequence[int]]) -> Tuple[int, ...]:
    try:
        node_indices = iter(node_indices)
    except TypeError:
        node_indices = [node_indices]
    node_indices = tuple(int(i) for i in node_indices)
    if len(set(node_indices)) < len(node_indices):                 # core.py:617-618
        raise ValueError("Node indices contains duplicates.")
    if min(node_indices) < 0 or max(node_indices) > mps._nqudits - 1:    # core.py:620-621
        raise IndexError("One or more invalid node indices.")
    return node_indices


def reduced_density_matrix_device(mps, node_indices: Union[int, Sequence[int]]):
    """Device tensor [d^k][d^k]: rho[(kets in the given order), (bras in the given order)].

    One sweep over the chain, the environment kept as ``env[O][a][a']`` with ``O`` the open
    (ket, bra) legs of the kept sites seen so far; per site two strided batched GEMMs:

        T[o, a', (p, b)]        = sum_a  env[o, a, a'] A[a, (p, b)]
        traced: env'[o, b, b']  = sum_{a', p} T[o, (a', p), b] conj(A[(a', p), b'])
        kept  : env'[o, (p, b), (p', b')] = sum_{a'} T[o, a', (p, b)] conj(A[a', (p', b')])
    """
    import torch
    keep = _indices(mps, node_indices)
    chain = mps._chain
    d, n = chain.d, chain.n
    if d ** (2 * len(keep)) > MAX_GEMM_BATCH + 1:
        raise ValueError(f"reduced density matrix on {len(keep)} sites of dimension {d} is too large")
    dev = chain.device
    if min(chain.bonds) == 0:                      # maxsvals=0 somewhere: the zero state
        return torch.zeros((d ** len(keep), d ** len(keep)), dtype=torch.complex64, device=dev)
    env = torch.ones((1, 1, 1), dtype=torch.complex64, device=dev)
    labels: List[Tuple[str, int]] = []
    for i in range(n):
        A = chain.site_view(i)                     # [a][p][b], dense in its slot
        a, _, b = A.shape
        O = env.shape[0]
        envT = env.transpose(1, 2).contiguous()    # [o][a'][a]
        T = torch.empty((O, a, d, b), dtype=torch.complex64, device=dev)     # [o][a'][p][b]
        _cgemm(envT.data_ptr(), a, 1, False, 0, A.data_ptr(), d * b, 1, False, 0, T, O * a, d * b, a, 1)
        if i in keep:
            out = torch.empty((O, d, b, d, b), dtype=torch.complex64, device=dev)
            _cgemm(T.data_ptr(), 1, d * b, False, a * d * b, A.data_ptr(), d * b, 1, True, 0, out, d * b, d * b, a, O)
            env = out.permute(0, 1, 3, 2, 4).reshape(O * d * d, b, b)
            labels += [("k", i), ("b", i)]
        else:
            out = torch.empty((O, b, b), dtype=torch.complex64, device=dev)
            _cgemm(T.data_ptr(), 1, b, False, a * d * b, A.data_ptr(), b, 1, True, 0, out, b, b, a * d, O)
            env = out
    k = len(keep)
    rho = env.reshape([d] * (2 * k)) if k else env.reshape(())
    order = [labels.index(("k", i)) for i in keep] + [labels.index(("b", i)) for i in keep]
    return rho.permute(order).reshape(d ** k, d ** k)


def reduced_density_matrix(mps, node_indices: Union[int, Sequence[int]]) -> np.ndarray:
    return reduced_density_matrix_device(mps, node_indices).cpu().numpy()


def site_marginals(mps) -> np.ndarray:
    """float64 [n][d]: diagonal of every single-site reduced density matrix."""
    import torch
    diags = [torch.diagonal(reduced_density_matrix_device(mps, i)).real for i in range(mps._nqudits)]
    return torch.stack(diags).cpu().numpy().astype(np.float64)


def sample(mps, nsamples: int, as_hist: bool = False, as_string: bool = False) -> Any:
    """``mpsim/core.py:684-721``.  The reference draws every site from its OWN marginal of the
    unconditioned state (``core.py:665`` reads ``self``; the conditioned copy it builds is never
    used for the probabilities), i.e. from the product of the single-site marginals -- kept, and
    with the same ``np.random.choice`` calls in the same order, so a seeded run returns the
    reference's draws.  The marginals are computed once per call instead of once per draw.  As
    in the reference an unnormalised state is refused by ``np.random.choice``; marginals within
    float32 rounding (1e-4) of 1 are renormalised in float64 first."""
    if not isinstance(nsamples, int):
        raise ValueError(f"Arg nsamples should be an int but is a {type(nsamples)}.")
    if nsamples <= 0:
        raise ValueError(f"Arg nsamples should be positive but is {nsamples}.")
    if as_hist:
        as_string = True
    probs = np.clip(site_marginals(mps), 0.0, None)
    tot = probs.sum(axis=1)
    if np.any(np.abs(tot - 1.0) > 1e-4):
        raise ValueError("probabilities do not sum to 1")
    probs /= tot[:, None]
    states = list(range(mps._qudit_dimension))
    raw = []
    for _ in range(nsamples):
        string = [np.random.choice(states, size=1, p=probs[i])[0] for i in range(mps._nqudits)]
        raw.append("".join(str(bit) for bit in string) if as_string else string)
    if as_hist:
        hist: Dict[str, int] = {}
        for bitstring in raw:
            hist[bitstring] = hist.get(bitstring, 0) + 1
        return hist
    return raw


def expectation(mps, observable) -> float:
    """``mpsim/core.py:723-751``: Re <psi| O |psi> by applying O to a copy (non-unitary one-qudit
    observables take the orthonormalise-and-renormalise path exactly like the reference)."""
    if not observable.is_hermitian():
        raise ValueError("Observable is not Hermitian.")
    if observable.qudit_dimension != mps._qudit_dimension:
        obs_dim, mps_dim = observable.qudit_dimension, mps._qudit_dimension
        raise ValueError(f"Dimension mismatch between observable and MPS. Observable is ({obs_dim}, {obs_dim}) "
                         f"but MPS has qudit dimension {mps_dim}.")
    mps_copy = mps.copy()
    mps_copy.apply(observable)
    return mps.inner_product(mps_copy).real


def from_wavefunction(cls, wavefunction: Any, nqudits: int, qudit_dimension: int = 2, tensor_prefix: str = "q",
                      device: Any = None):
    """``mpsim/core.py:245-328``: SVD across every cut, nothing truncated, sqrt(S) handed to both
    sides of the cut like ``tn.split_node``.  The vector goes to the device once; each cut is one
    ``mpsb_svd`` plus two diagonal ``mpsb_cgemm`` products."""
    import torch
    from mpsim_b200.ortho import _device_svd, _device_matmul
    if not isinstance(wavefunction, (list, tuple, np.ndarray)):
        raise TypeError("Invalid type for wavefunction.")
    wavefunction = np.array(wavefunction)
    if len(wavefunction.shape) != 1:
        raise ValueError("Invalid shape for wavefunction. Should be a vector.")
    if nqudits < 2:
        raise ValueError("At least two qudits are required.")
    if wavefunction.size != qudit_dimension ** nqudits:
        raise ValueError(
            "Mismatch between wavefunction, qudit_dimension, and nqudits. "
            f"Expected {qudit_dimension ** nqudits} elements in the wavefunction, but wavefunction has "
            f"{wavefunction.size} elements.")
    d = int(qudit_dimension)
    mps = cls(nqudits, d, tensor_prefix, device=device)
    dev = mps._chain.device
    rest = torch.from_numpy(wavefunction.astype(np.complex64)).to(dev).reshape(1, -1)
    for i in range(nqudits - 1):
        chi = rest.shape[0]
        mat = rest.reshape(chi * d, -1)
        u, svh, sv = _device_svd(mat, True)                      # u [m][k], S.Vh [k][n], k = min(m, n)
        sq = torch.sqrt(torch.clamp(sv, min=0.0))
        inv = torch.where(sq > 0, 1.0 / sq, torch.zeros_like(sq))
        left = _device_matmul(u, torch.diag(sq).to(torch.complex64))           # U sqrt(S)
        rest = _device_matmul(torch.diag(inv).to(torch.complex64), svh)        # sqrt(S) Vh
        mps._chain.set_site(i, left.reshape(chi, d, -1), 0)
    mps._chain.set_site(nqudits - 1, rest.reshape(rest.shape[0], d, 1), 0)
    mps._last_bond_from_right = False        # every bond comes out of a split, left factor first
    return mps
