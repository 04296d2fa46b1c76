"""Gate library with the reference's names, dtypes and quirks (``mpsim/gates.py``).

Functions (not constants) return fresh :class:`~mpsim_b200.node.Node` objects, as in the
reference (``gates.py:112-138``).  dtypes follow the reference: one-qubit gates and
``cphase`` are complex64, ``cnot`` / ``swap`` are float64 (``gates.py:195-232``).
"""

from typing import Any, Optional

import numpy as np

from mpsim_b200.node import Node, tensor_of


def _as_matrix(gate: Any) -> np.ndarray:
    if not (isinstance(gate, np.ndarray) or hasattr(gate, "tensor")):
        raise TypeError("Invalid type for gate.")
    gate = tensor_of(gate)
    if len(gate.shape) > 2:
        if len(set(gate.shape)) != 1:
            raise ValueError("Gate shape should be of the form (d, d, ..., d).")
        dim = int(np.sqrt(gate.size))
        gate = np.reshape(gate, (dim, dim))
    return gate


def is_unitary(gate: Any) -> bool:
    """``gates.py:15-33``: ``allclose(G^dag G, I, atol=1e-5)``."""
    gate = _as_matrix(gate)
    return bool(np.allclose(gate.conj().T @ gate, np.identity(gate.shape[0]), atol=1e-5))


def is_hermitian(gate: Any) -> bool:
    """``gates.py:36-52``."""
    gate = _as_matrix(gate)
    return bool(np.allclose(gate.conj().T, gate, atol=1e-5))


def is_projector(gate: Any) -> bool:
    """``gates.py:55-70`` (rank-one test, as in the reference)."""
    gate = _as_matrix(gate)
    return bool(np.linalg.matrix_rank(gate) == 1)


zero_state = np.array([1.0, 0.0], dtype=np.complex64)
one_state = np.array([0.0, 1.0], dtype=np.complex64)
plus_state = 1.0 / np.sqrt(2) * (zero_state + one_state)


def computational_basis_state(state: int, dim: int = 2) -> Node:
    """``gates.py:79-101``."""
    if state < 0:
        raise ValueError(f"Argument state should be positive but is {state}.")
    if dim < 0:
        raise ValueError(f"Argument dim should be positive but is {dim}.")
    if state >= dim:
        raise ValueError(f"Requires state < dim but state = {state} and dim = {dim}.")
    vector = np.zeros((dim,))
    vector[state] = 1.0
    return Node(vector, name=f"|{state}>")


_hmatrix = 1 / np.sqrt(2) * np.array([[1.0, 1.0], [1.0, -1.0]], dtype=np.complex64)
_imatrix = np.array([[1.0, 0.0], [0.0, 1.0]], dtype=np.complex64)
_xmatrix = np.array([[0.0, 1.0], [1.0, 0.0]], dtype=np.complex64)
_ymatrix = np.array([[0.0, -1j], [1j, 0.0]], dtype=np.complex64)
_zmatrix = np.array([[1.0, 0.0], [0.0, -1.0]], dtype=np.complex64)


def igate() -> Node:
    return Node(_imatrix.copy(), name="igate")


def xgate() -> Node:
    return Node(_xmatrix.copy(), name="xgate")


def ygate() -> Node:
    return Node(_ymatrix.copy(), name="ygate")


def zgate() -> Node:
    return Node(_zmatrix.copy(), name="zmat")


def hgate() -> Node:
    return Node(_hmatrix.copy(), name="hgate")


def rgate(seed: Optional[int] = None, angle_scale: float = 1.0) -> Node:
    """Random one-qubit gate of arXiv:2002.07730 as written in ``gates.py:141-164``,
    including its quirks: ``if seed:`` ignores seed 0, and the generator is
    ``mx X + my Y * mz Z`` with an elementwise product of Y and Z (``gates.py:149,162``)."""
    from scipy.linalg import expm
    if seed:
        np.random.seed(seed)
    theta, alpha, phi = np.random.rand(3) * 2 * np.pi
    mx = np.sin(alpha) * np.cos(phi)
    my = np.sin(alpha) * np.sin(phi)
    mz = np.cos(alpha)
    theta *= angle_scale
    unitary = expm(-1j * theta * (mx * _xmatrix + my * _ymatrix * mz * _zmatrix))
    return Node(unitary)


def computational_basis_projector(state: int, dim: int = 2) -> Node:
    """``gates.py:168-191``."""
    if state < 0:
        raise ValueError(f"Argument state should be positive but is {state}.")
    if dim < 0:
        raise ValueError(f"Argument dim should be positive but is {dim}.")
    if state >= dim:
        raise ValueError(f"Requires state < dim but state = {state} and dim = {dim}.")
    projector = np.zeros((dim, dim))
    projector[state, state] = 1.0
    return Node(projector, name=f"|{state}><{state}|")


_cnot_matrix = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0],
                         [0.0, 0.0, 0.0, 1.0], [0.0, 0.0, 1.0, 0.0]]).reshape(2, 2, 2, 2)
_swap_matrix = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0],
                         [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]).reshape(2, 2, 2, 2)


def cnot() -> Node:
    return Node(_cnot_matrix.copy(), name="cnot")


def swap() -> Node:
    return Node(_swap_matrix.copy(), name="swap")


def cphase(exp: float) -> Node:
    """``gates.py:220-232``: diag(1, 1, 1, exp(2 pi i exp)), complex64."""
    matrix = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0],
                       [0.0, 0.0, 0.0, np.exp(1j * 2 * np.pi * exp)]], dtype=np.complex64)
    return Node(matrix.reshape(2, 2, 2, 2), name="cphase")


def random_two_qubit_gate(seed: Optional[int] = None) -> Node:
    """``gates.py:235-245`` (scipy ``unitary_group``; ``if seed:`` ignores 0)."""
    from scipy.stats import unitary_group
    if seed:
        np.random.seed(seed)
    unitary = unitary_group.rvs(dim=4)
    return Node(np.reshape(unitary, (2, 2, 2, 2)).copy(), name="R2Q")


def haar_random_unitary_tensor(nqudits: int = 2, qudit_dimension: int = 2,
                               seed: Optional[int] = None, rng: Any = None) -> np.ndarray:
    """Mezzadri's algorithm as in ``gates.py:269-286``; returns the bare tensor."""
    if rng is None:
        rng = np.random.RandomState(seed)
    units = np.array([1, 1j])
    shape = (qudit_dimension ** nqudits, qudit_dimension ** nqudits)
    mat = np.sum(rng.randn(*(shape + (2,))) * units, axis=-1) / np.sqrt(2)
    qmat, rmat = np.linalg.qr(mat)
    diag = np.diag(rmat).copy()
    diag /= np.abs(diag)
    return np.reshape(qmat * diag, [qudit_dimension] * 2 * nqudits)


def haar_random_unitary(nqudits: int = 2, qudit_dimension: int = 2, name: str = "Haar",
                        seed: Optional[int] = None) -> Node:
    """``gates.py:248-286``."""
    return Node(haar_random_unitary_tensor(nqudits, qudit_dimension, seed=seed), name=name)
