"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Independent dense state-vector simulator (complex128, n <= ~24 qudits).  It stands in for
Cirq's ``final_wavefunction`` which the reference's tests use as ground truth
(``mpsim/mpsim_cirq/simulator_test.py:113-144, 274-305``; ``mpsim/core_test.py:1225-1258``)
but which is not installable here.  Ordering is big-endian (qudit 0 most significant), the
same as ``MPS.wavefunction`` (``mpsim/core.py:483-500``, pinned by ``core_test.py:358-367``).
"""

from typing import Sequence, Tuple

import numpy as np


class DenseState:
    def __init__(self, nqudits: int, qudit_dimension: int = 2) -> None:
        self.n = nqudits
        self.d = qudit_dimension
        self.psi = np.zeros([qudit_dimension] * nqudits, dtype=np.complex128)
        self.psi[(0,) * nqudits] = 1.0

    def apply(self, tensor: np.ndarray, indices: Sequence[int]) -> None:
        """``tensor`` has shape (d,)*2k with output axes first (mpsim edge convention,
        ``mpsim/core.py:43-63``): out axes 0..k-1 replace qudits ``indices`` in order."""
        if isinstance(indices, int):
            indices = (indices,)
        k = len(indices)
        tensor = np.asarray(tensor, dtype=np.complex128).reshape([self.d] * (2 * k))
        psi = np.tensordot(tensor, self.psi, [list(range(k, 2 * k)), list(indices)])
        # tensordot puts the k output axes first; move them back to ``indices``
        self.psi = np.moveaxis(psi, list(range(k)), list(indices))

    def run(self, operations: Sequence[Tuple[np.ndarray, Sequence[int]]]) -> "DenseState":
        for tensor, indices in operations:
            self.apply(tensor, indices)
        return self

    def wavefunction(self) -> np.ndarray:
        return self.psi.reshape(-1).copy()


def fidelity(a: np.ndarray, b: np.ndarray) -> float:
    """|<a|b>|^2 / (<a|a><b|b>)."""
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    num = abs(np.vdot(a, b)) ** 2
    den = (np.vdot(a, a).real * np.vdot(b, b).real)
    return float(num / den) if den > 0 else 0.0
