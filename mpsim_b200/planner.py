"""Static shape planner and moment dispatcher (host side, no GPU needed).

The kept singular-value count of a two-qudit application is data independent,
``k = min(maxsvals, d*chi_left, d*chi_right)`` -- zeros are kept
(``mpsim/core.py:1105-1137`` + tensornetwork's ``split_node_full_svd``; pinned by
``mpsim/core_test.py:932-944``).  Every tensor shape of a run is therefore known from the
circuit alone.  The planner

  * expands non-adjacent gates into the reference's swap networks
    (``mpsim/core.py:1030-1043, 1154-1158, 1163-1219``), each SWAP a full application with the
    same kwargs;
  * resolves ``maxsvals`` / ``fraction`` exactly as ``mpsim/core.py:1105-1130`` does;
  * tracks bond dimensions and assigns every adjacent application its ``(chiL, chiM, chiR, k)``;
  * schedules operations ASAP into layers of operations on disjoint sites (the reference
    applies them one by one, ``# TODO: Parallelize`` at ``mpsim/core.py:1245``) and groups each
    layer by shape class so one C-ABI call launches a whole group.
"""
from typing import Any, Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

SWAP_TENSOR = np.array([[1.0, 0, 0, 0], [0, 0, 1.0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]]).reshape(2, 2, 2, 2)


class App1(NamedTuple):
    """One-qudit application."""
    site: int
    gate_index: int          # index into the plan's gate list
    layer: int


class App2(NamedTuple):
    """Adjacent two-qudit application on (site, site+1)."""
    site: int
    gate_index: int
    chiL: int
    chiM: int
    chiR: int
    k: int
    left_canonical: bool
    layer: int
    source_op: int           # index of the user-level operation it came from
    is_swap: bool


def max_bond_dimensions(nqudits: int, d: int) -> List[int]:
    """``mpsim/core.py:235-242``."""
    mbd = [d ** (i + 1) for i in range(nqudits // 2)]
    mbd += list(reversed(mbd))
    if nqudits % 2 == 0:
        mbd.remove(d ** (nqudits // 2))
    return mbd


def resolve_truncation(kwargs: Dict[str, Any], nqudits: int, d: int, index: int) -> Tuple[bool, Optional[int]]:
    """keep_left_canonical and maxsvals of one adjacent application (``mpsim/core.py:1105-1130``)."""
    keep_left = kwargs["keep_left_canonical"] if "keep_left_canonical" in kwargs else True
    if "fraction" in kwargs and "maxsvals" in kwargs:
        raise ValueError("Only one of (fraction, maxsvals) can be provided as kwargs.")
    maxsvals = None
    if "fraction" in kwargs:
        fraction = kwargs.get("fraction")
        if not (0 <= fraction <= 1):
            raise ValueError("Keyword fraction must be between 0 and 1 but is", fraction)
        maxsvals = int(round(fraction * max_bond_dimensions(nqudits, d)[index]))
    if "maxsvals" in kwargs:
        maxsvals = int(kwargs.get("maxsvals"))
    return bool(keep_left), maxsvals


class Plan:
    """Result of planning a list of operations against a chain with given bond dimensions."""

    def __init__(self, nqudits: int, d: int, bonds: Sequence[int]) -> None:
        self.n = nqudits
        self.d = d
        self.bonds_in = list(bonds)          # len n+1, ends are 1
        self.bonds = list(bonds)
        self.caps = list(bonds)              # running max of every bond
        self.gates: List[np.ndarray] = []    # flattened complex64 tensors, d^2 or d^4 entries
        self.gate_src: List[Tuple[int, bool]] = []   # per gate: (user op index or -1 for a SWAP, flipped)
        self.apps1: List[App1] = []
        self.apps2: List[App2] = []
        self.order: List[Tuple[int, int]] = []    # program order: (1|2, index into apps1/apps2)
        self._ready = [0] * nqudits
        self.nlayers = 0

    # -- building ----------------------------------------------------------------------------
    def _add_gate(self, tensor: np.ndarray, src: int = -1, flipped: bool = False) -> int:
        self.gates.append(np.ascontiguousarray(tensor, dtype=np.complex64).reshape(-1))
        self.gate_src.append((src, flipped))
        return len(self.gates) - 1

    def add_one(self, tensor: np.ndarray, site: int, source_op: int = -1) -> None:
        layer = self._ready[site]
        self._ready[site] = layer + 1
        self.nlayers = max(self.nlayers, layer + 1)
        self.apps1.append(App1(site, self._add_gate(tensor, source_op), layer))
        self.order.append((1, len(self.apps1) - 1))

    def _add_adjacent(self, tensor: np.ndarray, site: int, kwargs: Dict[str, Any], source_op: int,
                      is_swap: bool, flipped: bool = False) -> None:
        d = self.d
        if d > MAX_DEVICE_QUDIT_DIMENSION:
            raise ValueError(f"two-qudit gates are implemented on the device for qudit dimension 2..{MAX_DEVICE_QUDIT_DIMENSION} "
                             f"(theta kernel instantiations), not {d}")
        keep_left, maxsvals = resolve_truncation(kwargs, self.n, d, site)
        chiL, chiM, chiR = self.bonds[site], self.bonds[site + 1], self.bonds[site + 2]
        full = min(d * chiL, d * chiR)
        k = full if maxsvals is None else min(maxsvals, full)
        if k < 0:
            k = 0
        layer = max(self._ready[site], self._ready[site + 1])
        self._ready[site] = self._ready[site + 1] = layer + 1
        self.nlayers = max(self.nlayers, layer + 1)
        gate_index = self._add_gate(tensor, -1 if is_swap else source_op, flipped)
        self.apps2.append(App2(site, gate_index, chiL, chiM, chiR, k, keep_left, layer, source_op, is_swap))
        self.order.append((2, len(self.apps2) - 1))
        self.bonds[site + 1] = k
        self.caps[site + 1] = max(self.caps[site + 1], k)

    def add_two(self, tensor: np.ndarray, i: int, j: int, kwargs: Dict[str, Any], source_op: int = -1) -> None:
        """Two-qudit gate on arbitrary (i, j): flip + swap network as ``mpsim/core.py:1030-1043,
        1154-1158``."""
        d = self.d
        tensor = np.asarray(tensor).reshape(d, d, d, d)
        flipped = j < i
        if flipped:
            tensor = np.transpose(tensor, (1, 0, 3, 2))
            i, j = j, i
        if i < j - 1 and d != 2:
            raise ValueError("SWAP routing of non-adjacent gates is only defined for qubits "
                             "(mpsim/core.py:1188 'SWAP is only for qubits').")
        for s in range(i, j - 1):                       # move_node_from_left_to_right(i, j-1)
            self._add_adjacent(SWAP_TENSOR, s, kwargs, source_op, True)
        self._add_adjacent(tensor, j - 1, kwargs, source_op, False, flipped)
        for s in range(j - 2, i - 1, -1):               # move_node_from_right_to_left(j-1, i)
            self._add_adjacent(SWAP_TENSOR, s, kwargs, source_op, True)

    # -- queries -----------------------------------------------------------------------------
    def layers(self) -> List[Dict[str, Any]]:
        """Per layer: the one-qudit applications and the two-qudit applications grouped by
        shape class ``(chiL, chiM, chiR, k, left_canonical)``."""
        out = [dict(one=[], two={}) for _ in range(self.nlayers)]
        for idx, a in enumerate(self.apps1):
            out[a.layer]["one"].append(idx)
        for idx, a in enumerate(self.apps2):
            key = (a.chiL, a.chiM, a.chiR, a.k, a.left_canonical)
            out[a.layer]["two"].setdefault(key, []).append(idx)
        return out

    def gate_table(self, width: int) -> np.ndarray:
        """All gates as one complex64 array [ngates][width] (zero padded)."""
        tab = np.zeros((len(self.gates), width), dtype=np.complex64)
        for g, t in enumerate(self.gates):
            tab[g, :t.size] = t
        return tab

    def counts(self) -> Dict[str, int]:
        return dict(one_qudit=len(self.apps1), adjacent_applications=len(self.apps2),
                    swaps=sum(a.is_swap for a in self.apps2), layers=self.nlayers)


MAX_DEVICE_QUDIT_DIMENSION = 4       # csrc/theta.cu: theta_kernel<2..4>


def plan_operations(nqudits: int, d: int, bonds: Sequence[int],
                    ops: Sequence[Tuple[np.ndarray, Tuple[int, ...], Dict[str, Any]]]) -> Plan:
    """ops: (tensor, qudit_indices, kwargs).  Validates like ``mpsim/core.py:785-805, 1003-1028,
    1260-1276`` before anything is launched."""
    plan = Plan(nqudits, d, bonds)
    for t, (tensor, indices, kwargs) in enumerate(ops):
        tensor = np.asarray(tensor)
        nq = len(indices)
        if nq not in (1, 2):
            raise ValueError(
                "Only one-qudit and two-qudit gates are supported. To apply a gate on three or more "
                "qudits, the gate must be compiled into a sequence of one- and two-qudit gates.")
        for i in indices:
            if i not in range(nqudits):
                raise ValueError(f"Input tensor indices={tuple(indices)} are out of bounds for an MPS "
                                 f"on {nqudits} qudits.")
        if tensor.ndim != 2 * nq:
            raise ValueError("Two qubit gate must have four free edges and zero connected edges."
                             if nq == 2 else
                             "Single qudit gate must have two free edges and zero connected edges.")
        if len(set(tensor.shape)) != 1:
            raise ValueError("All gate edges must have the same dimension." if nq == 2
                             else "Gate edge dimensions must be equal.")
        if tensor.shape[0] != d:
            raise ValueError(f"Gate edges have dimension {tensor.shape[0]} but should have MPS qudit "
                             f"dimension = {d}")
        if nq == 1:
            plan.add_one(tensor, indices[0], t)
        else:
            if indices[0] == indices[1]:
                raise ValueError("Node indices cannot be identical.")
            plan.add_two(tensor, indices[0], indices[1], kwargs, t)
    return plan
