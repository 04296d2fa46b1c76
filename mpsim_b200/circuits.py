"""Synthetic circuits for the named benchmark shapes (BASELINE.json ``configs``).

A circuit here is a list of :class:`Op` -- ``(tensor, qudit_indices, keep_left_canonical)``.
``tensor`` follows the reference's gate edge convention (``mpsim/core.py:43-63``): shape
``(d,)*2k``, output axes first.  Pure numpy; nothing here touches the GPU.
"""

from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from mpsim_b200 import gates as _gates


class Op(NamedTuple):
    tensor: np.ndarray
    indices: Tuple[int, ...]
    keep_left_canonical: bool = True


def haar_two_qubit(rng: np.random.RandomState) -> np.ndarray:
    """Haar-random 4x4 unitary as a (2,2,2,2) tensor (``mpsim/gates.py:248-286``)."""
    return _gates.haar_random_unitary_tensor(2, 2, rng=rng)


def brickwork(nqubits: int, depth: int, seed: int) -> List[Op]:
    """1D random brickwork circuit of Haar two-qubit gates.

    Layer pattern = the reference's sweep helpers (``mpsim/core.py:1348-1360``): even layers
    act on bonds (0,1),(2,3),... left to right with ``keep_left_canonical=True``; odd layers
    on (n-3,n-2)... i.e. ``for i in range(n-2, 0, -2): (i-1, i)`` right to left with
    ``keep_left_canonical=False``.
    """
    rng = np.random.RandomState(seed)
    ops: List[Op] = []
    for layer in range(depth):
        if layer % 2 == 0:
            for i in range(0, nqubits - 1, 2):
                ops.append(Op(haar_two_qubit(rng), (i, i + 1), True))
        else:
            for i in range(nqubits - 2, 0, -2):
                ops.append(Op(haar_two_qubit(rng), (i - 1, i), False))
    return ops


def haar_gate_stack(count: int, rng: "np.random.Generator") -> np.ndarray:
    """``count`` Haar-random 4x4 unitaries (Mezzadri, ``mpsim/gates.py:269-286``) drawn in one
    vectorised call from a ``np.random.Generator``; shape ``(count, 4, 4)`` complex64."""
    z = (rng.standard_normal((count, 4, 4)) + 1j * rng.standard_normal((count, 4, 4))) / np.sqrt(2)
    q, r = np.linalg.qr(z)
    dg = np.diagonal(r, axis1=1, axis2=2)
    return (q * (dg / np.abs(dg))[:, None, :]).astype(np.complex64)


def batch_member_gates(nops: int, member: int) -> np.ndarray:
    """Gates of circuit ``member`` of the batched workload (BASELINE.json configs[3]): the
    stream seeded ``1000 + member``; shape ``(nops, 16)`` complex64.  ``bench.py``'s GPU arm, its
    CPU arms and the golden fixtures all take their gates from here."""
    return haar_gate_stack(nops, np.random.default_rng(1000 + member)).reshape(nops, 16)


def brickwork_member(nqubits: int, depth: int, member: int) -> List[Op]:
    """Circuit ``member`` of the batched workload: the brickwork layer pattern of
    :func:`brickwork` with the gates of :func:`batch_member_gates`."""
    structure = brickwork(nqubits, depth, seed=0)
    gates = batch_member_gates(len(structure), member)
    return [Op(gates[t].reshape(2, 2, 2, 2), op.indices, op.keep_left_canonical) for t, op in enumerate(structure)]


def ghz(nqubits: int) -> List[Op]:
    """H(0) then CNOT(0, i), i = 1..n-1 (``mpsim/core_test.py:1236-1245``)."""
    ops = [Op(_gates.hgate().tensor, (0,))]
    for i in range(1, nqubits):
        ops.append(Op(_gates.cnot().tensor, (0, i)))
    return ops


def qft(nqubits: int) -> List[Op]:
    """QFT as in ``mpsim/core_test.py:1248-1258`` / ``mpsim_cirq/simulator_test.py:126-144``:
    for i = n-1..0: H(i), then CPhase(2**(j-i)) on (j, i) for j = i-1..0."""
    ops: List[Op] = []
    for i in range(nqubits - 1, -1, -1):
        ops.append(Op(_gates.hgate().tensor, (i,)))
        for j in range(i - 1, -1, -1):
            ops.append(Op(_gates.cphase(2.0 ** (j - i)).tensor, (j, i)))
    return ops


def ghz_qft(nqubits: int) -> List[Op]:
    """BASELINE.json config 2: GHZ preparation followed by the QFT chain."""
    return ghz(nqubits) + qft(nqubits)


def sycamore_snake(rows: int = 6, cols: int = 9, drop_last: int = 1, cycles: int = 14,
                   seed: int = 5) -> Tuple[int, List[Op]]:
    """BASELINE.json config 5 (synthetic layout): a rows x cols grid minus the
    last ``drop_last`` sites, qubits numbered in boustrophedon (snake) order, coupler pattern
    ABCDCDAB over ``cycles`` cycles, one Haar ("fSim-like") two-qubit gate per active coupler.
    Non-adjacent pairs in the 1D order are routed by the swap networks of
    ``mpsim/core.py:1035-1043, 1154-1158``.  Returns ``(nqubits, ops)``."""
    nq = rows * cols - drop_last
    index = {}
    k = 0
    for r in range(rows):
        cs = range(cols) if r % 2 == 0 else range(cols - 1, -1, -1)
        for c in cs:
            if k < nq:
                index[(r, c)] = k
            k += 1
    # four coupler classes: horizontal even/odd column, vertical even/odd row
    def couplers(kind: str) -> List[Tuple[int, int]]:
        out = []
        for r in range(rows):
            for c in range(cols):
                if kind in "AB":       # horizontal
                    if c + 1 < cols and (c % 2 == 0) == (kind == "A"):
                        a, b = (r, c), (r, c + 1)
                    else:
                        continue
                else:                  # vertical
                    if r + 1 < rows and (r % 2 == 0) == (kind == "C"):
                        a, b = (r, c), (r + 1, c)
                    else:
                        continue
                if a in index and b in index:
                    i, j = index[a], index[b]
                    out.append((min(i, j), max(i, j)))
        return sorted(out)
    pattern = "ABCDCDAB"
    rng = np.random.RandomState(seed)
    ops: List[Op] = []
    for cyc in range(cycles):
        for (i, j) in couplers(pattern[cyc % len(pattern)]):
            ops.append(Op(haar_two_qubit(rng), (i, j), True))
    return nq, ops


def grid_snake(rows: int, cols: int, warmup_depth: int, cycles: int, seed: int) -> Tuple[int, List[Op]]:
    """A small full-rank relative of :func:`sycamore_snake` for parity fixtures: ``warmup_depth``
    brickwork layers of Haar gates first (they saturate every bond, so that no later theta is
    rank deficient and the reference's result is well defined -- DESIGN.md section 2), then ``cycles``
    cycles of the ABCD coupler pattern on the snake-ordered ``rows x cols`` grid, whose vertical
    couplers are routed through swap networks."""
    nq = rows * cols
    ops = brickwork(nq, warmup_depth, seed)
    _, tail = sycamore_snake(rows, cols, 0, cycles, seed + 1)
    return nq, ops + tail


def count_adjacent_applications(ops: Sequence[Op]) -> int:
    """Number of adjacent-bond applications the reference executes for ``ops``: a gate at
    distance D costs 2(D-1) SWAPs + 1 (``mpsim/core.py:1036-1043, 1155-1158, 1187-1190``)."""
    total = 0
    for op in ops:
        if len(op.indices) == 2:
            dist = abs(op.indices[0] - op.indices[1])
            total += 2 * (dist - 1) + 1
    return total
