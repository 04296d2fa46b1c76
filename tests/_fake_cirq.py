"""Tiny duck-typed circuit objects implementing the protocol MPSimulator consumes
(all_qubits / all_operations / op.qubits / op._has_unitary_ / op._unitary_ /
_resolve_parameters_) -- stands in for cirq~=0.8, which is not installable here."""
import numpy as np


class Op:
    def __init__(self, qubits, unitary=None, param=None, maker=None):
        self.qubits = tuple(qubits)
        self._u, self._param, self._maker = unitary, param, maker

    def _has_unitary_(self):
        return self._u is not None

    def _unitary_(self):
        return None if self._u is None else np.array(self._u, copy=True)

    def resolve(self, resolver):
        if self._param is None or resolver is None:
            return self
        return Op(self.qubits, self._maker(resolver[self._param]))


class Circuit:
    def __init__(self, ops):
        self._ops = list(ops)

    def all_qubits(self):
        return frozenset(q for op in self._ops for q in op.qubits)

    def all_operations(self):
        return iter(self._ops)

    def _resolve_parameters_(self, resolver):
        return Circuit([op.resolve(resolver) for op in self._ops])


_H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
_CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)


def H(q):
    return Op((q,), _H)


def CNOT(a, b):
    return Op((a, b), _CNOT)


def CZPow(exponent, a, b):
    return Op((a, b), np.diag([1, 1, 1, np.exp(1j * np.pi * exponent)]))


def _rx(theta):
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -1j * s], [-1j * s, c]])


def Rx(param, q):
    return Op((q,), None if isinstance(param, str) else _rx(param), param if isinstance(param, str) else None, _rx)


def Toffoli(a, b, c):
    u = np.eye(8, dtype=complex)
    u[6:, 6:] = [[0, 1], [1, 0]]
    return Op((a, b, c), u)


# ---- the gate domain of mpsim/mpsim_cirq/simulator_test.py:279-292 and a random-circuit generator in the
# ---- spirit of cirq.testing.random_circuit (moments of operations on disjoint, randomly chosen qubits)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.diag([1, -1]).astype(complex)
_S = np.diag([1, 1j]).astype(complex)
_T = np.diag([1, np.exp(0.25j * np.pi)]).astype(complex)
_CZ = np.diag([1, 1, 1, -1]).astype(complex)
_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
_ISWAP = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=complex)


def _fsim(theta, phi):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [0, 0, 0, np.exp(-1j * phi)]], dtype=complex)


GATE_DOMAIN = [(_X, 1), (_Y, 1), (_Z, 1), (_H.astype(complex), 1), (_S, 1), (_T, 1),
               (_CNOT, 2), (_CZ, 2), (_SWAP, 2), (_CZ, 2), (_ISWAP, 2), (_fsim(0.2, 0.3), 2)]


def random_circuit(nqubits, n_moments, op_density, rng):
    """Every moment: shuffle the qubits, then place gates drawn from GATE_DOMAIN on consecutive free
    qubits of that order with probability ``op_density`` each (two-qubit gates therefore act on
    arbitrary, mostly non-adjacent pairs).  Every qubit is touched at least once."""
    ops = []
    for _ in range(n_moments):
        free = list(rng.permutation(nqubits))
        while free:
            u, nq = GATE_DOMAIN[rng.randint(len(GATE_DOMAIN))]
            if nq > len(free):
                u, nq = GATE_DOMAIN[rng.randint(6)]
            qs = [free.pop() for _ in range(nq)]
            if rng.rand() < op_density:
                ops.append(Op(tuple(int(q) for q in qs), u))
    ops += [Op((q,), np.eye(2, dtype=complex)) for q in range(nqubits)]
    return Circuit(ops)
