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
