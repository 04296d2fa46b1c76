import numpy as np


def density_matrix_from_state_vector(state, indices=None):
    """Reduced density matrix of a state vector on ``indices`` (big-endian qubits)."""
    state = np.asarray(state)
    n = int(round(np.log2(state.size)))
    psi = state.reshape([2] * n)
    if indices is None:
        indices = list(range(n))
    indices = list(indices)
    rest = [i for i in range(n) if i not in indices]
    psi = np.transpose(psi, indices + rest).reshape(2 ** len(indices), -1)
    return psi @ psi.conj().T
