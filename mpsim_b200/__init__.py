"""mpsim_b200 -- B200-native two-qudit gate application path for matrix product states.

Drop-in for the hot path of grmlarose/mpsim (``MPS`` / ``MPSOperation`` / ``MPSimulator``,
``h`` / ``cnot`` / ``apply``, ``maxsvals``, ``wavefunction()``, ``renormalize()``), backed by
hand-written sm_100a CUDA kernels reached through the C-ABI in ``include/mpsim_b200.h``.
"""
from mpsim_b200.node import Node
from mpsim_b200.gates import (
    igate, xgate, ygate, zgate, hgate, rgate, cnot, cphase, swap,
)
from mpsim_b200.core import MPS, MPSOperation, CannotConvertToMPSOperation
from mpsim_b200.batch import MPSBatch

__all__ = [
    "MPS", "MPSOperation", "MPSBatch", "CannotConvertToMPSOperation", "Node",
    "igate", "xgate", "ygate", "zgate", "hgate", "rgate", "cnot", "cphase", "swap",
]
__version__ = "0.1.0"
