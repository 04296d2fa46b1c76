import sys, os
sys.path.insert(0, ".")
import numpy as np
import mpsim_b200 as mp
from mpsim_b200.mpsim_cirq import MPSimulator
from tests._fake_cirq import Circuit, H, CNOT, CZPow
n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
ms = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ops = [H(0)] + [CNOT(0, i) for i in range(1, n)]
for i in range(n - 1, -1, -1):
    ops.append(H(i))
    for j in range(i - 1, -1, -1):
        ops.append(CZPow(2.0 ** (j - i), j, i))
mps = MPSimulator({"maxsvals": ms}).simulate(Circuit(ops))
print("env", {k: v for k, v in os.environ.items() if k.startswith("MPSB_")}, "n", n, "maxsvals", ms, "norm - 1 = %.3e" % (mps.norm() - 1.0), "max bond", max(mps.bond_dimensions()))
