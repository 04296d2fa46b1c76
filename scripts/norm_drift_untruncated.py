"""Norm of an UNTRUNCATED random brickwork circuit after many applications (it must stay 1: every kernel of the path
is then a unitary update): a systematic bias in theta, the isometry or the weighted factor shows up as a drift that
grows with the number of applications.   python scripts/norm_drift_untruncated.py [nqubits=12] [depth=80]"""
import os
import sys
import numpy as np
sys.path.insert(0, ".")
import mpsim_b200 as mp
from mpsim_b200 import circuits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 80
ops = circuits.brickwork(n, depth, seed=5)
mps = mp.MPS(n)
mps._execute([(o.tensor, o.indices, {"keep_left_canonical": o.keep_left_canonical}) for o in ops])
st = mps.last_status()
print("env", {k: v for k, v in os.environ.items() if k.startswith("MPSB_")}, f"n {n} depth {depth}: {len(ops)} applications, "
      f"max bond {max(mps.bond_dimensions())}, norm - 1 = {mps.norm() - 1.0:+.3e}, mean sweeps {st[:, 1].mean():.1f}, "
      f"not converged {(st[:, 0] != 0).sum()}")
