"""Times BASELINE.json configs[2] (100-qubit brickwork depth 20, chi=256) on one GPU."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import mpsim_b200 as mp
from mpsim_b200 import circuits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
chi = int(sys.argv[3]) if len(sys.argv) > 3 else 256
ops = circuits.brickwork(n, depth, seed=3)
triples = [(op.tensor, op.indices, {"maxsvals": chi, "keep_left_canonical": op.keep_left_canonical}) for op in ops]
mps = mp.MPS(n)
from mpsim_b200.planner import plan_operations
plan = plan_operations(n, 2, mps._chain.bonds, triples)
cp = mps._chain.compile(plan)
torch.cuda.synchronize()
for rep in range(2):
    mps._chain.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mps._chain.run(cp)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = cp.info.cpu().numpy()[: len(plan.apps2)]
    print(f"rep {rep}: {len(plan.apps2)} applications in {dt:.3f} s = {len(plan.apps2)/dt:.1f} apps/s; launches {len(cp.launches)}; "
          f"not converged {int((info[:,0]!=0).sum())}; sweeps mean {info[:,1].mean():.1f} max {info[:,1].max()}; norm {mps.norm():.6f}")
