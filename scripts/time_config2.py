"""Times BASELINE.json configs[1]: 50-qubit GHZ + QFT chain, maxsvals=128, one chain, through the
moment dispatcher (swap networks expanded: 1 274 logical two-qubit gates -> 42 826 adjacent
applications).  python scripts/time_config2.py [n] [maxsvals] [max_logical_ops]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import mpsim_b200 as mp
from mpsim_b200 import circuits
from mpsim_b200.planner import plan_operations
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 128
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ops = circuits.ghz_qft(n)
if limit:
    ops = ops[:limit]
triples = [(op.tensor, op.indices, {"maxsvals": chi, "keep_left_canonical": op.keep_left_canonical}) for op in ops]
mps = mp.MPS(n)
t0 = time.perf_counter()
plan = plan_operations(n, 2, mps._chain.bonds, triples)
cp = mps._chain.compile(plan)
t1 = time.perf_counter()
print(f"{len(ops)} operations -> {len(plan.apps2)} adjacent applications, {len(cp.launches)} calls; plan+compile {t1-t0:.2f} s", flush=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
mps._chain.run(cp)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
info = cp.info.cpu().numpy()[: len(plan.apps2)]
shapes = {}
for a in plan.apps2:
    shapes[(a.chiL, a.chiM, a.chiR, a.k)] = shapes.get((a.chiL, a.chiM, a.chiR, a.k), 0) + 1
top = sorted(shapes.items(), key=lambda kv: -kv[1])[:3]
print(f"{len(plan.apps2)} applications in {dt:.2f} s = {len(plan.apps2)/dt:.1f} apps/s; not converged {int((info[:,0]!=0).sum())}; "
      f"sweeps mean {info[:,1].mean():.1f} max {info[:,1].max()}; norm {mps.norm():.6f}; top shapes {top}")
