import os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
import mpsim_b200 as mp
from mpsim_b200 import circuits
n, depth = 12, 250
ops = circuits.brickwork(n, depth, seed=5)
mps = mp.MPS(n)
per_layer = []
t = 0
for layer in range(depth):
    cnt = (n // 2) if layer % 2 == 0 else (n // 2 - 1)
    chunk = ops[t:t + cnt]; t += cnt
    mps._execute([(o.tensor, o.indices, {"keep_left_canonical": o.keep_left_canonical}) for o in chunk])
    if layer % 10 == 9 or layer > 225:
        mx = [float(mps._chain.site_view(s).abs().max()) for s in range(n)]
        print(layer, "norm %.6e" % mps.norm(), "site max:", " ".join("%.1e" % m for m in mx))
