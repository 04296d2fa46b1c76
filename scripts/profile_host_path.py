"""cProfile of the gate-by-gate host path (one apply_two_qudit_gate call per gate, BASELINE configs[0]
shape: 20 qubits, depth 10, maxsvals 8): where the host time per call goes.
    python scripts/profile_host_path.py [maxsvals]
"""
import cProfile
import pstats
import sys
import time
sys.path.insert(0, ".")
import torch
import mpsim_b200 as mp
from mpsim_b200 import circuits

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, depth = 20, 10
ops = circuits.brickwork(n, depth, seed=1)
nodes = [mp.Node(o.tensor) for o in ops]


def run():
    mps = mp.MPS(n)
    for o, g in zip(ops, nodes):
        mps.apply_two_qudit_gate(g, *o.indices, keep_left_canonical=o.keep_left_canonical, maxsvals=chi)
    return mps


for _ in range(3):
    run()
torch.cuda.synchronize()
t0 = time.perf_counter()
m = run()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{len(ops)} calls: host {1e6 * (t1 - t0) / len(ops):.1f} us/call, drained {1e6 * (t2 - t0) / len(ops):.1f} us/call")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    run()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(40)
