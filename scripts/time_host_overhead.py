"""Host cost of the gate-by-gate API ('Host side of a plan'): BASELINE configs[0]
(20 qubits, depth 10, maxsvals = 64) applied one ``apply_two_qudit_gate`` call at a time, as the reference's own
tests and examples use it.  Prints wall time per call with the device drained after every call (host + device,
serialised), without draining (host and device overlap), and the device time alone (CUDA events), so that
wall(no drain) ~ max(host, device).
    python scripts/time_host_overhead.py [maxsvals=64] [repetitions=3]
"""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import torch
import mpsim_b200 as mp
from mpsim_b200 import circuits

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, depth = 20, 10
ops = circuits.brickwork(n, depth, seed=1)
nodes = [mp.Node(op.tensor) for op in ops]


def run(drain):
    mps = mp.MPS(n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for op, node in zip(ops, nodes):
        mps.apply_two_qudit_gate(node, *op.indices, maxsvals=chi, keep_left_canonical=op.keep_left_canonical)
        if drain:
            torch.cuda.synchronize()
    t_host = time.perf_counter() - t0          # time until the last call RETURNED
    e1.record()
    torch.cuda.synchronize()
    return t_host, time.perf_counter() - t0, e0.elapsed_time(e1) * 1e-3, mps.norm()


run(False)
for r in range(reps):
    th, tw, td, nrm = run(False)
    th2, tw2, td2, _ = run(True)
    print(f"{len(ops)} calls, maxsvals {chi}: calls returned after {1e6 * th / len(ops):7.1f} us/call (host), wall "
          f"{1e6 * tw / len(ops):7.1f} us/call, device span {1e6 * td / len(ops):7.1f} us/call;  drained after every "
          f"call: {1e6 * tw2 / len(ops):7.1f} us/call;  norm {nrm:.6f}")
