"""Times BASELINE.json configs[4]: the 53-qubit Sycamore-style circuit (6 x 9 grid minus one site,
snake order, coupler pattern ABCDCDAB, 14 cycles, Haar two-qubit gates, swap networks expanded:
317 logical gates -> 3 069 adjacent applications), chi = 1024, one chain, through the moment
dispatcher.  The run is TIME-BOXED: launches are issued one circuit layer at a time until
``budget_s`` seconds of device time have been spent; the line reports how far it got, the rate on
the full-size shape measured in the circuit itself, and the full-circuit time extrapolated from it.
python scripts/time_config5.py [budget_s] [chi] [cycles]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch
import mpsim_b200 as mp
from mpsim_b200 import circuits
from mpsim_b200.planner import plan_operations

budget_s = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 14

nq, ops = circuits.sycamore_snake(cycles=cycles)
triples = [(op.tensor, op.indices, {"maxsvals": chi, "keep_left_canonical": op.keep_left_canonical}) for op in ops]
mps = mp.MPS(nq)
chain = mps._chain
t0 = time.perf_counter()
plan = plan_operations(nq, 2, chain.bonds, triples)
cp = chain.compile(plan)
print(f"{len(ops)} logical gates -> {len(plan.apps2)} adjacent applications in {len(cp.launches)} calls; "
      f"plan+compile {time.perf_counter() - t0:.2f} s; slab {chain.slab.numel() * 8 / 1e6:.0f} MB", flush=True)

launches = list(cp.launches)


def napps(L):
    if L[0] == "g2":
        return int(L[2])
    if L[0] == "g2layer":
        return int(L[1]["ndesc"].sum())
    return 0


def full_shape(L):
    """number of applications of the full (chi, chi, chi, chi) shape in this call"""
    if L[0] == "g2":
        return int(L[2]) if (L[3], L[4], L[5], L[6]) == (chi, chi, chi, chi) else 0
    if L[0] == "g2layer":
        t = L[1]
        m = (t["chiL"] == chi) & (t["chiM"] == chi) & (t["chiR"] == chi) & (t["k"] == chi)
        return int(t["ndesc"][m].sum())
    return 0


done_apps = done_calls = 0
full_apps, full_ms = 0, 0.0          # calls made of full-shape applications only
spent_ms = 0.0
chain.upload_gates(cp)
torch.cuda.synchronize()
wall0 = time.perf_counter()
for L in launches:
    cp.launches = [L]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chain.run(cp, upload=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    spent_ms += ms
    n = napps(L)
    done_apps += n
    done_calls += 1
    if n and full_shape(L) == n:
        full_apps += n
        full_ms += ms
    if time.perf_counter() - wall0 > budget_s:
        break
cp.launches = launches
complete = done_calls == len(launches)

info = cp.info.cpu().numpy()[:done_apps]          # descriptors are laid out in launch order
total_apps = len(plan.apps2)
n_full_total = sum(1 for a in plan.apps2 if (a.chiL, a.chiM, a.chiR, a.k) == (chi, chi, chi, chi))
out = {
    "workload": f"{nq}-qubit Sycamore-style snake circuit, {cycles} cycles, chi={chi} (BASELINE.json configs[4]), one chain",
    "applications_total": total_apps, "applications_done": done_apps, "complete": complete,
    "device_s": spent_ms * 1e-3, "wall_s": time.perf_counter() - wall0,
    "applications_per_sec_so_far": done_apps / (spent_ms * 1e-3) if spent_ms else None,
    "svd_not_converged": int((info[:, 0] != 0).sum()) if done_apps else 0,
    "svd_mean_sweeps": float(info[:, 1].mean()) if done_apps else None,
    "svd_max_sweeps": int(info[:, 1].max()) if done_apps else None,
    "full_shape": {"count_in_circuit": n_full_total, "timed": full_apps,
                   "ms_per_application": full_ms / full_apps if full_apps else None},
}
if done_apps:
    bad = {}
    for pos in np.nonzero(info[:, 0] != 0)[0]:
        a = plan.apps2[cp.order2[int(pos)]]
        key = f"{a.chiL},{a.chiM},{a.chiR},{a.k}"
        bad[key] = bad.get(key, 0) + 1
    out["not_converged_shapes"] = bad
if full_apps and not complete:
    rest_full = n_full_total - sum(full_shape(L) for L in launches[:done_calls])
    rest_other = (total_apps - done_apps) - rest_full
    other_ms = (spent_ms - full_ms) / max(done_apps - full_apps, 1)
    out["extrapolated_full_circuit_s"] = (spent_ms + rest_full * full_ms / full_apps + rest_other * other_ms) * 1e-3
if complete:
    out["norm"] = float(mps.norm())
print(json.dumps(out), flush=True)
