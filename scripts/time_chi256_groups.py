"""Per shape-group time of BASELINE.json configs[2] (100 qubits, depth 20, chi=256): every
mpsb_apply_gate2 call of the compiled plan is timed with a device synchronise on both sides."""
import sys, time, collections
sys.path.insert(0, ".")
import torch
import mpsim_b200 as mp
from mpsim_b200 import circuits, _lib
from mpsim_b200.planner import plan_operations
n, depth, chi = 100, 20, 256
ops = circuits.brickwork(n, depth, seed=3)
triples = [(op.tensor, op.indices, {"maxsvals": chi, "keep_left_canonical": op.keep_left_canonical}) for op in ops]
mps = mp.MPS(n)
chain = mps._chain
plan = plan_operations(n, 2, chain.bonds, triples)
cp = chain.compile(plan)
lib = _lib.load(require_device=True)
for rep in range(2):
    chain.reset()
    chain.upload_gates(cp)
    ws = chain.workspace(cp.workspace_bytes)
    st = _lib.stream_ptr()
    tot = collections.OrderedDict()
    torch.cuda.synchronize()
    flat = []
    for L in cp.launches:          # layer calls expanded into their groups (timed one after the other)
        if L[0] == "g2layer":
            base = cp.desc2.data_ptr()
            flat += [("g2", int((g["descs"] - base) // _lib.GATE2_DESC.itemsize), int(g["ndesc"]), int(g["chiL"]),
                      int(g["chiM"]), int(g["chiR"]), int(g["k"]), int(g["left_canonical"])) for g in L[1]]
        else:
            flat.append(L)
    for L in flat:
        _, off, cnt, chiL, chiM, chiR, k, lc = L
        t0 = time.perf_counter()
        _lib.check(lib.mpsb_apply_gate2(cp.desc2.data_ptr() + off * _lib.GATE2_DESC.itemsize, cnt, 1, 2, chiL, chiM, chiR,
                                        k, lc, ws.data_ptr(), ws.numel(), cp.info.data_ptr() + off * 8, st))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        key = (chiL, chiM, chiR, k)
        a = tot.setdefault(key, [0, 0, 0.0])
        a[0] += 1; a[1] += cnt; a[2] += dt
    chain.bonds = list(cp.bonds_out)
total = sum(v[2] for v in tot.values())
print(f"total {total*1e3:.1f} ms over {len(cp.launches)} calls")
for key, (calls, jobs, t) in sorted(tot.items(), key=lambda kv: -kv[1][2]):
    print(f"  chi(L,M,R,k)={key}: calls {calls:3d} jobs {jobs:4d}  {t*1e3:8.1f} ms  ({100*t/total:4.1f} %)  {t*1e3/jobs:7.2f} ms/job")
