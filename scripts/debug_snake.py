"""Teacher-forced singular-value errors of the swap-network fixture, per application (which shapes are worst?)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import mpsim_b200 as mp
from tests import _baseline
from oracle.mps_oracle import OracleMPS

base = _baseline.Baseline("snake_4x4_chi96")
ora = OracleMPS(base.n, dtype=np.complex128)
mps = mp.MPS(base.n)
mps.record_singular_values(True)
rows = []
for op in base.ops:
    lo, hi = min(op.indices), max(op.indices)
    for s in range(lo, hi + 1):
        mps._chain.set_site(s, ora.sites[s])
    n0 = len(ora.trace)
    ora.apply_two_qudit_gate(op.tensor, *op.indices, maxsvals=base.chi, keep_left_canonical=op.keep_left_canonical)
    mps.apply_two_qudit_gate(mp.Node(op.tensor), *op.indices, maxsvals=base.chi, keep_left_canonical=op.keep_left_canonical)
    st = mps.last_status()
    for s, t, inf in zip(mps.last_singular_values(), ora.trace[n0:], st):
        ref = np.concatenate([t["s_kept"], t["s_trunc"]])
        err = np.abs(s["svals"] - ref)
        j = int(err.argmax())
        rows.append((err.max() / ref.max(), t["chi"], t["k"], s["is_swap"], j, ref[j] / ref.max(), int(inf[1]),
                     float(np.min(np.abs(np.diff(ref[:max(2, t["k"])]))) / ref.max()) if ref.size > 1 else 0.0))
rows.sort(key=lambda r: -r[0])
print("err/smax   chi(l,m,r)        k  swap  idx  sigma/smax  sweeps  min gap among kept")
for r in rows[:15]:
    print("%.2e  %-16s %4d  %-5s %4d  %.3e  %3d  %.2e" % r)
small = [r[0] for r in rows if 2 * max(r[1][0], r[1][2]) <= 128]
large = [r[0] for r in rows if 2 * max(r[1][0], r[1][2]) > 128]
print("single-CTA path: %d applications, worst %.2e;  block-Jacobi path: %d applications, worst %.2e"
      % (len(small), max(small) if small else 0, len(large), max(large) if large else 0))
