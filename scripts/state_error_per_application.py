"""Per-application backward error of the truncated split on real thetas (teacher-forced):
    || A_i' A_j'  -  U_k S_k V_k^H || / || theta ||
where theta comes from the complex128 oracle's sites, A_i' A_j' is what the GPU path wrote for that
application (= Q Q^H theta_gpu) and U_k S_k V_k^H the oracle's rank-k truncation.  Also the same quantity for
numpy's complex64 SVD (LAPACK cgesdd) on the same theta, for scale.
    python scripts/state_error_per_application.py <fixture>        (tests/golden/baseline/<fixture>.npz)
"""
import sys
import numpy as np
sys.path.insert(0, ".")
import mpsim_b200 as mp
from tests import _baseline
from oracle.mps_oracle import OracleMPS

name = sys.argv[1] if len(sys.argv) > 1 else "config3_member0"
base = _baseline.Baseline(name)
ora = OracleMPS(base.n, dtype=np.complex128)
mps = mp.MPS(base.n)
rows = []
for op in base.ops:
    i, j = op.indices
    if abs(i - j) != 1:                       # routed gates: let both sides run, compare nothing
        ora.apply_two_qudit_gate(op.tensor, i, j, maxsvals=base.chi, keep_left_canonical=op.keep_left_canonical)
        continue
    lo = min(i, j)
    for s in (lo, lo + 1):
        mps._chain.set_site(s, ora.sites[s])
    a, b = ora.sites[lo], ora.sites[lo + 1]
    g = op.tensor if i < j else np.transpose(op.tensor, (1, 0, 3, 2))
    theta = np.einsum("abpq,lpm,mqr->labr", g, a, b).reshape(a.shape[0] * 2, 2 * b.shape[2])
    ora.apply_two_qudit_gate(op.tensor, i, j, maxsvals=base.chi, keep_left_canonical=op.keep_left_canonical)
    mps.apply_two_qudit_gate(mp.Node(op.tensor), i, j, maxsvals=base.chi, keep_left_canonical=op.keep_left_canonical)
    k = ora.trace[-1]["k"]
    if k >= min(theta.shape):                  # nothing truncated: the projection is the identity
        continue
    ref = np.einsum("lpm,mqr->lpqr", ora.sites[lo], ora.sites[lo + 1]).reshape(theta.shape)
    al = mps._chain.site_view(lo).cpu().numpy().astype(np.complex128)
    ar = mps._chain.site_view(lo + 1).cpu().numpy().astype(np.complex128)
    got = np.einsum("lpm,mqr->lpqr", al, ar).reshape(theta.shape)
    nt = np.linalg.norm(theta)
    u4, s4, v4 = np.linalg.svd(theta.astype(np.complex64), full_matrices=False)
    lap = (u4[:, :k].astype(np.complex128) * s4[:k]) @ v4[:k].astype(np.complex128)
    s = np.concatenate([ora.trace[-1]["s_kept"], ora.trace[-1]["s_trunc"]])
    rows.append((np.linalg.norm(got - ref) / nt, np.linalg.norm(lap - ref) / nt, theta.shape, k,
                 (s[k - 1] - s[k]) / s[0] if k < s.size else 1.0, 2 * max(a.shape[0], b.shape[2]) > 128))
for large in (False, True):
    sel = [r for r in rows if r[5] == large]
    if not sel:
        continue
    e = np.array([r[0] for r in sel]); l = np.array([r[1] for r in sel])
    print(f"{name}: {'block-Jacobi' if large else 'single-CTA'} path, {len(sel)} truncated applications: backward error "
          f"median {np.median(e):.2e}, max {e.max():.2e};  LAPACK complex64 on the same thetas: median {np.median(l):.2e}, "
          f"max {l.max():.2e}")
    worst = sorted(sel, key=lambda r: -r[0])[:3]
    for r in worst:
        print(f"    worst: {r[0]:.2e} (LAPACK c64 {r[1]:.2e}) shape {r[2]} k {r[3]} gap at the cut {r[4]:.1e} sigma_max")
