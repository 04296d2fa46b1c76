"""Loader for tests/golden/*.npz (made by tests/golden/make_golden.py from the real reference)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    """Circuit fixtures (observables.npz has its own loader, tests/_observables.py)."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0]
                              for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))) if n != "observables")


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.n = int(z["nqudits"])
        self.maxsvals = int(z["maxsvals"])
        self.kwargs = {} if self.maxsvals < 0 else {"maxsvals": self.maxsvals}
        self.ops = []
        for nq, idx, left, t in zip(z["op_nq"], z["op_indices"], z["op_left"], z["op_tensors"]):
            nq = int(nq)
            tensor = t[: 4 ** nq].reshape([2] * (2 * nq))
            self.ops.append((tensor, tuple(int(i) for i in idx[:nq]), bool(left)))
        self.bond_dimensions = z["bond_dimensions"].tolist()
        self.norm = float(z["norm"])
        self.norms_after = z["norms_after_each_application"]
        self.wavefunction = z["wavefunction"] if "wavefunction" in z.files else None
        self.amp_indices = z["amp_indices"] if "amp_indices" in z.files else None
        self.amp_values = z["amp_values"] if "amp_values" in z.files else None
        ends = np.cumsum(z["s_kept_len"])
        self.s_kept = np.split(z["s_kept_flat"], ends[:-1]) if len(ends) else []
        ends = np.cumsum(z["s_trunc_len"])
        self.s_trunc = np.split(z["s_trunc_flat"], ends[:-1]) if len(ends) else []


def run_ops(mps, golden, two_qudit="apply_two_qudit_gate", one_qudit="apply_one_qudit_gate", wrap=None):
    """Drive any MPS-like object through a golden circuit."""
    for tensor, idx, left in golden.ops:
        g = wrap(tensor) if wrap is not None else tensor
        if len(idx) == 1:
            getattr(mps, one_qudit)(g, idx[0])
        else:
            kw = dict(golden.kwargs)
            if not left:
                kw["keep_left_canonical"] = False
            getattr(mps, two_qudit)(g, idx[0], idx[1], **kw)
    return mps


def first_rank_deficient(golden, rel=1e-9):
    """Index of the first application whose kept singular values include a numerical zero
    (None if every theta is full rank).  After it the reference is gauge-unstable."""
    for t, s in enumerate(golden.s_kept):
        if s.size and s.min() <= rel * max(s.max(), 1e-300):
            return t
    return None
