"""Loader for tests/golden/baseline/*.npz (made by tests/golden/make_golden_baseline.py): golden
vectors at the sizes BASELINE.json names.  The circuits are regenerated from mpsim_b200.circuits
(pure numpy) and checked against the stored SHA-256 of the gate tensors."""
import hashlib
import os

import numpy as np

from mpsim_b200 import circuits

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "baseline")

RECIPES = {
    "config3_member0": lambda: (40, circuits.brickwork_member(40, 20, 0), 64),
    "config3_member511": lambda: (40, circuits.brickwork_member(40, 20, 511), 64),
    "snake_4x4_chi96": lambda: circuits.grid_snake(4, 4, 16, 4, seed=21) + (96,),
    "config2_full": lambda: (100, circuits.brickwork(100, 20, seed=3), 256),
}


def available(name):
    return os.path.exists(os.path.join(DIR, name + ".npz"))


def gates_digest(ops):
    h = hashlib.sha256()
    for op in ops:
        h.update(np.ascontiguousarray(np.asarray(op.tensor, dtype=np.complex128)).tobytes())
        h.update(np.asarray(op.indices, dtype=np.int64).tobytes())
        h.update(b"L" if op.keep_left_canonical else b"R")
    return h.hexdigest()


def amplitudes_of(sites, bits):
    """<bits|psi> for every row of ``bits`` by a complex128 vector chain over [chiL][d][chiR] sites."""
    v = np.ones((bits.shape[0], 1), dtype=np.complex128)
    for i, a in enumerate(sites):
        sel = np.asarray(a).astype(np.complex128)[:, bits[:, i], :]          # [chiL, K, chiR]
        v = np.einsum("kl,lkr->kr", v, sel)
    return v[:, 0]


class Baseline:
    def __init__(self, name):
        z = np.load(os.path.join(DIR, name + ".npz"))
        self.name = name
        self.n, self.ops, self.chi = RECIPES[name]()
        assert int(z["nqudits"]) == self.n and int(z["maxsvals"]) == self.chi
        assert gates_digest(self.ops) == str(z["gates_sha256"]), "circuit generator drifted from the fixture"
        ends = np.cumsum(z["s_len"])
        self.svals = [s.astype(np.float64) for s in np.split(z["s_flat"], ends[:-1])]
        self.k = z["app_k"].tolist()
        self.app_index = z["app_index"].tolist()
        self.app_chi = [tuple(int(v) for v in c) for c in z["app_chi"]]
        self.bond_dimensions = z["bond_dimensions"].tolist()
        self.norm = float(z["norm"])
        self.amp_bits = z["amp_bits"]
        self.amp_values = z["amp_values"]
        self.wavefunction = z["wavefunction"] if "wavefunction" in z.files else None
        self.norms_after = z["norms_after_each_application"] if "norms_after_each_application" in z.files else None
        self.min_kept_over_max = float(z["min_kept_over_max"])
        self.reference_complex64_sigma_deviation = (float(z["reference_complex64_sigma_deviation"])
                                                    if "reference_complex64_sigma_deviation" in z.files else None)


def sigma_errors(svals, k, ref, floor=1e-3):
    """(error relative to the largest singular value, worst per-value relative error over the
    KEPT values >= floor * sigma_max) of one application."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(svals, dtype=np.float64)[: ref.size]
    smax = max(ref.max(), 1e-300)
    e_max = np.abs(got - ref).max() / smax
    sel = ref[:k] >= floor * smax
    e_rel = (np.abs(got[:k] - ref[:k])[sel] / ref[:k][sel]).max() if sel.any() else 0.0
    return float(e_max), float(e_rel)


def sampled_fidelity(a, b):
    """|<a|b>|^2 / (<a|a><b|b>) over the sampled amplitudes."""
    a = np.asarray(a, dtype=np.complex128); b = np.asarray(b, dtype=np.complex128)
    return float(abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real))
