"""Golden vectors AT THE SIZES BASELINE.json NAMES (build container only; nothing here runs on the GPU box).

    python tests/golden/make_golden_baseline.py [--skip-config2]

Writes tests/golden/baseline/*.npz:

* ``config3_member0`` / ``config3_member511`` -- members 0 and 511 of BASELINE configs[3] exactly as
  ``bench.py`` runs them (40 qubits, brickwork depth 20, maxsvals = 64, gates from
  ``circuits.batch_member_gates``);
* ``config2_full`` -- BASELINE configs[2] in full (100 qubits, depth 20, chi = 256, seed 3);
* ``snake_4x4_chi96`` -- a full-rank swap-network circuit that reaches the block-Jacobi path
  (16 qubits on a snake-ordered 4 x 4 grid: 16 warm-up brickwork layers, then 4 coupler cycles
  whose vertical couplers are routed through swap networks; maxsvals = 96, thetas up to 192 x 192).

Stored per fixture: every singular value (kept and discarded) of every adjacent application, kept
counts, final bond dimensions, final norm, 256 amplitudes at seeded bitstrings (plus the full
wavefunction for the 16-qubit circuit), and a checksum of the gate tensors (the circuits are
regenerated from ``mpsim_b200.circuits`` by the tests).

Engines.  The complex128 oracle (``oracle/mps_oracle.py``) produces every fixture.  For the three
circuits that the UNMODIFIED reference (``/root/reference/mpsim`` on ``oracle/tn_shim``) finishes in
minutes -- both configs[3] members and the snake circuit -- the reference itself is run too and must
agree with the oracle (singular values 1e-10 relative to the largest, norm 1e-10, amplitudes 1e-12):
that pins the oracle at these sizes; its per-application ``_norms`` (core.py:1160-1161) are stored
as ``norms_after_each_application``.  configs[2] would take the reference ~20 minutes of norm
bookkeeping alone and is produced by the (now pinned) oracle only.
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "baseline")
sys.path.insert(0, ROOT)

from mpsim_b200 import circuits  # noqa: E402  (pure numpy host code)
from oracle.mps_oracle import OracleMPS  # noqa: E402
from tests._baseline import amplitudes_of, gates_digest  # noqa: E402

N_AMPS = 256


def run_oracle(n, ops, chi):
    mps = OracleMPS(n, dtype=np.complex128)
    t0 = time.time()
    for op in ops:
        mps.apply_two_qudit_gate(np.asarray(op.tensor), *op.indices, maxsvals=chi,
                                 keep_left_canonical=op.keep_left_canonical)
    return mps, time.time() - t0


def run_reference(n, ops, chi, gate_dtype=np.complex128):
    """The unmodified reference on the tensornetwork shim; returns its observables and sites."""
    import importlib
    mg = importlib.import_module("make_golden")            # sets up the shim + the svd logger
    del mg._svd_log[:]
    tn = mg.tn
    mps = mg.ref_core.MPS(n)
    t0 = time.time()
    for op in ops:
        kw = {"maxsvals": chi}
        if not op.keep_left_canonical:
            kw["keep_left_canonical"] = False
        # complex128 gate values: numpy promotion then makes the reference compute in complex128 (with
        # complex64 gates it stays in complex64 and its own singular values are 2e-5 off)
        mps.apply_two_qudit_gate(tn.Node(np.array(op.tensor, dtype=gate_dtype)), op.indices[0], op.indices[1], **kw)
    dt = time.time() - t0
    svals = [np.concatenate([k, r]) for k, r in mg._svd_log]
    kept = [len(k) for k, _ in mg._svd_log]
    # site tensors in (left, phys, right) order for the amplitude chain
    sites = []
    for i in range(n):
        node = mps._nodes[i]                                # the live node: its edges tell the axis roles
        edges = node.get_all_edges()
        free = node.get_all_dangling().pop()
        order = []
        if i > 0:
            order.append(edges.index(mps.get_left_connected_edge_of(i)))
        order.append(edges.index(free))
        if i < n - 1:
            order.append(edges.index(mps.get_right_connected_edge_of(i)))
        t = np.transpose(np.asarray(node.tensor), order)
        if i == 0:
            t = t[None, :, :]
        if i == n - 1:
            t = t[:, :, None]
        sites.append(t)
    return dict(svals=svals, kept=kept, norm=float(mps.norm()), norms_after=np.array(mps._norms, dtype=np.float64),
                bonds=list(mps.bond_dimensions()), sites=sites, seconds=dt)


def make(name, n, ops, chi, recipe, with_reference, sv_dtype=np.float64, full_wavefunction=False):
    ora, secs = run_oracle(n, ops, chi)
    svals = [np.concatenate([t["s_kept"], t["s_trunc"]]) for t in ora.trace]
    kept = [t["k"] for t in ora.trace]
    rng = np.random.RandomState(2024)
    bits = rng.randint(0, 2, size=(N_AMPS, n)).astype(np.uint8)
    amps = amplitudes_of(ora.sites, bits)
    min_kept_rel = min(float(t["s_kept"].min() / t["s_kept"].max()) for t in ora.trace if t["k"])
    out = {
        "nqudits": np.int64(n), "maxsvals": np.int64(chi), "recipe": np.array(recipe),
        "gates_sha256": np.array(gates_digest(ops)), "napplications": np.int64(len(ora.trace)),
        "app_index": np.array([t["index"] for t in ora.trace], dtype=np.int64),
        "app_chi": np.array([t["chi"] for t in ora.trace], dtype=np.int64),
        "app_k": np.array(kept, dtype=np.int64),
        "s_len": np.array([len(s) for s in svals], dtype=np.int64),
        "s_flat": np.concatenate(svals).astype(sv_dtype),
        "bond_dimensions": np.array(ora.bond_dimensions(), dtype=np.int64),
        "norm": np.float64(ora.norm()), "amp_bits": bits, "amp_values": amps,
        "min_kept_over_max": np.float64(min_kept_rel), "engine": np.array("oracle (complex128)"),
    }
    if full_wavefunction:
        out["wavefunction"] = ora.wavefunction().astype(np.complex64)
    msg = f"{name}: n={n} applications={len(svals)} oracle {secs:.1f}s norm={float(out['norm']):.6e} " \
          f"min kept sigma/sigma_max={min_kept_rel:.2e}"
    if with_reference:
        ref = run_reference(n, ops, chi)
        assert ref["kept"] == kept and ref["bonds"] == ora.bond_dimensions()
        worst = max(np.abs(a - b).max() / b.max() for a, b in zip(ref["svals"], svals))
        assert worst < 1e-10, worst
        assert abs(ref["norm"] - float(out["norm"])) < 1e-10 * max(1.0, float(out["norm"]))
        ramps = amplitudes_of(ref["sites"], bits)
        da = np.abs(ramps - amps).max()
        assert da < 1e-12 + 1e-9 * np.abs(amps).max(), da
        out["norms_after_each_application"] = ref["norms_after"]
        out["engine"] = np.array("oracle (complex128), cross-checked against the unmodified reference on oracle/tn_shim")
        msg += f" | reference {ref['seconds']:.1f}s agrees: sigma {worst:.1e}, amplitudes {da:.1e}"
        # FINDING: handed complex64 gates (what a complex64 user passes) the reference never leaves
        # complex64 (LAPACK cgesdd) and its own singular values drift from the complex128 result
        r32 = run_reference(n, ops, chi, gate_dtype=np.complex64)
        dev = max(np.abs(a - b).max() / b.max() for a, b in zip(r32["svals"], svals))
        out["reference_complex64_sigma_deviation"] = np.float64(dev)
        out["reference_complex64_norm"] = np.float64(r32["norm"])
        msg += f" | reference in complex64: sigma {dev:.1e}, norm {r32['norm']:.6e}"
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(msg + f" ({os.path.getsize(path)} B)", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-config2", action="store_true")
    ap.add_argument("--no-reference", action="store_true")
    args = ap.parse_args()
    sys.path.insert(0, HERE)
    ref = not args.no_reference
    for member in (0, 511):
        make(f"config3_member{member}", 40, circuits.brickwork_member(40, 20, member), 64,
             f"circuits.brickwork_member(40, 20, {member}), maxsvals=64", ref)
    nq, ops = circuits.grid_snake(4, 4, 16, 4, seed=21)
    make("snake_4x4_chi96", nq, ops, 96, "circuits.grid_snake(4, 4, 16, 4, seed=21), maxsvals=96", ref,
         full_wavefunction=True)
    if not args.skip_config2:
        make("config2_full", 100, circuits.brickwork(100, 20, seed=3), 256,
             "circuits.brickwork(100, 20, seed=3), maxsvals=256", False, sv_dtype=np.float32)
