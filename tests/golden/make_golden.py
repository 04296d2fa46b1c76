"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/mpsim)
on top of oracle/tn_shim (a stand-in for the un-vendored tensornetwork==0.2.1, validated by the
reference's own 144 core/gates tests -- oracle/run_reference_tests.sh).

Run in the BUILD CONTAINER only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each fixture stores the circuit (gate tensors, indices, kwargs) and what the reference
produced: final wavefunction (or sampled amplitudes for n > 14), norm, bond dimensions and
-- captured by wrapping ``tn.split_node_full_svd`` at ``mpsim/core.py:1132`` -- the kept and
discarded singular values of every adjacent application.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tn_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
if not hasattr(np, "complex"):
    np.complex = complex  # removed in numpy>=1.24; used at mpsim/core.py:507,561

import warnings  # noqa: E402
warnings.filterwarnings("ignore", category=DeprecationWarning)

import importlib.util  # noqa: E402
import types  # noqa: E402

import tensornetwork as tn  # noqa: E402  (the shim)

# import mpsim.core / mpsim.gates without mpsim/__init__ pulling in mpsim_cirq (needs cirq)
pkg = types.ModuleType("mpsim")
pkg.__path__ = ["/root/reference/mpsim"]
sys.modules["mpsim"] = pkg
for name in ("gates", "core"):
    spec = importlib.util.spec_from_file_location(f"mpsim.{name}", f"/root/reference/mpsim/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"mpsim.{name}"] = mod
    spec.loader.exec_module(mod)
ref_core = sys.modules["mpsim.core"]
ref_gates = sys.modules["mpsim.gates"]

# circuit generators: pure numpy host code of the product (no CUDA touched)
spec = importlib.util.spec_from_file_location("_circ_node", os.path.join(ROOT, "mpsim_b200", "node.py"))
from mpsim_b200 import circuits  # noqa: E402

_svd_log = []
_orig_svd = tn.split_node_full_svd


def _logging_svd(node, left_edges, right_edges, max_singular_values=None, max_truncation_err=None, **kw):
    out = _orig_svd(node, left_edges, right_edges, max_singular_values=max_singular_values,
                    max_truncation_err=max_truncation_err, **kw)
    u, s, vh, rest = out
    _svd_log.append((np.real(np.diag(s.tensor)).astype(np.float64).copy(),
                     np.real(np.asarray(rest)).astype(np.float64).copy()))
    return out


ref_core.tn.split_node_full_svd = _logging_svd


def run_reference(n, ops, kwargs, use_apply=False):
    """ops: list of circuits.Op.  Returns dict of reference outputs."""
    del _svd_log[:]
    mps = ref_core.MPS(n)
    for op in ops:
        node = tn.Node(np.array(op.tensor, copy=True))
        kw = dict(kwargs)
        if len(op.indices) == 1:
            mps.apply_one_qudit_gate(node, op.indices[0])
        else:
            if not op.keep_left_canonical:
                kw["keep_left_canonical"] = False
            mps.apply_two_qudit_gate(node, op.indices[0], op.indices[1], **kw)
    out = {}
    out["bond_dimensions"] = np.array(mps.bond_dimensions(), dtype=np.int64)
    out["norm"] = np.float64(mps.norm())
    out["norms_after_each_application"] = np.array(mps._norms, dtype=np.float64)
    if n <= 14:
        out["wavefunction"] = np.asarray(mps.wavefunction()).astype(np.complex128)
    else:
        wf = np.asarray(mps.wavefunction()).astype(np.complex128)
        rng = np.random.RandomState(12345)
        # the largest amplitudes plus a random sample
        top = np.argsort(-np.abs(wf))[:128]
        rnd = rng.choice(wf.size, size=128, replace=False)
        idx = np.unique(np.concatenate([top, rnd]))
        out["amp_indices"] = idx.astype(np.int64)
        out["amp_values"] = wf[idx]
    kept = [k for k, _ in _svd_log]
    rest = [r for _, r in _svd_log]
    out["s_kept_flat"] = np.concatenate(kept) if kept else np.zeros(0)
    out["s_kept_len"] = np.array([len(k) for k in kept], dtype=np.int64)
    out["s_trunc_flat"] = np.concatenate(rest) if rest else np.zeros(0)
    out["s_trunc_len"] = np.array([len(r) for r in rest], dtype=np.int64)
    return out


def save(name, n, ops, kwargs):
    out = run_reference(n, ops, kwargs)
    out["nqudits"] = np.int64(n)
    out["op_nq"] = np.array([len(op.indices) for op in ops], dtype=np.int64)
    out["op_indices"] = np.array([list(op.indices) + [-1] * (2 - len(op.indices)) for op in ops], dtype=np.int64)
    out["op_left"] = np.array([op.keep_left_canonical for op in ops], dtype=np.bool_)
    tens = np.zeros((len(ops), 16), dtype=np.complex128)
    for i, op in enumerate(ops):
        t = np.asarray(op.tensor).reshape(-1)
        tens[i, :t.size] = t
    out["op_tensors"] = tens
    out["maxsvals"] = np.int64(kwargs.get("maxsvals", -1))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} ops={len(ops)} svds={len(out['s_kept_len'])} "
          f"bonds={out['bond_dimensions'].tolist()[:8]}... norm={float(out['norm']):.6f} "
          f"({os.path.getsize(path)} B)")


Op = circuits.Op
H = ref_gates._hmatrix
X = ref_gates._xmatrix
CNOT = ref_gates._cnot_matrix
SWAP = ref_gates._swap_matrix


def with_h_layers(n, ops, every):
    """Interleave a layer of Hadamards on all sites before every ``every``-th two-qubit op."""
    out = []
    for t, op in enumerate(ops):
        if t % every == 0:
            out += [Op(H, (i,)) for i in range(n)]
        out.append(op)
    return out


if __name__ == "__main__":
    # --- known answers of the reference's own tests / README ---------------------------
    save("bell", 2, [Op(H, (0,)), Op(CNOT, (0, 1))], {})                       # core_test.py:926-929
    save("bell_maxsvals1", 2, [Op(H, (0,)), Op(CNOT, (0, 1))], {"maxsvals": 1})  # README.md:48-53
    save("bell_maxsvals0", 2, [Op(H, (0,)), Op(CNOT, (0, 1))], {"maxsvals": 0})  # core_test.py:1093-1101
    save("cnot_product_bond2", 2, [Op(CNOT, (0, 1))], {})                       # core_test.py:932-944
    save("cnot_flipped", 2, [Op(X, (1,)), Op(CNOT, (1, 0))], {})                # core_test.py:442-472
    save("x0_cnot05", 6, [Op(X, (0,)), Op(CNOT, (0, 5))], {})                   # nonlocal -> index 33
    for n in (3, 5, 9):
        save(f"ghz_n{n}", n, circuits.ghz(n), {})                                # core_test.py:1236-1245
    for n in (3, 5, 7):
        save(f"qft_n{n}", n, circuits.qft(n), {})                                # core_test.py:1248-1258
    save("ghz_qft_n8_maxsvals4", 8, circuits.ghz_qft(8), {"maxsvals": 4})       # config 2 in miniature
    # qubit hopping with swaps, core_test.py:888-902
    n = 7
    save("hopping_n7", n, [Op(X, (0,))] + [Op(SWAP, (i, i + 1)) for i in range(n - 1)]
         + [Op(SWAP, (i - 1, i)) for i in range(n - 1, 0, -1)], {})
    # --- random brickwork, the benchmark pattern (core.py:1348-1360) ---------------------
    save("brick_n6_d4_full", 6, circuits.brickwork(6, 4, seed=11), {})
    save("brick_n8_d6_chi4", 8, circuits.brickwork(8, 6, seed=12), {"maxsvals": 4})
    save("brick_n10_d8_chi8", 10, circuits.brickwork(10, 8, seed=13), {"maxsvals": 8})
    save("brick_n12_d8_chi16_h", 12, with_h_layers(12, circuits.brickwork(12, 8, seed=14), 11), {"maxsvals": 16})
    save("brick_n12_d10_chi3", 12, circuits.brickwork(12, 10, seed=15), {"maxsvals": 3})
    # nonlocal Haar gates (swap networks with truncation)
    rng = np.random.RandomState(16)
    ops = []
    for t in range(12):
        i, j = rng.choice(9, size=2, replace=False)
        ops.append(Op(circuits.haar_two_qubit(rng), (int(i), int(j)), bool(t % 2)))
    save("nonlocal_n9_chi6", 9, ops, {"maxsvals": 6})
    # BASELINE.json config 1 in full: 20 qubits, depth 10, maxsvals 64, seed 1
    save("config1_n20_d10_chi64", 20, circuits.brickwork(20, 10, seed=1), {"maxsvals": 64})
