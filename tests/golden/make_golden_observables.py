"""Generates tests/golden/observables.npz: what the UNMODIFIED reference returns for the
SURVEY.md 8(f) rows 2-3 -- ``MPS.from_wavefunction`` (mpsim/core.py:245-328),
``reduced_density_matrix`` (596-652), ``sample`` (654-721), ``expectation`` (723-751) and
``inner_product`` between different states (507-561) -- on top of oracle/tn_shim.

Run in the BUILD CONTAINER only (needs /root/reference):

    python tests/golden/make_golden_observables.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (loads the reference on the shim; saves nothing on import)

tn, ref_core, ref_gates, circuits = mg.tn, mg.ref_core, mg.ref_gates, mg.circuits
Op = circuits.Op

PAULI_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
PAULI_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
PROJ0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)          # Hermitian, NOT unitary
ZX_MIX = 0.6 * PAULI_Z + 0.3 * np.array([[0, 1], [1, 0]], dtype=np.complex128)   # Hermitian, not unitary
ZZ = np.kron(PAULI_Z, PAULI_Z).reshape(2, 2, 2, 2)
XY = np.kron(np.array([[0, 1], [1, 0]]), PAULI_Y).reshape(2, 2, 2, 2)


def build_reference_state(n, ops, kwargs):
    mps = ref_core.MPS(n)
    for op in ops:
        node = tn.Node(np.array(op.tensor, copy=True))
        if len(op.indices) == 1:
            mps.apply_one_qudit_gate(node, op.indices[0])
        else:
            kw = dict(kwargs)
            if not op.keep_left_canonical:
                kw["keep_left_canonical"] = False
            mps.apply_two_qudit_gate(node, op.indices[0], op.indices[1], **kw)
    return mps


def ops_arrays(prefix, n, ops, kwargs, out):
    out[prefix + "nqudits"] = np.int64(n)
    out[prefix + "op_indices"] = np.array([list(op.indices) + [-1] * (2 - len(op.indices)) for op in ops], dtype=np.int64)
    out[prefix + "op_left"] = np.array([op.keep_left_canonical for op in ops], dtype=np.bool_)
    tens = np.zeros((len(ops), 16), dtype=np.complex128)
    for i, op in enumerate(ops):
        t = np.asarray(op.tensor).reshape(-1)
        tens[i, :t.size] = t
    out[prefix + "op_tensors"] = tens
    out[prefix + "maxsvals"] = np.int64(kwargs.get("maxsvals", -1))


STATES = {
    # name: (n, ops, kwargs)
    "brick6": (6, [Op(ref_gates._hmatrix, (i,)) for i in range(6)] + circuits.brickwork(6, 4, seed=21), {}),
    "brick8chi4": (8, circuits.brickwork(8, 6, seed=22), {"maxsvals": 4}),            # truncated: norm < 1
    "ghz5": (5, circuits.ghz(5), {}),
}
RDM_SETS = {
    "brick6": [(0,), (3,), (5,), (1, 2), (2, 1), (0, 5), (4, 0, 2), (1, 3, 4, 5)],
    "brick8chi4": [(0,), (7,), (3, 4), (6, 1), (2, 5, 0)],
    "ghz5": [(0,), (2,), (0, 4), (4, 2, 0)],
}
OBSERVABLES = [
    # (tensor, indices)
    (np.asarray(ref_gates._hmatrix, dtype=np.complex128), (0,)),
    (np.asarray(ref_gates._xmatrix, dtype=np.complex128), (2,)),
    (PAULI_Z, (1,)),
    (PAULI_Y, (4,)),
    (PROJ0, (3,)),                      # non-unitary one-qudit path (core.py:816-845)
    (ZX_MIX, (0,)),
    (ZZ, (1, 2)),
    (ZZ, (0, 4)),                       # non-adjacent: swap network (core.py:1035-1043)
    (XY, (3, 2)),                       # flipped indices (core.py:1031-1033)
    (np.asarray(ref_gates._cnot_matrix, dtype=np.complex128).reshape(2, 2, 2, 2), (2, 3)),
]


def main():
    out = {}
    for name, (n, ops, kwargs) in STATES.items():
        pre = f"{name}__"
        ops_arrays(pre, n, ops, kwargs, out)
        mps = build_reference_state(n, ops, kwargs)
        out[pre + "wavefunction"] = np.asarray(mps.wavefunction()).astype(np.complex128)
        out[pre + "norm"] = np.float64(mps.norm())
        for t, idx in enumerate(RDM_SETS[name]):
            arg = idx[0] if (len(idx) == 1 and t % 2 == 0) else list(idx)      # int and sequence forms
            out[pre + f"rdm{t}_indices"] = np.array(idx, dtype=np.int64)
            out[pre + f"rdm{t}"] = np.asarray(mps.reduced_density_matrix(node_indices=arg)).astype(np.complex128)
        out[pre + "n_rdm"] = np.int64(len(RDM_SETS[name]))
        vals = []
        for tensor, idx in OBSERVABLES:
            if max(idx) >= n:
                vals.append(np.nan)
                continue
            obs = ref_core.MPSOperation(tn.Node(np.array(tensor, copy=True)), idx)
            vals.append(float(mps.expectation(obs)))
        out[pre + "expectations"] = np.array(vals, dtype=np.float64)
        # the state must be untouched by all of the above (core_test.py:1552-1559)
        assert np.allclose(mps.wavefunction(), out[pre + "wavefunction"])
        np.random.seed(1234)
        try:
            out[pre + "samples_seed1234"] = np.array(mps.sample(nsamples=24), dtype=np.int64)
        except ValueError as e:
            # np.random.choice refuses marginals of a truncated (unnormalised) state (core.py:667)
            assert "do not sum to 1" in str(e)
            out[pre + "sample_raises"] = np.bool_(True)
        print(name, "norm", out[pre + "norm"], "expectations", np.round(vals, 5).tolist())
    n_obs = len(OBSERVABLES)
    obs_t = np.zeros((n_obs, 16), dtype=np.complex128)
    obs_i = -np.ones((n_obs, 2), dtype=np.int64)
    for t, (tensor, idx) in enumerate(OBSERVABLES):
        obs_t[t, :np.asarray(tensor).size] = np.asarray(tensor).reshape(-1)
        obs_i[t, :len(idx)] = idx
    out["obs_tensors"], out["obs_indices"] = obs_t, obs_i
    # inner products between DIFFERENT states (core.py:543-561 conjugates ``other``)
    a = build_reference_state(*STATES["brick6"])
    b = build_reference_state(6, circuits.brickwork(6, 3, seed=23), {})
    ops_arrays("brick6b__", 6, circuits.brickwork(6, 3, seed=23), {}, out)
    out["inner_brick6_brick6b"] = np.complex128(a.inner_product(b))
    out["inner_brick6b_brick6"] = np.complex128(b.inner_product(a))
    # from_wavefunction: random real/complex qubit vectors and one qutrit vector
    rng = np.random.RandomState(77)
    cases = [(2, 2), (3, 2), (5, 2), (7, 2), (8, 2), (3, 3), (4, 3), (2, 5)]
    for t, (n, d) in enumerate(cases):
        wf = rng.randn(d ** n) + 1j * rng.randn(d ** n)
        if t % 3 == 0:
            wf = np.abs(wf.real)
        wf = wf / np.linalg.norm(wf)
        mps = ref_core.MPS.from_wavefunction(wf, nqudits=n, qudit_dimension=d)
        out[f"fw{t}_input"] = wf.astype(np.complex128)
        out[f"fw{t}_nd"] = np.array([n, d], dtype=np.int64)
        out[f"fw{t}_bonds"] = np.array(mps.bond_dimensions(), dtype=np.int64)
        out[f"fw{t}_wavefunction"] = np.asarray(mps.wavefunction()).astype(np.complex128)
        out[f"fw{t}_norm"] = np.float64(mps.norm())
        # gauge check material: the left-most node as the reference built it (U sqrt(S))
        out[f"fw{t}_node0"] = np.asarray(mps.get_node(0).tensor).astype(np.complex128)
        site = int(rng.randint(n))
        out[f"fw{t}_rdm_site"] = np.int64(site)
        out[f"fw{t}_rdm"] = np.asarray(mps.reduced_density_matrix(node_indices=site)).astype(np.complex128)
    out["n_fw"] = np.int64(len(cases))
    path = os.path.join(HERE, "observables.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "B")


if __name__ == "__main__":
    main()
