"""Shared checks against tests/golden/observables.npz (made by
tests/golden/make_golden_observables.py from the UNMODIFIED reference): the same assertions run
on the CPU oracle (tests/test_oracle_observables.py) and, on the GPU box, on mpsim_b200 through
the C-ABI (tests/test_gpu_observables.py).  An *adapter* hides the two constructors:

    adapter.new(n, d)                         -> MPS-like in |0...0>
    adapter.from_wavefunction(wf, n, d)       -> MPS-like
    adapter.apply1(mps, tensor, i) / adapter.apply2(mps, tensor, i, j, **kw)
    adapter.expectation(mps, tensor, indices) -> float
"""
import os

import numpy as np

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "observables.npz"))
STATES = ("brick6", "brick8chi4", "ghz5")


def build_state(adapter, name):
    pre = name + "__"
    n = int(Z[pre + "nqudits"])
    mps = adapter.new(n, 2)
    maxsvals = int(Z[pre + "maxsvals"])
    for idx, left, t in zip(Z[pre + "op_indices"], Z[pre + "op_left"], Z[pre + "op_tensors"]):
        if idx[1] < 0:
            adapter.apply1(mps, t[:4].reshape(2, 2), int(idx[0]))
        else:
            kw = {} if maxsvals < 0 else {"maxsvals": maxsvals}
            if not left:
                kw["keep_left_canonical"] = False
            adapter.apply2(mps, t.reshape(2, 2, 2, 2), int(idx[0]), int(idx[1]), **kw)
    return mps


def check_state(adapter, name, atol):
    pre = name + "__"
    mps = build_state(adapter, name)
    wf0 = np.asarray(mps.wavefunction())
    np.testing.assert_allclose(wf0, Z[pre + "wavefunction"], rtol=0, atol=atol)
    assert abs(mps.norm() - float(Z[pre + "norm"])) < atol
    # reduced density matrices, int and sequence arguments, any index order (core.py:596-652)
    for t in range(int(Z[pre + "n_rdm"])):
        idx = [int(i) for i in Z[pre + f"rdm{t}_indices"]]
        arg = idx[0] if (len(idx) == 1 and t % 2 == 0) else idx
        rdm = np.asarray(mps.reduced_density_matrix(node_indices=arg))
        ref = Z[pre + f"rdm{t}"]
        assert rdm.shape == ref.shape
        np.testing.assert_allclose(rdm, ref, rtol=0, atol=atol)
    # expectation values incl. non-unitary, non-adjacent and flipped observables (core.py:723-751)
    n = int(Z[pre + "nqudits"])
    for tensor, idx, ref in zip(Z["obs_tensors"], Z["obs_indices"], Z[pre + "expectations"]):
        k = 1 if idx[1] < 0 else 2
        if max(idx) >= n:
            continue
        val = adapter.expectation(mps, tensor[:4 ** k].reshape([2] * (2 * k)), tuple(int(i) for i in idx[:k]))
        assert abs(val - ref) < 4 * atol, (idx, val, ref)
    # none of the above may change the state (core_test.py:1552-1559, 1595-1599)
    np.testing.assert_allclose(np.asarray(mps.wavefunction()), wf0, rtol=0, atol=0)
    # sampling: same numpy RNG stream -> the very same draws as the reference (core.py:654-721)
    np.random.seed(1234)
    if pre + "sample_raises" in Z.files:
        try:
            mps.sample(nsamples=2)
        except ValueError as e:
            assert "do not sum to 1" in str(e)
        else:
            raise AssertionError("sampling an unnormalised state must raise like the reference")
    else:
        got = np.array(mps.sample(nsamples=24), dtype=np.int64)
        np.testing.assert_array_equal(got, Z[pre + "samples_seed1234"])


def check_inner_products(adapter, atol):
    a = build_state(adapter, "brick6")
    b = build_state(adapter, "brick6b")
    assert abs(complex(a.inner_product(b)) - complex(Z["inner_brick6_brick6b"])) < atol
    assert abs(complex(b.inner_product(a)) - complex(Z["inner_brick6b_brick6"])) < atol


def check_from_wavefunction(adapter, atol):
    for t in range(int(Z["n_fw"])):
        n, d = (int(v) for v in Z[f"fw{t}_nd"])
        wf = Z[f"fw{t}_input"]
        mps = adapter.from_wavefunction(wf, n, d)
        assert mps.bond_dimensions() == Z[f"fw{t}_bonds"].tolist()
        np.testing.assert_allclose(np.asarray(mps.wavefunction()), Z[f"fw{t}_wavefunction"], rtol=0, atol=atol)
        assert abs(mps.norm() - float(Z[f"fw{t}_norm"])) < atol
        site = int(Z[f"fw{t}_rdm_site"])
        np.testing.assert_allclose(np.asarray(mps.reduced_density_matrix(site)), Z[f"fw{t}_rdm"], rtol=0, atol=atol)
