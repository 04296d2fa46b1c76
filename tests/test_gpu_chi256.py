"""Parity at the large-chi shapes of BASELINE.json (configs[2]: 100 qubits, depth 20, chi = 256;
configs[4]: chi = 1024, i.e. 2048 x 2048 thetas), all through the C-ABI on the GPU:

* a 22-qubit chi=256 brickwork circuit against the complex128 oracle (512 x 512 SVDs on the
  block-Jacobi path, theta on the tcgen05 kernel): kept counts exact, singular values of every
  application within 1e-5 sigma_max, fidelity >= 1 - 1e-5;
* configs[2] at FULL size through size-independent properties (the oracle needs minutes there):
  every SVD converged, bonds follow the static rule, updated sites are isometries, norm in (0, 1],
  the amplitude kernel agrees with a complex128 host contraction of the device's own sites;
* mpsb_svd at 1024 x 1024 and 2048 x 2048 against LAPACK singular values.
"""
import numpy as np
import pytest

from oracle.mps_oracle import OracleMPS
from oracle.dense_sim import fidelity

pytestmark = pytest.mark.gpu


def _triples(ops, chi):
    return [(op.tensor, op.indices, {"maxsvals": chi, "keep_left_canonical": op.keep_left_canonical}) for op in ops]


def test_chi256_circuit_vs_oracle():
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    n, depth, chi = 22, 12, 256
    ops = circuits.brickwork(n, depth, seed=3)
    mps = mp.MPS(n)
    mps.record_singular_values(True)
    mps._execute(_triples(ops, chi))
    assert (mps.last_status()[:, 0] == 0).all()
    svs = mps.last_singular_values()
    ora = OracleMPS(n, dtype=np.complex128)
    for op in ops:
        ora.apply_two_qudit_gate(op.tensor, *op.indices, maxsvals=chi, keep_left_canonical=op.keep_left_canonical)
    assert mps.bond_dimensions() == ora.bond_dimensions()
    assert max(mps.bond_dimensions()) == chi
    n512 = sum(1 for t in ora.trace if 2 * min(t["chi"][0], t["chi"][2]) == 512)
    assert n512 >= 8, n512                                                 # 512 x 512 thetas, truncated to 256
    for s, t in zip(svs, ora.trace):
        assert s["k"] == t["k"]
    # singular values with the SAME input on both sides (teacher-forced), 1e-5 of sigma_max per application
    tfm = mp.MPS(n)
    tfm.record_singular_values(True)
    tf = OracleMPS(n, dtype=np.complex128)
    worst = 0.0
    for op in ops:
        for s_ in op.indices:
            tfm._chain.set_site(s_, tf.sites[s_])
        tf.apply_two_qudit_gate(op.tensor, *op.indices, maxsvals=chi, keep_left_canonical=op.keep_left_canonical)
        tfm.apply_two_qudit_gate(mp.Node(op.tensor), *op.indices, maxsvals=chi, keep_left_canonical=op.keep_left_canonical)
        (s,), t = tfm.last_singular_values(), tf.trace[-1]
        ref = np.concatenate([t["s_kept"], t["s_trunc"]])
        worst = max(worst, np.abs(s["svals"] - ref).max() / ref.max())
    print(f"worst singular-value error {worst:.2e} sigma_max over {len(svs)} applications ({n512} of 512 x 512), same input")
    assert worst <= 1e-5, worst
    # the free-running state (errors of earlier applications compound): norm, amplitudes, fidelity
    assert abs(mps.norm() - ora.norm()) < 1e-4
    wf, wref = mps.wavefunction(), ora.wavefunction()
    assert np.abs(wf - wref).max() < 1e-4
    assert fidelity(wf, wref) >= 1 - 1e-5


def test_config3_full_size_properties():
    import torch
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    n, depth, chi = 100, 20, 256
    ops = circuits.brickwork(n, depth, seed=3)
    mps = mp.MPS(n)
    mps.record_singular_values(True)
    mps._execute(_triples(ops, chi))
    status = mps.last_status()
    assert len(status) == 990 and (status[:, 0] == 0).all()
    # static bond rule (core.py:1105-1130): min(maxsvals, d chi_L, d chi_R) from a product state
    expect = [min(chi, 2 ** min(i + 1, n - 1 - i, depth)) for i in range(n - 1)]
    assert mps.bond_dimensions() == expect
    for s in mps.last_singular_values():
        sv = s["svals"]
        assert s["k"] <= chi and np.all(np.diff(sv) <= 1e-6 * sv[0]) and sv[-1] >= 0      # sorted on the device
    nrm = mps.norm()
    assert 0.0 < nrm <= 1.0 + 1e-5
    # the last layer (odd bonds, keep_left_canonical) left isometries on its left sites
    last = [op for op in ops[-49:]]
    for op in last[::12]:
        i = min(op.indices)
        a = mps.site_tensor(i).to(torch.complex128)
        m = a.reshape(-1, a.shape[2])
        g = (m.conj().T @ m).cpu().numpy()
        iso = g if op.keep_left_canonical else None
        if iso is None:
            b = mps.site_tensor(i + 1).to(torch.complex128)
            mb = b.reshape(b.shape[0], -1)
            iso = (mb @ mb.conj().T).cpu().numpy()
        assert np.abs(iso - np.eye(iso.shape[0])).max() < 5e-5
    # amplitude kernel vs a complex128 host contraction of the same site tensors
    rng = np.random.RandomState(0)
    bits = rng.randint(0, 2, size=(4, n)).astype(np.uint8)
    amps = mps.amplitudes(bits)
    sites = [mps.site_tensor(i).cpu().numpy().astype(np.complex128) for i in range(n)]
    for b, amp in zip(bits, amps):
        v = np.ones(1, dtype=np.complex128)
        for i in range(n):
            v = v @ sites[i][:, b[i], :]
        assert abs(v[0] - amp) <= 1e-5 * max(abs(v[0]), 1e-30) + 1e-12


@pytest.mark.parametrize("m", [1024, 2048])
def test_svd_large_sizes_vs_lapack(m):
    from tests.test_gpu_kernels import _svd, _graded
    rng = np.random.RandomState(m)
    mats = np.stack([_graded(rng, m, m, 6.0)])
    left, right, sv, info = _svd(mats, m // 2, 1)
    assert (info[:, 0] == 0).all(), info
    sref = np.linalg.svd(mats[0].astype(np.complex128), compute_uv=False)
    assert np.abs(sv[0] - sref).max() <= 1e-5 * sref[0], np.abs(sv[0] - sref).max() / sref[0]
    iso = left[0]
    assert np.abs(iso.conj().T @ iso - np.eye(m // 2)).max() < 5e-5
    # exact projection: |Q Q^H X|_F^2 = sum of the kept sigma^2
    assert abs(np.linalg.norm(left[0] @ right[0]) ** 2 - (sref[: m // 2] ** 2).sum()) < 1e-4 * (sref ** 2).sum()
