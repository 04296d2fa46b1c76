"""GPU parity tests proper: the CUDA path (through the reference-shaped Python API, which calls
the C-ABI) against (1) golden vectors produced by the unmodified reference, (2) the reference's
own known answers, (3) the complex128 oracle on seeded random circuits.

Tolerances (BASELINE.json north_star): kept singular-value counts exact; singular values within
1e-5 (relative to the largest of the application); amplitudes within 1e-4 absolute; state
fidelity >= 1 - 1e-5."""
import numpy as np
import pytest

from oracle.mps_oracle import OracleMPS, cphase as ocphase
from oracle.dense_sim import DenseState, fidelity
from tests import _golden

pytestmark = pytest.mark.gpu

SV_TOL = 1e-5
AMP_TOL = 1e-4
FID_TOL = 1e-5


def _mps(n, **kw):
    import mpsim_b200
    return mpsim_b200.MPS(n, **kw)


def _node(t):
    import mpsim_b200
    return mpsim_b200.Node(np.array(t, copy=True))


@pytest.mark.parametrize("name", _golden.names())
def test_golden_reference_vectors(name):
    g = _golden.Golden(name)
    mps = _mps(g.n)
    mps.record_singular_values(True)
    sv_log = []
    for tensor, idx, left in g.ops:
        if len(idx) == 1:
            mps.apply_one_qudit_gate(_node(tensor), idx[0])
        else:
            kw = dict(g.kwargs)
            if not left:
                kw["keep_left_canonical"] = False
            mps.apply_two_qudit_gate(_node(tensor), idx[0], idx[1], **kw)
            sv_log += mps.last_singular_values()
            assert (mps.last_status()[:, 0] == 0).all()
    assert mps.bond_dimensions() == g.bond_dimensions                       # counts: exact
    assert [s["k"] for s in sv_log] == [len(s) for s in g.s_kept]
    first_null = _golden.first_rank_deficient(g)
    for t, (s, s_ref, r_ref) in enumerate(zip(sv_log, g.s_kept, g.s_trunc)):
        if first_null is not None and t > first_null:
            break           # reference is gauge-unstable after a kept zero
        ref = np.concatenate([s_ref, r_ref])
        scale = max(ref.max(), 1e-30) if ref.size else 1.0
        assert np.abs(s["svals"] - ref).max() <= SV_TOL * scale, (name, t)
    untruncated = all((r ** 2).sum() < 1e-20 for r in g.s_trunc)
    if not (first_null is None or untruncated):
        assert mps.norm() <= 1.0 + 1e-4
        return
    assert abs(mps.norm() - g.norm) < 1e-4
    if g.wavefunction is not None:
        wf = mps.wavefunction()
        np.testing.assert_allclose(wf, g.wavefunction, atol=AMP_TOL)
        if g.norm > 1e-6:
            assert fidelity(wf, g.wavefunction) >= 1 - FID_TOL
    else:
        bits = ((g.amp_indices[:, None] >> np.arange(g.n - 1, -1, -1)[None, :]) & 1).astype(np.uint8)
        amps = mps.amplitudes(bits)
        np.testing.assert_allclose(amps, g.amp_values, atol=AMP_TOL)
        wf = mps.wavefunction()
        np.testing.assert_allclose(wf[g.amp_indices], g.amp_values, atol=AMP_TOL)


# ---- the reference's known-answer tests, through the drop-in API --------------------------------
def test_initial_state_and_one_qubit_gates():               # core_test.py:299-406
    import mpsim_b200 as mp
    for n in (2, 3, 5):
        mps = mp.MPS(n)
        wf = np.zeros(2 ** n); wf[0] = 1
        assert np.allclose(mps.wavefunction(), wf)
        assert mps.bond_dimensions() == [1] * (n - 1)
        mps.x(-1)
        wf = np.zeros(2 ** n); wf[-1] = 1
        assert np.allclose(mps.wavefunction(), wf)
        mps = mp.MPS(n); mps.h(-1)
        assert np.allclose(mps.wavefunction(), np.ones(2 ** n) / 2 ** (n / 2), atol=1e-6)
    for gate, expected in ((mp.xgate(), [0, 1]), (mp.hgate(), [2 ** -0.5, 2 ** -0.5]), (mp.zgate(), [1, 0])):
        mps = mp.MPS(4); mps.apply_one_qudit_gate(gate, 2)
        assert np.allclose(mps.get_node(2).tensor.reshape(2), expected, atol=1e-6)


def test_cnot_truth_table_and_flip():                       # core_test.py:409-472
    import mpsim_b200 as mp
    for prep, ctrl_first, out in (((0,), True, 3), ((), True, 0), ((1,), True, 1), ((0, 1), True, 2),
                                  ((0,), False, 2), ((1,), False, 3), ((0, 1), False, 1)):
        mps = mp.MPS(2)
        for q in prep:
            mps.x(q)
        if ctrl_first:
            mps.apply_two_qudit_gate(mp.cnot(), 0, 1)
        else:
            mps.cnot(1, 0)
        correct = np.zeros(4); correct[out] = 1
        assert np.allclose(mps.wavefunction(), correct, atol=1e-6)


def test_bell_truncation_and_bond_growth():                 # core_test.py:905-973, README.md:48-53
    import mpsim_b200 as mp
    mps = mp.MPS(2); mps.h(0); mps.cnot(0, 1, fraction=0.5)
    assert np.allclose(mps.wavefunction(), [2 ** -0.5, 0, 0, 0], atol=1e-6)
    mps = mp.MPS(2); mps.h(0); mps.cnot(0, 1, fraction=1)
    assert np.allclose(mps.wavefunction(), [2 ** -0.5, 0, 0, 2 ** -0.5], atol=1e-6)
    mps = mp.MPS(2)
    assert mps.bond_dimension_of(0) == 1
    mps.h(0); mps.cnot(0, 1)
    assert mps.is_valid() and mps.bond_dimension_of(0) == 2
    mps.cnot(0, 1)
    assert mps.bond_dimension_of(0) == 2
    mps = mp.MPS(2); mps.x(0); mps.cnot(0, 1, max_singular_values=0.5)     # ignored kwarg, core_test.py:910
    assert np.allclose(mps.wavefunction(), [0, 0, 0, 1], atol=1e-6)
    mps = mp.MPS(4); mps.r(-1); mps.apply_two_qudit_gate(mp.cnot(), 0, 1, fraction=0.5)
    assert mps.bond_dimensions() == [1, 1, 1]
    mps = mp.MPS(2); mps.h(0); mps.cnot(0, 1, maxsvals=0)                  # core_test.py:1093-1101
    assert mps.bond_dimensions() == [0] and mps.norm() == 0.0
    assert np.allclose(mps.wavefunction(), 0)


@pytest.mark.parametrize("left", [True, False])
def test_three_cnots_is_swap(left):                         # core_test.py:854-873
    import mpsim_b200 as mp
    for n in range(2, 9):
        mps = mp.MPS(n); mps.x(0)
        mps.cnot(0, 1, keep_left_canonical=left)
        mps.h(-1); mps.cnot(0, 1, keep_left_canonical=left); mps.h(-1)
        mps.cnot(0, 1)
        correct = np.zeros(2 ** n); correct[2 ** (n - 2)] = 1
        assert np.allclose(mps.wavefunction(), correct, atol=1e-5)


def test_hopping_nonlocal_ghz_qft():                        # core_test.py:888-902, 1225-1258
    import mpsim_b200 as mp
    for n in (2, 5, 12):
        mps = mp.MPS(n); mps.x(0)
        for i in range(n - 1):
            mps.swap(i, i + 1, keep_left_canonical=True)
        for i in range(n - 1, 0, -1):
            mps.swap(i - 1, i, keep_left_canonical=True)
        correct = np.zeros(2 ** n); correct[2 ** (n - 1)] = 1
        assert mps.is_valid() and np.allclose(mps.wavefunction(), correct, atol=1e-5)
    for n in range(3, 10):
        mps = mp.MPS(n); mps.x(0); mps.cnot(0, n - 1)
        correct = np.zeros(2 ** n); correct[2 ** (n - 1) + 1] = 1
        assert np.allclose(mps.wavefunction(), correct, atol=1e-5)
        mps = mp.MPS(n); mps.h(0)
        for i in range(1, n):
            mps.cnot(0, i)
        correct = np.zeros(2 ** n); correct[0] = correct[-1] = 2 ** -0.5
        assert np.allclose(mps.wavefunction(), correct, atol=1e-5)
        mps = mp.MPS(n)
        for i in range(n - 1, -1, -1):
            mps.h(i)
            for j in range(i - 1, -1, -1):
                mps.apply_two_qudit_gate(mp.cphase(2 ** (j - i)), j, i)
        assert np.allclose(mps.wavefunction(), np.ones(2 ** n) / 2 ** (n / 2), atol=1e-5)


def test_norm_renormalize_inner_product():                  # core_test.py:976-1184
    import mpsim_b200 as mp
    mps = mp.MPS(2); mps.h(0); mps.cnot(0, 1, maxsvals=1)
    assert np.isclose(mps.norm(), 2 ** -0.5, atol=1e-6)
    mps.renormalize()
    assert np.isclose(mps.norm(), 1.0, atol=1e-6)
    assert np.allclose(mps.wavefunction(), [1, 0, 0, 0], atol=1e-6)
    mps.renormalize(to_norm=2.0)
    assert np.isclose(mps.norm(), 2.0, atol=1e-5)
    for bad in (-1.0, 0.0):
        with pytest.raises(ValueError):
            mps.renormalize(bad)
    z = mp.MPS(2); z.h(0); z.cnot(0, 1, maxsvals=0)
    with pytest.raises(ValueError):
        z.renormalize()
    a = mp.MPS(5); b = mp.MPS(5)
    assert np.isclose(a.inner_product(b), 1.0)
    b.x(2)
    assert np.isclose(a.inner_product(b), 0.0)
    a.h(-1); b.h(-1)
    assert np.isclose(abs(a.inner_product(b)), 0.0, atol=1e-6)
    c = a.copy()
    assert c == a and c is not a
    c.x(0)
    assert np.isclose(c.inner_product(a), 1.0, atol=1e-6)     # X|+> = |+>
    with pytest.raises(ValueError):
        a.inner_product(mp.MPS(4))


def test_errors():                                          # core_test.py:475-485, 699-712; simulator_test.py:16-19
    import mpsim_b200 as mp
    with pytest.raises(ValueError):
        mp.MPS(1)
    mps = mp.MPS(3)
    for call in (lambda: mps.cnot(0, 0), lambda: mps.cnot(0, 3), lambda: mps.cnot(-1, 1),
                 lambda: mps.cnot(0, 1, fraction=0.5, maxsvals=1), lambda: mps.cnot(0, 1, fraction=2),
                 lambda: mps.apply_one_qudit_gate(mp.xgate(), 3), lambda: mps.apply_one_qudit_gate(mp.cnot(), 0),
                 lambda: mps.apply_two_qudit_gate(mp.xgate(), 0, 1),
                 lambda: mps.apply_two_qudit_gate(mp.Node(np.zeros((3, 3, 3, 3))), 0, 1),
                 lambda: mps.apply(mp.MPSOperation(mp.Node(np.zeros((2,) * 6)), (0, 1, 2)))):
        with pytest.raises(ValueError):
            call()
    with pytest.raises(TypeError):
        mps.apply([mp.xgate()])
    assert np.allclose(mps.wavefunction(), [1] + [0] * 7)     # nothing was launched by the failures


def test_apply_with_operations_and_dispatcher():            # core_test.py:1187-1222
    import mpsim_b200 as mp
    rng = np.random.RandomState(7)
    from mpsim_b200.gates import haar_random_unitary_tensor as haar
    n = 9
    ops, raw = [], []
    for _ in range(40):
        if rng.rand() < 0.4:
            t, idx = haar(1, 2, rng=rng), (int(rng.randint(n)),)
        else:
            i, j = rng.choice(n, size=2, replace=False)
            t, idx = haar(2, 2, rng=rng), (int(i), int(j))
        ops.append(mp.MPSOperation(mp.Node(t), idx)); raw.append((t, idx))
    mps = mp.MPS(n); mps.apply(ops)
    dense = DenseState(n).run(raw).wavefunction()
    np.testing.assert_allclose(mps.wavefunction(), dense, atol=AMP_TOL)
    assert fidelity(mps.wavefunction(), dense) >= 1 - FID_TOL
    # one by one gives the same state as the batched dispatch
    seq = mp.MPS(n)
    for op in ops:
        seq.apply(op)
    np.testing.assert_allclose(seq.wavefunction(), dense, atol=AMP_TOL)


# ---- against the complex128 oracle on seeded random circuits -------------------------------------
@pytest.mark.parametrize("n,depth,chi,seed", [(8, 8, None, 1), (12, 10, 8, 2), (16, 12, 16, 3), (20, 10, 64, 1),
                                               (16, 14, 96, 4)])   # last: d*chi = 192 > 128, large-chi path
def test_brickwork_vs_oracle(n, depth, chi, seed):
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    ops = circuits.brickwork(n, depth, seed)
    kw = {} if chi is None else {"maxsvals": chi}
    ora = OracleMPS(n, dtype=np.complex128)
    for op in ops:
        ora.apply_two_qudit_gate(op.tensor, *op.indices, keep_left_canonical=op.keep_left_canonical, **kw)
    # (a) gate by gate with the SAME input on both sides (teacher-forced: the two sites a gate touches are
    # set to the oracle's before it is applied): every singular value of every application within 1e-5
    # of the application's largest, on the single-CTA path (d*chi <= 128) and on the block-Jacobi path
    mps = mp.MPS(n); mps.record_singular_values(True)
    tf = OracleMPS(n, dtype=np.complex128)
    for op in ops:
        for s_ in op.indices:
            mps._chain.set_site(s_, tf.sites[s_])
        tf.apply_two_qudit_gate(op.tensor, *op.indices, keep_left_canonical=op.keep_left_canonical, **kw)
        mps.apply_two_qudit_gate(mp.Node(op.tensor), *op.indices, keep_left_canonical=op.keep_left_canonical, **kw)
        (s,), t = mps.last_singular_values(), tf.trace[-1]
        assert s["k"] == t["k"]
        ref = np.concatenate([t["s_kept"], t["s_trunc"]])
        assert np.abs(s["svals"] - ref).max() <= SV_TOL * ref.max()
    assert mps.bond_dimensions() == ora.bond_dimensions()
    # free running (nothing reset in between): state, norm and amplitudes
    mps = mp.MPS(n)
    for op in ops:
        mps.apply_two_qudit_gate(mp.Node(op.tensor), *op.indices, keep_left_canonical=op.keep_left_canonical, **kw)
    assert mps.bond_dimensions() == ora.bond_dimensions()
    assert abs(mps.norm() - ora.norm()) < 1e-4
    wf, wref = mps.wavefunction(), ora.wavefunction()
    np.testing.assert_allclose(wf, wref, atol=AMP_TOL)
    assert fidelity(wf, wref) >= 1 - FID_TOL
    # (b) the same circuit through the moment dispatcher in one call
    disp = mp.MPS(n)
    disp._execute([(op.tensor, op.indices, dict(kw, keep_left_canonical=op.keep_left_canonical)) for op in ops])
    assert (disp.last_status()[:, 0] == 0).all()
    np.testing.assert_allclose(disp.wavefunction(), wref, atol=AMP_TOL)
    # amplitudes kernel agrees with the dense contraction
    rng = np.random.RandomState(0)
    idx = rng.choice(2 ** n, size=32, replace=False)
    bits = ((idx[:, None] >> np.arange(n - 1, -1, -1)[None, :]) & 1).astype(np.uint8)
    np.testing.assert_allclose(disp.amplitudes(bits), wref[idx], atol=AMP_TOL)


@pytest.mark.parametrize("d", [3, 4])
def test_qudit_circuit_vs_oracle(d):
    """Qudit dimensions beyond 2 (the reference is dimension-agnostic, mpsim/core.py:161-166 and its tests build
    d = 3 ... 5 states): Haar one- and two-qudit gates on adjacent sites in both orders, with and without
    truncation.  (Non-adjacent gates route through the QUBIT swap gate in the reference, core.py:1376-1380,
    and fail there for d > 2 -- here too, with the same ValueError.)"""
    import mpsim_b200 as mp
    from mpsim_b200.gates import haar_random_unitary_tensor
    n = 5
    rng = np.random.RandomState(10 + d)
    pairs = [(0, 1), (2, 3), (1, 2), (3, 4), (2, 1), (4, 3), (2, 3), (1, 0)]
    for kw in ({}, {"maxsvals": 2 * d}):
        mps = mp.MPS(n, qudit_dimension=d)
        ora = OracleMPS(n, qudit_dimension=d, dtype=np.complex128)
        for q in range(n):
            g = haar_random_unitary_tensor(1, d, rng=rng)
            mps.apply_one_qudit_gate(mp.Node(g), q); ora.apply_one_qudit_gate(g, q)
        for (a, b) in pairs:
            g = haar_random_unitary_tensor(2, d, rng=rng)
            mps.apply_two_qudit_gate(mp.Node(g), a, b, **kw); ora.apply_two_qudit_gate(g, a, b, **kw)
        assert mps.bond_dimensions() == ora.bond_dimensions()
        assert abs(mps.norm() - ora.norm()) < 1e-4
        np.testing.assert_allclose(mps.wavefunction(), ora.wavefunction(), atol=AMP_TOL)
    with pytest.raises(ValueError):
        mp.MPS(n, qudit_dimension=d).apply_two_qudit_gate(mp.Node(haar_random_unitary_tensor(2, d, rng=rng)), 0, 3)
    # two-qudit gates on d > 4 are refused on the host, before anything is launched (one-qudit gates,
    # from_wavefunction and the observables work for any d)
    with pytest.raises(ValueError, match="qudit dimension"):
        mp.MPS(3, qudit_dimension=5).apply_two_qudit_gate(mp.Node(haar_random_unitary_tensor(2, 5, rng=rng)), 0, 1)


def test_mps_on_a_non_current_device():
    """An MPS / MPSBatch built on cuda:1 while cuda:0 is current launches on cuda:1's stream against cuda:1's
    memory (every DeviceChain entry point switches device; the library's stream pool and pinned slots are per
    device).  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    torch.cuda.set_device(0)
    n = 8
    mps = mp.MPS(n, device="cuda:1")
    mps.h(0)
    for i in range(n - 1):
        mps.cnot(i, i + 1)
    mps.cnot(0, n - 1)                                       # swap network
    assert mps._chain.slab.device.index == 1 and torch.cuda.current_device() == 0
    ghz = np.zeros(2 ** n); ghz[0] = ghz[-2] = 2 ** -0.5     # last CNOT flips the last qubit of |1..1>
    np.testing.assert_allclose(mps.wavefunction(), ghz, atol=1e-6)
    assert abs(mps.norm() - 1) < 1e-6
    # a chi = 96 bond: the block-Jacobi path (library-owned streams and pinned read-back slots of device 1)
    ops = circuits.brickwork(16, 14, seed=4)
    big = mp.MPS(16, device="cuda:1")
    ref = mp.MPS(16, device="cuda:0")
    for m in (big, ref):
        m._execute([(o.tensor, o.indices, {"maxsvals": 96, "keep_left_canonical": o.keep_left_canonical}) for o in ops])
    assert (big.last_status()[:, 0] == 0).all() and big.bond_dimensions() == ref.bond_dimensions()
    assert abs(big.norm() - ref.norm()) < 1e-4
    assert torch.cuda.current_device() == 0


def test_batch_matches_oracle_per_member():
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    n, depth, chi, B = 10, 8, 8, 6
    structure = circuits.brickwork(n, depth, seed=0)
    batch = mp.MPSBatch(B, n)
    cp = batch.compile(structure, record_svals=True, maxsvals=chi)
    gates = np.zeros((len(structure), B, 16), dtype=np.complex64)
    members = [circuits.brickwork(n, depth, seed=1000 + b) for b in range(B)]
    for b, ops in enumerate(members):
        for t, op in enumerate(ops):
            gates[t, b] = op.tensor.reshape(-1)
    batch.stage_gates(cp, gates)
    batch.run(cp)
    norms = batch.norms()
    sv = batch.singular_values(cp)
    assert (batch.status(cp)[..., 0] == 0).all()
    for b, ops in enumerate(members):
        ora = OracleMPS(n, dtype=np.complex128)
        for op in ops:
            ora.apply_two_qudit_gate(op.tensor, *op.indices, keep_left_canonical=op.keep_left_canonical, maxsvals=chi)
        assert abs(norms[b] - ora.norm()) < 1e-4
        for t, tr in enumerate(ora.trace):
            ref = np.concatenate([tr["s_kept"], tr["s_trunc"]])
            assert np.abs(sv[t, b, :ref.size] - ref).max() <= SV_TOL * ref.max()
        np.testing.assert_allclose(batch.wavefunction(b), ora.wavefunction(), atol=AMP_TOL)
    # second run after reset reproduces the first (plan re-use)
    batch.reset(); batch.run(cp)
    np.testing.assert_allclose(batch.norms(), norms, atol=1e-6)
    batch.renormalize()
    np.testing.assert_allclose(batch.norms(), 1.0, atol=1e-5)


def test_simulator_ghz_qft_and_sweep():                     # simulator_test.py:113-144, 147-259
    import mpsim_b200 as mp
    from mpsim_b200.mpsim_cirq import MPSimulator
    from tests._fake_cirq import Circuit, H, CNOT, CZPow, Rx, Toffoli
    for n in range(3, 8):
        circ = Circuit([H(0)] + [CNOT(0, i) for i in range(1, n)])
        wf = MPSimulator().simulate(circ).wavefunction()
        correct = np.zeros(2 ** n); correct[0] = correct[-1] = 2 ** -0.5
        assert np.allclose(wf, correct, atol=1e-5)
        ops = []
        for i in range(n - 1, -1, -1):
            ops.append(H(i))
            for j in range(i - 1, -1, -1):
                ops.append(CZPow(2.0 ** (j - i), j, i))
        assert np.allclose(MPSimulator().simulate(Circuit(ops)).wavefunction(), np.ones(2 ** n) / 2 ** (n / 2), atol=1e-5)
    res = MPSimulator({"maxsvals": 1}).simulate(Circuit([H(0), CNOT(0, 1)]))      # simulator_test.py:77-94
    assert isinstance(res, mp.MPS) and np.allclose(res.wavefunction(), [2 ** -0.5, 0, 0, 0], atol=1e-6)
    sweep = MPSimulator().simulate_sweep(Circuit([Rx("t", 0), CNOT(0, 1)]), [{"t": 0.0}, {"t": np.pi}])
    assert np.allclose(np.abs(sweep[0].wavefunction()), [1, 0, 0, 0], atol=1e-6)
    assert np.allclose(np.abs(sweep[1].wavefunction()), [0, 0, 0, 1], atol=1e-6)
    with pytest.raises(ValueError):
        MPSimulator().simulate(Circuit([Toffoli(0, 1, 2)]))                         # simulator_test.py:262-271
    with pytest.raises(ValueError):
        MPSimulator().simulate(Circuit([H(0)]))                                     # one qubit: simulator_test.py:16-19
    with pytest.raises(ValueError):
        MPSimulator().simulate("not a circuit")


def test_simulator_ghz_qft_maxsvals128_block_jacobi():
    """The regime of BASELINE configs[1] (GHZ + QFT through MPSimulator, maxsvals = 128) at a size whose
    wavefunction is computable: 14 qubits, bonds inflate to 128 (zeros are kept, core_test.py:932-944), the
    swap networks run ~700 adjacent applications, those with d chi > 128 on the block-Jacobi path with
    rank-deficient thetas.  Nothing physical is truncated (the true ranks are tiny), so the state must equal
    the dense simulation."""
    from mpsim_b200.mpsim_cirq import MPSimulator
    from tests._fake_cirq import Circuit, H, CNOT, CZPow
    n = 14
    ops = [H(0)] + [CNOT(0, i) for i in range(1, n)]
    for i in range(n - 1, -1, -1):
        ops.append(H(i))
        for j in range(i - 1, -1, -1):
            ops.append(CZPow(2.0 ** (j - i), j, i))
    circ = Circuit(ops)
    dense = DenseState(n)
    for op in circ.all_operations():
        dense.apply(np.asarray(op._unitary_()).reshape((2,) * (2 * len(op.qubits))), op.qubits)
    mps = MPSimulator({"maxsvals": 128}).simulate(circ)
    assert max(mps.bond_dimensions()) == 128
    assert (mps.last_status()[:, 0] == 0).all()
    wf, ref = mps.wavefunction(), dense.wavefunction()
    np.testing.assert_allclose(wf, ref, atol=AMP_TOL)
    assert fidelity(wf, ref) >= 1 - FID_TOL
    assert abs(mps.norm() - 1.0) < 1e-4


@pytest.mark.parametrize("nqubits", [2, 4, 8])
def test_simulator_random_circuits(nqubits):                # simulator_test.py:274-305
    """50 random circuits of 25 moments over the reference's gate domain (X, Y, Z, H, S, T, CNOT, CZ, SWAP,
    CZPow, ISWAP, FSim(0.2, 0.3)) on random -- mostly non-adjacent -- qubit pairs, through MPSimulator,
    against a dense state-vector simulation (standing in for circuit.final_wavefunction())."""
    from mpsim_b200.mpsim_cirq import MPSimulator
    from tests._fake_cirq import random_circuit
    rng = np.random.RandomState(1)
    for _ in range(50):
        circ = random_circuit(nqubits, 25, 0.999, rng)
        dense = DenseState(nqubits)
        for op in circ.all_operations():
            dense.apply(np.asarray(op._unitary_()).reshape((2,) * (2 * len(op.qubits))), op.qubits)
        wf = MPSimulator().simulate(circ).wavefunction()
        np.testing.assert_allclose(wf, dense.wavefunction(), atol=2e-5)


def test_simulator_custom_gate_and_sweep_of_50():           # simulator_test.py:308-357
    from mpsim_b200.mpsim_cirq import MPSimulator
    from mpsim_b200.gates import haar_random_unitary_tensor
    from tests._fake_cirq import Circuit, Op, Rx
    u = haar_random_unitary_tensor(2, 2, rng=np.random.RandomState(1)).reshape(4, 4)
    wf = MPSimulator().simulate(Circuit([Op((0, 1), u)])).wavefunction()
    np.testing.assert_allclose(wf, u[:, 0], atol=1e-6)
    params = [{"theta": t} for t in np.linspace(0, 2 * np.pi, 50)]
    allmps = MPSimulator().simulate_sweep(Circuit([Rx("theta", 0), Rx("theta", 1)]), params)
    assert len(allmps) == 50
    for pr, mps in zip(params, allmps):
        c, s_ = np.cos(pr["theta"] / 2), -1j * np.sin(pr["theta"] / 2)
        np.testing.assert_allclose(mps.wavefunction(), np.kron([c, s_], [c, s_]), atol=1e-6)


def test_simulator_batched_sweep_matches_sequential():          # simulator.py:67-87 as ONE batched run
    """A parameter sweep compiles once and runs its resolvers as one MPSBatch; the results equal the
    reference's sequential loop (one MPS per resolver), including non-adjacent gates (swap networks),
    flipped control/target order and truncation."""
    import mpsim_b200 as mp
    from mpsim_b200.mpsim_cirq import MPSimulator
    from tests._fake_cirq import Circuit, H, CNOT, CZPow, Rx
    n = 7
    ops = [H(0)] + [CNOT(i, i + 1) for i in range(n - 1)] + [Rx("t", q) for q in range(n)]
    ops += [CZPow(0.3, 0, 4), CNOT(5, 2), Rx("u", 3), CNOT(6, 0)] + [Rx("t", q) for q in range(0, n, 2)]
    circ = Circuit(ops)
    params = [{"t": 0.17 * (i + 1), "u": 1.0 - 0.2 * i} for i in range(6)]
    for options in ({}, {"maxsvals": 3}):
        seq = MPSimulator(dict(options, batch_sweeps=False)).simulate_sweep(circ, params)
        bat = MPSimulator(dict(options)).simulate_sweep(circ, params)
        assert len(seq) == len(bat) == len(params)
        for a, b in zip(seq, bat):
            assert isinstance(b, mp.MPS) and a.bond_dimensions() == b.bond_dimensions()
            np.testing.assert_allclose(b.wavefunction(), a.wavefunction(), atol=2e-6)
            assert abs(a.norm() - b.norm()) < 1e-5
        bits = np.array([[0] * n, [1] * n, [1, 0, 1, 0, 1, 0, 1]], dtype=np.uint8)
        res = MPSimulator(dict(options)).simulate_sweep_batched(circ, params, amplitudes=bits)
        assert res.local_range == (0, len(params)) and res.total == len(params)
        idx = (bits.astype(np.int64) << np.arange(n - 1, -1, -1)[None, :]).sum(axis=1)
        for i, a in enumerate(seq):
            assert abs(res.norms[i] - a.norm()) < 1e-5
            np.testing.assert_allclose(res.amplitudes[i], a.wavefunction()[idx], atol=2e-6)
        np.testing.assert_allclose(res.mps(2).wavefunction(), seq[2].wavefunction(), atol=2e-6)
    # circuits whose structure differs between resolvers, or with a non-unitary gate, keep the sequential path
    with pytest.raises(ValueError):
        MPSimulator().simulate_sweep_batched(Circuit([H(0), Toffoli_or_none(0, 1, 2)]), params)


def Toffoli_or_none(a, b, c):
    from tests._fake_cirq import Toffoli
    return Toffoli(a, b, c)


def test_non_unitary_gate_orthonormalizes_and_renormalizes():      # core_test.py:1341-1428
    import mpsim_b200 as mp
    from mpsim_b200.gates import computational_basis_projector
    mps = mp.MPS(2); mps.h(0); mps.cnot(0, 1)
    mps.apply_one_qudit_gate(computational_basis_projector(0), 0)
    assert mps.bond_dimensions() == [1]                     # bond 2 -> 1 after the projector
    assert np.isclose(mps.norm(), 1.0, atol=1e-5)
    assert np.allclose(np.abs(mps.wavefunction()), [1, 0, 0, 0], atol=1e-5)
    ora = OracleMPS(3, dtype=np.complex128); m = mp.MPS(3)
    for o in (ora, m):
        o.h(0); o.cnot(0, 1); o.cnot(1, 2)
    proj = np.zeros((2, 2)); proj[1, 1] = 1
    ora.apply_one_qudit_gate(proj, 1); m.apply_one_qudit_gate(mp.Node(proj), 1)
    assert m.bond_dimensions() == ora.bond_dimensions()
    assert fidelity(m.wavefunction(), ora.wavefunction()) >= 1 - FID_TOL
    # a GENERIC rank-deficient state: Haar layers, then a projector in the middle of the chain.  The bonds
    # next to the projected site lose rank; the fp32 SVD returns the numerically zero singular values at
    # ~1e-7 sigma_max, above the reference's 1e-8 * norm cut, so the cut is clamped to the fp32 noise
    # floor (ortho.py) -- the bond dimensions must come out as in the complex128 reference
    from mpsim_b200 import circuits
    n = 8
    ops = circuits.brickwork(n, 6, seed=12)
    ora = OracleMPS(n, dtype=np.complex128); m = mp.MPS(n)
    for op in ops:
        ora.apply_two_qudit_gate(op.tensor, *op.indices, keep_left_canonical=op.keep_left_canonical)
        m.apply_two_qudit_gate(mp.Node(op.tensor), *op.indices, keep_left_canonical=op.keep_left_canonical)
    for site in (3, 4):
        ora.apply_one_qudit_gate(proj, site); m.apply_one_qudit_gate(mp.Node(proj), site)
    assert m.bond_dimensions() == ora.bond_dimensions(), (m.bond_dimensions(), ora.bond_dimensions())
    assert max(ora.bond_dimensions()) < 16                    # the projectors did reduce the middle bonds
    assert abs(m.norm() - ora.norm()) < 1e-4
    assert fidelity(m.wavefunction(), ora.wavefunction()) >= 1 - FID_TOL


# ---- site-scale drift of long circuits (complex64 storage) -----------------------------------------
def test_rebalance_is_exact_and_zero_sum():
    """``mpsb_rebalance_sites``: power-of-two shifts that sum to zero per chain -- every site ends up
    at about the same exponent and the wavefunction is unchanged bit for bit; a chain whose sites
    are within the spread is not touched at all."""
    import torch
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    n = 10
    mps = mp.MPS(n)
    ops = circuits.brickwork(n, 6, seed=11)
    mps._execute([(o.tensor, o.indices, {"keep_left_canonical": o.keep_left_canonical}) for o in ops])
    chain = mps._chain
    before = chain.slab.clone()
    chain.rebalance()                                   # sites within 2**32 of each other: untouched
    assert torch.equal(chain.slab, before)
    wf0 = mps.wavefunction()
    shifts = [40, -35, 0, 17, -22, 0, 0, 30, -30, 0]    # sums to zero: the state is the same
    for i, sh in enumerate(shifts):
        chain.set_site(i, chain.site_view(i).clone() * float(2.0 ** sh))
    np.testing.assert_array_equal(mps.wavefunction(), wf0)
    chain.rebalance()
    ex = [int(np.floor(np.log2(float(chain.site_view(i).abs().max())))) for i in range(n)]
    assert max(ex[:-1]) - min(ex[:-1]) <= 2, ex          # the last site takes the remainder of the sum
    assert abs(ex[-1] - ex[0]) <= n + 2, ex
    np.testing.assert_array_equal(mps.wavefunction(), wf0)
    assert mps.norm() == pytest.approx(1.0, abs=1e-5)


def test_long_untruncated_circuit_stays_in_range():
    """320 brickwork layers on 12 qubits, nothing truncated.  The reference's alternating
    left/right-canonical sweeps (core.py:1348-1360) let the scale of individual sites drift
    geometrically -- the complex128 oracle ends this circuit with one site at 1e-49 and two at
    1e24, beyond complex64 -- and run() rebalances the exponents on the way (powers of two whose
    product is one).  The state must stay finite, normalised and equal to the oracle's."""
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    n, depth = 12, 320
    ops = circuits.brickwork(n, depth, seed=5)
    ora = OracleMPS(n, dtype=np.complex128)
    for op in ops:
        ora.apply_two_qudit_gate(op.tensor, *op.indices, keep_left_canonical=op.keep_left_canonical)
    mx = [float(np.abs(s).max()) for s in ora.sites]
    assert min(mx) < 1e-38 and max(mx) > 1e18            # the drift is the algorithm's, not ours
    mps = mp.MPS(n)
    mps._execute([(o.tensor, o.indices, {"keep_left_canonical": o.keep_left_canonical}) for o in ops])
    assert (mps.last_status()[:, 0] == 0).all()
    ours = [float(mps._chain.site_view(i).abs().max()) for i in range(n)]
    assert all(np.isfinite(ours)) and min(ours) > 1e-12 and max(ours) < 1e12, ours
    assert abs(mps.norm() - 1.0) < 2e-4
    wf, wref = mps.wavefunction(), ora.wavefunction()
    assert np.isfinite(wf).all()
    assert fidelity(wf, wref) >= 1 - 1e-4
    # gate by gate (transient one-application plans) takes the same precaution
    seq = mp.MPS(n)
    for op in ops:
        seq.apply_two_qudit_gate(mp.Node(op.tensor), *op.indices, keep_left_canonical=op.keep_left_canonical)
    assert fidelity(seq.wavefunction(), wref) >= 1 - 1e-4
