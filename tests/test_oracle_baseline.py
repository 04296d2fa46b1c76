"""CPU: the oracle against the golden vectors at BASELINE sizes (tests/golden/baseline/).  The
configs[3] members and the swap-network circuit take seconds; configs[2] in full takes minutes and only
runs with MPSB_SLOW_TESTS=1.  Also pins the reference's norm-bookkeeping order for routed gates."""
import os

import numpy as np
import pytest

from oracle.mps_oracle import OracleMPS
from tests import _baseline


def _run(base, track=False):
    mps = OracleMPS(base.n, dtype=np.complex128, track_norms=track)
    for op in base.ops:
        mps.apply_two_qudit_gate(np.asarray(op.tensor), *op.indices, maxsvals=base.chi,
                                 keep_left_canonical=op.keep_left_canonical)
    return mps


@pytest.mark.parametrize("name", ["config3_member0", "config3_member511", "snake_4x4_chi96"]
                         + (["config2_full"] if os.environ.get("MPSB_SLOW_TESTS") else []))
def test_oracle_reproduces_baseline_fixture(name):
    base = _baseline.Baseline(name)
    mps = _run(base, track=base.norms_after is not None and base.n <= 16)
    assert [t["k"] for t in mps.trace] == base.k
    assert mps.bond_dimensions() == base.bond_dimensions
    tol = 1e-6 if name == "config2_full" else 1e-10              # configs[2] singular values are stored in float32
    for t, ref in zip(mps.trace, base.svals):
        got = np.concatenate([t["s_kept"], t["s_trunc"]])
        assert np.abs(got - ref).max() <= tol * ref.max()
    assert abs(mps.norm() - base.norm) <= 1e-10 * base.norm
    amps = _baseline.amplitudes_of(mps.sites, base.amp_bits)
    assert np.abs(amps - base.amp_values).max() <= 1e-9 * np.abs(base.amp_values).max()
    if mps.track_norms:
        # the reference's own _norms (core.py:1160-1161), swap networks included
        np.testing.assert_allclose(mps._norms, base.norms_after, rtol=1e-10)


def test_fixtures_record_the_complex64_reference_drift():
    """The reference handed complex64 gates computes in complex64 (numpy promotion) and is itself
    > 1e-5 sigma_max away from the complex128 result on every baseline circuit."""
    for name in ("config3_member0", "config3_member511", "snake_4x4_chi96"):
        dev = _baseline.Baseline(name).reference_complex64_sigma_deviation
        assert dev is not None and 1e-5 < dev < 1e-3
