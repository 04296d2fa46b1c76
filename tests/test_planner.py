"""CPU: static shape planner / moment dispatcher against the oracle's traces and the golden
vectors (kept counts, swap-network expansion, layering)."""
import numpy as np
import pytest

from mpsim_b200 import circuits, planner
from oracle.mps_oracle import OracleMPS
from tests import _golden


def _triples(g):
    out = []
    for tensor, idx, left in g.ops:
        kw = dict(g.kwargs)
        if len(idx) == 2 and not left:
            kw["keep_left_canonical"] = False
        out.append((tensor, idx, kw))
    return out


@pytest.mark.parametrize("name", _golden.names())
def test_plan_matches_reference_counts(name):
    g = _golden.Golden(name)
    plan = planner.plan_operations(g.n, 2, [1] * (g.n + 1), _triples(g))
    assert plan.bonds[1:-1] == g.bond_dimensions
    assert [a.k for a in plan.apps2] == [len(s) for s in g.s_kept]          # kept counts, exact
    assert len(plan.apps2) == circuits.count_adjacent_applications(
        [circuits.Op(t, i, l) for t, i, l in g.ops])


def test_plan_shapes_match_oracle_trace():
    ops = circuits.brickwork(14, 9, seed=5)
    mps = OracleMPS(14)
    for op in ops:
        mps.apply_two_qudit_gate(op.tensor, *op.indices, maxsvals=12, keep_left_canonical=op.keep_left_canonical)
    plan = planner.plan_operations(14, 2, [1] * 15, [(o.tensor, o.indices, {"maxsvals": 12, "keep_left_canonical": o.keep_left_canonical}) for o in ops])
    assert [(a.site, (a.chiL, a.chiM, a.chiR), a.k, a.left_canonical) for a in plan.apps2] == \
           [(t["index"], t["chi"], t["k"], t["left"]) for t in mps.trace]


def test_layers_are_disjoint_and_ordered():
    ops = circuits.ghz_qft(9)
    plan = planner.plan_operations(9, 2, [1] * 10, [(o.tensor, o.indices, {"maxsvals": 8}) for o in ops])
    last = {}
    for kind, idx in plan.order:          # program order per site must be increasing in layer
        a = plan.apps1[idx] if kind == 1 else plan.apps2[idx]
        sites = [a.site] if kind == 1 else [a.site, a.site + 1]
        for s in sites:
            assert a.layer >= last.get(s, -1) + 1 or last.get(s, -1) == -1 and a.layer >= 0
            assert a.layer > last.get(s, -1)
            last[s] = a.layer
    for layer in plan.layers():
        used = [plan.apps1[i].site for i in layer["one"]]
        for idxs in layer["two"].values():
            for i in idxs:
                used += [plan.apps2[i].site, plan.apps2[i].site + 1]
        assert len(used) == len(set(used))
    assert plan.counts()["adjacent_applications"] == circuits.count_adjacent_applications(ops)


def test_benchmark_config_counts():
    assert circuits.count_adjacent_applications(circuits.brickwork(20, 10, 1)) == 95
    assert circuits.count_adjacent_applications(circuits.brickwork(100, 20, 3)) == 990
    assert circuits.count_adjacent_applications(circuits.brickwork(40, 20, 1000)) == 390
    assert circuits.count_adjacent_applications(circuits.ghz_qft(50)) == 42826
    n, ops = circuits.sycamore_snake()
    assert n == 53 and len(ops) == 317 and circuits.count_adjacent_applications(ops) == 3069


def test_fraction_and_errors():
    plan = planner.Plan(4, 2, [1] * 5)
    plan.add_two(np.eye(4).reshape(2, 2, 2, 2), 0, 1, {"fraction": 0.5})
    assert plan.apps2[0].k == 1                                   # round(0.5 * 2) = 1
    with pytest.raises(ValueError):
        planner.resolve_truncation({"fraction": 0.5, "maxsvals": 2}, 4, 2, 0)
    with pytest.raises(ValueError):
        planner.resolve_truncation({"fraction": 1.5}, 4, 2, 0)
    with pytest.raises(ValueError):
        planner.plan_operations(3, 2, [1] * 4, [(np.zeros((2,) * 6), (0, 1, 2), {})])
    with pytest.raises(ValueError):
        planner.plan_operations(3, 2, [1] * 4, [(np.eye(4).reshape(2, 2, 2, 2), (0, 0), {})])
    with pytest.raises(ValueError):
        planner.plan_operations(3, 2, [1] * 4, [(np.eye(4).reshape(2, 2, 2, 2), (0, 3), {})])


def test_shard_range():
    from mpsim_b200.distributed import shard_range
    for total in (4096, 10, 7):
        for world in (1, 2, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def test_graph_edge_views_on_a_stub_chain():      # core_test.py:165-227 (the GPU port is in test_gpu_reference_ported.py)
    """The reference's graph accessors are views computed from the chain's bookkeeping: check them
    on a host stub of the device store (no kernels involved)."""
    import torch
    from mpsim_b200.core import MPS

    class StubChain:
        def __init__(self, n, d):
            self.n, self.d = n, d
            self.bonds = [1] * (n + 1)

        def site_view(self, i, b=0):
            t = torch.zeros((self.bonds[i], self.d, self.bonds[i + 1]), dtype=torch.complex64)
            t[0, 0, 0] = 1
            return t

    for n in (2, 3, 6):
        mps = MPS.__new__(MPS)
        mps._nqudits, mps._qudit_dimension, mps._prefix = n, 3, "q"
        mps._chain = StubChain(n, 3)
        mps._last_bond_from_right = n >= 3
        for i in range(n):
            e = mps.get_free_edge_of(i, copy=False)
            assert e.is_dangling() and e.node1.name == f"q{i}" and e.dimension == 3
        assert mps.get_left_connected_edge_of(0) is None and mps.get_right_connected_edge_of(n - 1) is None
        for i in range(1, n):
            assert mps.get_right_connected_edge_of(i - 1) == mps.get_left_connected_edge_of(i)
            assert mps.get_left_connected_edge_of(i) != mps.get_free_edge_of(i)
        last = mps.get_left_connected_edge_of(n - 1)
        names = (last.node1.name, last.node2.name)
        assert names == ((f"q{n - 1}", f"q{n - 2}") if n >= 3 else ("q0", "q1"))
        assert not last.is_dangling() and last.dimension == 1
