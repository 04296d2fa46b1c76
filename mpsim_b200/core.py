"""``MPS`` / ``MPSOperation`` with the reference's signatures, kwargs and errors
(``mpsim/core.py``), backed by the device store and the sm_100a kernels.

What differs from the reference, deliberately:
  * site tensors live on the GPU as complex64 in a canonical ``[chi_left][d][chi_right]``
    layout (the reference keeps ``tn.Node`` objects with no fixed axis order and silently
    promotes to complex128);
  * ``apply`` dispatches operations on disjoint sites concurrently (``# TODO: Parallelize``
    at ``mpsim/core.py:1245``);
  * the per-application norm bookkeeping of ``mpsim/core.py:1160-1161`` ("TODO: Remove") is
    opt-in: ``MPS(..., track_norms=True)``.
"""
from copy import deepcopy
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from mpsim_b200 import gates as _gates
from mpsim_b200.node import BondEdge, Node, tensor_of
from mpsim_b200.planner import Plan, max_bond_dimensions, plan_operations

BITSTRING = Union[Sequence[int], str]


class CannotConvertToMPSOperation(Exception):
    pass


class MPSOperation:
    """An operation which can act on a matrix product state (``mpsim/core.py:21-156``).

    Gate edge convention (``mpsim/core.py:43-63``): one-qudit gates contract axis 1 with the
    site, axis 0 becomes the new physical index; two-qudit gates on (i < j) contract axes 2, 3
    with sites i, j and axes 0, 1 become their new physical indices
    (``matrix.reshape(2, 2, 2, 2)`` satisfies it).  If i > j, 0 <-> 1 and 2 <-> 3.
    """

    def __init__(self, node: Any, qudit_indices: Union[int, Tuple[int, ...]], qudit_dimension: int = 2) -> None:
        self._node = node
        if isinstance(qudit_indices, (int, np.integer)):
            qudit_indices = (int(qudit_indices),)
        self._qudit_indices = tuple(qudit_indices)
        self._qudit_dimension = int(qudit_dimension)

    @property
    def qudit_indices(self) -> Tuple[int, ...]:
        return self._qudit_indices

    @property
    def qudit_dimension(self) -> int:
        return self._qudit_dimension

    @property
    def num_qudits(self) -> int:
        return len(self._qudit_indices)

    def node(self, copy: bool = True) -> Any:
        if not copy:
            return self._node
        return Node(np.array(tensor_of(self._node), copy=True), name=getattr(self._node, "name", None))

    def tensor(self, reshape_to_square_matrix: bool = True) -> np.ndarray:
        tensor = deepcopy(tensor_of(self._node))
        if reshape_to_square_matrix:
            dim = self._qudit_dimension ** self.num_qudits
            tensor = np.reshape(tensor, (dim, dim))
        return tensor

    def is_valid(self) -> bool:
        """``mpsim/core.py:113-129``: shape (d,)*2n and all edges free."""
        d = self._qudit_dimension
        t = tensor_of(self._node)
        if not t.shape == tuple([d] * 2 * self.num_qudits):
            return False
        has_nd = getattr(self._node, "has_nondangling_edge", None)
        if callable(has_nd) and has_nd():
            return False
        return True

    def is_unitary(self) -> bool:
        return _gates.is_unitary(self.tensor(reshape_to_square_matrix=True))

    def is_hermitian(self) -> bool:
        return _gates.is_hermitian(self.tensor(reshape_to_square_matrix=True))

    def is_single_qudit_operation(self) -> bool:
        return self.num_qudits == 1

    def is_two_qudit_operation(self) -> bool:
        return self.num_qudits == 2

    def __str__(self) -> str:
        return f"Tensor {getattr(self._node, 'name', '?')} on qudit(s) {self._qudit_indices}."


def _check_gate_edges(gate: Any, nfree: int, what: str) -> np.ndarray:
    """Edge-count validation of ``mpsim/core.py:791-796, 1013-1018`` for duck-typed gates."""
    t = tensor_of(gate)
    nd = getattr(gate, "get_all_nondangling", None)
    if t.ndim != nfree or (callable(nd) and len(nd()) != 0):
        raise ValueError(what)
    return t


class MPS:
    """Matrix product state on the GPU (``mpsim/core.py:159-1423``)."""

    _last_bond_from_right = False        # graph-view detail (see _bond_edge); set by __init__

    def __init__(self, nqudits: int, qudit_dimension: int = 2, tensor_prefix: str = "q",
                 track_norms: bool = False, device: Any = None) -> None:
        if nqudits < 2:                                              # core.py:184-187
            raise ValueError(f"Number of qudits must be greater than 2 but is {nqudits}.")
        from mpsim_b200.store import DeviceChain
        self._nqudits = int(nqudits)
        self._qudit_dimension = int(qudit_dimension)
        self._prefix = tensor_prefix
        self._chain = DeviceChain(self._nqudits, self._qudit_dimension, 1, device)
        self._max_bond_dimensions = max_bond_dimensions(self._nqudits, self._qudit_dimension)
        self._track_norms = bool(track_norms)
        self._norms: List[float] = []
        self._last = None            # last CompiledPlan (svals / status for inspection)
        self._record_svals = False
        self._last_bond_from_right = self._nqudits >= 3      # graph-view detail, see _bond_edge

    @staticmethod
    def from_wavefunction(wavefunction: Any, nqudits: int, qudit_dimension: int = 2, tensor_prefix: str = "q",
                          device: Any = None) -> "MPS":             # core.py:245-328
        from mpsim_b200 import observables
        return observables.from_wavefunction(MPS, wavefunction, nqudits, qudit_dimension, tensor_prefix, device)

    # ------------------------------------------------------------------ properties
    @property
    def nqudits(self) -> int:
        return self._nqudits

    @property
    def qudit_dimension(self) -> int:
        return self._qudit_dimension

    def bond_dimension_of(self, node_index: int) -> int:            # core.py:340-361
        if not self.is_valid():
            raise ValueError("MPS is invalid.")
        if node_index >= self._nqudits:
            raise ValueError(f"Index should be less than {self._nqudits} but is {node_index}.")
        if node_index == self._nqudits - 1:
            raise IndexError("list index out of range")          # get_node(node_index + 1), core.py:358
        return self._chain.bonds[node_index + 1]

    def bond_dimensions(self) -> List[int]:                          # core.py:363-365
        return [self.bond_dimension_of(i) for i in range(self._nqudits - 1)]

    def max_bond_dimension_of(self, edge_index: int) -> int:         # core.py:367-382
        if edge_index >= self._nqudits:
            raise ValueError(f"Edge index should be less than {self._nqudits} but is {edge_index}.")
        return self._max_bond_dimensions[edge_index]

    def max_bond_dimensions(self) -> List[int]:                      # core.py:384-386
        return self._max_bond_dimensions

    def is_valid(self) -> bool:
        """The device store is a chain by construction (``mpsim/core.py:388-419`` walks a graph)."""
        c = self._chain
        return c.n >= 2 and c.bonds[0] == 1 and c.bonds[-1] == 1 and len(c.bonds) == c.n + 1

    # ------------------------------------------------------------------ compatibility views
    def get_node(self, node_index: int, copy: bool = True) -> Node:
        """Host copy of one site as a ``Node`` in the reference's initial axis order
        (phys, left, right); chain ends drop their dummy bond (``mpsim/core.py:190-218``)."""
        n = self._nqudits
        i = range(n)[node_index]
        t = self._chain.site_view(i).permute(1, 0, 2).cpu().numpy()
        if i == 0:
            t = t.reshape(t.shape[0], t.shape[2])
        elif i == n - 1:
            t = t.reshape(t.shape[0], t.shape[1])
        return Node(t.copy(), name=self._prefix + str(i))

    def get_nodes(self, copy: bool = True) -> List[Node]:
        return [self.get_node(i) for i in range(self._nqudits)]

    def get_free_edge_of(self, node_index: int, copy: bool = True) -> BondEdge:      # core.py:443-451
        """The physical (dangling) leg of a site."""
        i = range(self._nqudits)[node_index]
        return BondEdge(self, ("phys", i), self.get_node(i), None, self._qudit_dimension)

    def _bond_edge(self, bond: int) -> BondEdge:
        """Bond between sites ``bond`` and ``bond + 1``.  ``node1`` / ``node2`` follow the order in
        which the reference connects the chain (``mpsim/core.py:220-229``): left site first, except
        the last bond of a chain of three or more sites, which is connected from the right end --
        until a two-site application re-creates it (``split_node`` puts the left factor first)."""
        left, right = self.get_node(bond), self.get_node(bond + 1)
        if self._last_bond_from_right and bond == self._nqudits - 2:
            left, right = right, left
        return BondEdge(self, ("bond", bond), left, right, self._chain.bonds[bond + 1])

    def get_left_connected_edge_of(self, node_index: int) -> Optional[BondEdge]:      # core.py:453-466
        i = range(self._nqudits)[node_index]
        return None if i == 0 else self._bond_edge(i - 1)

    def get_right_connected_edge_of(self, node_index: int) -> Optional[BondEdge]:     # core.py:468-481
        i = range(self._nqudits)[node_index]
        return None if i == self._nqudits - 1 else self._bond_edge(i)

    def site_tensor(self, index: int):
        """Device view of site ``index`` as a torch tensor [chi_left][d][chi_right]."""
        return self._chain.site_view(range(self._nqudits)[index])

    def _device_guard(self):
        """Context with the chain's device current (helpers outside DeviceChain take the current stream)."""
        import torch
        return torch.cuda.device(self._chain.device)

    # ------------------------------------------------------------------ contractions
    def wavefunction(self) -> np.ndarray:                            # core.py:483-500
        if not self.is_valid():
            raise ValueError("MPS is not valid.")
        self._chain.check_status()
        return self._chain.wavefunction(0).cpu().numpy()

    def wavefunction_device(self):
        return self._chain.wavefunction(0)

    def amplitudes(self, bitstrings) -> np.ndarray:
        """<bits|psi> for selected basis states (the only option once d**n is out of reach)."""
        self._chain.check_status()
        return self._chain.amplitudes(bitstrings)[0].cpu().numpy()

    def dagger(self) -> None:                                        # core.py:502-505
        self._chain.slab.conj_physical_()

    def inner_product(self, other: "MPS") -> complex:                # core.py:507-561
        if other._nqudits != self._nqudits:
            raise ValueError(
                f"Cannot compute inner product between self which has {self._nqudits} qudits and "
                f"other which has {other._nqudits} qudits.\nNumber of qudits must be equal.")
        if other._qudit_dimension != self._qudit_dimension:
            raise ValueError("Cannot compute inner product: qudit dimensions must be equal.")
        if not self.is_valid():
            raise ValueError("MPS is invalid.")
        if not other.is_valid():
            raise ValueError("Other MPS is invalid.")
        self._chain.check_status()
        return complex(self._chain.inner_products(other._chain)[0].item())

    def norm(self) -> float:                                         # core.py:563-565
        return float(np.sqrt(max(self.inner_product(self).real, 0.0)))

    def renormalize(self, to_norm: float = 1.0) -> None:             # core.py:567-594
        if to_norm < 0.0:
            raise ValueError(f"Arg to_norm must be positive but is {to_norm}")
        if np.isclose(to_norm, 0.0, atol=1e-15):
            raise ValueError(f"Arg to_norm = {to_norm} is too close to numerical zero.")
        norm = self.norm()
        if np.isclose(norm, 0.0, atol=1e-15):
            raise ValueError("Norm of MPS is numerically zero, cannot renormalize.")
        self._chain.scale([(to_norm / norm) ** (1 / self.nqudits)])

    def reduced_density_matrix(self, node_indices: Union[int, Sequence[int]]) -> np.ndarray:   # core.py:596-652
        from mpsim_b200 import observables
        with self._device_guard():
            return observables.reduced_density_matrix(self, node_indices)

    def sample(self, nsamples: int, as_hist: bool = False, as_string: bool = False) -> Any:      # core.py:684-721
        from mpsim_b200 import observables
        with self._device_guard():
            return observables.sample(self, nsamples, as_hist, as_string)

    def expectation(self, observable: MPSOperation) -> float:        # core.py:723-751
        from mpsim_b200 import observables
        with self._device_guard():
            return observables.expectation(self, observable)

    # ------------------------------------------------------------------ gate application
    def _execute(self, ops: Sequence[Tuple[np.ndarray, Tuple[int, ...], Dict[str, Any]]]) -> None:
        plan = plan_operations(self._nqudits, self._qudit_dimension, self._chain.bonds, ops)
        if not plan.order:
            return
        if any(a.site == self._nqudits - 2 for a in plan.apps2):
            self._last_bond_from_right = False
        if self._track_norms and plan.apps2:
            # the reference records the norm after EVERY adjacent application (core.py:1160-1161);
            # to reproduce that the plan is cut after each one
            self._execute_tracking(plan, ops)
            return
        cp = self._chain.compile(plan, record_svals=self._record_svals, transient=True)
        self._chain.run(cp)
        self._last = cp

    def _execute_tracking(self, plan: Plan, ops) -> None:
        """One compiled step per application so that the norm can be read in between.  Order of the
        recorded norms = the reference's: every SWAP of a swap network is itself an
        ``apply_two_qudit_gate`` call and appends its norm, while the routed gate appends only after
        its swap-back (``mpsim/core.py:1154-1161``) -- i.e. once more on the final state."""
        d = self._qudit_dimension
        last_of_op: Dict[int, int] = {}
        for pos, (kind, idx) in enumerate(plan.order):
            if kind == 2:
                last_of_op[plan.apps2[idx].source_op] = pos
        for pos, (kind, idx) in enumerate(plan.order):
            sub = Plan(self._nqudits, d, self._chain.bonds)
            if kind == 1:
                a = plan.apps1[idx]
                sub.add_one(plan.gates[a.gate_index].reshape(d, d), a.site)
            else:
                a = plan.apps2[idx]
                kw = {"keep_left_canonical": a.left_canonical, "maxsvals": a.k}
                sub._add_adjacent(plan.gates[a.gate_index].reshape(d, d, d, d), a.site, kw, a.source_op, a.is_swap)
            cp = self._chain.compile(sub, record_svals=self._record_svals, transient=True)
            self._chain.run(cp)
            self._last = cp
            if kind == 2:
                routed = last_of_op[a.source_op] != pos or a.is_swap     # part of a swap network
                if a.is_swap or not routed:
                    self._norms.append(self.norm())
                if routed and last_of_op[a.source_op] == pos:
                    self._norms.append(self.norm())                      # the routed gate's own entry

    def apply_one_qudit_gate(self, gate: Any, node_index: int, **kwargs: Any) -> None:
        """``mpsim/core.py:753-845``.  Unitary gates run entirely on the device; the non-unitary
        branch (orthonormalise + renormalise, core.py:816-817, 828-845) is SURVEY.md 8(f) row 1."""
        if not self.is_valid():
            raise ValueError("MPS is invalid.")
        if node_index not in range(self._nqudits):
            raise ValueError(f"Input tensor index={node_index} is out of bounds for an MPS on "
                             f"{self._nqudits} qudits.")
        t = _check_gate_edges(gate, 2, "Single qudit gate must have two free edges and zero connected edges.")
        if t.shape[0] != t.shape[1]:
            raise ValueError("Gate edge dimensions must be equal.")
        if t.shape[0] != self._qudit_dimension:
            raise ValueError(f"Gate edges have dimension {t.shape[0]} but should have MPS qudit "
                             f"dimension = {self._qudit_dimension}")
        renormalize_after = kwargs.get("renormalize_after_non_unitary") is not False
        ortho_after = kwargs.get("ortho_after_non_unitary") is not False
        unitary = _gates.is_unitary(t)
        if not unitary and renormalize_after:
            norm = self.norm()
        self._execute([(t, (int(node_index),), {})])
        if not unitary and ortho_after:
            from mpsim_b200 import ortho
            if node_index == 0:
                ortho.orthonormalize_right_edge_of(self, node_index)
            elif node_index == self._nqudits - 1:
                ortho.orthonormalize_left_edge_of(self, node_index)
            else:
                ortho.orthonormalize_right_edge_of(self, node_index)
                ortho.orthonormalize_left_edge_of(self, node_index)
        if not unitary and renormalize_after:
            self.renormalize(norm)

    def orthonormalize_right_edge_of(self, node_index: int, threshold: float = 1e-8) -> None:
        from mpsim_b200 import ortho
        ortho.orthonormalize_right_edge_of(self, node_index, threshold)

    def orthonormalize_left_edge_of(self, node_index: int, threshold: float = 1e-8) -> None:
        from mpsim_b200 import ortho
        ortho.orthonormalize_left_edge_of(self, node_index, threshold)

    def apply_one_qudit_gate_to_all(self, gate: Any) -> None:        # core.py:941-948
        t = tensor_of(gate)
        if _gates.is_unitary(t) and t.ndim == 2:
            self._execute([(t, (i,), {}) for i in range(self._nqudits)])
        else:
            for i in range(self._nqudits):
                self.apply_one_qudit_gate(gate, i)

    def apply_two_qudit_gate(self, gate: Any, node_index1: int, node_index2: int, **kwargs: Any) -> None:
        """``mpsim/core.py:950-1161``: kwargs ``keep_left_canonical`` (default True),
        ``maxsvals`` xor ``fraction``; unknown kwargs are ignored, as in the reference."""
        if not self.is_valid():
            raise ValueError("MPS is not valid.")
        if node_index1 not in range(self._nqudits) or node_index2 not in range(self._nqudits):
            raise ValueError(f"Input tensor indices={(node_index1, node_index2)} are out of bounds for an "
                             f"MPS on {self._nqudits} qudits.")
        if node_index1 == node_index2:
            raise ValueError("Node indices cannot be identical.")
        t = _check_gate_edges(gate, 4, "Two qubit gate must have four free edges and zero connected edges.")
        self._execute([(t, (int(node_index1), int(node_index2)), dict(kwargs))])

    def move_node_from_left_to_right(self, current_node_index: int, final_node_index: int, **kwargs: Any) -> None:
        """``mpsim/core.py:1163-1190``."""
        if current_node_index > final_node_index:
            raise ValueError("current_node_index should be smaller than final_node_index.")
        if current_node_index < 0:
            raise ValueError("current_node_index out of range.")
        if final_node_index >= self._nqudits:
            raise ValueError("final_node_index out of range.")
        ops = [(_gates.swap().tensor, (s, s + 1), dict(kwargs)) for s in range(current_node_index, final_node_index)]
        if ops:
            self._execute(ops)

    def move_node_from_right_to_left(self, current_node_index: int, final_node_index: int, **kwargs: Any) -> None:
        """``mpsim/core.py:1192-1219``."""
        if current_node_index < final_node_index:
            raise ValueError("current_node_index should be larger than final_node_index.")
        if current_node_index > self._nqudits:
            raise ValueError("current_node_index out of range.")
        if final_node_index < 0:
            raise ValueError("final_node_index out of range.")
        ops = [(_gates.swap().tensor, (s - 1, s), dict(kwargs))
               for s in range(current_node_index, final_node_index, -1)]
        if ops:
            self._execute(ops)

    def apply(self, operations: Union[MPSOperation, Sequence[MPSOperation]], **kwargs: Any) -> None:
        """``mpsim/core.py:1221-1276``.  Operations on disjoint sites are launched together."""
        try:
            operations = iter(operations)
        except TypeError:
            operations = (operations,)
        ops = []
        pending_non_unitary = False
        for op in operations:
            if not isinstance(op, MPSOperation):
                raise TypeError(f"Argument operation should be of type MPSOperation but is of type {type(op)}.")
            if not op.is_valid():
                raise ValueError("Input MPS Operation is not valid.")
            if op.is_single_qudit_operation():
                t = tensor_of(op.node(copy=False))
                if not _gates.is_unitary(t):
                    # non-unitary gates go through the sequential path (orthonormalise + renormalise)
                    if ops:
                        self._execute(ops)
                        ops = []
                    self.apply_one_qudit_gate(op.node(), *op.qudit_indices, **kwargs)
                    continue
                ops.append((t, op.qudit_indices, {}))
            elif op.is_two_qudit_operation():
                if op.qudit_dimension != self._qudit_dimension:
                    raise ValueError(f"Gate edges have dimension {op.qudit_dimension} but should have MPS "
                                     f"qudit dimension = {self._qudit_dimension}")
                ops.append((tensor_of(op.node(copy=False)), op.qudit_indices, dict(kwargs)))
            else:
                raise ValueError(
                    "Only one-qudit and two-qudit gates are supported. To apply a gate on three or more "
                    "qudits, the gate must be compiled into a sequence of one- and two-qudit gates.")
        del pending_non_unitary
        if ops:
            self._execute(ops)

    # ------------------------------------------------------------------ conveniences (qubits)
    def x(self, index: int) -> None:                                 # core.py:1279-1290
        if index == -1:
            self.apply_one_qudit_gate_to_all(_gates.xgate())
        else:
            self.apply_one_qudit_gate(_gates.xgate(), index)

    def h(self, index: int) -> None:                                 # core.py:1292-1303
        if index == -1:
            self.apply_one_qudit_gate_to_all(_gates.hgate())
        else:
            self.apply_one_qudit_gate(_gates.hgate(), index)

    def r(self, index: int, seed: Optional[int] = None, angle_scale: float = 1.0) -> None:   # core.py:1305-1321
        if index == -1:
            for i in range(self._nqudits):
                self.apply_one_qudit_gate(_gates.rgate(seed, angle_scale), i)
        else:
            self.apply_one_qudit_gate(_gates.rgate(seed, angle_scale), index)

    def cnot(self, a: int, b: int, **kwargs: Any) -> None:           # core.py:1324-1328
        self.apply_two_qudit_gate(_gates.cnot(), a, b, **kwargs)

    def haar_random(self, qudit1_index: int, qudit2_index: int, **kwargs: Any) -> None:      # core.py:1330-1346
        gate = _gates.haar_random_unitary(nqudits=2, qudit_dimension=self._qudit_dimension)
        self.apply_two_qudit_gate(gate, qudit1_index, qudit2_index, **kwargs)

    def sweep_haar_random_left_to_right(self, **kwargs: Any) -> None:        # core.py:1348-1353
        d = self._qudit_dimension
        self._execute([(_gates.haar_random_unitary_tensor(2, d), (i, i + 1), dict(kwargs, keep_left_canonical=True))
                       for i in range(0, self._nqudits - 1, 2)])

    def sweep_haar_random_right_to_left(self, **kwargs: Any) -> None:        # core.py:1355-1360
        d = self._qudit_dimension
        self._execute([(_gates.haar_random_unitary_tensor(2, d), (i - 1, i), dict(kwargs, keep_left_canonical=False))
                       for i in range(self._nqudits - 2, 0, -2)])

    def sweep_cnots_left_to_right(self, **kwargs: Any) -> None:              # core.py:1362-1367
        self._execute([(_gates.cnot().tensor, (i, i + 1), dict(kwargs, keep_left_canonical=True))
                       for i in range(0, self._nqudits - 1, 2)])

    def sweep_cnots_right_to_left(self, **kwargs: Any) -> None:              # core.py:1369-1374
        self._execute([(_gates.cnot().tensor, (i - 1, i), dict(kwargs, keep_left_canonical=False))
                       for i in range(self._nqudits - 2, 0, -2)])

    def swap(self, a: int, b: int, **kwargs: Any) -> None:           # core.py:1376-1380
        if b < a:
            a, b = b, a
        self.apply_two_qudit_gate(_gates.swap(), a, b, **kwargs)

    # ------------------------------------------------------------------ diagnostics
    def record_singular_values(self, on: bool = True) -> None:
        """Keep the singular values of every application of the next call(s) on the device."""
        self._record_svals = bool(on)

    def last_singular_values(self) -> List[Dict[str, Any]]:
        """Per adjacent application of the last call, in program order: site, shape, k, all
        singular values (descending; the first k were kept)."""
        cp = self._last
        if cp is None or cp.svals is None:
            return []
        sv = cp.svals.cpu().numpy()
        pos = {idx: p for p, idx in enumerate(cp.order2)}
        out = []
        d = self._qudit_dimension
        for idx, a in enumerate(cp.plan.apps2):
            mn = min(d * a.chiL, d * a.chiR)
            out.append(dict(index=a.site, chi=(a.chiL, a.chiM, a.chiR), k=a.k, left=a.left_canonical,
                            is_swap=a.is_swap, svals=sv[pos[idx], 0, :mn].astype(np.float64)))
        return out

    def last_status(self) -> np.ndarray:
        """int32 [napplications][2] = (status, sweeps) of the last call's SVDs (0 = converged)."""
        cp = self._last
        if cp is None:
            return np.zeros((0, 2), dtype=np.int32)
        n2 = len(cp.plan.apps2)
        return cp.info.cpu().numpy()[:n2]

    # ------------------------------------------------------------------ copies / comparison
    def copy(self) -> "MPS":                                         # core.py:1382-1384
        return self.__copy__()

    def __copy__(self) -> "MPS":                                     # core.py:1420-1423
        new = MPS.__new__(MPS)
        new._nqudits, new._qudit_dimension, new._prefix = self._nqudits, self._qudit_dimension, self._prefix
        new._chain = self._chain.clone()
        new._max_bond_dimensions = list(self._max_bond_dimensions)
        new._track_norms, new._norms = self._track_norms, []
        new._last, new._record_svals = None, False
        new._last_bond_from_right = self._last_bond_from_right
        return new

    @classmethod
    def _from_chain(cls, chain, tensor_prefix: str = "q") -> "MPS":
        """An MPS around an existing one-member ``DeviceChain`` (e.g. ``MPSBatch`` member views)."""
        assert chain.B == 1
        new = cls.__new__(cls)
        new._nqudits, new._qudit_dimension, new._prefix = chain.n, chain.d, tensor_prefix
        new._chain = chain
        new._max_bond_dimensions = max_bond_dimensions(chain.n, chain.d)
        new._track_norms, new._norms = False, []
        new._last, new._record_svals = None, False
        new._last_bond_from_right = chain.n >= 3
        return new

    def __str__(self) -> str:
        return "----".join(self._prefix + str(i) for i in range(self._nqudits))

    def __eq__(self, other: Any) -> bool:                            # core.py:1389-1418
        if not isinstance(other, MPS):
            return False
        if self is other:
            return True
        if other._qudit_dimension != self._qudit_dimension or other._nqudits != self._nqudits:
            return False
        if self._chain.bonds != other._chain.bonds:
            return False
        import torch
        for i in range(self._nqudits):
            if not torch.allclose(self._chain.site_view(i), other._chain.site_view(i).to(self._chain.device)):
                return False
        return True

    __hash__ = None
