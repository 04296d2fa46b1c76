"""Minimal stand-in for ``tensornetwork==0.2.1`` (numpy backend) -- TEST INFRASTRUCTURE ONLY.

Why this exists
---------------
The reference (``/root/reference/mpsim``) delegates all arithmetic of the two-qudit gate
path to the un-vendored PyPI package ``tensornetwork==0.2.1`` (``requirements.txt:2``),
which is not installed in this image and cannot be downloaded (no network).  This module
restates the *published* behaviour of the handful of graph operations that
``mpsim/core.py`` calls (55 call sites; list in SURVEY.md section 2.4) so that the
UNMODIFIED reference can be imported in the build container in order to

  * validate ``oracle/mps_oracle.py`` (the array-level restatement), and
  * generate the golden fixtures under ``tests/golden/`` (``tests/golden/make_golden.py``).

It is never imported by the product (``mpsim_b200``), never by ``bench.py`` and never on
the GPU box (``/root/reference`` does not exist there).  It is validated by running the
reference's own ``mpsim/core_test.py`` against it (see ``oracle/run_reference_tests.sh``).

Semantics restated (tensornetwork 0.2.1, ``network_components.py`` / ``network_operations.py``
/ ``backends/numpy/decompositions.py``):

  * ``contract(edge)``            = ``np.tensordot(edge.node1, edge.node2, [[axis1],[axis2]])``;
                                    result axes = node1's remaining, then node2's remaining.
  * ``contract_between(a, b)``    = ``np.tensordot`` over all shared edges; result axes = a's
                                    remaining (in order) then b's remaining.
  * ``flatten_edges_between``     = transpose shared axes to the end of both tensors (same
                                    order on both sides), merge them into one axis, reconnect.
  * ``split_node_full_svd``       = reorder to left+right edges, reshape to a matrix,
                                    ``np.linalg.svd(full_matrices=False)``, keep
                                    ``min(max_singular_values, n_above_trunc_err)`` values
                                    (no tolerance cut when only ``max_singular_values`` is
                                    given; zeros are kept), ``s`` cast to the tensor dtype,
                                    returned as a dense ``diag(s)`` node.
  * ``split_node``                = same SVD, ``sqrt(s)`` folded into both factors.
"""

from typing import Any, Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

__version__ = "0.2.1-shim"


class Edge:
    """Edge between two node axes (or a dangling edge when ``node2`` is None)."""

    def __init__(self, node1: "Node", axis1: int, name: Optional[str] = None,
                 node2: Optional["Node"] = None, axis2: Optional[int] = None) -> None:
        if (node2 is None) != (axis2 is None):
            raise ValueError("node2 and axis2 must either be both None or both not be None")
        self.node1 = node1
        self.axis1 = axis1
        self.node2 = node2
        self.axis2 = axis2
        self.name = name if name is not None else "__unnamed_edge__"

    def get_nodes(self) -> List[Optional["Node"]]:
        return [self.node1, self.node2]

    def is_dangling(self) -> bool:
        return self.node2 is None

    def is_trace(self) -> bool:
        return self.node1 is self.node2

    @property
    def dimension(self) -> int:
        return self.node1.tensor.shape[self.axis1]

    def update_axis(self, old_axis: int, old_node: "Node", new_axis: int,
                    new_node: "Node") -> None:
        if self.node1 is old_node and self.axis1 == old_axis:
            self.node1, self.axis1 = new_node, new_axis
        elif self.node2 is old_node and self.axis2 == old_axis:
            self.node2, self.axis2 = new_node, new_axis
        else:
            raise ValueError("Edge does not touch (old_node, old_axis).")

    def __xor__(self, other: "Edge") -> "Edge":
        return connect(self, other)

    def __str__(self) -> str:
        return self.name


class Node:
    """Tensor plus one :class:`Edge` per axis."""

    def __init__(self, tensor: Any, name: Optional[str] = None,
                 axis_names: Optional[List[str]] = None, backend: Any = None) -> None:
        if isinstance(tensor, Node):
            tensor = tensor.tensor
        self.tensor = np.asarray(tensor)
        self.name = name if name is not None else "__unnamed_node__"
        n = self.tensor.ndim
        if axis_names is not None and len(axis_names) != n:
            raise ValueError("axis_names is not the same length as the tensor shape.")
        self.axis_names = list(axis_names) if axis_names is not None else [str(i) for i in range(n)]
        self.edges = [Edge(self, i, self.axis_names[i]) for i in range(n)]

    # -- tensor access -------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.tensor.shape)

    def get_tensor(self) -> np.ndarray:
        return self.tensor

    def set_tensor(self, tensor: Any) -> None:
        self.tensor = np.asarray(tensor)

    def get_rank(self) -> int:
        return self.tensor.ndim

    # -- edge access ---------------------------------------------------------------------
    def get_edge(self, axis: Any) -> Edge:
        if isinstance(axis, str):
            axis = self.axis_names.index(axis)
        return self.edges[axis]

    def __getitem__(self, key: Any) -> Any:
        if isinstance(key, slice):
            return self.edges[key]
        return self.get_edge(key)

    def get_all_edges(self) -> List[Edge]:
        return list(self.edges)

    def get_all_dangling(self) -> Set[Edge]:
        return {e for e in self.edges if e.is_dangling()}

    def get_all_nondangling(self) -> Set[Edge]:
        return {e for e in self.edges if not e.is_dangling()}

    def has_nondangling_edge(self) -> bool:
        return any(not e.is_dangling() for e in self.edges)

    def has_dangling_edge(self) -> bool:
        return any(e.is_dangling() for e in self.edges)

    def add_edge(self, edge: Edge, axis: int, override: bool = False) -> None:
        self.edges[axis] = edge

    def fresh_edges(self, axis_names: Optional[List[str]] = None) -> None:
        """Give the node brand-new dangling edges (0.2.x does this to contracted nodes,
        which is why ``core_test.py:371-385`` can re-use one gate node many times)."""
        if axis_names is None:
            axis_names = [str(i) for i in range(self.tensor.ndim)]
        self.axis_names = list(axis_names)
        self.edges = [Edge(self, i, self.axis_names[i]) for i in range(self.tensor.ndim)]

    def reorder_edges(self, edge_order: List[Edge]) -> "Node":
        if set(edge_order) != set(self.edges) or len(edge_order) != len(self.edges):
            raise ValueError("Given edge order does not match expected edges.")
        if any(e.is_trace() for e in self.edges):
            raise NotImplementedError("reorder_edges with trace edges is not supported by the shim.")
        permutation = [self.edges.index(e) for e in edge_order]
        self.tensor = np.transpose(self.tensor, permutation)
        for new_axis, e in enumerate(edge_order):
            if e.node1 is self:
                e.axis1 = new_axis
            else:
                e.axis2 = new_axis
        self.edges = list(edge_order)
        self.axis_names = [self.axis_names[p] for p in permutation]
        return self

    def reorder_axes(self, perm: List[int]) -> "Node":
        return self.reorder_edges([self.edges[p] for p in perm])

    def __matmul__(self, other: "Node") -> "Node":
        return contract_between(self, other)

    def __str__(self) -> str:
        return self.name


# ---------------------------------------------------------------------------------------
# graph operations
# ---------------------------------------------------------------------------------------
def connect(edge1: Edge, edge2: Edge, name: Optional[str] = None) -> Edge:
    if edge1 is edge2:
        raise ValueError("Cannot connect an edge to itself.")
    for e in (edge1, edge2):
        if not e.is_dangling():
            raise ValueError("Edge '{}' is not a dangling edge.".format(e))
    if edge1.dimension != edge2.dimension:
        raise ValueError("Cannot connect edges of unequal dimension. "
                         "Dimension of edge '{}': {}, Dimension of edge '{}': {}.".format(
                             edge1, edge1.dimension, edge2, edge2.dimension))
    n1, a1 = edge1.node1, edge1.axis1
    n2, a2 = edge2.node1, edge2.axis1
    new_edge = Edge(n1, a1, name, n2, a2)
    n1.add_edge(new_edge, a1, override=True)
    n2.add_edge(new_edge, a2, override=True)
    return new_edge


def get_shared_edges(node1: Node, node2: Node) -> Set[Edge]:
    nodes = {node1, node2}
    shared = set()
    for e in node1.edges:
        if set(e.get_nodes()) == nodes:
            shared.add(e)
    return shared


def _remove_edges(edges: Set[Edge], node1: Node, node2: Node, new_node: Node) -> None:
    """Re-attach the surviving edges of node1 then node2 to ``new_node`` in order."""
    if node1 is node2:
        raise ValueError("node1 and node2 are the same ('{}'), use trace instead.".format(node1))
    node1_edges = node1.edges[:]
    node2_edges = node2.edges[:]
    node1_axis_names = node1.axis_names
    node2_axis_names = node2.axis_names
    remaining = []
    names = []
    for (i, e) in enumerate(node1_edges):
        if e not in edges:
            remaining.append((e, node1, i))
            names.append(node1_axis_names[i])
    for (i, e) in enumerate(node2_edges):
        if e not in edges:
            remaining.append((e, node2, i))
            names.append(node2_axis_names[i])
    new_node.edges = []
    new_node.axis_names = [str(i) for i in range(len(remaining))]
    for new_axis, (e, old_node, old_axis) in enumerate(remaining):
        e.update_axis(old_axis, old_node, new_axis, new_node)
        new_node.edges.append(e)
    node1.fresh_edges(node1_axis_names)
    node2.fresh_edges(node2_axis_names)


def _new_bare_node(tensor: np.ndarray, name: Optional[str]) -> Node:
    return Node(tensor, name=name)


def contract(edge: Edge, name: Optional[str] = None) -> Node:
    if edge.is_dangling():
        raise ValueError("Attempting to contract dangling edge '{}'".format(edge))
    if edge.is_trace():
        return _contract_trace_edge(edge, name)
    node1, node2 = edge.node1, edge.node2
    new_tensor = np.tensordot(node1.tensor, node2.tensor, [[edge.axis1], [edge.axis2]])
    new_node = _new_bare_node(new_tensor, name)
    _remove_edges({edge}, node1, node2, new_node)
    return new_node


def _contract_trace_edge(edge: Edge, name: Optional[str]) -> Node:
    node = edge.node1
    a1, a2 = edge.axis1, edge.axis2
    new_tensor = np.trace(node.tensor, axis1=a1, axis2=a2)
    new_node = _new_bare_node(new_tensor, name)
    remaining = [(e, i) for i, e in enumerate(node.edges) if e is not edge]
    new_node.edges = []
    for new_axis, (e, old_axis) in enumerate(remaining):
        e.update_axis(old_axis, node, new_axis, new_node)
        new_node.edges.append(e)
    new_node.axis_names = [str(i) for i in range(len(remaining))]
    node.fresh_edges(node.axis_names)
    return new_node


def outer_product(node1: Node, node2: Node, name: Optional[str] = None) -> Node:
    new_tensor = np.tensordot(node1.tensor, node2.tensor, 0)
    new_node = _new_bare_node(new_tensor, name)
    _remove_edges(set(), node1, node2, new_node)
    return new_node


def contract_between(node1: Node, node2: Node, name: Optional[str] = None,
                     allow_outer_product: bool = False,
                     output_edge_order: Optional[Sequence[Edge]] = None) -> Node:
    if node1 is node2:
        result = node1
        for e in [e for e in node1.edges if e.is_trace()]:
            if e in result.edges:
                result = _contract_trace_edge(e, name)
        return result
    shared = get_shared_edges(node1, node2)
    if not shared:
        if allow_outer_product:
            return outer_product(node1, node2, name)
        raise ValueError("No edges found between nodes '{}' and '{}' and "
                         "allow_outer_product=False.".format(node1, node2))
    axes1, axes2 = [], []
    for e in shared:
        if e.node1 is node1:
            axes1.append(e.axis1)
            axes2.append(e.axis2)
        else:
            axes1.append(e.axis2)
            axes2.append(e.axis1)
    new_tensor = np.tensordot(node1.tensor, node2.tensor, [axes1, axes2])
    new_node = _new_bare_node(new_tensor, name)
    _remove_edges(shared, node1, node2, new_node)
    if output_edge_order is not None:
        new_node.reorder_edges(list(output_edge_order))
    return new_node


def flatten_edges(edges: Sequence[Edge], new_edge_name: Optional[str] = None) -> Edge:
    edges = list(edges)
    if not edges:
        raise ValueError("At least 1 edge must be given.")
    if len(edges) == 1:
        return edges[0]
    expected_nodes = set(edges[0].get_nodes())
    for e in edges:
        if set(e.get_nodes()) != expected_nodes:
            raise ValueError("Two edges do not share the same nodes.")
    if len(expected_nodes) == 1:
        raise NotImplementedError("Flattening trace edges is not supported by the shim.")
    new_dangling = []
    for node in expected_nodes:
        if node is None:
            raise ValueError("Cannot flatten dangling edges in the shim.")
        flat_axes = [e.axis1 if e.node1 is node else e.axis2 for e in edges]
        keep_axes = [i for i in range(node.tensor.ndim) if i not in flat_axes]
        perm = keep_axes + flat_axes
        t = np.transpose(node.tensor, perm)
        keep_shape = t.shape[:len(keep_axes)]
        flat_dim = int(np.prod(t.shape[len(keep_axes):]))
        t = np.reshape(t, tuple(keep_shape) + (flat_dim,))
        kept_edges = [node.edges[i] for i in keep_axes]
        kept_names = [node.axis_names[i] for i in keep_axes]
        node.tensor = t
        new_axis = len(keep_axes)
        dangling = Edge(node, new_axis, "__Flattened_Edge__")
        for i, (e, old_axis) in enumerate(zip(kept_edges, keep_axes)):
            e.update_axis(old_axis, node, i, node)
        node.edges = kept_edges + [dangling]
        node.axis_names = kept_names + ["__flat__"]
        new_dangling.append(dangling)
    return connect(new_dangling[0], new_dangling[1], new_edge_name)


def flatten_edges_between(node1: Node, node2: Node) -> Optional[Edge]:
    shared = get_shared_edges(node1, node2)
    if shared:
        return flatten_edges(list(shared))
    return None


def check_connected(nodes: Iterable[Node]) -> None:
    nodes = list(nodes)
    if not nodes:
        return
    node_set = set(nodes)
    seen = {nodes[0]}
    stack = [nodes[0]]
    while stack:
        n = stack.pop()
        for e in n.edges:
            for other in e.get_nodes():
                if other is not None and other in node_set and other not in seen:
                    seen.add(other)
                    stack.append(other)
    if seen != node_set:
        raise ValueError("Non-connected graph")


def check_correct(nodes: Iterable[Node], check_connections: bool = True) -> None:
    for node in nodes:
        for i, e in enumerate(node.edges):
            if e.node1 is not node and e.node2 is not node:
                raise ValueError("Edge does not point back at its node.")
    if check_connections:
        check_connected(nodes)


def copy(nodes: Iterable[Node], conjugate: bool = False) -> Tuple[Dict[Node, Node], Dict[Edge, Edge]]:
    nodes = list(nodes)
    node_dict: Dict[Node, Node] = {}
    for n in nodes:
        t = np.conj(n.tensor) if conjugate else np.array(n.tensor, copy=True)
        node_dict[n] = Node(t, name=n.name, axis_names=list(n.axis_names))
    edge_dict: Dict[Edge, Edge] = {}
    for n in nodes:
        for e in n.edges:
            if e in edge_dict:
                continue
            n1, a1 = e.node1, e.axis1
            n2, a2 = e.node2, e.axis2
            if e.is_dangling() or n1 not in node_dict or n2 not in node_dict:
                # Dangling (or leaving the copied set): stays a dangling edge of the copy.
                owner, axis = (n1, a1) if n1 in node_dict else (n2, a2)
                new_edge = node_dict[owner].edges[axis]
                new_edge.name = e.name
                edge_dict[e] = new_edge
            else:
                new_edge = Edge(node_dict[n1], a1, e.name, node_dict[n2], a2)
                node_dict[n1].edges[a1] = new_edge
                node_dict[n2].edges[a2] = new_edge
                edge_dict[e] = new_edge
    return node_dict, edge_dict


# ---------------------------------------------------------------------------------------
# decompositions (numpy backend)
# ---------------------------------------------------------------------------------------
def _svd_decomposition(tensor: np.ndarray, split_axis: int,
                       max_singular_values: Optional[int],
                       max_truncation_error: Optional[float]):
    left_dims = tensor.shape[:split_axis]
    right_dims = tensor.shape[split_axis:]
    mat = np.reshape(tensor, [int(np.prod(left_dims)), int(np.prod(right_dims))])
    u, s, vh = np.linalg.svd(mat, full_matrices=False)
    if max_singular_values is None:
        max_singular_values = np.size(s)
    if max_truncation_error is not None:
        trunc_errs = np.sqrt(np.cumsum(np.square(s[::-1])))
        num_sing_vals_err = np.count_nonzero((trunc_errs > max_truncation_error).astype(np.int32))
    else:
        num_sing_vals_err = max_singular_values
    keep = min(max_singular_values, num_sing_vals_err)
    s = s.astype(mat.dtype)
    s_rest = s[keep:]
    s = s[:keep]
    u = u[:, :keep]
    vh = vh[:keep, :]
    dim_s = s.shape[0]
    u = np.reshape(u, list(left_dims) + [dim_s])
    vh = np.reshape(vh, [dim_s] + list(right_dims))
    return u, s, vh, s_rest


def _reattach(new_node: Node, axis: int, e: Edge, old_node: Node, old_axis: int) -> None:
    e.update_axis(old_axis, old_node, axis, new_node)
    new_node.edges[axis] = e


def split_node_full_svd(node: Node, left_edges: List[Edge], right_edges: List[Edge],
                        max_singular_values: Optional[int] = None,
                        max_truncation_err: Optional[float] = None,
                        left_name: Optional[str] = None, middle_name: Optional[str] = None,
                        right_name: Optional[str] = None,
                        left_edge_name: Optional[str] = None,
                        right_edge_name: Optional[str] = None):
    node.reorder_edges(list(left_edges) + list(right_edges))
    u, s, vh, trun_vals = _svd_decomposition(node.tensor, len(left_edges),
                                             max_singular_values, max_truncation_err)
    left_node = Node(u, name=left_name)
    singular_values_node = Node(np.diag(s), name=middle_name)
    right_node = Node(vh, name=right_name)
    for i, e in enumerate(list(left_edges)):
        _reattach(left_node, i, e, node, i)
    for i, e in enumerate(list(right_edges)):
        _reattach(right_node, i + 1, e, node, i + len(left_edges))
    connect(left_node.edges[-1], singular_values_node.edges[0], name=left_edge_name)
    connect(singular_values_node.edges[1], right_node.edges[0], name=right_edge_name)
    return left_node, singular_values_node, right_node, trun_vals


def split_node(node: Node, left_edges: List[Edge], right_edges: List[Edge],
               max_singular_values: Optional[int] = None,
               max_truncation_err: Optional[float] = None,
               left_name: Optional[str] = None, right_name: Optional[str] = None,
               edge_name: Optional[str] = None):
    node.reorder_edges(list(left_edges) + list(right_edges))
    u, s, vh, trun_vals = _svd_decomposition(node.tensor, len(left_edges),
                                             max_singular_values, max_truncation_err)
    sqrt_s = np.sqrt(s)
    u_s = u * sqrt_s
    vh_s = np.reshape(sqrt_s, [-1] + [1] * (vh.ndim - 1)) * vh
    left_node = Node(u_s, name=left_name)
    right_node = Node(vh_s, name=right_name)
    for i, e in enumerate(list(left_edges)):
        _reattach(left_node, i, e, node, i)
    for i, e in enumerate(list(right_edges)):
        _reattach(right_node, i + 1, e, node, i + len(left_edges))
    connect(left_node.edges[-1], right_node.edges[0], name=edge_name)
    return left_node, right_node, trun_vals


def conj(node: Node, name: Optional[str] = None) -> Node:
    return Node(np.conj(node.tensor), name=name)


def set_default_backend(backend: str) -> None:
    if backend != "numpy":
        raise ValueError("The shim only provides the numpy backend.")
