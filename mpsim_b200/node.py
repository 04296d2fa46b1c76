"""Gate container standing in for ``tensornetwork.Node`` on the gate-application path.

The reference passes gates around as ``tn.Node`` objects (``mpsim/gates.py:104-232``,
``mpsim/core.py:21-65``) but only ever reads ``.tensor`` / ``.name`` and the free-edge count
from them.  ``tensornetwork`` is optional here: any object exposing ``.tensor`` with shape
``(d,)*2k`` is accepted wherever a gate is expected (a real ``tn.Node`` included).
"""

from typing import Any, List, Optional

import numpy as np


class _FreeEdge:
    """A dangling edge of a gate node (gates never carry connected edges on this path)."""

    def __init__(self, node: "Node", axis: int) -> None:
        self.node1 = node
        self.node2 = None
        self.axis1 = axis
        self.axis2 = None

    @property
    def dimension(self) -> int:
        return self.node1.tensor.shape[self.axis1]

    def is_dangling(self) -> bool:
        return True


class Node:
    def __init__(self, tensor: Any, name: Optional[str] = None, **_: Any) -> None:
        if hasattr(tensor, "tensor") and not isinstance(tensor, np.ndarray):
            tensor = tensor.tensor
        self.tensor = np.asarray(tensor)
        self.name = name if name is not None else "__unnamed_node__"

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def edges(self) -> List[_FreeEdge]:
        return [_FreeEdge(self, i) for i in range(self.tensor.ndim)]

    def get_all_edges(self) -> List[_FreeEdge]:
        return self.edges

    def get_all_dangling(self) -> List[_FreeEdge]:
        return self.edges

    def get_all_nondangling(self) -> List[_FreeEdge]:
        return []

    def has_nondangling_edge(self) -> bool:
        return False

    def get_edge(self, axis: int) -> _FreeEdge:
        return self.edges[axis]

    def __getitem__(self, axis: int) -> _FreeEdge:
        return self.edges[axis]

    def set_tensor(self, tensor: Any) -> None:
        self.tensor = np.asarray(tensor)

    def copy(self) -> "Node":
        return Node(np.array(self.tensor, copy=True), name=self.name)

    def __str__(self) -> str:
        return self.name


class BondEdge:
    """A leg of the chain seen through the reference's graph accessors (``mpsim/core.py:443-481``
    return ``tn.Edge`` objects): a bond between two neighbouring sites, or the physical (dangling)
    leg of one site.  A compatibility VIEW -- the device store has no graph -- carrying what the
    reference's callers and tests read: ``node1`` / ``node2`` (host copies of the sites, with the
    reference's names), ``is_dangling()``, ``dimension``, and equality of two views of one bond."""

    def __init__(self, owner: Any, key: Any, node1: "Node", node2: Optional["Node"], dimension: int,
                 name: str = "__unnamed_edge__") -> None:
        self._owner, self._key = owner, key
        self.node1, self.node2 = node1, node2
        self.dimension = int(dimension)
        self.name = name

    def is_dangling(self) -> bool:
        return self.node2 is None

    def get_nodes(self) -> List[Optional["Node"]]:
        return [self.node1, self.node2]

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, BondEdge) and self._owner is other._owner and self._key == other._key

    def __hash__(self) -> int:
        return hash((id(self._owner), self._key))


def tensor_of(gate: Any) -> np.ndarray:
    """The array of a gate given as ``Node``, ``tn.Node`` or a plain array."""
    if isinstance(gate, np.ndarray):
        return gate
    if hasattr(gate, "tensor"):
        return np.asarray(gate.tensor)
    return np.asarray(gate)
