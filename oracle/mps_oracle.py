"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Array-level numpy restatement of the reference's MPS gate-application path
(``/root/reference/mpsim/core.py``; all ``file:line`` below are into ``/root/reference``).
The reference expresses these steps through ``tensornetwork==0.2.1`` graph calls; here the
same arithmetic is written directly with ``np.tensordot`` / ``np.linalg.svd`` (which is what
the numpy backend of tensornetwork 0.2.1 executes).

Pinning (see DESIGN.md "Oracle"): this restatement is checked in ``tests/test_oracle.py``
against golden vectors produced by running the UNMODIFIED reference itself in the build
container on top of ``oracle/tn_shim`` (script: ``tests/golden/make_golden.py``), and
against the reference's own known-answer tests (ported in ``tests/test_oracle.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product package ``mpsim_b200`` never does.

Site tensor convention used here (and by the CUDA store): ``A[i]`` has shape
``(chi_left, d, chi_right)``; chain ends carry a bond of dimension 1.  The reference keeps
no fixed axis order (it tracks the graph); every quantity compared in tests is
independent of that choice.
"""

from typing import Any, Dict, List, Optional, Sequence

import numpy as np


class OracleMPS:
    """Restatement of ``mpsim.MPS`` (``mpsim/core.py:159-1423``) for the hot path."""

    def __init__(self, nqudits: int, qudit_dimension: int = 2, dtype: Any = None,
                 track_norms: bool = False) -> None:
        # core.py:184-187
        if nqudits < 2:
            raise ValueError(f"Number of qudits must be greater than 2 but is {nqudits}.")
        self.nqudits = nqudits
        self.d = qudit_dimension
        # core.py:190-218: |0...0>, complex64.  ``dtype=None`` follows numpy promotion exactly
        # like the reference (complex64 sites become complex128 once a float64 CNOT/SWAP or a
        # complex128 unitary touches them, gates.py:195-212); a fixed dtype forces it.
        self.dtype = dtype
        base = np.complex64 if dtype is None else dtype
        site = np.zeros((1, qudit_dimension, 1), dtype=base)
        site[0, 0, 0] = 1.0
        self.sites: List[np.ndarray] = [site.copy() for _ in range(nqudits)]
        # core.py:235-242
        mbd = [qudit_dimension ** (i + 1) for i in range(nqudits // 2)]
        mbd += list(reversed(mbd))
        if nqudits % 2 == 0:
            mbd.remove(qudit_dimension ** (nqudits // 2))
        self._max_bond_dimensions = mbd
        self.track_norms = track_norms
        self._norms: List[float] = []          # core.py:243, 1160-1161
        #: one entry per adjacent application: dict(index, chi (l,m,r), k, s_kept, s_trunc, left)
        self.trace: List[Dict[str, Any]] = []

    # ------------------------------------------------------------------ bond bookkeeping
    def bond_dimension_of(self, i: int) -> int:           # core.py:340-361
        if i >= self.nqudits:
            raise ValueError(f"Index should be less than {self.nqudits} but is {i}.")
        return self.sites[i].shape[2]

    def bond_dimensions(self) -> List[int]:                # core.py:363-365
        return [self.bond_dimension_of(i) for i in range(self.nqudits - 1)]

    def max_bond_dimension_of(self, i: int) -> int:        # core.py:367-382
        if i >= self.nqudits:
            raise ValueError("Edge index out of range.")
        return self._max_bond_dimensions[i]

    def max_bond_dimensions(self) -> List[int]:            # core.py:384-386
        return self._max_bond_dimensions

    def _cast(self, x: np.ndarray) -> np.ndarray:
        return x if self.dtype is None else x.astype(self.dtype)

    # ------------------------------------------------------------------ contractions
    def wavefunction(self) -> np.ndarray:                  # core.py:483-500
        fin = self.sites[0]
        fin = fin.reshape(fin.shape[1], fin.shape[2])       # left bond of site 0 has dim 1
        for a in self.sites[1:]:
            fin = np.tensordot(fin, a, [[fin.ndim - 1], [0]])
            fin = fin.reshape(-1, a.shape[2])
        return fin.reshape(self.d ** self.nqudits)

    def inner_product(self, other: "OracleMPS") -> complex:   # core.py:507-561, <self|other>
        if other.nqudits != self.nqudits or other.d != self.d:
            raise ValueError("Cannot compute inner product: shape mismatch.")
        # core.py:543-546 conjugates the copy of ``other``'s nodes (sic) -- for
        # other is self (the only use on the hot path, norm()) this is <psi|psi>.
        env = np.ones((1, 1), dtype=np.result_type(self.sites[0], other.sites[0]))
        for a, b in zip(self.sites, other.sites):
            # env[x, y] (x: self bond, y: other bond)
            tmp = np.tensordot(env, a, [[0], [0]])                    # [y, p, x']
            env = np.tensordot(tmp, np.conj(b), [[0, 1], [0, 1]])     # [x', y']
        return complex(env.reshape(()))

    def norm(self) -> float:                               # core.py:563-565
        return float(np.sqrt(self.inner_product(self).real))

    def renormalize(self, to_norm: float = 1.0) -> None:   # core.py:567-594
        if to_norm < 0.0:
            raise ValueError(f"Arg to_norm must be positive but is {to_norm}")
        if np.isclose(to_norm, 0.0, atol=1e-15):
            raise ValueError(f"Arg to_norm = {to_norm} is too close to numerical zero.")
        if np.isclose(self.norm(), 0.0, atol=1e-15):
            raise ValueError("Norm of MPS is numerically zero, cannot renormalize.")
        norm = self.norm()
        f = (to_norm / norm) ** (1 / self.nqudits)
        self.sites = [self._cast(f * a) for a in self.sites]

    # ------------------------------------------------------------------ one-qudit gates
    def apply_one_qudit_gate(self, gate: np.ndarray, i: int, **kwargs: Any) -> None:
        """core.py:753-845.  ``gate[o, p]``: axis 1 contracts with the site (core.py:773-775)."""
        gate = np.asarray(gate)
        if i not in range(self.nqudits):                    # core.py:785-789
            raise ValueError(f"Input tensor index={i} is out of bounds.")
        if gate.ndim != 2:                                  # core.py:791-796
            raise ValueError("Single qudit gate must have two free edges and zero connected edges.")
        if gate.shape[0] != gate.shape[1]:                  # core.py:798-799
            raise ValueError("Gate edge dimensions must be equal.")
        if gate.shape[0] != self.d:                         # core.py:801-805
            raise ValueError("Gate edges have the wrong dimension.")
        renorm = kwargs.get("renormalize_after_non_unitary") is not False   # core.py:808-813
        ortho = kwargs.get("ortho_after_non_unitary") is not False
        unitary = _is_unitary(gate)
        if not unitary and renorm:                          # core.py:816-817
            norm = self.norm()
        a = self.sites[i]
        # core.py:820-826:  A'[l, o, r] = sum_p g[o, p] A[l, p, r]
        self.sites[i] = self._cast(np.transpose(np.tensordot(gate, a, [[1], [1]]), (1, 0, 2)))
        if not unitary and ortho:                           # core.py:829-841
            if i == 0:
                self.orthonormalize_right_edge_of(i)
            elif i == self.nqudits - 1:
                self.orthonormalize_left_edge_of(i)
            else:
                self.orthonormalize_right_edge_of(i)
                self.orthonormalize_left_edge_of(i)
        if not unitary and renorm:                          # core.py:844-845
            self.renormalize(norm)

    def apply_one_qudit_gate_to_all(self, gate: np.ndarray) -> None:   # core.py:941-948
        for i in range(self.nqudits):
            self.apply_one_qudit_gate(gate, i)

    def orthonormalize_right_edge_of(self, i: int, threshold: float = 1e-8) -> None:
        """core.py:847-892: SVD site i as (phys,left | right), push S.Vh into site i+1."""
        if not 0 <= i < self.nqudits - 1:
            raise ValueError("Invalid edge index.")
        a = self.sites[i]
        cl, d, cr = a.shape
        u, s, vh, _ = _svd_trunc(a.reshape(cl * d, cr), None, threshold * self.norm())
        k = s.shape[0]
        self.sites[i] = self._cast(u.reshape(cl, d, k))
        sv = np.diag(s) @ vh
        self.sites[i + 1] = self._cast(np.tensordot(sv, self.sites[i + 1], [[1], [0]]))

    def orthonormalize_left_edge_of(self, i: int, threshold: float = 1e-8) -> None:
        """core.py:894-939: SVD site i as (left | phys,right), push U.S into site i-1."""
        if not 0 < i <= self.nqudits - 1:
            raise ValueError("Invalid edge index.")
        a = self.sites[i]
        cl, d, cr = a.shape
        u, s, vh, _ = _svd_trunc(a.reshape(cl, d * cr), None, threshold * self.norm())
        k = s.shape[0]
        self.sites[i] = self._cast(vh.reshape(k, d, cr))
        us = u @ np.diag(s)
        self.sites[i - 1] = self._cast(np.tensordot(self.sites[i - 1], us, [[2], [0]]))

    # ------------------------------------------------------------------ two-qudit gates
    def apply_two_qudit_gate(self, gate: np.ndarray, i: int, j: int, **kwargs: Any) -> None:
        """core.py:950-1161.  ``gate[o1, o2, p, q]`` with the edge convention of core.py:986-992."""
        gate = np.asarray(gate)
        n = self.nqudits
        if i not in range(n) or j not in range(n):           # core.py:1003-1008
            raise ValueError(f"Input tensor indices={(i, j)} are out of bounds.")
        if i == j:                                           # core.py:1010-1011
            raise ValueError("Node indices cannot be identical.")
        if gate.ndim != 4:                                   # core.py:1013-1018
            raise ValueError("Two qubit gate must have four free edges and zero connected edges.")
        if len(set(gate.shape)) != 1:                        # core.py:1020-1022
            raise ValueError("All gate edges must have the same dimension.")
        if gate.shape[0] != self.d:                          # core.py:1024-1028
            raise ValueError("Gate edges have the wrong dimension.")

        if j < i:                                            # core.py:1031-1033
            gate = np.transpose(gate, (1, 0, 3, 2))
            i, j = j, i

        invert_swap_network = False                          # core.py:1036-1043
        if i < j - 1:
            invert_swap_network = True
            original_i = i
            self.move_node_from_left_to_right(i, j - 1, **kwargs)
            i = j - 1

        a, b = self.sites[i], self.sites[j]
        cl, d, cm = a.shape
        cr = b.shape[2]
        # core.py:1060-1062: theta[l, p, q, r] = sum_m A[l, p, m] B[m, q, r]
        theta = np.tensordot(a, b, [[2], [0]])
        # core.py:1065-1068: theta'[o1, o2, l, r] = sum_{p,q} G[o1, o2, p, q] theta[l, p, q, r]
        theta = np.tensordot(gate, theta, [[2, 3], [1, 2]])
        # core.py:1095-1102: rows = (gate edge 0, left bond), cols = (gate edge 1, right bond)
        mat = np.transpose(theta, (0, 2, 1, 3)).reshape(d * cl, d * cr)

        # core.py:1105-1130
        keep_left_canonical = kwargs["keep_left_canonical"] if "keep_left_canonical" in kwargs else True
        if "fraction" in kwargs and "maxsvals" in kwargs:
            raise ValueError("Only one of (fraction, maxsvals) can be provided as kwargs.")
        if "fraction" in kwargs:
            fraction = kwargs.get("fraction")
            if not (0 <= fraction <= 1):
                raise ValueError("Keyword fraction must be between 0 and 1 but is", fraction)
            maxsvals = int(round(fraction * self.max_bond_dimension_of(min(i, j))))
        else:
            maxsvals = None
        if "maxsvals" in kwargs:
            maxsvals = int(kwargs.get("maxsvals"))

        # core.py:1132-1137 -> tensornetwork split_node_full_svd -> np.linalg.svd, keep first k
        u, s, vh, s_rest = _svd_trunc(mat, maxsvals, None)
        k = s.shape[0]
        self.trace.append(dict(index=i, chi=(cl, cm, cr), k=k, left=bool(keep_left_canonical),
                               s_kept=np.real(s).astype(np.float64).copy(),
                               s_trunc=np.real(s_rest).astype(np.float64).copy()))
        # core.py:1140-1145 (S is a dense diag matrix in the reference; same product)
        if keep_left_canonical:
            new_left, new_right = u, np.diag(s) @ vh
        else:
            new_left, new_right = u @ np.diag(s), vh
        # rows of ``mat`` are (o1, l): store as [l, o1, k]
        self.sites[i] = self._cast(np.transpose(new_left.reshape(d, cl, k), (1, 0, 2)))
        self.sites[j] = self._cast(new_right.reshape(k, d, cr))

        if invert_swap_network:                              # core.py:1155-1158
            self.move_node_from_right_to_left(i, original_i, **kwargs)

        if self.track_norms:                                 # core.py:1160-1161
            self._norms.append(self.norm())

    def move_node_from_left_to_right(self, cur: int, fin: int, **kwargs: Any) -> None:
        """core.py:1163-1190."""
        if cur > fin:
            raise ValueError("current_node_index should be smaller than final_node_index.")
        if cur < 0:
            raise ValueError("current_node_index out of range.")
        if fin >= self.nqudits:
            raise ValueError("final_node_index out of range.")
        while cur < fin:
            self.swap(cur, cur + 1, **kwargs)
            cur += 1

    def move_node_from_right_to_left(self, cur: int, fin: int, **kwargs: Any) -> None:
        """core.py:1192-1219."""
        if cur < fin:
            raise ValueError("current_node_index should be larger than final_node_index.")
        if cur > self.nqudits:
            raise ValueError("current_node_index out of range.")
        if fin < 0:
            raise ValueError("final_node_index out of range.")
        while cur > fin:
            self.swap(cur - 1, cur, **kwargs)
            cur -= 1

    # ------------------------------------------------------------------ conveniences
    def x(self, i: int) -> None:                           # core.py:1279-1290
        if i == -1:
            self.apply_one_qudit_gate_to_all(XGATE)
        else:
            self.apply_one_qudit_gate(XGATE, i)

    def h(self, i: int) -> None:                           # core.py:1292-1303
        if i == -1:
            self.apply_one_qudit_gate_to_all(HGATE)
        else:
            self.apply_one_qudit_gate(HGATE, i)

    def cnot(self, a: int, b: int, **kwargs: Any) -> None:    # core.py:1324-1328
        self.apply_two_qudit_gate(CNOT, a, b, **kwargs)

    def swap(self, a: int, b: int, **kwargs: Any) -> None:    # core.py:1376-1380
        if b < a:
            a, b = b, a
        self.apply_two_qudit_gate(SWAP, a, b, **kwargs)

    def apply(self, operations: Sequence[Any], **kwargs: Any) -> None:
        """core.py:1221-1276; an operation here is ``(tensor, qudit_indices)``."""
        for tensor, indices in operations:
            tensor = np.asarray(tensor)
            if isinstance(indices, int):
                indices = (indices,)
            if tensor.shape != tuple([self.d] * 2 * len(indices)):     # core.py:113-129, 1260
                raise ValueError("Input MPS Operation is not valid.")
            if len(indices) == 1:
                self.apply_one_qudit_gate(tensor, *indices, **kwargs)
            elif len(indices) == 2:
                self.apply_two_qudit_gate(tensor, *indices, **kwargs)
            else:                                                        # core.py:1271-1276
                raise ValueError("Only one-qudit and two-qudit gates are supported.")

    # ------------------------------------------------------------------ SURVEY.md 8(f) rows 2-3
    @staticmethod
    def from_wavefunction(wavefunction: Any, nqudits: int, qudit_dimension: int = 2,
                          dtype: Any = None) -> "OracleMPS":      # core.py:245-328
        if not isinstance(wavefunction, (list, tuple, np.ndarray)):
            raise TypeError("Invalid type for wavefunction.")
        wavefunction = np.array(wavefunction)
        if len(wavefunction.shape) != 1:
            raise ValueError("Invalid shape for wavefunction. Should be a vector.")
        if nqudits < 2:
            raise ValueError("At least two qudits are required.")
        if wavefunction.size != qudit_dimension ** nqudits:
            raise ValueError("Mismatch between wavefunction, qudit_dimension, and nqudits.")
        d = qudit_dimension
        mps = OracleMPS(nqudits, d, dtype)
        # core.py:301-321: tn.split_node across every cut, no truncation; split_node hands
        # sqrt(S) to both sides (tensornetwork 0.2.1 network_operations.split_node)
        rest = wavefunction.reshape(1, -1)
        sites = []
        for _ in range(nqudits - 1):
            chi = rest.shape[0]
            mat = rest.reshape(chi * d, -1)
            u, s, vh = np.linalg.svd(mat, full_matrices=False)
            sq = np.sqrt(s).astype(mat.dtype)
            sites.append((u * sq[None, :]).reshape(chi, d, s.size))
            rest = sq[:, None] * vh
        sites.append(rest.reshape(rest.shape[0], d, 1))
        mps.sites = [mps._cast(a) for a in sites]
        return mps

    def reduced_density_matrix(self, node_indices: Any) -> np.ndarray:      # core.py:596-652
        try:
            node_indices = iter(node_indices)
        except TypeError:
            node_indices = [node_indices]
        node_indices = tuple(node_indices)
        if len(set(node_indices)) < len(node_indices):
            raise ValueError("Node indices contains duplicates.")
        if min(node_indices) < 0 or max(node_indices) > self.nqudits - 1:
            raise IndexError("One or more invalid node indices.")
        # env axes: open ket / bra legs in site order (labels), then ket bond, bra bond
        env = np.ones((1, 1), dtype=self.sites[0].dtype)
        labels: List[Any] = []
        for i, a in enumerate(self.sites):
            t = np.tensordot(env, a, [[env.ndim - 2], [0]])               # [..., a', p, b]
            if i in node_indices:
                env = np.tensordot(t, np.conj(a), [[t.ndim - 3], [0]])    # [..., p, b, p', b']
                nd = env.ndim
                env = np.moveaxis(env, nd - 2, nd - 3)                    # [..., p, p', b, b']
                labels += [("k", i), ("b", i)]
            else:
                env = np.tensordot(t, np.conj(a), [[t.ndim - 3, t.ndim - 2], [0, 1]])   # [..., b, b']
        env = env.reshape(env.shape[:-2])
        order = [labels.index(("k", i)) for i in node_indices] + [labels.index(("b", i)) for i in node_indices]
        env = np.transpose(env, order)
        n = len(node_indices)
        return env.reshape(self.d ** n, self.d ** n)

    def sample_once(self, as_string: bool = False) -> Any:                 # core.py:654-682
        # (sic) the reference draws every site from ITS OWN marginal of the unconditioned state
        # (core.py:665 uses ``self``, the conditioned ``copy`` is never read): the same draws here
        string = []
        states = list(range(self.d))
        for i in range(self.nqudits):
            qubit = self.reduced_density_matrix(i).diagonal().real
            string.append(np.random.choice(states, size=1, p=qubit)[0])
        if as_string:
            return "".join(str(bit) for bit in string)
        return string

    def sample(self, nsamples: int, as_hist: bool = False, as_string: bool = False) -> Any:   # core.py:684-721
        if not isinstance(nsamples, int):
            raise ValueError(f"Arg nsamples should be an int but is a {type(nsamples)}.")
        if nsamples <= 0:
            raise ValueError(f"Arg nsamples should be positive but is {nsamples}.")
        if as_hist:
            as_string = True
        raw = [self.sample_once(as_string) for _ in range(nsamples)]
        if as_hist:
            hist: Dict[Any, int] = {}
            for bitstring in raw:
                hist[bitstring] = hist.get(bitstring, 0) + 1
            return hist
        return raw

    def expectation(self, tensor: np.ndarray, indices: Sequence[int]) -> float:   # core.py:723-751
        k = len(indices)
        mat = np.asarray(tensor).reshape(self.d ** k, self.d ** k)
        if not np.allclose(mat, mat.conj().T):
            raise ValueError("Observable is not Hermitian.")
        cp = self.copy()
        if k == 1:
            cp.apply_one_qudit_gate(np.asarray(tensor), indices[0])
        else:
            cp.apply_two_qudit_gate(np.asarray(tensor), indices[0], indices[1])
        return self.inner_product(cp).real

    def copy(self) -> "OracleMPS":                          # core.py:1382-1384, 1420-1423
        new = OracleMPS(self.nqudits, self.d, self.dtype, self.track_norms)
        new.sites = [a.copy() for a in self.sites]
        return new


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _svd_trunc(mat: np.ndarray, max_singular_values: Optional[int],
               max_truncation_err: Optional[float]):
    """tensornetwork 0.2.1 ``backends/numpy/decompositions.py: svd_decomposition`` restated
    (call sites core.py:877, 924, 1132): thin SVD, keep ``min(max_singular_values,
    #values whose tail 2-norm exceeds max_truncation_err)``; no tolerance cut otherwise."""
    u, s, vh = np.linalg.svd(mat, full_matrices=False)
    if max_singular_values is None:
        max_singular_values = s.size
    if max_truncation_err is not None:
        trunc_errs = np.sqrt(np.cumsum(np.square(s[::-1])))
        n_err = int(np.count_nonzero(trunc_errs > max_truncation_err))
    else:
        n_err = max_singular_values
    keep = min(max_singular_values, n_err)
    s = s.astype(mat.dtype)
    return u[:, :keep], s[:keep], vh[:keep, :], s[keep:]


def _is_unitary(gate: np.ndarray) -> bool:                 # gates.py:15-33
    gate = np.asarray(gate)
    if gate.ndim > 2:
        dim = int(np.sqrt(gate.size))
        gate = gate.reshape(dim, dim)
    return bool(np.allclose(gate.conj().T @ gate, np.identity(gate.shape[0]), atol=1e-5))


# gates.py:104-232 (dtypes as in the reference: 1-qubit gates and cphase complex64,
# cnot / swap float64)
HGATE = (1 / np.sqrt(2) * np.array([[1.0, 1.0], [1.0, -1.0]], dtype=np.complex64))
IGATE = np.array([[1.0, 0.0], [0.0, 1.0]], dtype=np.complex64)
XGATE = np.array([[0.0, 1.0], [1.0, 0.0]], dtype=np.complex64)
YGATE = np.array([[0.0, -1j], [1j, 0.0]], dtype=np.complex64)
ZGATE = np.array([[1.0, 0.0], [0.0, -1.0]], dtype=np.complex64)
CNOT = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0], [0, 0, 1.0, 0]]).reshape(2, 2, 2, 2)
SWAP = np.array([[1.0, 0, 0, 0], [0, 0, 1.0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]]).reshape(2, 2, 2, 2)


def cphase(exp: float) -> np.ndarray:                      # gates.py:220-232
    m = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0],
                  [0, 0, 0, np.exp(1j * 2 * np.pi * exp)]], dtype=np.complex64)
    return m.reshape(2, 2, 2, 2)


def haar_random_unitary(nqudits: int = 2, qudit_dimension: int = 2,
                        seed: Optional[int] = None, rng: Any = None) -> np.ndarray:
    """gates.py:248-286 (Mezzadri, arXiv:math-ph/0609050): QR of a complex Ginibre matrix
    with the phases of diag(R) divided out.  Returns shape ``(d,)*2*nqudits``."""
    if rng is None:
        rng = np.random.RandomState(seed)
    units = np.array([1, 1j])
    shape = (qudit_dimension ** nqudits, qudit_dimension ** nqudits)
    mat = np.sum(rng.randn(*(shape + (2,))) * units, axis=-1) / np.sqrt(2)
    qmat, rmat = np.linalg.qr(mat)
    diag = np.diag(rmat).copy()
    diag /= np.abs(diag)
    return (qmat * diag).reshape([qudit_dimension] * 2 * nqudits)
