"""Device-resident site-tensor store and plan execution.

One ``DeviceChain`` holds ``nbatch`` independent MPS of identical shape (the single ``MPS``
object is the ``nbatch == 1`` case).  All site tensors live in ONE complex64 slab in HBM:

    slab[b, off_i : off_i + cap_i * d * cap_{i+1}]   holds site i of batch member b,
    dense row-major [chi_i][d][chi_{i+1}] in the first chi_i*d*chi_{i+1} entries,

where ``cap`` are bond capacities (the running maximum of each bond dimension, known
statically from the plan, padded to chi = maxsvals in the bulk).  Gates are applied in
place: the theta kernel writes to workspace before the split kernel overwrites the two
sites.  Replaces the Python list of ``tn.Node`` of ``mpsim/core.py:190-243``.

PyTorch is used only to own device memory and streams; every computation is a call into
``libmpsim_b200.so`` (``include/mpsim_b200.h``).
"""
import ctypes
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from mpsim_b200 import _lib
from mpsim_b200.planner import Plan

MAX_JOBS_PER_CALL = 65535
#: run() checks the spread of the site exponents after this many brickwork layers' worth of
#: two-site applications (n // 2 each), and rebalances a chain whose sites are further apart than
#: 2**REBALANCE_SPREAD_LOG2 (see DeviceChain.rebalance)
REBALANCE_EVERY_LAYERS = 32
REBALANCE_SPREAD_LOG2 = 32


def _torch():
    import torch
    return torch


def _on_device(fn):
    """Run a DeviceChain method with the chain's device current: the C-ABI launches on the stream it
    is handed, and ``_lib.stream_ptr`` / the library's per-device stream pool follow the current device."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        torch = _torch()
        if torch.cuda.current_device() == self.device.index:      # the usual case: no context switch
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapper


class SVDNotConverged(RuntimeWarning):
    """A Jacobi solve hit its sweep limit while still rotating (status word 1 of the C-ABI's ``info``):
    the factors are still an exact orthogonal projection of theta, but the kept subspace and singular
    values may differ from the reference's LAPACK SVD by more than the parity bound."""


class _Staging:
    """One pinned host buffer and its device twin.  A compiled plan's descriptor tables and gate table
    live in ONE such pair and go to the device in ONE asynchronous copy: the first version allocated
    and pinned a fresh gate buffer and made two blocking pageable uploads per compile, which serialised
    host and device on every ``apply_two_qudit_gate`` call."""

    def __init__(self, nbytes: int, device: Any, host=None, dev=None) -> None:
        torch = _torch()
        self.cap = int(nbytes)
        # (host, dev): views of a larger pair -- the slots of a chain's ring share ONE pinned allocation
        self.host = torch.zeros(self.cap, dtype=torch.uint8).pin_memory() if host is None else host
        self.dev = torch.empty(self.cap, dtype=torch.uint8, device=device) if dev is None else dev
        self.event = None            # recorded after the last run() that read this slot

    def wait(self) -> None:
        if self.event is not None:
            self.event.synchronize()
            self.event = None


#: bond capacity a single chain grows to at once (bounded by d**min(i, n - i)), see ensure_caps
_SINGLE_CHAIN_MIN_CAP = 64

#: small plans (gate-by-gate use of the API) draw their staging from a ring of reusable slots
_RING_SLOTS = 32
_RING_SLOT_BYTES = 64 * 1024


class CompiledPlan:
    """A plan bound to a chain's buffers: device descriptor tables + the launch list."""

    def __init__(self) -> None:
        self.layout_gen = -1
        self.owner = None
        self.launches: List[Tuple] = []
        self.desc1 = None
        self.desc2 = None
        self.staging = None          # _Staging holding desc1 | desc2 | gates
        self.gates_span = (0, 0)     # byte range of the gate table inside the staging buffers
        self.gates = None            # device view [ngates][nb_g][width] complex64
        self.gates_host = None       # pinned view of the same shape
        self.info = None             # device int32 [napps2 * B][2]
        self.svals = None            # device float32 [napps2][B][maxmn] or None
        self.svals_width = 0
        self.order2: List[int] = []  # sorted position -> index into plan.apps2
        self.workspace_bytes = 0
        self.plan: Optional[Plan] = None
        self.bonds_out: List[int] = []
        self.n_launch_calls = 0
        self.n_rebalance_calls = 0
        self.launch_bonds: List[Tuple[int, ...]] = []


class DeviceChain:
    def __init__(self, nqudits: int, d: int = 2, nbatch: int = 1, device: Any = None,
                 caps: Optional[Sequence[int]] = None) -> None:
        torch = _torch()
        _lib.load(require_device=True)
        self.n = int(nqudits)
        self.d = int(d)
        self.B = int(nbatch)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.bonds = [1] * (self.n + 1)
        self.caps = [1] * (self.n + 1)
        if caps is not None:
            self.caps = [max(1, int(c)) for c in caps]
            self.caps[0] = self.caps[-1] = 1
        self._layout_gen = 0
        self._status_flags = []          # device scalars: max SVD status of every run() since the last check
        self._apps_since_rebalance = 0
        self._rebalance_bufs = None
        self._ring: List[_Staging] = []
        self._ring_next = 0
        with torch.cuda.device(self.device):
            self._alloc(self.caps)
            self._workspace = None
            self.reset()

    # ------------------------------------------------------------------ layout
    def _offsets(self, caps: Sequence[int]) -> Tuple[List[int], int]:
        offs, total = [], 0
        for i in range(self.n):
            offs.append(total)
            elems = caps[i] * self.d * caps[i + 1]
            total += (elems + 15) // 16 * 16          # 128-byte aligned slots
        return offs, max(total, 16)

    def _alloc(self, caps: Sequence[int]) -> None:
        torch = _torch()
        self.caps = list(caps)
        self.offs, self.total = self._offsets(self.caps)
        self.slab = torch.zeros((self.B, self.total), dtype=torch.complex64, device=self.device)
        self._layout_gen += 1

    @_on_device
    def ensure_caps(self, caps: Sequence[int]) -> None:
        """Grow bond capacities (re-lays the slab out and copies the live tensors).  A single chain
        grows a bond to at least twice its capacity (bounded by what the bond can ever reach,
        d**min(i, n - i)), so that gate-by-gate use re-lays the slab out O(log chi) times instead of
        at every application that widens a bond; batches are sized exactly (their slabs are GBs)."""
        new = [max(a, b) for a, b in zip(self.caps, caps)]
        if new == self.caps:
            return
        if self.B == 1:
            # (every interior bond goes to at least _SINGLE_CHAIN_MIN_CAP on the first growth: a 20-qubit chain
            # at capacity 64 is 1.3 MB, and the first layers of a circuit would otherwise re-lay the slab out
            # -- n slice copies -- every time a gate widens a bond for the first time: 35 times in 95 calls)
            for i in range(1, self.n):
                e = min(i, self.n - i)
                bound = self.d ** e if e < 24 else 1 << 62
                want = max(new[i], _SINGLE_CHAIN_MIN_CAP)
                if new[i] > self.caps[i]:
                    want = max(want, 2 * self.caps[i])
                new[i] = max(new[i], min(want, bound))
        old_slab, old_offs = self.slab, self.offs
        self._alloc(new)
        for i in range(self.n):
            ne = self.bonds[i] * self.d * self.bonds[i + 1]
            if ne:
                self.slab[:, self.offs[i]:self.offs[i] + ne] = old_slab[:, old_offs[i]:old_offs[i] + ne]

    @_on_device
    def reset(self) -> None:
        """|0...0> on every batch member (``mpsim/core.py:190-218``)."""
        self.bonds = [1] * (self.n + 1)
        self._apps_since_rebalance = 0
        self.slab.zero_()
        idx = _torch().tensor(self.offs, device=self.device, dtype=_torch().long)
        self.slab[:, idx] = 1.0

    def site_elems(self, i: int) -> int:
        return self.bonds[i] * self.d * self.bonds[i + 1]

    def site_ptr(self, i: int) -> int:
        return self.slab.data_ptr() + self.offs[i] * 8

    def site_view(self, i: int, b: int = 0):
        ne = self.site_elems(i)
        return self.slab[b, self.offs[i]:self.offs[i] + ne].view(self.bonds[i], self.d, self.bonds[i + 1])

    @_on_device
    def set_site(self, i: int, tensor, b: Optional[int] = None) -> None:
        """Overwrite site i (all batch members, or one) with a [chiL][d][chiR] tensor."""
        torch = _torch()
        t = torch.as_tensor(tensor).to(device=self.device, dtype=torch.complex64).contiguous()
        cl, d, cr = t.shape
        assert d == self.d
        bonds = list(self.bonds)
        bonds[i], bonds[i + 1] = cl, cr
        self.ensure_caps(bonds)
        self.bonds = bonds
        ne = cl * d * cr
        if b is None:
            self.slab[:, self.offs[i]:self.offs[i] + ne] = t.reshape(1, -1)
        else:
            self.slab[b, self.offs[i]:self.offs[i] + ne] = t.reshape(-1)

    def clone(self) -> "DeviceChain":
        new = DeviceChain.__new__(DeviceChain)
        new.n, new.d, new.B, new.device = self.n, self.d, self.B, self.device
        new.bonds, new.caps = list(self.bonds), list(self.caps)
        new.offs, new.total = list(self.offs), self.total
        new.slab = self.slab.clone()
        new._workspace = None
        new._layout_gen = 0
        new._status_flags = list(self._status_flags)
        new._apps_since_rebalance, new._rebalance_bufs = self._apps_since_rebalance, None
        new._ring, new._ring_next = [], 0
        return new

    @_on_device
    def member(self, b: int) -> "DeviceChain":
        """Batch member ``b`` as a chain of its own (a copy: same layout, one slab row)."""
        new = DeviceChain.__new__(DeviceChain)
        new.n, new.d, new.B, new.device = self.n, self.d, 1, self.device
        new.bonds, new.caps = list(self.bonds), list(self.caps)
        new.offs, new.total = list(self.offs), self.total
        new.slab = self.slab[int(b):int(b) + 1].clone()
        new._workspace = None
        new._layout_gen = 0
        new._status_flags = []
        new._apps_since_rebalance, new._rebalance_bufs = self._apps_since_rebalance, None
        new._ring, new._ring_next = [], 0
        return new

    def _site_refs(self) -> np.ndarray:
        refs = np.zeros(self.n, dtype=_lib.SITE_REF)
        for i in range(self.n):
            refs[i] = (self.site_ptr(i), self.total, self.bonds[i], self.bonds[i + 1])
        return refs

    def _staging(self, nbytes: int) -> _Staging:
        """Staging for a plan: a slot of the ring for small plans (waits for the slot's previous user
        only when 32 later plans are already in flight), a buffer of its own for a large one."""
        if nbytes > _RING_SLOT_BYTES:
            return _Staging(nbytes, self.device)
        if not self._ring:
            torch = _torch()
            host = torch.zeros(_RING_SLOTS * _RING_SLOT_BYTES, dtype=torch.uint8).pin_memory()
            dev = torch.empty(_RING_SLOTS * _RING_SLOT_BYTES, dtype=torch.uint8, device=self.device)
            self._ring = [_Staging(_RING_SLOT_BYTES, self.device, host[i * _RING_SLOT_BYTES:(i + 1) * _RING_SLOT_BYTES],
                                   dev[i * _RING_SLOT_BYTES:(i + 1) * _RING_SLOT_BYTES]) for i in range(_RING_SLOTS)]
        slot = self._ring[self._ring_next]
        self._ring_next = (self._ring_next + 1) % _RING_SLOTS
        slot.wait()
        return slot

    def workspace(self, nbytes: int):
        torch = _torch()
        nbytes = max(int(nbytes), 256)
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._workspace

    # ------------------------------------------------------------------ plan compilation
    @_on_device
    def compile(self, plan: Plan, per_batch_gates: bool = False, record_svals: bool = False,
                transient: bool = False) -> CompiledPlan:
        """Bind ``plan`` (made against the chain's current bonds) to this chain's buffers.
        ``transient``: the plan is run once, right away (gate-by-gate use of the API): its staging may
        come from the chain's ring of reusable slots instead of a buffer of its own."""
        torch = _torch()
        lib = _lib.load(require_device=True)
        assert plan.bonds_in == self.bonds, "plan was made for different bond dimensions"
        self.ensure_caps(plan.caps)
        cp = CompiledPlan()
        cp.plan = plan
        cp.bonds_out = list(plan.bonds)
        d, B = self.d, self.B
        width = d ** 4
        nb_g = B if per_batch_gates else 1
        ng = len(plan.gates)
        layers = plan.layers()
        n1, n2 = len(plan.apps1), len(plan.apps2)
        # one staging pair: [desc1 | desc2 | gates], every part 128-byte aligned
        b1 = (max(n1, 1) * _lib.GATE1_DESC.itemsize + 127) // 128 * 128
        b2 = (max(n2, 1) * _lib.GATE2_DESC.itemsize + 127) // 128 * 128
        bg = max(ng, 1) * nb_g * width * 8
        st = self._staging(b1 + b2 + bg) if transient else _Staging(b1 + b2 + bg, self.device)
        cp.staging = st
        cp.gates_span = (b1 + b2, b1 + b2 + bg)
        shape = (max(ng, 1), nb_g, width)
        cp.gates = st.dev[b1 + b2: b1 + b2 + bg].view(torch.complex64).view(shape)
        cp.gates_host = st.host[b1 + b2: b1 + b2 + bg].view(torch.complex64).view(shape)
        gbase = cp.gates.data_ptr()
        gstride = nb_g * width * 8
        bs_gate = width if per_batch_gates else 0
        d1 = st.host[:max(n1, 1) * _lib.GATE1_DESC.itemsize].numpy().view(_lib.GATE1_DESC)
        d2 = st.host[b1: b1 + max(n2, 1) * _lib.GATE2_DESC.itemsize].numpy().view(_lib.GATE2_DESC)
        d1[:] = np.zeros(1, dtype=_lib.GATE1_DESC)[0]
        d2[:] = np.zeros(1, dtype=_lib.GATE2_DESC)[0]
        maxmn = max([min(d * a.chiL, d * a.chiR) for a in plan.apps2] + [1])
        if record_svals and n2:
            cp.svals = torch.zeros((n2, B, maxmn), dtype=torch.float32, device=self.device)
            cp.svals_width = maxmn
        cp.info = torch.zeros((max(n2, 1) * B, 2), dtype=torch.int32, device=self.device)
        # bonds evolve layer by layer; replay them to know each 1q application's shape
        bonds = list(plan.bonds_in)
        p1 = p2 = 0
        ws_need = 0
        launches = []
        launch_bonds: List[Tuple[int, ...]] = []      # bond dimensions once launch i (its whole layer) has run
        for layer in layers:
            if layer["one"]:
                start = p1
                max_elems = 1
                for idx in layer["one"]:
                    a = plan.apps1[idx]
                    ptr = self.site_ptr(a.site)
                    d1[p1] = (ptr, ptr, gbase + a.gate_index * gstride, self.total, self.total, bs_gate,
                              bonds[a.site], bonds[a.site + 1])
                    max_elems = max(max_elems, bonds[a.site] * d * bonds[a.site + 1])
                    p1 += 1
                launches.append(("g1", start, p1 - start, max_elems))
            layer_groups = []
            for key, idxs in layer["two"].items():
                chiL, chiM, chiR, k, lc = key
                start = p2
                for idx in idxs:
                    a = plan.apps2[idx]
                    pl, pr = self.site_ptr(a.site), self.site_ptr(a.site + 1)
                    sv = cp.svals.data_ptr() + p2 * B * maxmn * 4 if cp.svals is not None else 0
                    d2[p2] = (pl, pr, pl, pr, gbase + a.gate_index * gstride, sv,
                              self.total, self.total, self.total, self.total, bs_gate, maxmn)
                    cp.order2.append(idx)
                    p2 += 1
                count = p2 - start
                per_call = max(1, MAX_JOBS_PER_CALL // B)
                for c0 in range(0, count, per_call):
                    c = min(per_call, count - c0)
                    layer_groups.append(("g2", start + c0, c, chiL, chiM, chiR, k, int(lc)))
                    ws_need = max(ws_need, lib.mpsb_gate2_workspace_bytes(c, B, d, chiL, chiM, chiR, k))
            # A layer with several shape classes goes down as ONE mpsb_apply_gate2_layer call: its groups run
            # concurrently on library streams.  With block-Jacobi groups (d*chi > 128) the latency-bound one- or
            # two-matrix groups at the chain ends hide behind the main group (configs[2]: +30 %); with
            # single-CTA groups only, the ragged-edge groups fill the last partial wave of the main group's
            # launch (configs[3]: 153.5 k -> 155.2 k applications/s).
            if len(layer_groups) > 1:
                launches.append(("g2layer", layer_groups))
            else:
                launches += layer_groups
            # bonds after this layer
            for key, idxs in layer["two"].items():
                for idx in idxs:
                    a = plan.apps2[idx]
                    bonds[a.site + 1] = a.k
            snapshot = tuple(bonds)
            launch_bonds += [snapshot] * (len(launches) - len(launch_bonds))
        cp.desc1 = st.dev[:b1]
        cp.desc2 = st.dev[b1: b1 + b2]
        # group tables of the layer calls (host arrays; pointers into the uploaded descriptor table)
        p2base, pibase = cp.desc2.data_ptr(), cp.info.data_ptr()
        for li, L in enumerate(launches):
            if L[0] != "g2layer":
                continue
            tab = np.zeros(len(L[1]), dtype=_lib.GATE2_GROUP)
            for gi, (_, off, cnt, chiL, chiM, chiR, k, lc) in enumerate(L[1]):
                tab[gi] = (p2base + off * _lib.GATE2_DESC.itemsize, pibase + off * B * 8, cnt, chiL, chiM, chiR, k, lc)
            ws_need = max(ws_need, lib.mpsb_gate2_layer_workspace_bytes(tab.ctypes.data, len(tab), B, d))
            launches[li] = ("g2layer", tab)
        cp.launches = launches
        cp.launch_bonds = launch_bonds
        cp.workspace_bytes = int(ws_need)
        cp.slab_ptr = self.slab.data_ptr()
        cp.layout_gen = self._layout_gen
        cp.owner = self
        # stage the gates of the plan itself (callers may overwrite cp.gates_host and re-upload)
        tab = plan.gate_table(width)
        cp.gates_host.zero_()
        if ng:
            cp.gates_host[:ng] = torch.from_numpy(tab).unsqueeze(1)
        # descriptor tables (and these gates) to the device: one asynchronous copy, stream ordered
        # before every launch of run()
        st.dev[:b1 + b2].copy_(st.host[:b1 + b2], non_blocking=True)
        return cp

    @_on_device
    def upload_gates(self, cp: CompiledPlan) -> None:
        lo, hi = cp.gates_span
        cp.staging.dev[lo:hi].copy_(cp.staging.host[lo:hi], non_blocking=True)

    @_on_device
    def run(self, cp: CompiledPlan, upload: bool = True) -> None:
        """Launch a compiled plan on the current stream (asynchronous)."""
        lib = _lib.load(require_device=True)
        if cp.owner is not self or cp.layout_gen != self._layout_gen or cp.slab_ptr != self.slab.data_ptr():
            raise RuntimeError("chain buffers were re-laid out since the plan was compiled: compile it again")
        if cp.plan.bonds_in != self.bonds:
            raise RuntimeError("the chain's bond dimensions differ from the ones the plan was compiled for "
                               "(run the plan on the state it was planned against, e.g. after reset())")
        if upload:
            self.upload_gates(cp)
        ws = self.workspace(cp.workspace_bytes)
        st = _lib.stream_ptr()
        d, B = self.d, self.B
        cp.n_rebalance_calls = 0
        p1, p2, pi = cp.desc1.data_ptr(), cp.desc2.data_ptr(), cp.info.data_ptr()
        for li, L in enumerate(cp.launches):
            if L[0] == "g1":
                _, off, cnt, max_elems = L
                _lib.check(lib.mpsb_apply_gate1(p1 + off * _lib.GATE1_DESC.itemsize, cnt, B, d, max_elems, st),
                           "mpsb_apply_gate1")
            elif L[0] == "g2layer":
                tab = L[1]
                _lib.check(lib.mpsb_apply_gate2_layer(tab.ctypes.data, len(tab), B, d, ws.data_ptr(), ws.numel(), st),
                           "mpsb_apply_gate2_layer")
                self._apps_since_rebalance += int(tab["ndesc"].sum())
            else:
                _, off, cnt, chiL, chiM, chiR, k, lc = L
                _lib.check(lib.mpsb_apply_gate2(p2 + off * _lib.GATE2_DESC.itemsize, cnt, B, d, chiL, chiM, chiR,
                                                k, lc, ws.data_ptr(), ws.numel(), pi + off * B * 8, st),
                           "mpsb_apply_gate2")
                self._apps_since_rebalance += cnt
            if self._apps_since_rebalance >= REBALANCE_EVERY_LAYERS * max(1, self.n // 2):
                self.rebalance(bonds=cp.launch_bonds[li])
                cp.n_rebalance_calls += 1
        cp.n_launch_calls = len(cp.launches) + cp.n_rebalance_calls
        self.bonds = list(cp.bonds_out)
        if cp.staging is not None and cp.staging.cap == _RING_SLOT_BYTES:
            ev = _torch().cuda.Event()
            ev.record()
            cp.staging.event = ev        # the ring slot may be rewritten once this run has consumed it
        if cp.plan.apps2:
            self._status_flags.append(cp.info[: len(cp.plan.apps2) * B, 0].max())
            if len(self._status_flags) >= 256:
                self._status_flags = [_torch().stack(self._status_flags).max()]

    def rebalance(self, spread_log2: int = REBALANCE_SPREAD_LOG2, bonds: Optional[Sequence[int]] = None) -> None:
        """Even out the binary exponents of the sites of every chain (``mpsb_rebalance_sites``).

        The reference's alternating left/right-canonical sweeps fix only the PRODUCT of the site
        scales; in a long circuit the scale migrates between sites geometrically (complex128 has
        the exponent range for it, complex64 reaches its limits after ~240 brickwork layers).
        ``run()`` calls this every ``REBALANCE_EVERY_LAYERS`` layers' worth of applications; it
        moves nothing while the exponents are within ``spread_log2`` of each other, and what it
        moves are powers of two that multiply to one, so every contraction of the chain is
        unchanged bit for bit.  ``bonds``: the bond dimensions at this point of a running plan
        (default: the chain's current ones)."""
        lib = _lib.load(require_device=True)
        torch = _torch()
        self._apps_since_rebalance = 0
        bonds = list(self.bonds if bonds is None else bonds)
        refs = np.zeros(self.n, dtype=_lib.SITE_REF)
        for i in range(self.n):
            refs[i] = (self.site_ptr(i), self.total, bonds[i], bonds[i + 1])
        max_elems = max(bonds[i] * self.d * bonds[i + 1] for i in range(self.n))
        if self._rebalance_bufs is None or self._rebalance_bufs.numel() < self.n * self.B:
            self._rebalance_bufs = torch.empty(self.n * self.B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            refs_dev = _lib.to_device_bytes(refs, self.device)
            _lib.check(lib.mpsb_rebalance_sites(refs_dev.data_ptr(), self.n, self.B, self.d, int(spread_log2),
                                                self._rebalance_bufs.data_ptr(), max(max_elems, 1),
                                                _lib.stream_ptr(self.device)), "mpsb_rebalance_sites")

    def check_status(self) -> int:
        """Largest SVD status word of every run() since the last call (0 = all converged); warns
        (``SVDNotConverged``) when a solve hit its sweep limit.  Called by the host-returning paths
        (norm, wavefunction, amplitudes), which synchronise anyway."""
        if not self._status_flags:
            return 0
        worst = int(_torch().stack(self._status_flags).max().item())
        self._status_flags = []
        if worst != 0:
            import warnings
            warnings.warn("a truncated SVD reached its sweep limit while still rotating (status "
                          f"{worst}); singular values of that application may be off by more than 1e-5",
                          SVDNotConverged, stacklevel=3)
        return worst

    # ------------------------------------------------------------------ whole-chain contractions
    @_on_device
    def inner_products(self, other: Optional["DeviceChain"] = None):
        """<self|other> per batch member (sum self * conj(other), ``mpsim/core.py:543-561``):
        device complex64 [B]."""
        torch = _torch()
        lib = _lib.load(require_device=True)
        other = self if other is None else other
        if min(self.bonds) == 0 or min(other.bonds) == 0:
            return torch.zeros(self.B, dtype=torch.complex64, device=self.device)
        a, b = self._site_refs(), other._site_refs()
        need = lib.mpsb_inner_workspace_bytes(self.B, self.d, max(self.bonds), max(other.bonds))
        ws = self.workspace(need)
        out = torch.empty(self.B, dtype=torch.complex64, device=self.device)
        _lib.check(lib.mpsb_inner_products(a.ctypes.data, b.ctypes.data, self.n, self.B, self.d, out.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mpsb_inner_products")
        return out

    def norms(self):
        """device float32 [B]: sqrt(Re <psi|psi>)  (``mpsim/core.py:563-565``)."""
        torch = _torch()
        return torch.sqrt(torch.clamp(self.inner_products().real, min=0.0))

    @_on_device
    def scale(self, factors) -> None:
        """site <- factors[b] * site for all sites (``mpsim/core.py:590-594``)."""
        torch = _torch()
        lib = _lib.load(require_device=True)
        f = torch.as_tensor(factors, dtype=torch.float32, device=self.device).contiguous()
        refs = _lib.to_device_bytes(self._site_refs(), self.device)
        max_elems = max(self.site_elems(i) for i in range(self.n))
        # the call is limited to 65535 (site, batch) jobs: chunk over sites
        per = max(1, MAX_JOBS_PER_CALL // self.B)
        for s0 in range(0, self.n, per):
            c = min(per, self.n - s0)
            _lib.check(lib.mpsb_scale_sites(refs.data_ptr() + s0 * _lib.SITE_REF.itemsize, c, self.B, self.d,
                                            f.data_ptr(), max(max_elems, 1), _lib.stream_ptr()), "mpsb_scale_sites")

    @_on_device
    def wavefunction(self, b: int = 0):
        """device complex64 [d**n], big-endian (``mpsim/core.py:483-500``)."""
        torch = _torch()
        lib = _lib.load(require_device=True)
        refs = self._site_refs()
        need = lib.mpsb_wavefunction_workspace_bytes(refs.ctypes.data, self.n, self.d)
        ws = self.workspace(need)
        out = torch.empty(self.d ** self.n, dtype=torch.complex64, device=self.device)
        _lib.check(lib.mpsb_wavefunction(refs.ctypes.data, self.n, self.d, int(b), out.data_ptr(), ws.data_ptr(),
                                         ws.numel(), _lib.stream_ptr()), "mpsb_wavefunction")
        return out

    @_on_device
    def amplitudes(self, bitstrings):
        """device complex64 [B][nbits]: <bits|psi_b> for each row of ``bitstrings`` (nbits x n)."""
        torch = _torch()
        lib = _lib.load(require_device=True)
        bits = np.ascontiguousarray(np.asarray(bitstrings, dtype=np.uint8).reshape(-1, self.n))
        if bits.size and bits.max() >= self.d:
            raise ValueError("basis state digit out of range for the qudit dimension")
        nbits = bits.shape[0]
        out = torch.zeros((self.B, nbits), dtype=torch.complex64, device=self.device)
        if nbits == 0 or min(self.bonds) == 0:
            return out
        bits_dev = torch.from_numpy(bits).to(self.device)
        refs = _lib.to_device_bytes(self._site_refs(), self.device)
        for b0 in range(0, self.B, MAX_JOBS_PER_CALL):
            nb = min(MAX_JOBS_PER_CALL, self.B - b0)
            # batch offset folded into the site pointers through a shifted copy of the refs
            r = self._site_refs()
            r["site"] += np.uint64(b0 * self.total * 8)
            rdev = refs if b0 == 0 else _lib.to_device_bytes(r, self.device)
            _lib.check(lib.mpsb_amplitudes(rdev.data_ptr(), self.n, nb, self.d, max(self.bonds), bits_dev.data_ptr(),
                                           nbits, out.data_ptr() + b0 * nbits * 8, _lib.stream_ptr()),
                       "mpsb_amplitudes")
        return out
