import numpy as np, sys, os
sys.path.insert(0, ".")
import torch
import mpsim_b200 as mp
from mpsim_b200 import planner, _lib
from tests.test_gpu_kernels import _svd
n = 9
mps = mp.MPS(n); mps.h(0)
ops = [(mp.cnot().tensor, (0, i), {}) for i in range(1, n)]
plan = planner.plan_operations(n, 2, mps._chain.bonds, ops)
d = 2
prev = mps.norm()
for kind, idx in plan.order:
    a = plan.apps2[idx]
    before = mps.copy()
    sub = planner.Plan(n, d, mps._chain.bonds)
    sub._add_adjacent(plan.gates[a.gate_index].reshape(2, 2, 2, 2), a.site, {"keep_left_canonical": a.left_canonical, "maxsvals": a.k}, 0, a.is_swap)
    cp = mps._chain.compile(sub, record_svals=True); mps._chain.run(cp)
    nrm = mps.norm()
    if abs(nrm - prev) > 1e-5:
        print("norm drop at app", idx, a, prev, "->", nrm)
        A = before._chain.site_view(a.site).cpu().numpy().astype(np.complex128)
        B = before._chain.site_view(a.site + 1).cpu().numpy().astype(np.complex128)
        G = plan.gates[a.gate_index].reshape(2, 2, 2, 2).astype(np.complex128)
        th = np.einsum("xypq,lpm,mqr->lxyr", G, A, B).reshape(A.shape[0] * 2, B.shape[2] * 2)
        np.save("gpurun_out/bad_theta.npy", th.astype(np.complex64))
        print("theta shape", th.shape, "svals ref", np.linalg.svd(th, compute_uv=False))
        left, right, sv, info = _svd(th.astype(np.complex64)[None], a.k, 1)
        print("gpu svals", sv[0], "info", info)
        print("recon err", np.abs(left[0] @ right[0] - th).max(), "iso err", np.abs(left[0].conj().T @ left[0] - np.eye(a.k)).max())
        newA = mps._chain.site_view(a.site).cpu().numpy(); newB = mps._chain.site_view(a.site + 1).cpu().numpy()
        rec = np.einsum("lxk,kyr->lxyr", newA, newB).reshape(th.shape)
        print("in-situ recon err", np.abs(rec - th).max())
        break
    prev = nrm
