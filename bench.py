#!/usr/bin/env python
"""Benchmark of the two-qudit gate application path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle)

Workload (config.workload): BASELINE.json configs[3] -- a batch of independent 40-qubit random
brickwork circuits (depth 20, Haar gates, maxsvals = chi = 64), sharded one contiguous slice per
GPU with no cross-GPU traffic; weak scaling with --batch-per-gpu circuits per GPU (512 -> the
4096-circuit config at 8 GPUs).  A step = reset |0..0>, apply all 390 gates of every circuit of
the slice, compute every circuit's norm.  Unit of work = one adjacent two-qudit application
(theta + truncated SVD + absorb), as SURVEY.md 8(d).

The one JSON line printed by rank 0 follows the driver's contract (metric/value/unit/n_gpus/
steps/warmup/ms_per_step/higher_is_better/scaling/vs_baseline/dtype/data/config/clocks/e2e/
gpu_launches) plus "roofline" for the dominant kernel and "cpu_baseline".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram bytes per 128 x 128 job of svd_small_kernel and where that number comes from (ncu --set full capture)
SVD_SMALL_DRAM_BYTES_PER_JOB = (20.165e6 + 0.133e6) / 148.0
SVD_SMALL_TRAFFIC_SOURCE = ("profiles/r2_svd_small_ncu_full.txt (ncu --set full of the kernel as of commit 71017ec; 148 jobs: "
                            "20.16 MB read + 0.13 MB written; algorithmic 128 KB in + 128 KB out per job)")

METRIC = "two_qudit_gate_applications_per_sec"
UNIT = "applications/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch-per-gpu", type=int, default=512)
    p.add_argument("--nqubits", type=int, default=40)
    p.add_argument("--depth", type=int, default=20)
    p.add_argument("--chi", type=int, default=64)
    p.add_argument("--cpu-baseline-circuits", type=int, default=1)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the chi=256 single-chain timing")
    return p.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": f"batch of independent {args.nqubits}-qubit random brickwork circuits, depth {args.depth}, "
                    f"chi={args.chi} (BASELINE.json configs[3]), {args.batch_per_gpu} circuits per GPU",
        "nqubits": args.nqubits, "depth": args.depth, "chi": args.chi,
        "circuits_per_gpu": args.batch_per_gpu, "circuits_total": args.batch_per_gpu * n_gpus,
        "parallelism": f"batch-sharded x{n_gpus}, no data-path collective",
        "l2_policy": "inputs larger than L2 (site slab >= 1.3 GB per GPU at 512 circuits)",
    }


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def haar_gates(count, rng):
    """count Haar-random 4x4 unitaries (Mezzadri; mpsim/gates.py:269-286), vectorised: the workload's own
    generator (mpsim_b200.circuits.haar_gate_stack), shared by the GPU arm, the CPU arms and the fixtures."""
    from mpsim_b200 import circuits
    return circuits.haar_gate_stack(count, rng)


def flops_svd_lapack(m, n):
    """Golub-Reinsch thin count x4 for complex."""
    mx, mn = max(m, n), min(m, n)
    return 4.0 * (14.0 * mx * mn * mn + 8.0 * mn ** 3)


def flops_theta(d, cl, cm, cr):
    return 8.0 * d * d * cl * cm * cr + 8.0 * d ** 4 * cl * cr


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline (the oracle = restatement of the reference's path)
# ------------------------------------------------------------------------------------------------
def _oracle_one_circuit(payload):
    """Runs in a worker process with 1 BLAS thread: one circuit through the complex128 oracle."""
    nq, depth, chi, seed, track = payload
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1)
    except Exception:  # noqa: BLE001
        ctx = None
    from oracle.mps_oracle import OracleMPS
    from mpsim_b200 import circuits
    # circuit `seed - 1000` of the batched workload: the very gate arrays the GPU arm stages for that member
    ops = circuits.brickwork_member(nq, depth, seed - 1000)
    t0 = time.perf_counter()
    mps = OracleMPS(nq, dtype=np.complex128, track_norms=track)
    for op in ops:
        mps.apply_two_qudit_gate(op.tensor, *op.indices, maxsvals=chi, keep_left_canonical=op.keep_left_canonical)
    nrm = mps.norm()
    dt = time.perf_counter() - t0
    del ctx
    smax = np.array([max(float(t["s_kept"].max()) if t["s_kept"].size else 0.0,
                         float(t["s_trunc"].max()) if t["s_trunc"].size else 0.0) for t in mps.trace])
    skept_last = np.array([float(t["s_kept"][-1]) if t["s_kept"].size else 0.0 for t in mps.trace])
    return len(ops), dt, nrm, smax, skept_last


def cpu_reference_step(args, ncircuits, nprocs, track_norms, seed0=1000, pool=None):
    """ncircuits circuits of the workload through the oracle on nprocs host processes (``pool``: an
    already running multiprocessing pool -- process start-up stays outside the timed wall).
    Returns (applications, wall seconds)."""
    payloads = [(args.nqubits, args.depth, args.chi, seed0 + i, track_norms) for i in range(ncircuits)]
    t0 = time.perf_counter()
    if pool is None:
        res = [_oracle_one_circuit(p) for p in payloads]
    else:
        res = pool.map(_oracle_one_circuit, payloads, chunksize=1)
    wall = time.perf_counter() - t0
    cpu_reference_step.last = res
    return sum(r[0] for r in res), wall


def _pool_ready(_):
    import numpy  # noqa: F401  (import cost paid before the timed region)
    from oracle.mps_oracle import OracleMPS  # noqa: F401
    return os.getpid()


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the
    reference itself cannot be imported here -- tensornetwork==0.2.1 / cirq absent)
    on all host cores, one circuit per core per step (1 BLAS thread each).
    The worker pool is created and warmed before anything is timed; ``--warmup`` untimed steps run
    first (at least one)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    nprocs = max(1, cores)
    ncirc = nprocs
    warm = max(1, args.warmup)
    with mp.get_context("fork").Pool(nprocs) as pool:
        pool.map(_pool_ready, range(4 * nprocs))
        for _ in range(warm):
            cpu_reference_step(args, ncirc, nprocs, False, pool=pool)
        apps, wall = 0, 0.0
        for _ in range(args.steps):
            a, w = cpu_reference_step(args, ncirc, nprocs, False, pool=pool)
            apps += a; wall += w
        value = apps / wall
        fa, fw = cpu_reference_step(args, ncirc, nprocs, True, pool=pool)
    sample = (f"{ncirc} circuits per step (members 0..{ncirc - 1} of the batched workload: {args.nqubits} qubits, depth "
              f"{args.depth}, chi {args.chi}; {apps // max(args.steps, 1)} applications), complex128 numpy/LAPACK "
              f"restatement, one process per core with 1 BLAS thread, pool started before the timed region; no "
              f"per-application norm bookkeeping (conservative: the reference also calls norm() after every "
              f"application, core.py:1160-1161)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128",
        "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nprocs, "kind": "port", "sample": sample,
                         "faithful_value_with_norm_bookkeeping": fa / fw},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def time_dominant_kernel(torch, args, batch, reps=3):
    """Average launch duration of the dominant kernel (single-CTA Jacobi SVD of the 2chi x 2chi
    theta) measured with CUDA events on the launching stream, on theta matrices taken from the
    middle of the actual circuits (one per circuit of the slice)."""
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    chain = batch._chain
    d, chi, B = 2, args.chi, chain.B
    # disjoint bonds of the full shape (chi, chi, chi) in the middle of the chain
    sites = [i for i in range(0, chain.n - 1, 2)
             if chain.bonds[i] == chi and chain.bonds[i + 1] == chi and chain.bonds[i + 2] == chi]
    if not sites:
        return None
    dev = chain.device
    m = d * chi
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    # one CTA per matrix and SM: time a whole number of waves, as the launches of the workload are
    # (a layer of the circuit is ~19 bonds x 512 members = 66 waves; 512 jobs alone would be 3.46)
    want = 6 * nsm
    nd = min(len(sites), -(-want // B))
    gates = torch.from_numpy(haar_gates(B, np.random.default_rng(7)).reshape(B, 16)).to(dev)
    desc = np.zeros(nd, dtype=_lib.GATE2_DESC)
    for t in range(nd):
        desc[t] = (chain.site_ptr(sites[t]), chain.site_ptr(sites[t] + 1), 0, 0, gates.data_ptr(), 0,
                   chain.total, chain.total, 0, 0, 16, 0)
    ddesc = _lib.to_device_bytes(desc, dev)
    theta = torch.empty((nd * B, m, m), dtype=torch.complex64, device=dev)
    _lib.check(lib.mpsb_theta(ddesc.data_ptr(), nd, B, d, chi, chi, chi, theta.data_ptr(), None, 0, _lib.stream_ptr()))
    nj = min(nd * B, 65535)
    nj = max(nj // nsm * nsm, min(nj, nsm))
    left = torch.empty((nj, m, chi), dtype=torch.complex64, device=dev)
    right = torch.empty((nj, chi, m), dtype=torch.complex64, device=dev)
    info = torch.zeros((nj, 2), dtype=torch.int32, device=dev)
    ws = torch.empty(max(lib.mpsb_svd_workspace_bytes(nj, m, m), 256), dtype=torch.uint8, device=dev)

    def launch():
        _lib.check(lib.mpsb_svd(theta.data_ptr(), nj, m, m, chi, 1, left.data_ptr(), right.data_ptr(), None,
                                info.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    # mpsb_svd = one strided D2D copy + the SVD kernel; the copy is ~0.1% of the time
    ms = e0.elapsed_time(e1) / reps
    sweeps = float(info[:, 1].float().mean().item())
    out = {"ms_per_launch": ms, "jobs_per_launch": nj, "m": m, "n": m, "mean_sweeps": sweeps}
    # profiling build of the library only (MPSB_NVCC_EXTRA=-DMPSB_PROFILE): phase clocks of CTA 0
    import ctypes
    if hasattr(lib, "mpsb_debug_phase_clocks"):
        clk = (ctypes.c_longlong * 32)()
        if lib.mpsb_debug_phase_clocks(clk) == 0:
            names = ["load", "qr_R", "jacobi", "sort", "W=XV", "form_Q", "P=QhX", "write"]
            c = list(clk)
            out["phase_cycles_cta0"] = {n: c[i + 1] - c[i] for i, n in enumerate(names)}
            out["sweeps_cta0"] = int(info[0, 1].item())
    return out


def time_secondary(torch, args, batch):
    """Device-timed numbers for the other kernels of the path (CUDA events, same run)."""
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    chain = batch._chain
    d, chi, B = 2, args.chi, chain.B
    out = {}

    def ev(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    site = next((i for i in range(chain.n - 1)
                 if chain.bonds[i] == chi and chain.bonds[i + 1] == chi and chain.bonds[i + 2] == chi), None)
    if site is not None:
        gates = torch.from_numpy(haar_gates(B, np.random.default_rng(7)).reshape(B, 16)).to(chain.device)
        desc = np.zeros(1, dtype=_lib.GATE2_DESC)
        desc[0] = (chain.site_ptr(site), chain.site_ptr(site + 1), 0, 0, gates.data_ptr(), 0,
                   chain.total, chain.total, 0, 0, 16, 0)
        ddesc = _lib.to_device_bytes(desc, chain.device)
        m = d * chi
        theta = torch.empty((B, m, m), dtype=torch.complex64, device=chain.device)
        ms = ev(lambda: _lib.check(lib.mpsb_theta(ddesc.data_ptr(), 1, B, d, chi, chi, chi, theta.data_ptr(), None, 0,
                                                  _lib.stream_ptr())))
        fl = flops_theta(d, chi, chi, chi) * B
        out["theta_kernel_ffma"] = {"ms_per_launch": ms, "jobs": B, "tflops": fl / (ms * 1e-3) / 1e12,
                                    "note": "split-real FFMA tiles (used for d != 2 or chi < 32); flops = 8 d^2 chi^3 + 8 d^4 chi^2 per application"}
        wsb = lib.mpsb_theta_workspace_bytes(1, B, d, chi, chi, chi)
        if wsb:
            tws = torch.empty(wsb + 256, dtype=torch.uint8, device=chain.device)
            ms = ev(lambda: _lib.check(lib.mpsb_theta(ddesc.data_ptr(), 1, B, d, chi, chi, chi, theta.data_ptr(),
                                                      tws.data_ptr(), tws.numel(), _lib.stream_ptr())))
            out["theta_kernel"] = {"ms_per_launch": ms, "jobs": B, "tflops": fl / (ms * 1e-3) / 1e12,
                                   "note": "tcgen05 3xTF32 (operand split + TMA/TMEM GEMM + gate epilogue), the path mpsb_apply_gate2 uses"}
            del tws
        # one-qudit gate on that site of every member: 16*d*chiL*chiR bytes per site
        g1 = torch.from_numpy(np.tile(np.array([1, 1, 1, -1], dtype=np.complex64) / np.sqrt(2), (B, 1))).to(chain.device)
        d1 = np.zeros(1, dtype=_lib.GATE1_DESC)
        d1[0] = (chain.site_ptr(site), chain.site_ptr(site), g1.data_ptr(), chain.total, chain.total, 4, chi, chi)
        dd1 = _lib.to_device_bytes(d1, chain.device)
        ms = ev(lambda: _lib.check(lib.mpsb_apply_gate1(dd1.data_ptr(), 1, B, d, chi * d * chi, _lib.stream_ptr())))
        out["gate1_kernel"] = {"ms_per_launch": ms, "gbs": 16.0 * d * chi * chi * B / (ms * 1e-3) / 1e9,
                               "note": "one site (chi x 2 x chi) of every batch member: small launch, latency bound"}
    ms = ev(lambda: chain.norms(), reps=3)
    site_bytes = sum(8.0 * chain.site_elems(i) for i in range(chain.n)) * B
    # transfer chain: env <- A^H (env A) per site = 2 x 8 d chiL chiR max(chiL, chiR)-ish flops; exactly
    # 8 chiL chiL d chiR (env . A) + 8 chiR d chiL chiR (A^H . tmp) real flops per site and member
    norm_flops = sum(8.0 * chain.bonds[i] * chain.bonds[i] * d * chain.bonds[i + 1]
                     + 8.0 * chain.bonds[i + 1] * d * chain.bonds[i] * chain.bonds[i + 1] for i in range(chain.n)) * B
    out["norm_chain"] = {"ms": ms, "launches": 2 * chain.n + 2, "gbs_sites_read_once": site_bytes / (ms * 1e-3) / 1e9,
                         "tflops": norm_flops / (ms * 1e-3) / 1e12,
                         "note": "2 x 8 d chi^3 flops against 8 d chi^2 bytes per site: intensity 2 chi = 128 flop/B at "
                                 "chi = 64, i.e. FFMA bound (cgemm_kernel), not HBM bound; the GB/s figure is what the "
                                 "north star asks to be reported"}
    # renormalize scale pass (scale_kernel): every site read and written once, 16 d chiL chiR bytes per site
    ones = torch.ones(B, dtype=torch.float32, device=chain.device)
    ms = ev(lambda: chain.scale(ones), reps=3)
    out["scale_kernel"] = {"ms": ms, "gbs": 2.0 * site_bytes / (ms * 1e-3) / 1e9,
                           "note": "site <- f[b] * site over the whole slab (read + write), HBM bound"}
    # amplitude_kernel: <bits|psi_b> for 256 bitstrings of every member: one d-slice of every site per amplitude
    bits = np.random.default_rng(5).integers(0, 2, size=(256, chain.n)).astype(np.uint8)
    ms = ev(lambda: chain.amplitudes(bits), reps=3)
    slice_bytes = sum(8.0 * chain.bonds[i] * chain.bonds[i + 1] for i in range(chain.n)) * B * bits.shape[0]
    out["amplitude_kernel"] = {"ms": ms, "amplitudes": int(B * bits.shape[0]),
                               "gbs_slices_read": slice_bytes / (ms * 1e-3) / 1e9,
                               "note": "vector chain per (member, bitstring); the 256 bitstrings of a member re-read its "
                                       "sites through L2, so this is an L2/latency figure, not HBM"}
    return out


def time_wavefunction(torch):
    """mpsb_wavefunction (chain of cgemm_kernel launches): 8 d^n bytes written, on a
    24-qubit chi <= 64 state prepared by the library itself."""
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    n, depth, chi = 24, 12, 64
    mps = mp.MPS(n)
    mps._execute([(o.tensor, o.indices, {"maxsvals": chi, "keep_left_canonical": o.keep_left_canonical})
                  for o in circuits.brickwork(n, depth, seed=9)])
    for _ in range(3):
        wf = mps.wavefunction_device()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):                                       # every repetition timed on its own
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wf = mps.wavefunction_device()
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    out_bytes = 8.0 * 2 ** n
    # partial products of the two-sided contraction (api.cu: wf_plan picks the split that minimises them):
    # each is written once and read once
    bonds = mps._chain.bonds
    Ls = [2.0 ** (i + 1) * bonds[i + 1] for i in range(n)]
    Rs = [2.0 ** (n - i) * bonds[i] for i in range(n)]
    inter = min(sum(Ls[1:sp]) + sum(Rs[sp:n - 1]) for sp in range(1, n))
    del wf
    return {"nqubits": n, "chi": chi, "ms": ms, "ms_all_repetitions": times, "gbs_output_written": out_bytes / (ms * 1e-3) / 1e9,
            "gbs_with_partial_products": (out_bytes + 16.0 * inter) / (ms * 1e-3) / 1e9,
            "note": "left and right chains of strided complex GEMMs joined by one product (tcgen05 kernel); bytes = "
                    "8 d^n written (+ the partial products written and re-read)"}


def measure_tf32_peak(torch):
    """Dense TF32 tensor-core throughput measured live (cuBLAS fp32 GEMM with TF32 allowed): the
    'complex tensor-core peak' of SURVEY.md 8(d) is this / 3 (three TF32 products per fp32 product)."""
    n = 8192
    a = torch.randn((n, n), device="cuda", dtype=torch.float32)
    b = torch.randn((n, n), device="cuda", dtype=torch.float32)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def time_theta_tc(torch, tf32_peak):
    """theta contraction alone at the large shapes of BASELINE.json (chi = 256: configs[2], 50
    disjoint bonds of one layer; chi = 1024: configs[4], 8 bonds), tensor-core kernel, CUDA events.
    frac = achieved / (measured dense TF32 / 3)."""
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    out = {}
    for chi, jobs in ((256, 50), (1024, 8)):
        d = 2
        A = torch.randn((jobs, chi, d, chi), dtype=torch.complex64, device="cuda")
        Bm = torch.randn((jobs, chi, d, chi), dtype=torch.complex64, device="cuda")
        G = torch.from_numpy(haar_gates(jobs, np.random.default_rng(11)).reshape(jobs, 16)).cuda()
        desc = np.zeros(1, dtype=_lib.GATE2_DESC)
        desc[0] = (A.data_ptr(), Bm.data_ptr(), 0, 0, G.data_ptr(), 0, chi * d * chi, chi * d * chi, 0, 0, 16, 0)
        ddesc = _lib.to_device_bytes(desc, "cuda")
        theta = torch.empty((jobs, d * chi, d * chi), dtype=torch.complex64, device="cuda")
        ws = torch.empty(lib.mpsb_theta_workspace_bytes(1, jobs, d, chi, chi, chi) + 256, dtype=torch.uint8, device="cuda")

        def run():
            _lib.check(lib.mpsb_theta(ddesc.data_ptr(), 1, jobs, d, chi, chi, chi, theta.data_ptr(), ws.data_ptr(),
                                      ws.numel(), _lib.stream_ptr()))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = flops_theta(d, chi, chi, chi) * jobs
        tf = fl / (ms * 1e-3) / 1e12
        out[f"chi{chi}"] = {"jobs": jobs, "ms_per_launch": ms, "tflops_complex_equivalent": tf,
                            "complex_tensor_core_peak_tflops": tf32_peak / 3.0, "frac_of_complex_peak": tf / (tf32_peak / 3.0),
                            "note": "includes the operand split / transpose kernels; operands %.0f MB > L2 at chi=1024"
                                    % (2 * jobs * chi * d * chi * 8 / 1e6)}
        del A, Bm, theta, ws
    return out


def cpu_dominant_shape(chi, reps):
    """The reference's CPU arithmetic for ONE adjacent application of shape (chi, chi, chi, k = chi):
    the complex128 numpy/LAPACK oracle (mpsim/core.py:1060-1152 restated) on random sites, all host
    cores available to BLAS.  Reported next to the GPU number of the same shape; bounded sample."""
    from oracle.mps_oracle import OracleMPS
    rng = np.random.default_rng(chi)

    def rnd(*shape):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2 * chi)
    ora = OracleMPS(4, dtype=np.complex128)
    ora.sites = [rnd(1, 2, chi), rnd(chi, 2, chi), rnd(chi, 2, chi), rnd(chi, 2, 1)]
    gates = haar_gates(reps, rng).astype(np.complex128)
    t0 = time.perf_counter()
    for r in range(reps):
        ora.apply_two_qudit_gate(gates[r].reshape(2, 2, 2, 2), 1, 2, maxsvals=chi)
    dt = (time.perf_counter() - t0) / reps
    return {"ms_per_application": dt * 1e3, "applications_per_sec": 1.0 / dt, "kind": "port",
            "cores": len(os.sched_getaffinity(0)),
            "sample": f"{reps} application(s) of shape ({chi},{chi},{chi}), k={chi}, random sites, complex128 numpy/LAPACK "
                      "oracle with the host's BLAS threads"}


def time_chi256(torch):
    """BASELINE.json configs[2]: 100-qubit brickwork, depth 20, chi = 256, one chain (replicas only)."""
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    from mpsim_b200.planner import plan_operations
    n, depth, chi = 100, 20, 256
    ops = circuits.brickwork(n, depth, seed=3)
    mps = mp.MPS(n)
    plan = plan_operations(n, 2, mps._chain.bonds,
                           [(o.tensor, o.indices, {"maxsvals": chi, "keep_left_canonical": o.keep_left_canonical}) for o in ops])
    cp = mps._chain.compile(plan)
    mps._chain.run(cp)
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        mps._chain.reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mps._chain.run(cp, upload=False)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    info = cp.info.cpu().numpy()[: len(plan.apps2)]
    fl = sum(flops_svd_lapack(2 * a.chiL, 2 * a.chiR) + flops_theta(2, a.chiL, a.chiM, a.chiR) for a in plan.apps2)
    return {"workload": "100-qubit brickwork depth 20 chi=256 (BASELINE.json configs[2]), one chain",
            "applications": len(plan.apps2), "ms": ms, "ms_all_repetitions": times,
            "applications_per_sec": len(plan.apps2) / (ms * 1e-3),
            "lapack_equivalent_tflops": fl / (ms * 1e-3) / 1e12,
            "svd_not_converged": int((info[:, 0] != 0).sum()), "svd_mean_sweeps": float(info[:, 1].mean())}


def time_chi1024(torch, jobs=4):
    """chi = 1024 (the shape of BASELINE.json configs[4]): `jobs` adjacent applications on disjoint
    bonds with random Gaussian sites (flat spectrum), through mpsb_apply_gate2 (tensor-core theta
    + block-Jacobi SVD of the 2048 x 2048 theta + split), CUDA events, one warm-up call."""
    from mpsim_b200 import _lib
    lib = _lib.load(require_device=True)
    d, chi = 2, 1024
    A = torch.randn((jobs, chi, d, chi), dtype=torch.complex64, device="cuda") / np.sqrt(2 * chi)
    Bm = torch.randn((jobs, chi, d, chi), dtype=torch.complex64, device="cuda") / np.sqrt(2 * chi)
    A0, B0 = A.clone(), Bm.clone()
    G = torch.from_numpy(haar_gates(jobs, np.random.default_rng(13)).reshape(jobs, 16)).cuda()
    sv = torch.zeros((jobs, d * chi), dtype=torch.float32, device="cuda")
    desc = np.zeros(1, dtype=_lib.GATE2_DESC)
    se = chi * d * chi
    desc[0] = (A.data_ptr(), Bm.data_ptr(), A.data_ptr(), Bm.data_ptr(), G.data_ptr(), sv.data_ptr(), se, se, se, se, 16, d * chi)
    ddesc = _lib.to_device_bytes(desc, "cuda")
    info = torch.zeros((jobs, 2), dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.mpsb_gate2_workspace_bytes(1, jobs, d, chi, chi, chi, chi) + 256, dtype=torch.uint8, device="cuda")
    times = []
    for rep in range(4):                                     # first = warm-up
        A.copy_(A0); Bm.copy_(B0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.mpsb_apply_gate2(ddesc.data_ptr(), 1, jobs, d, chi, chi, chi, chi, 1, ws.data_ptr(), ws.numel(),
                                        info.data_ptr(), _lib.stream_ptr()), "mpsb_apply_gate2")
        e1.record(); torch.cuda.synchronize()
        if rep > 0:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    inf = info.cpu().numpy()
    fl = (flops_svd_lapack(d * chi, d * chi) + flops_theta(d, chi, chi, chi)) * jobs
    return {"workload": f"{jobs} disjoint bonds at chi=1024 (2048 x 2048 theta, k=1024), random sites",
            "applications": jobs, "ms": ms, "ms_all_repetitions": times, "applications_per_sec": jobs / (ms * 1e-3),
            "lapack_equivalent_tflops": fl / (ms * 1e-3) / 1e12,
            "svd_not_converged": int((inf[:, 0] != 0).sum()), "svd_mean_sweeps": float(inf[:, 1].mean())}


def measure_ffma_peak(torch):
    """FP32 FMA throughput from a register-only microkernel (mpsim_b200/csrc/bench/ffma_peak.cu, built by
    __graft_entry__.build() into libmpsb_bench.so): (FFMA TFLOP/s, FFMA2 TFLOP/s) or None."""
    import ctypes
    path = os.path.join(ROOT, "mpsim_b200", "libmpsb_bench.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    lib.mpsb_bench_ffma_peak.restype = ctypes.c_int
    lib.mpsb_bench_ffma_peak.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    nsm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    scratch = torch.empty(2 * nsm * 512, dtype=torch.float32, device="cuda")
    out = (ctypes.c_double * 2)()
    rc = lib.mpsb_bench_ffma_peak(out, scratch.data_ptr(), nsm, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return (float(out[0]), float(out[1])) if rc == 0 else None


def measure_fp32_peak(torch):
    """FP32 FFMA peak of this GPU measured live with torch (dependent-free FMA chains are not
    expressible in torch; use a large fp32 GEMM through cuBLAS as the FFMA proxy)."""
    n = 8192
    a = torch.randn((n, n), device="cuda", dtype=torch.float32)
    b = torch.randn((n, n), device="cuda", dtype=torch.float32)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        return 3 * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def run_our_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: mpsim_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import mpsim_b200 as mp
    from mpsim_b200 import circuits
    from mpsim_b200.distributed import shard_range, gather_slices

    n, depth, chi = args.nqubits, args.depth, args.chi
    B = args.batch_per_gpu
    total = B * world
    lo, hi = shard_range(total, rank, world)
    structure = circuits.brickwork(n, depth, seed=0)
    nops = len(structure)
    batch = mp.MPSBatch(B, n)
    cp = batch.compile(structure, maxsvals=chi)
    napps = len(cp.plan.apps2)
    # per-circuit Haar gates: circuit c of the global batch uses the stream seeded 1000 + c
    gates = np.empty((nops, B, 16), dtype=np.complex64)
    for b in range(B):
        gates[:, b, :] = circuits.batch_member_gates(nops, lo + b)        # stream seeded 1000 + member
    gates_pinned = torch.from_numpy(gates).pin_memory()
    # results of a step: every circuit's norm and NAMP amplitudes at seeded bitstrings; with N > 1 they are
    # all-gathered over NCCL INSIDE the timed region (the path's only collective)
    NAMP = 16
    amp_bits = np.random.default_rng(77).integers(0, 2, size=(NAMP, n)).astype(np.uint8)
    norms_host = torch.empty(total, dtype=torch.float32).pin_memory()
    amps_host = torch.empty((total, NAMP), dtype=torch.complex64).pin_memory()
    h2d = gates_pinned.numel() * 8
    d2h = total * 4 + total * NAMP * 8

    def results():
        norms = gather_slices(batch.norms_device(), total)
        amps = gather_slices(batch.amplitudes_device(amp_bits), total)
        return norms, amps

    def step_resident():
        batch.reset()
        batch.run(cp, upload=False)
        return results()

    def step_e2e():
        # public API with HOST buffers: gates in, norms and amplitudes out
        batch.stage_gates(cp, gates_pinned.numpy())
        batch.reset()
        batch.run(cp, upload=True)
        norms, amps = results()
        norms_host.copy_(norms, non_blocking=True)
        amps_host.copy_(amps, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return norms_host, amps_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item()), out

    batch.stage_gates(cp, gates)
    batch._chain.upload_gates(cp)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, (norms_all, amps_all) = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    status = batch.status(cp)
    not_converged = int((status[..., 0] != 0).sum())
    mean_sweeps = float(status[..., 1].mean())
    for _ in range(1):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    apps_per_step_total = napps * total
    value = apps_per_step_total * args.steps / (ms_res * 1e-3)
    e2e_value = apps_per_step_total * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        launches_per_step = 0
        for L in cp.launches:
            if L[0] == "g1":
                launches_per_step += 1
            elif L[0] == "g2layer":                   # the shape groups of one layer, concurrently on library streams
                for grp in L[1]:
                    tc = grp["chiL"] >= 32 and grp["chiR"] >= 32 and grp["chiM"] >= 16
                    launches_per_step += (3 if tc else 1) + 1
            else:
                _, _, _, chiL, chiM, chiR, _, _ = L
                tc = chiL >= 32 and chiR >= 32 and chiM >= 16        # api.cu: theta_uses_tc (d = 2)
                launches_per_step += (3 if tc else 1) + 1            # theta (split A, split/transpose B, GEMM | FFMA) + SVD
        launches_per_step += 2 * n + 2 + 1            # norm chain: 2 GEMMs per site + init + gather; amplitude kernel
        # dominant kernel roofline (measured live, CUDA events on the launching stream)
        dom = time_dominant_kernel(torch, args, batch)
        sgemm_peak = measure_fp32_peak(torch)
        ffma = measure_ffma_peak(torch)
        fp32_peak = ffma[0] if ffma else sgemm_peak
        roof = None
        if dom is not None:
            fl = flops_svd_lapack(dom["m"], dom["n"]) * dom["jobs_per_launch"]
            achieved = fl / (dom["ms_per_launch"] * 1e-3) / 1e12
            roof = {"kernel": "svd_small_kernel (single-CTA QR + one-sided Jacobi, 128x128 complex64)",
                    "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak,
                    "peak_source": ("FFMA microkernel (mpsim_b200/csrc/bench/ffma_peak.cu) measured live in this run"
                                    if ffma else "cuBLAS SGEMM measured live in this run (libmpsb_bench.so missing)"),
                    "sgemm_tflops_measured_here": sgemm_peak,
                    "ffma2_tflops_measured_here": ffma[1] if ffma else None,
                    "frac_of_sgemm": achieved / sgemm_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum per job of the 148-job ncu capture named in
                    # traffic_source, scaled to this launch's job count (not measured by this run)
                    "traffic": SVD_SMALL_DRAM_BYTES_PER_JOB * dom["jobs_per_launch"],
                    "traffic_source": SVD_SMALL_TRAFFIC_SOURCE,
                    "note": "achieved = LAPACK-equivalent flops 4(14 mx mn^2 + 8 mn^3) x jobs / CUDA-event launch "
                            "time over whole waves of thetas taken from the circuits; peak = fp32 FMA throughput "
                            "of a register-only microkernel measured live (MEASURED_PEAKS.json records only HBM "
                            "and bf16); the kernel is FMA/latency bound in shared memory, not HBM bound",
                    "ms_per_launch": dom["ms_per_launch"], "jobs_per_launch": dom["jobs_per_launch"],
                    "mean_sweeps": dom["mean_sweeps"]}
            if "phase_cycles_cta0" in dom:
                roof["phase_cycles_cta0"] = dom["phase_cycles_cta0"]
                roof["sweeps_cta0"] = dom["sweeps_cta0"]
        secondary = time_secondary(torch, args, batch)
        roof_theta = None
        # GPU side of the parity check done in the cpu_baseline leg below: member 0 run alone with its
        # singular values recorded
        chk = None
        if not args.no_cpu_baseline and lo == 0:
            one = mp.MPSBatch(1, n)
            cp1 = one.compile(structure, record_svals=True, maxsvals=chi)
            one.stage_gates(cp1, gates[:, :1])
            one.run(cp1)
            sv1 = one.singular_values(cp1)[:, 0]
            chk = {"norm": float(one.norms()[0]), "smax": sv1.max(axis=1),
                   "skept_last": np.array([sv1[t, a.k - 1] if a.k > 0 else 0.0 for t, a in enumerate(cp1.plan.apps2)]),
                   "k": [a.k for a in cp1.plan.apps2], "batch_norm": float(norms_all[0].item())}
            del one, cp1
        if not args.no_extra:
            del batch
            torch.cuda.empty_cache()
            secondary["wavefunction"] = time_wavefunction(torch)
            secondary["chi256"] = time_chi256(torch)
            if not args.no_cpu_baseline:
                secondary["chi256"]["cpu_baseline"] = cpu_dominant_shape(256, 6)
            tf32_peak = measure_tf32_peak(torch)
            secondary["theta_tensor_core"] = time_theta_tc(torch, tf32_peak)
            secondary["tf32_tflops_measured_here"] = tf32_peak
            secondary["chi1024"] = time_chi1024(torch)
            if not args.no_cpu_baseline:
                secondary["chi1024"]["cpu_baseline"] = cpu_dominant_shape(1024, 1)
            t = secondary["theta_tensor_core"]["chi1024"]
            roof_theta = {"kernel": "tc_cgemm_kernel + operand split (theta, chi=1024, 8 bonds)", "bound": "tensor",
                          "achieved": t["tflops_complex_equivalent"], "peak": t["complex_tensor_core_peak_tflops"],
                          "unit": "TFLOP/s", "frac": t["frac_of_complex_peak"], "traffic": 1.682e9 + 0.449e9,
                          "note": "achieved = (8 d^2 chi^3 + 8 d^4 chi^2) x bonds / CUDA-event time; peak = dense TF32 "
                                  "measured live in this run / 3 (3xTF32); traffic = dram read + write of the GEMM kernel "
                                  "in profiles/r1_tc_theta_chi1024_ncu_full.txt (tensor pipe 70 % active there)"}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        cpu = None
        if not args.no_cpu_baseline:
            a1, w1 = cpu_reference_step(args, args.cpu_baseline_circuits, 1, False)
            a2, w2 = cpu_reference_step(args, args.cpu_baseline_circuits, 1, True)
            # the same oracle run is the checker of member 0 (identical gate arrays): free-running bounds of
            # tests/test_gpu_baseline.py (a truncated circuit amplifies rounding; per-application parity at
            # 1e-5 is the teacher-forced test there)
            parity = None
            if chk is not None:
                _, _, o_norm, o_smax, o_last = cpu_reference_step.last[0]
                e_max = float(np.max(np.abs(chk["smax"] - o_smax) / np.maximum(o_smax, 1e-300)))
                e_last = float(np.max(np.abs(chk["skept_last"] - o_last) / np.maximum(o_smax, 1e-300)))
                parity = {"member": 0, "norm_gpu": chk["norm"], "norm_gpu_in_batch": chk["batch_norm"], "norm_oracle": o_norm,
                          "norm_rel_err": abs(chk["norm"] - o_norm) / o_norm,
                          "sigma_max_trace_err": e_max, "last_kept_sigma_trace_err": e_last,
                          "applications": len(chk["k"]), "mode": "free-running, complex64 vs complex128 oracle"}
                assert abs(chk["norm"] - o_norm) <= 2e-3 * o_norm, parity
                assert abs(chk["batch_norm"] - chk["norm"]) <= 1e-5 * chk["norm"] + 1e-7, parity
                assert e_max <= 5e-3 and e_last <= 5e-3, parity
            cpu = {"value": a1 / w1, "unit": UNIT, "cores": 1, "kind": "port", "parity_check": parity,
                   "sample": f"{args.cpu_baseline_circuits} circuit(s) of the same workload ({a1} applications) "
                             "through the complex128 numpy/LAPACK restatement of mpsim/core.py:950-1161, 1 process, "
                             "1 BLAS thread, without the per-application norm bookkeeping",
                   "faithful_value_with_norm_bookkeeping": a2 / w2,
                   "host_cores_available": len(os.sched_getaffinity(0))}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
            "cpu_baseline": cpu,
            "secondary": secondary,
            "roofline_theta": roof_theta,
            "applications_per_step": apps_per_step_total,
            "svd_not_converged": not_converged, "svd_mean_sweeps": mean_sweeps,
            "norm_mean": float(norms_all.float().mean().item()),
            "amplitude_abs_mean": float(amps_all.abs().mean().item()),
            "peaks": {"hbm_gbs": peaks.get("hbm_gbs"), "bf16_tflops": peaks.get("bf16_tflops"),
                      "fp32_ffma_tflops_measured_here": fp32_peak, "fp32_sgemm_tflops_measured_here": sgemm_peak},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_our_arm(args)


if __name__ == "__main__":
    main()
