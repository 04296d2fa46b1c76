"""ctypes binding of ``libmpsim_b200.so`` (the C-ABI declared in ``include/mpsim_b200.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
first device operation raises ``RuntimeError``.  Host-only logic (gates, planner, circuits)
imports without touching this module's loader.
"""
import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPSIM_B200_LIB: another build of the same library (A/B timing of kernel variants); no fallback either way
LIB_PATH = os.environ.get("MPSIM_B200_LIB") or os.path.join(_HERE, "libmpsim_b200.so")

MAX_SMALL_DIM = 128      # MPSB_MAX_SMALL_DIM of include/mpsim_b200.h

c_void_p, c_int, c_size_t, c_int64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_int64

# numpy mirrors of the descriptor structs (same field order / padding as the C header)
GATE2_DESC = np.dtype([
    ("site_l", np.uint64), ("site_r", np.uint64), ("out_l", np.uint64), ("out_r", np.uint64),
    ("gate", np.uint64), ("svals", np.uint64),
    ("bs_site_l", np.int64), ("bs_site_r", np.int64), ("bs_out_l", np.int64), ("bs_out_r", np.int64),
    ("bs_gate", np.int64), ("bs_svals", np.int64),
], align=True)
GATE1_DESC = np.dtype([
    ("site", np.uint64), ("out", np.uint64), ("gate", np.uint64),
    ("bs_site", np.int64), ("bs_out", np.int64), ("bs_gate", np.int64),
    ("chiL", np.int32), ("chiR", np.int32),
], align=True)
GATE2_GROUP = np.dtype([("descs", np.uint64), ("info", np.uint64), ("ndesc", np.int32), ("chiL", np.int32),
                        ("chiM", np.int32), ("chiR", np.int32), ("k", np.int32), ("left_canonical", np.int32)], align=True)
SITE_REF = np.dtype([("site", np.uint64), ("bs", np.int64), ("chiL", np.int32), ("chiR", np.int32)], align=True)
assert GATE2_DESC.itemsize == 96 and GATE1_DESC.itemsize == 56 and SITE_REF.itemsize == 24
assert GATE2_GROUP.itemsize == 40

#: every symbol include/mpsim_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mpsb_version": (c_int, []),
    "mpsb_last_error": (ctypes.c_char_p, []),
    "mpsb_device_info": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsb_gate2_workspace_bytes": (c_size_t, [c_int] * 7),
    "mpsb_apply_gate2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_size_t, c_void_p, c_void_p]),
    "mpsb_gate2_layer_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "mpsb_apply_gate2_layer": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mpsb_apply_gate1": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mpsb_inner_workspace_bytes": (c_size_t, [c_int] * 4),
    "mpsb_inner_products": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mpsb_scale_sites": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mpsb_rebalance_sites": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mpsb_wavefunction_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "mpsb_wavefunction": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mpsb_amplitudes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "mpsb_cgemm": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int64, c_void_p, c_int64, c_int64, c_int, c_int64,
                           c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "mpsb_cgemm_tc_workspace_bytes": (c_size_t, [c_int] * 4),
    "mpsb_cgemm_tc": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int,
                              c_void_p, c_size_t, c_void_p]),
    "mpsb_theta_workspace_bytes": (c_size_t, [c_int] * 6),
    "mpsb_theta": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mpsb_svd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mpsb_svd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_size_t, c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


def load(require_device: bool = False) -> ctypes.CDLL:
    """Load the shared library (once).  Raises RuntimeError -- never falls back -- when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m mpsim_b200.csrc.build` "
                "(or __graft_entry__.build()).  mpsim_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if require_device:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("mpsim_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
    return _lib


def check(rc: int, what: str = "") -> None:
    """Error convention of the C-ABI: <0 argument error -> ValueError, >0 CUDA error -> RuntimeError."""
    if rc == 0:
        return
    msg = load().mpsb_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg}" if what else msg)
    raise RuntimeError(f"{what}: CUDA error {rc}: {msg}" if what else f"CUDA error {rc}: {msg}")


def to_device_bytes(arr: np.ndarray, device):
    """Upload a (structured) numpy array as a uint8 device tensor; keeps it alive for the caller."""
    import torch
    host = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1))
    return host.to(device, non_blocking=False)


def stream_ptr(device=None) -> int:
    """Raw handle of torch's current stream on ``device`` (default: the current device)."""
    import torch
    if device is None:
        index = torch.cuda.current_device()
    else:
        index = torch.device(device).index
        if index is None:
            index = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(index)
