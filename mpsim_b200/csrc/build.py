"""Builds libmpsim_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ["api.cu", "cgemm.cu", "theta.cu", "svd_small.cu", "svd_large.cu", "site_ops.cu", "tc_gemm.cu"]
OUT = os.path.join(PKG, "libmpsim_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--shared", "-cudart", "shared",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [os.path.join(HERE, "common.cuh"), os.path.join(HERE, "tc_ptx.cuh"),
            os.path.join(os.path.dirname(PKG), "include", "mpsim_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # MPSB_NVCC_EXTRA / MPSB_LIB_OUT: variant builds for A/B timing (scripts/ab_large.sh), e.g.
    #   MPSB_NVCC_EXTRA=-DTA_STS=1 MPSB_LIB_OUT=mpsim_b200/libmpsim_b200_sts.so python -m mpsim_b200.csrc.build
    extra = os.environ.get("MPSB_NVCC_EXTRA", "").split()
    out = os.environ.get("MPSB_LIB_OUT") or OUT
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(HERE, s) for s in SOURCES] + ["-o", out]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libmpsim_b200.so")
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
