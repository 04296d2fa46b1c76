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


# measurement aid of bench.py (FFMA peak microkernel): its own small library, never loaded by the package
BENCH_SOURCES = [os.path.join("bench", "ffma_peak.cu")]
BENCH_OUT = os.path.join(PKG, "libmpsb_bench.so")


def build_bench(force: bool = False) -> str:
    srcs = [os.path.join(HERE, s) for s in BENCH_SOURCES]
    if not force and os.path.exists(BENCH_OUT) and all(os.path.getmtime(s) <= os.path.getmtime(BENCH_OUT) for s in srcs):
        return BENCH_OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    res = subprocess.run([nvcc] + NVCC_FLAGS + srcs + ["-o", BENCH_OUT], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed building libmpsb_bench.so")
    return BENCH_OUT


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
    if not os.environ.get("MPSB_LIB_OUT"):
        print(build_bench(force=True))
