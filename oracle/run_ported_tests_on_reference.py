"""TEST INFRASTRUCTURE (build container only: needs /root/reference).

Runs tests/test_gpu_reference_ported.py and tests/test_gates_ported.py -- the reference's own tests
ported to this package's API -- against the UNMODIFIED reference package on oracle/tn_shim instead of mpsim_b200, on the CPU.  A port that
does not hold on the reference itself is a wrong port; this is how the ports are checked before they
are trusted as parity tests on the GPU.  (mpsim_cirq is left out: it needs the real Cirq.)

    python oracle/run_ported_tests_on_reference.py
"""
import inspect
import os
import shutil
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/mpsim"


def main() -> int:
    if not os.path.isdir(REF):
        print("reference not present: nothing to check")
        return 0
    import numpy as np
    if not hasattr(np, "complex"):
        np.complex = complex                  # mpsim/core.py:507,561 predate numpy 1.24
    tmp = tempfile.mkdtemp()
    try:
        pkg = os.path.join(tmp, "mpsim")
        shutil.copytree(REF, pkg, ignore=shutil.ignore_patterns("mpsim_cirq", "*_test.py"))
        init = os.path.join(pkg, "__init__.py")
        with open(init) as f:
            lines = [ln for ln in f if "mpsim_cirq" not in ln]
        with open(init, "w") as f:
            f.writelines(lines)
        sys.path[:0] = [os.path.join(HERE, "tn_shim"), tmp, ROOT]
        warnings.simplefilter("ignore")
        import mpsim
        import mpsim.gates
        import tests.test_gpu_reference_ported as ported
        import tests.test_gates_ported as gates_ported
        ported._mp = lambda: mpsim
        gates_ported._g = lambda: mpsim.gates
        ok = bad = 0
        cases = [(n, f) for mod in (ported, gates_ported) for n, f in inspect.getmembers(mod, inspect.isfunction)]
        for name, fn in cases:
            if not name.startswith("test_"):
                continue
            combos = [{}]
            for mark in getattr(fn, "pytestmark", []):
                if mark.name == "parametrize":
                    combos = [dict(c, **{mark.args[0]: v}) for c in combos for v in mark.args[1]]
            for kw in combos:
                try:
                    fn(**kw)
                    ok += 1
                except Exception as e:        # noqa: BLE001 -- report every failing port
                    bad += 1
                    print(f"FAILS ON THE REFERENCE: {name} {kw}: {type(e).__name__}: {e}")
        print(f"{ok} ported test cases hold on the unmodified reference, {bad} do not")
        return 1 if bad else 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
