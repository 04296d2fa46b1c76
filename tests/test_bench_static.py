"""bench.py cannot run here (no GPU): check statically that every function it calls by name exists (a
refactoring once dropped two measurement functions unnoticed), and run its CPU arm on a tiny workload."""
import ast
import builtins
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_calls_only_defined_names():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    known = {n.name for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
    for n in ast.walk(tree):
        if isinstance(n, ast.ImportFrom):
            known |= {a.asname or a.name for a in n.names}
        elif isinstance(n, ast.Import):
            known |= {(a.asname or a.name).split(".")[0] for a in n.names}
        elif isinstance(n, (ast.arg,)):
            known.add(n.arg)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
            known.add(n.id)
    called = {n.func.id for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name)}
    missing = sorted(c for c in called if c not in known and not hasattr(builtins, c))
    assert not missing, missing


def test_reference_arm_runs_on_a_tiny_workload():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--nqubits", "10", "--depth", "4", "--chi", "8"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["warmup"] >= 1
