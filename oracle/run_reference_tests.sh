#!/bin/bash
# Runs the UNMODIFIED reference's own core/gates tests against oracle/tn_shim (build container
# only: needs /root/reference).  np.complex was removed in numpy>=1.24; the reference uses it at
# mpsim/core.py:507,561, so conftest re-adds the alias.  Output is summarised in DESIGN.md.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
TMP="$(mktemp -d)"
cp -r /root/reference/mpsim "$TMP/mpsim"
# the cirq adapter subpackage needs real cirq; core tests do not
rm -rf "$TMP/mpsim/mpsim_cirq"
cat > "$TMP/conftest.py" <<'PY'
import numpy as np
if not hasattr(np, "complex"):
    np.complex = complex
PY
cd "$TMP"
PYTHONPATH="$HERE/tn_shim:$TMP" python -m pytest -q -x -p no:cacheprovider mpsim/core_test.py mpsim/gates_test.py "$@"
