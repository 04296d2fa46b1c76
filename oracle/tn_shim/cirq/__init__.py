"""Stub of the two cirq symbols ``mpsim/core_test.py`` imports -- TEST INFRASTRUCTURE ONLY.

cirq~=0.8 is not installable here; this stub lets the reference's own core tests run
against ``oracle/tn_shim/tensornetwork``.  Never imported by the product.
"""
from . import qis  # noqa: F401
