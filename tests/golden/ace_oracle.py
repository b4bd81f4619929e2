#!/usr/bin/env python
"""`ACE <parameter file>` for the fixture generator: the reference-format reader/writer of pyaceqd_b200.ace_cli with
the CPU ORACLE as propagation backend (test infrastructure -- lets the UNMODIFIED reference workflows run here)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oracle_backend import OracleEngine  # noqa: E402
from pyaceqd_b200.ace_cli import run_param_file  # noqa: E402

try:
    run_param_file(sys.argv[1], engine=OracleEngine())
except Exception as exc:  # noqa: BLE001
    sys.stderr.write("ACE (oracle): {}: {}\n".format(type(exc).__name__, exc))
    sys.exit(1)
