"""Run one of the reference's driver scripts UNMODIFIED on this implementation:

    python -m system_identification_b200.run /path/to/system_identification/spot_identification.py
    python -m system_identification_b200.run /path/to/system_identification/demo/solo_identification.py

The scripts do `from src.solver import Solver` / `from src.sys_identification import SystemIdentification`
(reference demo/solo_identification.py:5-6).  `python script.py` puts the SCRIPT's directory first on sys.path, so for
the copy of spot_identification.py that sits at the root of the reference checkout the reference's own src/ (pinocchio,
cvxpy, MOSEK) would win over this repository's re-exporting src/ whatever PYTHONPATH says.  This launcher executes the
file with runpy instead, with this repository first on the path (equivalent: `PYTHONSAFEPATH=1 PYTHONPATH=<this repo>
python script.py`, or `python -P`).  The script itself is not touched: it still computes its workspace from its own
location (os.path.realpath(__file__), reference demo/solo_identification.py:58-59).
"""
from __future__ import annotations

import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_script(path, run_name="__main__"):
    """Execute `path` with this repository's `src` package taking precedence; returns the script's globals."""
    path = os.path.abspath(path)
    drop = {os.path.dirname(path), ""}
    saved = list(sys.path)
    saved_src = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    try:
        sys.path[:] = [ROOT] + [p for p in sys.path if p not in drop and os.path.abspath(p) != ROOT]
        for k in saved_src:
            del sys.modules[k]
        return runpy.run_path(path, run_name=run_name)
    finally:
        sys.path[:] = saved


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 1:
        print(__doc__)
        return 2
    sys.argv = [argv[0]] + argv[1:]
    run_script(argv[0])
    return 0


if __name__ == "__main__":
    sys.exit(main())
