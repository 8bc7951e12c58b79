"""ctypes loader of the CPU oracle (oracle/trg_oracle.h, prefix `orc_`).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this. The product package (trg-planner_b200/) holds no
reference to the oracle; the oracle merely re-uses the product's generic facade binding class,
because both libraries export the same C facade of the reference's TRG class.
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ORACLE_LIB = ROOT / "oracle" / "liboracle.so"
ORACLE_REFKD_LIB = ROOT / "oracle" / "_ref" / "liboracle_refkd.so"   # linked against the reference's own kdtree.c


def oracle(params, ref_kdtree: bool = False):
    """One oracle TRG instance; `ref_kdtree` selects the build on top of the verbatim reference kd-tree."""
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    import _pkg
    _pkg.load()
    from trg_planner_b200.binding import TrgFacade
    return TrgFacade(ORACLE_REFKD_LIB if ref_kdtree else ORACLE_LIB, "orc", params)
