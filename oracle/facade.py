"""ctypes loader of the CPU oracle (oracle/trg_oracle.h, prefix `orc_`).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this. The product package (trg-planner_b200/) holds no
reference to the oracle; the oracle merely re-uses the product's generic facade binding class,
because all these libraries export the same C facade of the reference's TRG class.

Three builds, strongest first:
  "ref"    oracle/_ref/libtrg_ref.so     the reference's OWN unmodified trg.cpp + kdtree.c compiled
                                         where they lie against oracle/shim/ (ref_harness.cpp)
  "refkd"  oracle/_ref/liboracle_refkd.so the restated trg.cpp (trg_oracle.cpp) on the reference's kdtree.c
  "port"   oracle/liboracle.so           the restatement on the restated kd-tree (richer outputs:
                                         wireEdge stage, float64 twin, counters)
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ORACLE_LIB = ROOT / "oracle" / "liboracle.so"
ORACLE_REFKD_LIB = ROOT / "oracle" / "_ref" / "liboracle_refkd.so"   # linked against the reference's own kdtree.c
REFERENCE_LIB = ROOT / "oracle" / "_ref" / "libtrg_ref.so"           # the reference's own trg.cpp + kdtree.c

LIBS = {"port": ORACLE_LIB, "refkd": ORACLE_REFKD_LIB, "ref": REFERENCE_LIB}


def available(kind: str) -> bool:
    return LIBS[kind].exists()


def _facade_cls():
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    import _pkg
    _pkg.load()
    from trg_planner_b200.binding import TrgFacade
    return TrgFacade


def _pinned_cls():
    import ctypes as C

    import numpy as np
    TrgFacade = _facade_cls()

    class PinnedOracle(TrgFacade):
        """The reference's own trg.cpp (libtrg_ref.so). `wireEdge` does not say where it returned and has
        no float64 twin, so `edge_eval` also asks the restated oracle, REQUIRES it to agree bit for bit
        with the reference on which edges exist and on their (weight, dist), and returns the
        restatement's richer record (stage, npts, weight64). Everything else is the reference alone."""
        kind = "ref"

        def __init__(self, params):
            super().__init__(REFERENCE_LIB, "orc", params)
            self.lib.orc_save_graph.argtypes = [C.c_void_p, C.c_char_p]
            self.lib.orc_load_graph.argtypes = [C.c_void_p, C.c_char_p]
            self._twin = None
            self._maps = {}
            self._fed = set()

        def set_global_map(self, pts):
            super().set_global_map(pts)
            self._maps["global"] = (None, np.array(pts, np.float32, copy=True))
            self._fed.discard("global")

        def set_local_map(self, sx, sy, pts):
            super().set_local_map(sx, sy, pts)
            self._maps["local"] = ((sx, sy), np.array(pts, np.float32, copy=True))
            self._fed.discard("local")

        def _twin_for(self, type_):
            if self._twin is None:
                self._twin = TrgFacade(ORACLE_LIB, "orc", self.params)
            if type_ not in self._fed and type_ in self._maps:
                where, pts = self._maps[type_]
                if where is None:
                    self._twin.set_global_map(pts)
                else:
                    self._twin.set_local_map(where[0], where[1], pts)
                self._fed.add(type_)
            return self._twin

        def edge_eval(self, p1, p2, type_="global"):
            ref = super().edge_eval(p1, p2, type_)
            port = self._twin_for(type_).edge_eval(p1, p2, type_)
            ok = ref["stage"] == 0
            if not (np.array_equal(ok, port["stage"] == 0)
                    and np.array_equal(ref["weight"][ok], port["weight"][ok])
                    and np.array_equal(ref["dist"][ok], port["dist"][ok])):
                raise AssertionError("restated wireEdge disagrees with the reference's own trg.cpp")
            return port

        def save_graph(self, path):
            self.lib.orc_save_graph(self.h, str(path).encode())

        def load_graph(self, path):
            self.lib.orc_load_graph(self.h, str(path).encode())

    return PinnedOracle


def oracle(params, ref_kdtree: bool = False, kind: str | None = None):
    """One oracle TRG instance. `kind` in {"port", "refkd", "ref"}; `ref_kdtree=True` == "refkd".
    Default (no kind): the reference's own code when oracle/_ref/libtrg_ref.so exists, else the port."""
    if kind is None:
        kind = "refkd" if ref_kdtree else ("ref" if available("ref") else "port")
    if kind == "ref":
        return _pinned_cls()(params)
    o = _facade_cls()(LIBS[kind], "orc", params)
    o.kind = kind
    return o


