"""Import helper: load the hyphen-named package directory as module `trg_planner_b200`."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent


def load():
    name = "trg_planner_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = ROOT / "trg-planner_b200"
    spec = importlib.util.spec_from_file_location(name, pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_oracle():
    """The CPU oracle's loader (oracle/facade.py) - test infrastructure: tests, smoke(), bench cpu_baseline."""
    name = "trg_oracle_facade"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, ROOT / "oracle" / "facade.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
