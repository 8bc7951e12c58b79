"""trg-planner_b200: B200-native TRG construction + risk-aware path queries.

The compute path is C++/CUDA (csrc/, host/) behind the C ABIs in include/; this Python
package only holds the ctypes binding, the YAML parameter sets and the synthetic-map
generators used by tests/ and bench.py. The directory name has a hyphen (fixed by the
project layout), so it is imported under the module name `trg_planner_b200` via `_pkg.py`.
"""
from . import params, terrain  # noqa: F401
from . import binding  # noqa: F401
from .binding import TrgFacade, load_params_yaml, load_pcd, product, save_pcd  # noqa: F401
from .params import INDOOR, MOUNTAIN, TrgParams  # noqa: F401
